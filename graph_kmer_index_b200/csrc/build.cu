// build.cu -- K2: CollisionFreeKmerIndex construction on the device.
//
// Reference: collision_free_kmer_index.py:422-467 (from_flat_kmers) and :267-293 (set_frequencies):
//   hashes = kmers % modulo; sorting = argsort(hashes); gather 5 columns; run heads -> hashes_to_index,
//   run lengths -> n_kmers; frequencies = #distinct ref_offsets per k-mer.
// Two paths, same result (the canonical order of SURVEY.md section 8c(ii): inside a bucket the entries keep their input
// order; numpy's default argsort is not stable, so the reference's own payload order is only defined up to a permutation
// inside each bucket):
//   * binned (the usual case, see bin_finish_small_kernel): bins of consecutive buckets -> histogram + scan -> ONE scatter of
//     whole 32-byte records, each with one 256-bit store -> per-bin ordering by (bucket, input index) in registers -> payload
//     columns and both dense tables streamed out.  24 (in) + 32 (records out) + 32 (records in) + 24 (out) bytes per entry
//     + 8 * modulo for the tables.
//   * radix (fallback for skewed inputs, and the machinery behind gki_group_by_key / gki_partition_by_bucket_range):
//     (bucket key << 32 | index) elements -> stable LSD radix sort, up to 10 bits per pass (3 passes at the default modulo)
//     -> run heads / tails into the zeroed dense tables -> one gather of interleaved payload records through the permutation.
// Frequencies (cfki:267-293) are bucket-local scans over the sorted columns on both paths.
//
// Roofline: HBM streaming, compulsory traffic 50*N + 8*modulo bytes.
#include <algorithm>
#include <stdlib.h>
#include "common.cuh"

namespace gki {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RADIX_MAX = 1024;   // digits of up to 10 bits: a 29-bit bucket id (default modulo) sorts in 3 passes

// sort element = (bucket key << 32) | input index: one 8-byte item per entry instead of two 4-byte arrays halves
// the number of scattered stores of a pass
__device__ __forceinline__ uint32_t elem_key(unsigned long long e) { return (uint32_t)(e >> 32); }
__device__ __forceinline__ uint32_t elem_idx(unsigned long long e) { return (uint32_t)e; }

// key = bucket - bucket_lo (whole-index build: bucket_lo = 0; hash-range partitioned build: the rank's first bucket),
// or, with part_size != 0, key = bucket / part_size (the owner of the bucket range, for the all-to-all)
__global__ void bucket_keys_kernel(const uint64_t *__restrict__ kmers, int64_t n, FastMod fm, uint32_t bucket_lo, uint32_t part_size,
                                   unsigned long long *__restrict__ elems) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t b = fastmod(__ldg(kmers + i), fm);
        uint32_t key = part_size ? b / part_size : b - bucket_lo;
        elems[i] = ((unsigned long long)key << 32) | (unsigned long long)(uint32_t)i;
    }
}

// first sorted position whose key is >= p, for p = 0..n_parts (bounds[p]); counts follow by difference
__global__ void part_bounds_kernel(const unsigned long long *__restrict__ sorted, int64_t n, int n_parts, long long *__restrict__ bounds) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n_parts) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (elem_key(__ldg(sorted + mid)) < (uint32_t)p) lo = mid + 1; else hi = mid;
    }
    bounds[p] = lo;
}
__global__ void part_counts_kernel(const long long *__restrict__ bounds, int n_parts, long long *__restrict__ counts) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_parts) counts[p] = bounds[p + 1] - bounds[p];
}
__global__ void extract_perm_kernel(const unsigned long long *__restrict__ sorted, int64_t n, uint32_t *__restrict__ perm) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        perm[i] = elem_idx(__ldg(sorted + i));
}
// hashes_to_index holds local positions after a range build; the global position adds the entries of the lower ranks
__global__ void add_offset_nonempty_kernel(int32_t *__restrict__ h2i, const uint32_t *__restrict__ nk, int64_t len, int32_t offset) {
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < len; b += (int64_t)gridDim.x * blockDim.x)
        if (nk[b]) h2i[b] += offset;
}

// per-tile digit histogram, stored digit-major: hist[d * n_tiles + tile]
__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const unsigned long long *__restrict__ elems, int64_t n, int shift,
                                                                int radix, uint32_t *__restrict__ hist, int64_t n_tiles) {
    __shared__ uint32_t h[RADIX_MAX];
    for (int d = threadIdx.x; d < radix; d += RS_THREADS) h[d] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    const uint32_t dmask = (uint32_t)radix - 1u;
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {
        int64_t i = base + it * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(elem_key(__ldg(elems + i)) >> shift) & dmask], 1u);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < radix; d += RS_THREADS) hist[(int64_t)d * n_tiles + blockIdx.x] = h[d];
}

// lanes of `okmask` holding the same BITS-bit digit.  MATCH.ANY issues about once per 124 cycles per SM on B200
// (ncu: profiles/r1/radix_scatter_match_any.txt), BITS ballots + logic ops are several times faster.
template <int BITS> __device__ __forceinline__ uint32_t digit_peers(uint32_t d, uint32_t okmask) {
    uint32_t peers = okmask;
#pragma unroll
    for (int b = 0; b < BITS; b++) {
        const bool bit = (d >> b) & 1u;
        const uint32_t votes = __ballot_sync(okmask, bit);
        peers &= bit ? votes : ~votes;
    }
    return peers;
}

// stable scatter of one tile.  Element order inside a tile: warp-major, then item, then lane -- i.e. memory order.
template <int BITS>
__global__ void __launch_bounds__(RS_THREADS)
    radix_scatter_kernel(const unsigned long long *__restrict__ in, int64_t n, int shift, const uint32_t *__restrict__ offsets,
                         int64_t n_tiles, unsigned long long *__restrict__ out) {
    constexpr int radix = 1 << BITS;
    __shared__ uint32_t warp_hist[RS_WARPS][radix];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int w = 0; w < RS_WARPS; w++)
        for (int d = threadIdx.x; d < radix; d += RS_THREADS) warp_hist[w][d] = 0;
    __syncthreads();
    const int64_t warp_base = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (RS_ITEMS * 32);
    const uint32_t dmask = (uint32_t)radix - 1u;
    unsigned long long e[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {   // all loads first: RS_ITEMS independent requests in flight
        int64_t i = warp_base + it * 32 + lane;
        e[it] = i < n ? __ldg(in + i) : 0ull;
    }
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {
        int64_t i = warp_base + it * 32 + lane;
        bool ok = i < n;
        uint32_t okmask = __ballot_sync(0xffffffffu, ok);
        rank[it] = 0;
        if (ok) {
            uint32_t d = (elem_key(e[it]) >> shift) & dmask;
            uint32_t peers = digit_peers<BITS>(d, okmask);
            uint32_t before = __popc(peers & ((1u << lane) - 1u));
            int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (lane == leader) {
                base = warp_hist[warp][d];
                warp_hist[warp][d] = base + __popc(peers);
            }
            base = __shfl_sync(peers, base, leader);
            rank[it] = base + before;
        }
        __syncwarp();
    }
    __syncthreads();
    for (int d = threadIdx.x; d < radix; d += RS_THREADS) {   // per-warp counts -> start positions in the output
        uint32_t run = offsets[(int64_t)d * n_tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c = warp_hist[w][d];
            warp_hist[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {
        int64_t i = warp_base + it * 32 + lane;
        if (i < n) out[warp_hist[warp][(elem_key(e[it]) >> shift) & dmask] + rank[it]] = e[it];
    }
}

// Stable sort of the packed elements by the low `bits` bits of their key.  On return *sorted points into the
// scratch buffers owned by `bufs`.
struct SortBuffers {
    Scratch elems[2], hist;
};

static int radix_sort_elems(SortBuffers &bufs, int64_t n, int bits, const unsigned long long **sorted, cudaStream_t s) {
    // bufs.elems[0] holds the input
    const int64_t n_tiles = (n + RS_TILE - 1) / RS_TILE;
    if (bits < 1) bits = 1;
    const int passes = (bits + 9) / 10;
    int digit_bits = (bits + passes - 1) / passes;   // <= 10
    if (digit_bits < 8) digit_bits = 8;
    const int radix = 1 << digit_bits;
    GKI_TRY(bufs.elems[1].alloc((size_t)n * 8, s));
    GKI_TRY(bufs.hist.alloc((size_t)radix * n_tiles * 4, s));
    int cur = 0;
    for (int p = 0; p < passes; p++) {
        const int shift = digit_bits * p;
        radix_hist_kernel<<<(unsigned)n_tiles, RS_THREADS, 0, s>>>(bufs.elems[cur].as<unsigned long long>(), n, shift, radix,
                                                                  bufs.hist.as<uint32_t>(), n_tiles);
        GKI_CHECK_LAUNCH();
        GKI_TRY(exclusive_scan_u32(bufs.hist.as<uint32_t>(), bufs.hist.as<uint32_t>(), (int64_t)radix * n_tiles, nullptr, s));
        const unsigned long long *src = bufs.elems[cur].as<unsigned long long>();
        unsigned long long *dst = bufs.elems[cur ^ 1].as<unsigned long long>();
        if (digit_bits == 8) radix_scatter_kernel<8><<<(unsigned)n_tiles, RS_THREADS, 0, s>>>(src, n, shift, bufs.hist.as<uint32_t>(), n_tiles, dst);
        else if (digit_bits == 9) radix_scatter_kernel<9><<<(unsigned)n_tiles, RS_THREADS, 0, s>>>(src, n, shift, bufs.hist.as<uint32_t>(), n_tiles, dst);
        else radix_scatter_kernel<10><<<(unsigned)n_tiles, RS_THREADS, 0, s>>>(src, n, shift, bufs.hist.as<uint32_t>(), n_tiles, dst);
        GKI_CHECK_LAUNCH();
        cur ^= 1;
    }
    *sorted = bufs.elems[cur].as<unsigned long long>();
    return GKI_OK;
}

// run heads: hashes_to_index[bucket] = first sorted position (cfki:444-454).  The head also walks its run (runs are a
// handful of entries unless modulo is tiny) and writes n_kmers[bucket]; runs longer than RUN_WALK are left to
// run_tails_kernel, which is only launched when *long_runs != 0.
constexpr int RUN_WALK = 32;
__global__ void run_heads_kernel(const unsigned long long *__restrict__ sorted, int64_t n, int32_t *__restrict__ h2i,
                                 uint32_t *__restrict__ nk, unsigned int *__restrict__ long_runs) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t k = elem_key(__ldg(sorted + i));
        if (i == 0 || elem_key(__ldg(sorted + i - 1)) != k) {
            h2i[k] = (int32_t)i;
            int len = 1;
            while (len <= RUN_WALK && i + len < n && elem_key(__ldg(sorted + i + len)) == k) len++;
            if (len <= RUN_WALK) nk[k] = (uint32_t)len;
            else *long_runs = 1u;
        }
    }
}
// run tails: n_kmers[bucket] = run length (cfki:455-457)
__global__ void run_tails_kernel(const unsigned long long *__restrict__ sorted, int64_t n, const int32_t *__restrict__ h2i,
                                 uint32_t *__restrict__ nk, const unsigned int *__restrict__ long_runs) {
    if (*long_runs == 0u) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t k = elem_key(__ldg(sorted + i));
        if (i == n - 1 || elem_key(__ldg(sorted + i + 1)) != k) nk[k] = (uint32_t)(i + 1 - h2i[k]);
    }
}

// cfki:436-440: the payload columns follow the sort.  Gathering four columns through a random permutation costs four
// HBM fetches (64-128 B each) per entry; interleaving them first into one record per entry (streaming pass) makes it
// ONE fetch per entry.  Record = {kmer, node, af} (16 B) or {kmer, ref, node, af, pad} (32 B) when ref_offsets move too.
template <bool WITH_REF>
__global__ void pack_records_kernel(int64_t n, const uint64_t *__restrict__ kmers, const uint32_t *__restrict__ nodes,
                                    const uint64_t *__restrict__ ref, const float *__restrict__ af, uint4 *__restrict__ rec) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t km = __ldg(kmers + i);
        uint32_t nd = nodes ? __ldg(nodes + i) : 0u;
        uint32_t a = af ? __float_as_uint(__ldg(af + i)) : 0u;
        if (WITH_REF) {
            uint64_t r = __ldg(ref + i);
            rec[2 * i] = make_uint4((uint32_t)km, (uint32_t)(km >> 32), (uint32_t)r, (uint32_t)(r >> 32));
            rec[2 * i + 1] = make_uint4(nd, a, 0u, 0u);
        } else {
            rec[i] = make_uint4((uint32_t)km, (uint32_t)(km >> 32), nd, a);
        }
    }
}

template <bool WITH_REF>
__global__ void gather_records_kernel(const unsigned long long *__restrict__ sorted, int64_t n, const uint4 *__restrict__ rec,
                                      uint64_t *__restrict__ kmers_o, uint32_t *__restrict__ nodes_o, uint64_t *__restrict__ ref_o,
                                      float *__restrict__ af_o, uint32_t *__restrict__ perm_o) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t p = elem_idx(__ldg(sorted + i));
        uint4 a = __ldg(rec + (WITH_REF ? 2 * (int64_t)p : (int64_t)p));
        uint32_t nd, afb;
        if (WITH_REF) {
            uint4 b = __ldg(rec + 2 * (int64_t)p + 1);
            if (ref_o) ref_o[i] = ((uint64_t)a.w << 32) | a.z;
            nd = b.x;
            afb = b.y;
        } else {
            nd = a.z;
            afb = a.w;
        }
        if (kmers_o) kmers_o[i] = ((uint64_t)a.y << 32) | a.x;
        if (nodes_o) nodes_o[i] = nd;
        if (af_o) af_o[i] = __uint_as_float(afb);
        if (perm_o) perm_o[i] = p;
    }
}

__global__ void gather_payload_kernel(const unsigned long long *__restrict__ sorted, int64_t n, const uint64_t *__restrict__ kmers,
                                      const uint32_t *__restrict__ nodes, const uint64_t *__restrict__ ref,
                                      const float *__restrict__ af, uint64_t *__restrict__ kmers_o, uint32_t *__restrict__ nodes_o,
                                      uint64_t *__restrict__ ref_o, float *__restrict__ af_o, uint32_t *__restrict__ perm_o) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t p = elem_idx(__ldg(sorted + i));
        if (kmers_o) kmers_o[i] = __ldg(kmers + p);
        if (nodes_o) nodes_o[i] = __ldg(nodes + p);
        if (ref_o) ref_o[i] = __ldg(ref + p);
        if (af_o) af_o[i] = __ldg(af + p);
        if (perm_o) perm_o[i] = p;
    }
}

template <typename T> __global__ void gather_kernel(const T *__restrict__ src, const uint32_t *__restrict__ perm, int64_t n, T *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __ldg(src + __ldg(perm + i));
}

// ---- binned build: one scatter pass, then every bin is ordered on its own --------------------------------------------------
// Bucket ids of an index are hashes, close to uniform over [0, modulo): a bin of 2^shift consecutive buckets receives a
// predictable handful of entries.  So instead of three radix passes plus a random gather: (1) histogram of the bins (RED on an
// L2-resident counter array), exclusive scan; (2) every entry is written once, as a whole record, to the next free slot of its
// bin (returning atomic on the bin's cursor; the measured ceiling for random 32-byte stores is ~20 G/s,
// profiles/r1/calibrate_scatter.jsonl); (3) one warp per bin brings its records into (bucket, input index) order in shared
// memory -- the canonical stable order, whatever order the atomics produced -- and streams out the payload columns and BOTH
// dense tables, empty buckets included (no memset, no random table scatter).  Bins that overflow the per-warp capacity (tiny
// modulo, one k-mer repeated hundreds of times) send the whole build to the radix path.
constexpr int BIN_CAP = 512;          // records per bin the finish kernel can order
constexpr int BIN_MAX_SHIFT = 8;      // at most 256 buckets per bin
constexpr int BIN_WARPS = 4;
struct BinParams {
    FastMod fm;
    uint32_t bucket_lo, shift, n_bins;
    uint64_t table_len;
};
__device__ __forceinline__ uint32_t bin_of(uint64_t kmer, const BinParams &p) { return (fastmod(kmer, p.fm) - p.bucket_lo) >> p.shift; }

__global__ void bin_hist_kernel(const uint64_t *__restrict__ kmers, int64_t n, BinParams p, uint32_t *__restrict__ counts) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(counts + bin_of(__ldg(kmers + i), p), 1u);
}
__global__ void bin_max_kernel(const uint32_t *__restrict__ counts, uint32_t n_bins, uint32_t *__restrict__ max_out) {
    uint32_t m = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_bins; i += gridDim.x * blockDim.x) m = max(m, counts[i]);
#pragma unroll
    for (int d = 16; d; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(max_out, m);
}
// record = {kmer, ref_offset, node | af << 32, input index} = 32 bytes = one sector, written with ONE 256-bit store: random
// full-sector stores issued as one instruction run at 42 G/s on this part, twice the rate of the same bytes issued as two
// 128-bit stores or of 16-byte records (profiles/r1/calibrate_scatter.jsonl) -- so the record is 32 bytes even when only
// k-mers and nodes are wanted.  The bin cursors start at the bins' first slots, so the returning atomic is the position.
struct BinRecord {
    unsigned long long kmer, ref, node_af, index;
};
__device__ __forceinline__ BinRecord load_record(const BinRecord *p) {
    BinRecord r;
    asm("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(r.kmer), "=l"(r.ref), "=l"(r.node_af), "=l"(r.index) : "l"(p));
    return r;
}
__global__ void bin_scatter_kernel(int64_t n, const uint64_t *__restrict__ kmers, const uint32_t *__restrict__ nodes,
                                   const uint64_t *__restrict__ ref, const float *__restrict__ af, BinParams p,
                                   uint32_t *__restrict__ cursor, BinRecord *__restrict__ rec) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t km = __ldg(kmers + i);
        const size_t pos = atomicAdd(cursor + bin_of(km, p), 1u);
        const unsigned long long nd = nodes ? __ldg(nodes + i) : 0u;
        const unsigned long long a = af ? __float_as_uint(__ldg(af + i)) : 0u;
        asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(rec + pos), "l"((unsigned long long)km),
                     "l"((unsigned long long)(ref ? __ldg(ref + i) : 0ull)), "l"(nd | (a << 32)), "l"((unsigned long long)i)
                     : "memory");
    }
}
__global__ void __launch_bounds__(BIN_WARPS * 32)
bin_finish_kernel(BinParams p, const uint32_t *__restrict__ bin_start, const BinRecord *__restrict__ rec, int32_t *__restrict__ h2i,
                  uint32_t *__restrict__ nk, uint64_t *__restrict__ kmers_o, uint32_t *__restrict__ nodes_o, uint64_t *__restrict__ ref_o,
                  float *__restrict__ af_o, uint32_t *__restrict__ perm_o) {
    __shared__ uint32_t s_cnt[BIN_WARPS][1 << BIN_MAX_SHIFT], s_excl[BIN_WARPS][1 << BIN_MAX_SHIFT], s_idx[BIN_WARPS][BIN_CAP];
    __shared__ uint16_t s_bucket[BIN_WARPS][BIN_CAP], s_order[BIN_WARPS][BIN_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *cnt = s_cnt[warp], *excl = s_excl[warp], *idx = s_idx[warp];
    uint16_t *bucket = s_bucket[warp], *order = s_order[warp];
    const uint32_t B = 1u << p.shift;
    for (uint32_t bin = blockIdx.x * BIN_WARPS + warp; bin < p.n_bins; bin += gridDim.x * BIN_WARPS) {
        const uint32_t start = __ldg(bin_start + bin), c = __ldg(bin_start + bin + 1) - start;
        const uint64_t first_bucket = (uint64_t)bin << p.shift;
        for (uint32_t b = lane; b < B; b += 32) cnt[b] = 0;
        __syncwarp();
        for (uint32_t j = lane; j < c; j += 32) {
            const BinRecord a = load_record(rec + start + j);
            const uint32_t bl = (uint32_t)((fastmod(a.kmer, p.fm) - p.bucket_lo) - first_bucket);
            bucket[j] = (uint16_t)bl;
            idx[j] = (uint32_t)a.index;
            atomicAdd(cnt + bl, 1u);
        }
        __syncwarp();
        // exclusive scan of the bucket counts: every lane owns B/32 consecutive buckets (all of them in lane 0.. when B < 32)
        const uint32_t per = (B + 31) / 32;
        uint32_t mine = 0;
        for (uint32_t q = 0; q < per; q++) {
            const uint32_t b = lane * per + q;
            if (b < B) mine += cnt[b];
        }
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        uint32_t run = incl - mine;
        for (uint32_t q = 0; q < per; q++) {
            const uint32_t b = lane * per + q;
            if (b < B) {
                excl[b] = run;
                run += cnt[b];
            }
        }
        __syncwarp();
        // both tables, every bucket of the bin (cfki:444-457): position of the run head, run length; 0 / 0 when empty
        for (uint32_t b = lane; b < B; b += 32) {
            const uint64_t g = first_bucket + b;
            if (g < p.table_len) {
                const uint32_t len = cnt[b];
                nk[g] = len;
                h2i[g] = len ? (int32_t)(start + excl[b]) : 0;
            }
        }
        // group the records by bucket (any order), then order every bucket's run by input index: the stable order
        __syncwarp();
        for (uint32_t j = lane; j < c; j += 32) order[atomicAdd(excl + bucket[j], 1u)] = (uint16_t)j;   // excl[b] becomes the run's end
        __syncwarp();
        for (uint32_t b = lane; b < B; b += 32) {
            const uint32_t len = cnt[b], lo = excl[b] - len;
            for (uint32_t a = 1; a < len; a++) {   // insertion sort: runs are a handful of entries
                const uint16_t moving = order[lo + a];
                const uint32_t key = idx[moving];
                uint32_t q = a;
                while (q > 0 && idx[order[lo + q - 1]] > key) {
                    order[lo + q] = order[lo + q - 1];
                    q--;
                }
                order[lo + q] = moving;
            }
        }
        __syncwarp();
        for (uint32_t q = lane; q < c; q += 32) {
            const size_t dst = (size_t)start + q;
            const BinRecord a = load_record(rec + start + order[q]);
            if (kmers_o) kmers_o[dst] = a.kmer;
            if (ref_o) ref_o[dst] = a.ref;
            if (nodes_o) nodes_o[dst] = (uint32_t)a.node_af;
            if (af_o) af_o[dst] = __uint_as_float((uint32_t)(a.node_af >> 32));
            if (perm_o) perm_o[dst] = (uint32_t)a.index;
        }
        __syncwarp();
    }
}

// The common case -- at most 64 records in every bin, usually under 32 -- needs no per-bucket arrays: with one record per lane
// the rank of a record in (bucket, input index) order comes from eight ballots over the bits of its bucket id (lanes with a
// smaller bucket, lanes with the same one) plus a walk over the one or two lanes that share its bucket.  Records stay in
// registers from the single load to the final stores; both tables are zero-filled for the bin's buckets first and the run heads
// then overwrite their entries.  Bins with 33..64 records take two records per lane and count ranks against shared memory.
constexpr int BIN_SMALL_CAP = 64;
__global__ void __launch_bounds__(BIN_WARPS * 32)
bin_finish_small_kernel(BinParams p, const uint32_t *__restrict__ bin_start, const BinRecord *__restrict__ rec, int32_t *__restrict__ h2i,
                        uint32_t *__restrict__ nk, uint64_t *__restrict__ kmers_o, uint32_t *__restrict__ nodes_o,
                        uint64_t *__restrict__ ref_o, float *__restrict__ af_o, uint32_t *__restrict__ perm_o) {
    __shared__ unsigned long long s_comp[BIN_WARPS][BIN_SMALL_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long *comp = s_comp[warp];
    const uint32_t B = 1u << p.shift;
    auto emit = [&](size_t dst, const BinRecord &a) {
        if (kmers_o) kmers_o[dst] = a.kmer;
        if (ref_o) ref_o[dst] = a.ref;
        if (nodes_o) nodes_o[dst] = (uint32_t)a.node_af;
        if (af_o) af_o[dst] = __uint_as_float((uint32_t)(a.node_af >> 32));
        if (perm_o) perm_o[dst] = (uint32_t)a.index;
    };
    for (uint32_t bin = blockIdx.x * BIN_WARPS + warp; bin < p.n_bins; bin += gridDim.x * BIN_WARPS) {
        const uint32_t start = __ldg(bin_start + bin), c = __ldg(bin_start + bin + 1) - start;
        const uint64_t first_bucket = (uint64_t)bin << p.shift;
        // zero both tables over the bin's buckets (cfki:451-452: np.zeros)
        if (B >= 4 && first_bucket + B <= p.table_len && (((uintptr_t)h2i | (uintptr_t)nk) & 15) == 0) {   // caller's tables may be views
            for (uint32_t q = lane; q < B / 4; q += 32) {
                ((uint4 *)(h2i + first_bucket))[q] = make_uint4(0u, 0u, 0u, 0u);
                ((uint4 *)(nk + first_bucket))[q] = make_uint4(0u, 0u, 0u, 0u);
            }
        } else {
            for (uint32_t b = lane; b < B; b += 32)
                if (first_bucket + b < p.table_len) {
                    h2i[first_bucket + b] = 0;
                    nk[first_bucket + b] = 0u;
                }
        }
        BinRecord a0{0, 0, 0, 0}, a1{0, 0, 0, 0};
        uint32_t bl0 = 0xffffu, bl1 = 0xffffu, id0 = 0, id1 = 0;
        const bool have0 = (uint32_t)lane < c, have1 = (uint32_t)lane + 32u < c;
        if (have0) {
            a0 = load_record(rec + start + lane);
            bl0 = (uint32_t)((fastmod(a0.kmer, p.fm) - p.bucket_lo) - first_bucket);
            id0 = (uint32_t)a0.index;
        }
        if (have1) {
            a1 = load_record(rec + start + 32 + lane);
            bl1 = (uint32_t)((fastmod(a1.kmer, p.fm) - p.bucket_lo) - first_bucket);
            id1 = (uint32_t)a1.index;
        }
        __syncwarp();   // the zero fill is ordered before the run heads' stores below
        if (c <= 32) {
            const uint32_t active = __ballot_sync(0xffffffffu, have0);
            uint32_t eq = active, lt = 0;
#pragma unroll
            for (int bit = BIN_MAX_SHIFT - 1; bit >= 0; bit--) {   // MSB first: lanes with a smaller bucket id, lanes with mine
                const uint32_t ones = __ballot_sync(0xffffffffu, (bl0 >> bit) & 1u);
                if ((bl0 >> bit) & 1u) {
                    lt |= eq & ~ones;
                    eq &= ones;
                } else {
                    eq &= ~ones;
                }
            }
            uint32_t lower_same = 0;   // lanes of my bucket whose record came earlier in the input
            // every lane walks the lanes of its own bucket (one or two of them, itself included)
            uint32_t walk = have0 ? eq : 0u;
            while (__any_sync(0xffffffffu, walk != 0u)) {
                const int src = walk ? __ffs(walk) - 1 : 0;
                const uint32_t other = __shfl_sync(0xffffffffu, id0, src);
                if (walk) {
                    lower_same += other < id0;
                    walk &= walk - 1u;
                }
            }
            if (have0) {
                const uint32_t rank = __popc(lt & active) + lower_same;
                if (lower_same == 0) {
                    nk[first_bucket + bl0] = __popc(eq);
                    h2i[first_bucket + bl0] = (int32_t)(start + rank);
                }
                emit((size_t)start + rank, a0);
            }
        } else {
            comp[lane] = have0 ? (((unsigned long long)bl0 << 32) | id0) : ~0ull;
            comp[32 + lane] = have1 ? (((unsigned long long)bl1 << 32) | id1) : ~0ull;
            __syncwarp();
            const unsigned long long c0 = comp[lane], c1 = comp[32 + lane];
            uint32_t r0 = 0, r1 = 0, len0 = 0, len1 = 0, low0 = 0, low1 = 0;
            for (uint32_t i = 0; i < c; i++) {
                const unsigned long long ci = comp[i];
                const uint32_t bi = (uint32_t)(ci >> 32);
                r0 += ci < c0;
                r1 += ci < c1;
                len0 += bi == bl0;
                len1 += bi == bl1;
                low0 += (bi == bl0) & (ci < c0);
                low1 += (bi == bl1) & (ci < c1);
            }
            if (have0) {
                if (low0 == 0) {
                    nk[first_bucket + bl0] = len0;
                    h2i[first_bucket + bl0] = (int32_t)(start + r0);
                }
                emit((size_t)start + r0, a0);
            }
            if (have1) {
                if (low1 == 0) {
                    nk[first_bucket + bl1] = len1;
                    h2i[first_bucket + bl1] = (int32_t)(start + r1);
                }
                emit((size_t)start + r1, a1);
            }
            __syncwarp();
        }
    }
}

// ---- slab build: one append-scatter into L2-friendly slabs, then every slab is ordered in shared memory -------------------
// The 16-entry bins above make the scatter a random 32-byte store per entry (8 M open bins: every store is its own partial line
// in L2 and in DRAM) and leave ~300 warp instructions of ordering per bin.  Here a bin is a range of `nb` consecutive buckets
// expected to hold ~3/4 of SLAB_CAP entries, and owns a fixed-capacity slab of SLAB_CAP records:
//   (1) scatter: bucket -> bin (multiply-high division), returning atomic on the bin's counter = slot in its slab, ONE 256-bit
//       store.  There is no histogram pass (the capacity is fixed), and with tens of thousands of bins instead of millions the
//       line every counter currently appends to stays in L2 until its four records have arrived: DRAM sees whole lines.
//   (2) exclusive scan of the bin counters -> output position of every bin.
//   (3) finish: a CTA takes a bin: ONE bulk copy (cp.async.bulk, mbarrier) brings its slab into shared memory; per-bucket counts
//       by shared-memory atomics on packed 16-bit counters (the value returned is the record's arrival rank in its bucket);
//       in-place exclusive scan; records of a bucket shared by several entries are ranked by input index (the stable order,
//       whatever order the atomics produced); then the payload columns leave in output order with coalesced stores, and BOTH
//       dense tables are written for every bucket of the bin, empty ones included (no memset, no table scatter).
// A bin that receives more than SLAB_CAP records (tiny modulo, one k-mer repeated thousands of times) raises a flag and the build
// falls back to the 16-entry-bin path above, then to the radix path.
// What bounds the scatter (profiles/r2/calibrate_append.jsonl, calibrate_store_groups.jsonl, slab_scatter_l2_hints.log): a 32-byte store
// whose sector is not already in L2 costs one request out of ~36-40 G/s chip-wide, for 8 M tiny bins and for 40 K append cursors alike
// and whatever the L2 eviction hint; stores to resident sectors retire at 160+ G/s and whole-line stores cost one request per line.  A
// windowed form that exploited this (entries partitioned into windows of 296 bins with smem-staged tiles, then a persistent kernel
// scattering each window into an L2-resident ring of slabs and finishing the previous window) was built, was bit-exact, and ran 3x
// SLOWER (9.8 ms vs 3.0 ms at 60 M entries): with only a window's 296 bin counters live, the slot-assigning atomics serialise per
// address (~110 ns each).  It was removed; profiles/r2/window_build_experiment_*.
constexpr int SLAB_CAP = 2048;        // records per slab (64 KB of shared memory)
constexpr int SLAB_THREADS = 512;
constexpr int SLAB_NB_MAX = 16384;    // buckets per bin (packed 16-bit counters in shared memory)
struct SlabParams {
    FastMod fm;
    uint64_t nb_magic;      // ceil(2^64 / nb): bucket / nb = umulhi64(bucket, nb_magic), exact for bucket < 2^32, nb <= 2^16
    uint64_t table_len;
    uint32_t bucket_lo, nb, n_bins, chunk;   // chunk: counters per thread in the scan (even, chunk / 2 odd: conflict-free word stride)
    int32_t position_offset;
};
__device__ __forceinline__ uint32_t slab_bin_of(uint32_t bucket, const SlabParams &p) { return (uint32_t)__umul64hi((uint64_t)bucket, p.nb_magic); }

// Rows of one k-mer sit next to each other in a FlatKmers as the finder emits it (one row per node of the k-mer's path,
// kmer_finder.py:223-240), so neighbouring lanes often hold the same bin.  In the RUNS instantiation the first lane of such a run
// reserves the whole run with ONE returning atomic and the run's records leave as one store request to consecutive slots: a run of two
// costs one request out of the ~36 G/s budget instead of two (profiles/r2/calibrate_store_groups.jsonl; 60 M entries in runs of two:
// 2.98 -> 2.39 ms, 1 B: 44.6 -> 34.2 ms, profiles/r2/slab_scatter_runs.log).  The order of the slots inside a slab is irrelevant: the
// finish pass ranks the records of a bucket by input index.  On rows without such runs the shuffles cost 3-5 %, so a sampling kernel
// looks at 4096 neighbouring pairs first and BOTH instantiations are launched: the one the sample did not choose returns at once.
// The reservation has two steps, so that the atomics of a lane's UNROLL records are all in flight before the first result is used.
struct SlabRun {
    uint32_t base;    // value returned by the run head's atomic (meaningful on the head lane), then the record's slot
    uint32_t head;    // lane of the run's head | valid << 5 | bucket inside the bin << 6
};
__device__ __forceinline__ SlabRun slab_reserve_issue(uint32_t *__restrict__ count, uint32_t bin, bool valid, int lane) {
    const uint32_t prev = __shfl_up_sync(0xffffffffu, valid ? bin : 0xffffffffu, 1);
    const bool head = lane == 0 || prev != bin || !valid;      // an invalid lane is a run of its own (its neighbour above sees bin 0xffffffff)
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    const int head_lane = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
    SlabRun r;
    r.base = 0;
    r.head = (uint32_t)head_lane | ((uint32_t)valid << 5);
    if (head && valid) {
        const uint32_t above = lane == 31 ? 0u : (heads & (0xffffffffu << (lane + 1)));
        r.base = atomicAdd(count + bin, (uint32_t)((above ? __ffs(above) - 1 : 32) - lane));
    }
    return r;
}
__device__ __forceinline__ uint32_t slab_reserve_slot(const SlabRun &r, int lane) {   // SLAB_CAP or more: invalid lane or overflowing slab
    const int head_lane = (int)(r.head & 31u);
    const uint32_t base = __shfl_sync(0xffffffffu, r.base, head_lane);
    return ((r.head >> 5) & 1u) ? base + (uint32_t)(lane - head_lane) : (uint32_t)SLAB_CAP;
}
// *runs = 1 iff at least 1/8 of 4096 pairs of neighbouring rows, spread over the input, hold the same k-mer.  keys: the k-mer column
// (stride_words = 1) or the k-mer field of 32-byte records (stride_words = 4).
constexpr int SLAB_SAMPLE_PAIRS = 4096;
__global__ void __launch_bounds__(256) slab_sample_runs_kernel(const uint64_t *__restrict__ keys, int stride_words, int64_t n, uint32_t *__restrict__ runs) {
    __shared__ uint32_t equal;
    if (threadIdx.x == 0) equal = 0;
    __syncthreads();
    uint32_t mine = 0;
    for (int q = threadIdx.x; q < SLAB_SAMPLE_PAIRS && n >= 2; q += blockDim.x) {
        const int64_t i = (int64_t)(((unsigned __int128)(uint64_t)(n - 1) * (uint64_t)q) / SLAB_SAMPLE_PAIRS);
        mine += __ldg(keys + i * stride_words) == __ldg(keys + (i + 1) * stride_words);
    }
    if (mine) atomicAdd(&equal, mine);
    __syncthreads();
    if (threadIdx.x == 0) *runs = equal >= SLAB_SAMPLE_PAIRS / 8;
}

template <int UNROLL, int MINB, bool RUNS>
__global__ void __launch_bounds__(256, MINB)
slab_scatter_kernel(int64_t n, const uint64_t *__restrict__ kmers, const uint32_t *__restrict__ nodes, const uint64_t *__restrict__ ref,
                    const float *__restrict__ af, SlabParams p, uint32_t *__restrict__ count, BinRecord *__restrict__ slab,
                    uint32_t *__restrict__ overflow, const uint32_t *__restrict__ runs) {
    if ((__ldg(runs) != 0u) != RUNS) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    // RUNS: warp-uniform trip count, the lanes of a warp stay together for the shuffles
    for (int64_t base = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; (RUNS ? base - lane : base) < n; base += stride * UNROLL) {
        unsigned long long km[UNROLL];
        uint32_t bin[UNROLL];
        SlabRun run[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t i = base + u * stride;
            km[u] = i < n ? __ldg(kmers + i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t i = base + u * stride;
            const uint32_t b = fastmod(km[u], p.fm) - p.bucket_lo;
            bin[u] = slab_bin_of(b, p);
            const bool valid = i < n && bin[u] < p.n_bins;
            if (RUNS) {
                run[u] = slab_reserve_issue(count, bin[u], valid, lane);
            } else {
                run[u].base = valid ? atomicAdd(count + bin[u], 1u) : (uint32_t)SLAB_CAP;
                run[u].head = 0;
            }
            run[u].head |= (b - bin[u] * p.nb) << 6;
        }
        if (RUNS) {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) run[u].base = slab_reserve_slot(run[u], lane);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t i = base + u * stride;
            if (i >= n) continue;
            if (run[u].base >= (uint32_t)SLAB_CAP) {
                *overflow = 1u;
                continue;
            }
            const unsigned long long nd = nodes ? __ldg(nodes + i) : 0u;
            const unsigned long long a = af ? __float_as_uint(__ldg(af + i)) : 0u;
            const unsigned long long r = ref ? __ldg(ref + i) : 0ull;
            const unsigned long long tag = (unsigned long long)(uint32_t)i | ((unsigned long long)(run[u].head >> 6) << 32);
            BinRecord *dst = slab + (size_t)bin[u] * SLAB_CAP + run[u].base;
            asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"(km[u]), "l"(r), "l"(nd | (a << 32)), "l"(tag) : "memory");
        }
    }
}

struct SlabOut {
    int32_t *h2i;
    uint32_t *nk;
    uint64_t *kmers_o;
    uint32_t *nodes_o;
    uint64_t *ref_o;
    float *af_o;
    uint32_t *perm_o;
    uint16_t *freq_o;   // NULL: frequencies not wanted (or zeroed by the caller)
};
struct SlabSmem {
    BinRecord *rec;        // SLAB_CAP records, filled by the bulk copy
    uint32_t *sidx;        // input index of the record placed at a slot
    uint16_t *order;       // order[output rank] = record
    uint32_t *cnt32;       // SLAB_THREADS * chunk packed 16-bit bucket counters -> offsets
    uint64_t *bar;         // mbarrier of the bulk copy
    uint32_t *warp_tot;    // SLAB_THREADS / 32 + 1 words
};
__device__ __forceinline__ SlabSmem slab_smem_layout(unsigned char *base, uint64_t *bar, uint32_t *warp_tot) {
    SlabSmem m;
    m.rec = (BinRecord *)base;
    m.sidx = (uint32_t *)(base + (size_t)SLAB_CAP * 32);
    m.order = (uint16_t *)(m.sidx + SLAB_CAP);
    m.cnt32 = (uint32_t *)(m.order + SLAB_CAP);
    m.bar = bar;
    m.warp_tot = warp_tot;
    return m;
}
// One bin: c records at src (global, 32-byte aligned) -> output positions gstart.. and the tables of the bin's buckets.  Called by
// every thread of the CTA; ends with a CTA barrier after which the shared memory may be reused.
__device__ __forceinline__ void slab_finish_bin(const SlabParams &p, const SlabSmem &m, uint32_t &phase, const BinRecord *src, uint32_t c,
                                                uint32_t gstart, uint32_t bin, const SlabOut &o, bool tables_vec) {
    constexpr int T = SLAB_THREADS, R = SLAB_CAP / SLAB_THREADS;
    BinRecord *rec = m.rec;
    uint32_t *sidx = m.sidx, *cnt32 = m.cnt32, *warp_tot = m.warp_tot;
    uint16_t *order = m.order;
    uint64_t *bar = m.bar;
    const uint16_t *start16 = (const uint16_t *)cnt32;
    int32_t *h2i = o.h2i;
    uint32_t *nk = o.nk, *nodes_o = o.nodes_o, *perm_o = o.perm_o;
    uint64_t *kmers_o = o.kmers_o, *ref_o = o.ref_o;
    float *af_o = o.af_o;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t words = p.chunk / 2;                                       // counter words per thread
    if (tid == 0 && c) {
        fence_proxy_async();
        mbar_expect_tx(bar, c * 32u);
        bulk_g2s(rec, src, c * 32u, bar);
    }
    for (uint32_t w = tid; w < words * T; w += T) cnt32[w] = 0u;
    __syncthreads();
    uint32_t bl[R], arr[R], idx[R];
    if (c) {
        mbar_wait(bar, phase);
        phase ^= 1u;
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const uint32_t j = tid + r * T;
        bl[r] = 0; arr[r] = 0; idx[r] = 0;
        if (j < c) {
            const unsigned long long w = rec[j].index;
            idx[r] = (uint32_t)w;
            bl[r] = (uint32_t)(w >> 32);
            const uint32_t sh = (bl[r] & 1u) << 4;
            arr[r] = (atomicAdd(cnt32 + (bl[r] >> 1), 1u << sh) >> sh) & 0xffffu;
        }
    }
    __syncthreads();
    // in-place exclusive scan of the packed counters: a thread owns `chunk` consecutive buckets
    uint32_t sum = 0;
    for (uint32_t w = 0; w < words; w++) {
        const uint32_t v = cnt32[tid * words + w];
        sum += (v & 0xffffu) + (v >> 16);
    }
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t wv = lane < T / 32 ? warp_tot[lane] : 0u;
        uint32_t winc = wv;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        if (lane < T / 32) warp_tot[lane] = winc - wv;
    }
    __syncthreads();
    uint32_t run = warp_tot[warp] + inc - sum;
    for (uint32_t w = 0; w < words; w++) {
        const uint32_t v = cnt32[tid * words + w];
        const uint32_t lo = run, hi = run + (v & 0xffffu);
        run = hi + (v >> 16);
        cnt32[tid * words + w] = lo | (hi << 16);
    }
    __syncthreads();
    // slot of every record inside the bin: grouped by bucket, arrival order inside a bucket
#pragma unroll
    for (int r = 0; r < R; r++) {
        const uint32_t j = tid + r * T;
        if (j < c) {
            arr[r] += start16[bl[r]];
            sidx[arr[r]] = idx[r];
        }
    }
    // both tables, every bucket of the bin (cfki:444-457): position of the run head, run length; 0 / 0 when empty
    const uint64_t g0 = (uint64_t)bin * p.nb;
    const uint32_t nb_here = (uint32_t)min((uint64_t)p.nb, p.table_len - g0);
    const int32_t head0 = (int32_t)gstart + p.position_offset;
    if (tables_vec) {
        for (uint32_t b = tid * 4; b < nb_here; b += T * 4) {
            const uint2 sw = *(const uint2 *)(start16 + b);
            const uint32_t s0 = sw.x & 0xffffu, s1 = sw.x >> 16, s2 = sw.y & 0xffffu, s3 = sw.y >> 16, s4 = start16[b + 4];
            const uint4 len = make_uint4(s1 - s0, s2 - s1, s3 - s2, s4 - s3);
            const uint4 head = make_uint4(len.x ? head0 + s0 : 0, len.y ? head0 + s1 : 0, len.z ? head0 + s2 : 0, len.w ? head0 + s3 : 0);
            if (b + 4 <= nb_here) {
                *(uint4 *)(nk + g0 + b) = len;
                *(uint4 *)(h2i + g0 + b) = head;
            } else {
                nk[g0 + b] = len.x;
                h2i[g0 + b] = (int32_t)head.x;
                if (b + 1 < nb_here) {
                    nk[g0 + b + 1] = len.y;
                    h2i[g0 + b + 1] = (int32_t)head.y;
                }
                if (b + 2 < nb_here) {
                    nk[g0 + b + 2] = len.z;
                    h2i[g0 + b + 2] = (int32_t)head.z;
                }
            }
        }
    } else {
        for (uint32_t b = tid; b < nb_here; b += T) {
            const uint32_t s0 = start16[b], len = start16[b + 1] - s0;
            nk[g0 + b] = len;
            h2i[g0 + b] = len ? head0 + (int32_t)s0 : 0;
        }
    }
    __syncthreads();
    // stable order inside a bucket: rank by input index among the records sharing the bucket
#pragma unroll
    for (int r = 0; r < R; r++) {
        const uint32_t j = tid + r * T;
        if (j < c) {
            const uint32_t lo = start16[bl[r]], hi = start16[bl[r] + 1];
            uint32_t rank = lo;
            if (hi - lo > 1)
                for (uint32_t q = lo; q < hi; q++) rank += sidx[q] < idx[r];   // (four at a time: no faster, the pass is not issue-bound)
            order[rank] = (uint16_t)j;
        }
    }
    __syncthreads();
    // set_frequencies (cfki:267-293) while the bucket's records are at hand: first[q] = no earlier entry of the bucket holds the same
    // (k-mer, ref offset) pair; frequency = number of first-flagged entries of the bucket with the entry's k-mer (uint16, wraps).
    // The flags reuse sidx, which is dead once `order` is complete.
    uint8_t *first = (uint8_t *)sidx;
    if (o.freq_o) {
        for (uint32_t q = tid; q < c; q += T) {
            const uint4 *src = (const uint4 *)(rec + order[q]);
            const uint4 a = src[0];
            const uint32_t lo = start16[src[1].w];
            uint8_t f = 1;
            for (uint32_t e = lo; e < q; e++) {
                const uint4 k = *(const uint4 *)(rec + order[e]);
                if (k.x == a.x && k.y == a.y && k.z == a.z && k.w == a.w) {
                    f = 0;
                    break;
                }
            }
            first[q] = f;
        }
        __syncthreads();
    }
    for (uint32_t q = tid; q < c; q += T) {
        const uint4 *src = (const uint4 *)(rec + order[q]);
        const uint4 a = src[0], b = src[1];
        const size_t dst = (size_t)gstart + q;
        if (kmers_o) kmers_o[dst] = ((uint64_t)a.y << 32) | a.x;
        if (ref_o) ref_o[dst] = ((uint64_t)a.w << 32) | a.z;
        if (nodes_o) nodes_o[dst] = b.x;
        if (af_o) af_o[dst] = __uint_as_float(b.y);
        if (perm_o) perm_o[dst] = b.z;
        if (o.freq_o) {
            const uint32_t lo = start16[b.w], hi = start16[b.w + 1];
            uint32_t f = 0;
            for (uint32_t e = lo; e < hi; e++) {
                const uint2 k = *(const uint2 *)(rec + order[e]);
                f += (k.x == a.x && k.y == a.y) ? first[e] : 0u;
            }
            o.freq_o[dst] = (uint16_t)f;
        }
    }
    __syncthreads();   // the next bin's bulk copy and counter reset overwrite what was just read
}

__global__ void __launch_bounds__(SLAB_THREADS, 2)
slab_finish_kernel(SlabParams p, const uint32_t *__restrict__ count, const uint32_t *__restrict__ bin_start, const BinRecord *__restrict__ slab, SlabOut o,
                   const uint32_t *__restrict__ overflow) {
    if (__ldg(overflow)) return;   // a slab overflowed in the scatter: the caller falls back to another path
    extern __shared__ __align__(128) unsigned char slab_smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t warp_tot[SLAB_THREADS / 32 + 1];
    const SlabSmem m = slab_smem_layout(slab_smem, &bar, warp_tot);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    const bool tables_vec = (p.nb % 4 == 0) && ((((uintptr_t)o.h2i | (uintptr_t)o.nk) & 15) == 0);
    for (uint32_t bin = blockIdx.x; bin < p.n_bins; bin += gridDim.x) {
        const uint32_t c = min(__ldg(count + bin), (uint32_t)SLAB_CAP);
        const uint32_t nxt = bin + gridDim.x;      // this CTA's next slab starts its way from HBM to L2 now (1 B entries: 32.8 -> 31.6 ms)
        if (threadIdx.x == 32 && nxt < p.n_bins) {
            const uint32_t cn = min(__ldg(count + nxt), (uint32_t)SLAB_CAP);
            if (cn) bulk_prefetch_l2(slab + (size_t)nxt * SLAB_CAP, cn * 32u);
        }
        slab_finish_bin(p, m, phase, slab + (size_t)bin * SLAB_CAP, c, __ldg(bin_start + bin), bin, o, tables_vec);
    }
}

// ---- hash-range partitioned build, sender side: ONE stable partition pass that emits whole 32-byte records ----------------------
// record = {kmer, ref_offset, node | af << 32, bucket}.  Every rank orders its FlatKmers shard by the rank that owns the bucket
// (fan-out = number of ranks, so each warp writes long runs per owner: whole lines) keeping the input order inside an owner, and
// the records travel with ONE all-to-all instead of one gather + one all-to-all per column.  The receiver builds its slice straight
// from the records (slab_scatter_records_kernel): its input index is the position in the received array, which is the global
// input order (rank-major) restricted to its bucket range.
constexpr int PP_THREADS = 256;
constexpr int PP_ITEMS = 8;
constexpr int PP_TILE = PP_THREADS * PP_ITEMS;
constexpr int PP_MAX_PARTS = 32;
__global__ void __launch_bounds__(PP_THREADS) part_hist_kernel(const uint64_t *__restrict__ kmers, int64_t n, FastMod fm, uint32_t part_size, int n_parts,
                                                              uint32_t *__restrict__ hist, int64_t n_tiles) {
    __shared__ uint32_t h[PP_MAX_PARTS];
    if (threadIdx.x < PP_MAX_PARTS) h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * PP_TILE;
#pragma unroll
    for (int it = 0; it < PP_ITEMS; it++) {
        const int64_t i = base + it * PP_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[fastmod(__ldg(kmers + i), fm) / part_size], 1u);
    }
    __syncthreads();
    if (threadIdx.x < n_parts) hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}
// element order inside a tile: warp-major, then item, then lane -- i.e. memory order (stable)
__global__ void __launch_bounds__(PP_THREADS)
part_scatter_records_kernel(int64_t n, const uint64_t *__restrict__ kmers, const uint32_t *__restrict__ nodes, const uint64_t *__restrict__ ref,
                            const float *__restrict__ af, FastMod fm, uint32_t part_size, int n_parts, const uint32_t *__restrict__ offsets,
                            int64_t n_tiles, BinRecord *__restrict__ out) {
    constexpr int WARPS = PP_THREADS / 32;
    __shared__ uint32_t warp_hist[WARPS][PP_MAX_PARTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < WARPS * PP_MAX_PARTS; d += PP_THREADS) (&warp_hist[0][0])[d] = 0;
    __syncthreads();
    const int64_t warp_base = (int64_t)blockIdx.x * PP_TILE + (int64_t)warp * (PP_ITEMS * 32);
    unsigned long long km[PP_ITEMS];
    uint32_t bucket[PP_ITEMS], rank[PP_ITEMS];
#pragma unroll
    for (int it = 0; it < PP_ITEMS; it++) {
        const int64_t i = warp_base + it * 32 + lane;
        km[it] = i < n ? __ldg(kmers + i) : 0ull;
    }
#pragma unroll
    for (int it = 0; it < PP_ITEMS; it++) {
        const int64_t i = warp_base + it * 32 + lane;
        const bool ok = i < n;
        const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
        bucket[it] = fastmod(km[it], fm);
        rank[it] = 0;
        if (ok) {
            const uint32_t d = bucket[it] / part_size;
            const uint32_t peers = digit_peers<5>(d, okmask);
            const uint32_t before = __popc(peers & ((1u << lane) - 1u));
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (lane == leader) {
                base = warp_hist[warp][d];
                warp_hist[warp][d] = base + __popc(peers);
            }
            base = __shfl_sync(peers, base, leader);
            rank[it] = base + before;
        }
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < n_parts) {   // per-warp counts -> start positions in the output
        uint32_t run = offsets[(int64_t)threadIdx.x * n_tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            const uint32_t c = warp_hist[w][threadIdx.x];
            warp_hist[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < PP_ITEMS; it++) {
        const int64_t i = warp_base + it * 32 + lane;
        if (i >= n) continue;
        const unsigned long long nd = nodes ? __ldg(nodes + i) : 0u;
        const unsigned long long a = af ? __float_as_uint(__ldg(af + i)) : 0u;
        const unsigned long long r = ref ? __ldg(ref + i) : 0ull;
        BinRecord *dst = out + warp_hist[warp][bucket[it] / part_size] + rank[it];
        asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"(km[it]), "l"(r), "l"(nd | (a << 32)), "l"((unsigned long long)bucket[it]) : "memory");
    }
}
__global__ void part_counts_from_offsets_kernel(const uint32_t *__restrict__ offsets, int64_t n_tiles, int n_parts, int64_t n, long long *__restrict__ counts) {
    const int p = threadIdx.x;
    if (p >= n_parts) return;
    const long long lo = offsets[(int64_t)p * n_tiles], hi = p + 1 < n_parts ? offsets[(int64_t)(p + 1) * n_tiles] : n;
    counts[p] = hi - lo;
}
// receiver: the slab scatter fed by records (index = position in the record array)
template <int UNROLL, int MINB, bool RUNS>
__global__ void __launch_bounds__(256, MINB)
slab_scatter_records_kernel(int64_t n, const BinRecord *__restrict__ in, SlabParams p, uint32_t *__restrict__ count, BinRecord *__restrict__ slab,
                            uint32_t *__restrict__ overflow, const uint32_t *__restrict__ runs) {
    if ((__ldg(runs) != 0u) != RUNS) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (int64_t base = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; (RUNS ? base - lane : base) < n; base += stride * UNROLL) {
        BinRecord rec[UNROLL];
        uint32_t bin[UNROLL];
        SlabRun run[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t i = base + u * stride;
            rec[u] = i < n ? load_record(in + i) : BinRecord{0, 0, 0, 0};
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t i = base + u * stride;
            const uint32_t b = (uint32_t)rec[u].index - p.bucket_lo;
            bin[u] = slab_bin_of(b, p);
            const bool valid = i < n && bin[u] < p.n_bins;
            if (RUNS) {
                run[u] = slab_reserve_issue(count, bin[u], valid, lane);
            } else {
                run[u].base = valid ? atomicAdd(count + bin[u], 1u) : (uint32_t)SLAB_CAP;
                run[u].head = 0;
            }
            run[u].head |= (b - bin[u] * p.nb) << 6;
        }
        if (RUNS) {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) run[u].base = slab_reserve_slot(run[u], lane);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t i = base + u * stride;
            if (i >= n) continue;
            if (run[u].base >= (uint32_t)SLAB_CAP) {
                *overflow = 1u;
                continue;
            }
            asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(slab + (size_t)bin[u] * SLAB_CAP + run[u].base), "l"(rec[u].kmer), "l"(rec[u].ref),
                         "l"(rec[u].node_af), "l"((unsigned long long)(uint32_t)i | ((unsigned long long)(run[u].head >> 6) << 32))
                         : "memory");
        }
    }
}
// fallback paths take columns
__global__ void unpack_records_kernel(const BinRecord *__restrict__ in, int64_t n, uint64_t *__restrict__ kmers, uint32_t *__restrict__ nodes,
                                      uint64_t *__restrict__ ref, float *__restrict__ af) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const BinRecord r = load_record(in + i);
        kmers[i] = r.kmer;
        if (nodes) nodes[i] = (uint32_t)r.node_af;
        if (ref) ref[i] = r.ref;
        if (af) af[i] = __uint_as_float((uint32_t)(r.node_af >> 32));
    }
}

// set_frequencies (cfki:267-293), pass 1: first[e] = 1 iff no earlier entry of the bucket has the same
// (k-mer, ref_offset) pair
// (the bucket of entry e is the key of its sort element, or, after the binned build, kmers[e] % modulo - bucket_lo)
// Both passes compare an entry with the other entries of its bucket, quadratic in the bucket length: buckets of more than
// HEAVY_BUCKET entries (one k-mer thousands of times: poly-A, N runs; or a tiny modulo) are left to the hash-set passes below
// and counted in *n_heavy.
constexpr uint32_t HEAVY_BUCKET = 256;
__global__ void freq_first_kernel(const unsigned long long *__restrict__ sorted, const uint64_t *__restrict__ kmers,
                                  const uint64_t *__restrict__ ref, const int32_t *__restrict__ h2i, const uint32_t *__restrict__ nk, int64_t n,
                                  uint8_t *__restrict__ first, FastMod fm, uint32_t bucket_lo, unsigned long long *__restrict__ n_heavy) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        uint64_t km = __ldg(kmers + e);
        const uint32_t b = sorted ? elem_key(__ldg(sorted + e)) : fastmod(km, fm) - bucket_lo;
        if (nk[b] > HEAVY_BUCKET) {
            atomicAdd(n_heavy, 1ull);
            continue;
        }
        int64_t s = h2i[b];
        uint64_t ro = ref ? __ldg(ref + e) : 0ull;
        uint8_t f = 1;
        for (int64_t c = e - 1; c >= s; c--) {
            if (__ldg(kmers + c) == km && (!ref || __ldg(ref + c) == ro)) {
                f = 0;
                break;
            }
        }
        first[e] = f;
    }
}
// pass 2: frequency[e] = number of first-flagged entries of the bucket with the same k-mer (uint16, wraps)
__global__ void freq_count_kernel(const unsigned long long *__restrict__ sorted, const uint64_t *__restrict__ kmers,
                                  const int32_t *__restrict__ h2i, const uint32_t *__restrict__ nk, const uint8_t *__restrict__ first,
                                  int64_t n, uint16_t *__restrict__ freq, FastMod fm, uint32_t bucket_lo) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        uint64_t km = __ldg(kmers + e);
        uint32_t b = sorted ? elem_key(__ldg(sorted + e)) : fastmod(km, fm) - bucket_lo;
        if (nk[b] > HEAVY_BUCKET) continue;
        int64_t s = h2i[b], t = s + nk[b];
        uint32_t c = 0;
        for (int64_t a = s; a < t; a++) c += (__ldg(kmers + a) == km) & first[a];
        freq[e] = (uint16_t)c;
    }
}

// ---- the same two passes for heavy buckets, linear in their length: open-addressing sets of ENTRY NUMBERS ----
// A slot holds the number of the entry that claimed it (compare-and-swap on an empty slot); a later entry compares its key with the
// key of the slot's entry (the columns are read-only here) and walks on if they differ.  Pass 1 claims a slot per distinct
// (k-mer, ref offset) pair: the claimant is the pair's one `first` entry (the reference counts a set, cfki:288: which entry
// stands for the pair does not matter).  Pass 2 claims a slot per distinct k-mer and adds the first-flagged entries to its
// counter, lanes of a warp holding the same slot add once.  Pass 3 reads the counter back for every entry.
constexpr uint32_t HSET_EMPTY = 0xffffffffu;
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}
// slot of the entry with key (km, ro) in `slots` (mask + 1 slots, a power of two); INSERT: entry e claims an empty slot (claimed = true),
// otherwise HSET_EMPTY when the key is absent.  key1 == nullptr: the key is km alone.
template <bool INSERT>
__device__ __forceinline__ uint32_t hset_slot(uint32_t *__restrict__ slots, uint32_t mask, uint32_t e, uint64_t km, uint64_t ro,
                                              const uint64_t *__restrict__ key0, const uint64_t *__restrict__ key1, bool &claimed) {
    uint32_t h = (uint32_t)mix64(km + mix64(key1 ? ro + 0x9E3779B97F4A7C15ull : 0ull)) & mask;
    claimed = false;
    for (;;) {
        uint32_t o = __ldcg(slots + h);
        if (o == HSET_EMPTY) {
            if (!INSERT) return HSET_EMPTY;
            o = atomicCAS(slots + h, HSET_EMPTY, e);
            if (o == HSET_EMPTY) {
                claimed = true;
                return h;
            }
        }
        if (__ldg(key0 + o) == km && (!key1 || __ldg(key1 + o) == ro)) return h;
        h = (h + 1u) & mask;
    }
}
__global__ void freq_heavy_pairs_kernel(const unsigned long long *__restrict__ sorted, const uint64_t *__restrict__ kmers,
                                        const uint64_t *__restrict__ ref, const uint32_t *__restrict__ nk, int64_t n, FastMod fm,
                                        uint32_t bucket_lo, uint32_t *__restrict__ slots, uint32_t mask, uint8_t *__restrict__ first) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t km = __ldg(kmers + e);
        const uint32_t b = sorted ? elem_key(__ldg(sorted + e)) : fastmod(km, fm) - bucket_lo;
        if (nk[b] <= HEAVY_BUCKET) continue;
        bool claimed;
        hset_slot<true>(slots, mask, (uint32_t)e, km, ref ? __ldg(ref + e) : 0ull, kmers, ref, claimed);
        first[e] = claimed;
    }
}
// COUNT: first-flagged entries are added to their k-mer's counter; otherwise the counter is read back into freq
template <bool COUNT>
__global__ void freq_heavy_kmers_kernel(const unsigned long long *__restrict__ sorted, const uint64_t *__restrict__ kmers,
                                        const uint32_t *__restrict__ nk, int64_t n, FastMod fm, uint32_t bucket_lo,
                                        uint32_t *__restrict__ slots, uint32_t mask, const uint8_t *__restrict__ first,
                                        uint32_t *__restrict__ counts, uint16_t *__restrict__ freq) {
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count (the lanes vote)
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e - lane < n; e += (int64_t)gridDim.x * blockDim.x) {
        uint32_t slot = HSET_EMPTY;
        if (e < n) {
            const uint64_t km = __ldg(kmers + e);
            const uint32_t b = sorted ? elem_key(__ldg(sorted + e)) : fastmod(km, fm) - bucket_lo;
            if (nk[b] > HEAVY_BUCKET && (!COUNT || first[e])) {
                bool claimed;
                slot = COUNT ? hset_slot<true>(slots, mask, (uint32_t)e, km, 0ull, kmers, nullptr, claimed)
                             : hset_slot<false>(slots, mask, (uint32_t)e, km, 0ull, kmers, nullptr, claimed);
                if (!COUNT) freq[e] = slot == HSET_EMPTY ? (uint16_t)0 : (uint16_t)__ldcg(counts + slot);
            }
        }
        if (COUNT && __any_sync(0xffffffffu, slot != HSET_EMPTY)) {
            const uint32_t same = __match_any_sync(0xffffffffu, slot);
            if (slot != HSET_EMPTY && lane == __ffs(same) - 1) atomicAdd(counts + slot, (uint32_t)__popc(same));
        }
    }
}

// flat_kmers.py:98-125: keep[i] = 0 for the first occurrence of a hash, 1 for every later one.  A set of entry numbers keyed by
// the hash (as above) with the smallest entry number seen per slot; linear whatever the multiplicity of a hash.
__global__ void first_occurrence_min_kernel(const uint64_t *__restrict__ hashes, int64_t n, uint32_t *__restrict__ slots, uint32_t mask,
                                            uint32_t *__restrict__ min_entry) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool claimed;
        const uint32_t slot = hset_slot<true>(slots, mask, (uint32_t)i, __ldg(hashes + i), 0ull, hashes, nullptr, claimed);
        if (__ldcg(min_entry + slot) > (uint32_t)i) atomicMin(min_entry + slot, (uint32_t)i);
    }
}
__global__ void non_first_kernel(const uint64_t *__restrict__ hashes, int64_t n, uint32_t *__restrict__ slots, uint32_t mask,
                                 const uint32_t *__restrict__ min_entry, uint8_t *__restrict__ keep) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool claimed;
        const uint32_t slot = hset_slot<false>(slots, mask, (uint32_t)i, __ldg(hashes + i), 0ull, hashes, nullptr, claimed);
        keep[i] = __ldg(min_entry + slot) != (uint32_t)i;
    }
}

// ---- side indexes keyed by node / reference position (reverse_kmer_index.py:59-84, reference_kmer_index.py:81-121) ----
template <typename K>
__global__ void group_keys_kernel(const K *__restrict__ keys, int64_t n, unsigned long long n_keys, unsigned long long *__restrict__ elems,
                                  unsigned int *__restrict__ bad) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long k = (unsigned long long)__ldg(keys + i);
        if (k >= n_keys) {
            *bad = 1u;
            k = 0;
        }
        elems[i] = (k << 32) | (unsigned long long)(uint32_t)i;
    }
}
// reference_kmer_index.py:92-118: ref_position_to_index[p] = first sorted position of p for every present position but
// the smallest (ediff1d(..., to_begin=0) leaves the first run unmarked), then fill_zeros_from_end gives every other
// slot the value of the next marked slot to its right.  Run head i >= 1 therefore owns the slots (key[i-1], key[i]],
// and the head of the second run also owns [0, key[0]].  Short ranges are filled in place, long ones are queued.
constexpr int FILL_INLINE = 32;
__global__ void ref_heads_fill_kernel(const unsigned long long *__restrict__ sorted, int64_t n, uint32_t *__restrict__ first,
                                      uint4 *__restrict__ work, unsigned int *__restrict__ n_work) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (i == 0) continue;
        const uint32_t k = elem_key(__ldg(sorted + i)), kp = elem_key(__ldg(sorted + i - 1));
        if (k == kp) continue;
        const uint32_t lo = (kp == elem_key(__ldg(sorted))) ? 0u : kp + 1u;
        if (k - lo < (uint32_t)FILL_INLINE) {
            for (uint32_t r = lo; r <= k; r++) first[r] = (uint32_t)i;
        } else {
            work[atomicAdd(n_work, 1u)] = make_uint4(lo, k, (uint32_t)i, 0u);
        }
    }
}
__global__ void ref_long_fill_kernel(const uint4 *__restrict__ work, const unsigned int *__restrict__ n_work, uint32_t *__restrict__ first) {
    const unsigned int nw = *n_work;
    for (unsigned int w = blockIdx.x; w < nw; w += gridDim.x) {
        const uint4 item = work[w];
        for (uint64_t r = (uint64_t)item.x + threadIdx.x; r <= item.y; r += blockDim.x) first[r] = item.z;
    }
}
__global__ void count_long_gaps_kernel(const unsigned long long *__restrict__ sorted, int64_t n, unsigned int *__restrict__ n_long) {
    unsigned int local = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (i == 0) continue;
        const uint32_t k = elem_key(__ldg(sorted + i)), kp = elem_key(__ldg(sorted + i - 1));
        if (k == kp) continue;
        const uint32_t lo = (kp == elem_key(__ldg(sorted))) ? 0u : kp + 1u;
        local += (k - lo >= (uint32_t)FILL_INLINE);
    }
    if (local) atomicAdd(n_long, local);
}

// shared with count.cu: stable sort of n packed (key << 32 | payload) elements held in `a` by the low key_bits of the key
int radix_sort_packed(Scratch &a, Scratch &b, Scratch &hist, int64_t n, int key_bits, const unsigned long long **sorted, cudaStream_t s) {
    SortBuffers bufs;
    bufs.elems[0].ptr = a.ptr;  bufs.elems[0].stream = a.stream;  a.ptr = nullptr;   // take ownership for the duration of the sort
    int rc = radix_sort_elems(bufs, n, key_bits, sorted, s);
    a.ptr = bufs.elems[0].ptr;  bufs.elems[0].ptr = nullptr;                          // hand both buffers back to the caller
    b.ptr = bufs.elems[1].ptr;  b.stream = s;  bufs.elems[1].ptr = nullptr;
    hist.ptr = bufs.hist.ptr;   hist.stream = s;  bufs.hist.ptr = nullptr;
    return rc;
}

static int bit_length(uint64_t v) {
    int b = 0;
    while (v) {
        b++;
        v >>= 1;
    }
    return b;
}

static int sort_by_bucket(const uint64_t *d_kmers, int64_t n, uint64_t modulo, uint32_t bucket_lo, uint32_t part_size, uint64_t max_key,
                          SortBuffers &bufs, const unsigned long long **sorted, cudaStream_t s) {
    GKI_TRY(bufs.elems[0].alloc((size_t)n * 8, s));
    int grid = grid_for(n, 256 * 4, device_info().sms * 16);
    bucket_keys_kernel<<<grid, 256, 0, s>>>(d_kmers, n, make_fastmod(modulo), bucket_lo, part_size, bufs.elems[0].as<unsigned long long>());
    GKI_CHECK_LAUNCH();
    return radix_sort_elems(bufs, n, bit_length(max_key), sorted, s);
}

}  // namespace gki

using namespace gki;

extern "C" {

static int build_range(const uint64_t *kmers, const uint32_t *nodes, const uint64_t *ref_offsets, const float *af, int64_t n,
                       uint64_t modulo, uint64_t bucket_lo, uint64_t bucket_hi, int64_t position_offset, int32_t flags,
                       int32_t *hashes_to_index, uint32_t *n_kmers, uint64_t *kmers_out, uint32_t *nodes_out, uint64_t *ref_out,
                       float *af_out, uint16_t *freq_out, uint32_t *perm_out, gki_stream_t stream, const void *records = nullptr) {
    // records (device, gki_partition_pack layout) replace the four input columns when given
    CallScope call(stream);
    cudaStream_t s = call.stream;
    if (records) {
        GKI_REQUIRE(is_device_ptr(records) && (((uintptr_t)records & 31) == 0), GKI_ERR_INVALID, "gki_index_build_records: records must be 32-byte aligned device memory");
        kmers = (const uint64_t *)records;   // placeholders for the presence checks below; never dereferenced as columns
        nodes = (const uint32_t *)records;
        ref_offsets = (const uint64_t *)records;
        af = (const float *)records;
    }
    GKI_REQUIRE(bucket_lo < bucket_hi && bucket_hi <= modulo, GKI_ERR_INVALID, "gki_index_build: bad bucket range");
    GKI_REQUIRE(position_offset >= 0 && position_offset + n < (1ll << 31), GKI_ERR_UNSUPPORTED, "gki_index_build: positions must stay < 2^31");
    const uint64_t table_len = bucket_hi - bucket_lo;
    GKI_REQUIRE(n >= 1, GKI_ERR_INVALID, "gki_index_build: empty FlatKmers (the reference raises IndexError, cfki:455)");
    GKI_REQUIRE(n < (1ll << 31), GKI_ERR_UNSUPPORTED, "gki_index_build: n must be < 2^31 (int32 hashes_to_index, cfki:453)");
    GKI_REQUIRE(modulo >= 1 && modulo < (1ull << 32), GKI_ERR_UNSUPPORTED, "gki_index_build: need 1 <= modulo < 2^32");
    GKI_REQUIRE(kmers && hashes_to_index && n_kmers, GKI_ERR_INVALID, "gki_index_build: kmers / table outputs are NULL");
    GKI_REQUIRE((!nodes_out || nodes) && (!ref_out || ref_offsets) && (!af_out || af), GKI_ERR_INVALID,
                "gki_index_build: output column requested without its input");
    const bool want_freq = freq_out && !(flags & GKI_BUILD_SKIP_FREQUENCIES);

    DevIn d_kmers, d_nodes, d_ref, d_af;
    if (!records) {
        GKI_TRY(d_kmers.stage(kmers, (size_t)n * 8, s));
        GKI_TRY(d_nodes.stage(nodes_out ? nodes : nullptr, (size_t)n * 4, s));
        GKI_TRY(d_ref.stage((ref_out || want_freq) ? ref_offsets : nullptr, (size_t)n * 8, s));
        GKI_TRY(d_af.stage(af_out ? af : nullptr, (size_t)n * 4, s));
    }
    auto columns_from_records = [&]() -> int {   // the fallback paths read columns
        if (!records || d_kmers.dptr) return GKI_OK;
        GKI_TRY(d_kmers.scratch.alloc((size_t)n * 8, s));
        d_kmers.dptr = d_kmers.scratch.ptr;
        if (nodes_out) {
            GKI_TRY(d_nodes.scratch.alloc((size_t)n * 4, s));
            d_nodes.dptr = d_nodes.scratch.ptr;
        }
        if (ref_out || want_freq) {
            GKI_TRY(d_ref.scratch.alloc((size_t)n * 8, s));
            d_ref.dptr = d_ref.scratch.ptr;
        }
        if (af_out) {
            GKI_TRY(d_af.scratch.alloc((size_t)n * 4, s));
            d_af.dptr = d_af.scratch.ptr;
        }
        unpack_records_kernel<<<grid_for(n, 256 * 4, device_info().sms * 16), 256, 0, s>>>((const BinRecord *)records, n, (uint64_t *)d_kmers.dptr,
                                                                                         (uint32_t *)d_nodes.dptr, (uint64_t *)d_ref.dptr, (float *)d_af.dptr);
        GKI_CHECK_LAUNCH();
        return GKI_OK;
    };
    DevOut o_h2i, o_nk, o_kmers, o_nodes, o_ref, o_af, o_freq, o_perm;
    GKI_TRY(o_h2i.prepare(hashes_to_index, (size_t)table_len * 4, s));
    GKI_TRY(o_nk.prepare(n_kmers, (size_t)table_len * 4, s));
    GKI_TRY(o_kmers.prepare(kmers_out, (size_t)n * 8, s));
    GKI_TRY(o_nodes.prepare(nodes_out, (size_t)n * 4, s));
    GKI_TRY(o_ref.prepare(ref_out, (size_t)n * 8, s));
    GKI_TRY(o_af.prepare(af_out, (size_t)n * 4, s));
    GKI_TRY(o_freq.prepare(freq_out, (size_t)n * 2, s));
    GKI_TRY(o_perm.prepare(perm_out, (size_t)n * 4, s));

    const int grid_n = grid_for(n, 256 * 4, device_info().sms * 16);
    // the frequency pass needs sorted k-mers (and ref offsets) even if the caller did not ask for them
    Scratch tmp_kmers, tmp_ref;
    uint64_t *kmers_sorted = o_kmers.as<uint64_t>();
    uint64_t *ref_sorted = o_ref.as<uint64_t>();
    if (want_freq && !kmers_sorted) {
        GKI_TRY(tmp_kmers.alloc((size_t)n * 8, s));
        kmers_sorted = tmp_kmers.as<uint64_t>();
    }
    if (want_freq && !ref_sorted && (d_ref.dptr || records)) {
        GKI_TRY(tmp_ref.alloc((size_t)n * 8, s));
        ref_sorted = tmp_ref.as<uint64_t>();
    }
    const FastMod fm = make_fastmod(modulo);

    bool binned = false;
    SortBuffers bufs;
    const unsigned long long *sorted = nullptr;
    const char *force_path = getenv("GKI_BUILD_PATH");   // tests force a path: "slab" (default first choice), "binned", "radix"
    const bool allow_slab = !force_path || !strcmp(force_path, "slab");
    const bool allow_binned = (!force_path || !strcmp(force_path, "binned") || !strcmp(force_path, "slab"));
    // ---- slab path (see slab_finish_kernel): bins of nb buckets expected to hold ~3/4 of a slab ----
    bool fold_offset = false, freq_done = false;
    if (n >= (1 << 15) && allow_slab) {
        SlabParams sp;
        sp.fm = fm;
        sp.bucket_lo = (uint32_t)bucket_lo;
        sp.table_len = table_len;
        const char *mean_env = experiment_knob("GKI_SLAB_MEAN");
        const double mean = mean_env ? atof(mean_env) : 0.75 * SLAB_CAP;
        double nb_f = mean * (double)table_len / (double)n;
        uint32_t nb = nb_f >= (double)SLAB_NB_MAX ? (uint32_t)SLAB_NB_MAX : (uint32_t)nb_f;
        nb &= ~3u;
        if (nb < 4) nb = 4;
        sp.nb = nb;
        sp.nb_magic = (~0ull) / nb + 1ull;
        const uint64_t n_bins = (table_len + nb - 1) / nb;
        uint32_t chunk = (nb + 1 + SLAB_THREADS - 1) / SLAB_THREADS;
        chunk += chunk & 1u;
        if (((chunk / 2) & 1u) == 0) chunk += 2;
        sp.chunk = chunk;
        sp.n_bins = (uint32_t)n_bins;
        fold_offset = true;          // (the separate frequency pass of the fallback paths reads hashes_to_index as local positions)
        sp.position_offset = (int32_t)position_offset;
        const size_t smem = (size_t)SLAB_CAP * 38 + (size_t)SLAB_THREADS * chunk * 2;
        const SlabOut so{o_h2i.as<int32_t>(), o_nk.as<uint32_t>(), kmers_sorted, o_nodes.as<uint32_t>(), ref_sorted, o_af.as<float>(), o_perm.as<uint32_t>(),
                         want_freq ? o_freq.as<uint16_t>() : nullptr};
        Scratch counts, starts, slab, flag;
        if (n_bins < (1ull << 31) / SLAB_CAP * 64 && smem <= device_info().smem_optin && slab.try_alloc((size_t)n_bins * SLAB_CAP * sizeof(BinRecord), s)) {
            GKI_TRY(counts.alloc((size_t)(n_bins + 1) * 4, s));
            GKI_TRY(starts.alloc((size_t)(n_bins + 1) * 4, s));
            GKI_TRY(flag.alloc(8, s));                      // [overflow, runs]
            GKI_CUDA(cudaMemsetAsync(counts.ptr, 0, (size_t)(n_bins + 1) * 4, s));
            GKI_CUDA(cudaMemsetAsync(flag.ptr, 0, 8, s));
            // the scatter kernel for rows with / without runs of one k-mer: chosen on the device by a sample of neighbouring pairs
            uint32_t *runs_flag = flag.as<uint32_t>() + 1;
            const char *force_runs = getenv("GKI_SLAB_RUNS");   // tests force a kernel: "0" / "1"
            if (force_runs) {
                const uint32_t v = force_runs[0] == '1';
                GKI_CUDA(cudaMemcpyAsync(runs_flag, &v, 4, cudaMemcpyHostToDevice, s));
            } else {
                slab_sample_runs_kernel<<<1, 256, 0, s>>>(records ? (const uint64_t *)records : d_kmers.as<uint64_t>(), records ? 4 : 1, n, runs_flag);
                GKI_CHECK_LAUNCH();
            }
            // Both slab kernels get many more CTAs than are resident: what a CTA takes is then decided as slots free up, and SMs that
            // run slower (the far die, a busy L2 slice) take less.  One resident wave with a static stride: 60 M entries 2.99 / 2.35 ms
            // (shuffled rows / rows in finder order), 1 B 43.7 / 31.4 ms; with these grids 2.84 / 2.20 and 41.1 / 29.3 ms
            // (profiles/r2/slab_grid_oversubscription.log).
            int scatter_mult = 32, finish_mult = 32;
            if (const char *e = experiment_knob("GKI_SLAB_SCATTER_MULT")) scatter_mult = atoi(e) > 0 ? atoi(e) : scatter_mult;
            if (const char *e = experiment_knob("GKI_SLAB_FINISH_MULT")) finish_mult = atoi(e) > 0 ? atoi(e) : finish_mult;
            const int sgrid = grid_for(n, 256 * 4, device_info().sms * 8 * scatter_mult);
            if (records) {
                slab_scatter_records_kernel<4, 5, false><<<sgrid, 256, 0, s>>>(n, (const BinRecord *)records, sp, counts.as<uint32_t>(), slab.as<BinRecord>(), flag.as<uint32_t>(), runs_flag);
                GKI_CHECK_LAUNCH();
                slab_scatter_records_kernel<3, 5, true><<<sgrid, 256, 0, s>>>(n, (const BinRecord *)records, sp, counts.as<uint32_t>(), slab.as<BinRecord>(), flag.as<uint32_t>(), runs_flag);
            } else {
                slab_scatter_kernel<4, 5, false><<<sgrid, 256, 0, s>>>(n, d_kmers.as<uint64_t>(), d_nodes.as<uint32_t>(), d_ref.as<uint64_t>(), d_af.as<float>(), sp, counts.as<uint32_t>(), slab.as<BinRecord>(), flag.as<uint32_t>(), runs_flag);
                GKI_CHECK_LAUNCH();
                slab_scatter_kernel<3, 5, true><<<sgrid, 256, 0, s>>>(n, d_kmers.as<uint64_t>(), d_nodes.as<uint32_t>(), d_ref.as<uint64_t>(), d_af.as<float>(), sp, counts.as<uint32_t>(), slab.as<BinRecord>(), flag.as<uint32_t>(), runs_flag);
            }
            GKI_CHECK_LAUNCH();
            GKI_TRY(exclusive_scan_u32(counts.as<uint32_t>(), starts.as<uint32_t>(), (int64_t)n_bins + 1, nullptr, s));
            // the finish pass is queued behind the scatter without a host round trip: it looks at the overflow flag itself and
            // leaves at once when a slab overflowed (the fallback paths below then write every output)
            GKI_CUDA(cudaFuncSetAttribute(slab_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int grid = (int)std::min<uint64_t>(n_bins, (uint64_t)device_info().sms * 2 * finish_mult);
            slab_finish_kernel<<<grid, SLAB_THREADS, smem, s>>>(sp, counts.as<uint32_t>(), starts.as<uint32_t>(), slab.as<BinRecord>(), so, flag.as<uint32_t>());
            GKI_CHECK_LAUNCH();
            uint32_t overflow = 0;
            GKI_CUDA(cudaMemcpyAsync(&overflow, flag.ptr, 4, cudaMemcpyDeviceToHost, s));
            GKI_CUDA(cudaStreamSynchronize(s));
            if (!overflow) {
                binned = true;
                freq_done = want_freq;   // computed in the finish pass
            } else {
                fold_offset = false;
            }
        } else {
            fold_offset = false;
        }
    }

    if (!binned) GKI_TRY(columns_from_records());
    // ---- binned path (see bin_finish_kernel): applies when every bin of 2^shift buckets holds at most BIN_CAP entries ----
    if (!binned && n >= (1 << 15) && allow_binned) {
        BinParams bp;
        bp.fm = fm;
        bp.bucket_lo = (uint32_t)bucket_lo;
        bp.table_len = table_len;
        bp.shift = 0;   // about 16 entries per bin (one record per lane in the finish kernel), and few enough bins for their
                        // counters to stay in L2
        while (bp.shift < (uint32_t)BIN_MAX_SHIFT &&
               ((((uint64_t)n << bp.shift) / table_len) < 16 || ((table_len + (1ull << bp.shift) - 1) >> bp.shift) > (8ull << 20)))
            bp.shift++;
        const uint64_t n_bins = (table_len + (1ull << bp.shift) - 1) >> bp.shift;
        if (n_bins <= (8ull << 20) || bp.shift == (uint32_t)BIN_MAX_SHIFT) {
            bp.n_bins = (uint32_t)n_bins;
            Scratch counts, starts, cursor, maxbuf;
            GKI_TRY(counts.alloc((size_t)(n_bins + 1) * 4, s));
            GKI_TRY(starts.alloc((size_t)(n_bins + 1) * 4, s));
            GKI_TRY(maxbuf.alloc(4, s));
            GKI_CUDA(cudaMemsetAsync(counts.ptr, 0, (size_t)(n_bins + 1) * 4, s));
            GKI_CUDA(cudaMemsetAsync(maxbuf.ptr, 0, 4, s));
            bin_hist_kernel<<<grid_n, 256, 0, s>>>(d_kmers.as<uint64_t>(), n, bp, counts.as<uint32_t>());
            GKI_CHECK_LAUNCH();
            bin_max_kernel<<<grid_for((int64_t)n_bins, 256 * 4, device_info().sms * 8), 256, 0, s>>>(counts.as<uint32_t>(), bp.n_bins, maxbuf.as<uint32_t>());
            GKI_CHECK_LAUNCH();
            uint32_t max_bin = 0;
            GKI_CUDA(cudaMemcpyAsync(&max_bin, maxbuf.ptr, 4, cudaMemcpyDeviceToHost, s));
            GKI_CUDA(cudaStreamSynchronize(s));
            if (max_bin <= (uint32_t)BIN_CAP) {
                binned = true;
                GKI_TRY(exclusive_scan_u32(counts.as<uint32_t>(), starts.as<uint32_t>(), (int64_t)n_bins + 1, nullptr, s));   // starts[n_bins] = n
                GKI_TRY(cursor.alloc((size_t)n_bins * 4, s));
                GKI_CUDA(cudaMemcpyAsync(cursor.ptr, starts.ptr, (size_t)n_bins * 4, cudaMemcpyDeviceToDevice, s));
                Scratch records;
                GKI_TRY(records.alloc((size_t)n * sizeof(BinRecord), s));
                const int finish_grid = grid_for((int64_t)n_bins, BIN_WARPS, device_info().sms * 16);
                bin_scatter_kernel<<<grid_n, 256, 0, s>>>(n, d_kmers.as<uint64_t>(), d_nodes.as<uint32_t>(), d_ref.as<uint64_t>(), d_af.as<float>(), bp,
                                                          cursor.as<uint32_t>(), records.as<BinRecord>());
                GKI_CHECK_LAUNCH();
                if (max_bin <= (uint32_t)BIN_SMALL_CAP)
                    bin_finish_small_kernel<<<finish_grid, BIN_WARPS * 32, 0, s>>>(bp, starts.as<uint32_t>(), records.as<BinRecord>(), o_h2i.as<int32_t>(), o_nk.as<uint32_t>(),
                                                                                     kmers_sorted, o_nodes.as<uint32_t>(), ref_sorted, o_af.as<float>(), o_perm.as<uint32_t>());
                else
                    bin_finish_kernel<<<finish_grid, BIN_WARPS * 32, 0, s>>>(bp, starts.as<uint32_t>(), records.as<BinRecord>(), o_h2i.as<int32_t>(), o_nk.as<uint32_t>(),
                                                                               kmers_sorted, o_nodes.as<uint32_t>(), ref_sorted, o_af.as<float>(), o_perm.as<uint32_t>());
                GKI_CHECK_LAUNCH();
            }
        }
    }

    if (!binned) {
    // ---- radix path: stable LSD sort of (bucket << 32 | index) elements, run heads, payload gather ----
    // (a k-mer outside [bucket_lo, bucket_hi) would wrap to a huge key; the caller routes entries by range first)
    GKI_TRY(sort_by_bucket(d_kmers.as<uint64_t>(), n, modulo, (uint32_t)bucket_lo, 0u, table_len - 1, bufs, &sorted, s));

    GKI_CUDA(cudaMemsetAsync(o_h2i.dptr, 0, (size_t)table_len * 4, s));
    GKI_CUDA(cudaMemsetAsync(o_nk.dptr, 0, (size_t)table_len * 4, s));
    Scratch long_runs;
    GKI_TRY(long_runs.alloc(4, s));
    GKI_CUDA(cudaMemsetAsync(long_runs.ptr, 0, 4, s));
    run_heads_kernel<<<grid_n, 256, 0, s>>>(sorted, n, o_h2i.as<int32_t>(), o_nk.as<uint32_t>(), long_runs.as<unsigned int>());
    GKI_CHECK_LAUNCH();
    // buckets holding more than RUN_WALK entries (tiny modulo, or one heavily repeated k-mer): the tails pass exits
    // immediately unless the heads pass flagged one
    run_tails_kernel<<<grid_n, 256, 0, s>>>(sorted, n, o_h2i.as<int32_t>(), o_nk.as<uint32_t>(), long_runs.as<unsigned int>());
    GKI_CHECK_LAUNCH();

    const int payload_columns = (kmers_sorted != nullptr) + (o_nodes.dptr != nullptr) + (ref_sorted != nullptr) + (o_af.dptr != nullptr);
    if (payload_columns >= 2) {   // one interleaved record per entry: one random HBM fetch instead of one per column
        const bool with_ref = ref_sorted != nullptr;
        Scratch records;
        GKI_TRY(records.alloc((size_t)n * (with_ref ? 32 : 16), s));
        if (with_ref) {
            pack_records_kernel<true><<<grid_n, 256, 0, s>>>(n, d_kmers.as<uint64_t>(), d_nodes.as<uint32_t>(), d_ref.as<uint64_t>(), d_af.as<float>(), records.as<uint4>());
            GKI_CHECK_LAUNCH();
            gather_records_kernel<true><<<grid_n, 256, 0, s>>>(sorted, n, records.as<uint4>(), kmers_sorted, o_nodes.as<uint32_t>(), ref_sorted, o_af.as<float>(), o_perm.as<uint32_t>());
            GKI_CHECK_LAUNCH();
        } else {
            pack_records_kernel<false><<<grid_n, 256, 0, s>>>(n, d_kmers.as<uint64_t>(), d_nodes.as<uint32_t>(), nullptr, d_af.as<float>(), records.as<uint4>());
            GKI_CHECK_LAUNCH();
            gather_records_kernel<false><<<grid_n, 256, 0, s>>>(sorted, n, records.as<uint4>(), kmers_sorted, o_nodes.as<uint32_t>(), nullptr, o_af.as<float>(), o_perm.as<uint32_t>());
            GKI_CHECK_LAUNCH();
        }
    } else {
        gather_payload_kernel<<<grid_n, 256, 0, s>>>(sorted, n, d_kmers.as<uint64_t>(), d_nodes.as<uint32_t>(), d_ref.as<uint64_t>(),
                                                     d_af.as<float>(), kmers_sorted, o_nodes.as<uint32_t>(), ref_sorted, o_af.as<float>(),
                                                     o_perm.as<uint32_t>());
        GKI_CHECK_LAUNCH();
    }
    }

    if (freq_out && !freq_done) {
        if (!want_freq) {
            GKI_CUDA(cudaMemsetAsync(o_freq.dptr, 0, (size_t)n * 2, s));   // cfki:270-274
        } else {
            Scratch first, heavy;
            GKI_TRY(first.alloc((size_t)n, s));
            GKI_TRY(heavy.alloc(8, s));
            GKI_CUDA(cudaMemsetAsync(heavy.ptr, 0, 8, s));
            freq_first_kernel<<<grid_n, 256, 0, s>>>(sorted, kmers_sorted, ref_sorted, o_h2i.as<int32_t>(), o_nk.as<uint32_t>(), n, first.as<uint8_t>(), fm,
                                                     (uint32_t)bucket_lo, heavy.as<unsigned long long>());
            GKI_CHECK_LAUNCH();
            freq_count_kernel<<<grid_n, 256, 0, s>>>(sorted, kmers_sorted, o_h2i.as<int32_t>(), o_nk.as<uint32_t>(), first.as<uint8_t>(), n,
                                                     o_freq.as<uint16_t>(), fm, (uint32_t)bucket_lo);
            GKI_CHECK_LAUNCH();
            unsigned long long n_heavy = 0;
            GKI_CUDA(cudaMemcpyAsync(&n_heavy, heavy.ptr, 8, cudaMemcpyDeviceToHost, s));
            GKI_CUDA(cudaStreamSynchronize(s));
            if (n_heavy) {   // entries of buckets longer than HEAVY_BUCKET: the hash-set passes (linear in the bucket length)
                uint64_t cap = 1024;
                while (cap < 2 * n_heavy) cap <<= 1;
                Scratch pairs, kslots, kcounts;
                GKI_TRY(pairs.alloc((size_t)cap * 4, s));
                GKI_TRY(kslots.alloc((size_t)cap * 4, s));
                GKI_TRY(kcounts.alloc((size_t)cap * 4, s));
                GKI_CUDA(cudaMemsetAsync(pairs.ptr, 0xff, (size_t)cap * 4, s));
                GKI_CUDA(cudaMemsetAsync(kslots.ptr, 0xff, (size_t)cap * 4, s));
                GKI_CUDA(cudaMemsetAsync(kcounts.ptr, 0, (size_t)cap * 4, s));
                const uint32_t hmask = (uint32_t)(cap - 1);
                freq_heavy_pairs_kernel<<<grid_n, 256, 0, s>>>(sorted, kmers_sorted, ref_sorted, o_nk.as<uint32_t>(), n, fm, (uint32_t)bucket_lo,
                                                               pairs.as<uint32_t>(), hmask, first.as<uint8_t>());
                GKI_CHECK_LAUNCH();
                freq_heavy_kmers_kernel<true><<<grid_n, 256, 0, s>>>(sorted, kmers_sorted, o_nk.as<uint32_t>(), n, fm, (uint32_t)bucket_lo, kslots.as<uint32_t>(),
                                                                     hmask, first.as<uint8_t>(), kcounts.as<uint32_t>(), o_freq.as<uint16_t>());
                GKI_CHECK_LAUNCH();
                freq_heavy_kmers_kernel<false><<<grid_n, 256, 0, s>>>(sorted, kmers_sorted, o_nk.as<uint32_t>(), n, fm, (uint32_t)bucket_lo, kslots.as<uint32_t>(),
                                                                      hmask, first.as<uint8_t>(), kcounts.as<uint32_t>(), o_freq.as<uint16_t>());
                GKI_CHECK_LAUNCH();
            }
        }
    }
    if (position_offset && !fold_offset) {
        add_offset_nonempty_kernel<<<grid_for((int64_t)table_len, 256 * 4, device_info().sms * 16), 256, 0, s>>>(
            o_h2i.as<int32_t>(), o_nk.as<uint32_t>(), (int64_t)table_len, (int32_t)position_offset);
        GKI_CHECK_LAUNCH();
    }
    GKI_TRY(o_h2i.finish(s));
    GKI_TRY(o_nk.finish(s));
    GKI_TRY(o_kmers.finish(s));
    GKI_TRY(o_nodes.finish(s));
    GKI_TRY(o_ref.finish(s));
    GKI_TRY(o_af.finish(s));
    GKI_TRY(o_freq.finish(s));
    GKI_TRY(o_perm.finish(s));
    return call.finish();
}

int gki_index_build(const uint64_t *kmers, const uint32_t *nodes, const uint64_t *ref_offsets, const float *af, int64_t n,
                    uint64_t modulo, int32_t flags, int32_t *hashes_to_index, uint32_t *n_kmers, uint64_t *kmers_out,
                    uint32_t *nodes_out, uint64_t *ref_out, float *af_out, uint16_t *freq_out, uint32_t *perm_out,
                    gki_stream_t stream) {
    GKI_REQUIRE(modulo >= 1 && modulo < (1ull << 32), GKI_ERR_UNSUPPORTED, "gki_index_build: need 1 <= modulo < 2^32");
    return build_range(kmers, nodes, ref_offsets, af, n, modulo, 0, modulo, 0, flags, hashes_to_index, n_kmers, kmers_out, nodes_out,
                       ref_out, af_out, freq_out, perm_out, stream);
}

int gki_index_build_range(const uint64_t *kmers, const uint32_t *nodes, const uint64_t *ref_offsets, const float *af, int64_t n,
                          uint64_t modulo, uint64_t bucket_lo, uint64_t bucket_hi, int64_t position_offset, int32_t flags,
                          int32_t *hashes_to_index, uint32_t *n_kmers, uint64_t *kmers_out, uint32_t *nodes_out, uint64_t *ref_out,
                          float *af_out, uint16_t *freq_out, gki_stream_t stream) {
    GKI_REQUIRE(modulo >= 1 && modulo < (1ull << 32), GKI_ERR_UNSUPPORTED, "gki_index_build_range: need 1 <= modulo < 2^32");
    return build_range(kmers, nodes, ref_offsets, af, n, modulo, bucket_lo, bucket_hi, position_offset, flags, hashes_to_index, n_kmers,
                       kmers_out, nodes_out, ref_out, af_out, freq_out, nullptr, stream);
}

int gki_index_build_records(const void *records, int64_t n, uint64_t modulo, uint64_t bucket_lo, uint64_t bucket_hi, int64_t position_offset,
                            int32_t flags, int32_t *hashes_to_index, uint32_t *n_kmers, uint64_t *kmers_out, uint32_t *nodes_out, uint64_t *ref_out,
                            float *af_out, uint16_t *freq_out, gki_stream_t stream) {
    GKI_REQUIRE(records, GKI_ERR_INVALID, "gki_index_build_records: records is NULL");
    GKI_REQUIRE(modulo >= 1 && modulo < (1ull << 32), GKI_ERR_UNSUPPORTED, "gki_index_build_records: need 1 <= modulo < 2^32");
    return build_range(nullptr, nullptr, nullptr, nullptr, n, modulo, bucket_lo, bucket_hi, position_offset, flags, hashes_to_index, n_kmers,
                       kmers_out, nodes_out, ref_out, af_out, freq_out, nullptr, stream, records);
}

int gki_partition_pack(const uint64_t *kmers, const uint32_t *nodes, const uint64_t *ref_offsets, const float *af, int64_t n, uint64_t modulo,
                       int32_t n_parts, void *records_out, int64_t *counts_out, gki_stream_t stream) {
    CallScope call(stream);
    cudaStream_t s = call.stream;
    GKI_REQUIRE(n >= 0 && n < (1ll << 31) && n_parts >= 1 && n_parts <= PP_MAX_PARTS && modulo >= 1 && modulo < (1ull << 32) && counts_out &&
                    (n == 0 || (kmers && records_out)),
                GKI_ERR_INVALID, "gki_partition_pack: bad arguments (n_parts <= 32)");
    GKI_REQUIRE(n == 0 || (is_device_ptr(records_out) && (((uintptr_t)records_out & 31) == 0)), GKI_ERR_INVALID,
                "gki_partition_pack: records_out must be 32-byte aligned device memory");
    DevOut o_counts;
    GKI_TRY(o_counts.prepare(counts_out, (size_t)n_parts * 8, s));
    GKI_CUDA(cudaMemsetAsync(o_counts.dptr, 0, (size_t)n_parts * 8, s));
    if (n > 0) {
        DevIn d_kmers, d_nodes, d_ref, d_af;
        GKI_TRY(d_kmers.stage(kmers, (size_t)n * 8, s));
        GKI_TRY(d_nodes.stage(nodes, (size_t)n * 4, s));
        GKI_TRY(d_ref.stage(ref_offsets, (size_t)n * 8, s));
        GKI_TRY(d_af.stage(af, (size_t)n * 4, s));
        const uint32_t part_size = (uint32_t)((modulo + n_parts - 1) / n_parts);
        const int64_t n_tiles = (n + PP_TILE - 1) / PP_TILE;
        const FastMod fm = make_fastmod(modulo);
        Scratch hist;
        GKI_TRY(hist.alloc((size_t)n_parts * n_tiles * 4, s));
        part_hist_kernel<<<(unsigned)n_tiles, PP_THREADS, 0, s>>>(d_kmers.as<uint64_t>(), n, fm, part_size, n_parts, hist.as<uint32_t>(), n_tiles);
        GKI_CHECK_LAUNCH();
        GKI_TRY(exclusive_scan_u32(hist.as<uint32_t>(), hist.as<uint32_t>(), (int64_t)n_parts * n_tiles, nullptr, s));
        part_counts_from_offsets_kernel<<<1, PP_MAX_PARTS, 0, s>>>(hist.as<uint32_t>(), n_tiles, n_parts, n, (long long *)o_counts.dptr);
        GKI_CHECK_LAUNCH();
        part_scatter_records_kernel<<<(unsigned)n_tiles, PP_THREADS, 0, s>>>(n, d_kmers.as<uint64_t>(), d_nodes.as<uint32_t>(), d_ref.as<uint64_t>(), d_af.as<float>(),
                                                                             fm, part_size, n_parts, hist.as<uint32_t>(), n_tiles, (BinRecord *)records_out);
        GKI_CHECK_LAUNCH();
    }
    GKI_TRY(o_counts.finish(s));
    return call.finish();
}

int gki_partition_by_bucket_range(const uint64_t *kmers, int64_t n, uint64_t modulo, int32_t n_parts, uint32_t *perm_out,
                                  int64_t *counts_out, gki_stream_t stream) {
    CallScope call(stream);
    cudaStream_t s = call.stream;
    GKI_REQUIRE(n >= 0 && n < (1ll << 31) && n_parts >= 1 && n_parts <= 1024 && modulo >= 1 && modulo < (1ull << 32) && counts_out &&
                    (n == 0 || (kmers && perm_out)),
                GKI_ERR_INVALID, "gki_partition_by_bucket_range: bad arguments");
    DevOut o_counts;
    GKI_TRY(o_counts.prepare(counts_out, (size_t)n_parts * 8, s));
    GKI_CUDA(cudaMemsetAsync(o_counts.dptr, 0, (size_t)n_parts * 8, s));
    if (n > 0) {
        DevIn d_kmers;
        DevOut o_perm;
        GKI_TRY(d_kmers.stage(kmers, (size_t)n * 8, s));
        GKI_TRY(o_perm.prepare(perm_out, (size_t)n * 4, s));
        const uint32_t part_size = (uint32_t)((modulo + n_parts - 1) / n_parts);
        SortBuffers bufs;
        const unsigned long long *sorted;
        GKI_TRY(sort_by_bucket(d_kmers.as<uint64_t>(), n, modulo, 0u, part_size, (uint64_t)n_parts - 1, bufs, &sorted, s));
        Scratch bounds;
        GKI_TRY(bounds.alloc((size_t)(n_parts + 1) * 8, s));
        part_bounds_kernel<<<(n_parts + 1 + 255) / 256, 256, 0, s>>>(sorted, n, n_parts, bounds.as<long long>());
        GKI_CHECK_LAUNCH();
        part_counts_kernel<<<(n_parts + 255) / 256, 256, 0, s>>>(bounds.as<long long>(), n_parts, (long long *)o_counts.dptr);
        GKI_CHECK_LAUNCH();
        extract_perm_kernel<<<grid_for(n, 256 * 4, device_info().sms * 16), 256, 0, s>>>(sorted, n, o_perm.as<uint32_t>());
        GKI_CHECK_LAUNCH();
        GKI_TRY(o_perm.finish(s));
    }
    GKI_TRY(o_counts.finish(s));
    return call.finish();
}

int gki_group_by_key(const void *keys, int32_t key_size, int64_t n, uint64_t n_keys, int32_t flags, uint32_t *perm_out,
                     uint32_t *first_out, uint32_t *count_out, gki_stream_t stream) {
    CallScope call(stream);
    cudaStream_t s = call.stream;
    const bool ref_mode = (flags & GKI_GROUP_REFERENCE_INDEX) != 0;
    GKI_REQUIRE((key_size == 4 || key_size == 8) && n >= 1 && n < (1ll << 32) - 1 && n_keys >= 1 && n_keys <= (1ull << 32) && keys,
                GKI_ERR_INVALID, "gki_group_by_key: need 4- or 8-byte keys, 1 <= n < 2^32-1, 1 <= n_keys <= 2^32");
    GKI_REQUIRE(!(ref_mode && count_out), GKI_ERR_INVALID, "gki_group_by_key: count_out is not defined with GKI_GROUP_REFERENCE_INDEX");
    DevIn d_keys;
    DevOut o_perm, o_first, o_count;
    GKI_TRY(d_keys.stage(keys, (size_t)n * key_size, s));
    GKI_TRY(o_perm.prepare(perm_out, (size_t)n * 4, s));
    GKI_TRY(o_first.prepare(first_out, (size_t)n_keys * 4, s));
    GKI_TRY(o_count.prepare(count_out, (size_t)n_keys * 4, s));
    Scratch flagbuf;
    GKI_TRY(flagbuf.alloc(16, s));
    GKI_CUDA(cudaMemsetAsync(flagbuf.ptr, 0, 16, s));
    unsigned int *d_flags = flagbuf.as<unsigned int>();   // [0] key out of range, [1] long runs / long gaps, [2] work items
    SortBuffers bufs;
    GKI_TRY(bufs.elems[0].alloc((size_t)n * 8, s));
    const int grid = grid_for(n, 256 * 4, device_info().sms * 16);
    if (key_size == 4) group_keys_kernel<uint32_t><<<grid, 256, 0, s>>>(d_keys.as<uint32_t>(), n, n_keys, bufs.elems[0].as<unsigned long long>(), d_flags);
    else group_keys_kernel<unsigned long long><<<grid, 256, 0, s>>>(d_keys.as<unsigned long long>(), n, n_keys, bufs.elems[0].as<unsigned long long>(), d_flags);
    GKI_CHECK_LAUNCH();
    const unsigned long long *sorted;
    GKI_TRY(radix_sort_elems(bufs, n, bit_length(n_keys - 1) ? bit_length(n_keys - 1) : 1, &sorted, s));
    if (o_perm.dptr) {
        extract_perm_kernel<<<grid, 256, 0, s>>>(sorted, n, o_perm.as<uint32_t>());
        GKI_CHECK_LAUNCH();
    }
    if (o_first.dptr) GKI_CUDA(cudaMemsetAsync(o_first.dptr, 0, (size_t)n_keys * 4, s));
    if (o_count.dptr) GKI_CUDA(cudaMemsetAsync(o_count.dptr, 0, (size_t)n_keys * 4, s));
    Scratch work;
    if (ref_mode && o_first.dptr) {
        count_long_gaps_kernel<<<grid, 256, 0, s>>>(sorted, n, d_flags + 1);
        GKI_CHECK_LAUNCH();
        unsigned int host_flags[2];
        GKI_CUDA(cudaMemcpyAsync(host_flags, d_flags, 8, cudaMemcpyDeviceToHost, s));
        GKI_CUDA(cudaStreamSynchronize(s));
        GKI_REQUIRE(host_flags[0] == 0, GKI_ERR_INVALID, "gki_group_by_key: a key is >= n_keys");
        GKI_TRY(work.alloc(((size_t)host_flags[1] + 1) * 16, s));
        ref_heads_fill_kernel<<<grid, 256, 0, s>>>(sorted, n, o_first.as<uint32_t>(), work.as<uint4>(), d_flags + 2);
        GKI_CHECK_LAUNCH();
        if (host_flags[1]) {
            ref_long_fill_kernel<<<device_info().sms * 8, 256, 0, s>>>(work.as<uint4>(), d_flags + 2, o_first.as<uint32_t>());
            GKI_CHECK_LAUNCH();
        }
    } else if (o_first.dptr || o_count.dptr) {
        Scratch tmp_first, tmp_count;   // run_heads writes both tables
        int32_t *h2i = (int32_t *)o_first.dptr;
        uint32_t *nk = (uint32_t *)o_count.dptr;
        if (!h2i) {
            GKI_TRY(tmp_first.alloc((size_t)n_keys * 4, s));
            h2i = tmp_first.as<int32_t>();
        }
        if (!nk) {
            GKI_TRY(tmp_count.alloc((size_t)n_keys * 4, s));
            nk = tmp_count.as<uint32_t>();
        }
        run_heads_kernel<<<grid, 256, 0, s>>>(sorted, n, h2i, nk, d_flags + 1);
        GKI_CHECK_LAUNCH();
        run_tails_kernel<<<grid, 256, 0, s>>>(sorted, n, h2i, nk, d_flags + 1);
        GKI_CHECK_LAUNCH();
        unsigned int bad = 0;
        GKI_CUDA(cudaMemcpyAsync(&bad, d_flags, 4, cudaMemcpyDeviceToHost, s));
        GKI_CUDA(cudaStreamSynchronize(s));
        GKI_REQUIRE(bad == 0, GKI_ERR_INVALID, "gki_group_by_key: a key is >= n_keys");
    }
    GKI_TRY(o_perm.finish(s));
    GKI_TRY(o_first.finish(s));
    GKI_TRY(o_count.finish(s));
    return call.finish();
}

int gki_gather(const void *src, int32_t item_size, const uint32_t *perm, int64_t n, void *out, gki_stream_t stream) {
    CallScope call(stream);
    cudaStream_t s = call.stream;
    GKI_REQUIRE(n >= 0 && (n == 0 || (src && perm && out)), GKI_ERR_INVALID, "gki_gather: bad arguments");
    GKI_REQUIRE(item_size == 1 || item_size == 2 || item_size == 4 || item_size == 8, GKI_ERR_INVALID, "gki_gather: item_size must be 1, 2, 4 or 8");
    if (n == 0) return GKI_OK;
    DevIn d_src, d_perm;
    DevOut o;
    GKI_TRY(d_src.stage(src, (size_t)n * item_size, s));
    GKI_TRY(d_perm.stage(perm, (size_t)n * 4, s));
    GKI_TRY(o.prepare(out, (size_t)n * item_size, s));
    int grid = grid_for(n, 256 * 4, device_info().sms * 16);
    switch (item_size) {
        case 1: gather_kernel<uint8_t><<<grid, 256, 0, s>>>(d_src.as<uint8_t>(), d_perm.as<uint32_t>(), n, o.as<uint8_t>()); break;
        case 2: gather_kernel<uint16_t><<<grid, 256, 0, s>>>(d_src.as<uint16_t>(), d_perm.as<uint32_t>(), n, o.as<uint16_t>()); break;
        case 4: gather_kernel<uint32_t><<<grid, 256, 0, s>>>(d_src.as<uint32_t>(), d_perm.as<uint32_t>(), n, o.as<uint32_t>()); break;
        default: gather_kernel<uint64_t><<<grid, 256, 0, s>>>(d_src.as<uint64_t>(), d_perm.as<uint32_t>(), n, o.as<uint64_t>()); break;
    }
    GKI_CHECK_LAUNCH();
    GKI_TRY(o.finish(s));
    return call.finish();
}

int gki_mark_non_first_occurrences(const uint64_t *hashes, int64_t n, uint8_t *keep, gki_stream_t stream) {
    CallScope call(stream);
    cudaStream_t s = call.stream;
    GKI_REQUIRE(n >= 0 && n < (1ll << 31) && (n == 0 || (hashes && keep)), GKI_ERR_INVALID, "gki_mark_non_first_occurrences: bad arguments");
    if (n == 0) return GKI_OK;
    DevIn d_h;
    DevOut o;
    GKI_TRY(d_h.stage(hashes, (size_t)n * 8, s));
    GKI_TRY(o.prepare(keep, (size_t)n, s));
    uint64_t cap = 1024;
    while (cap < 2 * (uint64_t)n) cap <<= 1;
    Scratch slots, min_entry;
    GKI_TRY(slots.alloc((size_t)cap * 4, s));
    GKI_TRY(min_entry.alloc((size_t)cap * 4, s));
    GKI_CUDA(cudaMemsetAsync(slots.ptr, 0xff, (size_t)cap * 4, s));
    GKI_CUDA(cudaMemsetAsync(min_entry.ptr, 0xff, (size_t)cap * 4, s));
    const int grid = grid_for(n, 256 * 4, device_info().sms * 16);
    first_occurrence_min_kernel<<<grid, 256, 0, s>>>(d_h.as<uint64_t>(), n, slots.as<uint32_t>(), (uint32_t)(cap - 1), min_entry.as<uint32_t>());
    GKI_CHECK_LAUNCH();
    non_first_kernel<<<grid, 256, 0, s>>>(d_h.as<uint64_t>(), n, slots.as<uint32_t>(), (uint32_t)(cap - 1), min_entry.as<uint32_t>(), o.as<uint8_t>());
    GKI_CHECK_LAUNCH();
    GKI_TRY(o.finish(s));
    return call.finish();
}

}  // extern "C"
