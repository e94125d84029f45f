// count.cu -- K3 counting: CounterKmerIndex.count_kmers / get_node_counts (collision_free_kmer_index.py:14-40)
// and the fused read path (read_kmers.py:14-26 -> cfki:33-37).
//
// The reference keeps one counter per distinct index k-mer (npstructures.Counter keyed by k-mer, cfki:27) and
// get_node_counts = bincount(nodes, weights=counter[kmers]) (cfki:39-40).  The counts only depend on k-mer
// equality, so the device is free to key the counters however it likes.  Measured on B200 (profiles/r1):
// random 8-byte gathers run at ~217 G/s while the table fits L2 (<= 48 MB) and at ~37 G/s from HBM (>= 4 GB
// tables, 64-byte fetch granularity); the first version of this kernel, which probed the reference's
// modulo-bucket tables behind a 56 MB bucket bitmap, was bound by exactly those two rates (132 GB of DRAM reads
// per 2.4 G queries).  Hence this layout:
//
//   * Bloom filter, register-blocked (one 64-bit word per key, filter_k bits), sized <= ~32 MB so it stays in
//     L2: one L2 access decides most absent k-mers.
//   * bucketised open-addressing table over the distinct k-mers: bucket = 4 slots x {key, cnt[2]} = one 64-byte
//     line = ONE HBM access per surviving probe; the hit's RED lands on the line that was just fetched.
//   * canonical keys: key = min(x, revcomp_k(x)), cnt[o] with o = (x != key).  The forward and reverse-complement
//     hashes of a read position share the key, so a position (2 queries) costs one filter access, at most one
//     table access and one 64-bit RED (+1 on both orientations).
//   * multiply-xorshift hash + multiply-high range reduction: no 64-bit modulo in the hot loop.
#include <stdlib.h>
#include "index.cuh"
#include "reads_tile.cuh"

namespace gki {

constexpr int COUNT_THREADS = 256;

// ------------------------------------------------------------------ hashing / keys
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x *= 0x9E3779B97F4A7C15ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    return x ^ (x >> 32);
}

struct Key {
    unsigned long long c;   // table key
    uint32_t o;             // orientation (0: the query is the key itself)
    bool ok;                // false: cannot be in the table
};

// key of one query for a table of mode k (k == 0: raw)
__device__ __forceinline__ Key make_key(uint64_t q, int k) {
    Key key;
    if (k == 0) {
        key.c = q;
        key.o = 0;
        key.ok = true;
    } else {
        uint64_t r = revcomp_hash(q, k);
        key.c = q < r ? q : r;
        key.o = q != key.c;
        key.ok = (q >> (2 * k)) == 0;     // index k-mers are < 4^k (checked at build); anything larger is absent
    }
    return key;
}

__device__ __forceinline__ bool filter_pass(const TableView &t, uint64_t h) {
    if (!t.filter) return true;
    uint32_t word = __umulhi((uint32_t)h, t.filter_words);
    uint32_t hi = (uint32_t)(h >> 32);
    unsigned long long mask = 1ull << (hi & 63);
    if (t.filter_k > 1) mask |= 1ull << ((hi >> 6) & 63);
    if (t.filter_k > 2) mask |= 1ull << ((hi >> 12) & 63);
    return (__ldg(t.filter + word) & mask) == mask;
}
__device__ __forceinline__ unsigned long long filter_mask(const TableView &t, uint64_t h) {
    uint32_t hi = (uint32_t)(h >> 32);
    unsigned long long mask = 1ull << (hi & 63);
    if (t.filter_k > 1) mask |= 1ull << ((hi >> 6) & 63);
    if (t.filter_k > 2) mask |= 1ull << ((hi >> 12) & 63);
    return mask;
}
__device__ __forceinline__ uint32_t home_bucket(const TableView &t, uint64_t h) {
    return __umulhi((uint32_t)(h >> 32), t.n_buckets);
}

// slot holding key c, or nullptr.  Probing visits buckets linearly and stops at the first non-full bucket.
__device__ __forceinline__ Slot *find_slot(const TableView &t, unsigned long long c, uint64_t h) {
    if (c == SLOT_EMPTY) {   // raw mode only: the one value that collides with the empty marker has its own slot,
        Slot *sp = t.slots + (size_t)t.n_buckets * SLOTS_PER_BUCKET;   // whose key field is 1 iff that value is indexed
        return __ldg(&sp->key) == 1ull ? sp : nullptr;
    }
    uint32_t b = home_bucket(t, h);
    for (uint32_t tries = 0; tries < t.n_buckets; tries++) {
        Slot *base = t.slots + (size_t)b * SLOTS_PER_BUCKET;
        unsigned long long key[SLOTS_PER_BUCKET];   // keys never change while a counting kernel runs: all four loads in flight
#pragma unroll
        for (int i = 0; i < SLOTS_PER_BUCKET; i++) key[i] = __ldg(&base[i].key);
#pragma unroll
        for (int i = 0; i < SLOTS_PER_BUCKET; i++) {
            if (key[i] == c) return base + i;
            if (key[i] == SLOT_EMPTY) return nullptr;
        }
        b = (b + 1 == t.n_buckets) ? 0 : b + 1;
    }
    return nullptr;
}

// one independent query
__device__ __forceinline__ void count_one(const TableView &t, uint64_t q) {
    Key key = make_key(q, t.k);
    if (!key.ok) return;
    uint64_t h = mix64(key.c);
    if (!filter_pass(t, h)) return;
    Slot *s = find_slot(t, key.c, h);
    if (s) atomicAdd(&s->cnt[key.o], 1u);
}

// ------------------------------------------------------------------ table construction
__global__ void table_init_kernel(Slot *__restrict__ slots, size_t n_slots) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_slots; i += (size_t)gridDim.x * blockDim.x) {
        Slot s;
        s.key = SLOT_EMPTY;
        s.cnt[0] = 0;
        s.cnt[1] = 0;
        slots[i] = s;
    }
}

// number of distinct k-mers: entry e is a representative iff no earlier entry of its bucket holds the same k-mer
__global__ void count_distinct_kernel(IndexView ix, int64_t n, unsigned long long *__restrict__ out) {
    unsigned int local = 0;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        uint64_t km = __ldg(ix.kmers + e);
        uint2 cell = __ldg(ix.cells + fastmod(km, ix.fm));
        bool rep = true;
        for (int64_t c = e - 1; c >= (int64_t)cell.x; c--)
            if (__ldg(ix.kmers + c) == km) {
                rep = false;
                break;
            }
        local += rep;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, (unsigned long long)local);
}

__global__ void table_insert_kernel(TableView t, const uint64_t *__restrict__ kmers, int64_t n, unsigned long long *__restrict__ filter,
                                    unsigned int *__restrict__ failed) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        Key key = make_key(__ldg(kmers + e), t.k);
        if (key.c == SLOT_EMPTY) {
            t.slots[(size_t)t.n_buckets * SLOTS_PER_BUCKET].key = 1ull;   // mark the special slot as present
            continue;
        }
        uint64_t h = mix64(key.c);
        if (filter) atomicOr(filter + __umulhi((uint32_t)h, t.filter_words), filter_mask(t, h));
        uint32_t b = home_bucket(t, h);
        bool placed = false;
        for (uint32_t tries = 0; tries < t.n_buckets && !placed; tries++) {
            Slot *base = t.slots + (size_t)b * SLOTS_PER_BUCKET;
            for (int i = 0; i < SLOTS_PER_BUCKET && !placed; i++) {
                unsigned long long cur = *(volatile unsigned long long *)&base[i].key;
                if (cur == SLOT_EMPTY) cur = atomicCAS(&base[i].key, SLOT_EMPTY, key.c);
                placed = (cur == SLOT_EMPTY) || (cur == key.c);
            }
            b = (b + 1 == t.n_buckets) ? 0 : b + 1;
        }
        if (!placed) atomicAdd(failed, 1u);
    }
}

__global__ void table_reset_kernel(Slot *__restrict__ slots, size_t n_slots) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_slots; i += (size_t)gridDim.x * blockDim.x)
        *(unsigned long long *)&slots[i].cnt[0] = 0ull;
}

// ------------------------------------------------------------------ counting kernels
__global__ void __launch_bounds__(COUNT_THREADS) count_kmers_kernel(TableView t, const uint64_t *__restrict__ queries, int64_t nq) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x)
        count_one(t, __ldg(queries + i));
}

// Fused K1 -> K3.  A warp owns a read; each lane takes WPL windows.  PAIRED (both strands, read k == table k):
// forward and reverse-complement hash of a fully valid window share one canonical key -> one probe, one RED.
constexpr int WPL = 4;

template <bool BOTH, bool PAIRED>
__global__ void __launch_bounds__(COUNT_THREADS) count_reads_kernel(TableView t, ReadBatch b) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint64_t mask = kmer_mask(b.k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for_each_tile(b, smem_raw, [&](int64_t tile, const TileSmem &ts) {
        int64_t r0 = tile * (int64_t)b.tile_reads;
        for (int r = warp; r < b.tile_reads && r0 + r < b.n_reads; r += nwarps) {
            const uint64_t *cw = ts.codes + (size_t)r * b.words;
            const uint64_t *vw = ts.valid + (size_t)r * b.words;
            for (int base = 0; base < b.nk; base += 32 * WPL) {
                uint64_t fwd[WPL], rc[WPL], h[WPL];
                unsigned long long c[WPL];
                uint32_t live = 0, slow = 0;
#pragma unroll
                for (int u = 0; u < WPL; u++) {
                    int i = base + u * 32 + lane;
                    bool ok = i < b.nk;
                    int ii = ok ? i : 0;
                    fwd[u] = extract_window(cw, ii, mask);
                    if (BOTH) {
                        uint64_t v = extract_window(vw, ii, mask);
                        rc[u] = revcomp_hash_masked(fwd[u], v, b.k);
                        if (PAIRED) {
                            bool clean = v == mask;            // no N: rc is the true reverse complement of fwd
                            c[u] = fwd[u] < rc[u] ? fwd[u] : rc[u];
                            live |= (uint32_t)(ok && clean) << u;
                            slow |= (uint32_t)(ok && !clean) << u;
                        } else {
                            slow |= (uint32_t)ok << u;
                        }
                    } else {
                        slow |= (uint32_t)ok << u;
                    }
                }
                if (PAIRED) {
                    // stage 1: filter words (L2) for all windows of the lane
                    uint64_t fw[WPL];
                    unsigned long long fm[WPL];
#pragma unroll
                    for (int u = 0; u < WPL; u++) {
                        h[u] = mix64(c[u]);
                        fm[u] = filter_mask(t, h[u]);
                        fw[u] = (t.filter && ((live >> u) & 1u)) ? __ldg(t.filter + __umulhi((uint32_t)h[u], t.filter_words)) : ~0ull;
                    }
#pragma unroll
                    for (int u = 0; u < WPL; u++) live &= ~((uint32_t)((fw[u] & fm[u]) != fm[u]) << u);
                    // stage 2: table (HBM) for the survivors
#pragma unroll
                    for (int u = 0; u < WPL; u++) {
                        if (!((live >> u) & 1u)) continue;
                        Slot *s = find_slot(t, c[u], h[u]);
                        if (!s) continue;
                        if (fwd[u] == rc[u]) atomicAdd(&s->cnt[0], 2u);                                   // palindrome (even k)
                        else atomicAdd((unsigned long long *)&s->cnt[0], 0x0000000100000001ull);        // +1 on both orientations
                    }
                }
                if (slow) {   // windows with non-ACGT bases, forward-only mode, or read k != table k: independent queries
#pragma unroll
                    for (int u = 0; u < WPL; u++) {
                        if (!((slow >> u) & 1u)) continue;
                        count_one(t, fwd[u]);
                        if (BOTH) count_one(t, rc[u]);
                    }
                }
            }
        }
    });
}

// ------------------------------------------------------------------ counters -> per-entry / per-node
__device__ __forceinline__ uint32_t kmer_count(const TableView &t, uint64_t km, bool wrap16) {
    Key key = make_key(km, t.k);
    uint32_t w = 0;
    if (key.ok) {
        Slot *s = find_slot(t, key.c, mix64(key.c));
        if (s) w = *(volatile uint32_t *)&s->cnt[key.o];
    }
    return wrap16 ? (w & 0xFFFFu) : w;
}

__global__ void node_counts_kernel(TableView t, const uint64_t *__restrict__ kmers, const uint32_t *__restrict__ nodes, int64_t n,
                                   double *__restrict__ out, int64_t n_out, bool wrap16) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w = kmer_count(t, __ldg(kmers + e), wrap16);
        uint32_t node = __ldg(nodes + e);
        if (w && (int64_t)node < n_out) atomicAdd(out + node, (double)w);
    }
}

__global__ void query_counts_kernel(TableView t, const uint64_t *__restrict__ queries, int64_t nq, uint32_t *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = kmer_count(t, __ldg(queries + i), false);
}

// ------------------------------------------------------------------ host side
void destroy_count_table(gki_index *ix) {
    cudaFree(ix->table.slots);
    cudaFree((void *)ix->table.filter);
    ix->table = TableView{};
    ix->table_bytes = ix->filter_bytes = 0;
}

// Build the table on first use.  k > 0 selects canonical keys (needs every index k-mer < 4^k), k == 0 raw keys.
static int ensure_table(gki_index *ix, int k, cudaStream_t s) {
    if (ix->table.slots) return GKI_OK;
    if (k < 0 || k > 31) k = 0;
    if (k && (ix->max_kmer >> (2 * k)) != 0) k = 0;     // index values wider than k bases: canonical form undefined
    if (const char *e = getenv("GKI_TABLE_RAW")) if (atoi(e)) k = 0;
    Scratch counters;
    GKI_TRY(counters.alloc(16, s));
    GKI_CUDA(cudaMemsetAsync(counters.ptr, 0, 16, s));
    const int grid_n = grid_for(ix->n, 256 * 4, device_info().sms * 16);
    count_distinct_kernel<<<grid_n, 256, 0, s>>>(ix->view(), ix->n, (unsigned long long *)counters.ptr);
    GKI_CHECK_LAUNCH();
    unsigned long long distinct = 0;
    GKI_CUDA(cudaMemcpyAsync(&distinct, counters.ptr, 8, cudaMemcpyDeviceToHost, s));
    GKI_CUDA(cudaStreamSynchronize(s));
    ix->n_distinct = (int64_t)distinct;

    TableView t{};
    t.k = k;
    uint64_t buckets = (distinct + 1) / 2 + 16;          // 4 slots per bucket -> load factor <= 0.5
    GKI_REQUIRE(buckets < (1ull << 32), GKI_ERR_UNSUPPORTED, "count table: too many distinct k-mers");
    t.n_buckets = (uint32_t)buckets;
    size_t n_slots = (size_t)buckets * SLOTS_PER_BUCKET + 1;
    GKI_CUDA(cudaMalloc((void **)&t.slots, n_slots * sizeof(Slot)));
    ix->table_bytes = n_slots * sizeof(Slot);
    table_init_kernel<<<grid_for((int64_t)n_slots, 256 * 4, device_info().sms * 16), 256, 0, s>>>(t.slots, n_slots);
    GKI_CHECK_LAUNCH();

    // Bloom filter: as many bits per key as fit the L2 budget (<= 16); below 1.5 bits per key it filters nothing
    size_t budget = (size_t)48 << 20;   // measured: random gathers stay at the L2 rate up to 48 MB (profiles/r1)
    if (const char *e = getenv("GKI_FILTER_MAX_MB")) budget = (size_t)atoi(e) << 20;
    size_t want = (size_t)distinct * 2;                  // 16 bits per key
    size_t fbytes = want < budget ? want : budget;
    fbytes = (fbytes + 7) & ~(size_t)7;
    if (fbytes < 64) fbytes = 64;
    double bits_per_key = distinct ? (double)fbytes * 8.0 / (double)distinct : 16.0;
    unsigned long long *filter = nullptr;
    if (budget > 0 && bits_per_key >= 1.5) {
        GKI_CUDA(cudaMalloc((void **)&filter, fbytes));
        GKI_CUDA(cudaMemsetAsync(filter, 0, fbytes, s));
        t.filter = (const uint64_t *)filter;
        t.filter_words = (uint32_t)(fbytes / 8);
        t.filter_k = bits_per_key >= 5.0 ? 3 : (bits_per_key >= 3.0 ? 2 : 1);
        if (const char *e = getenv("GKI_FILTER_K")) t.filter_k = atoi(e) < 1 ? 1 : (atoi(e) > 3 ? 3 : atoi(e));
        ix->filter_bytes = fbytes;
    }
    table_insert_kernel<<<grid_n, 256, 0, s>>>(t, ix->kmers, ix->n, filter, (unsigned int *)counters.ptr + 2);
    GKI_CHECK_LAUNCH();
    unsigned int failed = 0;
    GKI_CUDA(cudaMemcpyAsync(&failed, (unsigned int *)counters.ptr + 2, 4, cudaMemcpyDeviceToHost, s));
    GKI_CUDA(cudaStreamSynchronize(s));
    ix->table = t;
    GKI_REQUIRE(failed == 0, GKI_ERR_CUDA, "count table: %u insertions failed", failed);
    return GKI_OK;
}

static int ensure_staging(gki_index *ix, size_t bytes) {
    if (!ix->copy_stream) {
        GKI_CUDA(cudaStreamCreateWithFlags(&ix->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            GKI_CUDA(cudaEventCreateWithFlags(&ix->ready[i], cudaEventDisableTiming));
            GKI_CUDA(cudaEventCreateWithFlags(&ix->done[i], cudaEventDisableTiming));
        }
    }
    if (ix->stage_bytes < bytes) {
        for (int i = 0; i < 2; i++) {
            if (ix->stage[i]) GKI_CUDA(cudaFree(ix->stage[i]));
            ix->stage[i] = nullptr;
            GKI_CUDA(cudaMalloc(&ix->stage[i], bytes));
        }
        ix->stage_bytes = bytes;
    }
    return GKI_OK;
}

static int launch_count_kmers(gki_index *ix, const uint64_t *dq, int64_t nq, cudaStream_t s) {
    if (nq <= 0) return GKI_OK;
    int grid = grid_for(nq, COUNT_THREADS * 4, device_info().sms * 8);
    count_kmers_kernel<<<grid, COUNT_THREADS, 0, s>>>(ix->table, dq, nq);
    GKI_CHECK_LAUNCH();
    return GKI_OK;
}

template <bool BOTH, bool PAIRED> static int launch_count_reads_t(gki_index *ix, const ReadBatch &b, size_t smem, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        GKI_CUDA(cudaFuncSetAttribute(count_reads_kernel<BOTH, PAIRED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set = true;
    }
    int blocks_per_sm = 0;
    GKI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, count_reads_kernel<BOTH, PAIRED>, COUNT_THREADS, smem));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    int grid = grid_for(b.n_tiles, 1, device_info().sms * blocks_per_sm);
    count_reads_kernel<BOTH, PAIRED><<<grid, COUNT_THREADS, smem, s>>>(ix->table, b);
    GKI_CHECK_LAUNCH();
    return GKI_OK;
}

// reads: device rows
static int launch_count_reads(gki_index *ix, const uint8_t *dreads, int64_t n_reads, int32_t read_len, int64_t stride, int32_t k,
                              int32_t both, cudaStream_t s) {
    ReadBatch b;
    size_t smem;
    make_read_batch(dreads, n_reads, read_len, stride, k, b, smem);
    GKI_REQUIRE(smem <= 64 * 1024, GKI_ERR_UNSUPPORTED, "gki_count_reads: read_len %d too long for the tile path", read_len);
    if (!both) return launch_count_reads_t<false, false>(ix, b, smem, s);
    if (ix->table.k == k) return launch_count_reads_t<true, true>(ix, b, smem, s);
    return launch_count_reads_t<true, false>(ix, b, smem, s);
}

}  // namespace gki

using namespace gki;

extern "C" {

int gki_prepare_counting(gki_index_t *ix, int32_t k, gki_stream_t stream) {
    GKI_REQUIRE(ix && k >= 0 && k <= 31, GKI_ERR_INVALID, "gki_prepare_counting: bad arguments");
    return ensure_table(ix, k, (cudaStream_t)stream);
}

int gki_reset_counts(gki_index_t *ix, gki_stream_t stream) {
    GKI_REQUIRE(ix, GKI_ERR_INVALID, "gki_reset_counts: index is NULL");
    if (!ix->table.slots) return GKI_OK;   // nothing counted yet
    size_t n_slots = (size_t)ix->table.n_buckets * SLOTS_PER_BUCKET + 1;
    table_reset_kernel<<<grid_for((int64_t)n_slots, 256 * 4, device_info().sms * 16), 256, 0, (cudaStream_t)stream>>>(ix->table.slots, n_slots);
    GKI_CHECK_LAUNCH();
    return GKI_OK;
}

int gki_count_kmers(gki_index_t *ix, const uint64_t *queries, int64_t nq, gki_stream_t stream) {
    GKI_REQUIRE(ix && nq >= 0 && (nq == 0 || queries), GKI_ERR_INVALID, "gki_count_kmers: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (nq == 0) return GKI_OK;
    GKI_TRY(ensure_table(ix, 0, s));
    if (is_device_ptr(queries)) return launch_count_kmers(ix, queries, nq, s);
    // host queries: chunked, double-buffered H2D overlapped with the probe kernel
    const int64_t chunk = 4 << 20;   // 4 Mi queries = 32 MiB
    GKI_TRY(ensure_staging(ix, (size_t)chunk * 8));
    int c = 0;
    for (int64_t off = 0; off < nq; off += chunk, ++c) {
        int bsel = c & 1;
        int64_t cnt = nq - off < chunk ? nq - off : chunk;
        if (c >= 2) GKI_CUDA(cudaStreamWaitEvent(ix->copy_stream, ix->done[bsel], 0));
        GKI_CUDA(cudaMemcpyAsync(ix->stage[bsel], queries + off, (size_t)cnt * 8, cudaMemcpyHostToDevice, ix->copy_stream));
        GKI_CUDA(cudaEventRecord(ix->ready[bsel], ix->copy_stream));
        GKI_CUDA(cudaStreamWaitEvent(s, ix->ready[bsel], 0));
        GKI_TRY(launch_count_kmers(ix, (const uint64_t *)ix->stage[bsel], cnt, s));
        GKI_CUDA(cudaEventRecord(ix->done[bsel], s));
    }
    GKI_CUDA(cudaStreamSynchronize(s));
    return GKI_OK;
}

int gki_count_reads(gki_index_t *ix, const uint8_t *reads, int64_t n_reads, int32_t read_len, int64_t row_stride, int32_t k,
                    int32_t both_strands, gki_stream_t stream) {
    GKI_REQUIRE(ix && n_reads >= 0 && read_len >= 0 && row_stride >= read_len, GKI_ERR_INVALID, "gki_count_reads: bad arguments");
    GKI_REQUIRE(k >= 1 && k <= 31, GKI_ERR_INVALID, "gki_count_reads: k must be in [1, 31], got %d", k);
    cudaStream_t s = (cudaStream_t)stream;
    if (n_reads == 0 || read_len < k) return GKI_OK;
    GKI_REQUIRE(reads, GKI_ERR_INVALID, "gki_count_reads: reads is NULL");
    GKI_TRY(ensure_table(ix, k, s));
    if (is_device_ptr(reads)) return launch_count_reads(ix, reads, n_reads, read_len, row_stride, k, both_strands, s);
    // host reads: rows are compacted to dense device rows (so every full tile is one TMA bulk copy) in chunks;
    // the copy of chunk c+1 overlaps the count kernel of chunk c
    int64_t chunk_reads = ((32ll << 20) / (read_len > 0 ? read_len : 1)) & ~31ll;
    if (chunk_reads < 32) chunk_reads = 32;
    GKI_TRY(ensure_staging(ix, (size_t)chunk_reads * read_len + 16));
    int c = 0;
    for (int64_t off = 0; off < n_reads; off += chunk_reads, ++c) {
        int bsel = c & 1;
        int64_t cnt = n_reads - off < chunk_reads ? n_reads - off : chunk_reads;
        if (c >= 2) GKI_CUDA(cudaStreamWaitEvent(ix->copy_stream, ix->done[bsel], 0));
        if (row_stride == read_len)
            GKI_CUDA(cudaMemcpyAsync(ix->stage[bsel], reads + off * row_stride, (size_t)cnt * read_len, cudaMemcpyHostToDevice, ix->copy_stream));
        else
            GKI_CUDA(cudaMemcpy2DAsync(ix->stage[bsel], read_len, reads + off * row_stride, row_stride, read_len, cnt, cudaMemcpyHostToDevice, ix->copy_stream));
        GKI_CUDA(cudaEventRecord(ix->ready[bsel], ix->copy_stream));
        GKI_CUDA(cudaStreamWaitEvent(s, ix->ready[bsel], 0));
        GKI_TRY(launch_count_reads(ix, (const uint8_t *)ix->stage[bsel], cnt, read_len, read_len, k, both_strands, s));
        GKI_CUDA(cudaEventRecord(ix->done[bsel], s));
    }
    GKI_CUDA(cudaStreamSynchronize(s));
    return GKI_OK;
}

int gki_node_counts(gki_index_t *ix, double *out, int64_t n_out, int32_t flags, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(ix && out && n_out >= 0, GKI_ERR_INVALID, "gki_node_counts: bad arguments");
    GKI_REQUIRE(n_out > ix->max_node, GKI_ERR_OVERFLOW, "gki_node_counts: n_out %lld <= max node id %lld", (long long)n_out, (long long)ix->max_node);
    DevOut o;
    GKI_TRY(o.prepare(out, (size_t)n_out * 8, call.stream));
    GKI_CUDA(cudaMemsetAsync(o.dptr, 0, (size_t)n_out * 8, call.stream));
    if (ix->table.slots) {
        node_counts_kernel<<<grid_for(ix->n, 256 * 4, device_info().sms * 16), 256, 0, call.stream>>>(
            ix->table, ix->kmers, ix->nodes, ix->n, o.as<double>(), n_out, (flags & GKI_COUNTS_WRAP_UINT16) != 0);
        GKI_CHECK_LAUNCH();
    }
    GKI_TRY(o.finish(call.stream));
    return call.finish();
}

int gki_entry_counts(gki_index_t *ix, uint32_t *out, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(ix && out, GKI_ERR_INVALID, "gki_entry_counts: bad arguments");
    DevOut o;
    GKI_TRY(o.prepare(out, (size_t)ix->n * 4, call.stream));
    if (ix->table.slots) {
        query_counts_kernel<<<grid_for(ix->n, 256 * 4, device_info().sms * 16), 256, 0, call.stream>>>(ix->table, ix->kmers, ix->n, o.as<uint32_t>());
        GKI_CHECK_LAUNCH();
    } else {
        GKI_CUDA(cudaMemsetAsync(o.dptr, 0, (size_t)ix->n * 4, call.stream));
    }
    GKI_TRY(o.finish(call.stream));
    return call.finish();
}

int gki_query_counts(gki_index_t *ix, const uint64_t *queries, int64_t nq, uint32_t *out, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(ix && nq >= 0 && (nq == 0 || (queries && out)), GKI_ERR_INVALID, "gki_query_counts: bad arguments");
    if (nq == 0) return GKI_OK;
    DevIn q;
    GKI_TRY(q.stage(queries, (size_t)nq * 8, call.stream));
    DevOut o;
    GKI_TRY(o.prepare(out, (size_t)nq * 4, call.stream));
    if (ix->table.slots) {
        query_counts_kernel<<<grid_for(nq, 256 * 2, device_info().sms * 16), 256, 0, call.stream>>>(ix->table, q.as<uint64_t>(), nq, o.as<uint32_t>());
        GKI_CHECK_LAUNCH();
    } else {
        GKI_CUDA(cudaMemsetAsync(o.dptr, 0, (size_t)nq * 4, call.stream));
    }
    GKI_TRY(o.finish(call.stream));
    return call.finish();
}

}  // extern "C"
