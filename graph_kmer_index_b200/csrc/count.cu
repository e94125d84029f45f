// count.cu -- K3 counting: CounterKmerIndex.count_kmers / get_node_counts (collision_free_kmer_index.py:14-40)
// and the fused read path (read_kmers.py:14-26 -> cfki:33-37).
//
// The reference keeps one counter per distinct index k-mer (npstructures.Counter keyed by k-mer, cfki:27) and
// get_node_counts = bincount(nodes, weights=counter[kmers]) (cfki:39-40).  The counts only depend on k-mer
// equality, so the device is free to key the counters however it likes.  Measured on B200 (profiles/r1):
// random 8-byte gathers run at ~290 G/s while the table fits L2 (L1TEX-bound, one sector request per clock per SM)
// and at ~37 G/s from HBM, each HBM access filling a whole 128-byte line; the first version of this kernel, which
// probed the reference's modulo-bucket tables behind a 56 MB bucket bitmap, was bound by exactly those rates
// (132 GB of DRAM reads per 2.4 G queries).  Hence this layout:
//
//   * Bloom filter, register-blocked (one 32-bit word per key, filter_k bits), 32 MB so it stays in L2: one L2
//     access decides most absent k-mers.  Filters of indexes too large for that live in HBM and are addressed by
//     the k-mer's minimizer, so that consecutive windows of a read share a line (filter_m, kmer_minimizer).
//   * bucketised open-addressing table over the distinct k-mers: bucket = key[4] | cnt[4][2] = 64 bytes, two buckets
//     per 128-byte line (a full bucket overflows into its line mate first); a probe reads the four keys with ONE
//     256-bit load, a hit adds one RED on the other half of the bucket.
//   * canonical keys: key = min(x, revcomp_k(x)), one count per orientation o = (x != key).  The forward and reverse-complement
//     hashes of a read position share the key, so a position (2 queries) costs one filter access, at most one
//     table access and one 32-bit RED (cnt[0] counts orientation 0, cnt[1] holds the DIFFERENCE of the two orientations,
//     so "+1 on both" touches cnt[0] alone and no carry can run between the two counts; see add_both_orientations).
//   * multiply-fold hash + multiply-high range reduction: no 64-bit modulo in the hot loop.
//
// The fused kernel is warp-autonomous: every warp streams its own tiles of reads into shared memory with TMA
// bulk copies on its own mbarrier, packs them to 2 bits per base (or receives them packed) and walks them with a
// rolling window -- no CTA-wide barrier anywhere, so a warp that waits on HBM never stalls its neighbours.
// Host batches reach it through the packing lanes at the end of this file (count_reads_host_pipeline).
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "index.cuh"
#include "ingest.h"
#include "reads_tile.cuh"

namespace gki {

constexpr int COUNT_THREADS = 256;
constexpr int COUNT_WARPS = COUNT_THREADS / 32;

// ------------------------------------------------------------------ hashing / keys
struct Hash {
    uint32_t hi;   // well-mixed: home bucket
    uint32_t f;    // folded: Bloom word + bit positions
};
__device__ __forceinline__ Hash hash_key(unsigned long long key) {
    unsigned long long a = key * 0x9E3779B97F4A7C15ull;
    Hash h;
    h.hi = (uint32_t)(a >> 32);
    h.f = h.hi ^ (uint32_t)a;
    return h;
}
__device__ __forceinline__ uint32_t home_bucket(const TableView &t, const Hash &h) { return __umulhi(h.hi, t.n_buckets); }
__device__ __forceinline__ uint32_t filter_word(const TableView &t, const Hash &h) {
    uint32_t g = h.f * 0x85EBCA6Bu;
    g ^= g >> 15;
    return __umulhi(g, t.filter_words);
}
// FK: bits per key when known at compile time (the fused kernel is instantiated for 3), 0: read t.filter_k
template <int FK = 0>
__device__ __forceinline__ uint32_t filter_mask(const TableView &t, const Hash &h) {
    uint32_t b = h.hi * 0xC2B2AE35u;     // bit positions from the high half, word from the fold: 64 bits of the product are used
    uint32_t mask = 1u << (b >> 27);
    if (FK ? FK > 1 : t.filter_k > 1) mask |= 1u << ((b >> 22) & 31);
    if (FK ? FK > 2 : t.filter_k > 2) mask |= 1u << ((b >> 17) & 31);
    return mask;
}

// ---- minimizer-addressed filter words ----
// A filter that does not fit L2 costs one HBM line per probe when the word is chosen by the key's hash (C3: 4.5 G lines per
// step).  Choosing the 32-byte sector by the k-mer's minimizer instead -- the canonical m-mer, m = k - 16, with the smallest
// hash among the 17 the k-mer holds -- makes the ~9 consecutive windows of a read that share a minimizer probe the same
// sector; the word inside the sector and the bits still come from the key's hash.  Strand-symmetric: a k-mer and its
// reverse complement hold the same canonical m-mers.
constexpr int MZ_W = 17;
__device__ __forceinline__ uint32_t revcomp_m(uint32_t x, int m) {   // m <= 15 bases in 32 bits
    uint32_t y = __brev(~x);
    y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
    return y >> (2 * (16 - m));
}
__device__ __forceinline__ uint32_t mmer_hash(uint32_t fwd, uint32_t rc) {
    uint32_t c = (fwd < rc ? fwd : rc) * 0x9E3779B1u;
    return c ^ (c >> 16);
}
__device__ __forceinline__ uint32_t kmer_minimizer(unsigned long long x, int m) {   // smallest mmer_hash over the k-mer's 17 m-mers
    const uint32_t mmask = (1u << (2 * m)) - 1u;
    uint32_t fwd = (uint32_t)x & mmask, rc = revcomp_m(fwd, m);
    uint32_t best = mmer_hash(fwd, rc);
    x >>= 2 * m;
#pragma unroll 4
    for (int j = 1; j < MZ_W; j++) {
        const uint32_t nb = (uint32_t)x & 3u;
        x >>= 2;
        fwd = (fwd >> 2) | (nb << (2 * (m - 1)));
        rc = ((rc << 2) | (3u - nb)) & mmask;
        best = min(best, mmer_hash(fwd, rc));
    }
    return best;
}
__device__ __forceinline__ uint32_t filter_word_mz(const TableView &t, const Hash &h, uint32_t minimizer) {
    uint32_t s = minimizer * 0x85EBCA6Bu;   // the minimum of 17 hashes is far from uniform: mix again before the range reduction
    s ^= s >> 13;
    s *= 0xC2B2AE35u;
    s ^= s >> 16;
    return __umulhi(s, t.filter_words >> 3) * 8u + (h.f >> 29);
}
__device__ __forceinline__ uint32_t filter_word_of(const TableView &t, const Hash &h, unsigned long long key) {
    return t.filter_m ? filter_word_mz(t, h, kmer_minimizer(key, t.filter_m)) : filter_word(t, h);
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

// L2 eviction-priority hints (createpolicy + ld.global.L2::cache_hint): the Bloom filter is the only structure with
// reuse, table lines are touched once per probe.  HINTS bit 0: filter loads evict_last, bit 1: table loads evict_first.
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint32_t ld_u32_hint(const uint32_t *ptr, unsigned long long pol) {
    uint32_t v;
    asm("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ unsigned long long ld_u64_hint(const unsigned long long *ptr, unsigned long long pol) {
    unsigned long long v;
    asm("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(ptr), "l"(pol));
    return v;
}

// The four keys of a bucket in ONE 256-bit load (LDG.E.256 on sm_100a).  ncu on the first table layout showed four
// separate 8-byte loads of one line, issued back to back, each paying its own 32-byte HBM fetch (L2 does not merge
// in-flight misses of different requests): 128 B of DRAM reads per probe instead of 32.
__device__ __forceinline__ void load_bucket_keys(const Bucket *bk, unsigned long long (&key)[SLOTS_PER_BUCKET]) {
    asm("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(key[0]), "=l"(key[1]), "=l"(key[2]), "=l"(key[3]) : "l"(bk));
}
__device__ __forceinline__ void load_bucket_keys_64B(const Bucket *bk, unsigned long long (&key)[SLOTS_PER_BUCKET]) {   // experiment: 64-byte L2 fill
    asm("ld.global.nc.L2::64B.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(key[0]), "=l"(key[1]), "=l"(key[2]), "=l"(key[3]) : "l"(bk));
}
__device__ __forceinline__ void load_bucket_keys_hint(const Bucket *bk, unsigned long long (&key)[SLOTS_PER_BUCKET], unsigned long long pol) {
    asm("ld.global.nc.L2::cache_hint.v4.u64 {%0, %1, %2, %3}, [%4], %5;" : "=l"(key[0]), "=l"(key[1]), "=l"(key[2]), "=l"(key[3]) : "l"(bk), "l"(pol));
}

// Probe order: HBM fills L2 in whole 128-byte lines on this part (measured: 127 B of DRAM reads per random 8-byte gather,
// profiles/r1/calibrate_gather_ncu.txt), so a line holds two 64-byte buckets and the second one is free once the first
// was fetched: home bucket, then its line mate, then the next line starting with the same half.  n_buckets is even.
__device__ __forceinline__ uint32_t next_bucket(const TableView &t, uint32_t b, uint32_t tries) {
    if (!(tries & 1u)) return b ^ 1u;
    b = (b ^ 1u) + 2u;
    return b >= t.n_buckets ? b - t.n_buckets : b;
}

// counter pair {cnt[0], cnt[1]} of key c (c != SLOT_EMPTY) starting at bucket b, or nullptr.  Probing follows
// next_bucket and stops at the first non-full bucket.
template <int HINT = 0>
__device__ __forceinline__ uint32_t *find_slot_from(const TableView &t, unsigned long long c, uint32_t b, unsigned long long pol = 0,
                                                    uint32_t tries0 = 0) {
    for (uint32_t tries = tries0; tries < t.n_buckets; tries++) {
        Bucket *bk = t.buckets + b;
        unsigned long long key[SLOTS_PER_BUCKET];   // keys never change while a counting kernel runs
        if (HINT == 1) load_bucket_keys_hint(bk, key, pol);
        else if (HINT == 2) load_bucket_keys_64B(bk, key);
        else load_bucket_keys(bk, key);
#pragma unroll
        for (int i = 0; i < SLOTS_PER_BUCKET; i++) {
            if (key[i] == c) return bk->cnt[i];
            if (key[i] == SLOT_EMPTY) return nullptr;
        }
        b = next_bucket(t, b, tries);
    }
    return nullptr;
}

// counter pair of key c, or nullptr
__device__ __forceinline__ uint32_t *find_slot(const TableView &t, unsigned long long c, const Hash &h) {
    if (c == SLOT_EMPTY) {   // raw mode only: the one value that collides with the empty marker has its own bucket,
        Bucket *sp = t.buckets + t.n_buckets;   // whose key[0] is 1 iff that value is indexed
        return __ldg(&sp->key[0]) == 1ull ? sp->cnt[0] : nullptr;
    }
    return find_slot_from(t, c, home_bucket(t, h));
}
// slot id <-> counter pair (get_node_counts regroups the entries by slot id)
__device__ __forceinline__ unsigned long long slot_id(const TableView &t, const uint32_t *cnt) {
    const size_t off = (size_t)(cnt - (const uint32_t *)t.buckets);       // in u32 units: bucket * 16 + 8 + 2 * i
    return (unsigned long long)(off >> 4) * SLOTS_PER_BUCKET + (((off & 15) - 8) >> 1);
}
__device__ __forceinline__ uint32_t *slot_counters(const TableView &t, unsigned long long slot) {
    return t.buckets[slot >> 2].cnt[slot & 3];
}
// Counter pair of a slot: cnt[0] = count of orientation 0, cnt[1] = count of orientation 1 MINUS count of orientation 0, both
// modulo 2^32.  A read position counted on both strands adds one to each orientation, i.e. +1 on cnt[0] alone: one 32-bit RED,
// and no carry can run from one orientation's count into the other's (a single 64-bit add of {1, 1} on two plain counters did
// exactly that after 2^32 hits).  Every orientation's count is exact modulo 2^32 (the reference's Counter wraps at 2^16, cfki:27).
__device__ __forceinline__ void add_both_orientations(uint32_t *cnt, uint32_t n) { atomicAdd(cnt, n); }
__device__ __forceinline__ void add_one_orientation(const TableView &t, uint32_t *cnt, uint32_t o) {
    if (o) {
        atomicAdd(cnt + 1, 1u);
    } else {
        atomicAdd(cnt, 1u);
        if (t.k) atomicAdd(cnt + 1, 0xffffffffu);   // raw keys (k == 0) have orientation 0 only: the difference is never read
    }
}
__device__ __forceinline__ uint32_t read_orientation(const uint32_t *cnt, uint32_t o) {
    const uint32_t a = *(const volatile uint32_t *)cnt;
    return o ? a + *(const volatile uint32_t *)(cnt + 1) : a;
}

// one independent query
__device__ __forceinline__ void count_one(const TableView &t, uint64_t q) {
    Key key = make_key(q, t.k);
    if (!key.ok) return;
    Hash h = hash_key(key.c);
    if (t.filter && key.c != SLOT_EMPTY) {   // the value that collides with the empty marker has its own bucket and no filter bits
        uint32_t m = filter_mask(t, h);
        if ((__ldg(t.filter + filter_word_of(t, h, key.c)) & m) != m) return;
    }
    uint32_t *cnt = find_slot(t, key.c, h);
    if (cnt) add_one_orientation(t, cnt, key.o);
}

// ------------------------------------------------------------------ table construction
__global__ void table_init_kernel(Bucket *__restrict__ buckets, size_t n_buckets) {
    // one thread per 8-byte word: words 0-3 of a bucket are keys (empty), words 4-7 counters (zero)
    unsigned long long *w = (unsigned long long *)buckets;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_buckets * 8; i += (size_t)gridDim.x * blockDim.x)
        w[i] = (i & 4) ? 0ull : SLOT_EMPTY;
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

__global__ void table_insert_kernel(TableView t, const uint64_t *__restrict__ kmers, int64_t n, uint32_t *__restrict__ filter,
                                    unsigned int *__restrict__ failed, FastMod fm) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t km = __ldg(kmers + e);
        // the entries are in bucket order: a k-mer that an earlier entry of the same bucket already holds (variant k-mers come with
        // two or more nodes) has nothing to insert -- half of the random table and filter accesses at c3.  The look back is bounded;
        // an entry it cannot decide about is simply inserted again.
        {
            const uint32_t bucket = fastmod(km, fm);
            bool seen = false;
            for (int64_t c = e - 1; c >= 0 && c >= e - 16; c--) {
                const uint64_t other = __ldg(kmers + c);
                if (other == km) {
                    seen = true;
                    break;
                }
                if (fastmod(other, fm) != bucket) break;
            }
            if (seen) continue;
        }
        Key key = make_key(km, t.k);
        if (key.c == SLOT_EMPTY) {
            t.buckets[t.n_buckets].key[0] = 1ull;   // mark the special bucket as present
            continue;
        }
        Hash h = hash_key(key.c);
        if (filter) atomicOr(filter + filter_word_of(t, h, key.c), filter_mask(t, h));
        uint32_t b = home_bucket(t, h);
        bool placed = false;
        for (uint32_t tries = 0; tries < t.n_buckets && !placed; tries++) {
            Bucket *bk = t.buckets + b;
            for (int i = 0; i < SLOTS_PER_BUCKET && !placed; i++) {
                unsigned long long cur = *(volatile unsigned long long *)&bk->key[i];
                if (cur == SLOT_EMPTY) cur = atomicCAS(&bk->key[i], SLOT_EMPTY, key.c);
                placed = (cur == SLOT_EMPTY) || (cur == key.c);
            }
            b = next_bucket(t, b, tries);
        }
        if (!placed) atomicAdd(failed, 1u);
    }
}

__global__ void table_reset_kernel(Bucket *__restrict__ buckets, size_t n_buckets) {
    // the counter half (32 B) of every bucket, 16 bytes per thread
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_buckets * 2; i += (size_t)gridDim.x * blockDim.x)
        *(uint4 *)((char *)&buckets[i >> 1].cnt[0][0] + 16 * (i & 1)) = make_uint4(0u, 0u, 0u, 0u);
}

// ------------------------------------------------------------------ counting kernels
// CounterKmerIndex.count_kmers (cfki:33-37) on an array of hashes.  Same scheme as the fused kernel: a warp takes 4 x 32
// queries at a time (coalesced loads, four filter words in flight per lane), the few that pass the filter are compacted into a
// per-warp queue by ballot and the table is probed 32 survivors at a time -- a thread per query with the probe inline leaves
// most of a warp waiting on HBM for the one lane in six that has a survivor (94 G queries/s; this form: see DESIGN.md).
constexpr int CK_U = 4;       // queries per lane and round
constexpr int CK_QCAP = 64;   // survivor queue capacity per warp
__global__ void __launch_bounds__(COUNT_THREADS) count_kmers_kernel(TableView t, const uint64_t *__restrict__ queries, int64_t nq) {
    __shared__ unsigned long long s_key[COUNT_WARPS][CK_QCAP];
    __shared__ uint32_t s_meta[COUNT_WARPS][CK_QCAP];   // home bucket | orientation << 31
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long *qkey = s_key[warp];
    uint32_t *qmeta = s_meta[warp];
    uint32_t qn = 0, qhead = 0;
    auto probe_queue = [&](uint32_t n) {
        if ((uint32_t)lane < n) {
            const uint32_t idx = (qhead + lane) & (CK_QCAP - 1);
            const uint32_t meta = qmeta[idx];
            uint32_t *cnt = find_slot_from(t, qkey[idx], meta & 0x7fffffffu);
            if (cnt) add_one_orientation(t, cnt, meta >> 31);
        }
        qhead += n;
        __syncwarp();
    };
    const int64_t round = 32 * CK_U;
    const int64_t n_rounds = (nq + round - 1) / round;
    const uint32_t below = (1u << lane) - 1u;
    for (int64_t r = (int64_t)blockIdx.x * COUNT_WARPS + warp; r < n_rounds; r += (int64_t)gridDim.x * COUNT_WARPS) {
        Key key[CK_U];
        uint32_t home[CK_U], fw[CK_U], fm[CK_U];
        uint32_t live = 0;
#pragma unroll
        for (int u = 0; u < CK_U; u++) {
            const int64_t i = r * round + u * 32 + lane;
            key[u] = make_key(i < nq ? __ldg(queries + i) : 0ull, t.k);
            bool ok = i < nq && key[u].ok;
            if (ok && key[u].c == SLOT_EMPTY) {   // raw mode only: the one value that collides with the empty marker
                count_one(t, key[u].c);
                ok = false;
            }
            const Hash h = hash_key(key[u].c);
            home[u] = home_bucket(t, h);
            fm[u] = filter_mask(t, h);
            fw[u] = (t.filter && ok) ? __ldg(t.filter + filter_word_of(t, h, key[u].c)) : 0xffffffffu;
            live |= (uint32_t)ok << u;
        }
#pragma unroll
        for (int u = 0; u < CK_U; u++) live &= ~((uint32_t)((fw[u] & fm[u]) != fm[u]) << u);
#pragma unroll
        for (int u = 0; u < CK_U; u++) {
            const bool mine = (live >> u) & 1u;
            const uint32_t votes = __ballot_sync(0xffffffffu, mine);
            if (mine) {
                const uint32_t slot_idx = (qn + __popc(votes & below)) & (CK_QCAP - 1);
                qkey[slot_idx] = key[u].c;
                qmeta[slot_idx] = home[u] | (key[u].o << 31);
            }
            qn += __popc(votes);
            if (qn - qhead >= 32) {
                __syncwarp();
                probe_queue(32);
            }
        }
    }
    __syncwarp();
    if (qn != qhead) probe_queue(qn - qhead);
}

// ---- fused K1 -> K3, warp-autonomous ----
struct WarpBatch {
    const uint8_t *reads;
    int64_t n_reads;
    int64_t row_stride;
    int64_t n_wtiles;       // ceil(n_reads / rpw)
    int32_t read_len;
    int32_t k;
    int32_t nk;             // read_len - k + 1
    int32_t words;          // ceil(read_len / 32) + 1
    int32_t rpw;            // reads per warp tile
    int32_t bulk_ok;        // dense + aligned: full tiles are one TMA bulk copy
    uint32_t mz_off;        // minimizer filters: byte offset (16-aligned) of the per-warp m-mer hash array, mz_len entries
    uint32_t mz_len;
    uint32_t inv_halves;    // ceil(65536 / (2 * words)): task / (2 * words) == task * inv_halves >> 16 for every task of a tile
    uint32_t stage_bytes;   // bytes of one ASCII stage (16-byte multiple, incl. slack)
    uint32_t warp_bytes;    // shared memory per warp
};

constexpr int WPL = 4;   // consecutive windows per lane
// survivor queue capacity per warp (< 32 waiting + <= 32 pushed per ballot).  Kept as small as the scheme allows: shared memory
// comes out of the L1 that tracks the outstanding gathers, and a 4x larger queue measurably slowed the launch (9.2 -> 9.5 ms)
constexpr int QCAP = 64;

// PACKED: the batch is already 2 bits per base (gki_pack_reads layout: read r owns b.words 64-bit words, every byte was
// one of ACGTacgt) -- the tile lands straight in the code words, double-buffered, and the pack phase disappears.
// KODD: k is odd, no k-mer equals its own reverse complement (no palindrome bookkeeping).  FK: filter bits per key (0: runtime).
// MINZ: the filter is minimizer-addressed (t.filter_m): per read the hashes of all canonical m-mers go to shared memory
// and every window takes the minimum of its 17.
template <bool BOTH, bool PAIRED, int MINB, int HINTS = 0, bool PACKED = false, bool KODD = false, int FK = 0, bool MINZ = false>
__global__ void __launch_bounds__(COUNT_THREADS, MINB) count_reads_kernel(TableView t, WarpBatch b) {
    const unsigned long long pol_last = (HINTS & 1) ? l2_policy_evict_last() : 0ull;
    const unsigned long long pol_first = (HINTS & 2) ? l2_policy_evict_first() : 0ull;
    const unsigned long long pol_stream = (HINTS & 4) ? l2_policy_evict_first() : 0ull;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per-warp shared memory: [mbarrier | ASCII stage | codes | valid | dirty flags].  One ASCII stage is enough: it is
    // dead once packed, so the next tile's bulk copy is issued right after packing and lands while the warp walks.
    unsigned char *wbase = smem_raw + (size_t)warp * b.warp_bytes;
    uint64_t *bar = (uint64_t *)wbase;
    uint8_t *ascii = wbase + 16;
    uint64_t *codes = (uint64_t *)(wbase + 16 + (size_t)b.stage_bytes);
    uint64_t *valid = codes + (size_t)b.rpw * b.words;
    if (PACKED) {   // [2 mbarriers | code stage 0 | code stage 1 | queue | -]
        codes = (uint64_t *)(wbase + 16);
        valid = (uint64_t *)(wbase + 16 + (size_t)b.stage_bytes);
    }
    unsigned long long *qkey = PACKED ? (unsigned long long *)(wbase + 16 + 2 * (size_t)b.stage_bytes)
                                      : (unsigned long long *)(valid + (size_t)b.rpw * b.words);   // survivor queue: keys
    uint32_t *qmeta = (uint32_t *)(qkey + QCAP);                                           //   home bucket | palindrome << 31
    uint32_t *dirty = qmeta + QCAP;
    uint32_t *mzh = (uint32_t *)(wbase + b.mz_off);
    const uint64_t mask = kmer_mask(b.k);
    const uint32_t tile_bytes = PACKED ? (uint32_t)b.rpw * (uint32_t)b.words * 8u : (uint32_t)b.rpw * (uint32_t)b.read_len;
    uint32_t qn = 0, qhead = 0;   // warp-uniform: entries pushed / consumed so far
    auto probe_queue = [&](uint32_t n) {   // the first n queued survivors, one per lane: bucket line from HBM, compare, RED
        if (!(HINTS & 16) && (uint32_t)lane < n) {   // HINTS bit 4: experiment only, survivors are dropped
            const uint32_t idx = (qhead + lane) & (QCAP - 1);
            const unsigned long long key = qkey[idx];
            const uint32_t meta = qmeta[idx];
            uint32_t *cnt = find_slot_from<(HINTS & 2) ? 1 : ((HINTS & 64) ? 2 : 0)>(t, key, meta & 0x7fffffffu, pol_first);
            if (cnt) add_both_orientations(cnt, (!KODD && (meta >> 31)) ? 2u : 1u);   // palindrome (even k): both strands are orientation 0
        }
        qhead += n;
        __syncwarp();
    };

    if (lane == 0) {
        mbar_init(bar, 1);
        if (PACKED) mbar_init(bar + 1, 1);
        fence_mbar_init();
    }
    __syncwarp();
    // A CTA owns a contiguous range of tiles and its warps take them one by one from a shared counter: a warp that drew costly reads
    // (reads from the indexed sequence probe the table at every window) takes fewer tiles, and the CTA's slots are not held by one
    // slow warp while seven have finished.
    __shared__ uint32_t cta_taken;
    const int64_t per_cta = (b.n_wtiles + gridDim.x - 1) / gridDim.x;
    const int64_t c0 = (int64_t)blockIdx.x * per_cta, c1 = min(c0 + per_cta, b.n_wtiles);
    if (threadIdx.x == 0) cta_taken = COUNT_WARPS;       // the first tile of every warp is its own
    __syncthreads();
    auto take = [&]() -> int64_t {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(&cta_taken, 1u);
        return c0 + (int64_t)__shfl_sync(0xffffffffu, t, 0);
    };
    int64_t wt = c0 + warp;
    auto uses_bulk = [&](int64_t tile) { return b.bulk_ok && (tile + 1) * (int64_t)b.rpw <= b.n_reads; };
    auto issue = [&](int64_t tile, int stage) {   // stage: PACKED only (code stage and its mbarrier)
        void *dst = PACKED ? (void *)(stage ? valid : codes) : (void *)ascii;
        uint64_t *mb = PACKED ? bar + stage : bar;
        fence_proxy_async();
        mbar_expect_tx(mb, tile_bytes);
        if (HINTS & 4) bulk_g2s_hint(dst, b.reads + tile * (int64_t)b.rpw * b.row_stride, tile_bytes, mb, pol_stream);
        else bulk_g2s(dst, b.reads + tile * (int64_t)b.rpw * b.row_stride, tile_bytes, mb);
    };
    if (wt < c1 && lane == 0 && uses_bulk(wt)) issue(wt, 0);
    uint32_t phase = 0;   // PACKED: bit s is the parity of stage s

    for (int it = 0; wt < c1; ++it) {
        int64_t next = c1;    // taken from the counter when its copy is issued
        const int64_t r0 = wt * (int64_t)b.rpw;
        const int n_here = (int)min((int64_t)b.rpw, b.n_reads - r0);
        const uint64_t *tile_codes = codes;
        if (PACKED) {
            const int st = it & 1;
            tile_codes = st ? valid : codes;
            __syncwarp();
            next = take();
            if (next < c1 && lane == 0 && uses_bulk(next)) issue(next, st ^ 1);   // the other stage was walked last round
            if (uses_bulk(wt)) {
                mbar_wait(bar + st, (phase >> st) & 1u);
                phase ^= 1u << st;
            } else {   // unaligned batch or tail tile: the warp copies its words itself
                const unsigned long long *src = (const unsigned long long *)(b.reads + r0 * b.row_stride);
                for (int i = lane; i < n_here * b.words; i += 32) ((uint64_t *)tile_codes)[i] = __ldg(src + i);
                __syncwarp();
            }
        } else if (uses_bulk(wt)) {
            mbar_wait(bar, phase);
            phase ^= 1;
        } else {   // strided / unaligned / tail tile: the warp copies its rows itself
            for (int r = 0; r < n_here; r++) {
                const uint8_t *src = b.reads + (r0 + r) * b.row_stride;
                for (int i = lane; i < b.read_len; i += 32) ascii[(size_t)r * b.read_len + i] = __ldg(src + i);
            }
            __syncwarp();
        }
        // ---- pack to 2 bits per base (+ validity), flag reads that contain a non-ACGT byte ----
        if (!PACKED && lane < n_here) dirty[lane] = 0;
        __syncwarp();
        // one task = 16 bases = one 32-bit half of a code word (the pad word included): 8 reads x 12 halves fill the warp
        // three times over, where whole-word tasks left half of it idle in their second round
        const int halves = 2 * b.words;
        for (int task = lane; !PACKED && task < n_here * halves; task += 32) {
            const int r = (int)(((uint32_t)task * b.inv_halves) >> 16), h = task - r * halves;   // task / halves (exact, checked on the host)
            const int first = h * 16;
            const int nb = min(16, b.read_len - first);
            uint32_t c32 = 0, v32 = 0;
            if (nb > 0) {
                const uint32_t addr = (uint32_t)r * (uint32_t)b.read_len + (uint32_t)first;
                const uint32_t *aligned = (const uint32_t *)(ascii + (addr & ~3u));
                const uint32_t sh = (addr & 3u) * 8u;
                uint32_t lo = aligned[0];
                const int nq = (nb + 3) >> 2;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (q < nq) {
                        const uint32_t hi = aligned[q + 1];
                        uint32_t c8, v8;
                        encode4(__funnelshift_r(lo, hi, sh), c8, v8);
                        lo = hi;
                        c32 |= c8 << (8 * q);
                        v32 |= v8 << (8 * q);
                    }
                }
                const uint32_t m = nb < 16 ? ((1u << (2 * nb)) - 1u) : ~0u;
                c32 &= m;
                v32 &= m;
                if (v32 != m) dirty[r] = 1;
            }
            ((uint32_t *)codes)[(size_t)r * halves + h] = c32;
            ((uint32_t *)valid)[(size_t)r * halves + h] = v32;
        }
        __syncwarp();
        if (!PACKED) {
            next = take();
            if (next < c1 && lane == 0 && uses_bulk(next)) issue(next, 0);   // prefetch: overlaps the walk below
        }
        // ---- walk the reads ----
        for (int r = 0; r < n_here; r++) {
            const uint64_t *cw = tile_codes + (size_t)r * b.words;
            const uint64_t *vw = valid + (size_t)r * b.words;
            if (PAIRED && (PACKED || !dirty[r])) {
                // clean read: lane owns WPL consecutive windows, rolled from one extraction; the reverse-complement
                // hash is rolled alongside (kmer_hashing.py:24-28 == bit reversal of the complemented window)
                if (MINZ) {   // hashes of the read's canonical m-mers, five consecutive positions per lane
                    const int m = t.filter_m, n_pos = b.read_len - m + 1;
                    const uint32_t mmask = (1u << (2 * m)) - 1u;
                    __syncwarp();   // the previous read's windows are done with the array
                    for (int p0 = lane * 5; p0 < (int)b.mz_len; p0 += 160) {
                        uint32_t hv[5] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
                        if (p0 < n_pos) {
                            uint64_t win = extract_window(cw, p0, (1ull << (2 * (m + 4))) - 1ull);
                            uint32_t fwd = (uint32_t)win & mmask, rcm = revcomp_m(fwd, m);
                            hv[0] = mmer_hash(fwd, rcm);
                            win >>= 2 * m;
#pragma unroll
                            for (int j = 1; j < 5; j++) {
                                const uint32_t nb = (uint32_t)win & 3u;
                                win >>= 2;
                                fwd = (fwd >> 2) | (nb << (2 * (m - 1)));
                                rcm = ((rcm << 2) | (3u - nb)) & mmask;
                                if (p0 + j < n_pos) hv[j] = mmer_hash(fwd, rcm);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 5; j++)
                            if (p0 + j < (int)b.mz_len) mzh[p0 + j] = hv[j];
                    }
                    __syncwarp();
                }
                for (int base = 0; base < b.nk; base += 32 * WPL) {
                    const int i0 = base + lane * WPL;
                    uint32_t mz[WPL] = {0, 0, 0, 0};   // MINZ: minimizer of each of the lane's windows
                    if (MINZ && i0 < b.nk) {
                        uint32_t v[MZ_W + 3];
#pragma unroll
                        for (int j = 0; j < (MZ_W + 3) / 4; j++) {
                            const uint4 q = *(const uint4 *)(mzh + i0 + 4 * j);
                            v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
                        }
                        uint32_t common = v[3];
#pragma unroll
                        for (int j = 4; j < MZ_W; j++) common = min(common, v[j]);
                        mz[0] = min(common, min(v[0], min(v[1], v[2])));
                        mz[1] = min(common, min(v[1], min(v[2], v[MZ_W])));
                        mz[2] = min(common, min(v[2], min(v[MZ_W], v[MZ_W + 1])));
                        mz[3] = min(common, min(v[MZ_W], min(v[MZ_W + 1], v[MZ_W + 2])));
                    }
                    // the lane's first window starts on a byte boundary of the packed read (i0 is a multiple of 4 bases): three byte
                    // permutes bring 96 bits from there into place, window u is those bits shifted by the constant 2u (as in K1, hash.cu);
                    // the reverse-complement hash of the first window is a bit reversal and rolls on with the base that enters on top
                    const uint32_t *cw32 = (const uint32_t *)cw;
                    const int wi = i0 >> 4;
                    const uint32_t sel = 0x3210u + 0x1111u * (((uint32_t)i0 >> 2) & 3u);
                    const uint32_t w0 = cw32[wi], w1 = cw32[wi + 1], w2 = cw32[wi + 2], w3 = cw32[wi + 3];   // (beyond the read: pad word, next read, the arrays behind)
                    const uint32_t a0 = __byte_perm(w0, w1, sel), a1 = __byte_perm(w1, w2, sel), a2 = __byte_perm(w2, w3, sel);
                    uint64_t x = (((uint64_t)a1 << 32) | a0) & mask, rc = revcomp_hash(x, b.k);
                    unsigned long long c[WPL];
                    uint32_t home[WPL], fw[WPL], fm[WPL];
                    uint32_t live = 0, pal = 0;
#pragma unroll
                    for (int u = 0; u < WPL; u++) {
                        if (u) {
                            x = (((uint64_t)__funnelshift_r(a1, a2, 2 * u) << 32) | __funnelshift_r(a0, a1, 2 * u)) & mask;
                            rc = ((rc << 2) | (3u - ((x >> (2 * (b.k - 1))) & 3u))) & mask;
                        }
                        bool ok = i0 + u < b.nk;
                        c[u] = x < rc ? x : rc;
                        if (!KODD) pal |= (uint32_t)(x == rc) << u;
                        const Hash h = hash_key(c[u]);
                        home[u] = home_bucket(t, h);
                        fm[u] = filter_mask<FK>(t, h);
                        const uint32_t word = MINZ ? filter_word_mz(t, h, mz[u]) : filter_word(t, h);
                        fw[u] = (t.filter && ok) ? ((HINTS & 1) ? ld_u32_hint(t.filter + word, pol_last) : __ldg(t.filter + word)) : 0xffffffffu;
                        live |= (uint32_t)ok << u;
                    }
#pragma unroll
                    for (int u = 0; u < WPL; u++) live &= ~((uint32_t)((fw[u] & fm[u]) != fm[u]) << u);
                    // survivors (few, scattered over lanes) are compacted into the warp's queue by ballot (a shuffle prefix sum
                    // over the four windows needs fewer instructions, but its dependent chain made the launch slower), and the
                    // table is probed 32 survivors at a time, so every HBM round trip is shared by a full warp
                    const uint32_t below = (1u << lane) - 1u;
#pragma unroll
                    for (int u = 0; u < WPL; u++) {
                        const bool mine = (live >> u) & 1u;
                        const uint32_t votes = __ballot_sync(0xffffffffu, mine);
                        if (mine) {
                            const uint32_t slot_idx = (qn + __popc(votes & below)) & (QCAP - 1);
                            qkey[slot_idx] = c[u];
                            qmeta[slot_idx] = KODD ? home[u] : (home[u] | (((pal >> u) & 1u) << 31));
                        }
                        qn += __popc(votes);
                        if (qn - qhead >= 32) {
                            __syncwarp();
                            probe_queue(32);
                        }
                    }
                }
            } else {
                // general path: reads with non-ACGT bytes, forward-only counting, or read k != table k --
                // every window is extracted on its own and each strand is an independent query
                for (int i = lane; i < b.nk; i += 32) {
                    uint64_t x = extract_window(cw, i, mask);
                    count_one(t, x);
                    if (BOTH) count_one(t, PACKED ? revcomp_hash(x, b.k) : revcomp_hash_masked(x, extract_window(vw, i, mask), b.k));
                }
            }
        }
        __syncwarp();
        wt = next;
    }
    __syncwarp();
    if (qn != qhead) probe_queue(qn - qhead);   // drain the queue
}

// ------------------------------------------------------------------ counters -> per-entry / per-node
__device__ __forceinline__ uint32_t kmer_count(const TableView &t, uint64_t km, bool wrap16) {
    Key key = make_key(km, t.k);
    uint32_t w = 0;
    if (key.ok) {
        uint32_t *cnt = find_slot(t, key.c, hash_key(key.c));
        if (cnt) w = read_orientation(cnt, key.o);
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

// ---- entries regrouped by slot: get_node_counts as a streaming pass ----
__global__ void entry_slots_kernel(TableView t, const uint64_t *__restrict__ kmers, int64_t n, unsigned long long *__restrict__ elems) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        Key key = make_key(__ldg(kmers + e), t.k);
        uint32_t *cnt = find_slot(t, key.c, hash_key(key.c));    // every index k-mer is in the table
        unsigned long long slot = cnt ? slot_id(t, cnt) : 0ull;
        elems[e] = (slot << 32) | (unsigned long long)(uint32_t)e;
    }
}
__global__ void csr_fill_kernel(TableView t, const unsigned long long *__restrict__ sorted, const uint64_t *__restrict__ kmers,
                                const uint32_t *__restrict__ nodes, int64_t n, uint32_t *__restrict__ cs_slot, uint32_t *__restrict__ cs_node) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long el = __ldg(sorted + i);
        uint32_t e = (uint32_t)el;
        Key key = make_key(__ldg(kmers + e), t.k);
        cs_slot[i] = (uint32_t)(el >> 32);
        cs_node[i] = __ldg(nodes + e) | (key.o << 31);
    }
}
__global__ void node_counts_csr_kernel(TableView t, const uint32_t *__restrict__ cs_slot, const uint32_t *__restrict__ cs_node, int64_t n,
                                       double *__restrict__ out, int64_t n_out, bool wrap16) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t nd = __ldg(cs_node + i);
        uint32_t w = read_orientation(slot_counters(t, __ldg(cs_slot + i)), nd >> 31);
        if (wrap16) w &= 0xFFFFu;
        const uint32_t node = nd & 0x7fffffffu;
        if (w && (int64_t)node < n_out) atomicAdd(out + node, (double)w);
    }
}

// get_node_counts when the node-count vector is larger than L2 keeps (C3: 50 M nodes = 400 MB of float64): every addition is then a
// read-modify-write of a random HBM line, two DRAM accesses each against a ceiling of ~37 G random accesses/s (590 M additions:
// 31 ms).  Instead the per-entry weights are materialised once in slot order (streaming), and the additions are made in passes over
// node ranges whose counts stay in L2: each pass streams (node, weight) pairs -- 8 bytes per entry -- and adds the ones of its range.
// The (node, weight) pairs are packed into ONE 32-bit word -- node in the low node_bits, weight above -- so a pass streams 4 bytes per
// entry; the few weights that do not fit (a k-mer counted >= 2^(32 - node_bits) times) are stored as 0 there and added right away
// with a plain atomic (they are rare, and then the counter really is hot).
__global__ void entry_weights_kernel(TableView t, const uint32_t *__restrict__ cs_slot, const uint32_t *__restrict__ cs_node, int64_t n,
                                     uint32_t *__restrict__ packed, int node_bits, double *__restrict__ out, int64_t n_out, bool wrap16) {
    const uint32_t w_limit = 1u << (32 - node_bits);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t nd = __ldg(cs_node + i);
        uint32_t v = read_orientation(slot_counters(t, __ldg(cs_slot + i)), nd >> 31);
        if (wrap16) v &= 0xFFFFu;
        const uint32_t node = nd & 0x7fffffffu;
        if (v >= w_limit) {
            if ((int64_t)node < n_out) atomicAdd(out + node, (double)v);
            v = 0;
        }
        packed[i] = node | (v << node_bits);
    }
}
struct U32x8 {
    uint32_t v[8];
};
__device__ __forceinline__ U32x8 ld_stream_256(const void *p) {   // streamed once per pass: must not displace the counts of the pass's range
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    U32x8 r;
    r.v[0] = (uint32_t)a; r.v[1] = (uint32_t)(a >> 32); r.v[2] = (uint32_t)b; r.v[3] = (uint32_t)(b >> 32);
    r.v[4] = (uint32_t)c; r.v[5] = (uint32_t)(c >> 32); r.v[6] = (uint32_t)d; r.v[7] = (uint32_t)(d >> 32);
    return r;
}
__global__ void node_counts_slice_kernel(const uint32_t *__restrict__ packed, int64_t n, int node_bits, double *__restrict__ out, uint32_t lo, uint32_t hi) {
    const uint32_t node_mask = (1u << node_bits) - 1u;
    const int64_t n8 = n >> 3;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n8; q += (int64_t)gridDim.x * blockDim.x) {
        const U32x8 p = ld_stream_256(packed + 8 * q);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t node = p.v[j] & node_mask, w = p.v[j] >> node_bits;
            if (w && node >= lo && node < hi) atomicAdd(out + node, (double)w);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
        const uint32_t v = packed[(n8 << 3) + threadIdx.x];
        const uint32_t node = v & node_mask, w = v >> node_bits;
        if (w && node >= lo && node < hi) atomicAdd(out + node, (double)w);
    }
}

// The passes above stream ALL (node, weight) words once per node range (c3: 9 passes of 4 GB).  Which range a word belongs to depends
// only on the entry's node, so every tile of NB_TILE entries can write its words grouped by range, at places that are the same in
// every call (nb_seg, built once: per-tile histogram of the ranges + one exclusive scan in range-major order); each range is then
// read once.  A tile is ordered in shared memory (slot = the range's first place in the tile + a shared-memory atomic) and leaves with
// coalesced stores.  Ranges are 2^nb_shift nodes (32 MB of counts).  c3 (1 B entries, 50 M nodes; ncu, profiles/r2/node_counts_ranged_launches.txt):
// 60 GB -> 32 GB per call, 10.4 -> 9.1 ms: the words pass 6.1 ms (24 GB of table + entry list read, 4 GB written; 6.6 ms with 256 threads per tile), the twelve ranges
// 0.25 ms each, which is the rate of float64 REDs on L2-resident lines (49 M additions per range = 196 G/s).
constexpr int NB_TILE = 8192, NB_THREADS = 512, NB_MAX_RANGES = 64;   // (gki_index::nb_bounds holds NB_MAX_RANGES + 1 values)
__global__ void __launch_bounds__(NB_THREADS) node_range_hist_kernel(const uint32_t *__restrict__ cs_node, int64_t n, int shift, int ranges, int64_t n_tiles,
                                                                     uint32_t *__restrict__ hist) {
    __shared__ uint32_t h[NB_MAX_RANGES];
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (threadIdx.x < NB_MAX_RANGES) h[threadIdx.x] = 0;
        __syncthreads();
        const int64_t e0 = tile * NB_TILE, e1 = min(n, e0 + NB_TILE);
        for (int64_t e = e0 + threadIdx.x; e < e1; e += NB_THREADS) atomicAdd(&h[(__ldg(cs_node + e) & 0x7fffffffu) >> shift], 1u);
        __syncthreads();
        if ((int)threadIdx.x < ranges) hist[(int64_t)threadIdx.x * n_tiles + tile] = h[threadIdx.x];
        __syncthreads();
    }
}
__global__ void __launch_bounds__(NB_THREADS) entry_weights_ranged_kernel(TableView t, const uint32_t *__restrict__ cs_slot, const uint32_t *__restrict__ cs_node,
                                                                          int64_t n, const uint32_t *__restrict__ seg, int shift, int ranges, int64_t n_tiles,
                                                                          uint32_t *__restrict__ packed, int node_bits, double *__restrict__ out, int64_t n_out,
                                                                          bool wrap16) {
    __shared__ uint32_t stage[NB_TILE];
    __shared__ uint32_t first[NB_MAX_RANGES + 1], cursor[NB_MAX_RANGES], dst[NB_MAX_RANGES], len[NB_MAX_RANGES];
    constexpr int U = 8;                 // entries per thread in flight: loads of a batch are issued before the first is used
    const uint32_t w_limit = 1u << (32 - node_bits);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t e0 = tile * NB_TILE, e1 = min(n, e0 + NB_TILE);
        if ((int)threadIdx.x < ranges) {
            const int64_t at = (int64_t)threadIdx.x * n_tiles + tile;
            const uint32_t a = __ldg(seg + at);
            dst[threadIdx.x] = a;
            len[threadIdx.x] = __ldg(seg + at + 1) - a;
        }
        __syncthreads();
        if (threadIdx.x == 0) {          // where each range starts inside the tile (a few dozen values)
            uint32_t run = 0;
            for (int r = 0; r < ranges; r++) {
                first[r] = run;
                cursor[r] = run;
                run += len[r];
            }
            first[ranges] = run;
        }
        __syncthreads();
        for (int64_t base = e0 + threadIdx.x; base < e1; base += NB_THREADS * U) {
            uint32_t nd[U], sl[U], v[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int64_t e = base + (int64_t)u * NB_THREADS;
                nd[u] = e < e1 ? __ldg(cs_node + e) : 0u;
                sl[u] = e < e1 ? __ldg(cs_slot + e) : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; u++) v[u] = base + (int64_t)u * NB_THREADS < e1 ? read_orientation(slot_counters(t, sl[u]), nd[u] >> 31) : 0u;
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (base + (int64_t)u * NB_THREADS >= e1) continue;
                uint32_t w = wrap16 ? (v[u] & 0xFFFFu) : v[u];
                const uint32_t node = nd[u] & 0x7fffffffu;
                if (w >= w_limit) {
                    if ((int64_t)node < n_out) atomicAdd(out + node, (double)w);
                    w = 0;
                }
                stage[atomicAdd(&cursor[node >> shift], 1u)] = node | (w << node_bits);
            }
        }
        __syncthreads();
        for (int r = 0; r < ranges; r++) {
            const uint32_t lo = first[r], n_r = len[r], to = dst[r];
            for (uint32_t j = threadIdx.x; j < n_r; j += NB_THREADS) packed[(size_t)to + j] = stage[lo + j];
        }
        __syncthreads();
    }
}
// the words of one node range: packed[lo .. hi)
__global__ void node_counts_range_kernel(const uint32_t *__restrict__ packed, int64_t lo, int64_t hi, int node_bits, double *__restrict__ out) {
    const uint32_t node_mask = (1u << node_bits) - 1u;
    const int64_t a = (lo + 7) & ~(int64_t)7, b = hi & ~(int64_t)7;       // 32-byte aligned interior, 256-bit streaming loads
    if (a >= b) {
        for (int64_t i = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
            const uint32_t v = __ldg(packed + i);
            if (v >> node_bits) atomicAdd(out + (v & node_mask), (double)(v >> node_bits));
        }
        return;
    }
    for (int64_t q = (a >> 3) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < (b >> 3); q += (int64_t)gridDim.x * blockDim.x) {
        const U32x8 p = ld_stream_256(packed + 8 * q);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t w = p.v[j] >> node_bits;
            if (w) atomicAdd(out + (p.v[j] & node_mask), (double)w);
        }
    }
    if (blockIdx.x == 0) {
        for (int64_t i = lo + threadIdx.x; i < a; i += blockDim.x) {
            const uint32_t v = __ldg(packed + i);
            if (v >> node_bits) atomicAdd(out + (v & node_mask), (double)(v >> node_bits));
        }
        for (int64_t i = b + threadIdx.x; i < hi; i += blockDim.x) {
            const uint32_t v = __ldg(packed + i);
            if (v >> node_bits) atomicAdd(out + (v & node_mask), (double)(v >> node_bits));
        }
    }
}

__global__ void query_counts_kernel(TableView t, const uint64_t *__restrict__ queries, int64_t nq, uint32_t *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = kmer_count(t, __ldg(queries + i), false);
}

// ------------------------------------------------------------------ host side
static void destroy_pipeline(gki_index *ix);

void destroy_count_table(gki_index *ix) {
    destroy_pipeline(ix);
    cudaFree(ix->table.buckets);
    cudaFree((void *)ix->table.filter);
    cudaFree(ix->cs_slot);
    cudaFree(ix->cs_node);
    cudaFree(ix->nb_seg);
    ix->cs_slot = ix->cs_node = ix->nb_seg = nullptr;
    ix->nb_ranges = 0;
    ix->table = TableView{};
    ix->table_bytes = ix->filter_bytes = 0;
}

// stage times of the one-off table build, printed with GKI_BUILD_DEBUG (profiles/r2/prepare_counting_stages_c3.txt)
struct StageClock {
    cudaStream_t s;
    bool on;
    std::chrono::steady_clock::time_point t0;
    explicit StageClock(cudaStream_t stream) : s(stream), on(getenv("GKI_BUILD_DEBUG") != nullptr) {
        if (on) {
            cudaStreamSynchronize(s);
            t0 = std::chrono::steady_clock::now();
        }
    }
    void lap(const char *what) {
        if (!on) return;
        cudaStreamSynchronize(s);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[gki prepare_counting] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// Build the table on first use.  k > 0 selects canonical keys (needs every index k-mer < 4^k), k == 0 raw keys.
static int ensure_table(gki_index *ix, int k, cudaStream_t s) {
    if (ix->table.buckets) return GKI_OK;
    StageClock clock(s);
    if (const char *e = experiment_knob("GKI_L2_FETCH")) {   // experiment knob: L2 fetch granularity for misses (32 / 64 / 128 bytes)
        GKI_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e)));
    }
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
    clock.lap("count distinct k-mers");

    TableView t{};
    t.k = k;
    uint64_t buckets = ((distinct + 1) / 2 + 17) & ~1ull;   // 4 slots per bucket -> load factor <= 0.5; even (next_bucket)
    GKI_REQUIRE(buckets < (1ull << 31), GKI_ERR_UNSUPPORTED, "count table: too many distinct k-mers");
    t.n_buckets = (uint32_t)buckets;
    const size_t n_slots = ((size_t)buckets + 1) * SLOTS_PER_BUCKET;
    // Everything is built into locals: the table, the slot-ordered entry list and the byte counts are published together once
    // every step has succeeded, so a failed build (an allocation, an insertion) leaves no half-made table for the next call to find.
    uint32_t *cs_slot = nullptr, *cs_node = nullptr, *filter = nullptr;
    size_t table_bytes = ((size_t)buckets + 1) * sizeof(Bucket), filter_bytes = 0;
    unsigned int failed = 0;
    auto build_rest = [&]() -> int {
        GKI_CUDA(cudaMalloc((void **)&t.buckets, ((size_t)buckets + 1) * sizeof(Bucket)));
        table_init_kernel<<<grid_for((int64_t)(buckets + 1) * 8, 256 * 4, device_info().sms * 16), 256, 0, s>>>(t.buckets, (size_t)buckets + 1);
        GKI_CHECK_LAUNCH();

        // Bloom filter: as many bits per key as fit the L2 budget (<= 16); below 1.5 bits per key it filters nothing
        size_t budget = (size_t)32 << 20;   // measured optimum at c2 (profiles/r1/tune_filter_v4.jsonl): flat from 32 to 40 MB, worse
                                            // below (false positives) and above (the filter starts missing L2)
        if (const char *e = getenv("GKI_FILTER_MAX_MB")) budget = (size_t)atoi(e) << 20;
        size_t want = (size_t)distinct * 2;                  // 16 bits per key
        // Indexes too large for an L2-resident filter still profit from an HBM-resident one (measured at 500 M distinct k-mers:
        // kernel 297 ms without, 164 ms with a 512 MB filter): a filter miss costs one small fetch from a compact region instead
        // of a bucket line from the 32x larger table.  16 bits per key: the fetch costs the same whatever the filter's size, and
        // every false positive is a table line (c3, count kernel: 51.4 ms at 8 bits per key, 47.3 at 12, 46.1 at 16 and at 24).
        if (!getenv("GKI_FILTER_MAX_MB") && (size_t)distinct > budget) budget = (size_t)distinct * 2;
        size_t fbytes = want < budget ? want : budget;
        fbytes = (fbytes + 31) & ~(size_t)31;   // whole 32-byte sectors
        if (fbytes < 64) fbytes = 64;
        double bits_per_key = distinct ? (double)fbytes * 8.0 / (double)distinct : 16.0;
        if (budget > 0 && bits_per_key >= 1.5) {
            GKI_CUDA(cudaMalloc((void **)&filter, fbytes));
            GKI_CUDA(cudaMemsetAsync(filter, 0, fbytes, s));
            t.filter = filter;
            t.filter_words = (uint32_t)(fbytes / 4);
            // a filter that cannot stay in L2 is addressed by minimizer (odd k in 27..31: 17 m-mers of m = k - 16 <= 15 bases)
            const bool mz_possible = k >= 27 && k <= 31 && (k & 1);
            bool mz = mz_possible && fbytes > ((size_t)48 << 20);
            if (const char *e = getenv("GKI_FILTER_MZ")) mz = mz_possible && atoi(e) != 0;
            t.filter_m = mz ? k - 16 : 0;
            t.filter_k = bits_per_key >= 5.0 ? 3 : (bits_per_key >= 3.0 ? 2 : 1);
            if (const char *e = getenv("GKI_FILTER_K")) t.filter_k = atoi(e) < 1 ? 1 : (atoi(e) > 3 ? 3 : atoi(e));
            filter_bytes = fbytes;
        }
        clock.lap("allocate + initialise table");
        table_insert_kernel<<<grid_n, 256, 0, s>>>(t, ix->kmers, ix->n, filter, (unsigned int *)counters.ptr + 2, ix->view().fm);
        GKI_CHECK_LAUNCH();
        GKI_CUDA(cudaMemcpyAsync(&failed, (unsigned int *)counters.ptr + 2, 4, cudaMemcpyDeviceToHost, s));
        GKI_CUDA(cudaStreamSynchronize(s));
        GKI_REQUIRE(failed == 0, GKI_ERR_CUDA, "count table: %u insertions failed", failed);
        clock.lap("insert keys + filter bits");
        // regroup the entries by slot (one-time sort) so that get_node_counts reads the counters in table order
        if (n_slots < (1ull << 32) && ix->max_node < (1ll << 31) && !getenv("GKI_NO_CSR")) {
            Scratch a, b, hist;
            GKI_TRY(a.alloc((size_t)ix->n * 8, s));
            entry_slots_kernel<<<grid_n, 256, 0, s>>>(t, ix->kmers, ix->n, a.as<unsigned long long>());
            GKI_CHECK_LAUNCH();
            clock.lap("slot of every entry");
            int bits = 1;
            while ((1ull << bits) < n_slots) bits++;
            const unsigned long long *sorted = nullptr;
            GKI_TRY(radix_sort_packed(a, b, hist, ix->n, bits, &sorted, s));
            clock.lap("sort entries by slot");
            GKI_CUDA(cudaMalloc((void **)&cs_slot, (size_t)ix->n * 4));
            GKI_CUDA(cudaMalloc((void **)&cs_node, (size_t)ix->n * 4));
            csr_fill_kernel<<<grid_n, 256, 0, s>>>(t, sorted, ix->kmers, ix->nodes, ix->n, cs_slot, cs_node);
            GKI_CHECK_LAUNCH();
            GKI_CUDA(cudaStreamSynchronize(s));
            clock.lap("slot-ordered entry list");
            table_bytes += (size_t)ix->n * 8;
        }
        return GKI_OK;
    };
    const int rc = build_rest();
    if (rc != GKI_OK) {
        cudaStreamSynchronize(s);
        cudaFree(t.buckets);
        cudaFree(filter);
        cudaFree(cs_slot);
        cudaFree(cs_node);
        return rc;
    }
    ix->table = t;
    ix->filter_bytes = filter_bytes;
    ix->cs_slot = cs_slot;
    ix->cs_node = cs_node;
    ix->table_bytes = table_bytes;
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
    int grid_mult = 4;   // more CTAs than are resident (120 M queries: 1.21 -> 1.10 ms; 16: 1.11, 64: 1.22)
    if (const char *e = experiment_knob("GKI_COUNTK_GRID_MULT")) grid_mult = atoi(e) > 0 ? atoi(e) : grid_mult;
    int grid = grid_for((nq + 32 * CK_U - 1) / (32 * CK_U), COUNT_WARPS, device_info().sms * 8 * grid_mult);
    count_kmers_kernel<<<grid, COUNT_THREADS, 0, s>>>(ix->table, dq, nq);
    GKI_CHECK_LAUNCH();
    return GKI_OK;
}

static std::mutex g_launch_mutex;   // the packing lanes launch from their own host threads

template <bool BOTH, bool PAIRED, int MINB, int HINTS = 0, bool PACKED = false, bool KODD = false, int FK = 0, bool MINZ = false>
static int launch_count_reads_t(gki_index *ix, const WarpBatch &b, cudaStream_t s) {
    std::lock_guard<std::mutex> lock(g_launch_mutex);
    const size_t smem = (size_t)b.warp_bytes * COUNT_WARPS;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        GKI_CUDA(cudaFuncSetAttribute(count_reads_kernel<BOTH, PAIRED, MINB, HINTS, PACKED, KODD, FK, MINZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    int blocks_per_sm = 0;
    GKI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, count_reads_kernel<BOTH, PAIRED, MINB, HINTS, PACKED, KODD, FK, MINZ>, COUNT_THREADS, smem));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    if (const char *e = experiment_knob("GKI_COUNT_CTAS")) {   // experiment knob: fewer resident CTAs per SM (occupancy sensitivity)
        const int v = atoi(e);
        if (v >= 1 && v < blocks_per_sm) blocks_per_sm = v;
    }
    // The grid is 16 times what is resident.  Reads drawn from the indexed sequence cost several times what other reads cost (a
    // table line per window), so warps that stride over the batch in one persistent wave finish up to 10 % apart and the launch
    // waits for the slowest; CTAs of ~16 tiles per warp are dealt out by the hardware as slots free up.  c2: 9.10 -> 8.40 ms,
    // c3: 46.0 -> 40.0 ms (profiles/r2/count_kernel_grid_oversubscription.log; 8: 8.57, 16: 8.43, 32: 8.41 ms).
    int grid_mult = 16;
    if (const char *e = experiment_knob("GKI_COUNT_GRID_MULT")) grid_mult = atoi(e) > 0 ? atoi(e) : grid_mult;
    int grid = grid_for(b.n_wtiles, COUNT_WARPS, device_info().sms * blocks_per_sm * grid_mult);
    count_reads_kernel<BOTH, PAIRED, MINB, HINTS, PACKED, KODD, FK, MINZ><<<grid, COUNT_THREADS, smem, s>>>(ix->table, b);
    GKI_CHECK_LAUNCH();
    return GKI_OK;
}

// the production instantiations of the both-strands kernel: odd k and 3 filter bits (the defaults) are compile-time facts
template <bool PACKED> static int launch_paired(gki_index *ix, const WarpBatch &b, cudaStream_t s) {
    const bool fk3 = ix->table.filter && ix->table.filter_k == 3;
    if (ix->table.filter && ix->table.filter_m) {   // minimizer-addressed filter: odd k in 27..31 by construction (ensure_table)
        return fk3 ? launch_count_reads_t<true, true, 4, 0, PACKED, true, 3, true>(ix, b, s) : launch_count_reads_t<true, true, 4, 0, PACKED, true, 0, true>(ix, b, s);
    }
#ifdef GKI_EXPERIMENT_KNOBS   // 3 or 5 resident CTAs per SM with the register budget that goes with them (85 / 48 instead of 64)
    if (const char *e = experiment_knob("GKI_COUNT_MINB")) {
        if (!PACKED && (b.k & 1) && fk3 && atoi(e) == 3) return launch_count_reads_t<true, true, 3, 0, false, true, 3>(ix, b, s);
        if (!PACKED && (b.k & 1) && fk3 && atoi(e) == 5) return launch_count_reads_t<true, true, 5, 0, false, true, 3>(ix, b, s);
    }
#endif
    if (b.k & 1) return fk3 ? launch_count_reads_t<true, true, 4, 0, PACKED, true, 3>(ix, b, s) : launch_count_reads_t<true, true, 4, 0, PACKED, true, 0>(ix, b, s);
    return launch_count_reads_t<true, true, 4, 0, PACKED, false, 0>(ix, b, s);
}

// reads: device rows
static int launch_count_reads(gki_index *ix, const uint8_t *dreads, int64_t n_reads, int32_t read_len, int64_t stride, int32_t k,
                              int32_t both, cudaStream_t s) {
    WarpBatch b{};
    b.reads = dreads;
    b.n_reads = n_reads;
    b.row_stride = stride;
    b.read_len = read_len;
    b.k = k;
    b.nk = read_len - k + 1;
    b.words = (read_len + 31) / 32 + 1;
    const bool minz = both && ix->table.k == k && ix->table.filter && ix->table.filter_m;
    b.mz_len = minz ? (uint32_t)((read_len + 8) & ~3) : 0u;   // m-mer positions of a read, padded so that every window's 20 loads stay inside
    auto mz_offset = [&](int rpw) {
        size_t stage = (((size_t)rpw * read_len + 15) & ~(size_t)15) + 16;
        size_t bytes = 16 + stage + 2 * (size_t)rpw * b.words * 8 + (size_t)QCAP * 12 + (size_t)rpw * 4;
        return (bytes + 15) & ~(size_t)15;
    };
    auto warp_bytes = [&](int rpw) { return mz_offset(rpw) + (size_t)b.mz_len * 4; };
    int rpw = 8;
    if (const char *e = experiment_knob("GKI_RPW")) rpw = atoi(e) < 1 ? 1 : (atoi(e) > 32 ? 32 : atoi(e));
    while (rpw > 1 && warp_bytes(rpw) * COUNT_WARPS > 56 * 1024) rpw >>= 1;
    GKI_REQUIRE(warp_bytes(rpw) * COUNT_WARPS <= 200 * 1024, GKI_ERR_UNSUPPORTED, "gki_count_reads: read_len %d too long for the tile path", read_len);
    b.rpw = rpw;
    b.mz_off = (uint32_t)mz_offset(rpw);
    b.inv_halves = 65536u / (2u * (uint32_t)b.words) + 1u;
    for (uint32_t task = 0; task < (uint32_t)rpw * 2u * (uint32_t)b.words; task++)
        GKI_REQUIRE(((task * b.inv_halves) >> 16) == task / (2u * (uint32_t)b.words), GKI_ERR_UNSUPPORTED, "gki_count_reads: read_len %d too long for the tile path", read_len);
    b.stage_bytes = (uint32_t)((((size_t)rpw * read_len + 15) & ~(size_t)15) + 16);
    b.warp_bytes = (uint32_t)warp_bytes(rpw);
    b.n_wtiles = (n_reads + rpw - 1) / rpw;
    b.bulk_ok = (stride == read_len) && (((uintptr_t)dreads & 15) == 0) && (((int64_t)rpw * read_len) % 16 == 0);
    if (!both) return launch_count_reads_t<false, false, 4>(ix, b, s);
    if (ix->table.k != k) return launch_count_reads_t<true, false, 4>(ix, b, s);
#ifdef GKI_EXPERIMENT_KNOBS   // kernel variants of profiles/tune_count.py: 7 = L2 eviction hints, 16 = drop the survivors, 64 = 64-byte L2 fills
    int hints = 0;
    if (const char *e = experiment_knob("GKI_HINTS")) hints = atoi(e);
    if (hints == 7) return launch_count_reads_t<true, true, 4, 7>(ix, b, s);
    if (hints == 16) return launch_count_reads_t<true, true, 4, 16>(ix, b, s);
    if (hints == 64) return launch_count_reads_t<true, true, 4, 64>(ix, b, s);
#endif
    return launch_paired<false>(ix, b, s);
}

// packed reads (gki_pack_reads layout): device rows of ceil(read_len / 32) 64-bit words
static int launch_count_packed_reads(gki_index *ix, const uint64_t *dpacked, int64_t n_reads, int32_t read_len, int32_t k, int32_t both,
                                     cudaStream_t s) {
    WarpBatch b{};
    b.reads = (const uint8_t *)dpacked;
    b.n_reads = n_reads;
    b.read_len = read_len;
    b.k = k;
    b.nk = read_len - k + 1;
    b.words = (read_len + 31) / 32;
    b.row_stride = (int64_t)b.words * 8;
    const bool minz = both && ix->table.k == k && ix->table.filter && ix->table.filter_m;
    b.mz_len = minz ? (uint32_t)((read_len + 8) & ~3) : 0u;
    auto mz_offset = [&](int rpw) {
        size_t stage = (size_t)rpw * b.words * 8 + 16;
        return (16 + 2 * stage + (size_t)QCAP * 12 + 15) & ~(size_t)15;
    };
    auto warp_bytes = [&](int rpw) { return mz_offset(rpw) + (size_t)b.mz_len * 4; };
    int rpw = 8;
    while (rpw > 1 && warp_bytes(rpw) * COUNT_WARPS > 56 * 1024) rpw >>= 1;
    GKI_REQUIRE(warp_bytes(rpw) * COUNT_WARPS <= 200 * 1024, GKI_ERR_UNSUPPORTED, "gki_count_packed_reads: read_len %d too long for the tile path", read_len);
    b.rpw = rpw;
    b.mz_off = (uint32_t)mz_offset(rpw);
    b.stage_bytes = (uint32_t)((size_t)rpw * b.words * 8 + 16);
    b.warp_bytes = (uint32_t)warp_bytes(rpw);
    b.n_wtiles = (n_reads + rpw - 1) / rpw;
    b.bulk_ok = (((uintptr_t)dpacked & 15) == 0) && (((int64_t)rpw * b.words * 8) % 16 == 0);
    if (!both) return launch_count_reads_t<false, false, 4, 0, true>(ix, b, s);
    if (ix->table.k != k) return launch_count_reads_t<true, false, 4, 0, true>(ix, b, s);
    return launch_paired<true>(ix, b, s);
}

// ---- host reads through CPU packing lanes + the copy engine ----
// PCIe carries ASCII reads at ~52 GB/s, a quarter of what the count kernel consumes.  Host threads ("lanes") therefore
// pack chunks to 2 bits per base (ingest.cpp) into pinned buffers -- 3.75x fewer bytes on the bus -- while the calling thread
// keeps the copy engine busy with plain ASCII chunks; both take chunks from one atomic counter, so the split follows
// whatever the host cores and the bus deliver.  Reads with a non-ACGT byte travel as ASCII next to their chunk; a chunk
// with more than 1/16 such reads is left to the ASCII path.
struct PackLane {
    cudaStream_t stream = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr};
    uint64_t *h_packed[2] = {nullptr, nullptr}, *d_packed[2] = {nullptr, nullptr};
    uint8_t *h_dirty[2] = {nullptr, nullptr}, *d_dirty[2] = {nullptr, nullptr};
    bool used[2] = {false, false};
    int tog = 0;
};

// one batch handed to the lanes
struct PackJob {
    const uint8_t *reads = nullptr;
    const int64_t *row_offsets = nullptr;   // sequence lines of a file: row r starts at reads + row_offsets[r]
    int64_t n_reads = 0, row_stride = 0, n_units = 0;
    int32_t read_len = 0, k = 0, both = 0;
};

constexpr int64_t UNIT_READS = 32768;      // a lane packs one unit at a time (about 1 ms of CPU work at 150 bp)
constexpr int64_t ASCII_UNITS = 4;         // the copy engine moves four units per transfer

struct Pipeline {
    gki_index *ix = nullptr;
    int n_lanes = 0;
    int32_t read_len = 0;
    int64_t dirty_cap = 0;
    std::vector<PackLane> lanes;
    std::vector<std::thread> threads;
    cudaEvent_t start = nullptr;
    // job hand-off: workers sleep on cv until generation changes; the caller sleeps on cv_done until all have finished
    std::mutex mu;
    std::condition_variable cv, cv_done;
    uint64_t generation = 0;
    int running = 0;
    bool quit = false;
    PackJob job;
    std::atomic<int64_t> next{0};
    std::atomic<int> failed{0};
    uint64_t calls = 0;              // large host batches seen; rate[m]: bytes/s of the last one run in mode m: packing lanes alone (0),
    double rate[3] = {0.0, 0.0, 0.0};   // lanes + the copy-engine lane (1), the copy engine alone (2)
    std::vector<int64_t> deferred;   // units left to the ASCII path (too many dirty reads); guarded by mu
    std::string error;               // guarded by mu

    int lane_body(int li);
    void worker(int li);
};

int Pipeline::lane_body(int li) {
    PackLane &L = lanes[(size_t)li];
    const PackJob j = job;
    const size_t words = (size_t)(j.read_len + 31) / 32;
    while (!failed.load(std::memory_order_relaxed)) {
        const int64_t u = next.fetch_add(1);
        if (u >= j.n_units) break;
        const int64_t r0 = u * UNIT_READS, r1 = r0 + UNIT_READS < j.n_reads ? r0 + UNIT_READS : j.n_reads;
        const int t = L.tog;
        if (L.used[t]) GKI_CUDA(cudaEventSynchronize(L.done[t]));
        int64_t n_dirty = 0;
        const int64_t clean = pack_rows(j.reads, j.row_stride, j.row_offsets, j.read_len, r0, r1, L.h_packed[t], L.h_dirty[t], nullptr, dirty_cap, &n_dirty, 0);
        if (n_dirty > dirty_cap) {
            std::lock_guard<std::mutex> lock(mu);
            deferred.push_back(u);
            continue;
        }
        if (clean) {
            GKI_CUDA(cudaMemcpyAsync(L.d_packed[t], L.h_packed[t], (size_t)clean * words * 8, cudaMemcpyHostToDevice, L.stream));
            GKI_TRY(launch_count_packed_reads(ix, L.d_packed[t], clean, j.read_len, j.k, j.both, L.stream));
        }
        if (n_dirty) {
            GKI_CUDA(cudaMemcpyAsync(L.d_dirty[t], L.h_dirty[t], (size_t)n_dirty * j.read_len, cudaMemcpyHostToDevice, L.stream));
            GKI_TRY(launch_count_reads(ix, L.d_dirty[t], n_dirty, j.read_len, j.read_len, j.k, j.both, L.stream));
        }
        GKI_CUDA(cudaEventRecord(L.done[t], L.stream));
        L.used[t] = true;
        L.tog ^= 1;
    }
    return GKI_OK;
}

void Pipeline::worker(int li) {
    cudaSetDevice(ix->device);
    uint64_t seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lock(mu);
            cv.wait(lock, [&] { return quit || generation != seen; });
            if (quit) return;
            seen = generation;
        }
        const int rc = lane_body(li);
        std::lock_guard<std::mutex> lock(mu);
        if (rc != GKI_OK) {
            if (error.empty()) error = gki_last_error();
            failed.store(1);
        }
        if (--running == 0) cv_done.notify_all();
    }
}

static void destroy_pipeline(gki_index *ix) {
    Pipeline *p = (Pipeline *)ix->pipeline;
    if (!p) return;
    {
        std::lock_guard<std::mutex> lock(p->mu);
        p->quit = true;
    }
    p->cv.notify_all();
    for (auto &t : p->threads) t.join();
    for (PackLane &L : p->lanes) {
        for (int j = 0; j < 2; j++) {
            cudaFreeHost(L.h_packed[j]);
            cudaFreeHost(L.h_dirty[j]);
            cudaFree(L.d_packed[j]);
            cudaFree(L.d_dirty[j]);
            if (L.done[j]) cudaEventDestroy(L.done[j]);
        }
        if (L.stream) cudaStreamDestroy(L.stream);
    }
    if (p->start) cudaEventDestroy(p->start);
    delete p;
    ix->pipeline = nullptr;
}

static int allocate_pipeline(Pipeline *p, int32_t read_len) {
    const size_t words = (size_t)(read_len + 31) / 32;
    GKI_CUDA(cudaEventCreateWithFlags(&p->start, cudaEventDisableTiming));
    for (PackLane &L : p->lanes) {
        GKI_CUDA(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
        for (int j = 0; j < 2; j++) {
            GKI_CUDA(cudaEventCreateWithFlags(&L.done[j], cudaEventDisableTiming));
            GKI_CUDA(cudaHostAlloc((void **)&L.h_packed[j], (size_t)UNIT_READS * words * 8 + 16, cudaHostAllocDefault));
            GKI_CUDA(cudaHostAlloc((void **)&L.h_dirty[j], (size_t)p->dirty_cap * read_len + 16, cudaHostAllocDefault));
            GKI_CUDA(cudaMalloc((void **)&L.d_packed[j], (size_t)UNIT_READS * words * 8 + 16));
            GKI_CUDA(cudaMalloc((void **)&L.d_dirty[j], (size_t)p->dirty_cap * read_len + 16));
        }
    }
    return GKI_OK;
}

static int ensure_pipeline(gki_index *ix, int n_lanes, int32_t read_len) {
    Pipeline *p = (Pipeline *)ix->pipeline;
    if (p && p->n_lanes == n_lanes && p->read_len == read_len) return GKI_OK;
    destroy_pipeline(ix);
    p = new Pipeline();
    ix->pipeline = p;
    p->ix = ix;
    p->n_lanes = n_lanes;
    p->read_len = read_len;
    p->dirty_cap = UNIT_READS / 16 + 1;
    p->lanes.resize((size_t)n_lanes);
    const int rc = allocate_pipeline(p, read_len);
    if (rc != GKI_OK) {   // a half-built pipeline has no worker threads: never leave it behind
        destroy_pipeline(ix);
        return rc;
    }
    for (int i = 0; i < n_lanes; i++) p->threads.emplace_back(&Pipeline::worker, p, i);
    return GKI_OK;
}

// row_offsets != NULL: the rows are sequence lines of a mapped file (reads + row_offsets[r]); they are not dense, so only the
// packing lanes take units and the calling thread handles what they defer.
static int count_reads_host_pipeline(gki_index *ix, const uint8_t *reads, const int64_t *row_offsets, int64_t n_reads, int32_t read_len,
                                     int64_t row_stride, int32_t k, int32_t both, cudaStream_t s, int n_lanes) {
    GKI_TRY(ensure_pipeline(ix, n_lanes, read_len));
    Pipeline *p = (Pipeline *)ix->pipeline;
    GKI_TRY(ensure_staging(ix, (size_t)ASCII_UNITS * UNIT_READS * read_len + 16));
    GKI_CUDA(cudaEventRecord(p->start, s));            // lane kernels follow whatever the caller queued on s
    for (PackLane &L : p->lanes) GKI_CUDA(cudaStreamWaitEvent(L.stream, p->start, 0));
    // The copy engine and the packing threads read the same host memory, and which mix pays depends on the host and on how many
    // ranks share it: on the 16-thread bench box 14 packers alone are faster than packers + an ASCII transfer lane (13.7 ms vs 15.5 ms
    // per 1.5 GB), on the 24-thread two-GPU box the copy lane adds 19 %, and with eight ranks on one host (three lanes each, all
    // reading one memory system) every GPU's own PCIe link alone can be the fastest way in.  So the pipeline tries its three modes
    // -- lanes alone (0), lanes + copy lane (1), copy engine alone (2) -- twice each on its first large calls and keeps the fastest
    // (bytes per second of the whole call), looking at another mode again every 64th call.  GKI_PIPELINE_DMA=0/1/2 overrides.
    int mode = row_offsets ? 0 : 1;
    const bool adaptive = !row_offsets && !getenv("GKI_PIPELINE_DMA");
    if (const char *e = getenv("GKI_PIPELINE_DMA")) mode = row_offsets ? 0 : (atoi(e) < 0 ? 0 : (atoi(e) > 2 ? 2 : atoi(e)));
    if (adaptive) {
        const uint64_t call_no = p->calls++;
        int best = 0;
        for (int m = 1; m < 3; m++)
            if (p->rate[m] > p->rate[best]) best = m;
        if (call_no == 0) mode = 1;                                   // first call: warm-up, not recorded
        else if (call_no <= 6) mode = (int)(call_no % 3);             // every mode twice (the better trial is kept)
        else if (call_no % 64 == 63) mode = (best + 1 + (int)((call_no / 64) & 1)) % 3;   // a look at a losing mode
        else mode = best;
    }
    const bool dma_lane = mode >= 1, use_lanes = mode <= 1;
    {
        std::lock_guard<std::mutex> lock(p->mu);
        p->job.reads = reads;
        p->job.row_offsets = row_offsets;
        p->job.n_reads = n_reads;
        p->job.row_stride = row_stride;
        p->job.n_units = (n_reads + UNIT_READS - 1) / UNIT_READS;
        p->job.read_len = read_len;
        p->job.k = k;
        p->job.both = both;
        p->next.store(0);
        p->failed.store(0);
        p->deferred.clear();
        p->error.clear();
        p->running = use_lanes ? n_lanes : 0;
        if (use_lanes) p->generation++;
    }
    if (use_lanes) p->cv.notify_all();
    // the calling thread feeds the copy engine with ASCII transfers (double-buffered staging, kernels on s)
    int slot = 0;
    bool staged[2] = {false, false};
    auto ascii_units = [&](int64_t u0, int64_t n_units) -> int {
        const int64_t r0 = u0 * UNIT_READS;
        const int64_t r1 = r0 + n_units * UNIT_READS < n_reads ? r0 + n_units * UNIT_READS : n_reads, cnt = r1 - r0;
        if (cnt <= 0) return GKI_OK;
        if (staged[slot]) GKI_CUDA(cudaEventSynchronize(ix->done[slot]));   // paces this lane at bus speed
        if (row_offsets) {   // gather the lines into a dense block first (rare: only units the lanes deferred)
            std::vector<uint8_t> dense((size_t)cnt * read_len);
            for (int64_t r = 0; r < cnt; r++) memcpy(dense.data() + (size_t)r * read_len, reads + row_offsets[r0 + r], (size_t)read_len);
            GKI_CUDA(cudaMemcpyAsync(ix->stage[slot], dense.data(), dense.size(), cudaMemcpyHostToDevice, ix->copy_stream));
            GKI_CUDA(cudaStreamSynchronize(ix->copy_stream));   // the block goes out of scope
        } else if (row_stride == read_len)
            GKI_CUDA(cudaMemcpyAsync(ix->stage[slot], reads + r0 * row_stride, (size_t)cnt * read_len, cudaMemcpyHostToDevice, ix->copy_stream));
        else
            GKI_CUDA(cudaMemcpy2DAsync(ix->stage[slot], read_len, reads + r0 * row_stride, row_stride, read_len, cnt, cudaMemcpyHostToDevice, ix->copy_stream));
        GKI_CUDA(cudaEventRecord(ix->ready[slot], ix->copy_stream));
        GKI_CUDA(cudaStreamWaitEvent(s, ix->ready[slot], 0));
        GKI_TRY(launch_count_reads(ix, (const uint8_t *)ix->stage[slot], cnt, read_len, read_len, k, both, s));
        GKI_CUDA(cudaEventRecord(ix->done[slot], s));
        staged[slot] = true;
        slot ^= 1;
        return GKI_OK;
    };
    int rc = GKI_OK;
    const auto t_start = std::chrono::steady_clock::now();
    while (dma_lane && rc == GKI_OK && !p->failed.load(std::memory_order_relaxed)) {
        const int64_t u = p->next.fetch_add(ASCII_UNITS);
        if (u >= p->job.n_units) break;
        rc = ascii_units(u, u + ASCII_UNITS <= p->job.n_units ? ASCII_UNITS : p->job.n_units - u);
    }
    if (rc != GKI_OK) p->failed.store(1);
    {
        std::unique_lock<std::mutex> lock(p->mu);
        p->cv_done.wait(lock, [&] { return p->running == 0; });
    }
    for (size_t i = 0; rc == GKI_OK && !p->failed.load() && i < p->deferred.size(); i++) rc = ascii_units(p->deferred[i], 1);
    for (PackLane &L : p->lanes) cudaStreamSynchronize(L.stream);
    cudaStreamSynchronize(s);
    if (adaptive && rc == GKI_OK && !p->failed.load()) {
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        if (secs > 0 && p->calls > 1) {
            const double rate = (double)n_reads * read_len / secs;
            double &slot = p->rate[mode];
            slot = (p->calls <= 7 && slot > rate) ? slot : rate;   // exploration keeps the better of its two trials per mode
            if (getenv("GKI_PIPELINE_DEBUG")) fprintf(stderr, "[gki pipeline] call %llu mode=%d %.1f GB/s (best: lanes %.1f, lanes+copy %.1f, copy %.1f)\n",
                                                      (unsigned long long)p->calls, mode, rate / 1e9, p->rate[0] / 1e9, p->rate[1] / 1e9, p->rate[2] / 1e9);
        }
    }
    if (rc == GKI_OK && p->failed.load()) {
        set_error("%s", p->error.empty() ? "gki_count_reads: a packing lane failed" : p->error.c_str());
        rc = GKI_ERR_CUDA;
    }
    return rc;
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
    if (!ix->table.buckets) return GKI_OK;   // nothing counted yet
    const size_t nb = (size_t)ix->table.n_buckets + 1;
    table_reset_kernel<<<grid_for((int64_t)nb * 2, 256 * 4, device_info().sms * 16), 256, 0, (cudaStream_t)stream>>>(ix->table.buckets, nb);
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
    // (pageable arrays: every chunk crosses through the threaded pinned-buffer copy of runtime.cu, in larger chunks)
    const bool threaded = wants_parallel_copy(queries, (size_t)nq * 8);
    const int64_t chunk = threaded ? (16 << 20) : (4 << 20);   // 4 Mi queries = 32 MiB
    GKI_TRY(ensure_staging(ix, (size_t)chunk * 8));
    int c = 0;
    for (int64_t off = 0; off < nq; off += chunk, ++c) {
        int bsel = c & 1;
        int64_t cnt = nq - off < chunk ? nq - off : chunk;
        if (c >= 2) GKI_CUDA(cudaStreamWaitEvent(ix->copy_stream, ix->done[bsel], 0));
        if (threaded) GKI_TRY(parallel_host_copy(ix->stage[bsel], const_cast<uint64_t *>(queries + off), (size_t)cnt * 8, true, ix->copy_stream));
        else GKI_CUDA(cudaMemcpyAsync(ix->stage[bsel], queries + off, (size_t)cnt * 8, cudaMemcpyHostToDevice, ix->copy_stream));
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
    // host reads, large batch: CPU packing lanes + the copy engine share the chunks (count_reads_host_pipeline)
    {
        const int n_lanes = default_pack_threads();
        if (n_lanes > 0 && n_reads >= 16 * UNIT_READS) return count_reads_host_pipeline(ix, reads, nullptr, n_reads, read_len, row_stride, k, both_strands, s, n_lanes);
    }
    // host reads, small batch: rows are compacted to dense device rows (so every full tile is one TMA bulk copy) in
    // chunks; the copy of chunk c+1 overlaps the count kernel of chunk c
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

// Host DRAM read bandwidth seen by n_threads threads sweeping `bytes` bytes of host memory with 64-bit loads (a plain sum): the
// ceiling the packing lanes of gki_count_reads run against, measured on the caller's own buffer (bench.py reports e2e.host_frac).
int gki_host_read_bandwidth(const void *host, int64_t bytes, int32_t n_threads, double *gb_per_s) {
    GKI_REQUIRE(host && bytes >= 4096 && gb_per_s && !is_device_ptr(host), GKI_ERR_INVALID, "gki_host_read_bandwidth: need a host buffer of >= 4096 bytes");
    int T = n_threads > 0 ? n_threads : default_pack_threads() + 1;
    if (T > 256) T = 256;
    const uint64_t *words = (const uint64_t *)(((uintptr_t)host + 7) & ~(uintptr_t)7);
    const int64_t n_words = (bytes - 8) / 8;
    std::vector<uint64_t> sums((size_t)T * 8, 0);
    auto body = [&](int t) {
        const int64_t w0 = n_words * t / T, w1 = n_words * (t + 1) / T;
        sums[(size_t)t * 8] = sweep_words(words + w0, w1 - w0);
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> threads;
    for (int t = 1; t < T; t++) threads.emplace_back(body, t);
    body(0);
    for (auto &th : threads) th.join();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    uint64_t total = 0;
    for (int t = 0; t < T; t++) total += sums[(size_t)t * 8];
    if (total == 0x9e3779b97f4a7c15ull) fprintf(stderr, " ");   // keep the sum alive
    *gb_per_s = (double)n_words * 8.0 / secs / 1e9;
    return GKI_OK;
}

int gki_pack_reads(const uint8_t *reads, int64_t n_reads, int32_t read_len, int64_t row_stride, uint64_t *packed, int64_t *dirty_index,
                   int64_t dirty_cap, int64_t *n_clean, int64_t *n_dirty, int32_t n_threads, int32_t flags) {
    GKI_REQUIRE(n_reads >= 0 && read_len >= 1 && row_stride >= read_len && n_clean && n_dirty && dirty_cap >= 0 && (n_reads == 0 || (reads && packed)),
                GKI_ERR_INVALID, "gki_pack_reads: bad arguments");
    GKI_REQUIRE(n_reads == 0 || (!is_device_ptr(reads) && !is_device_ptr(packed)), GKI_ERR_INVALID, "gki_pack_reads: host buffers only");
    const int64_t words = (read_len + 31) / 32;
    int T = n_threads > 0 ? n_threads : default_pack_threads() + 1;
    if (T > 64) T = 64;
    if ((int64_t)T > n_reads / 4096 + 1) T = (int)(n_reads / 4096 + 1);
    std::vector<int64_t> clean((size_t)T, 0);
    std::vector<std::vector<int64_t>> dirty((size_t)T);
    auto body = [&](int t) {
        const int64_t r0 = n_reads * t / T, r1 = n_reads * (t + 1) / T;
        dirty[(size_t)t].resize((size_t)(r1 - r0 < dirty_cap ? r1 - r0 : dirty_cap));
        int64_t nd = 0;
        clean[(size_t)t] = pack_rows(reads, row_stride, nullptr, read_len, r0, r1, packed + r0 * words, nullptr, dirty[(size_t)t].data(),
                                     (int64_t)dirty[(size_t)t].size(), &nd, flags & GKI_PACK_FORCE_SCALAR);
        if ((size_t)nd < dirty[(size_t)t].size()) dirty[(size_t)t].resize((size_t)nd);
        dirty[(size_t)t].push_back(nd);   // last element: the slice's dirty count
    };
    std::vector<std::thread> threads;
    for (int t = 1; t < T; t++) threads.emplace_back(body, t);
    body(0);
    for (auto &th : threads) th.join();
    int64_t acc = 0, nd_total = 0, listed = 0;
    for (int t = 0; t < T; t++) {   // close the gaps the dirty rows left between the slices
        const int64_t r0 = n_reads * t / T;
        if (acc != r0 && clean[(size_t)t]) memmove(packed + acc * words, packed + r0 * words, (size_t)clean[(size_t)t] * words * 8);
        acc += clean[(size_t)t];
        nd_total += dirty[(size_t)t].back();
        for (size_t i = 0; i + 1 < dirty[(size_t)t].size() && listed < dirty_cap; i++)
            if (dirty_index) dirty_index[listed++] = dirty[(size_t)t][i];
    }
    *n_clean = acc;
    *n_dirty = nd_total;
    return GKI_OK;
}

int gki_count_packed_reads(gki_index_t *ix, const uint64_t *packed, int64_t n_reads, int32_t read_len, int32_t k, int32_t both_strands,
                           gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(ix && n_reads >= 0 && read_len >= 0, GKI_ERR_INVALID, "gki_count_packed_reads: bad arguments");
    GKI_REQUIRE(k >= 1 && k <= 31, GKI_ERR_INVALID, "gki_count_packed_reads: k must be in [1, 31], got %d", k);
    if (n_reads == 0 || read_len < k) return GKI_OK;
    GKI_REQUIRE(packed, GKI_ERR_INVALID, "gki_count_packed_reads: packed is NULL");
    GKI_TRY(ensure_table(ix, k, call.stream));
    const int64_t row_bytes = (int64_t)((read_len + 31) / 32) * 8;
    if (!is_device_ptr(packed) && n_reads * row_bytes >= ((int64_t)48 << 20)) {
        // a large host batch crosses the bus in chunks: the copy of chunk c+1 overlaps the count kernel of chunk c
        cudaStream_t s = call.stream;
        int64_t chunk_reads = (((int64_t)32 << 20) / row_bytes) & ~31ll;
        GKI_TRY(ensure_staging(ix, (size_t)chunk_reads * row_bytes + 16));
        int c = 0;
        for (int64_t off = 0; off < n_reads; off += chunk_reads, ++c) {
            const int bsel = c & 1;
            const int64_t cnt = n_reads - off < chunk_reads ? n_reads - off : chunk_reads;
            if (c >= 2) GKI_CUDA(cudaStreamWaitEvent(ix->copy_stream, ix->done[bsel], 0));
            GKI_CUDA(cudaMemcpyAsync(ix->stage[bsel], (const char *)packed + off * row_bytes, (size_t)(cnt * row_bytes), cudaMemcpyHostToDevice, ix->copy_stream));
            GKI_CUDA(cudaEventRecord(ix->ready[bsel], ix->copy_stream));
            GKI_CUDA(cudaStreamWaitEvent(s, ix->ready[bsel], 0));
            GKI_TRY(launch_count_packed_reads(ix, (const uint64_t *)ix->stage[bsel], cnt, read_len, k, both_strands, s));
            GKI_CUDA(cudaEventRecord(ix->done[bsel], s));
        }
        GKI_CUDA(cudaStreamSynchronize(s));
        return GKI_OK;
    }
    DevIn d;
    GKI_TRY(d.stage(packed, (size_t)(n_reads * row_bytes), call.stream));
    GKI_TRY(launch_count_packed_reads(ix, d.as<uint64_t>(), n_reads, read_len, k, both_strands, call.stream));
    return call.finish();
}

struct gki_fastx {
    gki::FastxFile *file;
};

int gki_fastx_open(const char *path, gki_fastx_t **out, int64_t *n_reads, int32_t *max_len, int32_t *format) {
    GKI_REQUIRE(path && out, GKI_ERR_INVALID, "gki_fastx_open: bad arguments");
    std::string err;
    FastxFile *f = fastx_open(path, 0, err);
    GKI_REQUIRE(f != nullptr, GKI_ERR_INVALID, "gki_fastx_open: %s", err.c_str());
    *out = new gki_fastx{f};
    if (n_reads) *n_reads = (int64_t)f->offsets.size();
    if (max_len) *max_len = f->max_len;
    if (format) *format = f->format;
    return GKI_OK;
}

int gki_fastx_lines(const gki_fastx_t *fx, int64_t *offsets, int32_t *lengths) {
    GKI_REQUIRE(fx && fx->file, GKI_ERR_INVALID, "gki_fastx_lines: bad handle");
    if (offsets) memcpy(offsets, fx->file->offsets.data(), fx->file->offsets.size() * 8);
    if (lengths) memcpy(lengths, fx->file->lengths.data(), fx->file->lengths.size() * 4);
    return GKI_OK;
}

int gki_fastx_close(gki_fastx_t *fx) {
    if (fx) {
        fastx_close(fx->file);
        delete fx;
    }
    return GKI_OK;
}

int gki_count_fastx(gki_index_t *ix, const gki_fastx_t *fx, int32_t k, int32_t both_strands, int64_t *n_kmers, gki_stream_t stream) {
    GKI_REQUIRE(ix && fx && fx->file, GKI_ERR_INVALID, "gki_count_fastx: bad arguments");
    GKI_REQUIRE(k >= 1 && k <= 31, GKI_ERR_INVALID, "gki_count_fastx: k must be in [1, 31], got %d", k);
    cudaStream_t s = (cudaStream_t)stream;
    const FastxFile &f = *fx->file;
    int64_t total = 0;
    // lines grouped by length: the fused kernel walks equal-length rows; sequencing runs have one or a handful of lengths
    std::vector<int32_t> order_len;
    std::vector<OffsetVector> groups;
    const bool uniform = !f.offsets.empty() && f.min_len == f.max_len;   // the usual case: the file's own offset array is the group
    if (uniform) {
        if (f.max_len >= k) order_len.push_back(f.max_len);
    } else {
        std::vector<int32_t> slot((size_t)f.max_len + 1, -1);
        for (size_t i = 0; i < f.offsets.size(); i++) {
            const int32_t len = f.lengths[i];
            if (len < k) continue;   // no k-mers (read_kmers.py:67-70 on a read shorter than k is documented as a deviation)
            if (slot[(size_t)len] < 0) {
                slot[(size_t)len] = (int32_t)groups.size();
                groups.emplace_back();
                order_len.push_back(len);
            }
            groups[(size_t)slot[(size_t)len]].push_back(f.offsets[i]);
        }
    }
    if (!order_len.empty()) GKI_TRY(ensure_table(ix, k, s));
    const int n_lanes = default_pack_threads();
    for (size_t g = 0; g < order_len.size(); g++) {
        const int32_t len = order_len[g];
        const OffsetVector &rows = uniform ? f.offsets : groups[g];
        total += (int64_t)rows.size() * (len - k + 1) * (both_strands ? 2 : 1);
        if (n_lanes > 0 && (int64_t)rows.size() >= 4 * UNIT_READS) {
            GKI_TRY(count_reads_host_pipeline(ix, f.data, rows.data(), (int64_t)rows.size(), len, 0, k, both_strands, s, n_lanes));
        } else {   // few lines of this length: a dense copy through the ordinary host path
            std::vector<uint8_t> dense(rows.size() * (size_t)len);
            for (size_t r = 0; r < rows.size(); r++) memcpy(dense.data() + r * (size_t)len, f.data + rows[r], (size_t)len);
            GKI_TRY(gki_count_reads(ix, dense.data(), (int64_t)rows.size(), len, len, k, both_strands, stream));
        }
    }
    if (n_kmers) *n_kmers = total;
    return GKI_OK;
}

int gki_node_counts(gki_index_t *ix, double *out, int64_t n_out, int32_t flags, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(ix && out && n_out >= 0, GKI_ERR_INVALID, "gki_node_counts: bad arguments");
    GKI_REQUIRE(n_out > ix->max_node, GKI_ERR_OVERFLOW, "gki_node_counts: n_out %lld <= max node id %lld", (long long)n_out, (long long)ix->max_node);
    DevOut o;
    GKI_TRY(o.prepare(out, (size_t)n_out * 8, call.stream));
    GKI_CUDA(cudaMemsetAsync(o.dptr, 0, (size_t)n_out * 8, call.stream));
    if (ix->table.buckets) {
        const bool wrap = (flags & GKI_COUNTS_WRAP_UINT16) != 0;
        int node_grid_mult = 16;   // CTAs dealt out as slots free up instead of one resident wave (c3: 10.9 -> 10.4 ms, c2: 0.49 -> 0.41 ms)
        if (const char *e = experiment_knob("GKI_NODE_GRID_MULT")) node_grid_mult = atoi(e) > 0 ? atoi(e) : node_grid_mult;
        const int grid = grid_for(ix->n, 256 * 4, device_info().sms * 16 * node_grid_mult);
        size_t slice_bytes = (size_t)48 << 20;   // node counts per pass: must stay in L2 (126 MB) next to the streams (measured at c3: 32 MB 18.9 ms, 48 MB 15.2, 64 MB 15.4, 96 MB 17.6)
        if (const char *e = getenv("GKI_NODE_SLICE_MB")) slice_bytes = (size_t)(atoi(e) > 0 ? atoi(e) : 0) << 20;
        if (ix->cs_slot && slice_bytes && (size_t)n_out * 8 > slice_bytes + (slice_bytes >> 1) && ix->max_node < (1ll << 28)) {
            int node_bits = 1;
            while ((1ll << node_bits) <= ix->max_node) node_bits++;
            Scratch w;
            GKI_TRY(w.alloc((size_t)ix->n * 4, call.stream));
            // ranges of 2^22 nodes (32 MB of counts), words grouped by range tile by tile (GKI_NODE_RANGED=0: the passes over all words)
            const int shift = 22;
            const int ranges = (int)(ix->max_node >> shift) + 1;
            const char *ranged_env = getenv("GKI_NODE_RANGED");
            if (ranges <= NB_MAX_RANGES && (int64_t)ix->n < (1ll << 32) && !(ranged_env && ranged_env[0] == '0')) {
                const int64_t n_tiles = (ix->n + NB_TILE - 1) / NB_TILE;
                const int tgrid = (int)std::min<int64_t>(n_tiles, (int64_t)device_info().sms * 8 * node_grid_mult);
                if (!ix->nb_seg || ix->nb_ranges != ranges || ix->nb_shift != shift || ix->nb_tiles != n_tiles) {
                    cudaFree(ix->nb_seg);
                    ix->nb_seg = nullptr;
                    const size_t cells = (size_t)ranges * n_tiles + 1;
                    Scratch hist;
                    GKI_TRY(hist.alloc(cells * 4, call.stream));
                    GKI_CUDA(cudaMemsetAsync(hist.ptr, 0, cells * 4, call.stream));
                    node_range_hist_kernel<<<tgrid, NB_THREADS, 0, call.stream>>>(ix->cs_node, ix->n, shift, ranges, n_tiles, hist.as<uint32_t>());
                    GKI_CHECK_LAUNCH();
                    GKI_CUDA(cudaMalloc((void **)&ix->nb_seg, cells * 4));
                    GKI_TRY(exclusive_scan_u32(hist.as<uint32_t>(), ix->nb_seg, (int64_t)cells, nullptr, call.stream));
                    // first word of every range: ranges + 1 values of nb_seg, n_tiles apart
                    GKI_CUDA(cudaMemcpy2DAsync(ix->nb_bounds, 4, ix->nb_seg, (size_t)n_tiles * 4, 4, ranges + 1, cudaMemcpyDeviceToHost, call.stream));
                    GKI_CUDA(cudaStreamSynchronize(call.stream));
                    ix->nb_ranges = ranges;
                    ix->nb_shift = shift;
                    ix->nb_tiles = n_tiles;
                }
                entry_weights_ranged_kernel<<<tgrid, NB_THREADS, 0, call.stream>>>(ix->table, ix->cs_slot, ix->cs_node, ix->n, ix->nb_seg, shift, ranges, n_tiles,
                                                                                     w.as<uint32_t>(), node_bits, o.as<double>(), n_out, wrap);
                GKI_CHECK_LAUNCH();
                for (int r = 0; r < ranges; r++) {
                    const int64_t lo = ix->nb_bounds[r], hi = ix->nb_bounds[r + 1];
                    if (lo >= hi) continue;
                    const int rgrid = grid_for((hi - lo) / 8 + 1, 256, device_info().sms * 16 * node_grid_mult);
                    node_counts_range_kernel<<<rgrid, 256, 0, call.stream>>>(w.as<uint32_t>(), lo, hi, node_bits, o.as<double>());
                    GKI_CHECK_LAUNCH();
                }
            } else {
            entry_weights_kernel<<<grid, 256, 0, call.stream>>>(ix->table, ix->cs_slot, ix->cs_node, ix->n, w.as<uint32_t>(), node_bits, o.as<double>(), n_out, wrap);
            GKI_CHECK_LAUNCH();
            const int64_t passes = ((int64_t)n_out * 8 + (int64_t)slice_bytes - 1) / (int64_t)slice_bytes;
            const int64_t per = (n_out + passes - 1) / passes;
            const int grid8 = grid_for(ix->n / 8 + 1, 256, device_info().sms * 16 * node_grid_mult);
            for (int64_t p = 0; p < passes; p++) {
                const int64_t lo = p * per, hi = std::min<int64_t>(n_out, lo + per);
                if (lo >= hi) break;
                node_counts_slice_kernel<<<grid8, 256, 0, call.stream>>>(w.as<uint32_t>(), ix->n, node_bits, o.as<double>(), (uint32_t)lo, (uint32_t)hi);
                GKI_CHECK_LAUNCH();
            }
            }
        } else if (ix->cs_slot) node_counts_csr_kernel<<<grid, 256, 0, call.stream>>>(ix->table, ix->cs_slot, ix->cs_node, ix->n, o.as<double>(), n_out, wrap);
        else node_counts_kernel<<<grid, 256, 0, call.stream>>>(ix->table, ix->kmers, ix->nodes, ix->n, o.as<double>(), n_out, wrap);
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
    if (ix->table.buckets) {
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
    if (ix->table.buckets) {
        query_counts_kernel<<<grid_for(nq, 256 * 2, device_info().sms * 16), 256, 0, call.stream>>>(ix->table, q.as<uint64_t>(), nq, o.as<uint32_t>());
        GKI_CHECK_LAUNCH();
    } else {
        GKI_CUDA(cudaMemsetAsync(o.dptr, 0, (size_t)nq * 4, call.stream));
    }
    GKI_TRY(o.finish(call.stream));
    return call.finish();
}

}  // extern "C"
