// index.cu -- K3: device-resident CollisionFreeKmerIndex, batched probe, per-k-mer counting, node counts.
//
// Reference semantics: collision_free_kmer_index.py:303-315 (probe: bucket = kmer % modulo, slice
// [hashes_to_index[b], +n_kmers[b]), compare full k-mers), :14-40 (CounterKmerIndex: one counter per distinct
// index k-mer, get_node_counts = bincount(nodes, weights=counter[kmers])), :210-216 (map_kmers / has_kmers),
// cython_kmer_index.pyx:47-109 (hit list).
//
// Device layout (differs from the npz layout, results do not):
//   cells[b]   = {hashes_to_index[b], n_kmers[b]} interleaved -> one 8-byte load (one 32-B sector) per probe
//                instead of two loads from two 1.8 GB tables;
//   bitmap     = 1 bit per bucket "non-empty" (modulo/8 bytes = 56.6 MB at the default modulo): it stays
//                resident in B200's 126 MB L2, so a query whose bucket is empty never goes to HBM.  Built only
//                when at most half of the buckets are occupied;
//   counts[e]  = counter of the distinct k-mer whose FIRST entry in its bucket is e (the representative);
//                a hit costs exactly one RED.ADD.  get_node_counts resolves every entry to its representative.
// The probe is bound by random 32-byte sector accesses (L2 for the bitmap, HBM for cells / chain / counter),
// not by streaming bandwidth; lanes keep NQ independent probes in flight to cover the latency.
#include <stdlib.h>
#include "reads_tile.cuh"

namespace gki {

struct IndexView {
    const uint2 *cells;
    const uint32_t *bitmap;   // bit g set <=> some bucket in [g << bitmap_shift, (g+1) << bitmap_shift) is non-empty
    const uint64_t *kmers;
    uint32_t *counts;
    FastMod fm;
    uint32_t bitmap_shift;
};

// L2 eviction-priority hints: the bitmap is the only structure with reuse, everything else streams through.
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
template <bool HINT> __device__ __forceinline__ uint32_t ld_bitmap(const uint32_t *p, uint64_t pol) {
    if (!HINT) return __ldg(p);
    uint32_t v;
    asm("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
template <bool HINT> __device__ __forceinline__ uint2 ld_cell(const uint2 *p, uint64_t pol) {
    if (!HINT) return __ldg(p);
    uint2 v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    return v;
}
template <bool HINT> __device__ __forceinline__ uint64_t ld_kmer(const uint64_t *p, uint64_t pol) {
    if (!HINT) return __ldg(p);
    uint64_t v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}

}  // namespace gki

struct gki_index {
    int device = 0;
    int64_t n = 0;
    uint64_t modulo = 0;
    gki::FastMod fm{};
    uint2 *cells = nullptr;
    uint32_t *bitmap = nullptr;
    uint32_t bitmap_shift = 0;
    size_t bitmap_bytes = 0;
    int l2_mode = 1;          // bit 0: per-load L2 eviction hints, bit 1: persisting access-policy window on the bitmap
    uint64_t *kmers = nullptr;
    uint32_t *nodes = nullptr;
    uint32_t *counts = nullptr;
    uint64_t *ref_offsets = nullptr;
    uint16_t *freq = nullptr;
    float *af = nullptr;
    int64_t max_node = -1;
    int64_t nonempty = 0;
    size_t device_bytes = 0;
    // host-buffer streaming (gki_count_reads / gki_count_kmers with host pointers)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
    void *stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;

    gki::IndexView view() const { return gki::IndexView{cells, bitmap, kmers, counts, fm, bitmap_shift}; }
};

namespace gki {

constexpr int COUNT_THREADS = 256;
constexpr int NQ = 8;   // independent probes in flight per lane

// ------------------------------------------------------------------ build of the device layout
__global__ void make_cells_kernel(const int32_t *__restrict__ h2i, const uint32_t *__restrict__ nk, uint64_t modulo,
                                  uint2 *__restrict__ cells, uint32_t *__restrict__ bitmap,
                                  unsigned long long *__restrict__ nonempty) {
    const uint64_t words = (modulo + 31) / 32;
    const uint64_t warp_global = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    unsigned long long local = 0;
    for (uint64_t w = warp_global; w < words; w += n_warps) {
        uint64_t b = w * 32 + lane;
        uint32_t n = 0;
        if (b < modulo) {
            n = __ldg(nk + b);
            cells[b] = make_uint2((uint32_t)__ldg(h2i + b), n);
        }
        uint32_t bits = __ballot_sync(0xffffffffu, n != 0);
        if (lane == 0) {
            if (bitmap) bitmap[w] = bits;
            local += __popc(bits);
        }
    }
    if (lane == 0 && local) atomicAdd(nonempty, local);
}

// coarse[g] = OR of the 2^shift fine bits [g << shift, (g+1) << shift)
__global__ void coarsen_bitmap_kernel(const uint32_t *__restrict__ fine, uint64_t fine_words, uint32_t shift,
                                      uint32_t *__restrict__ coarse, uint64_t coarse_words) {
    const uint32_t span = 1u << shift;
    const uint32_t ones = span >= 32 ? 0xffffffffu : ((1u << span) - 1u);
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < coarse_words; w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t out = 0;
        for (uint32_t i = 0; i < 32; i++) {
            uint64_t first_bit = (w * 32 + i) << shift;
            uint64_t word = first_bit >> 5;
            uint32_t v = word < fine_words ? __ldg(fine + word) : 0u;
            out |= (uint32_t)(((v >> (first_bit & 31)) & ones) != 0u) << i;
        }
        coarse[w] = out;
    }
}

__global__ void max_u32_kernel(const uint32_t *__restrict__ v, int64_t n, unsigned int *__restrict__ out) {
    unsigned int m = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = max(m, __ldg(v + i));
#pragma unroll
    for (int d = 16; d; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// ------------------------------------------------------------------ probe
// NQ independent probes per lane, staged so that the loads of one stage are all issued before the first
// use: bitmap words (L2) -> cells (HBM) -> first chain k-mer (HBM) -> rest of the chain (rare).
struct Policies {
    uint64_t last, first;
};
template <bool HINT> __device__ __forceinline__ Policies make_policies() {
    Policies p{0, 0};
    if (HINT) {
        p.last = policy_evict_last();
        p.first = policy_evict_first();
    }
    return p;
}

template <bool BITMAP, bool HINT>
__device__ __forceinline__ void probe_count_batch(const IndexView &ix, const Policies &pol, const uint64_t (&q)[NQ], uint32_t live) {
    uint32_t b[NQ];
#pragma unroll
    for (int j = 0; j < NQ; j++) b[j] = fastmod(q[j], ix.fm);
    if (BITMAP) {
        uint32_t w[NQ];
#pragma unroll
        for (int j = 0; j < NQ; j++) w[j] = ((live >> j) & 1u) ? ld_bitmap<HINT>(ix.bitmap + (b[j] >> (5 + ix.bitmap_shift)), pol.last) : 0u;
#pragma unroll
        for (int j = 0; j < NQ; j++) live &= ~((((w[j] >> ((b[j] >> ix.bitmap_shift) & 31)) & 1u) ^ 1u) << j);
        if (!live) return;
    }
    uint2 cell[NQ];
#pragma unroll
    for (int j = 0; j < NQ; j++) cell[j] = ((live >> j) & 1u) ? ld_cell<HINT>(ix.cells + b[j], pol.first) : make_uint2(0u, 0u);
#pragma unroll
    for (int j = 0; j < NQ; j++) live &= ~((uint32_t)(cell[j].y == 0u) << j);
    if (!live) return;
    uint64_t first[NQ];
#pragma unroll
    for (int j = 0; j < NQ; j++) first[j] = ((live >> j) & 1u) ? ld_kmer<HINT>(ix.kmers + cell[j].x, pol.first) : 0ull;
#pragma unroll
    for (int j = 0; j < NQ; j++) {
        if (!((live >> j) & 1u)) continue;
        if (first[j] == q[j]) {
            atomicAdd(ix.counts + cell[j].x, 1u);
            continue;
        }
        for (uint32_t e = 1; e < cell[j].y; e++) {
            if (ld_kmer<HINT>(ix.kmers + cell[j].x + e, pol.first) == q[j]) {
                atomicAdd(ix.counts + cell[j].x + e, 1u);
                break;
            }
        }
    }
}

template <bool BITMAP, bool HINT>
__global__ void __launch_bounds__(COUNT_THREADS) count_kmers_kernel(IndexView ix, const uint64_t *__restrict__ queries,
                                                                    int64_t nq) {
    const Policies pol = make_policies<HINT>();
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t i0 = g; i0 < nq; i0 += T * NQ) {
        uint64_t q[NQ];
        uint32_t live = 0;
#pragma unroll
        for (int j = 0; j < NQ; j++) {
            int64_t idx = i0 + (int64_t)j * T;
            bool ok = idx < nq;
            q[j] = ok ? __ldg(queries + idx) : 0ull;
            live |= (uint32_t)ok << j;
        }
        probe_count_batch<BITMAP, HINT>(ix, pol, q, live);
    }
}

// Fused K1 -> K3: tiles of reads are staged + packed (reads_tile.cuh); a warp owns a read, each lane takes
// NQ/2 windows and probes their forward and reverse-complement hashes.
template <bool BITMAP, bool BOTH, bool HINT>
__global__ void __launch_bounds__(COUNT_THREADS) count_reads_kernel(IndexView ix, ReadBatch b) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Policies pol = make_policies<HINT>();
    const uint64_t mask = kmer_mask(b.k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    constexpr int WPL = BOTH ? NQ / 2 : NQ;   // windows per lane per batch
    for_each_tile(b, smem_raw, [&](int64_t tile, const TileSmem &t) {
        int64_t r0 = tile * (int64_t)b.tile_reads;
        for (int r = warp; r < b.tile_reads && r0 + r < b.n_reads; r += nwarps) {
            const uint64_t *cw = t.codes + (size_t)r * b.words;
            const uint64_t *vw = t.valid + (size_t)r * b.words;
            for (int base = 0; base < b.nk; base += 32 * WPL) {
                uint64_t q[NQ];
                uint32_t live = 0;
#pragma unroll
                for (int u = 0; u < WPL; u++) {
                    int i = base + u * 32 + lane;
                    bool ok = i < b.nk;
                    int ii = ok ? i : 0;
                    uint64_t x = extract_window(cw, ii, mask);
                    if (BOTH) {
                        uint64_t v = extract_window(vw, ii, mask);
                        q[2 * u] = x;
                        q[2 * u + 1] = revcomp_hash_masked(x, v, b.k);
                        live |= (ok ? 3u : 0u) << (2 * u);
                    } else {
                        q[u] = x;
                        live |= (uint32_t)ok << u;
                    }
                }
                probe_count_batch<BITMAP, HINT>(ix, pol, q, live);
            }
        }
    });
}

// ------------------------------------------------------------------ counters -> per-entry / per-node
// counter[kmers[e]]: the representative of e is the first entry of e's bucket holding the same k-mer.
__device__ __forceinline__ uint32_t entry_weight(const IndexView &ix, int64_t e, bool wrap16) {
    uint64_t km = __ldg(ix.kmers + e);
    uint2 cell = __ldg(ix.cells + fastmod(km, ix.fm));
    uint32_t w = 0;
    for (uint32_t j = 0; j < cell.y; j++) {
        if (__ldg(ix.kmers + cell.x + j) == km) {
            w = ix.counts[cell.x + j];
            break;
        }
    }
    return wrap16 ? (w & 0xFFFFu) : w;
}

__global__ void node_counts_kernel(IndexView ix, const uint32_t *__restrict__ nodes, int64_t n, double *__restrict__ out,
                                   int64_t n_out, bool wrap16) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w = entry_weight(ix, e, wrap16);
        uint32_t node = __ldg(nodes + e);
        if (w && (int64_t)node < n_out) atomicAdd(out + node, (double)w);
    }
}

__global__ void entry_counts_kernel(IndexView ix, int64_t n, uint32_t *__restrict__ out) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = entry_weight(ix, e, false);
}

// ------------------------------------------------------------------ map_kmers / has_kmers / hit list
struct Gates {
    int32_t skip_bucket0;
    int64_t max_bucket;      // < 0: off
    int32_t max_frequency;   // < 0: off
};

// calls f(entry) for every gated match of q
template <typename F>
__device__ __forceinline__ void for_each_hit(const IndexView &ix, const uint16_t *__restrict__ freq, const Gates &g,
                                             uint64_t q, F &&f) {
    uint32_t b = fastmod(q, ix.fm);
    if (g.skip_bucket0 && b == 0) return;
    uint2 cell = __ldg(ix.cells + b);
    if (cell.y == 0) return;
    if (g.max_bucket >= 0 && (int64_t)cell.y > g.max_bucket) return;
    for (uint32_t j = 0; j < cell.y; j++) {
        uint32_t e = cell.x + j;
        if (__ldg(ix.kmers + e) != q) continue;
        if (g.max_frequency >= 0 && freq && (int32_t)__ldg(freq + e) > g.max_frequency) continue;
        f(e);
    }
}

__global__ void map_kmers_kernel(IndexView ix, const uint32_t *__restrict__ nodes, const uint16_t *__restrict__ freq, Gates g,
                                 const uint64_t *__restrict__ queries, int64_t nq, unsigned long long *__restrict__ out,
                                 int64_t n_nodes) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x)
        for_each_hit(ix, freq, g, __ldg(queries + i), [&](uint32_t e) {
            uint32_t node = __ldg(nodes + e);
            if ((int64_t)node < n_nodes) atomicAdd(out + node, 1ull);
        });
}

__global__ void has_kmers_kernel(IndexView ix, Gates g, const uint64_t *__restrict__ queries, int64_t nq,
                                 uint8_t *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
        uint8_t found = 0;
        for_each_hit(ix, nullptr, g, __ldg(queries + i), [&](uint32_t) { found = 1; });
        out[i] = found;
    }
}

__global__ void hits_count_kernel(IndexView ix, const uint16_t *__restrict__ freq, Gates g,
                                  const uint64_t *__restrict__ queries, int64_t nq, uint32_t *__restrict__ n_hits) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t c = 0;
        for_each_hit(ix, freq, g, __ldg(queries + i), [&](uint32_t) { c++; });
        n_hits[i] = c;
    }
}

__global__ void hits_fill_kernel(IndexView ix, const uint32_t *__restrict__ nodes, const uint64_t *__restrict__ ref,
                                 const uint16_t *__restrict__ freq, const float *__restrict__ af, Gates g,
                                 const uint64_t *__restrict__ queries, int64_t nq, const uint64_t *__restrict__ offsets,
                                 uint64_t *__restrict__ out, int64_t capacity) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t c = offsets[i];
        for_each_hit(ix, freq, g, __ldg(queries + i), [&](uint32_t e) {
            if ((int64_t)c < capacity) {
                out[0 * capacity + c] = __ldg(nodes + e);
                out[1 * capacity + c] = ref ? __ldg(ref + e) : 0ull;
                out[2 * capacity + c] = (uint64_t)i;
                out[3 * capacity + c] = freq ? __ldg(freq + e) : 0ull;
                // pyx:106: <uint64>(1000 * float32) -- float32 product, truncation
                out[4 * capacity + c] = af ? (uint64_t)(__fmul_rn(1000.0f, __ldg(af + e))) : 0ull;
            }
            c++;
        });
    }
}

// positions form of the hit list: entry index + query index per hit (cfki:303-315 `hit_positions + start`)
__global__ void hits_fill_entries_kernel(IndexView ix, const uint16_t *__restrict__ freq, Gates g,
                                         const uint64_t *__restrict__ queries, int64_t nq, const uint64_t *__restrict__ offsets,
                                         int64_t *__restrict__ entries, int64_t *__restrict__ qidx, int64_t capacity) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t c = offsets[i];
        for_each_hit(ix, freq, g, __ldg(queries + i), [&](uint32_t e) {
            if ((int64_t)c < capacity) {
                entries[c] = (int64_t)e;
                if (qidx) qidx[c] = i;
            }
            c++;
        });
    }
}

// counter[query] for arbitrary queries (0 when absent): `self.counter[keys]` of cfki:40 for any keys
__global__ void query_counts_kernel(IndexView ix, const uint64_t *__restrict__ queries, int64_t nq, uint32_t *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t q = __ldg(queries + i);
        uint2 cell = __ldg(ix.cells + fastmod(q, ix.fm));
        uint32_t w = 0;
        for (uint32_t j = 0; j < cell.y; j++)
            if (__ldg(ix.kmers + cell.x + j) == q) {
                w = ix.counts[cell.x + j];
                break;
            }
        out[i] = w;
    }
}

template <typename T> static int dev_alloc_copy(T **dst, const T *src, size_t count, size_t &total, cudaStream_t s) {
    *dst = nullptr;
    if (!src || count == 0) return GKI_OK;
    GKI_CUDA(cudaMalloc((void **)dst, count * sizeof(T)));
    total += count * sizeof(T);
    GKI_CUDA(cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyDefault, s));
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
    int grid = grid_for(nq, COUNT_THREADS * NQ, device_info().sms * 8);
    const bool hint = ix->l2_mode & 1;
    if (ix->bitmap) {
        if (hint) count_kmers_kernel<true, true><<<grid, COUNT_THREADS, 0, s>>>(ix->view(), dq, nq);
        else count_kmers_kernel<true, false><<<grid, COUNT_THREADS, 0, s>>>(ix->view(), dq, nq);
    } else {
        if (hint) count_kmers_kernel<false, true><<<grid, COUNT_THREADS, 0, s>>>(ix->view(), dq, nq);
        else count_kmers_kernel<false, false><<<grid, COUNT_THREADS, 0, s>>>(ix->view(), dq, nq);
    }
    GKI_CHECK_LAUNCH();
    return GKI_OK;
}

template <bool BITMAP, bool BOTH, bool HINT> static int launch_count_reads_t(gki_index *ix, const ReadBatch &b, size_t smem, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        GKI_CUDA(cudaFuncSetAttribute(count_reads_kernel<BITMAP, BOTH, HINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set = true;
    }
    int blocks_per_sm = 0;
    GKI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, count_reads_kernel<BITMAP, BOTH, HINT>, COUNT_THREADS, smem));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid_for(b.n_tiles, 1, device_info().sms * blocks_per_sm));
    cfg.blockDim = dim3(COUNT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    if (BITMAP && (ix->l2_mode & 2)) {   // keep the bitmap in the persisting L2 set-aside
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = (void *)ix->bitmap;
        attr[0].val.accessPolicyWindow.num_bytes = ix->bitmap_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    GKI_CUDA(cudaLaunchKernelEx(&cfg, count_reads_kernel<BITMAP, BOTH, HINT>, ix->view(), b));
    count_launch();
    return GKI_OK;
}
template <bool BITMAP, bool BOTH> static int launch_count_reads_h(gki_index *ix, const ReadBatch &b, size_t smem, cudaStream_t s) {
    return (ix->l2_mode & 1) ? launch_count_reads_t<BITMAP, BOTH, true>(ix, b, smem, s) : launch_count_reads_t<BITMAP, BOTH, false>(ix, b, smem, s);
}

// reads: dense device rows
static int launch_count_reads(gki_index *ix, const uint8_t *dreads, int64_t n_reads, int32_t read_len, int64_t stride,
                              int32_t k, int32_t both, cudaStream_t s) {
    ReadBatch b;
    size_t smem;
    make_read_batch(dreads, n_reads, read_len, stride, k, b, smem);
    GKI_REQUIRE(smem <= 64 * 1024, GKI_ERR_UNSUPPORTED, "gki_count_reads: read_len %d too long for the tile path", read_len);
    if (ix->bitmap) return both ? launch_count_reads_h<true, true>(ix, b, smem, s) : launch_count_reads_h<true, false>(ix, b, smem, s);
    return both ? launch_count_reads_h<false, true>(ix, b, smem, s) : launch_count_reads_h<false, false>(ix, b, smem, s);
}

}  // namespace gki

using namespace gki;

extern "C" {

int gki_index_create(const int32_t *hashes_to_index, const uint32_t *n_kmers, const uint64_t *kmers, const uint32_t *nodes,
                     const uint64_t *ref_offsets, const uint16_t *frequencies, const float *af, int64_t n, uint64_t modulo,
                     int32_t flags, gki_index_t **out, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(out, GKI_ERR_INVALID, "gki_index_create: out is NULL");
    *out = nullptr;
    GKI_REQUIRE(hashes_to_index && n_kmers && kmers && nodes, GKI_ERR_INVALID, "gki_index_create: NULL index array");
    GKI_REQUIRE(n >= 1 && n < (1ll << 31), GKI_ERR_UNSUPPORTED, "gki_index_create: need 1 <= n < 2^31 (int32 offsets, cfki:453)");
    GKI_REQUIRE(modulo >= 1 && modulo < (1ull << 32), GKI_ERR_UNSUPPORTED, "gki_index_create: need 1 <= modulo < 2^32");
    cudaStream_t s = call.stream;
    gki_index *ix = new gki_index();
    struct Guard { gki_index *p; ~Guard() { if (p) gki_index_destroy(p); } } guard{ix};
    GKI_CUDA(cudaGetDevice(&ix->device));
    ix->n = n;
    ix->modulo = modulo;
    ix->fm = make_fastmod(modulo);
    size_t total = 0;
    GKI_TRY(dev_alloc_copy(&ix->kmers, kmers, (size_t)n, total, s));
    GKI_TRY(dev_alloc_copy(&ix->nodes, nodes, (size_t)n, total, s));
    GKI_TRY(dev_alloc_copy(&ix->ref_offsets, ref_offsets, (size_t)n, total, s));
    GKI_TRY(dev_alloc_copy(&ix->freq, frequencies, (size_t)n, total, s));
    GKI_TRY(dev_alloc_copy(&ix->af, af, (size_t)n, total, s));
    GKI_CUDA(cudaMalloc((void **)&ix->counts, (size_t)n * 4));
    GKI_CUDA(cudaMemsetAsync(ix->counts, 0, (size_t)n * 4, s));
    GKI_CUDA(cudaMalloc((void **)&ix->cells, (size_t)modulo * 8));
    total += (size_t)n * 4 + (size_t)modulo * 8;
    const size_t bitmap_words = (size_t)((modulo + 31) / 32);
    const bool want_bitmap = !(flags & GKI_INDEX_NO_BITMAP);
    if (want_bitmap) GKI_CUDA(cudaMalloc((void **)&ix->bitmap, bitmap_words * 4));
    {
        DevIn h2i, nk;
        GKI_TRY(h2i.stage(hashes_to_index, (size_t)modulo * 4, s));
        GKI_TRY(nk.stage(n_kmers, (size_t)modulo * 4, s));
        Scratch counters;
        GKI_TRY(counters.alloc(16, s));
        GKI_CUDA(cudaMemsetAsync(counters.ptr, 0, 16, s));
        int grid = grid_for((int64_t)modulo, 256, device_info().sms * 16);
        make_cells_kernel<<<grid, 256, 0, s>>>(h2i.as<int32_t>(), nk.as<uint32_t>(), modulo, ix->cells, ix->bitmap,
                                               (unsigned long long *)counters.ptr);
        GKI_CHECK_LAUNCH();
        max_u32_kernel<<<grid_for(n, 256 * 8, device_info().sms * 8), 256, 0, s>>>(ix->nodes, n, (unsigned int *)counters.ptr + 2);
        GKI_CHECK_LAUNCH();
        unsigned long long host_counters[2];
        GKI_CUDA(cudaMemcpyAsync(host_counters, counters.ptr, 16, cudaMemcpyDeviceToHost, s));
        GKI_CUDA(cudaStreamSynchronize(s));
        ix->nonempty = (int64_t)host_counters[0];
        ix->max_node = (int64_t)(host_counters[1] & 0xffffffffull);
    }
    // keep the bitmap only where it pays: at most half of the buckets occupied (or forced)
    if (ix->bitmap && !(flags & GKI_INDEX_FORCE_BITMAP) && (uint64_t)ix->nonempty * 2 > modulo) {
        GKI_CUDA(cudaFree(ix->bitmap));
        ix->bitmap = nullptr;
    }
    if (const char *e = getenv("GKI_L2_MODE")) ix->l2_mode = atoi(e);
    ix->bitmap_bytes = bitmap_words * 4;
    if (ix->bitmap) {
        // Coarsen the bitmap (one bit per 2^shift buckets) until it fits the L2 budget: a filter that misses L2
        // costs a 64-byte HBM fetch per query, a coarser one only lets a few more queries through to the cells.
        size_t budget = (size_t)24 << 20;
        if (const char *e = getenv("GKI_BITMAP_MAX_MB")) budget = (size_t)atoi(e) << 20;
        uint32_t shift = 0;
        while (shift < 5 && (bitmap_words * 4 >> shift) > budget) shift++;
        if (const char *e = getenv("GKI_BITMAP_SHIFT")) shift = (uint32_t)atoi(e);
        if (shift > 5) shift = 5;
        if (shift) {
            uint64_t coarse_words = ((((modulo + ((1ull << shift) - 1)) >> shift) + 31) / 32);
            uint32_t *coarse = nullptr;
            GKI_CUDA(cudaMalloc((void **)&coarse, coarse_words * 4));
            coarsen_bitmap_kernel<<<grid_for((int64_t)coarse_words, 256, device_info().sms * 16), 256, 0, s>>>(ix->bitmap, bitmap_words, shift, coarse, coarse_words);
            GKI_CHECK_LAUNCH();
            GKI_CUDA(cudaStreamSynchronize(s));
            GKI_CUDA(cudaFree(ix->bitmap));
            ix->bitmap = coarse;
            ix->bitmap_shift = shift;
            ix->bitmap_bytes = coarse_words * 4;
        }
        total += ix->bitmap_bytes;
        if (ix->l2_mode & 2) {
            cudaDeviceProp prop;
            GKI_CUDA(cudaGetDeviceProperties(&prop, ix->device));
            GKI_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)prop.persistingL2CacheMaxSize));
        }
    }
    ix->device_bytes = total;
    guard.p = nullptr;
    *out = ix;
    return call.finish();
}

int gki_index_destroy(gki_index_t *ix) {
    if (!ix) return GKI_OK;
    cudaFree(ix->cells);
    cudaFree(ix->bitmap);
    cudaFree(ix->kmers);
    cudaFree(ix->nodes);
    cudaFree(ix->counts);
    cudaFree(ix->ref_offsets);
    cudaFree(ix->freq);
    cudaFree(ix->af);
    for (int i = 0; i < 2; i++) {
        cudaFree(ix->stage[i]);
        if (ix->ready[i]) cudaEventDestroy(ix->ready[i]);
        if (ix->done[i]) cudaEventDestroy(ix->done[i]);
    }
    if (ix->copy_stream) cudaStreamDestroy(ix->copy_stream);
    delete ix;
    return GKI_OK;
}

int gki_index_info(const gki_index_t *ix, int64_t *n, uint64_t *modulo, int64_t *max_node, int64_t *device_bytes,
                   int32_t *has_bitmap, int64_t *nonempty_buckets) {
    GKI_REQUIRE(ix, GKI_ERR_INVALID, "gki_index_info: index is NULL");
    if (n) *n = ix->n;
    if (modulo) *modulo = ix->modulo;
    if (max_node) *max_node = ix->max_node;
    if (device_bytes) *device_bytes = (int64_t)ix->device_bytes;
    if (has_bitmap) *has_bitmap = ix->bitmap != nullptr;
    if (nonempty_buckets) *nonempty_buckets = ix->nonempty;
    return GKI_OK;
}

int gki_reset_counts(gki_index_t *ix, gki_stream_t stream) {
    GKI_REQUIRE(ix, GKI_ERR_INVALID, "gki_reset_counts: index is NULL");
    GKI_CUDA(cudaMemsetAsync(ix->counts, 0, (size_t)ix->n * 4, (cudaStream_t)stream));
    return GKI_OK;
}

int gki_count_kmers(gki_index_t *ix, const uint64_t *queries, int64_t nq, gki_stream_t stream) {
    GKI_REQUIRE(ix && nq >= 0 && (nq == 0 || queries), GKI_ERR_INVALID, "gki_count_kmers: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (nq == 0) return GKI_OK;
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
    if (is_device_ptr(reads)) return launch_count_reads(ix, reads, n_reads, read_len, row_stride, k, both_strands, s);
    // host reads: rows are compacted to dense device rows (so every full tile is one TMA bulk copy) in
    // chunks; copy of chunk c+1 overlaps the count kernel of chunk c
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
    node_counts_kernel<<<grid_for(ix->n, 256 * 4, device_info().sms * 16), 256, 0, call.stream>>>(
        ix->view(), ix->nodes, ix->n, o.as<double>(), n_out, (flags & GKI_COUNTS_WRAP_UINT16) != 0);
    GKI_CHECK_LAUNCH();
    GKI_TRY(o.finish(call.stream));
    return call.finish();
}

int gki_entry_counts(gki_index_t *ix, uint32_t *out, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(ix && out, GKI_ERR_INVALID, "gki_entry_counts: bad arguments");
    DevOut o;
    GKI_TRY(o.prepare(out, (size_t)ix->n * 4, call.stream));
    entry_counts_kernel<<<grid_for(ix->n, 256 * 4, device_info().sms * 16), 256, 0, call.stream>>>(ix->view(), ix->n, o.as<uint32_t>());
    GKI_CHECK_LAUNCH();
    GKI_TRY(o.finish(call.stream));
    return call.finish();
}

int gki_map_kmers(gki_index_t *ix, const uint64_t *queries, int64_t nq, uint64_t *node_counts, int64_t n_nodes, int32_t flags,
                  int32_t max_frequency, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(ix && nq >= 0 && n_nodes >= 0 && (nq == 0 || queries) && (n_nodes == 0 || node_counts), GKI_ERR_INVALID, "gki_map_kmers: bad arguments");
    GKI_REQUIRE(max_frequency < 0 || ix->freq, GKI_ERR_INVALID, "gki_map_kmers: frequency gate needs an index created with frequencies");
    if (nq == 0 || n_nodes == 0) return GKI_OK;
    DevIn q;
    GKI_TRY(q.stage(queries, (size_t)nq * 8, call.stream));
    DevOut o;
    GKI_TRY(o.prepare(node_counts, (size_t)n_nodes * 8, call.stream));
    if (o.host) GKI_CUDA(cudaMemcpyAsync(o.dptr, node_counts, (size_t)n_nodes * 8, cudaMemcpyHostToDevice, call.stream));
    Gates g{(flags & GKI_PROBE_SKIP_BUCKET0) ? 1 : 0, -1, max_frequency};
    map_kmers_kernel<<<grid_for(nq, 256 * 2, device_info().sms * 16), 256, 0, call.stream>>>(
        ix->view(), ix->nodes, ix->freq, g, q.as<uint64_t>(), nq, (unsigned long long *)o.dptr, n_nodes);
    GKI_CHECK_LAUNCH();
    GKI_TRY(o.finish(call.stream));
    return call.finish();
}

int gki_has_kmers(gki_index_t *ix, const uint64_t *queries, int64_t nq, uint8_t *out, int32_t flags, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(ix && nq >= 0 && (nq == 0 || (queries && out)), GKI_ERR_INVALID, "gki_has_kmers: bad arguments");
    if (nq == 0) return GKI_OK;
    DevIn q;
    GKI_TRY(q.stage(queries, (size_t)nq * 8, call.stream));
    DevOut o;
    GKI_TRY(o.prepare(out, (size_t)nq, call.stream));
    Gates g{(flags & GKI_PROBE_SKIP_BUCKET0) ? 1 : 0, -1, -1};
    has_kmers_kernel<<<grid_for(nq, 256 * 2, device_info().sms * 16), 256, 0, call.stream>>>(ix->view(), g, q.as<uint64_t>(), nq, o.as<uint8_t>());
    GKI_CHECK_LAUNCH();
    GKI_TRY(o.finish(call.stream));
    return call.finish();
}

static int lookup_impl(gki_index_t *ix, const uint64_t *queries, int64_t nq, int32_t flags, int64_t max_bucket, int32_t max_frequency,
                       uint64_t *out, int64_t *entries, int64_t *qidx, int64_t capacity, int64_t *n_hits, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(ix && nq >= 0 && (nq == 0 || queries) && n_hits && capacity >= 0, GKI_ERR_INVALID, "gki_lookup_hits: bad arguments");
    GKI_REQUIRE(max_frequency < 0 || ix->freq, GKI_ERR_INVALID, "gki_lookup_hits: frequency gate needs an index created with frequencies");
    *n_hits = 0;
    if (nq == 0) return GKI_OK;
    cudaStream_t s = call.stream;
    DevIn q;
    GKI_TRY(q.stage(queries, (size_t)nq * 8, s));
    Gates g{(flags & GKI_PROBE_SKIP_BUCKET0) ? 1 : 0, max_bucket, max_frequency};
    Scratch per_query, offsets, total;
    GKI_TRY(per_query.alloc((size_t)nq * 4, s));
    GKI_TRY(offsets.alloc((size_t)nq * 8, s));
    GKI_TRY(total.alloc(8, s));
    int grid = grid_for(nq, 256 * 2, device_info().sms * 16);
    hits_count_kernel<<<grid, 256, 0, s>>>(ix->view(), ix->freq, g, q.as<uint64_t>(), nq, per_query.as<uint32_t>());
    GKI_CHECK_LAUNCH();
    GKI_TRY(exclusive_scan_u32_to_u64(per_query.as<uint32_t>(), offsets.as<uint64_t>(), nq, total.as<uint64_t>(), s));
    uint64_t host_total = 0;
    GKI_CUDA(cudaMemcpyAsync(&host_total, total.ptr, 8, cudaMemcpyDeviceToHost, s));
    GKI_CUDA(cudaStreamSynchronize(s));
    *n_hits = (int64_t)host_total;
    if ((!out && !entries) || host_total == 0) return GKI_OK;
    GKI_REQUIRE(capacity >= (int64_t)host_total, GKI_ERR_OVERFLOW, "gki_lookup_hits: capacity %lld < %llu hits", (long long)capacity, (unsigned long long)host_total);
    DevOut o, oe, oq;
    GKI_TRY(o.prepare(out, (size_t)capacity * 5 * 8, s));
    GKI_TRY(oe.prepare(entries, (size_t)capacity * 8, s));
    GKI_TRY(oq.prepare(qidx, (size_t)capacity * 8, s));
    if (out) {
        hits_fill_kernel<<<grid, 256, 0, s>>>(ix->view(), ix->nodes, ix->ref_offsets, ix->freq, ix->af, g, q.as<uint64_t>(), nq,
                                              offsets.as<uint64_t>(), o.as<uint64_t>(), capacity);
        GKI_CHECK_LAUNCH();
    }
    if (entries) {
        hits_fill_entries_kernel<<<grid, 256, 0, s>>>(ix->view(), ix->freq, g, q.as<uint64_t>(), nq, offsets.as<uint64_t>(),
                                                      oe.as<int64_t>(), oq.as<int64_t>(), capacity);
        GKI_CHECK_LAUNCH();
    }
    GKI_TRY(o.finish(s));
    GKI_TRY(oe.finish(s));
    GKI_TRY(oq.finish(s));
    GKI_CUDA(cudaStreamSynchronize(s));
    return GKI_OK;
}

int gki_lookup_hits(gki_index_t *ix, const uint64_t *queries, int64_t nq, int32_t flags, int64_t max_bucket, int32_t max_frequency,
                    uint64_t *out, int64_t capacity, int64_t *n_hits, gki_stream_t stream) {
    return lookup_impl(ix, queries, nq, flags, max_bucket, max_frequency, out, nullptr, nullptr, capacity, n_hits, stream);
}

int gki_lookup_entries(gki_index_t *ix, const uint64_t *queries, int64_t nq, int32_t flags, int64_t max_bucket, int32_t max_frequency,
                       int64_t *entries, int64_t *query_index, int64_t capacity, int64_t *n_hits, gki_stream_t stream) {
    return lookup_impl(ix, queries, nq, flags, max_bucket, max_frequency, nullptr, entries, query_index, capacity, n_hits, stream);
}

int gki_query_counts(gki_index_t *ix, const uint64_t *queries, int64_t nq, uint32_t *out, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(ix && nq >= 0 && (nq == 0 || (queries && out)), GKI_ERR_INVALID, "gki_query_counts: bad arguments");
    if (nq == 0) return GKI_OK;
    DevIn q;
    GKI_TRY(q.stage(queries, (size_t)nq * 8, call.stream));
    DevOut o;
    GKI_TRY(o.prepare(out, (size_t)nq * 4, call.stream));
    query_counts_kernel<<<grid_for(nq, 256 * 2, device_info().sms * 16), 256, 0, call.stream>>>(ix->view(), q.as<uint64_t>(), nq, o.as<uint32_t>());
    GKI_CHECK_LAUNCH();
    GKI_TRY(o.finish(call.stream));
    return call.finish();
}

}  // extern "C"
