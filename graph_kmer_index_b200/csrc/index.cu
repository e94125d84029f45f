// index.cu -- device-resident CollisionFreeKmerIndex in the reference's bucket layout + the batched lookups
// that need entry positions: map_kmers, has_kmers, hit lists.  (Counting lives in count.cu.)
//
// Reference semantics: collision_free_kmer_index.py:303-315 (probe: bucket = kmer % modulo, slice
// [hashes_to_index[b], +n_kmers[b]), compare full k-mers), :210-216 (map_kmers / has_kmers),
// :354-391 (multi-k-mer loops), cython_kmer_index.pyx:47-109 (hit list).
#include "index.cuh"

namespace gki {

// ------------------------------------------------------------------ build of the device layout
__global__ void make_cells_kernel(const int32_t *__restrict__ h2i, const uint32_t *__restrict__ nk, uint64_t modulo,
                                  uint2 *__restrict__ cells, unsigned long long *__restrict__ nonempty) {
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
        local += __popc(__ballot_sync(0xffffffffu, n != 0));
    }
    if (lane == 0 && local) atomicAdd(nonempty, local);
}

__global__ void index_maxima_kernel(const uint32_t *__restrict__ nodes, const uint64_t *__restrict__ kmers, int64_t n,
                                    unsigned int *__restrict__ max_node, unsigned long long *__restrict__ max_kmer) {
    unsigned int m = 0;
    unsigned long long mk = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        m = max(m, __ldg(nodes + i));
        mk = max(mk, (unsigned long long)__ldg(kmers + i));
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
        mk = max(mk, __shfl_xor_sync(0xffffffffu, mk, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(max_node, m);
        atomicMax(max_kmer, mk);
    }
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

template <typename T> static int dev_alloc_copy(T **dst, const T *src, size_t count, size_t &total, cudaStream_t s) {
    *dst = nullptr;
    if (!src || count == 0) return GKI_OK;
    GKI_CUDA(cudaMalloc((void **)dst, count * sizeof(T)));
    total += count * sizeof(T);
    if (!is_device_ptr(src) && wants_parallel_copy(src, count * sizeof(T)))   // multi-GB tables of an index held in numpy arrays
        GKI_TRY(parallel_host_copy(*dst, const_cast<T *>(src), count * sizeof(T), true, s));
    else
        GKI_CUDA(cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyDefault, s));
    return GKI_OK;
}

}  // namespace gki

using namespace gki;

extern "C" {

int gki_index_create(const int32_t *hashes_to_index, const uint32_t *n_kmers, const uint64_t *kmers, const uint32_t *nodes,
                     const uint64_t *ref_offsets, const uint16_t *frequencies, const float *af, int64_t n, uint64_t modulo,
                     int32_t flags, gki_index_t **out, gki_stream_t stream) {
    CallScope call(stream);
    (void)flags;
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
    GKI_CUDA(cudaMalloc((void **)&ix->cells, (size_t)modulo * 8));
    total += (size_t)modulo * 8;
    {
        DevIn h2i, nk;
        GKI_TRY(h2i.stage(hashes_to_index, (size_t)modulo * 4, s));
        GKI_TRY(nk.stage(n_kmers, (size_t)modulo * 4, s));
        Scratch counters;   // [0] non-empty buckets (u64), [1] max k-mer (u64), [2] max node (u32)
        GKI_TRY(counters.alloc(32, s));
        GKI_CUDA(cudaMemsetAsync(counters.ptr, 0, 32, s));
        unsigned long long *c64 = (unsigned long long *)counters.ptr;
        int grid = grid_for((int64_t)modulo, 256, device_info().sms * 16);
        make_cells_kernel<<<grid, 256, 0, s>>>(h2i.as<int32_t>(), nk.as<uint32_t>(), modulo, ix->cells, c64);
        GKI_CHECK_LAUNCH();
        index_maxima_kernel<<<grid_for(n, 256 * 8, device_info().sms * 8), 256, 0, s>>>(ix->nodes, ix->kmers, n, (unsigned int *)(c64 + 2), c64 + 1);
        GKI_CHECK_LAUNCH();
        unsigned long long host_counters[3];
        GKI_CUDA(cudaMemcpyAsync(host_counters, counters.ptr, 24, cudaMemcpyDeviceToHost, s));
        GKI_CUDA(cudaStreamSynchronize(s));
        ix->nonempty = (int64_t)host_counters[0];
        ix->max_kmer = host_counters[1];
        ix->max_node = (int64_t)(host_counters[2] & 0xffffffffull);
    }
    ix->device_bytes = total;
    guard.p = nullptr;
    *out = ix;
    return call.finish();
}

int gki_index_destroy(gki_index_t *ix) {
    if (!ix) return GKI_OK;
    destroy_count_table(ix);
    cudaFree(ix->cells);
    cudaFree(ix->kmers);
    cudaFree(ix->nodes);
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
    if (device_bytes) *device_bytes = (int64_t)(ix->device_bytes + ix->table_bytes + ix->filter_bytes);
    if (has_bitmap) *has_bitmap = ix->table.filter != nullptr;
    if (nonempty_buckets) *nonempty_buckets = ix->nonempty;
    return GKI_OK;
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

}  // extern "C"
