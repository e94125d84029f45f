// finder.cu -- DenseKmerFinder on the device (BASELINE config 5).
//
// Reference: graph_kmer_index/kmer_finder.py:179-434 -- a recursive depth-first walk of the variant graph from every
// "critical" position (critical_graph_paths.py:42-104: where all paths have merged for >= k bases), carrying a rolling
// k-mer hash (kf:15-34), de-duplicating path windows through the set `_positions_treated` (kf:311-319), pruning by the
// number of variant nodes in the window (kf:383-417) and emitting one row per (k-mer, node touched by its window).
//
// Here: the searches are independent except where one overruns the next start (a critical position at offset 0 is never
// seen by the `offset + 1` test of kf:333), so starting points are grouped into chains and ONE THREAD walks one chain,
// iteratively: the recursion stack becomes an explicit stack of window snapshots in global memory, `_positions_treated`
// a global open-addressing set of 128-bit fingerprints.  Two passes over the same walk -- count rows, exclusive scan,
// write rows -- reproduce the reference's row ORDER exactly (chains are concatenated in starting-point order, a chain
// emits in depth-first order).  The interior of long nodes (the reference's _process_whole_node fast path, kf:349-381)
// is not walked: the thread reserves the rows and a second, fully parallel kernel fills them.
#include <vector>
#include "common.cuh"

namespace gki {

constexpr int F_WMAX = 80;      // window elements: k <= 31 real bases + dummy-node markers in between
constexpr int F_DISTINCT = 64;  // distinct nodes in a window
constexpr int F_DEPTH = 64;     // branch points with unexplored children along one path

enum FinderError { F_OK = 0, F_ERR_WINDOW = 1, F_ERR_DEPTH = 2, F_ERR_TABLE = 4, F_ERR_LINEAR = 8, F_ERR_DISTINCT = 16, F_ERR_DYNAMIC = 32 };

struct FinderGraph {
    const int64_t *seq_off;
    const uint8_t *seq;
    const int64_t *edge_off;
    const int32_t *edges;
    const uint8_t *is_linear;
    const double *af;
    const uint16_t *crit_index;   // critical_graph_paths.py:11-19: offset per node, 0 for nodes that are not critical
    int64_t crit_len;
    const uint8_t *store;         // only_store_nodes as a per-node flag, or NULL
    const uint8_t *force;         // nodes whose (already reduced) successor list is followed whatever max_variant_nodes says, or NULL
};

struct FinderParams {
    int32_t k, max_variant_nodes, one_node, early_stop;
    const int32_t *start_nodes;   // starting points in processing order
    const int32_t *start_offsets;
    const int64_t *chain_first;   // chain c = starting points [chain_first[c], chain_first[c+1])
    int64_t n_chains;
    unsigned long long *treated;  // 2 * treated_slots u64, zero = empty
    uint64_t treated_mask;
    unsigned char *stacks;        // per-thread frame stacks
    unsigned int *error;
};

struct Frame {
    int32_t next_edge, end_edge;  // CSR positions of the remaining children
    int32_t nonempty, wlen;
    unsigned long long hash;
    int32_t wnode[F_WMAX];
    int8_t wbase[F_WMAX];
};

struct FinderOut {
    long long *kmers;
    int32_t *nodes, *start_nodes;
    int16_t *start_offsets;
    double *af;
    // bulk jobs (interior of long nodes): node, first offset, count, first row
    long long *jobs;
    unsigned long long *n_jobs;
};

__device__ __forceinline__ unsigned long long mixa(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// insert (a, b) into the treated set; returns true if it was already there
__device__ __forceinline__ bool treated_test_and_set(const FinderParams &p, unsigned long long a, unsigned long long b) {
    if (a == 0) a = 1;
    uint64_t slot = (a ^ (b << 1)) & p.treated_mask;
    for (uint64_t tries = 0; tries <= p.treated_mask; tries++) {
        unsigned long long *e = p.treated + 2 * slot;
        unsigned long long cur = *(volatile unsigned long long *)e;
        if (cur == 0) {
            cur = atomicCAS(e, 0ull, a);
            if (cur == 0) {
                *(volatile unsigned long long *)(e + 1) = b;
                return false;
            }
        }
        if (cur == a && *(volatile unsigned long long *)(e + 1) == b) return true;
        slot = (slot + 1) & p.treated_mask;
    }
    atomicOr(p.error, (unsigned int)F_ERR_TABLE);
    return false;
}

// One chain of starting points, walked exactly like kf:254-417.  FILL == false: count rows / jobs only.
template <bool FILL>
__device__ void walk_chain(const FinderGraph &g, const FinderParams &p, int64_t chain, Frame *stack, long long row_base, long long &n_rows,
                           long long &n_jobs_local, const FinderOut &out) {
    const int k = p.k;
    const unsigned long long top_weight = 1ull << (2 * (k - 1));
    int32_t wnode[F_WMAX];
    int8_t wbase[F_WMAX];
    int32_t distinct[F_DISTINCT];
    long long rows = 0;

    for (int64_t sp = p.chain_first[chain]; sp < p.chain_first[chain + 1]; sp++) {
        const int32_t crit_node = p.start_nodes[sp], crit_offset = p.start_offsets[sp];
        int wlen = 0, wfirst = 0;      // window = w*[wfirst .. wfirst + wlen); compacted when it reaches the end of the array
        int nonempty = 0, depth = 0;
        int32_t node = crit_node;
        int64_t offset = crit_offset;
        if (!p.early_stop && offset >= k - 1) offset -= k - 1;                       // kf:229-230
        unsigned long long hash = 0;

        while (true) {
            // ---------------- search_from(node, offset, hash) ----------------
            const int64_t s0 = g.seq_off[node];
            const int64_t size = g.seq_off[node + 1] - s0;
            bool stopped = false;
            auto push = [&](int8_t b, int32_t nd) {
                if (wfirst + wlen == F_WMAX) {
                    for (int i = 0; i < wlen; i++) { wnode[i] = wnode[wfirst + i]; wbase[i] = wbase[wfirst + i]; }
                    wfirst = 0;
                }
                if (wlen == F_WMAX) { atomicOr(p.error, (unsigned int)F_ERR_WINDOW); return; }
                wnode[wfirst + wlen] = nd;
                wbase[wfirst + wlen] = b;
                wlen++;
            };
            if (offset == 0 && size == 0) push(-1, node);                            // dummy node, kf:261-265
            while (offset < size) {
                if (offset == k + 2 && size > offset + k + 1 && !p.early_stop) {      // _process_whole_node, kf:349-381
                    const long long count = size - 1 - offset;
                    if (FILL) {
                        unsigned long long j = atomicAdd(out.n_jobs, 1ull);
                        out.jobs[4 * j] = node; out.jobs[4 * j + 1] = offset; out.jobs[4 * j + 2] = count; out.jobs[4 * j + 3] = row_base + rows;
                    }
                    n_jobs_local++;
                    rows += count;
                    const int64_t first = size - 2 - (k - 1);                        // window = the k bases ending at size-2
                    hash = 0;
                    wfirst = 0; wlen = k;
                    for (int j = 0; j < k; j++) {
                        uint8_t b = g.seq[s0 + first + j];
                        hash |= (unsigned long long)b << (2 * j);
                        wbase[j] = (int8_t)b;
                        wnode[j] = node;
                    }
                    offset = size - 1;
                }
                // _get_first_base_in_path, kf:419-434
                int first_base = 0;
                if (nonempty >= k) {
                    first_base = wbase[wfirst];
                    if (wlen > 1)
                        while (wlen > 1 && wbase[wfirst + 1] == -1) { wfirst++; wlen--; }
                }
                const int base = g.seq[s0 + offset];
                if (nonempty >= k) {
                    wfirst++; wlen--;
                    hash = (hash - (unsigned long long)first_base) / 4 + (unsigned long long)base * top_weight;   // update_hash, kf:31
                } else {
                    hash += ((unsigned long long)base) << (2 * nonempty);                                         // kf:27
                }
                push((int8_t)base, node);
                nonempty++;
                // distinct nodes of the window (np.unique at kf:134, frozenset at kf:311)
                int nd = 0;
                unsigned long long fa = 0, fb = 0;
                for (int i = 0; i < wlen; i++) {
                    int32_t v = wnode[wfirst + i];
                    bool seen = false;
                    for (int j = 0; j < nd; j++) seen |= distinct[j] == v;
                    if (!seen) {
                        if (nd == F_DISTINCT) { atomicOr(p.error, (unsigned int)F_ERR_DISTINCT); break; }
                        distinct[nd++] = v;
                        fa += mixa((unsigned long long)(uint32_t)v + 0x9E3779B97F4A7C15ull);
                        fb += mixa((unsigned long long)(uint32_t)v * 0xD6E8FEB86659FD93ull + 12345ull);
                    }
                }
                const unsigned long long ka = mixa(fa ^ ((unsigned long long)(uint32_t)node << 32) ^ (unsigned long long)offset * 0x9E3779B97F4A7C15ull);
                const unsigned long long kb = mixa(fb + (unsigned long long)offset) ^ ((unsigned long long)(uint32_t)node * 0xC2B2AE3D27D4EB4Full);
                const bool at_critical = node == crit_node && offset == crit_offset;
                const bool seen_before = treated_test_and_set(p, ka, kb);
                if (!at_critical && seen_before && wlen >= k) { stopped = true; break; }                          // kf:312-317
                if (nonempty >= k) {                                                                              // _add_kmer, kf:128-168
                    // ascending node order (np.unique), minimum allele frequency over the window's nodes
                    for (int i = 1; i < nd; i++) {
                        int32_t v = distinct[i];
                        int j = i - 1;
                        while (j >= 0 && distinct[j] > v) { distinct[j + 1] = distinct[j]; j--; }
                        distinct[j + 1] = v;
                    }
                    double af = g.af[distinct[0]];
                    for (int i = 1; i < nd; i++) af = fmin(af, g.af[distinct[i]]);
                    const int n_emit = p.one_node ? 1 : nd;
                    for (int i = 0; i < n_emit; i++) {
                        if (g.store && !g.store[distinct[i]]) continue;
                        if (FILL) {
                            const long long r = row_base + rows;
                            out.kmers[r] = (long long)hash;
                            out.nodes[r] = distinct[i];
                            out.start_nodes[r] = node;
                            out.start_offsets[r] = (int16_t)offset;
                            out.af[r] = af;
                        }
                        rows++;
                    }
                    if (p.early_stop) { stopped = true; break; }
                }
                if (!(node == crit_node && offset + 1 == crit_offset) && node < g.crit_len && (int64_t)g.crit_index[node] == offset + 1) {   // kf:333-340
                    stopped = true;
                    break;
                }
                offset++;
            }
            // ---------------- _search_next_nodes, kf:383-417 ----------------
            int32_t e0 = 0, e1 = 0;
            if (!stopped) {
                e0 = (int32_t)g.edge_off[node];
                e1 = (int32_t)g.edge_off[node + 1];
                if (e1 > e0) {
                    int n_var = 0;
                    // distinct[] may be stale after a dummy node was pushed: recount on the window
                    int nd = 0;
                    for (int i = 0; i < wlen; i++) {
                        int32_t v = wnode[wfirst + i];
                        bool seen = false;
                        for (int j = 0; j < nd; j++) seen |= distinct[j] == v;
                        if (!seen && nd < F_DISTINCT) { distinct[nd++] = v; n_var += !g.is_linear[v]; }
                    }
                    const bool force_follow = g.force && g.force[node];   // kf:385-388: the successors are only_follow_nodes
                    if (!force_follow && n_var >= p.max_variant_nodes) {   // only the linear-ref continuation is allowed
                        int32_t lin = -1, n_lin = 0;
                        for (int32_t e = e0; e < e1; e++)
                            if (g.is_linear[g.edges[e]]) { lin = e; n_lin++; }
                        if (n_lin != 1) { atomicOr(p.error, (unsigned int)F_ERR_LINEAR); e1 = e0; }
                        else { e0 = lin; e1 = lin + 1; }
                    }
                }
            }
            if (e1 > e0) {
                if (e1 - e0 > 1) {            // remember the state for the siblings (kf:410-417 restores it after each child)
                    if (depth == F_DEPTH) { atomicOr(p.error, (unsigned int)F_ERR_DEPTH); break; }
                    Frame &f = stack[depth++];
                    f.next_edge = e0 + 1; f.end_edge = e1; f.nonempty = nonempty; f.wlen = wlen; f.hash = hash;
                    for (int i = 0; i < wlen; i++) { f.wnode[i] = wnode[wfirst + i]; f.wbase[i] = wbase[wfirst + i]; }
                }
                node = g.edges[e0];
                offset = 0;
                continue;
            }
            // return to the closest branch point that still has an unexplored child
            bool resumed = false;
            while (depth > 0) {
                Frame &f = stack[depth - 1];
                if (f.next_edge < f.end_edge) {
                    node = g.edges[f.next_edge++];
                    offset = 0;
                    hash = f.hash; nonempty = f.nonempty; wlen = f.wlen; wfirst = 0;
                    for (int i = 0; i < wlen; i++) { wnode[i] = f.wnode[i]; wbase[i] = f.wbase[i]; }
                    if (f.next_edge == f.end_edge) depth--;      // last child: the frame is not needed again
                    resumed = true;
                    break;
                }
                depth--;
            }
            if (!resumed) break;
        }
    }
    n_rows = rows;
}

template <bool FILL>
__global__ void __launch_bounds__(64) finder_kernel(FinderGraph g, FinderParams p, const long long *__restrict__ row_base, long long *__restrict__ rows_out,
                                                    long long *__restrict__ jobs_out, FinderOut out) {
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    Frame *stack = (Frame *)(p.stacks + (size_t)tid * sizeof(Frame) * F_DEPTH);
    for (int64_t chain = tid; chain < p.n_chains; chain += (int64_t)gridDim.x * blockDim.x) {
        long long n_rows = 0, n_jobs = 0;
        walk_chain<FILL>(g, p, chain, stack, FILL ? row_base[chain] : 0, n_rows, n_jobs, out);
        if (!FILL) { rows_out[chain] = n_rows; jobs_out[chain] = n_jobs; }
    }
}

// rows of the interior of long nodes: every window lies inside the node (kf:349-381)
__global__ void finder_bulk_kernel(FinderGraph g, int k, const long long *__restrict__ jobs, unsigned long long n_jobs, FinderOut out) {
    for (unsigned long long j = blockIdx.x; j < n_jobs; j += gridDim.x) {
        const int32_t node = (int32_t)jobs[4 * j];
        const int64_t first = jobs[4 * j + 1], count = jobs[4 * j + 2], row0 = jobs[4 * j + 3];
        const int64_t s0 = g.seq_off[node];
        const double af = g.af[node];
        for (int64_t i = threadIdx.x; i < count; i += blockDim.x) {
            const int64_t off = first + i;
            unsigned long long h = 0;
            for (int b = 0; b < k; b++) h |= (unsigned long long)g.seq[s0 + off - k + 1 + b] << (2 * b);
            out.kmers[row0 + i] = (long long)h;
            out.nodes[row0 + i] = node;
            out.start_nodes[row0 + i] = node;
            out.start_offsets[row0 + i] = (int16_t)off;
            out.af[row0 + i] = af;
        }
    }
}

__global__ void exclusive_scan_ll_single(const long long *__restrict__ in, long long *__restrict__ out, int64_t n, long long *__restrict__ total) {
    // chains are few (one per critical position): a single-thread scan is not on any hot path
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        long long run = 0;
        for (int64_t i = 0; i < n; i++) { out[i] = run; run += in[i]; }
        *total = run;
    }
}

}  // namespace gki

using namespace gki;

struct gki_finder {
    FinderGraph g{};
    FinderParams p{};
    std::vector<void *> dev;   // every device allocation, freed by gki_finder_destroy
    long long total_rows = 0, total_jobs = 0;
    long long *row_base = nullptr;
    int grid = 0;
    int64_t n_nodes = 0;
};

extern "C" {

int gki_finder_destroy(gki_finder *f) {
    if (!f) return GKI_OK;
    for (void *d : f->dev) cudaFree(d);
    delete f;
    return GKI_OK;
}

// Pass 1: upload the graph + starting points, count the rows.  *n_rows receives the number of output rows.
int gki_finder_prepare(const int64_t *seq_offsets, const uint8_t *seq, const int64_t *edge_offsets, const int32_t *edges,
                       const uint8_t *is_linear, const double *allele_frequencies, int64_t n_nodes, const uint16_t *crit_index,
                       int64_t crit_len, const uint8_t *store_flags, const uint8_t *force_follow_flags, const int32_t *start_nodes,
                       const int32_t *start_offsets, int64_t n_starts, const int64_t *chain_first, int64_t n_chains, int32_t k,
                       int32_t max_variant_nodes, int32_t one_node_per_kmer, int32_t early_stop, int64_t treated_slots, gki_finder **out, int64_t *n_rows,
                       gki_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    GKI_REQUIRE(out && n_rows && seq_offsets && edge_offsets && is_linear && allele_frequencies && n_nodes >= 1 && k >= 1 && k <= 31 &&
                    n_chains >= 0 && n_starts >= 0 && treated_slots >= 16 && (treated_slots & (treated_slots - 1)) == 0,
                GKI_ERR_INVALID, "gki_finder_prepare: bad arguments");
    *out = nullptr;
    *n_rows = 0;
    gki_finder *f = new gki_finder();
    struct Guard { gki_finder *p; ~Guard() { if (p) gki_finder_destroy(p); } } guard{f};
    f->n_nodes = n_nodes;
    int64_t h_seq_total = 0, h_edge_total = 0;
    GKI_CUDA(cudaMemcpy(&h_seq_total, seq_offsets + n_nodes, 8, cudaMemcpyDefault));
    GKI_CUDA(cudaMemcpy(&h_edge_total, edge_offsets + n_nodes, 8, cudaMemcpyDefault));
    auto upload = [&](const void *src, size_t bytes, void **dst) -> int {
        *dst = nullptr;
        if (!src) return GKI_OK;
        GKI_CUDA(cudaMalloc(dst, bytes ? bytes : 16));
        f->dev.push_back(*dst);
        if (bytes) GKI_CUDA(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyDefault, s));
        return GKI_OK;
    };
    GKI_TRY(upload(seq_offsets, (size_t)(n_nodes + 1) * 8, (void **)&f->g.seq_off));
    GKI_TRY(upload(seq, (size_t)h_seq_total, (void **)&f->g.seq));
    GKI_TRY(upload(edge_offsets, (size_t)(n_nodes + 1) * 8, (void **)&f->g.edge_off));
    GKI_TRY(upload(edges, (size_t)h_edge_total * 4, (void **)&f->g.edges));
    GKI_TRY(upload(is_linear, (size_t)n_nodes, (void **)&f->g.is_linear));
    GKI_TRY(upload(allele_frequencies, (size_t)n_nodes * 8, (void **)&f->g.af));
    GKI_TRY(upload(crit_index, (size_t)crit_len * 2, (void **)&f->g.crit_index));
    f->g.crit_len = crit_index ? crit_len : 0;
    GKI_TRY(upload(store_flags, (size_t)n_nodes, (void **)&f->g.store));
    GKI_TRY(upload(force_follow_flags, (size_t)n_nodes, (void **)&f->g.force));
    GKI_TRY(upload(start_nodes, (size_t)n_starts * 4, (void **)&f->p.start_nodes));
    GKI_TRY(upload(start_offsets, (size_t)n_starts * 4, (void **)&f->p.start_offsets));
    GKI_TRY(upload(chain_first, (size_t)(n_chains + 1) * 8, (void **)&f->p.chain_first));
    f->p.k = k;
    f->p.max_variant_nodes = max_variant_nodes;
    f->p.one_node = one_node_per_kmer;
    f->p.early_stop = early_stop;
    f->p.n_chains = n_chains;
    f->p.treated_mask = (uint64_t)treated_slots - 1;
    if (n_chains == 0) {
        guard.p = nullptr;
        *out = f;
        return GKI_OK;
    }
    const int threads = 64;
    int64_t want_blocks = (n_chains + threads - 1) / threads;
    int max_blocks = device_info().sms * 8;
    f->grid = (int)(want_blocks < max_blocks ? want_blocks : max_blocks);
    void *tmp;
    GKI_CUDA(cudaMalloc(&tmp, (size_t)treated_slots * 16));
    f->dev.push_back(tmp);
    f->p.treated = (unsigned long long *)tmp;
    GKI_CUDA(cudaMalloc(&tmp, (size_t)f->grid * threads * sizeof(Frame) * F_DEPTH));
    f->dev.push_back(tmp);
    f->p.stacks = (unsigned char *)tmp;
    GKI_CUDA(cudaMalloc(&tmp, 16));
    f->dev.push_back(tmp);
    f->p.error = (unsigned int *)tmp;
    GKI_CUDA(cudaMemsetAsync(f->p.error, 0, 16, s));
    GKI_CUDA(cudaMemsetAsync(f->p.treated, 0, (size_t)treated_slots * 16, s));
    long long *rows, *jobs, *totals;
    GKI_CUDA(cudaMalloc((void **)&rows, (size_t)n_chains * 8));
    f->dev.push_back(rows);
    GKI_CUDA(cudaMalloc((void **)&jobs, (size_t)n_chains * 8));
    f->dev.push_back(jobs);
    GKI_CUDA(cudaMalloc((void **)&f->row_base, (size_t)n_chains * 8));
    f->dev.push_back(f->row_base);
    GKI_CUDA(cudaMalloc((void **)&totals, 16));
    f->dev.push_back(totals);
    FinderOut none{};
    finder_kernel<false><<<f->grid, threads, 0, s>>>(f->g, f->p, nullptr, rows, jobs, none);
    GKI_CHECK_LAUNCH();
    exclusive_scan_ll_single<<<1, 1, 0, s>>>(rows, f->row_base, n_chains, totals);
    GKI_CHECK_LAUNCH();
    exclusive_scan_ll_single<<<1, 1, 0, s>>>(jobs, jobs, n_chains, totals + 1);
    GKI_CHECK_LAUNCH();
    long long h_totals[2];
    unsigned int h_err = 0;
    GKI_CUDA(cudaMemcpyAsync(h_totals, totals, 16, cudaMemcpyDeviceToHost, s));
    GKI_CUDA(cudaMemcpyAsync(&h_err, f->p.error, 4, cudaMemcpyDeviceToHost, s));
    GKI_CUDA(cudaStreamSynchronize(s));
    GKI_REQUIRE(!(h_err & F_ERR_LINEAR), GKI_ERR_INVALID, "Not 1 linear ref next nodes (the reference asserts the same, kmer_finder.py:403)");
    GKI_REQUIRE(h_err == 0, GKI_ERR_UNSUPPORTED, "finder: graph exceeds a fixed limit (error mask %u: 1 window > %d elements, 2 more than %d open branch "
                "points, 4 treated-set full, 16 more than %d nodes in a window)", h_err, F_WMAX, F_DEPTH, F_DISTINCT);
    f->total_rows = h_totals[0];
    f->total_jobs = h_totals[1];
    *n_rows = f->total_rows;
    guard.p = nullptr;
    *out = f;
    return GKI_OK;
}

// Pass 2: write the rows (arrays of *n_rows elements from gki_finder_prepare), in the reference's order.
int gki_finder_fill(gki_finder *f, int64_t *kmers, int32_t *nodes, int32_t *start_nodes, int16_t *start_offsets, double *allele_frequencies,
                    gki_stream_t stream) {
    CallScope call(stream);
    cudaStream_t s = call.stream;
    GKI_REQUIRE(f, GKI_ERR_INVALID, "gki_finder_fill: finder is NULL");
    if (f->total_rows == 0 || f->p.n_chains == 0) return GKI_OK;
    GKI_REQUIRE(kmers && nodes && start_nodes && start_offsets && allele_frequencies, GKI_ERR_INVALID, "gki_finder_fill: NULL output");
    DevOut o_k, o_n, o_sn, o_so, o_af;
    const size_t n = (size_t)f->total_rows;
    GKI_TRY(o_k.prepare(kmers, n * 8, s));
    GKI_TRY(o_n.prepare(nodes, n * 4, s));
    GKI_TRY(o_sn.prepare(start_nodes, n * 4, s));
    GKI_TRY(o_so.prepare(start_offsets, n * 2, s));
    GKI_TRY(o_af.prepare(allele_frequencies, n * 8, s));
    Scratch jobs, n_jobs;
    GKI_TRY(jobs.alloc((size_t)(f->total_jobs + 1) * 32, s));
    GKI_TRY(n_jobs.alloc(8, s));
    GKI_CUDA(cudaMemsetAsync(n_jobs.ptr, 0, 8, s));
    GKI_CUDA(cudaMemsetAsync(f->p.treated, 0, (size_t)(f->p.treated_mask + 1) * 16, s));
    FinderOut out{(long long *)o_k.dptr, o_n.as<int32_t>(), o_sn.as<int32_t>(), o_so.as<int16_t>(), o_af.as<double>(), jobs.as<long long>(),
                  n_jobs.as<unsigned long long>()};
    finder_kernel<true><<<f->grid, 64, 0, s>>>(f->g, f->p, f->row_base, nullptr, nullptr, out);
    GKI_CHECK_LAUNCH();
    if (f->total_jobs > 0) {
        int grid = (int)(f->total_jobs < device_info().sms * 8 ? f->total_jobs : device_info().sms * 8);
        finder_bulk_kernel<<<grid, 128, 0, s>>>(f->g, f->p.k, jobs.as<long long>(), (unsigned long long)f->total_jobs, out);
        GKI_CHECK_LAUNCH();
    }
    GKI_TRY(o_k.finish(s));
    GKI_TRY(o_n.finish(s));
    GKI_TRY(o_sn.finish(s));
    GKI_TRY(o_so.finish(s));
    GKI_TRY(o_af.finish(s));
    GKI_CUDA(cudaStreamSynchronize(s));
    return GKI_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// CriticalGraphPaths.from_graph (critical_graph_paths.py:42-104): a walk along the linear reference keeping the branch
// depth and the bases since the last join.  Inherently sequential (each step follows an edge): one thread per chromosome.
namespace gki {
__global__ void critical_paths_kernel(const int64_t *__restrict__ seq_off, const int64_t *__restrict__ edge_off, const int32_t *__restrict__ edges,
                                      const uint8_t *__restrict__ is_linear, const int32_t *__restrict__ n_in_edges,
                                      const int64_t *__restrict__ chrom_starts, int n_chrom, int k, uint32_t *__restrict__ out_nodes,
                                      uint16_t *__restrict__ out_offsets, int64_t capacity, unsigned long long *__restrict__ count,
                                      unsigned int *__restrict__ error) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    unsigned long long n = 0;
    for (int c = 0; c < n_chrom; c++) {      // chromosomes in order: the reference appends to one list
        int64_t current = chrom_starts[c];
        long long depth = 0, bp = 0;
        while (true) {
            long long prev_depth = depth;
            depth -= n_in_edges[current];
            if (prev_depth > 1 && depth == 0) bp = 0;
            long long size = seq_off[current + 1] - seq_off[current];
            if (depth == 0 && size != 0 && bp <= k && bp + size >= k) {
                if ((int64_t)n < capacity) {
                    out_nodes[n] = (uint32_t)current;
                    out_offsets[n] = (uint16_t)(k - bp - 1);
                }
                n++;
            }
            int64_t e0 = edge_off[current], e1 = edge_off[current + 1];
            depth += e1 - e0;
            if (e1 == e0) break;
            if (e1 - e0 == 1) {
                bp += size;
                current = edges[e0];
            } else {
                int64_t lin = -1;
                int n_lin = 0;
                for (int64_t e = e0; e < e1; e++)
                    if (is_linear[edges[e]]) { lin = edges[e]; n_lin++; }
                if (n_lin != 1) { *error = 1u; *count = n; return; }
                current = lin;
            }
        }
    }
    *count = n;
}
}  // namespace gki

extern "C" int gki_critical_paths(const int64_t *seq_offsets, const int64_t *edge_offsets, const int32_t *edges, const uint8_t *is_linear,
                                  const int32_t *n_in_edges, int64_t n_nodes, const int64_t *chromosome_start_nodes, int32_t n_chromosomes,
                                  int32_t k, uint32_t *nodes_out, uint16_t *offsets_out, int64_t capacity, int64_t *n_out,
                                  gki_stream_t stream) {
    using namespace gki;
    CallScope call(stream);
    cudaStream_t s = call.stream;
    GKI_REQUIRE(seq_offsets && edge_offsets && is_linear && n_in_edges && chromosome_start_nodes && n_out && n_nodes >= 1 && capacity >= 0,
                GKI_ERR_INVALID, "gki_critical_paths: bad arguments");
    int64_t n_edges = 0;
    GKI_CUDA(cudaMemcpy(&n_edges, edge_offsets + n_nodes, 8, cudaMemcpyDefault));
    DevIn so, eo, ed, lin, nin, cs;
    GKI_TRY(so.stage(seq_offsets, (size_t)(n_nodes + 1) * 8, s));
    GKI_TRY(eo.stage(edge_offsets, (size_t)(n_nodes + 1) * 8, s));
    GKI_TRY(ed.stage(edges, (size_t)n_edges * 4, s));
    GKI_TRY(lin.stage(is_linear, (size_t)n_nodes, s));
    GKI_TRY(nin.stage(n_in_edges, (size_t)n_nodes * 4, s));
    GKI_TRY(cs.stage(chromosome_start_nodes, (size_t)n_chromosomes * 8, s));
    DevOut on, oo;
    GKI_TRY(on.prepare(nodes_out, (size_t)capacity * 4, s));
    GKI_TRY(oo.prepare(offsets_out, (size_t)capacity * 2, s));
    Scratch counters;
    GKI_TRY(counters.alloc(16, s));
    GKI_CUDA(cudaMemsetAsync(counters.ptr, 0, 16, s));
    critical_paths_kernel<<<1, 32, 0, s>>>(so.as<int64_t>(), eo.as<int64_t>(), ed.as<int32_t>(), lin.as<uint8_t>(), nin.as<int32_t>(), cs.as<int64_t>(),
                                           n_chromosomes, k, on.as<uint32_t>(), oo.as<uint16_t>(), capacity, (unsigned long long *)counters.ptr,
                                           (unsigned int *)counters.ptr + 2);
    GKI_CHECK_LAUNCH();
    unsigned long long h[2];
    GKI_CUDA(cudaMemcpyAsync(h, counters.ptr, 16, cudaMemcpyDeviceToHost, s));
    GKI_TRY(on.finish(s));
    GKI_TRY(oo.finish(s));
    GKI_CUDA(cudaStreamSynchronize(s));
    GKI_REQUIRE((h[1] & 0xffffffffull) == 0, GKI_ERR_INVALID, "Did not find 1 next node on the linear reference (critical_graph_paths.py:95-99 raises too)");
    *n_out = (int64_t)h[0];
    return GKI_OK;
}
