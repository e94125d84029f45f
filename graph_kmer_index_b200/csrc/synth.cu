// synth.cu -- deterministic synthetic workloads (bit-identical to graph_kmer_index_b200/synthetic.py) and
// the calibration micro-benchmarks (random-gather ceiling for K3, streaming copy for K1/K2).
#include "common.cuh"

namespace gki {

constexpr uint64_t SEED_GENOME = 1, SEED_NODES = 2, SEED_AF = 3, SEED_READS = 4, SEED_BASES = 5, SEED_NS = 6;
constexpr uint64_t PERM_MULT = 2654435761ull, PERM_ADD = 12345ull;

__device__ __forceinline__ uint32_t packed_base(uint64_t seed, uint64_t idx) {
    return (uint32_t)(rnd(seed, idx >> 5) >> ((idx & 31) << 1)) & 3u;
}

__global__ void synth_genome_kernel(uint8_t *__restrict__ codes, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        codes[i] = (uint8_t)packed_base(SEED_GENOME, (uint64_t)i);
}

__global__ void synth_flat_kmers_kernel(const uint8_t *__restrict__ g, int64_t n, uint64_t n_nodes, int k,
                                        uint64_t *__restrict__ hashes, uint32_t *__restrict__ nodes,
                                        uint64_t *__restrict__ ref, float *__restrict__ af) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        uint64_t p = ((uint64_t)j * PERM_MULT + PERM_ADD) % (uint64_t)n;
        uint64_t u = p >> 1;
        uint64_t h = 0;
        for (int b = 0; b < k; b++) h |= (uint64_t)__ldg(g + u + b) << (2 * b);
        hashes[j] = h;
        if (nodes) nodes[j] = (uint32_t)(rnd(SEED_NODES, p) % n_nodes);
        if (ref) ref[j] = u;
        if (af) af[j] = (float)((rnd(SEED_AF, p) & 1023ull) + 1ull) * (1.0f / 1024.0f);
    }
}

__global__ void synth_reads_kernel(const uint8_t *__restrict__ g, int64_t glen, int64_t first_read, int64_t n_reads, int L,
                                   uint32_t p_hit, uint32_t n_rate, uint8_t *__restrict__ out) {
    const int64_t total = n_reads * L;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t rl = i / L;
        int m = (int)(i - rl * L);
        uint64_t r = (uint64_t)(first_read + rl);
        uint64_t h = rnd(SEED_READS, r);
        bool from_genome = (h % 1000ull) < p_hit && glen >= L;
        uint64_t flat = r * (uint64_t)L + (uint64_t)m;
        uint32_t code;
        if (from_genome) {
            uint64_t span = (uint64_t)(glen - L + 1);
            int64_t start = (int64_t)((h >> 16) % span);
            bool strand = (h >> 12) & 1ull;
            code = strand ? 3u - __ldg(g + start + (L - 1 - m)) : __ldg(g + start + m);
        } else {
            code = packed_base(SEED_BASES, flat);
        }
        uint8_t c = (uint8_t)("ACGT"[code]);
        if (n_rate && (rnd(SEED_NS, flat) % 1000ull) < n_rate) c = 'N';
        out[i] = c;
    }
}

// mode bit 0: each gather is followed by a second one whose address depends on the loaded value (2-deep chain like
// cell -> chain); bit 1: loads carry the L2::64B fill-size qualifier (default fill on B200 is the whole 128-byte line).
// Addresses come from a 32-bit multiply-xorshift hash and a multiply-high range reduction, a handful of
// instructions per gather, so the kernel measures the memory system and not the address arithmetic.
template <bool FILL64>
__device__ __forceinline__ uint64_t gather_load(const uint64_t *p) {
    uint64_t v;
    if (FILL64) asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else v = __ldg(p);
    return v;
}
__device__ __forceinline__ uint32_t gather_slot(uint32_t x, uint32_t n_words) {
    x *= 0x9E3779B1u;
    x ^= x >> 15;
    x *= 0x85EBCA6Bu;
    x ^= x >> 13;
    return __umulhi(x, n_words);
}
template <bool FILL64>
__global__ void random_gather_kernel(const uint64_t *__restrict__ table, uint32_t n_words, int64_t n_gathers, int dependent,
                                     uint64_t *__restrict__ sink) {
    uint64_t acc = 0;
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_gathers; i += 4 * T) {
        uint64_t v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int64_t idx = i + j * T;
            v[j] = idx < n_gathers ? gather_load<FILL64>(table + gather_slot((uint32_t)idx, n_words)) : 0;
        }
        if (dependent) {
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = gather_load<FILL64>(table + gather_slot((uint32_t)v[j] + j, n_words));
        }
#pragma unroll
        for (int j = 0; j < 4; j++) acc += v[j];
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}

// ---- calibration of random stores / atomics (the measurements the K2 design rests on) ----
__global__ void scatter_store_kernel(uint4 *__restrict__ dst, uint32_t n_slots, int64_t n, int bytes) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t s = gather_slot((uint32_t)i, n_slots);
        const uint4 v = make_uint4((uint32_t)i, s, 1u, 2u);
        if (bytes == 256) {   // one 256-bit store of a 32-byte record
            asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst + 2 * (size_t)s), "l"((unsigned long long)v.x), "l"((unsigned long long)v.y),
                         "l"((unsigned long long)v.z), "l"((unsigned long long)v.w)
                         : "memory");
        } else if (bytes == 32) {
            dst[2 * (size_t)s] = v;
            dst[2 * (size_t)s + 1] = v;
        } else if (bytes == 16) {
            dst[(size_t)s] = v;
        } else {
            ((uint2 *)dst)[(size_t)s] = make_uint2(v.x, v.y);
        }
    }
}
__global__ void atomic_rank_kernel(uint32_t *__restrict__ bins, uint32_t n_bins, int64_t n, int returning, uint32_t *__restrict__ sink) {
    uint32_t acc = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t *p = bins + gather_slot((uint32_t)i, n_bins);
        if (returning) acc += atomicAdd(p, 1u);
        else atomicAdd(p, 1u);
    }
    if (acc == 0xdeadbeefu) *sink = acc;
}
__global__ void atomic_scatter_kernel(uint32_t *__restrict__ bins, uint32_t n_bins, uint32_t per_bin, uint4 *__restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t b = gather_slot((uint32_t)i, n_bins);
        const uint32_t r = atomicAdd(bins + b, 1u);
        const size_t s = (size_t)b * per_bin + (r % per_bin);
        const uint4 v = make_uint4((uint32_t)i, b, r, 2u);
        dst[2 * s] = v;
        dst[2 * s + 1] = v;
    }
}

// scattered 32-byte stores, `group` consecutive lanes writing `group` consecutive records (group = 4: one whole 128-byte line per
// request); launched with one 1024-thread CTA per SM on `ctas` SMs (dynamic shared memory keeps a second CTA off the SM)
__global__ void __launch_bounds__(1024) scatter_group_kernel(uint4 *__restrict__ dst, uint32_t n_slots, int64_t n, int group) {
    extern __shared__ unsigned char pad_smem[];
    if (n < 0) pad_smem[threadIdx.x] = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t g = (uint32_t)(i / group), r = (uint32_t)(i % group);
        const size_t s = (size_t)gather_slot(g, n_slots / group) * group + r;
        asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst + 2 * s), "l"((unsigned long long)i), "l"((unsigned long long)g),
                     "l"((unsigned long long)r), "l"(1ull)
                     : "memory");
    }
}

// append to n_bins cursors: streamed 8-byte key load -> returning atomic on the bin's cursor -> ONE 256-bit store at the cursor.
// With few enough bins the open lines of all cursors stay in L2 and leave it as whole lines (the slab scatter of build.cu).
__global__ void atomic_append256_kernel(const uint64_t *__restrict__ keys, uint32_t *__restrict__ bins, uint32_t n_bins, uint32_t per_bin,
                                        uint4 *__restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = __ldg(keys + i);
        const uint32_t b = (uint32_t)(((key >> 32) * (unsigned long long)n_bins) >> 32);
        const uint32_t r = atomicAdd(bins + b, 1u);
        const size_t s = (size_t)b * per_bin + (r < per_bin ? r : per_bin - 1);
        asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst + 2 * s), "l"(key), "l"((unsigned long long)i), "l"((unsigned long long)b),
                     "l"((unsigned long long)r)
                     : "memory");
    }
}

__global__ void fill_kernel(uint64_t *__restrict__ t, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) t[i] = splitmix64(i);
}

__global__ void copy_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

}  // namespace gki

using namespace gki;

extern "C" {

int gki_synth_genome(uint8_t *codes, int64_t length, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(length >= 0 && (length == 0 || codes), GKI_ERR_INVALID, "gki_synth_genome: bad arguments");
    if (length == 0) return GKI_OK;
    DevOut o;
    GKI_TRY(o.prepare(codes, (size_t)length, call.stream));
    synth_genome_kernel<<<grid_for(length, 256 * 4, device_info().sms * 16), 256, 0, call.stream>>>(o.as<uint8_t>(), length);
    GKI_CHECK_LAUNCH();
    GKI_TRY(o.finish(call.stream));
    return call.finish();
}

int gki_synth_flat_kmers(const uint8_t *genome_codes, int64_t n_entries, int64_t n_nodes, int32_t k, uint64_t *hashes,
                         uint32_t *nodes, uint64_t *ref_offsets, float *af, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(genome_codes && hashes && n_entries >= 1 && n_nodes >= 1 && k >= 1 && k <= 31, GKI_ERR_INVALID, "gki_synth_flat_kmers: bad arguments");
    cudaStream_t s = call.stream;
    DevIn g;
    GKI_TRY(g.stage(genome_codes, (size_t)((n_entries + 1) / 2 + k - 1), s));
    DevOut oh, on, orf, oa;
    GKI_TRY(oh.prepare(hashes, (size_t)n_entries * 8, s));
    GKI_TRY(on.prepare(nodes, (size_t)n_entries * 4, s));
    GKI_TRY(orf.prepare(ref_offsets, (size_t)n_entries * 8, s));
    GKI_TRY(oa.prepare(af, (size_t)n_entries * 4, s));
    synth_flat_kmers_kernel<<<grid_for(n_entries, 256 * 2, device_info().sms * 16), 256, 0, s>>>(
        g.as<uint8_t>(), n_entries, (uint64_t)n_nodes, k, oh.as<uint64_t>(), on.as<uint32_t>(), orf.as<uint64_t>(), oa.as<float>());
    GKI_CHECK_LAUNCH();
    GKI_TRY(oh.finish(s));
    GKI_TRY(on.finish(s));
    GKI_TRY(orf.finish(s));
    GKI_TRY(oa.finish(s));
    return call.finish();
}

int gki_synth_reads(const uint8_t *genome_codes, int64_t genome_len, int64_t first_read, int64_t n_reads, int32_t read_len,
                    int32_t p_hit_permille, int32_t n_permille, uint8_t *reads, gki_stream_t stream) {
    CallScope call(stream);
    GKI_REQUIRE(genome_codes && reads && n_reads >= 0 && read_len >= 1 && genome_len >= 1, GKI_ERR_INVALID, "gki_synth_reads: bad arguments");
    if (n_reads == 0) return GKI_OK;
    cudaStream_t s = call.stream;
    DevIn g;
    GKI_TRY(g.stage(genome_codes, (size_t)genome_len, s));
    DevOut o;
    GKI_TRY(o.prepare(reads, (size_t)n_reads * read_len, s));
    synth_reads_kernel<<<grid_for(n_reads * read_len, 256 * 4, device_info().sms * 16), 256, 0, s>>>(
        g.as<uint8_t>(), genome_len, first_read, n_reads, read_len, (uint32_t)p_hit_permille, (uint32_t)n_permille, o.as<uint8_t>());
    GKI_CHECK_LAUNCH();
    GKI_TRY(o.finish(s));
    return call.finish();
}

int gki_calibrate_random_gather(int64_t table_bytes, int64_t n_gathers, int32_t dependent_loads, float *ms) {
    GKI_REQUIRE(table_bytes >= 8 && table_bytes <= (32ll << 30) && n_gathers >= 1 && ms, GKI_ERR_INVALID,
                "gki_calibrate_random_gather: bad arguments");
    uint64_t n_words = (uint64_t)table_bytes / 8;
    const int dependent = dependent_loads & 1;
    const bool fill64 = (dependent_loads & 2) != 0;
    uint64_t *table = nullptr, *sink = nullptr;
    GKI_CUDA(cudaMalloc((void **)&table, n_words * 8));
    GKI_CUDA(cudaMalloc((void **)&sink, 8));
    int grid = device_info().sms * 8;
    fill_kernel<<<grid, 256>>>(table, n_words);
    GKI_CHECK_LAUNCH();
    cudaEvent_t a, b;
    GKI_CUDA(cudaEventCreate(&a));
    GKI_CUDA(cudaEventCreate(&b));
    auto run = [&](int64_t n) {
        if (fill64) random_gather_kernel<true><<<grid, 256>>>(table, (uint32_t)n_words, n, dependent, sink);
        else random_gather_kernel<false><<<grid, 256>>>(table, (uint32_t)n_words, n, dependent, sink);
    };
    run(n_gathers / 8 + 1);   // warm-up
    GKI_CHECK_LAUNCH();
    GKI_CUDA(cudaEventRecord(a));
    run(n_gathers);
    GKI_CHECK_LAUNCH();
    GKI_CUDA(cudaEventRecord(b));
    GKI_CUDA(cudaEventSynchronize(b));
    GKI_CUDA(cudaEventElapsedTime(ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(table);
    cudaFree(sink);
    return GKI_OK;
}

int gki_calibrate_scatter(int64_t n, int32_t mode, int64_t n_bins, float *ms) {
    GKI_REQUIRE(n >= 1 && n < (1ll << 31) && ms && mode >= 0 && mode <= 7 && n_bins >= 0 && n_bins < (1ll << 31), GKI_ERR_INVALID,
                "gki_calibrate_scatter: bad arguments");
    GKI_REQUIRE(mode < 3 || mode == 6 || n_bins >= 1, GKI_ERR_INVALID, "gki_calibrate_scatter: the atomic modes need n_bins");
    uint4 *dst = nullptr;
    uint32_t *bins = nullptr, *sink = nullptr;
    uint64_t *keys = nullptr;
    const uint32_t per_bin = mode == 5 ? (uint32_t)(n / n_bins) + 1 : mode == 7 ? (uint32_t)(((n / n_bins) * 5) / 4 + 64) : 0;
    const size_t dst_bytes = (mode == 5 || mode == 7) ? (size_t)n_bins * per_bin * 32 : (size_t)n * 32;
    if (mode <= 2 || mode >= 5) GKI_CUDA(cudaMalloc((void **)&dst, dst_bytes));
    if (mode == 7) {
        GKI_CUDA(cudaMalloc((void **)&keys, (size_t)n * 8));
        fill_kernel<<<device_info().sms * 8, 256>>>(keys, (uint64_t)n);
        GKI_CHECK_LAUNCH();
    }
    if ((mode >= 3 && mode <= 5) || mode == 7) {
        GKI_CUDA(cudaMalloc((void **)&bins, (size_t)n_bins * 4));
        GKI_CUDA(cudaMemset(bins, 0, (size_t)n_bins * 4));
    }
    GKI_CUDA(cudaMalloc((void **)&sink, 4));
    const int grid = device_info().sms * 16;
    auto run = [&]() {
        if (mode == 6) scatter_store_kernel<<<grid, 256>>>(dst, (uint32_t)n, n, 256);
        else if (mode == 0) scatter_store_kernel<<<grid, 256>>>(dst, (uint32_t)n, n, 32);
        else if (mode == 1) scatter_store_kernel<<<grid, 256>>>(dst, (uint32_t)n, n, 16);
        else if (mode == 2) scatter_store_kernel<<<grid, 256>>>(dst, (uint32_t)n, n, 8);
        else if (mode == 3) atomic_rank_kernel<<<grid, 256>>>(bins, (uint32_t)n_bins, n, 1, sink);
        else if (mode == 4) atomic_rank_kernel<<<grid, 256>>>(bins, (uint32_t)n_bins, n, 0, sink);
        else if (mode == 7) {
            cudaMemsetAsync(bins, 0, (size_t)n_bins * 4);
            atomic_append256_kernel<<<grid, 256>>>(keys, bins, (uint32_t)n_bins, per_bin, dst, n);
        } else atomic_scatter_kernel<<<grid, 256>>>(bins, (uint32_t)n_bins, per_bin, dst, n);
    };
    cudaEvent_t a, b;
    GKI_CUDA(cudaEventCreate(&a));
    GKI_CUDA(cudaEventCreate(&b));
    run();   // warm-up
    GKI_CHECK_LAUNCH();
    GKI_CUDA(cudaEventRecord(a));
    run();
    GKI_CHECK_LAUNCH();
    GKI_CUDA(cudaEventRecord(b));
    GKI_CUDA(cudaEventSynchronize(b));
    GKI_CUDA(cudaEventElapsedTime(ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(dst);
    cudaFree(bins);
    cudaFree(sink);
    cudaFree(keys);
    return GKI_OK;
}

// mode bits: group (1, 2, 4, 8 lanes per run of records); n_slots: number of 32-byte slots of the destination (small = L2-resident);
// ctas: SMs used (one 1024-thread CTA each)
int gki_calibrate_store_groups(int64_t n, int32_t group, int64_t n_slots, int32_t ctas, float *ms) {
    GKI_REQUIRE(n >= 1 && ms && group >= 1 && group <= 32 && n_slots >= group && n_slots < (1ll << 32) && ctas >= 1, GKI_ERR_INVALID,
                "gki_calibrate_store_groups: bad arguments");
    uint4 *dst = nullptr;
    GKI_CUDA(cudaMalloc((void **)&dst, (size_t)n_slots * 32));
    const size_t smem = 120 * 1024;
    GKI_CUDA(cudaFuncSetAttribute(scatter_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b;
    GKI_CUDA(cudaEventCreate(&a));
    GKI_CUDA(cudaEventCreate(&b));
    scatter_group_kernel<<<ctas, 1024, smem>>>(dst, (uint32_t)n_slots, n, group);
    GKI_CHECK_LAUNCH();
    GKI_CUDA(cudaEventRecord(a));
    scatter_group_kernel<<<ctas, 1024, smem>>>(dst, (uint32_t)n_slots, n, group);
    GKI_CHECK_LAUNCH();
    GKI_CUDA(cudaEventRecord(b));
    GKI_CUDA(cudaEventSynchronize(b));
    GKI_CUDA(cudaEventElapsedTime(ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(dst);
    return GKI_OK;
}

int gki_calibrate_copy(int64_t bytes, float *ms) {
    GKI_REQUIRE(bytes >= 16 && ms, GKI_ERR_INVALID, "gki_calibrate_copy: bad arguments");
    int64_t n = bytes / 16;
    uint4 *in = nullptr, *out = nullptr;
    GKI_CUDA(cudaMalloc((void **)&in, (size_t)n * 16));
    GKI_CUDA(cudaMalloc((void **)&out, (size_t)n * 16));
    GKI_CUDA(cudaMemset(in, 1, (size_t)n * 16));
    int grid = device_info().sms * 16;
    cudaEvent_t a, b;
    GKI_CUDA(cudaEventCreate(&a));
    GKI_CUDA(cudaEventCreate(&b));
    copy_kernel<<<grid, 256>>>(in, out, n);
    GKI_CHECK_LAUNCH();
    GKI_CUDA(cudaEventRecord(a));
    copy_kernel<<<grid, 256>>>(in, out, n);
    GKI_CHECK_LAUNCH();
    GKI_CUDA(cudaEventRecord(b));
    GKI_CUDA(cudaEventSynchronize(b));
    GKI_CUDA(cudaEventElapsedTime(ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(in);
    cudaFree(out);
    return GKI_OK;
}

}  // extern "C"
