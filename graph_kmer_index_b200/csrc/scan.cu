// scan.cu -- device-wide exclusive prefix sum (reduce-then-scan, 3 kernels per level).
// Used by the radix sort's digit offsets (build.cu) and by the hit-list offsets (index.cu).
#include "common.cuh"

namespace gki {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename TO>
__device__ __forceinline__ TO block_exclusive_scan(TO v, TO *warp_tot, TO &block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    TO inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        TO t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        TO w = (lane < SCAN_THREADS / 32) ? warp_tot[lane] : 0;
        TO winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            TO t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = winc - w;
        if (lane == 31) warp_tot[SCAN_THREADS / 32] = winc;
    }
    __syncthreads();
    block_total = warp_tot[SCAN_THREADS / 32];
    return warp_tot[warp] + inc - v;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const TI *__restrict__ in, int64_t n, TO *__restrict__ sums) {
    __shared__ TO warp_tot[SCAN_THREADS / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    TO s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) s += (TO)in[base + i];
    TO total;
    block_exclusive_scan<TO>(s, warp_tot, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN_THREADS)
    scan_tile_apply(const TI *__restrict__ in, int64_t n, const TO *__restrict__ tile_offsets, TO *__restrict__ out) {
    __shared__ TO warp_tot[SCAN_THREADS / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    TO v[SCAN_ITEMS];
    TO s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < n) ? (TO)in[base + i] : 0;
        s += v[i];
    }
    TO total;
    TO run = block_exclusive_scan<TO>(s, warp_tot, total) + (tile_offsets ? tile_offsets[blockIdx.x] : 0);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

// single-block scan for the top level (n <= SCAN_TILE); also writes the grand total
template <typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) scan_single(TO *__restrict__ data, int64_t n, TO *__restrict__ total_out) {
    __shared__ TO warp_tot[SCAN_THREADS / 32 + 1];
    int64_t base = (int64_t)threadIdx.x * SCAN_ITEMS;
    TO v[SCAN_ITEMS];
    TO s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < n) ? data[base + i] : 0;
        s += v[i];
    }
    TO total;
    TO run = block_exclusive_scan<TO>(s, warp_tot, total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == 0 && total_out) *total_out = total;
}

template <typename TI, typename TO>
static int scan_impl(const TI *in, TO *out, int64_t n, TO *total, cudaStream_t s) {
    if (n <= 0) {
        if (total) GKI_CUDA(cudaMemsetAsync(total, 0, sizeof(TO), s));
        return GKI_OK;
    }
    int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    Scratch sums;
    GKI_TRY(sums.alloc(sizeof(TO) * (size_t)tiles, s));
    scan_tile_sums<TI, TO><<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(in, n, sums.as<TO>());
    GKI_CHECK_LAUNCH();
    if (tiles <= SCAN_TILE) {
        scan_single<TO><<<1, SCAN_THREADS, 0, s>>>(sums.as<TO>(), tiles, total);
        GKI_CHECK_LAUNCH();
    } else {
        GKI_TRY((scan_impl<TO, TO>(sums.as<TO>(), sums.as<TO>(), tiles, total, s)));
    }
    scan_tile_apply<TI, TO><<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(in, n, sums.as<TO>(), out);
    GKI_CHECK_LAUNCH();
    return GKI_OK;
}

int exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total, cudaStream_t s) {
    return scan_impl<uint32_t, uint32_t>(in, out, n, total, s);
}
int exclusive_scan_u32_to_u64(const uint32_t *in, uint64_t *out, int64_t n, uint64_t *total, cudaStream_t s) {
    return scan_impl<uint32_t, uint64_t>(in, out, n, total, s);
}

}  // namespace gki
