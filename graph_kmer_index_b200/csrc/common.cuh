// common.cuh -- shared device/host helpers of libgki.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/gki.h"

namespace gki {

// ------------------------------------------------------------------ errors / bookkeeping
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define GKI_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            gki::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return GKI_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define GKI_CHECK_LAUNCH()                                                                     \
    do {                                                                                       \
        gki::count_launch();                                                                   \
        GKI_CUDA(cudaGetLastError());                                                          \
    } while (0)

#define GKI_REQUIRE(cond, code, ...)                                                           \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            gki::set_error(__VA_ARGS__);                                                       \
            return code;                                                                       \
        }                                                                                      \
    } while (0)

#define GKI_TRY(expr)                                                                          \
    do {                                                                                       \
        int _s = (expr);                                                                       \
        if (_s != GKI_OK) return _s;                                                           \
    } while (0)

// Set when a call touched host memory (staged input or host output): the entry point then
// synchronises the stream before returning so that host results are complete.
struct CallScope {
    cudaStream_t stream;
    explicit CallScope(gki_stream_t s);
    int finish();   // sync if any host buffer was involved in this call
};
void note_host_io();

// ------------------------------------------------------------------ experiment knobs
// Environment variables that only exist for the sweeps under profiles/ (kernel variants, occupancy, tile sizes) are read only in a
// library built with -DGKI_EXPERIMENT_KNOBS; the shipped build answers "unset" and compiles none of the variant kernels.
// Operational and test-forcing variables (GKI_PACK_THREADS, GKI_PIPELINE_DMA, GKI_BUILD_PATH, GKI_FILTER_*, ...) use getenv directly.
#ifdef GKI_EXPERIMENT_KNOBS
#include <stdlib.h>
inline const char *experiment_knob(const char *name) { return getenv(name); }
#else
inline const char *experiment_knob(const char *) { return nullptr; }
#endif

// ------------------------------------------------------------------ device properties
struct DeviceInfo {
    int device = -1;
    int sms = 148;
    int l2_bytes = 0;
    size_t smem_optin = 0;
};
const DeviceInfo &device_info();

// ------------------------------------------------------------------ host<->device staging
// Where does a caller pointer live?
bool is_device_ptr(const void *p);
// large pageable host arrays (runtime.cu): true when `host` should go through parallel_host_copy, which moves it through pinned double
// buffers on several host threads (waits for the work queued on `s`, returns when the copy is complete)
bool wants_parallel_copy(const void *host, size_t bytes);
int parallel_host_copy(void *dev, void *host, size_t bytes, bool to_device, cudaStream_t s);

// RAII device scratch from the stream-ordered pool.
struct Scratch {
    void *ptr = nullptr;
    cudaStream_t stream = nullptr;
    ~Scratch() { release(); }
    int alloc(size_t bytes, cudaStream_t s);
    bool try_alloc(size_t bytes, cudaStream_t s);   // false (and no error recorded) when the pool cannot supply the bytes
    void release();
    template <typename T> T *as() { return (T *)ptr; }
};

// An input array as seen by kernels: either the caller's device pointer or a staged device copy.
struct DevIn {
    Scratch scratch;
    const void *dptr = nullptr;
    int stage(const void *p, size_t bytes, cudaStream_t s);   // p may be NULL -> dptr NULL
    template <typename T> const T *as() const { return (const T *)dptr; }
};

// An output array: kernels write dptr; finish() copies back to a host destination if needed.
struct DevOut {
    Scratch scratch;
    void *dptr = nullptr;
    void *host = nullptr;
    size_t bytes = 0;
    int prepare(void *p, size_t nbytes, cudaStream_t s);       // p may be NULL -> dptr NULL
    int finish(cudaStream_t s);                                // D2H (async, or complete on return for large pageable arrays); caller syncs the stream
    template <typename T> T *as() { return (T *)dptr; }
};

// ------------------------------------------------------------------ fast x % m for m < 2^32
struct FastMod {
    uint64_t m;      // modulus
    uint64_t magic;  // floor((2^64-1)/m)
};
inline FastMod make_fastmod(uint64_t m) { return FastMod{m, m ? (~0ull) / m : 0}; }

__device__ __forceinline__ uint32_t fastmod(uint64_t x, const FastMod fm) {
    // q <= floor(x/m) <= q + 2  (see DESIGN.md): at most two corrections
    uint64_t q = __umul64hi(x, fm.magic);
    uint64_t r = x - q * fm.m;
    if (r >= fm.m) r -= fm.m;
    if (r >= fm.m) r -= fm.m;
    return (uint32_t)r;
}

// ------------------------------------------------------------------ base encoding (flat_kmers.py:134-145)
// returns code in bits 0-1 and validity (is one of ACGTacgt) in bit 2
__device__ __forceinline__ uint32_t encode_base(uint32_t c) {
    uint32_t lc = c | 0x20u;
    uint32_t x = (lc >> 1) & 3u;          // a:0 c:1 g:3 t:2
    uint32_t code = x ^ (x >> 1);         // a:0 c:1 g:2 t:3
    bool valid = (lc == 'a') | (lc == 'c') | (lc == 'g') | (lc == 't');
    return valid ? (code | 4u) : 0u;
}

// reverse the order of the 32 2-bit pairs of x
__device__ __forceinline__ uint64_t reverse_pairs(uint64_t x) {
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}
// kmer_hashing.py:24-28 in O(1): complement every base, reverse the base order within k
__device__ __forceinline__ uint64_t revcomp_hash(uint64_t x, int k) { return reverse_pairs(~x) >> (2 * (32 - k)); }
// hash of the reverse-complemented window when some bases are not ACGT (they stay code 0 on both
// strands, read_kmers.py:21-26 + flat_kmers.py:134-145): v has 0b11 for every valid base
__device__ __forceinline__ uint64_t revcomp_hash_masked(uint64_t x, uint64_t v, int k) {
    return reverse_pairs(~x & v) >> (2 * (32 - k));
}
__host__ __device__ __forceinline__ uint64_t kmer_mask(int k) { return (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1ull); }

// 2k-bit window starting at base `pos` of a packed 2-bit stream (base i at bits 2*(i%32) of word i/32).
// words[] must be readable one word past the last base.
__device__ __forceinline__ uint64_t extract_window(const uint64_t *words, int pos, uint64_t mask) {
    int w = pos >> 5;
    int s = (pos & 31) << 1;
    uint64_t lo = words[w], hi = words[w + 1];
    return ((lo >> s) | ((hi << 1) << (63 - s))) & mask;
}

// ------------------------------------------------------------------ PTX: mbarrier + bulk async copy (TMA 1D)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// same with an L2 eviction policy (createpolicy): read batches stream through L2 once and should not displace
// resident structures
// asks L2 for `bytes` (multiple of 16) at a 16-byte aligned global address; no shared memory, no barrier
__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// ------------------------------------------------------------------ misc
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t rnd(uint64_t seed, uint64_t i) { return splitmix64(i + seed * 0x632BE59BD9B4E019ull); }

inline int grid_for(int64_t work_items, int per_block, int max_blocks) {
    int64_t b = (work_items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

// device-wide exclusive scan (scan.cu). T in {uint32_t, uint64_t}. In-place allowed. total (device ptr, may be NULL).
int exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total, cudaStream_t s);
int exclusive_scan_u32_to_u64(const uint32_t *in, uint64_t *out, int64_t n, uint64_t *total, cudaStream_t s);

// stable LSD radix sort (build.cu) of n packed (key << 32 | payload) u64 elements held in `a` by the low key_bits bits of
// the key; `b` and `hist` receive the scratch the sort allocated (the result *sorted points into `a` or `b`).
int radix_sort_packed(Scratch &a, Scratch &b, Scratch &hist, int64_t n, int key_bits, const unsigned long long **sorted, cudaStream_t s);

}  // namespace gki
