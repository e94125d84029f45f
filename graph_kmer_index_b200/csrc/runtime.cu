// runtime.cu -- error state, device info, host<->device staging helpers, misc C-ABI entry points.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <sys/mman.h>
#include <thread>
#include <vector>
#include "common.cuh"

namespace gki {

static thread_local char g_error[1024] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
static thread_local bool g_host_io = false;
void note_host_io() { g_host_io = true; }
CallScope::CallScope(gki_stream_t s) : stream((cudaStream_t)s) { g_host_io = false; }
int CallScope::finish() {
    if (g_host_io) {
        g_host_io = false;
        GKI_CUDA(cudaStreamSynchronize(stream));
    }
    return GKI_OK;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

const DeviceInfo &device_info() {
    static thread_local DeviceInfo info;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return info;
    if (info.device != dev) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) == cudaSuccess) {
            info.device = dev;
            info.sms = p.multiProcessorCount;
            info.l2_bytes = p.l2CacheSize;
            info.smem_optin = p.sharedMemPerBlockOptin;
            // keep freed scratch in the stream-ordered pool instead of returning it to the OS
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t thr = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
        }
    }
    return info;
}

bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int Scratch::alloc(size_t bytes, cudaStream_t s) {
    release();
    stream = s;
    if (bytes == 0) bytes = 16;
    device_info();
    GKI_CUDA(cudaMallocAsync(&ptr, bytes, s));
    return GKI_OK;
}
bool Scratch::try_alloc(size_t bytes, cudaStream_t s) {
    release();
    stream = s;
    device_info();
    if (cudaMallocAsync(&ptr, bytes ? bytes : 16, s) != cudaSuccess) {
        ptr = nullptr;
        cudaGetLastError();
        return false;
    }
    return true;
}
void Scratch::release() {
    if (ptr) {
        cudaFreeAsync(ptr, stream);
        ptr = nullptr;
    }
}

// ------------------------------------------------------------------ large pageable buffers
// cudaMemcpyAsync on pageable memory is staged by the driver on one host thread (and a fresh destination array is
// page-faulted in by that thread): a few GB/s for the multi-GB columns and tables of an index built from numpy arrays.
// Above a size threshold the copy is done here instead: worker threads move chunks between the caller's memory and
// pinned double buffers, each with its own stream, so the host-side memcpy (and the page faults) run on several cores
// while the copy engine runs.  Synchronous: returns once the data has arrived.
namespace {
struct HostCopyPool {
    std::mutex mu;                 // one parallel copy at a time per process
    int device = -1, threads = 0;
    size_t chunk = 0;
    uint8_t *pinned = nullptr;     // threads x 2 chunks
    std::vector<cudaStream_t> streams;
    std::vector<cudaEvent_t> events;   // 2 per thread
    cudaEvent_t ready = nullptr;
};
HostCopyPool g_copy_pool;

int64_t env_i64(const char *name, int64_t fallback) {
    const char *v = getenv(name);
    return (v && *v) ? atoll(v) : fallback;
}
int host_copy_threads() {   // read per call: cheap next to a copy of tens of MB, and tests switch the path with the environment
    const int64_t hw = (int64_t)std::thread::hardware_concurrency();
    const int64_t ranks = std::max<int64_t>(1, env_i64("LOCAL_WORLD_SIZE", 1));
    return (int)std::min<int64_t>(64, env_i64("GKI_HOST_COPY_THREADS", std::max<int64_t>(0, std::min<int64_t>(8, hw / ranks - 1))));
}
size_t host_copy_min_bytes() { return (size_t)env_i64("GKI_HOST_COPY_MIN_BYTES", 32ll << 20); }
bool is_pageable_host_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}
// the pool for the current device (caller holds g_copy_pool.mu)
int ensure_copy_pool(HostCopyPool &pool) {
    int dev = 0;
    GKI_CUDA(cudaGetDevice(&dev));
    const int threads = host_copy_threads();
    const size_t chunk = (size_t)std::max<int64_t>(4096, env_i64("GKI_HOST_COPY_CHUNK_BYTES", 4ll << 20));
    if (pool.pinned && pool.device == dev && pool.threads == threads && pool.chunk == chunk) return GKI_OK;
    if (pool.pinned) {
        cudaFreeHost(pool.pinned);
        for (auto st : pool.streams) cudaStreamDestroy(st);
        for (auto ev : pool.events) cudaEventDestroy(ev);
        if (pool.ready) cudaEventDestroy(pool.ready);
        pool.pinned = nullptr;
        pool.ready = nullptr;
        pool.streams.clear();
        pool.events.clear();
    }
    GKI_CUDA(cudaHostAlloc((void **)&pool.pinned, (size_t)threads * 2 * chunk, cudaHostAllocDefault));
    pool.streams.resize(threads);
    pool.events.resize(2 * threads);
    for (auto &st : pool.streams) GKI_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto &ev : pool.events) GKI_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    GKI_CUDA(cudaEventCreateWithFlags(&pool.ready, cudaEventDisableTiming));
    pool.device = dev;
    pool.threads = threads;
    pool.chunk = chunk;
    return GKI_OK;
}
}  // namespace

bool wants_parallel_copy(const void *host, size_t bytes) {
    return host_copy_threads() > 0 && bytes >= host_copy_min_bytes() && is_pageable_host_ptr(host);
}

// to_device: host -> dev, else dev -> host.  Work queued on `s` before the call is waited for; the call returns when the copy is complete.
int parallel_host_copy(void *dev, void *host, size_t bytes, bool to_device, cudaStream_t s) {
    HostCopyPool &pool = g_copy_pool;
    std::lock_guard<std::mutex> lock(pool.mu);
    GKI_TRY(ensure_copy_pool(pool));
    GKI_CUDA(cudaEventRecord(pool.ready, s));
    if (!to_device && env_i64("GKI_HOST_COPY_HUGEPAGES", 1)) {
        // the destination is usually a fresh allocation that the copy threads fault in page by page: ask for 2 MB pages on its aligned
        // interior (transparent huge pages in `madvise` mode; a refusal costs nothing)
        const uintptr_t two_mb = (uintptr_t)2 << 20;
        const uintptr_t lo = ((uintptr_t)host + two_mb - 1) & ~(two_mb - 1), hi = ((uintptr_t)host + bytes) & ~(two_mb - 1);
        if (hi > lo) madvise((void *)lo, hi - lo, MADV_HUGEPAGE);
    }
    const size_t chunk = pool.chunk;
    const int64_t n_chunks = (int64_t)((bytes + chunk - 1) / chunk);
    std::atomic<int64_t> next{0};
    std::atomic<int> failed{(int)cudaSuccess};
    auto worker = [&](int t) {
        auto check = [&](cudaError_t e) {
            if (e != cudaSuccess) {
                int none = (int)cudaSuccess;
                failed.compare_exchange_strong(none, (int)e);
            }
            return e == cudaSuccess;
        };
        if (!check(cudaSetDevice(pool.device))) return;
        cudaStream_t st = pool.streams[t];
        uint8_t *slot[2] = {pool.pinned + (size_t)(2 * t) * chunk, pool.pinned + (size_t)(2 * t + 1) * chunk};
        cudaEvent_t ev[2] = {pool.events[2 * t], pool.events[2 * t + 1]};
        if (!check(cudaStreamWaitEvent(st, pool.ready, 0))) return;
        auto span = [&](int64_t i) { return std::min(chunk, bytes - (size_t)i * chunk); };
        if (to_device) {
            bool used[2] = {false, false};
            for (int cur = 0;; cur ^= 1) {
                const int64_t i = next.fetch_add(1);
                if (i >= n_chunks || failed.load() != (int)cudaSuccess) break;
                if (used[cur] && !check(cudaEventSynchronize(ev[cur]))) break;      // the slot's previous chunk has left
                memcpy(slot[cur], (const uint8_t *)host + (size_t)i * chunk, span(i));
                if (!check(cudaMemcpyAsync((uint8_t *)dev + (size_t)i * chunk, slot[cur], span(i), cudaMemcpyHostToDevice, st))) break;
                if (!check(cudaEventRecord(ev[cur], st))) break;
                used[cur] = true;
            }
        } else {
            int64_t prev = -1;
            for (int cur = 0;; cur ^= 1) {       // chunk i travels while chunk i-1 is copied out of the other slot
                int64_t i = next.fetch_add(1);
                if (i >= n_chunks || failed.load() != (int)cudaSuccess) i = -1;
                if (i >= 0) {
                    if (!check(cudaMemcpyAsync(slot[cur], (const uint8_t *)dev + (size_t)i * chunk, span(i), cudaMemcpyDeviceToHost, st))) i = -1;
                    else if (!check(cudaEventRecord(ev[cur], st))) i = -1;
                }
                if (prev >= 0 && check(cudaEventSynchronize(ev[cur ^ 1])))
                    memcpy((uint8_t *)host + (size_t)prev * chunk, slot[cur ^ 1], span(prev));
                if (i < 0) break;
                prev = i;
            }
        }
        check(cudaStreamSynchronize(st));
    };
    std::vector<std::thread> pool_threads;
    const int n_threads = (int)std::min<int64_t>(pool.threads, n_chunks);
    for (int t = 1; t < n_threads; t++) pool_threads.emplace_back(worker, t);
    worker(0);
    for (auto &th : pool_threads) th.join();
    if (failed.load() != (int)cudaSuccess) {
        set_error("parallel host copy: %s", cudaGetErrorString((cudaError_t)failed.load()));
        return GKI_ERR_CUDA;
    }
    return GKI_OK;
}

int DevIn::stage(const void *p, size_t bytes, cudaStream_t s) {
    dptr = nullptr;
    if (!p) return GKI_OK;
    if (bytes == 0 || is_device_ptr(p)) {
        dptr = p;
        return GKI_OK;
    }
    note_host_io();
    GKI_TRY(scratch.alloc(bytes, s));
    if (wants_parallel_copy(p, bytes)) GKI_TRY(parallel_host_copy(scratch.ptr, const_cast<void *>(p), bytes, true, s));
    else GKI_CUDA(cudaMemcpyAsync(scratch.ptr, p, bytes, cudaMemcpyHostToDevice, s));
    dptr = scratch.ptr;
    return GKI_OK;
}

int DevOut::prepare(void *p, size_t nbytes, cudaStream_t s) {
    dptr = nullptr;
    host = nullptr;
    bytes = nbytes;
    if (!p) return GKI_OK;
    if (nbytes == 0 || is_device_ptr(p)) {
        dptr = p;
        return GKI_OK;
    }
    host = p;
    note_host_io();
    GKI_TRY(scratch.alloc(nbytes, s));
    dptr = scratch.ptr;
    return GKI_OK;
}
int DevOut::finish(cudaStream_t s) {
    if (host && bytes) {
        if (wants_parallel_copy(host, bytes)) GKI_TRY(parallel_host_copy(dptr, host, bytes, false, s));
        else GKI_CUDA(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, s));
    }
    return GKI_OK;
}

}  // namespace gki

extern "C" {

const char *gki_last_error(void) { return gki::g_error; }
int gki_version(void) { return 100; }
int64_t gki_launch_count(void) { return gki::g_launches.load(); }

int gki_device_count(int *count) {
    GKI_REQUIRE(count, GKI_ERR_INVALID, "count is NULL");
    *count = 0;
    GKI_CUDA(cudaGetDeviceCount(count));
    return GKI_OK;
}
int gki_set_device(int device) {
    GKI_CUDA(cudaSetDevice(device));
    return GKI_OK;
}

}  // extern "C"
