// runtime.cu -- error state, device info, host<->device staging helpers, misc C-ABI entry points.
#include <stdarg.h>
#include <atomic>
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
void Scratch::release() {
    if (ptr) {
        cudaFreeAsync(ptr, stream);
        ptr = nullptr;
    }
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
    GKI_CUDA(cudaMemcpyAsync(scratch.ptr, p, bytes, cudaMemcpyHostToDevice, s));
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
    if (host && bytes) GKI_CUDA(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, s));
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
