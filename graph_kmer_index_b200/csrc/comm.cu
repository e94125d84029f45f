// comm.cu -- the multi-GPU step of the path through the C ABI: node-count vectors of the ranks are combined by
// ONE NCCL all-reduce.  Reference analogue: the parent process summing the workers' results, shared_mem.py:164-171 / cfki:222-232.
// NCCL is bound at run time (dlopen of libnccl.so.2: the library the host application already has loaded -- torch's bundled one
// under Python -- or the system's), so libgki.so carries no link-time dependency on it and single-GPU hosts never touch it.
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace gki {
namespace {
struct NcclId {
    char internal[128];   // ncclUniqueId (NCCL_UNIQUE_ID_BYTES)
};
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;

int load_nccl() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.lib) return GKI_OK;
    const char *names[] = {getenv("GKI_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *name : names) {
        if (!name || !*name) continue;
        lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    GKI_REQUIRE(lib, GKI_ERR_UNSUPPORTED, "NCCL not found (libnccl.so.2; set GKI_NCCL_LIBRARY to its path): %s", dlerror());
    NcclApi api;
    api.lib = lib;
    api.GetUniqueId = (int (*)(NcclId *))dlsym(lib, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
    api.CommDestroy = (int (*)(void *))dlsym(lib, "ncclCommDestroy");
    api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(lib, "ncclAllReduce");
    api.GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
    GKI_REQUIRE(api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString, GKI_ERR_UNSUPPORTED,
                "NCCL library lacks an expected symbol");
    g_nccl = api;
    return GKI_OK;
}
}  // namespace
}  // namespace gki

using namespace gki;

#define GKI_NCCL(expr)                                                                              \
    do {                                                                                            \
        int _r = (expr);                                                                            \
        if (_r != 0) {                                                                              \
            gki::set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
            return GKI_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

extern "C" {

int gki_nccl_unique_id(void *id128) {
    GKI_REQUIRE(id128, GKI_ERR_INVALID, "gki_nccl_unique_id: NULL buffer");
    GKI_TRY(load_nccl());
    NcclId id;
    GKI_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, sizeof(id.internal));
    return GKI_OK;
}

int gki_nccl_comm_create(const void *id128, int32_t rank, int32_t world_size, void **comm_out) {
    GKI_REQUIRE(id128 && comm_out && world_size >= 1 && rank >= 0 && rank < world_size, GKI_ERR_INVALID, "gki_nccl_comm_create: bad arguments");
    GKI_TRY(load_nccl());
    NcclId id;
    memcpy(id.internal, id128, sizeof(id.internal));
    void *comm = nullptr;
    GKI_NCCL(g_nccl.CommInitRank(&comm, world_size, id, rank));
    *comm_out = comm;
    return GKI_OK;
}

int gki_nccl_comm_destroy(void *comm) {
    if (!comm) return GKI_OK;
    GKI_TRY(load_nccl());
    GKI_NCCL(g_nccl.CommDestroy(comm));
    return GKI_OK;
}

int gki_allreduce_counts(void *nccl_comm, void *counts, int64_t n, int32_t dtype, gki_stream_t stream) {
    GKI_REQUIRE(nccl_comm && n >= 0 && (n == 0 || counts), GKI_ERR_INVALID, "gki_allreduce_counts: bad arguments");
    GKI_REQUIRE(dtype == GKI_COUNTS_FLOAT64 || dtype == GKI_COUNTS_UINT64, GKI_ERR_INVALID, "gki_allreduce_counts: dtype must be GKI_COUNTS_FLOAT64 or GKI_COUNTS_UINT64");
    GKI_REQUIRE(n == 0 || is_device_ptr(counts), GKI_ERR_INVALID, "gki_allreduce_counts: counts must be device memory");
    if (n == 0) return GKI_OK;
    GKI_TRY(load_nccl());
    const int nccl_type = dtype == GKI_COUNTS_FLOAT64 ? 8 /* ncclFloat64 */ : 5 /* ncclUint64 */;
    GKI_NCCL(g_nccl.AllReduce(counts, counts, (size_t)n, nccl_type, 0 /* ncclSum */, nccl_comm, (cudaStream_t)stream));
    return GKI_OK;
}

int gki_release_scratch(void) {
    GKI_CUDA(cudaDeviceSynchronize());
    int dev = 0;
    GKI_CUDA(cudaGetDevice(&dev));
    cudaMemPool_t pool;
    GKI_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    GKI_CUDA(cudaMemPoolTrimTo(pool, 0));
    return GKI_OK;
}

}  // extern "C"
