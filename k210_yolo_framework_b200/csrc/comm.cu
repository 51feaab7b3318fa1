// Multi-GPU collation of the detection records: ONE ncclAllGather per batch, behind the C-ABI.
//
// The reference is single-process (keras_inference.py:12-17); images shard across ranks and the only exchange on the
// path is the gather of the fixed-size per-image record blocks.  The library owns its communicator (k2y_comm), built
// from a 128-byte NCCL unique id that rank 0 creates and the host side distributes (torch.distributed broadcast, MPI, a
// file — whatever launched the ranks).  NCCL is bound at run time (dlopen of libnccl.so.2: the copy the process already
// has loaded, e.g. torch's, is reused; K2Y_NCCL_LIB overrides the name), so the library still loads on a box without
// NCCL and single-GPU users never touch it.
#include <dlfcn.h>
#include <stdlib.h>

#include <mutex>

#include "common.h"

namespace {

typedef int nccl_result_t;  // ncclResult_t, ncclSuccess == 0
typedef void *nccl_comm_t;  // ncclComm_t
struct nccl_unique_id {
    char internal[128];     // ncclUniqueId (NCCL_UNIQUE_ID_BYTES)
};
constexpr int NCCL_INT8 = 0;  // ncclInt8 / ncclChar

struct NcclApi {
    void *handle = nullptr;
    nccl_result_t (*GetUniqueId)(nccl_unique_id *) = nullptr;
    nccl_result_t (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
    nccl_result_t (*CommDestroy)(nccl_comm_t) = nullptr;
    nccl_result_t (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(nccl_result_t) = nullptr;
    nccl_result_t (*GetVersion)(int *) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return K2Y_OK;
    const char *names[3] = {getenv("K2Y_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        if (!nm || !nm[0]) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        k2y::set_error("NCCL not found (dlopen libnccl.so.2 failed: %s); set K2Y_NCCL_LIB", dlerror());
        return K2Y_ERR_STATE;
    }
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GetErrorString) {
        k2y::set_error("libnccl lacks one of ncclGetUniqueId/ncclCommInitRank/ncclCommDestroy/ncclAllGather/ncclGetErrorString");
        dlclose(h);
        return K2Y_ERR_STATE;
    }
    g_nccl = a;
    return K2Y_OK;
}

#define K2Y_NCCL_CHECK(expr)                                                                          \
    do {                                                                                              \
        nccl_result_t _r = (expr);                                                                    \
        if (_r != 0) {                                                                                \
            k2y::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r));  \
            return K2Y_ERR_CUDA;                                                                      \
        }                                                                                             \
    } while (0)

}  // namespace

struct k2y_comm {
    nccl_comm_t comm = nullptr;
    int world = 0, rank = 0, device = 0;
};

extern "C" int k2y_comm_unique_id(unsigned char *id128) {
    if (!id128) {
        k2y::set_error("k2y_comm_unique_id: null buffer");
        return K2Y_ERR_INVALID;
    }
    int rc = load_nccl();
    if (rc != K2Y_OK) return rc;
    nccl_unique_id id;
    K2Y_NCCL_CHECK(g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, sizeof(id.internal));
    return K2Y_OK;
}

extern "C" int k2y_comm_create(const unsigned char *id128, int world, int rank, int device, k2y_comm **out) {
    if (!id128 || !out || world < 1 || rank < 0 || rank >= world) {
        k2y::set_error("k2y_comm_create: bad arguments (world %d, rank %d)", world, rank);
        return K2Y_ERR_INVALID;
    }
    int rc = load_nccl();
    if (rc != K2Y_OK) return rc;
    K2Y_CUDA_CHECK(cudaSetDevice(device));
    nccl_unique_id id;
    memcpy(id.internal, id128, sizeof(id.internal));
    k2y_comm *c = new k2y_comm();
    c->world = world;
    c->rank = rank;
    c->device = device;
    nccl_result_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != 0) {
        k2y::set_error("ncclCommInitRank(world %d, rank %d): %s", world, rank, g_nccl.GetErrorString(r));
        delete c;
        return K2Y_ERR_CUDA;
    }
    *out = c;
    return K2Y_OK;
}

extern "C" int k2y_comm_destroy(k2y_comm *c) {
    if (!c) return K2Y_OK;
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
    return K2Y_OK;
}

extern "C" int k2y_comm_info(const k2y_comm *c, int *world, int *rank, int *nccl_version) {
    if (!c) {
        k2y::set_error("k2y_comm_info: comm is NULL");
        return K2Y_ERR_INVALID;
    }
    if (world) *world = c->world;
    if (rank) *rank = c->rank;
    if (nccl_version) {
        *nccl_version = 0;
        if (g_nccl.GetVersion) g_nccl.GetVersion(nccl_version);
    }
    return K2Y_OK;
}

// gather_dev: [world][bytes_per_rank]; this rank's block (written by k2y_detect_keras_strided straight into it) is at
// rank * bytes_per_rank.  In-place all-gather, asynchronous on `stream`.
extern "C" int k2y_allgather_detections(k2y_comm *c, void *gather_dev, size_t bytes_per_rank, void *stream) {
    if (!c || !c->comm || !gather_dev || bytes_per_rank == 0) {
        k2y::set_error("k2y_allgather_detections: bad arguments");
        return K2Y_ERR_INVALID;
    }
    if (c->world == 1) return K2Y_OK;
    char *base = reinterpret_cast<char *>(gather_dev);
    K2Y_NCCL_CHECK(g_nccl.AllGather(base + (size_t)c->rank * bytes_per_rank, base, bytes_per_rank, NCCL_INT8, c->comm,
                                    (cudaStream_t)stream));
    return K2Y_OK;
}
