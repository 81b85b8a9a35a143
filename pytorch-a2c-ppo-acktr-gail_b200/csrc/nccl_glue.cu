// NCCL glue of the environment-sharded update: the flat gradient buffer is all-reduced (sum)
// once per optimiser step.  The exchange is part of the op list a2cb_run_ops launches: a FORK op hands a finished
// segment of the gradient buffer to NCCL on a side stream while the remaining weight-gradient kernels keep running on
// the main stream (GRU + heads first, then fc, the convolutions last), a JOIN op makes the main stream wait for all of
// them before the clip + optimiser kernels.  The entry points take an ALREADY INITIALISED ncclComm_t (the host side
// owns rendezvous: a2cb_nccl_unique_id / a2cb_nccl_comm_init_rank wrap ncclGetUniqueId / ncclCommInitRank for callers
// that have none).  libnccl.so.2 -- the copy torch already loaded -- is resolved with dlopen at first use, so the
// library links and loads on boxes without NCCL (single-GPU use never touches it).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "decls.cuh"

namespace a2cb {

namespace {

struct NcclUniqueId { char internal[128]; };          // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
typedef int (*fn_get_uid)(NcclUniqueId*);
typedef int (*fn_init_rank)(void**, int, NcclUniqueId, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
    void* handle = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
} g_nccl;

bool nccl_load(const char* path) {
    if (g_nccl.handle) return true;
    void* h = nullptr;
    if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);     // the copy torch loaded
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("nccl: cannot load libnccl.so.2: %s", dlerror()); return false; }
    g_nccl.get_uid = (fn_get_uid)dlsym(h, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
    g_nccl.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
    g_nccl.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
    g_nccl.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.allreduce || !g_nccl.destroy) {
        set_error("nccl: libnccl.so.2 lacks an expected symbol");
        return false;
    }
    g_nccl.handle = h;
    return true;
}

int nccl_fail(const char* what, int rc) {
    set_error("nccl: %s failed: %s", what, g_nccl.errstr ? g_nccl.errstr(rc) : "?");
    return A2CB_ERR_NCCL;
}

}  // namespace

// exchange context: the communicator + a side stream + the events that order it against the launching stream
struct CommCtx {
    void* comm;
    void* p2p;                     // peer-memory exchange (p2p.cu): takes the forked segments it has room for
    cudaStream_t side;
    cudaEvent_t ready[8], done[8];
    int pending;                   // forks since the last join
};

int nccl_load_library(const char* path) { return nccl_load(path) ? A2CB_OK : A2CB_ERR_NCCL; }

int nccl_unique_id(void* out128) {
    A2CB_REQUIRE(out128, "nccl_unique_id: null pointer");
    if (!nccl_load(nullptr)) return A2CB_ERR_NCCL;
    NcclUniqueId id;
    const int rc = g_nccl.get_uid(&id);
    if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(out128, &id, sizeof(id));
    return A2CB_OK;
}

int nccl_comm_init_rank(void** comm, int world, int rank, const void* id128) {
    A2CB_REQUIRE(comm && id128 && world >= 1 && rank >= 0 && rank < world, "nccl_comm_init_rank: bad arguments");
    if (!nccl_load(nullptr)) return A2CB_ERR_NCCL;
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    const int rc = g_nccl.init_rank(comm, world, id, rank);
    return rc == 0 ? A2CB_OK : nccl_fail("ncclCommInitRank", rc);
}

int nccl_comm_destroy(void* comm) {
    if (!comm || !g_nccl.handle) return A2CB_OK;
    const int rc = g_nccl.destroy(comm);
    return rc == 0 ? A2CB_OK : nccl_fail("ncclCommDestroy", rc);
}

int nccl_allreduce_sum_f32(void* comm, float* buf, long count, cudaStream_t st) {
    A2CB_REQUIRE(comm && buf && count > 0, "nccl_allreduce_sum_f32: bad arguments");
    if (!nccl_load(nullptr)) return A2CB_ERR_NCCL;
    const int rc = g_nccl.allreduce(buf, buf, (size_t)count, 7 /* ncclFloat32 */, 0 /* ncclSum */, comm, st);
    return rc == 0 ? A2CB_OK : nccl_fail("ncclAllReduce", rc);
}

int comm_ctx_create(void** ctx, void* comm) {
    A2CB_REQUIRE(ctx && comm, "comm_ctx_create: bad arguments");
    CommCtx* c = (CommCtx*)calloc(1, sizeof(CommCtx));
    A2CB_REQUIRE(c, "comm_ctx_create: out of memory");
    c->comm = comm;
    cudaError_t e = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
    for (int i = 0; i < 8 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&c->ready[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { set_error("comm_ctx_create: %s", cudaGetErrorString(e)); free(c); return A2CB_ERR_CUDA; }
    *ctx = c;
    return A2CB_OK;
}

int comm_ctx_set_p2p(void* ctx, void* p2p) {
    CommCtx* c = (CommCtx*)ctx;
    A2CB_REQUIRE(c, "comm_ctx_set_p2p: null context");
    c->p2p = p2p;
    return A2CB_OK;
}

int comm_ctx_destroy(void* ctx) {
    CommCtx* c = (CommCtx*)ctx;
    if (!c) return A2CB_OK;
    for (int i = 0; i < 8; ++i) { cudaEventDestroy(c->ready[i]); cudaEventDestroy(c->done[i]); }
    cudaStreamDestroy(c->side);
    free(c);
    return A2CB_OK;
}

// FORK: everything launched so far on `st` is ordered before an all-reduce of buf[0..count) on the side stream
int comm_fork_allreduce(void* ctx, float* buf, long count, cudaStream_t st) {
    CommCtx* c = (CommCtx*)ctx;
    A2CB_REQUIRE(c && buf && count > 0 && c->pending < 8, "comm_fork_allreduce: bad arguments (at most 8 forks per join)");
    const int i = c->pending++;
    cudaError_t e = cudaEventRecord(c->ready[i], st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->side, c->ready[i], 0);
    if (e != cudaSuccess) { set_error("comm_fork_allreduce: %s", cudaGetErrorString(e)); return A2CB_ERR_CUDA; }
    const int rc = (c->p2p && count <= p2p_capacity(c->p2p)) ? p2p_allreduce_sum_f32(c->p2p, buf, count, c->side)
                                                             : nccl_allreduce_sum_f32(c->comm, buf, count, c->side);
    if (rc != A2CB_OK) return rc;
    e = cudaEventRecord(c->done[i], c->side);
    if (e != cudaSuccess) { set_error("comm_fork_allreduce: %s", cudaGetErrorString(e)); return A2CB_ERR_CUDA; }
    return A2CB_OK;
}

// JOIN: `st` waits for every all-reduce forked since the last join
int comm_join(void* ctx, cudaStream_t st) {
    CommCtx* c = (CommCtx*)ctx;
    A2CB_REQUIRE(c, "comm_join: null context");
    for (int i = 0; i < c->pending; ++i) {
        const cudaError_t e = cudaStreamWaitEvent(st, c->done[i], 0);
        if (e != cudaSuccess) { set_error("comm_join: %s", cudaGetErrorString(e)); return A2CB_ERR_CUDA; }
    }
    c->pending = 0;
    return A2CB_OK;
}

}  // namespace a2cb
