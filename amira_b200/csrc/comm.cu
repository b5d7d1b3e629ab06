// comm.cu -- multi-GPU plumbing of libamira_gmg.so: NCCL over NVLink 5 / NVSwitch, one process per GPU.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so that the single-GPU library has no NCCL
// dependency; inside a torch process the already-loaded bundled libnccl is picked up.  Only the
// collectives the sharded build needs are wrapped: fixed-size all-gather, in-place all-reduce(max),
// and the two variable-size exchanges (all-to-all-v, all-gather-v) as grouped ncclSend / ncclRecv --
// on NVSwitch every peer is one hop at full bandwidth, so a flat exchange is the right schedule.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace amira {

namespace {

struct NcclApi {
    void *dl = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.dl) return AMIRA_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *dl = nullptr;
    for (const char *n : names) {
        dl = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (dl) break;
    }
    if (!dl) {
        set_error("cannot load libnccl.so.2: %s", dlerror());
        return AMIRA_E_NCCL;
    }
#define BIND(field, sym)                                                        \
    do {                                                                        \
        *(void **)(&g_nccl.field) = dlsym(dl, sym);                             \
        if (!g_nccl.field) {                                                    \
            set_error("libnccl is missing %s", sym);                            \
            return AMIRA_E_NCCL;                                                \
        }                                                                       \
    } while (0)
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(AllGather, "ncclAllGather");
    BIND(AllReduce, "ncclAllReduce");
    BIND(Send, "ncclSend");
    BIND(Recv, "ncclRecv");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
    BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    g_nccl.dl = dl;
    return AMIRA_OK;
}

#define AMIRA_NCCL(expr)                                                                         \
    do {                                                                                         \
        ncclResult_t _r = (expr);                                                                \
        if (_r != ncclSuccess) {                                                                 \
            set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, g_nccl.GetErrorString(_r)); \
            return AMIRA_E_NCCL;                                                                 \
        }                                                                                        \
    } while (0)

}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

int comm_unique_id(void *out_128_bytes) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (!out_128_bytes) return AMIRA_E_ARG;
    AMIRA_TRY(load_nccl());
    ncclUniqueId id;
    AMIRA_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out_128_bytes, &id, sizeof(id));
    return AMIRA_OK;
}

int comm_create(Comm **out, const void *unique_id, int rank, int world) {
    if (!out || !unique_id || world < 1 || rank < 0 || rank >= world) {
        set_error("bad arguments to amira_gmg_comm_init");
        return AMIRA_E_ARG;
    }
    AMIRA_TRY(load_nccl());
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    Comm *c = new Comm();
    c->rank = rank;
    c->world = world;
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete c;
        return AMIRA_E_NCCL;
    }
    *out = c;
    return AMIRA_OK;
}

void comm_destroy(Comm *c) {
    if (!c) return;
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
}

int comm_rank(const Comm *c) { return c ? c->rank : 0; }
int comm_world(const Comm *c) { return c ? c->world : 1; }

int comm_allgather(Comm *c, const void *d_send, void *d_recv, size_t bytes_per_rank, cudaStream_t st) {
    AMIRA_NCCL(g_nccl.AllGather(d_send, d_recv, bytes_per_rank, ncclInt8, c->comm, st));
    return AMIRA_OK;
}

int comm_allreduce_max_i32(Comm *c, int *d_buf, int n, cudaStream_t st) {
    AMIRA_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)n, ncclInt32, ncclMax, c->comm, st));
    return AMIRA_OK;
}

// send_off / recv_off: host arrays of world+1 element offsets into d_send / d_recv
int comm_alltoallv(Comm *c, const void *d_send, const int64_t *send_off, void *d_recv, const int64_t *recv_off,
                   size_t elem_bytes, cudaStream_t st) {
    const char *s = (const char *)d_send;
    char *r = (char *)d_recv;
    const int me = c->rank;
    const size_t self = (size_t)(send_off[me + 1] - send_off[me]) * elem_bytes;
    if (self) AMIRA_CUDA(cudaMemcpyAsync(r + recv_off[me] * elem_bytes, s + send_off[me] * elem_bytes, self,
                                         cudaMemcpyDeviceToDevice, st));
    AMIRA_NCCL(g_nccl.GroupStart());
    for (int p = 0; p < c->world; ++p) {
        if (p == me) continue;
        const size_t ns = (size_t)(send_off[p + 1] - send_off[p]) * elem_bytes;
        const size_t nr = (size_t)(recv_off[p + 1] - recv_off[p]) * elem_bytes;
        if (ns) AMIRA_NCCL(g_nccl.Send(s + send_off[p] * elem_bytes, ns, ncclInt8, p, c->comm, st));
        if (nr) AMIRA_NCCL(g_nccl.Recv(r + recv_off[p] * elem_bytes, nr, ncclInt8, p, c->comm, st));
    }
    AMIRA_NCCL(g_nccl.GroupEnd());
    return AMIRA_OK;
}

// every rank contributes n_send elements; recv_off (host, world+1) places rank p's block in d_recv
int comm_allgatherv(Comm *c, const void *d_send, int64_t n_send, void *d_recv, const int64_t *recv_off,
                    size_t elem_bytes, cudaStream_t st) {
    char *r = (char *)d_recv;
    const int me = c->rank;
    if (n_send) AMIRA_CUDA(cudaMemcpyAsync(r + recv_off[me] * elem_bytes, d_send, (size_t)n_send * elem_bytes,
                                           cudaMemcpyDeviceToDevice, st));
    AMIRA_NCCL(g_nccl.GroupStart());
    for (int p = 0; p < c->world; ++p) {
        if (p == me) continue;
        const size_t nr = (size_t)(recv_off[p + 1] - recv_off[p]) * elem_bytes;
        if (n_send) AMIRA_NCCL(g_nccl.Send(d_send, (size_t)n_send * elem_bytes, ncclInt8, p, c->comm, st));
        if (nr) AMIRA_NCCL(g_nccl.Recv(r + recv_off[p] * elem_bytes, nr, ncclInt8, p, c->comm, st));
    }
    AMIRA_NCCL(g_nccl.GroupEnd());
    return AMIRA_OK;
}

}  // namespace amira

extern "C" int amira_gmg_nccl_unique_id(void *out_128_bytes) { return amira::comm_unique_id(out_128_bytes); }
