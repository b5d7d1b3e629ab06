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
    // peer-memory window: one cudaMalloc'ed buffer per rank, mapped into every other rank of the node
    // through CUDA IPC, so that the routing kernels store records straight into the owner's memory
    // over NVLink instead of staging them for ncclSend / ncclRecv
    bool p2p_ok = true;          // until a mapping attempt fails on some rank
    void *win_local = nullptr;
    size_t win_bytes = 0;
    void *win_peer[64] = {};
    void *d_scratch = nullptr;   // 128 B per rank: IPC handles / flags travel through NCCL
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
    if (cudaMalloc(&c->d_scratch, 128 * (size_t)(world + 1)) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc of the communicator scratch failed");
        g_nccl.CommDestroy(c->comm);
        delete c;
        return AMIRA_E_NOMEM;
    }
    cudaMemset(c->d_scratch, 0, 128 * (size_t)(world + 1));
    *out = c;
    return AMIRA_OK;
}

int comm_barrier(Comm *c, cudaStream_t st);

static void window_unmap_peers(Comm *c) {
    for (int p = 0; p < c->world; ++p) {
        if (p != c->rank && c->win_peer[p]) cudaIpcCloseMemHandle(c->win_peer[p]);
        c->win_peer[p] = nullptr;
    }
}

static void window_close(Comm *c) {
    window_unmap_peers(c);
    if (c->win_local) cudaFree(c->win_local);
    c->win_local = nullptr;
    c->win_bytes = 0;
}

// Collective.  Makes every rank's window at least `need_bytes` large (the same value on all ranks)
// and maps all windows into this process.  Returns AMIRA_OK with comm_p2p(c) == false if peer
// mapping is not possible on this node (the caller then uses the NCCL send/recv path).
int comm_window_ensure(Comm *c, size_t need_bytes, cudaStream_t st) {
    if (!c->p2p_ok) return AMIRA_OK;
    if (need_bytes <= c->win_bytes) return AMIRA_OK;
    AMIRA_CUDA(cudaStreamSynchronize(st));
    if (c->win_local) {
        // an exported allocation may only be freed once every importer has unmapped it
        window_unmap_peers(c);
        AMIRA_TRY(comm_barrier(c, st));
        AMIRA_CUDA(cudaStreamSynchronize(st));
        window_close(c);
    }
    const size_t bytes = need_bytes + need_bytes / 4 + (1 << 20);
    int ok = 1;
    if (cudaMalloc(&c->win_local, bytes) != cudaSuccess) {
        cudaGetLastError();
        c->win_local = nullptr;
        ok = 0;
    }
    struct Msg {
        cudaIpcMemHandle_t handle;
        int ok;
        int pad[15];
    } mine, all[64];
    static_assert(sizeof(Msg) == 128, "message is 128 bytes");
    memset(&mine, 0, sizeof(mine));
    if (ok && cudaIpcGetMemHandle(&mine.handle, c->win_local) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
    }
    mine.ok = ok;
    char *d = (char *)c->d_scratch;
    AMIRA_CUDA(cudaMemcpyAsync(d, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    AMIRA_NCCL(g_nccl.AllGather(d, d + 128, sizeof(Msg), ncclInt8, c->comm, st));
    AMIRA_CUDA(cudaMemcpyAsync(all, d + 128, sizeof(Msg) * c->world, cudaMemcpyDeviceToHost, st));
    AMIRA_CUDA(cudaStreamSynchronize(st));
    for (int p = 0; p < c->world; ++p) ok &= all[p].ok;
    if (ok) {
        for (int p = 0; p < c->world && ok; ++p) {
            if (p == c->rank) {
                c->win_peer[p] = c->win_local;
            } else if (cudaIpcOpenMemHandle(&c->win_peer[p], all[p].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                c->win_peer[p] = nullptr;
                ok = 0;
            }
        }
    }
    // every rank must take the same path: agree on the outcome
    int *flag = (int *)c->d_scratch;
    AMIRA_CUDA(cudaMemcpyAsync(flag, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
    AMIRA_NCCL(g_nccl.AllReduce(flag, flag, 1, ncclInt32, ncclMin, c->comm, st));
    AMIRA_CUDA(cudaMemcpyAsync(&ok, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    AMIRA_CUDA(cudaStreamSynchronize(st));
    if (!ok) {
        window_close(c);
        c->p2p_ok = false;
        return AMIRA_OK;
    }
    c->win_bytes = bytes;
    return AMIRA_OK;
}

bool comm_p2p(const Comm *c) { return c && c->p2p_ok && c->world > 1 && !getenv("AMIRA_NO_P2P"); }
void *comm_window(const Comm *c, int p) { return c->win_peer[p]; }

// all ranks' earlier work on `st` (including stores into peer windows) is complete and visible
// when the barrier completes on `st`
int comm_barrier(Comm *c, cudaStream_t st) {
    int *flag = (int *)c->d_scratch + 16;
    AMIRA_NCCL(g_nccl.AllReduce(flag, flag, 1, ncclInt32, ncclMax, c->comm, st));
    return AMIRA_OK;
}

void comm_destroy(Comm *c) {
    if (!c) return;
    window_close(c);
    if (c->d_scratch) cudaFree(c->d_scratch);
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
