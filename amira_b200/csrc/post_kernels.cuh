// post_kernels.cuh -- everything after the insert kernel: first-seen order, node / edge arrays,
// per-read lists, node -> reads and node -> edges transposes, components, filters.
//
// No kernel here takes an element count from the host that the host would first have to read back
// from the device: node, edge and window counts live in the device-side `sizes` array (Cnt::p), so a
// whole build is enqueued without a single host synchronisation (and can be captured in a CUDA
// graph).  Launch grids are sized from host-known upper bounds (call count, table capacities) and the
// kernels grid-stride up to the device-side count.
#pragma once

#include "gmg_kernels.cuh"
#include "incidence.cuh"

namespace amira {

// an element count that lives on the device (p != nullptr) or is known to the host (v)
struct Cnt {
    const long long *p;
    long long v;
    __device__ __forceinline__ long long get() const { return p ? *p : v; }
};


__global__ void k_fill_u32(uint32_t *__restrict__ a, const Cnt n, const long long mul, const long long add, const uint32_t v) {
    const long long m = n.get() * mul + add;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) a[i] = v;
}

__global__ void k_fill_u64(unsigned long long *__restrict__ a, const Cnt n, const long long mul, const long long add,
                           const unsigned long long v) {
    const long long m = n.get() * mul + add;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) a[i] = v;
}

// ---- per-read window offsets (scan functors) ---------------------------------------------------
// win_off = exclusive scan of the per-read window counts (in place)
struct WinOffLoad {
    const int64_t *nwin;
    __device__ __forceinline__ unsigned long long operator()(long long i) const { return (unsigned long long)nwin[i]; }
};
struct WinOffStore {
    int64_t *win_off;
    long long R;
    long long *sizes;
    __device__ __forceinline__ void operator()(long long i, unsigned long long excl, unsigned long long) const {
        win_off[i] = (int64_t)excl;
        if (i == R) sizes[SZ_W] = (long long)excl;
    }
};

// ---- first-seen order ---------------------------------------------------------------------------
// bit p of bm_node is set iff a node was first seen at call p; bm_ea likewise for undirected edge
// entries (first pair at p), bm_eb additionally when the entry is not a self-edge (it then expands to
// two directed edges).  The rank of a node / edge in upstream's dict order is the number of set bits
// below its own: a prefix popcount, no sort.
__global__ void k_mark_first(const NodeView nv, const EdgeView ev, unsigned int *__restrict__ bm_node,
                             unsigned int *__restrict__ bm_ea, unsigned int *__restrict__ bm_eb) {
    const unsigned int stride = gridDim.x * blockDim.x;
    const unsigned int ncap = nv.cap, ecap = ev.cap;
    const unsigned int n = max(ncap, ecap);
    for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += stride) {
        if (s < ncap) {
            unsigned long long w = nv.w(s);
            if (w != EMPTY64) {
                unsigned long long p = (w >> 1) & P_MASK;
                atomicOr(&bm_node[p >> 5], 1u << (p & 31));
            }
        }
        if (s < ecap) {
            unsigned long long key, eord;
            unsigned int ecov;
            if (ev.get(s, key, eord, ecov)) {
                unsigned long long p = eord >> 2;
                atomicOr(&bm_ea[p >> 5], 1u << (p & 31));
                unsigned int lo = (unsigned int)(key >> 32), hi = (unsigned int)((key & 0xFFFFFFFFull) >> 1);
                if (lo != hi) atomicOr(&bm_eb[p >> 5], 1u << (p & 31));
            }
        }
    }
}

// one scan for both ranks: node count in the upper 31 bits of the scanned value, directed-edge count below
struct RankLoad {
    const unsigned int *bm_node, *bm_ea, *bm_eb;
    __device__ __forceinline__ unsigned long long operator()(long long i) const {
        return ((unsigned long long)__popc(bm_node[i]) << 31) | (unsigned long long)(__popc(bm_ea[i]) + __popc(bm_eb[i]));
    }
};
// A build the insert kernel gave up on (palindromic gene-mer, full table, stale call count) is POISONED:
// its tables are not a graph.  Nothing waits for the host to notice, so the counts the later passes run on
// are zeroed here and they all become no-ops; the host sees the status words and raises or retries.
__device__ __forceinline__ bool poisoned(const int *status) {
    return (status[ST_ERR] | status[ST_OVERFLOW_N] | status[ST_OVERFLOW_E] | status[ST_UNPACK] | status[ST_STALE]) != 0;
}
struct RankStore {
    int *pref_node, *pref_edge;
    long long n_words;
    long long *sizes;
    const int *status;
    __device__ __forceinline__ void operator()(long long i, unsigned long long excl, unsigned long long) const {
        const int nn = (int)(excl >> 31), ne = (int)(excl & 0x7FFFFFFFull);
        pref_node[i] = nn;
        pref_edge[i] = ne;
        if (i == n_words) {
            const bool bad = poisoned(status);
            sizes[SZ_NODES] = bad ? 0 : nn;
            sizes[SZ_EDGES] = bad ? 0 : ne;
            if (bad) sizes[SZ_W] = 0;
        }
    }
};

// Node arrays in first-seen order, in two steps so that nothing is written scattered (a 4-byte store to a random
// place of a fresh array costs a 32-byte sector fill and a write-back): k_rank_nodes gives every occupied slot
// its rank and records the inverse (node -> slot, one L2-resident array); k_emit_nodes then runs over the NODES
// in order, gathers the slot and writes every array coalesced.
__global__ void k_rank_nodes(const NodeView nv, const unsigned int *__restrict__ bm_node, const int *__restrict__ pref_node,
                             uint32_t *__restrict__ node_slot) {
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < nv.cap; s += stride) {
        const unsigned long long w = nv.w(s);
        if (w == EMPTY64) continue;
        const unsigned long long p = (w >> 1) & P_MASK;
        const int idx = pref_node[p >> 5] + __popc(bm_node[p >> 5] & ((1u << (p & 31)) - 1u));
        nv.a(s) = (unsigned int)idx;
        node_slot[idx] = s;
    }
}

// node_cov[idx] = windows counted by the insert kernel
__global__ void k_emit_nodes(const NodeView nv, const int32_t *__restrict__ ids, int k, const Cnt n_nodes,
                             const uint32_t *__restrict__ node_slot, int32_t *__restrict__ node_key,
                             uint32_t *__restrict__ node_cov, int8_t *__restrict__ node_dir, uint8_t *__restrict__ link,
                             const NodeSlot16 *__restrict__ tab16, const int key_bits) {
    const long long N = n_nodes.get();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < N; idx += stride) {
        const unsigned int s = node_slot[idx];
        const unsigned long long w = nv.w(s);
        const unsigned long long p = (w >> 1) & P_MASK;
        const int neg = (int)(w & 1ull);
        node_cov[idx] = nv.c(s) + 1u;
        node_dir[idx] = neg ? -1 : 1;
        link[idx] = 0;
        if (tab16) {
            // 16-byte slots: the canonical gene-mer is in the slot (top 22 bits in `word`, low 63 in `key`), no need
            // to go back to the ids
            const unsigned long long top = w >> FP_SHIFT, keylow = tab16[s].key;
            const unsigned long long klo = keylow | ((top & 1ull) << 63), khi = top >> 1;
            const unsigned long long mask = (1ull << key_bits) - 1ull;
            const int bias = 1 << (key_bits - 1);
            for (int j = 0; j < k; ++j) {
                const int sh = (k - 1 - j) * key_bits;
                unsigned long long f;
                if (sh >= 64) f = khi >> (sh - 64);
                else if (sh == 0) f = klo;
                else f = (klo >> sh) | (khi << (64 - sh));
                node_key[idx * k + j] = (int)(f & mask) - bias;
            }
        } else {
            for (int j = 0; j < k; ++j)
                node_key[idx * k + j] = neg ? -ids[p + (k - 1 - j)] : ids[p + j];
        }
    }
}

// ---- union-find -----------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(int32_t *parent, int x) {
    // path halving (every other node on the way is re-hung under its grandparent, and the walk moves on to the
    // grandparent: one level per load); races only ever replace a parent by one of its ancestors
    while (true) {
        const int p = ((volatile int32_t *)parent)[x];
        if (p == x) return x;
        const int gp = ((volatile int32_t *)parent)[p];
        if (gp == p) return p;
        parent[x] = gp;
        x = gp;
    }
}

// Roots are linked by a hashed priority, not by index: first-seen node indices follow the reads, so
// linking by index would build list-shaped trees (node i+1 under node i) and serialise every find.
// Random linking keeps the expected depth logarithmic; the component's first node is recovered
// afterwards with an atomicMin per root (k_cc_flatten).
__device__ __forceinline__ unsigned int uf_prio(int x) {
    unsigned int v = (unsigned int)x * 0x9E3779B1u;
    v ^= v >> 15;
    v *= 0x85EBCA77u;
    v ^= v >> 13;
    return v;
}

__device__ __forceinline__ void uf_union(int32_t *parent, int a, int b) {
    int ra = uf_find(parent, a), rb = uf_find(parent, b);
    while (ra != rb) {
        const unsigned int pa = uf_prio(ra), pb = uf_prio(rb);
        if (pa < pb || (pa == pb && ra < rb)) {
            int t = ra;
            ra = rb;
            rb = t;
        }
        // hook the root of higher priority value under the other one
        int old = atomicCAS(&parent[ra], ra, rb);
        if (old == ra) return;
        ra = uf_find(parent, old);
        rb = uf_find(parent, rb);
    }
}

// Edge arrays in first-seen order: each undirected table entry expands to upstream's forward edge S->T and
// reverse edge T->S (-td, -sd), or to one self edge counted twice (construct_graph.py:246-277).  As for the
// nodes, k_rank_edges first records which table entry every DIRECTED edge comes from (bit 31: the reverse edge),
// then k_emit_edges runs over the directed edges in order and writes coalesced; the adjacency degree of the
// source side of every directed edge is counted on the way.
constexpr uint32_t EDGE_REV = 0x80000000u;
__global__ void k_rank_edges(const EdgeView ev, const unsigned int *__restrict__ bm_ea, const unsigned int *__restrict__ bm_eb,
                             const int *__restrict__ pref_edge, uint32_t *__restrict__ edge_slot, const int *__restrict__ status) {
    if (poisoned(status)) return;
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < ev.cap; s += stride) {
        unsigned long long key, ord;
        unsigned int ecov;
        if (!ev.get(s, key, ord, ecov)) continue;
        const unsigned long long p = ord >> 2;
        const unsigned int below = (1u << (p & 31)) - 1u;
        const int idx = pref_edge[p >> 5] + __popc(bm_ea[p >> 5] & below) + __popc(bm_eb[p >> 5] & below);
        const unsigned int lo = (unsigned int)(key >> 32), hi = (unsigned int)((key & 0xFFFFFFFFull) >> 1);
        edge_slot[idx] = s;
        if (lo != hi) edge_slot[idx + 1] = s | EDGE_REV;
    }
}

__global__ void k_emit_edges(const EdgeView ev, const NodeView nv, const uint32_t *__restrict__ edge_slot, const Cnt n_edges,
                             const Cnt n_nodes, int32_t *__restrict__ e_src, int32_t *__restrict__ e_tgt,
                             int8_t *__restrict__ e_sd, int8_t *__restrict__ e_td, uint32_t *__restrict__ e_cov,
                             unsigned long long *__restrict__ deg, uint8_t *__restrict__ link) {
    const long long E = n_edges.get(), N = n_nodes.get();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < E; idx += stride) {
        const uint32_t v = edge_slot[idx];
        const bool rev = (v & EDGE_REV) != 0;
        unsigned long long key, ord;
        unsigned int ecov;
        ev.get(v & ~EDGE_REV, key, ord, ecov);
        const unsigned int lo = (unsigned int)(key >> 32), hi = (unsigned int)((key & 0xFFFFFFFFull) >> 1);
        const int rel = (key & 1ull) ? 1 : -1;
        const bool src_hi = (ord >> 1) & 1ull;
        const int fsrc = (int)nv.a(src_hi ? hi : lo), ftgt = (int)nv.a(src_hi ? lo : hi);
        const int fsd = (ord & 1ull) ? -1 : 1, ftd = rel * fsd;
        // forward edge S->T, or the reverse edge T->S with directions (-td, -sd); S == T: forward and reverse are
        // the same Edge object, incremented twice per pair
        const int src = rev ? ftgt : fsrc, tgt = rev ? fsrc : ftgt;
        const int sd = rev ? -ftd : fsd, td = rev ? -fsd : ftd;
        e_src[idx] = src;
        e_tgt[idx] = tgt;
        e_sd[idx] = (int8_t)sd;
        e_td[idx] = (int8_t)td;
        e_cov[idx] = lo != hi ? ecov + 1u : 2u * (ecov + 1u);
        atomicAdd(&deg[src + (sd < 0 ? N : 0)], 1ull);
        // consecutive first-seen nodes joined by an edge (three quarters of all adjacencies: reads walk paths)
        // form RUNS; the components pass unites runs, not nodes
        if (!rev) {
            if (src - tgt == 1) link[src] = 1;
            else if (tgt - src == 1) link[tgt] = 1;
        }
    }
}

// ---- node -> reads offsets: exclusive scan of the coverages in node order (the lists themselves: incidence.cuh)
struct CovLoad {
    const uint32_t *cov;
    __device__ __forceinline__ unsigned long long operator()(long long i) const { return cov[i]; }
};
struct CovStore {
    int64_t *reads_off;
    Cnt n;
    long long *sizes;
    __device__ __forceinline__ void operator()(long long i, unsigned long long excl, unsigned long long) const {
        reads_off[i] = (int64_t)excl;
        if (i == n.get()) sizes[SZ_INC] = (long long)excl;
    }
};

// lazy removal of the equal neighbours the sort left (a gene-mer that occurs twice on one read): unique
// counts -> offsets (scan), then one warp per node copies its list without them
struct UniqLoad {
    const int64_t *raw_off;
    const uint32_t *dups;
    __device__ __forceinline__ unsigned long long operator()(long long i) const {
        return (unsigned long long)(raw_off[i + 1] - raw_off[i]) - dups[i];
    }
};
struct OffStore {
    int64_t *off;
    Cnt n;
    long long *total;  // nullable
    __device__ __forceinline__ void operator()(long long i, unsigned long long excl, unsigned long long) const {
        off[i] = (int64_t)excl;
        if (total && i == n.get()) *total = (long long)excl;
    }
};

__global__ void k_compact_unique(const int64_t *__restrict__ raw_off, const uint32_t *__restrict__ in,
                                 const int64_t *__restrict__ out_off, uint32_t *__restrict__ out, const Cnt n_nodes) {
    const long long N = n_nodes.get();
    const int lane = threadIdx.x & 31;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long node = (((long long)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; node < N; node += n_warps) {
        const long long a = raw_off[node], n = raw_off[node + 1] - a;
        long long o = out_off[node];
        for (long long base = 0; base < n; base += 32) {
            const long long i = base + lane;
            uint32_t v = 0;
            bool keep = false;
            if (i < n) {
                v = in[a + i];
                keep = (i == 0) || in[a + i - 1] != v;
            }
            const unsigned int m = __ballot_sync(0xffffffffu, keep);
            if (keep) out[o + __popc(m & ((1u << lane) - 1u))] = v;
            o += __popc(m);
        }
    }
}

// ---- adjacency: forward list of node n = edges with source n and stored source direction +1, in edge
// creation order (construct_graph.py:287-298); backward list likewise with -1.  Counting-sort transpose:
// degrees (k_emit_edges / k_adj_count), scan, scatter through a cursor, segments sorted by edge index.
__global__ void k_adj_count(const int32_t *__restrict__ e_src, const int8_t *__restrict__ e_sd, const Cnt n_edges,
                            const Cnt n_nodes, unsigned long long *__restrict__ deg) {
    const long long E = n_edges.get(), N = n_nodes.get();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride)
        atomicAdd(&deg[e_src[e] + (e_sd[e] < 0 ? N : 0)], 1ull);
}

struct DegLoad {
    const unsigned long long *deg;
    __device__ __forceinline__ unsigned long long operator()(long long i) const { return deg[i]; }
};
struct DegStore {
    int64_t *adj_off;
    unsigned long long *cursor;
    Cnt n_nodes;
    long long *sizes;
    __device__ __forceinline__ void operator()(long long i, unsigned long long excl, unsigned long long) const {
        adj_off[i] = (int64_t)excl;
        cursor[i] = excl;
        if (i == n_nodes.get()) sizes[SZ_FW] = (long long)excl;
    }
};

__global__ void k_adj_scatter(const int32_t *__restrict__ e_src, const int8_t *__restrict__ e_sd, const Cnt n_edges,
                              const Cnt n_nodes, unsigned long long *__restrict__ cursor, uint32_t *__restrict__ adj_edges) {
    const long long E = n_edges.get(), N = n_nodes.get();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
        const unsigned long long pos = atomicAdd(&cursor[e_src[e] + (e_sd[e] < 0 ? N : 0)], 1ull);
        adj_edges[pos] = (uint32_t)e;
    }
}

// ---- components -------------------------------------------------------------------------------------
// union-find over the emitted edges in first-seen order (each undirected adjacency once)
// run_id (nullable): run of every node; parent then spans the runs.  done (nullable): the small-graph pass below
// has already united everything.
__global__ void k_union_edges(const int32_t *__restrict__ e_src, const int32_t *__restrict__ e_tgt, const Cnt n_edges,
                              int32_t *__restrict__ parent, const int32_t *__restrict__ run_id,
                              const unsigned long long *__restrict__ done) {
    if (done && *done) return;
    const long long E = n_edges.get();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
        const int s = e_src[e], t = e_tgt[e];
        if (s >= t) continue;
        if (run_id) {
            if (t - s == 1) continue;  // same run by construction
            const int a = run_id[s], b = run_id[t];
            if (a != b) uf_union(parent, a, b);
        } else {
            uf_union(parent, s, t);
        }
    }
}

// ---- small graphs (the regime Amira lives in: a few 10^4 nodes) ---------------------------------------------
// A union is a chain of dependent loads, ~log n rounds of them while all edges hook at once; at L2 latency that is
// ~80 us for a 30 000-node graph, a quarter of the whole build.  When the runs fit the shared memory of ONE CTA
// the same unions run at shared-memory latency: the run-leaving edges are first collected into a compact list
// (all SMs), then one CTA unites them in shared memory and writes every run's root back.
constexpr int UF_SMALL_RUNS = 48 * 1024;   // runs (4 bytes each) one CTA holds
constexpr int UF_SMALL_THREADS = 1024;

__global__ void k_collect_run_edges(const int32_t *__restrict__ e_src, const int32_t *__restrict__ e_tgt, const Cnt n_edges,
                                    const Cnt n_nodes, const int32_t *__restrict__ run_id, unsigned long long *__restrict__ list,
                                    unsigned long long *__restrict__ counter) {
    const long long E = n_edges.get(), N = n_nodes.get();
    if (N == 0 || run_id[N - 1] >= UF_SMALL_RUNS) return;  // too many runs: k_union_edges does the work
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (long long e0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll; e0 < E; e0 += stride) {
        const long long e = e0 + lane;
        int a = 0, b = 0;
        if (e < E) {
            const int s = e_src[e], t = e_tgt[e];
            if (s < t && t - s != 1) {
                a = run_id[s];
                b = run_id[t];
            }
        }
        const unsigned int m = __ballot_sync(0xffffffffu, a != b);
        if (!m) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (a != b) list[base + __popc(m & ((1u << lane) - 1u))] = ((unsigned long long)(unsigned int)a << 32) | (unsigned int)b;
    }
}

__device__ __forceinline__ int uf_find_shared(int32_t *parent, int x) {
    while (true) {
        const int p = ((volatile int32_t *)parent)[x];
        if (p == x) return x;
        const int gp = ((volatile int32_t *)parent)[p];
        if (gp == p) return p;
        parent[x] = gp;
        x = gp;
    }
}

__global__ void __launch_bounds__(UF_SMALL_THREADS, 1)
k_union_small(const unsigned long long *__restrict__ list, const unsigned long long *__restrict__ counter, const Cnt n_nodes,
              const int32_t *__restrict__ run_id, int32_t *__restrict__ parent, unsigned long long *__restrict__ done) {
    extern __shared__ int32_t s_parent[];
    const long long N = n_nodes.get();
    const int n_runs = N ? run_id[N - 1] + 1 : 0;
    if (n_runs > UF_SMALL_RUNS) return;  // done stays 0
    for (int i = threadIdx.x; i < n_runs; i += UF_SMALL_THREADS) s_parent[i] = i;
    __syncthreads();
    const long long n = (long long)*counter;
    for (long long i = threadIdx.x; i < n; i += UF_SMALL_THREADS) {
        const unsigned long long pr = list[i];
        int ra = uf_find_shared(s_parent, (int)(pr >> 32)), rb = uf_find_shared(s_parent, (int)(pr & 0xFFFFFFFFull));
        while (ra != rb) {
            const unsigned int pa = uf_prio(ra), pb = uf_prio(rb);
            if (pa < pb || (pa == pb && ra < rb)) {
                const int t = ra;
                ra = rb;
                rb = t;
            }
            const int old = atomicCAS(&s_parent[ra], ra, rb);
            if (old == ra) break;
            ra = uf_find_shared(s_parent, old);
            rb = uf_find_shared(s_parent, rb);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_runs; i += UF_SMALL_THREADS) {
        int r = i;
        for (int p = s_parent[r]; p != r; p = s_parent[r]) r = p;
        parent[i] = r;
    }
    if (threadIdx.x == 0) *done = 1ull;
}

// runs of consecutive linked nodes: run_id = (number of run starts up to and including the node) - 1
struct RunLoad {
    const uint8_t *link;
    __device__ __forceinline__ unsigned long long operator()(long long i) const { return link[i] ? 0ull : 1ull; }
};
struct RunStore {
    int32_t *run_id, *parent;
    Cnt n;
    __device__ __forceinline__ void operator()(long long i, unsigned long long excl, unsigned long long v) const {
        if (i >= n.get()) return;
        const int32_t r = (int32_t)(excl + v) - 1;
        run_id[i] = r;
        if (v) parent[r] = r;
    }
};

// root of every node (read-only walk: the unions are over, trees are shallow thanks to the random
// linking) and each component's first node (cmin starts at 0xFFFFFFFF)
__global__ void k_cc_flatten(const int32_t *__restrict__ parent, const Cnt n_nodes, unsigned int *__restrict__ cmin,
                             uint32_t *__restrict__ root, const int32_t *__restrict__ run_id) {
    const long long N = n_nodes.get();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        int r = run_id ? run_id[i] : (int)i;
        for (int p = parent[r]; p != r; p = parent[r]) r = p;
        root[i] = (uint32_t)r;
        // one atomic per (warp, component): the lanes of a warp hold ascending node indices, so the lowest lane of
        // every group carries the group's minimum; threads run roughly in index order, so after the first few
        // updates the minimum is final and the remaining warps of a (giant) component skip the atomic
        const unsigned int peers = __match_any_sync(__activemask(), r);
        if ((threadIdx.x & 31) == __ffs(peers) - 1 && (unsigned int)i < ((volatile unsigned int *)cmin)[r])
            atomicMin(&cmin[r], (unsigned int)i);
    }
}

struct FirstLoad {
    const uint32_t *root;
    const unsigned int *cmin;
    __device__ __forceinline__ unsigned long long operator()(long long i) const {
        return cmin[root[i]] == (unsigned int)i ? 1ull : 0ull;
    }
};
struct FirstStore {
    int *first_rank;
    Cnt n;
    long long *sizes;
    __device__ __forceinline__ void operator()(long long i, unsigned long long excl, unsigned long long) const {
        first_rank[i] = (int)excl;
        if (i == n.get()) sizes[SZ_COMPS] = (long long)excl;
    }
};

// component ids 1, 2, ... in order of each component's first node (construct_graph.py:920-927);
// comp holds the roots on entry
__global__ void k_cc_number(const unsigned int *__restrict__ cmin, const int *__restrict__ first_rank, const Cnt n_nodes,
                            uint32_t *__restrict__ comp) {
    const long long N = n_nodes.get();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride)
        comp[i] = (uint32_t)first_rank[cmin[comp[i]]] + 1u;
}

// ---- filters (the host knows the sizes of the graph it filters; the sizes AFTER the filter stay on
// the device) -----------------------------------------------------------------------------------------
__global__ void k_component_max(const uint32_t *__restrict__ node_cov, const uint32_t *__restrict__ comp,
                                int64_t n_nodes, uint32_t *__restrict__ comp_max) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes) {
        // one giant component is the common case: read first, the maximum only ever grows
        const uint32_t c = comp[i], v = node_cov[i];
        if (v > ((volatile uint32_t *)comp_max)[c]) atomicMax(&comp_max[c], v);
    }
}

// keep flags: mode 0 = coverage >= thr (filter_graph), mode 1 = component max >= thr
__global__ void k_node_keep(const uint32_t *__restrict__ node_cov, const uint32_t *__restrict__ comp,
                            const uint32_t *__restrict__ comp_max, int64_t n_nodes, uint32_t thr, int mode,
                            int *__restrict__ keep) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes) keep[i] = mode == 0 ? (node_cov[i] >= thr) : (comp_max[comp[i]] >= thr);
    else if (i == n_nodes) keep[i] = 0;
}

__global__ void k_edge_keep(const int32_t *__restrict__ e_src, const int32_t *__restrict__ e_tgt,
                            const uint32_t *__restrict__ e_cov, const int *__restrict__ node_keep, int64_t n_edges,
                            uint32_t thr, int *__restrict__ keep) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_edges) keep[e] = (e_cov[e] >= thr) && node_keep[e_src[e]] && node_keep[e_tgt[e]];
    else if (e == n_edges) keep[e] = 0;
}

struct KeepLoad {
    const int *keep;
    __device__ __forceinline__ unsigned long long operator()(long long i) const { return keep[i] ? 1ull : 0ull; }
};
struct KeepStore {
    int *newidx;
    long long n;
    long long *total;
    __device__ __forceinline__ void operator()(long long i, unsigned long long excl, unsigned long long) const {
        newidx[i] = (int)excl;
        if (i == n) *total = (long long)excl;
    }
};

// upstream's remove_node dies with TypeError when a doomed node has two edges to one neighbour
// (get_edge_hashes_between_nodes returns lists, construct_graph.py:383-386, 479-482)
__global__ void k_multi_edge_check(const int32_t *__restrict__ e_src, const int32_t *__restrict__ e_tgt,
                                   const uint32_t *__restrict__ adj_edges, const int64_t *__restrict__ adj_off,
                                   const int *__restrict__ node_keep, int64_t n_nodes, int *__restrict__ status) {
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_nodes || node_keep[n]) return;
    // the node's forward edges live at [adj_off[n], adj_off[n+1]), its backward ones at [adj_off[N+n], ...)
    for (int pass = 0; pass < 2; ++pass) {
        int64_t a0 = adj_off[n + pass * n_nodes], a1 = adj_off[n + pass * n_nodes + 1];
        for (int64_t i = a0; i < a1; ++i) {
            int32_t t = e_tgt[adj_edges[i]];
            for (int pass2 = pass; pass2 < 2; ++pass2) {
                int64_t b0 = pass2 == pass ? i + 1 : adj_off[n + n_nodes], b1 = adj_off[n + pass2 * n_nodes + 1];
                for (int64_t j = b0; j < b1; ++j)
                    if (e_tgt[adj_edges[j]] == t) status[ST_ERR] = AMIRA_E_MULTI_EDGE;
            }
        }
    }
}

__global__ void k_compact_nodes(const int *__restrict__ keep, const int *__restrict__ newidx, int64_t n_nodes, int k,
                                const int32_t *__restrict__ key_in, const uint32_t *__restrict__ cov_in,
                                const int8_t *__restrict__ dir_in, const uint32_t *__restrict__ comp_in,
                                const int64_t *__restrict__ roff_in, int32_t *__restrict__ key_out,
                                uint32_t *__restrict__ cov_out, int8_t *__restrict__ dir_out,
                                uint32_t *__restrict__ comp_out, int64_t *__restrict__ rcount_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes || !keep[i]) return;
    const int o = newidx[i];
    for (int j = 0; j < k; ++j) key_out[(int64_t)o * k + j] = key_in[i * k + j];
    cov_out[o] = cov_in[i];
    dir_out[o] = dir_in[i];
    comp_out[o] = comp_in[i];
    rcount_out[o] = roff_in[i + 1] - roff_in[i];
}

struct CountLoad {
    const int64_t *cnt;
    __device__ __forceinline__ unsigned long long operator()(long long i) const { return (unsigned long long)cnt[i]; }
};

// one warp per surviving node copies its read list
__global__ void k_compact_incidence(const int *__restrict__ keep, const int *__restrict__ newidx, int64_t n_nodes,
                                    const int64_t *__restrict__ roff_in, const uint32_t *__restrict__ reads_in,
                                    const int64_t *__restrict__ roff_out, uint32_t *__restrict__ reads_out) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_nodes || !keep[warp]) return;
    const int64_t a = roff_in[warp], n = roff_in[warp + 1] - a, b = roff_out[newidx[warp]];
    for (int64_t i = lane; i < n; i += 32) reads_out[b + i] = reads_in[a + i];
}

__global__ void k_compact_edges(const int *__restrict__ keep, const int *__restrict__ newidx,
                                const int *__restrict__ node_newidx, int64_t n_edges,
                                const int32_t *__restrict__ src_in, const int32_t *__restrict__ tgt_in,
                                const int8_t *__restrict__ sd_in, const int8_t *__restrict__ td_in,
                                const uint32_t *__restrict__ cov_in, int32_t *__restrict__ src_out,
                                int32_t *__restrict__ tgt_out, int8_t *__restrict__ sd_out,
                                int8_t *__restrict__ td_out, uint32_t *__restrict__ cov_out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges || !keep[e]) return;
    const int o = newidx[e];
    src_out[o] = node_newidx[src_in[e]];
    tgt_out[o] = node_newidx[tgt_in[e]];
    sd_out[o] = sd_in[e];
    td_out[o] = td_in[e];
    cov_out[o] = cov_in[e];
}

// remove_node_from_reads (construct_graph.py:442-461): windows of removed nodes become None and
// their reads join _readsToCorrect
__global__ void k_mask_windows(const int *__restrict__ node_keep, const int *__restrict__ node_newidx,
                               int32_t *__restrict__ win_node, int8_t *__restrict__ win_dir,
                               const int32_t *__restrict__ win_read, const int32_t read_base,
                               int32_t *__restrict__ win_start, int32_t *__restrict__ win_end, const long long W,
                               uint8_t *__restrict__ to_correct) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < W; w += stride) {
        const int n = win_node[w];
        if (n < 0) continue;
        if (node_keep[n]) {
            win_node[w] = node_newidx[n];
        } else {
            win_node[w] = -1;
            win_dir[w] = 0;
            if (win_start) {
                win_start[w] = -1;
                win_end[w] = -1;
            }
            to_correct[win_read[w] - read_base] = 1;
        }
    }
}

}  // namespace amira
