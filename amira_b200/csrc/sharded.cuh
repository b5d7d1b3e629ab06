// sharded.cuh -- the multi-GPU half of the build (included by gmg.cu; one process per GPU).
//
// Reads are sharded contiguously in rank order, so the global first-seen order of upstream's dicts
// (construct_graph.py:45-100) is the order of GLOBAL call positions call_base[rank] + p.  Each rank
// first builds its local node / edge tables exactly as on one GPU (k_insert_windows).  Then:
//
//   nodes   one record per locally-unique gene-mer (canonical key, count, first global position,
//           first direction) is routed to the owner rank hash(key) mod world with an all-to-all;
//           the owner merges (sum of counts, minimum position) in a hash table; the merged
//           records are published to every rank, sorted by first position = upstream's `_nodes`
//           order, and every rank maps the global nodes onto its local table by probing it.
//   edges   the same with one record per locally-unique undirected adjacency, keyed on GLOBAL
//           node indices (lo, hi, sd*td), carrying the pair count and the first pair event.
//
// Every rank ends with the identical global node / edge tables (adjacency and components are
// computed redundantly on them: they are U-sized); per-read lists and the node -> read incidence
// stay on the rank that owns the reads.  Exchange volume is O(locally unique), not O(windows).
#pragma once

namespace amira {

constexpr int MAX_WORLD = 64;

struct NodeRec {             // 16 B
    unsigned long long ord;  // (global call position of the first window) << 1 | first window was the reverse complement
    unsigned int cov;
    unsigned int pad;
};

__device__ __forceinline__ int owner_of(unsigned long long h, int world) {
    return (int)((((h >> 20) & 0xFFFFFull) * (unsigned long long)world) >> 20);
}

// position in the per-destination send range for every active lane: lanes of a warp that go to the same
// destination share one atomicAdd (a per-record atomic on `world` counters serialises at L2)
__device__ __forceinline__ long long reserve_for_dest(unsigned long long *counts, int d) {
    const unsigned int peers = __match_any_sync(__activemask(), d);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(&counts[d], (unsigned long long)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    return (long long)base + __popc(peers & ((1u << lane) - 1u));
}

__device__ __forceinline__ long long reserve_one(unsigned long long *counter) {
    const unsigned int peers = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    return (long long)base + __popc(peers & ((1u << lane) - 1u));
}

__device__ __forceinline__ void load_canonical(const int32_t *ids, unsigned long long word, int k, int32_t *out) {
    const unsigned long long p = (word >> 1) & P_MASK;
    const int neg = (int)(word & 1ull);
    for (int j = 0; j < k; ++j) out[j] = neg ? -ids[p + (k - 1 - j)] : ids[p + j];
}

// Where the records for each owner go: with peer windows (CUDA IPC over NVLink) these are addresses
// in the OWNER's memory, so the routing kernel is the dispatch -- no staging buffer, no send/recv;
// without them they all point into the local send buffer of the NCCL all-to-all.
struct NodeDst {
    int32_t *key[MAX_WORLD];
    NodeRec *meta[MAX_WORLD];
    long long start[MAX_WORLD];  // first record index of this rank's block at that owner
};
struct EdgeDst {
    EdgeSlot *rec[MAX_WORLD];
    long long start[MAX_WORLD];
};

// pass 1 / pass 2 over the local node table: count per owner, then scatter to the owners
template <bool SCATTER>
__global__ void k_node_route(const NodeView nv, const int32_t *__restrict__ ids, int k,
                             int world, long long call_base, unsigned long long *__restrict__ counts,
                             const NodeDst dst) {
    __shared__ unsigned int s_cnt[MAX_WORLD];
    if (!SCATTER) {
        for (int i = threadIdx.x; i < world; i += blockDim.x) s_cnt[i] = 0;
        __syncthreads();
    }
    const unsigned int stride = gridDim.x * blockDim.x;
    int32_t key[MAX_K];
    for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < nv.cap; s += stride) {
        const unsigned long long w = nv.w(s);
        if (w == EMPTY64) continue;
        load_canonical(ids, w, k, key);
        const int d = owner_of(canonical_hash(key, k, 0), world);
        if (!SCATTER) {
            atomicAdd(&s_cnt[d], 1u);
        } else {
            const long long i = dst.start[d] + reserve_for_dest(counts, d);
            int32_t *kd = dst.key[d] + i * k;
            for (int j = 0; j < k; ++j) kd[j] = key[j];
            NodeRec r;
            r.ord = ((unsigned long long)(call_base + (long long)((w >> 1) & P_MASK)) << 1) | (w & 1ull);
            r.cov = nv.c(s) + 1u;
            r.pad = 0;
            dst.meta[d][i] = r;
        }
    }
    if (!SCATTER) {
        __syncthreads();
        for (int i = threadIdx.x; i < world; i += blockDim.x)
            if (s_cnt[i]) atomicAdd(&counts[i], (unsigned long long)s_cnt[i]);
    }
}

// ---- first-seen order of merged records without a sort ------------------------------------------------
// Every merged node (edge) has a distinct first global call position: a bitmap over the global call range, a prefix
// popcount, and the rank of a record is the number of set bits below its own -- the same device as on one GPU
// (post_kernels.cuh), over calls_global bits.  perm[rank] = record.
struct NodePos {
    const NodeRec *meta;
    __device__ __forceinline__ unsigned long long operator()(long long i) const { return meta[i].ord >> 1; }
};
struct EdgePos {
    const EdgeSlot *recs;
    __device__ __forceinline__ unsigned long long operator()(long long i) const { return recs[i].ord >> 2; }
};

template <class Pos>
__global__ void k_mark_ord(const Pos pos, long long n, unsigned int *__restrict__ bm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long p = pos(i);
    atomicOr(&bm[p >> 5], 1u << (p & 31));
}

struct BmLoad {
    const unsigned int *bm;
    __device__ __forceinline__ unsigned long long operator()(long long i) const { return (unsigned long long)__popc(bm[i]); }
};
struct BmStore {
    int *pref;
    __device__ __forceinline__ void operator()(long long i, unsigned long long excl, unsigned long long) const { pref[i] = (int)excl; }
};

template <class Pos>
__global__ void k_rank_ord(const Pos pos, long long n, const unsigned int *__restrict__ bm, const int *__restrict__ pref,
                           unsigned int *__restrict__ perm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long p = pos(i);
    perm[pref[p >> 5] + __popc(bm[p >> 5] & ((1u << (p & 31)) - 1u))] = (unsigned int)i;
}

// records (canonical keys, k ids each, in the order they arrived) -> table; the slot names one record of the
// gene-mer (whichever: it only serves key comparisons) and collects the weighted count and, in ord_min, the
// earliest first occurrence (position << 1 | direction) -- no sort of the received records is needed
__global__ void k_insert_records(const BuildParams P, long long n, const NodeRec *__restrict__ meta,
                                 unsigned long long *__restrict__ ord_min) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int32_t *win = P.ids + r * P.k;
    const unsigned long long h = canonical_hash(win, P.k, 0);
    const unsigned long long mine = ((h >> FP_SHIFT) << FP_SHIFT) | ((unsigned long long)(r * P.k) << 1);
    const unsigned int slot = node_insert(P, win, 0, false, 0ull, 0ull, h, mine);
    if (meta) {
        atomicAdd(&P.ntab[slot].cov, meta[r].cov);
        const unsigned long long o = meta[r].ord;
        if (o < ((volatile unsigned long long *)ord_min)[slot]) atomicMin(&ord_min[slot], o);
    }
}

__global__ void k_pack_merged_nodes(const NodeSlot *__restrict__ tab, unsigned int cap, const int32_t *__restrict__ keys,
                                    const unsigned long long *__restrict__ ord_min, int k, unsigned long long *__restrict__ counter,
                                    int32_t *__restrict__ out_key, NodeRec *__restrict__ out_meta) {
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += stride) {
        const unsigned long long w = tab[s].word;
        if (w == EMPTY64) continue;
        const long long r = (long long)(((w >> 1) & P_MASK) / (unsigned long long)k);
        const long long i = reserve_one(counter);
        for (int j = 0; j < k; ++j) out_key[i * k + j] = keys[r * k + j];
        NodeRec o;
        o.ord = ord_min[s];
        o.cov = tab[s].cov + 1u;   // counts from 0xFFFFFFFF: the sum of the merged weights
        o.pad = 0;
        out_meta[i] = o;
    }
}

__global__ void k_finalize_nodes(const unsigned int *__restrict__ perm, const int32_t *__restrict__ g_key,
                                 const NodeRec *__restrict__ g_meta, int k, long long n, int32_t *__restrict__ node_key,
                                 uint32_t *__restrict__ node_cov, int8_t *__restrict__ node_dir,
                                 uint8_t *__restrict__ link) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long s = perm[i];
    for (int j = 0; j < k; ++j) node_key[i * k + j] = g_key[s * k + j];
    node_cov[i] = g_meta[s].cov;
    node_dir[i] = (g_meta[s].ord & 1ull) ? -1 : 1;
    link[i] = 0;
}

// find-only probe of the LOCAL node table (any layout) for a canonical key; -1 if this rank never saw it.
// Mirrors the probe sequences of node_insert16 / node_insert; after the insert kernel every key is published.
__device__ __forceinline__ long long node_find_local(const BuildParams &L, bool n16, const int32_t *key) {
    const int k = L.k, kb = L.key_bits;
    if (kb > 0) {
        const unsigned int bias = 1u << (kb - 1), lim = (1u << kb) - 1u;
        unsigned long long klo = 0, khi = 0;
        for (int i = 0; i < k; ++i) {
            const unsigned int u = (unsigned int)key[i] + bias;
            if (u >= lim) return -1;  // wider than any id of this rank's reads
            khi = (khi << kb) | (klo >> (64 - kb));
            klo = (klo << kb) | (unsigned long long)u;
        }
        const unsigned long long h = n16 ? (unsigned long long)packed_hash32(klo, khi) : packed_hash(klo, khi);
        if (n16) {
            const unsigned int nb = L.ncap >> 1;
            unsigned int b = (unsigned int)(((unsigned long long)(unsigned int)h * nb) >> 32);
            const unsigned int top = (unsigned int)(((khi << 1) | (klo >> 63)) & ((1u << (64 - FP_SHIFT)) - 1u));
            const unsigned long long keylow = klo & 0x7FFFFFFFFFFFFFFFull;
            for (unsigned int probe = 0; probe < MAX_PROBES; ++probe) {
                const NodeSlot16 *B = L.ntab16 + 2 * (size_t)b;
                for (int i = 0; i < 2; ++i) {
                    const unsigned long long cur = B[i].word;
                    if (cur == EMPTY64) return -1;
                    if ((unsigned int)(cur >> FP_SHIFT) == top && B[i].key == keylow) return 2ll * b + i;
                }
                if (++b == nb) b = 0;
            }
            return -1;
        }
        unsigned int s = (unsigned int)(((unsigned long long)(unsigned int)h * L.ncap) >> 32);
        for (unsigned int probe = 0; probe < MAX_PROBES; ++probe) {
            const NodeSlot &S = L.ntab[s];
            if (S.word == EMPTY64) return -1;
            if (S.klo == klo && S.khi == khi) return s;
            if (++s == L.ncap) s = 0;
        }
        return -1;
    }
    const unsigned long long h = canonical_hash(key, k, 0);
    const unsigned int fp = (unsigned int)(h >> FP_SHIFT);
    unsigned int s = (unsigned int)(((unsigned long long)(unsigned int)h * L.ncap) >> 32);
    for (unsigned int probe = 0; probe < MAX_PROBES; ++probe) {
        const unsigned long long cur = L.ntab[s].word;
        if (cur == EMPTY64) return -1;
        if ((unsigned int)(cur >> FP_SHIFT) == fp && same_as_representative(L, key, 0, cur)) return s;
        if (++s == L.ncap) s = 0;
    }
    return -1;
}

// global node j -> the local slot that holds the same gene-mer (if this rank saw it): the slot learns
// its global node index, the global node its local coverage.  The probed table is the rank's own
// L2-resident node table -- no table over the (much larger) global node set is ever built.
__global__ void k_global_to_local(const BuildParams L, int n16, const NodeView nv, const int32_t *__restrict__ node_key,
                                  long long n_global, uint32_t *__restrict__ cov_local) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_global) return;
    const long long s = node_find_local(L, n16 != 0, node_key + j * L.k);
    if (s < 0) return;
    nv.a((unsigned int)s) = (unsigned int)j;
    cov_local[j] = nv.c((unsigned int)s) + 1u;
}

// merged records -> the same offset in every rank's window (stores over NVLink, 4-byte granularity
// because a record block may start at any multiple of 4 bytes)
struct PubDst {
    uint32_t *ptr[MAX_WORLD];
};
__global__ void k_publish(const uint32_t *__restrict__ src, long long n_words, int world, const PubDst dst) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) {
        const uint32_t v = src[i];
        for (int p = 0; p < world; ++p) dst.ptr[p][i] = v;
    }
}

// ---- edges ----------------------------------------------------------------------------------------
__device__ __forceinline__ EdgeSlot global_edge_record(unsigned long long key, unsigned long long ord, unsigned int cov,
                                                      const NodeView &nv, long long call_base) {
    const unsigned int lo = (unsigned int)(key >> 32), hi = (unsigned int)((key & 0xFFFFFFFFull) >> 1);
    const bool src_hi = (ord >> 1) & 1ull;
    const unsigned int gs = nv.a(src_hi ? hi : lo), gt = nv.a(src_hi ? lo : hi);
    EdgeSlot r;
    r.key = ((unsigned long long)min(gs, gt) << 32) | ((unsigned long long)max(gs, gt) << 1) | (key & 1ull);
    r.ord = ((unsigned long long)(call_base + (long long)(ord >> 2)) << 2) | ((unsigned long long)(gs > gt) << 1) |
            (ord & 1ull);
    r.cov = cov + 1u;
    r.pad[0] = r.pad[1] = r.pad[2] = 0;
    return r;
}

template <bool SCATTER>
__global__ void k_edge_route(const EdgeView ev, const NodeView nv, int world, long long call_base,
                             unsigned long long *__restrict__ counts, const EdgeDst dst) {
    __shared__ unsigned int s_cnt[MAX_WORLD];
    if (!SCATTER) {
        for (int i = threadIdx.x; i < world; i += blockDim.x) s_cnt[i] = 0;
        __syncthreads();
    }
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < ev.cap; s += stride) {
        unsigned long long ekey, eord;
        unsigned int ecov;
        if (!ev.get(s, ekey, eord, ecov)) continue;
        const EdgeSlot r = global_edge_record(ekey, eord, ecov, nv, call_base);
        const int d = owner_of(mix64(r.key), world);
        if (!SCATTER) atomicAdd(&s_cnt[d], 1u);
        else dst.rec[d][dst.start[d] + reserve_for_dest(counts, d)] = r;
    }
    if (!SCATTER) {
        __syncthreads();
        for (int i = threadIdx.x; i < world; i += blockDim.x)
            if (s_cnt[i]) atomicAdd(&counts[i], (unsigned long long)s_cnt[i]);
    }
}

__global__ void k_merge_edges(const EdgeSlot *__restrict__ recs, long long n, EdgeSlot *__restrict__ tab,
                              unsigned int cap, int *__restrict__ status) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const EdgeSlot r = recs[i];
    unsigned int s = (unsigned int)(((unsigned long long)(unsigned int)mix64(r.key) * cap) >> 32);
    for (unsigned int probe = 0; probe < MAX_PROBES; ++probe) {
        unsigned long long cur = tab[s].key;
        if (cur == EMPTY64) {
            unsigned long long old = atomicCAS(&tab[s].key, EMPTY64, r.key);
            cur = (old == EMPTY64) ? r.key : old;
        }
        if (cur == r.key) {
            atomicMin(&tab[s].ord, r.ord);
            atomicAdd(&tab[s].cov, r.cov);
            return;
        }
        if (++s == cap) s = 0;
    }
    status[ST_OVERFLOW_E] = 1;
}

__global__ void k_pack_merged_edges(const EdgeSlot *__restrict__ tab, unsigned int cap,
                                    unsigned long long *__restrict__ counter, EdgeSlot *__restrict__ out) {
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += stride) {
        EdgeSlot e = tab[s];
        if (e.key == EMPTY64) continue;
        e.cov += 1u;   // counts from 0xFFFFFFFF: the sum of the merged pair counts
        out[reserve_one(counter)] = e;
    }
}

// directed edges per undirected record, in first-pair order (2, or 1 for a self edge) -> offsets (scan functors)
struct FanLoad {
    const unsigned int *perm;
    const EdgeSlot *recs;
    __device__ __forceinline__ unsigned long long operator()(long long i) const {
        const unsigned long long key = recs[perm[i]].key;
        return ((unsigned int)(key >> 32) == (unsigned int)((key & 0xFFFFFFFFull) >> 1)) ? 1ull : 2ull;
    }
};
struct FanStore {
    int *pref;
    __device__ __forceinline__ void operator()(long long i, unsigned long long excl, unsigned long long) const {
        pref[i] = (int)excl;
    }
};

__global__ void k_emit_edges_sorted(const unsigned int *__restrict__ perm, const EdgeSlot *__restrict__ recs,
                                    const int *__restrict__ pref, long long n, int32_t *__restrict__ e_src,
                                    int32_t *__restrict__ e_tgt, int8_t *__restrict__ e_sd, int8_t *__restrict__ e_td,
                                    uint32_t *__restrict__ e_cov, uint8_t *__restrict__ link) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const EdgeSlot e = recs[perm[i]];
    const int idx = pref[i];
    const unsigned int lo = (unsigned int)(e.key >> 32), hi = (unsigned int)((e.key & 0xFFFFFFFFull) >> 1);
    const int rel = (e.key & 1ull) ? 1 : -1;
    const bool src_hi = (e.ord >> 1) & 1ull;
    const int src = (int)(src_hi ? hi : lo), tgt = (int)(src_hi ? lo : hi);
    const int sd = (e.ord & 1ull) ? -1 : 1, td = rel * sd;
    if (lo != hi) {
        e_src[idx] = src; e_tgt[idx] = tgt; e_sd[idx] = (int8_t)sd; e_td[idx] = (int8_t)td; e_cov[idx] = e.cov;
        e_src[idx + 1] = tgt; e_tgt[idx + 1] = src; e_sd[idx + 1] = (int8_t)-td; e_td[idx + 1] = (int8_t)-sd;
        e_cov[idx + 1] = e.cov;
        // consecutive nodes joined by an edge form runs; the components pass unites runs (post_kernels.cuh)
        if (src - tgt == 1) link[src] = 1;
        else if (tgt - src == 1) link[tgt] = 1;
    } else {
        e_src[idx] = src; e_tgt[idx] = src; e_sd[idx] = (int8_t)sd; e_td[idx] = (int8_t)td; e_cov[idx] = 2u * e.cov;
    }
}

__global__ void k_add_i32(int32_t *__restrict__ a, long long n, int32_t v) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] += v;
}

}  // namespace amira
