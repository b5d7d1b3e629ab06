// gmg_kernels.cuh -- hand-written sm_100a kernels of the GeneMerGraph build.
//
// Upstream (pure Python) walks reads one at a time and does ~28 SHA-256 per gene-mer
// (amira/construct_graph.py:45-100).  Here the reads are a CSR of signed int32 gene ids and the
// whole build is a handful of data-parallel passes:
//
//   k_read_windows     per read: window count (construct_read.py:41-43), short-read flag, and the
//                      read that owns each 1024-call tile boundary
//   k_insert_windows   THE hot kernel.  One CTA per 1024-call tile: coalesced 128-bit staging of
//                      the tile (+k halo) in shared memory, read boundaries by a block max-scan,
//                      per-window canonicalisation against the reverse complement
//                      (construct_gene_mer.py:4-39), hash, insert into the node table with
//                      first-seen tracking (atomicCAS / atomicMin on one 64-bit word), coverage
//                      RED.ADD, per-window outputs, then adjacent-pair edges from the slot numbers
//                      staged in shared memory (construct_graph.py:246-324)
//   k_mark_first / k_popcount / k_emit_nodes / k_emit_edges
//                      first-seen order without a sort: a bitmap over call positions + a prefix
//                      popcount gives every node / edge its rank in upstream's dict order
//   k_remap_windows    slot -> node index for the per-read node lists (construct_graph.py:165-178)
//   k_incidence_*      node -> unique ascending reads (construct_node.py:64-67) from a stable radix
//                      sort of (node, read) by node
//   k_cc_*             connected components with lock-free union-find, numbered in first-node
//                      order (construct_graph.py:911-927)
//   k_filter_*         coverage / component thresholds + order-preserving stream compaction
//                      (construct_graph.py:496-540, 950-958)
#pragma once

#include <cub/cub.cuh>

#include "common.cuh"

namespace amira {

constexpr int INS_THREADS = 256;
constexpr int INS_ITEMS = 4;
constexpr int INS_TILE = INS_THREADS * INS_ITEMS;  // calls per tile
constexpr int MAX_K = 64;
constexpr int NR_CAP = INS_TILE + 2;               // reads whose offsets are cached per tile
constexpr unsigned int INVALID_VAL = 0xFFFFFFFFu;
constexpr unsigned int MAX_PROBES = 1u << 15;

enum { ST_ERR = 0, ST_OVERFLOW_N = 1, ST_OVERFLOW_E = 2, ST_COUNT = 4 };
enum { SZ_W = 0, SZ_NODES = 1, SZ_EDGES = 2, SZ_INC = 3, SZ_SHORT = 4, SZ_FW = 5, SZ_BW = 6, SZ_COUNT = 8 };

struct BuildParams {
    const int32_t *ids;
    const int64_t *off;
    const int64_t *win_off;
    const int32_t *tile_r0;
    const int32_t *ps;
    const int32_t *pe;
    int64_t G, R, n_tiles;
    int k;
    NodeSlot *ntab;
    unsigned int ncap;
    EdgeSlot *etab;
    unsigned int ecap;
    int32_t *win_node;
    int8_t *win_dir;
    int32_t *win_read;
    int32_t *win_start;
    int32_t *win_end;
    int *status;
    int64_t read_base;  // global index of this shard's first read (multi-GPU)
};

__host__ __device__ __forceinline__ int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

struct MaxOp {
    __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; }
};

// ---------------------------------------------------------------------------------------------
__global__ void k_read_windows(const int64_t *__restrict__ off, int64_t R, int k, int64_t G,
                               int64_t *__restrict__ nwin, uint8_t *__restrict__ is_short,
                               uint8_t *__restrict__ to_correct, int32_t *__restrict__ tile_r0,
                               long long *__restrict__ sizes, int *__restrict__ status) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool shortr = false;
    if (r < R) {
        int64_t a = off[r], b = off[r + 1];
        if (b < a || a < 0 || b > G) {
            status[ST_ERR] = AMIRA_E_ARG;
            b = a;
        }
        int64_t L = b - a;
        int64_t nw = L >= k ? L - k + 1 : 0;
        nwin[r] = nw;
        shortr = nw == 0;
        is_short[r] = shortr;
        to_correct[r] = 0;
        // this read owns every tile whose first call lies inside it
        for (int64_t t = (a + INS_TILE - 1) / INS_TILE; t * INS_TILE < b; ++t) tile_r0[t] = (int32_t)r;
    } else if (r == R) {
        nwin[R] = 0;
    }
    unsigned m = __ballot_sync(0xffffffffu, shortr);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd((unsigned long long *)&sizes[SZ_SHORT], (unsigned long long)__popc(m));
}

// ---------------------------------------------------------------------------------------------
// node table insert-or-find.  win = the window's k ids in shared memory.
__device__ __forceinline__ unsigned int node_insert(const BuildParams &P, const int32_t *win, int dirneg,
                                                    unsigned long long h, unsigned long long mine) {
    const unsigned int cap = P.ncap;
    const int k = P.k;
    unsigned int s = (unsigned int)(((unsigned long long)(unsigned int)h * cap) >> 32);
    const unsigned int fp = (unsigned int)(mine >> FP_SHIFT);
    for (unsigned int probe = 0; probe < MAX_PROBES; ++probe) {
        unsigned long long cur = __ldcg(&P.ntab[s].word);
        if (cur == EMPTY64) {
            unsigned long long old = atomicCAS(&P.ntab[s].word, EMPTY64, mine);
            if (old == EMPTY64) return s;
            cur = old;
        }
        if ((unsigned int)(cur >> FP_SHIFT) == fp) {
            // same fingerprint: compare against the slot's representative window in the input
            const int64_t q = (int64_t)((cur >> 1) & P_MASK);
            const int qneg = (int)(cur & 1ull);
            bool same = true;
            for (int j = 0; j < k; ++j) {
                int a = dirneg ? -win[k - 1 - j] : win[j];
                int b = qneg ? -__ldg(P.ids + q + (k - 1 - j)) : __ldg(P.ids + q + j);
                if (a != b) {
                    same = false;
                    break;
                }
            }
            if (same) {
                if (mine < cur) atomicMin(&P.ntab[s].word, mine);  // keep the first occurrence
                return s;
            }
        }
        if (++s == cap) s = 0;
    }
    P.status[ST_OVERFLOW_N] = 1;
    return 0;
}

// hash of the canonical form of a window (dirneg: the window is the reverse complement of it)
__device__ __forceinline__ unsigned long long canonical_hash(const int32_t *win, int k, int dirneg) {
    unsigned long long h = 0x9e3779b97f4a7c15ULL;
    for (int i = 0; i < k; ++i) {
        int g = dirneg ? -win[k - 1 - i] : win[i];
        h = (h ^ (unsigned long long)(unsigned int)g) * 0x100000001b3ULL;
        h ^= h >> 29;
    }
    return mix64(h);
}

__device__ __forceinline__ void edge_insert(const BuildParams &P, unsigned long long key, unsigned long long ord) {
    const unsigned int cap = P.ecap;
    unsigned long long h = mix64(key);
    unsigned int s = (unsigned int)(((unsigned long long)(unsigned int)h * cap) >> 32);
    for (unsigned int probe = 0; probe < MAX_PROBES; ++probe) {
        unsigned long long cur = __ldcg(&P.etab[s].key);
        if (cur == EMPTY64) {
            unsigned long long old = atomicCAS(&P.etab[s].key, EMPTY64, key);
            cur = (old == EMPTY64) ? key : old;
        }
        if (cur == key) {
            if (ord < __ldcg(&P.etab[s].ord)) atomicMin(&P.etab[s].ord, ord);
            atomicAdd(&P.etab[s].cov, 1u);
            return;
        }
        if (++s == cap) s = 0;
    }
    P.status[ST_OVERFLOW_E] = 1;
}

__global__ void __launch_bounds__(INS_THREADS) k_insert_windows(const BuildParams P) {
    __shared__ __align__(16) int32_t s_ids[INS_TILE + MAX_K + 4];
    __shared__ unsigned int s_val[INS_TILE + 1];  // slot | dirneg << 31 per window start, INVALID_VAL if none
    __shared__ int s_j[INS_TILE];                 // read (relative to the tile's first read) of each call
    __shared__ long long s_off[NR_CAP + 1];
    __shared__ long long s_woff[NR_CAP];
    typedef cub::BlockScan<int, INS_THREADS> Scan;
    __shared__ typename Scan::TempStorage s_scan;

    const int tid = threadIdx.x;
    const int k = P.k;
    for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        const int64_t t0 = tile * INS_TILE;
        const int len = (int)imin64(INS_TILE, P.G - t0);
        const int n_load = (int)imin64(len + k, P.G - t0);
        const int r_lo = P.tile_r0[tile];
        const int r_hi = (tile + 1 < P.n_tiles) ? P.tile_r0[tile + 1] : (int)(P.R - 1);
        const int nr = r_hi - r_lo + 1;

        // ---- stage the tile (+ halo) with 128-bit loads; t0 is a multiple of 1024 calls
        {
            const int32_t *src = P.ids + t0;
            const int n4 = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) ? (n_load >> 2) : 0;
            const int4 *src4 = reinterpret_cast<const int4 *>(src);
            int4 *dst4 = reinterpret_cast<int4 *>(s_ids);
            for (int i = tid; i < n4; i += INS_THREADS) dst4[i] = __ldg(src4 + i);
            for (int i = (n4 << 2) + tid; i < n_load; i += INS_THREADS) s_ids[i] = __ldg(src + i);
        }
        for (int i = tid; i < INS_TILE; i += INS_THREADS) s_j[i] = 0;
        for (int i = tid; i <= nr && i <= NR_CAP; i += INS_THREADS) s_off[i] = P.off[r_lo + i];
        for (int i = tid; i < nr && i < NR_CAP; i += INS_THREADS) s_woff[i] = P.win_off[r_lo + i];
        __syncthreads();
        // ---- read boundaries: mark each later read's first call, then an inclusive max-scan
        for (int j = 1 + tid; j < nr; j += INS_THREADS) {
            long long o = (j <= NR_CAP ? s_off[j] : P.off[r_lo + j]) - t0;
            if (o < len) atomicMax(&s_j[(int)o], j);
        }
        __syncthreads();
        {
            int items[INS_ITEMS];
#pragma unroll
            for (int i = 0; i < INS_ITEMS; ++i) items[i] = s_j[tid * INS_ITEMS + i];
            Scan(s_scan).InclusiveScan(items, items, MaxOp());
#pragma unroll
            for (int i = 0; i < INS_ITEMS; ++i) s_j[tid * INS_ITEMS + i] = items[i];
        }
        __syncthreads();

        // ---- windows: pl == len is the halo window that only serves the last pair of the tile
        for (int pl = tid; pl <= len; pl += INS_THREADS) {
            const bool halo = (pl == len);
            const int j = s_j[halo ? pl - 1 : pl];
            const int64_t p = t0 + pl;
            const long long re = (j + 1 <= NR_CAP) ? s_off[j + 1] : P.off[r_lo + j + 1];
            unsigned int val = INVALID_VAL;
            if (p + k <= re) {
                const int32_t *win = s_ids + pl;
                int dir = 0;
                for (int i = 0; i < k; ++i) {
                    int f = win[i], c = -win[k - 1 - i];
                    if (f != c) {
                        dir = f < c ? 1 : -1;
                        break;
                    }
                }
                if (dir == 0) {
                    P.status[ST_ERR] = AMIRA_E_PALINDROME;  // construct_gene_mer.py:23-25
                } else {
                    const int dirneg = dir < 0;
                    const unsigned long long h = canonical_hash(win, k, dirneg);
                    const unsigned long long mine =
                        ((h >> FP_SHIFT) << FP_SHIFT) | ((unsigned long long)p << 1) | (unsigned long long)dirneg;
                    const unsigned int slot = node_insert(P, win, dirneg, h, mine);
                    val = slot | ((unsigned int)dirneg << 31);
                    if (!halo) {
                        atomicAdd(&P.ntab[slot].cov, 1u);
                        const long long rs = (j <= NR_CAP) ? s_off[j] : P.off[r_lo + j];
                        const long long wo = (j < NR_CAP) ? s_woff[j] : P.win_off[r_lo + j];
                        const int64_t w = wo + (p - rs);
                        P.win_node[w] = (int32_t)slot;
                        P.win_dir[w] = (int8_t)dir;
                        P.win_read[w] = (int32_t)(r_lo + j);
                        if (P.ps) {
                            P.win_start[w] = __ldg(P.ps + p);
                            P.win_end[w] = __ldg(P.pe + p + k - 1);
                        }
                    }
                }
            }
            s_val[pl] = val;
        }
        __syncthreads();
        // ---- adjacent pairs of the same read
        for (int pl = tid; pl < len; pl += INS_THREADS) {
            const unsigned int a = s_val[pl], b = s_val[pl + 1];
            if (a == INVALID_VAL || b == INVALID_VAL) continue;
            if (pl + 1 < len && s_j[pl + 1] != s_j[pl]) continue;
            const unsigned int sa = a & 0x7FFFFFFFu, sb = b & 0x7FFFFFFFu;
            const unsigned int sdneg = a >> 31, tdneg = b >> 31;
            const unsigned int lo = min(sa, sb), hi = max(sa, sb);
            const unsigned long long key =
                ((unsigned long long)lo << 32) | ((unsigned long long)hi << 1) | (unsigned long long)(sdneg == tdneg);
            const unsigned long long ord =
                ((unsigned long long)(t0 + pl) << 2) | ((unsigned long long)(sa > sb) << 1) | sdneg;
            edge_insert(P, key, ord);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// first-seen order: bit p of bm_node is set iff a node was first seen at call p; bm_ea likewise for
// undirected edge entries (first pair at p), bm_eb additionally when the entry is not a self-edge
// (it then expands to two directed edges).
__global__ void k_mark_first(const NodeSlot *__restrict__ ntab, unsigned int ncap,
                             const EdgeSlot *__restrict__ etab, unsigned int ecap,
                             unsigned int *__restrict__ bm_node, unsigned int *__restrict__ bm_ea,
                             unsigned int *__restrict__ bm_eb) {
    const unsigned int stride = gridDim.x * blockDim.x;
    const unsigned int n = max(ncap, ecap);
    for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += stride) {
        if (s < ncap) {
            unsigned long long w = ntab[s].word;
            if (w != EMPTY64) {
                unsigned long long p = (w >> 1) & P_MASK;
                atomicOr(&bm_node[p >> 5], 1u << (p & 31));
            }
        }
        if (s < ecap) {
            unsigned long long key = etab[s].key;
            if (key != EMPTY64) {
                unsigned long long p = etab[s].ord >> 2;
                atomicOr(&bm_ea[p >> 5], 1u << (p & 31));
                unsigned int lo = (unsigned int)(key >> 32), hi = (unsigned int)((key & 0xFFFFFFFFull) >> 1);
                if (lo != hi) atomicOr(&bm_eb[p >> 5], 1u << (p & 31));
            }
        }
    }
}

__global__ void k_popcount(const unsigned int *__restrict__ bm_node, const unsigned int *__restrict__ bm_ea,
                           const unsigned int *__restrict__ bm_eb, int64_t n_words, int *__restrict__ cnt_node,
                           int *__restrict__ cnt_edge) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_words) {
        cnt_node[i] = __popc(bm_node[i]);
        cnt_edge[i] = __popc(bm_ea[i]) + __popc(bm_eb[i]);
    } else if (i == n_words) {
        cnt_node[i] = 0;
        cnt_edge[i] = 0;
    }
}

__global__ void k_collect_sizes(const int64_t *__restrict__ win_off, int64_t R, const int *__restrict__ pref_node,
                                const int *__restrict__ pref_edge, int64_t n_words, long long *__restrict__ sizes) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        sizes[SZ_W] = win_off[R];
        sizes[SZ_NODES] = pref_node[n_words];
        sizes[SZ_EDGES] = pref_edge[n_words];
    }
}

__global__ void k_emit_nodes(NodeSlot *__restrict__ ntab, unsigned int ncap, const int32_t *__restrict__ ids, int k,
                             const unsigned int *__restrict__ bm_node, const int *__restrict__ pref_node,
                             int32_t *__restrict__ node_key, uint32_t *__restrict__ node_cov,
                             int8_t *__restrict__ node_dir, int32_t *__restrict__ parent) {
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < ncap; s += stride) {
        unsigned long long w = ntab[s].word;
        if (w == EMPTY64) continue;
        const unsigned long long p = (w >> 1) & P_MASK;
        const int neg = (int)(w & 1ull);
        const int idx = pref_node[p >> 5] + __popc(bm_node[p >> 5] & ((1u << (p & 31)) - 1u));
        ntab[s].aux = (unsigned int)idx;
        node_cov[idx] = ntab[s].cov + 1u;
        node_dir[idx] = neg ? -1 : 1;
        parent[idx] = idx;
        for (int j = 0; j < k; ++j)
            node_key[(int64_t)idx * k + j] = neg ? -ids[p + (k - 1 - j)] : ids[p + j];
    }
}

__device__ __forceinline__ int uf_find(int32_t *parent, int x) {
    // path halving; races only ever replace a parent by one of its ancestors
    while (true) {
        int p = ((volatile int32_t *)parent)[x];
        if (p == x) return x;
        int gp = ((volatile int32_t *)parent)[p];
        if (gp != p) parent[x] = gp;
        x = p;
    }
}

__device__ __forceinline__ void uf_union(int32_t *parent, int a, int b) {
    int ra = uf_find(parent, a), rb = uf_find(parent, b);
    while (ra != rb) {
        if (ra < rb) {
            int t = ra;
            ra = rb;
            rb = t;
        }
        // hook the larger root under the smaller: the final root is the component's first node
        int old = atomicCAS(&parent[ra], ra, rb);
        if (old == ra) return;
        ra = uf_find(parent, old);
        rb = uf_find(parent, rb);
    }
}

__global__ void k_emit_edges(const EdgeSlot *__restrict__ etab, unsigned int ecap, const NodeSlot *__restrict__ ntab,
                             const unsigned int *__restrict__ bm_ea, const unsigned int *__restrict__ bm_eb,
                             const int *__restrict__ pref_edge, int32_t *__restrict__ e_src,
                             int32_t *__restrict__ e_tgt, int8_t *__restrict__ e_sd, int8_t *__restrict__ e_td,
                             uint32_t *__restrict__ e_cov, int32_t *__restrict__ parent) {
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < ecap; s += stride) {
        const unsigned long long key = etab[s].key;
        if (key == EMPTY64) continue;
        const unsigned long long ord = etab[s].ord;
        const unsigned long long p = ord >> 2;
        const unsigned int below = (1u << (p & 31)) - 1u;
        const int idx = pref_edge[p >> 5] + __popc(bm_ea[p >> 5] & below) + __popc(bm_eb[p >> 5] & below);
        const unsigned int lo = (unsigned int)(key >> 32), hi = (unsigned int)((key & 0xFFFFFFFFull) >> 1);
        const int rel = (key & 1ull) ? 1 : -1;
        const bool src_hi = (ord >> 1) & 1ull;
        const int src = (int)ntab[src_hi ? hi : lo].aux, tgt = (int)ntab[src_hi ? lo : hi].aux;
        const int sd = (ord & 1ull) ? -1 : 1, td = rel * sd;
        const uint32_t cov = etab[s].cov + 1u;
        if (lo != hi) {
            // forward edge S->T, then the reverse edge T->S with directions (-td, -sd)
            e_src[idx] = src; e_tgt[idx] = tgt; e_sd[idx] = (int8_t)sd; e_td[idx] = (int8_t)td; e_cov[idx] = cov;
            e_src[idx + 1] = tgt; e_tgt[idx + 1] = src; e_sd[idx + 1] = (int8_t)-td; e_td[idx + 1] = (int8_t)-sd;
            e_cov[idx + 1] = cov;
            uf_union(parent, src, tgt);
        } else {
            // S == T: forward and reverse are the same Edge object, incremented twice per pair
            e_src[idx] = src; e_tgt[idx] = src; e_sd[idx] = (int8_t)sd; e_td[idx] = (int8_t)td; e_cov[idx] = 2u * cov;
        }
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void k_remap_windows(const NodeSlot *__restrict__ ntab, int32_t *__restrict__ win_node, int64_t W) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < W; w += stride)
        win_node[w] = (int32_t)__ldg(&ntab[win_node[w]].aux);
}

// after the stable sort by node: duplicates (same node, same read) are adjacent.  Count them per
// node (rare) and flag the survivors.
__global__ void k_incidence_flags(const int32_t *__restrict__ keys, const int32_t *__restrict__ vals, int64_t W,
                                  uint8_t *__restrict__ flags, uint32_t *__restrict__ dups) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < W; i += stride) {
        bool first = (i == 0) || keys[i] != keys[i - 1] || vals[i] != vals[i - 1];
        flags[i] = first;
        if (!first) atomicAdd(&dups[keys[i]], 1u);
    }
}

__global__ void k_incidence_counts(const uint32_t *__restrict__ node_cov, const uint32_t *__restrict__ dups,
                                   int64_t n_nodes, int64_t *__restrict__ reads_off) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes) reads_off[i] = (int64_t)node_cov[i] - (int64_t)dups[i];
    else if (i == n_nodes) reads_off[i] = 0;
}

// ---------------------------------------------------------------------------------------------
// adjacency: forward list of node n = edges with source n and stored source direction +1, in edge
// creation order (construct_graph.py:287-298); backward list likewise with -1.
__global__ void k_adj_keys(const int32_t *__restrict__ e_src, const int8_t *__restrict__ e_sd, int64_t n_edges,
                           int64_t n_nodes, uint32_t *__restrict__ keys, int32_t *__restrict__ vals,
                           int64_t *__restrict__ deg) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    uint32_t key = (uint32_t)e_src[e] + (e_sd[e] < 0 ? (uint32_t)n_nodes : 0u);
    keys[e] = key;
    vals[e] = (int32_t)e;
    atomicAdd((unsigned long long *)&deg[key], 1ull);
}

// ---------------------------------------------------------------------------------------------
__global__ void k_cc_roots(int32_t *__restrict__ parent, int64_t n_nodes, int *__restrict__ is_root) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes) {
        int r = uf_find(parent, (int)i);
        parent[i] = r;
        is_root[i] = (r == (int)i);
    } else if (i == n_nodes) {
        is_root[i] = 0;
    }
}

__global__ void k_cc_number(const int32_t *__restrict__ parent, const int *__restrict__ root_rank, int64_t n_nodes,
                            uint32_t *__restrict__ comp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes) {
        int r = parent[i];
        r = parent[r];  // k_cc_roots flattened against a moving target; one more hop is enough
        comp[i] = (uint32_t)root_rank[r] + 1u;
    }
}

// ---------------------------------------------------------------------------------------------
// filters
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

// upstream's remove_node dies with TypeError when a doomed node has two edges to one neighbour
// (get_edge_hashes_between_nodes returns lists, construct_graph.py:383-386, 479-482)
__global__ void k_multi_edge_check(const int32_t *__restrict__ e_src, const int32_t *__restrict__ e_tgt,
                                   const int32_t *__restrict__ adj_edges, const int64_t *__restrict__ adj_off,
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
    if (i == 0) rcount_out[newidx[n_nodes]] = 0;
    if (i >= n_nodes || !keep[i]) return;
    const int o = newidx[i];
    for (int j = 0; j < k; ++j) key_out[(int64_t)o * k + j] = key_in[i * k + j];
    cov_out[o] = cov_in[i];
    dir_out[o] = dir_in[i];
    comp_out[o] = comp_in[i];
    rcount_out[o] = roff_in[i + 1] - roff_in[i];
}

// one warp per surviving node copies its read list
__global__ void k_compact_incidence(const int *__restrict__ keep, const int *__restrict__ newidx, int64_t n_nodes,
                                    const int64_t *__restrict__ roff_in, const int32_t *__restrict__ reads_in,
                                    const int64_t *__restrict__ roff_out, int32_t *__restrict__ reads_out) {
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
                               const int32_t *__restrict__ win_read, int32_t *__restrict__ win_start,
                               int32_t *__restrict__ win_end, int64_t W, uint8_t *__restrict__ to_correct) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < W; w += stride) {
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
            to_correct[win_read[w]] = 1;
        }
    }
}

__global__ void k_filter_sizes(const int *__restrict__ node_newidx, int64_t n_nodes, const int *__restrict__ edge_newidx,
                               int64_t n_edges, long long *__restrict__ sizes) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        sizes[SZ_NODES] = node_newidx[n_nodes];
        sizes[SZ_EDGES] = edge_newidx[n_edges];
    }
}

// ---------------------------------------------------------------------------------------------
// atomic roofline micro-benchmark: random-address RED.ADD.u32 / CAS.b64 into a table
__global__ void k_atomic_red(unsigned int *__restrict__ table, unsigned long long n_slots, unsigned long long n_ops,
                             unsigned long long seed) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_ops; i += stride) {
        unsigned long long h = mix64(i + seed);
        atomicAdd(&table[(unsigned long long)(((h >> 32) * n_slots) >> 32) ], 1u);
    }
}

__global__ void k_atomic_cas(unsigned long long *__restrict__ table, unsigned long long n_slots,
                             unsigned long long n_ops, unsigned long long seed) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_ops; i += stride) {
        unsigned long long h = mix64(i + seed);
        acc += atomicCAS(&table[(unsigned long long)(((h >> 32) * n_slots) >> 32)], EMPTY64, h | 1ull);
    }
    if (acc == 0x123456789abcdefULL) table[0] = acc;  // keep the returns alive
}

}  // namespace amira
