// incidence.cuh -- per-read node lists and the node -> reads transpose (Node.listOfReads,
// construct_node.py:64-67: the ascending list of the reads that touch a node) in two streaming passes.
//
// The transpose is a stable counting sort of the W windows by node.  The insert kernel has counted the
// windows per node, so the final CSR offsets `reads_off` (scan of the coverages in node order) are known
// before a single read is placed.  The node axis is cut into UNITS: maximal runs of consecutive nodes
// whose lists start inside one INC_C-window cell of `reads_off` and whose indices share one INC_S-node
// cell, i.e. unit(n) = reads_off[n] / INC_C + n / INC_S (non-decreasing in n).  A unit's lists are one
// contiguous piece of `reads` of at most INC_C + (last list) entries and fit the shared memory of one CTA.
//
//   k_unit_table   unit of every table slot (next to its node index, one 8-byte gather per window later),
//                  first node of every unit
//   k_partition    streams the windows once: slot -> (node, unit) gather, writes the per-read node lists
//                  (construct_graph.py:165-178) and deals (node, read) records into per-BUCKET regions
//                  (bucket = unit >> g, at most INC_NB_MAX of them; a bucket's region is exactly its piece of
//                  the final array, so no histogram pass is needed).  Records of a tile are grouped by
//                  bucket in shared memory and leave in runs; a tile reserves its runs with one atomic per
//                  non-empty bucket.
//   k_unit_lists   one CTA per unit: two sweeps over the bucket's records (count, place) put every read
//                  into a VALUE sub-bucket of its node's list (lists above 32 entries are cut into ~n/4
//                  equal-width read ranges; 16-bit cursors, two to a word), the sub-buckets -- a handful
//                  of entries each -- are sorted with register networks, and the piece is written out
//                  coalesced.  Nothing is sorted in global memory, nothing is read twice from DRAM.
//
// A unit whose last list is so long that the piece outgrows shared memory (a gene-mer on more than
// INC_BIG windows) is done in two parts, the last list on its own; a single list above INC_CAPV entries
// falls back to placing the reads in global memory and sorting them with segsort.cuh (its work lists).
#pragma once

#include "common.cuh"
#include "segsort.cuh"

namespace amira {

constexpr int INC_S = 4096;                  // nodes per unit at most
constexpr int INC_CAPV = 16384;              // list entries a unit sorts in shared memory
constexpr int INC_BIG = 4096;                // a last list up to this long never overflows a unit
constexpr int INC_C = INC_CAPV - INC_BIG;    // window cell of the unit function
constexpr int INC_NB_MAX = 4096;             // buckets the partition pass deals into (its small-CTA form)
constexpr int INC_NB_BIG = 16384;            // ... in its one-CTA-per-SM form, for inputs with more units than that
constexpr int INC_SUB_MIN = 8;               // lists up to this long are one sub-bucket
constexpr int INC_THREADS = 512;             // two CTAs per SM: one sorts while the other streams its records in
constexpr int INC_BATCH = 8;                 // records a thread has in flight
constexpr int INC_CUR_WORDS = (INC_S + INC_CAPV / 2) / 2 + 8;  // 16-bit cursors, two to a word
constexpr size_t INC_SMEM = sizeof(uint32_t) * (INC_CAPV + 4) + sizeof(uint16_t) * (INC_S + 8) + (INC_S + 8) + sizeof(uint32_t) * INC_CUR_WORDS;

// The partition pass comes in three shapes.  Up to INC_NB_MAX units (~45M gene calls): small CTAs, three to an SM,
// on tiles of 2048 windows -- on high-coverage graphs the pass is bound by L2 transactions whatever its shape, and
// small CTAs leave room for the second stream (C5 shard: 1.94 ms per build against 2.03 ms with one big CTA per SM);
// or one CTA per SM on tiles of 8192 windows (512 threads x 16 windows in registers) for LOW-coverage graphs, where a
// tile touches most buckets with a record or two each: the reservations of one bucket are atomics on ONE address, L2
// serialises those (~50 ns each), and the tile size sets the pace (C3: 0.34 ms with 2048-window tiles, 0.20 ms with
// 8192).  Beyond INC_NB_MAX units: 16384 buckets, whose counters take most of an SM's shared memory, 1024 threads.
template <int NB_, int THREADS_, int CTAS_, int ITEMS_ = 8>
struct PartShape {
    static constexpr int NB = NB_, THREADS = THREADS_, CTAS = CTAS_, ITEMS = ITEMS_, TILE = THREADS_ * ITEMS_;
    // (capping the small shape at 64 registers, to leave the second stream more of the register file, was measured
    // SLOWER: 2.03 ms per build against 1.94 ms at the 72 registers the compiler takes)
    static constexpr int REG_CTAS = CTAS_;
    static constexpr int TOUCH = TILE < NB_ ? TILE : NB_;
    static constexpr size_t SMEM = 2 * sizeof(uint32_t) * NB_ + 2 * sizeof(uint16_t) * TOUCH;
};
using PartSmall = PartShape<INC_NB_MAX, 256, 3, 8>;
using PartWide = PartShape<INC_NB_MAX, 512, 1, 16>;
constexpr int PART_WIDE_BELOW = 16;   // windows per node (of the previous build on the handle) below which PartWide runs
using PartBig = PartShape<INC_NB_BIG, 1024, 1>;

struct UnitPlan {
    int nb_max;        // INC_NB_MAX or INC_NB_BIG: which shape of the partition pass runs
    int g;             // bucket = unit >> g
    int n_units;       // entries of unit_lo minus one; a multiple of 1 << g
    int n_buckets;     // n_units >> g
    uint32_t read_lo;  // smallest read index of this rank's windows
    uint32_t rscale;   // (read - read_lo) * rscale spreads this rank's reads over [0, 2^32)
};

__device__ __forceinline__ int unit_of(const int64_t *__restrict__ reads_off, long long n) {
    return (int)(reads_off[n] / INC_C) + (int)(n / INC_S);
}

// first node of every unit (unit_lo[q] .. unit_lo[q+1] are the nodes of unit q; units nobody maps to are
// empty) and the unit of every occupied table slot
__global__ void k_unit_table(const NodeView nv, const int64_t *__restrict__ reads_off, const long long *__restrict__ n_nodes_ptr,
                             const UnitPlan plan, int *__restrict__ unit_lo, int *__restrict__ status) {
    const long long N = *n_nodes_ptr;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long span = max((long long)nv.cap, N + 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < span; i += stride) {
        if (i <= N) {
            const int prev = i ? unit_of(reads_off, i - 1) : -1;
            const int cur = i < N ? unit_of(reads_off, i) : plan.n_units;
            if (i < N && cur >= plan.n_units) status[ST_ERR] = AMIRA_E_STATE;  // cannot happen: the host sized the table from upper bounds
            for (int q = max(prev, -1) + 1; q <= min(cur, plan.n_units); ++q) unit_lo[q] = (int)i;
        }
        if (i < (long long)nv.cap && nv.w((unsigned int)i) != EMPTY64) {
            const long long n = nv.a((unsigned int)i);
            nv.base((unsigned int)i) = n < N ? (unsigned int)unit_of(reads_off, n) : 0u;
        }
    }
}

// start of every bucket's region of the record array (= its piece of the final `reads`)
__global__ void k_bucket_base(const int *__restrict__ unit_lo, const int64_t *__restrict__ reads_off, const UnitPlan plan,
                              uint32_t *__restrict__ bucket_base, unsigned int *__restrict__ bucket_cursor) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < plan.nb_max) {
        bucket_base[b] = b < plan.n_buckets ? (uint32_t)reads_off[unit_lo[b << plan.g]] : 0u;
        bucket_cursor[b] = 0;
    }
}

// ---- pass 1: per-read node lists + (node, read) records dealt into the bucket regions -----------------
// Per tile: gather, rank every window within (tile, bucket) through a shared-memory counter, reserve one run per
// touched bucket in the bucket's region (one global atomic each), store the records at run start + rank.
// The slots and reads of the NEXT tile are loaded before the current one is ranked, so the streams never stop
// while a tile waits for its gathers, its reservations and its barriers.
template <class S>
__global__ void __launch_bounds__(S::THREADS, S::REG_CTAS)
k_partition(const uint2 *__restrict__ info, const int32_t *__restrict__ win_slot, const int32_t *__restrict__ win_read,
            int32_t *__restrict__ win_node, const long long *__restrict__ sizes, const uint32_t *__restrict__ bucket_base,
            const UnitPlan plan, unsigned int *__restrict__ bucket_cursor, uint2 *__restrict__ rec) {
    extern __shared__ __align__(16) unsigned char p_smem[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(p_smem);  // windows of the tile per bucket (zero between tiles)
    constexpr int PART_THREADS = S::THREADS, PART_ITEMS = S::ITEMS, PART_TILE = S::TILE, PART_TOUCH = S::TOUCH;
    uint32_t *run0 = hist + S::NB;                          // where the tile's run starts in the record array
    uint16_t (*touched)[PART_TOUCH] = reinterpret_cast<uint16_t (*)[PART_TOUCH]>(run0 + S::NB);  // buckets the tile touched (double-buffered by tile parity)
    __shared__ uint32_t n_touched[2];
    const long long W = sizes[SZ_W];
    const int g = plan.g;
    const int tid = threadIdx.x;
    for (int b = tid; b < S::NB; b += PART_THREADS) hist[b] = 0;
    if (tid < 2) n_touched[tid] = 0;
    __syncthreads();
    const long long n_tiles = (W + PART_TILE - 1) / PART_TILE;
    int32_t slot[PART_ITEMS], slot_n[PART_ITEMS];
    uint32_t rd[PART_ITEMS], rd_n[PART_ITEMS];
#define PART_LOAD(sl, rr, tile)                                                        \
    {                                                                                  \
        const long long w0_ = (tile) * PART_TILE;                                      \
        _Pragma("unroll") for (int i = 0; i < PART_ITEMS; ++i) {                       \
            const long long w_ = w0_ + i * PART_THREADS + tid;                         \
            sl[i] = w_ < W ? __ldcs(win_slot + w_) : -1;                               \
            rr[i] = w_ < W ? (uint32_t)__ldcs(win_read + w_) : 0u;                     \
        }                                                                              \
    }
    PART_LOAD(slot, rd, (long long)blockIdx.x);
    int par = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, par ^= 1) {
        const long long w0 = t * PART_TILE;
        const int cnt = (int)min((long long)PART_TILE, W - w0);
        uint32_t node[PART_ITEMS], br[PART_ITEMS];
#pragma unroll
        for (int i = 0; i < PART_ITEMS; ++i) {
            uint2 inf = make_uint2(0u, 0u);
            if (slot[i] >= 0) inf = info[slot[i]];
            node[i] = inf.x;
            br[i] = inf.y >> g;
        }
        PART_LOAD(slot_n, rd_n, t + gridDim.x);
#pragma unroll
        for (int i = 0; i < PART_ITEMS; ++i) {
            const int j = i * PART_THREADS + tid;
            if (j < cnt) {
                __stcs(win_node + w0 + j, (int32_t)node[i]);
                const uint32_t r = atomicAdd(&hist[br[i]], 1u);
                if (r == 0) touched[par][atomicAdd(&n_touched[par], 1u)] = (uint16_t)br[i];
                br[i] |= r << 14;  // bucket (< 2^14) | rank within (tile, bucket)
            }
        }
        __syncthreads();
        {
            const int nt = (int)n_touched[par];
            for (int q = tid; q < nt; q += PART_THREADS) {
                const int b = touched[par][q];
                const uint32_t c = hist[b];
                hist[b] = 0;
                run0[b] = bucket_base[b] + atomicAdd(&bucket_cursor[b], c);
            }
            if (tid == 0) n_touched[par ^ 1] = 0;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PART_ITEMS; ++i) {
            const int j = i * PART_THREADS + tid;
            if (j < cnt) rec[(size_t)(run0[br[i] & 16383u] + (br[i] >> 14))] = make_uint2(node[i], rd[i]);
            slot[i] = slot_n[i];
            rd[i] = rd_n[i];
        }
    }
#undef PART_LOAD
}

// ---- pass 2 ----------------------------------------------------------------------------------------------
// sub-buckets of a list of n entries: one up to INC_SUB_MIN, else the power of two in (n/4, n/2]
__device__ __forceinline__ int inc_nsub(long long n) {
    return n <= INC_SUB_MIN ? 1 : 1 << (32 - __clz((int)((n - 1) >> 2)));
}

__device__ __forceinline__ uint32_t cur_get(const uint32_t *cur, int s) { return (cur[s >> 1] >> ((s & 1) << 4)) & 0xFFFFu; }

// ascending sort of x[0 .. m) by one warp, any m: bitonic merges of doubling size in which every comparison
// points the same way, so the slots past m behave as +infinity and are never touched
__device__ __forceinline__ void warp_sort_any(uint32_t *x, const int m, const int lane) {
    for (int k = 2; (k >> 1) < m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < m; i += 32) {
                const int p = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
                if (p > i && p < m) {
                    const uint32_t a = x[i], b = x[p];
                    if (a > b) {
                        x[i] = b;
                        x[p] = a;
                    }
                }
            }
            __syncwarp();
        }
    }
}

// one warp per 32 consecutive segments of [seg_lo, seg_hi), sorted in place in global memory (the body of
// k_segsort_main for a range of segments): the fallback of a unit that outgrows shared memory
__device__ __forceinline__ void segsort_range(const SegJob &J, const SegWork &work, const long long seg_lo, const long long seg_hi,
                                              const int warp, const int n_warps, const int lane) {
    unsigned long long my_dups = 0;
    for (long long base = seg_lo + (long long)warp * 32; base < seg_hi; base += (long long)n_warps * 32) {
        const long long s = base + lane;
        long long o = 0, n = 0;
        if (s < seg_hi) {
            o = J.off[s];
            n = J.off[s + 1] - o;
        }
        unsigned int d = 0;
        if (n >= 1 && n <= 8) {
            uint32_t v[8];
            uint32_t *x = J.a + o;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = i < n ? x[i] : SEG_PAD;
            sort8(v);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < n) {
                    x[i] = v[i];
                    if (i > 0 && v[i] == v[i - 1]) ++d;
                }
        } else if (n > SEG_BITONIC_MAX) {
            const bool big = n > SEG_WARP_MAX;
            const unsigned int pos = atomicAdd(&work.counters[big ? 1 : 0], 1u);
            if ((long long)pos < work.cap) work.list[big ? work.cap - 1 - pos : pos] = s;
        }
        unsigned int mid = __ballot_sync(0xffffffffu, n > 8 && n <= SEG_BITONIC_MAX);
        while (mid) {
            const int l = __ffs(mid) - 1;
            mid &= mid - 1;
            const long long so = __shfl_sync(0xffffffffu, o, l);
            const int sn = (int)__shfl_sync(0xffffffffu, n, l);
            uint32_t *x = J.a + so;
            unsigned int sd;
            if (sn <= 32) sd = warp_sort_segment<1>(x, x, sn, lane);
            else if (sn <= 64) sd = warp_sort_segment<2>(x, x, sn, lane);
            else if (sn <= 128) sd = warp_sort_segment<4>(x, x, sn, lane);
            else sd = warp_sort_segment<8>(x, x, sn, lane);
            if (lane == l) d = sd;
        }
        if (J.dups && s < seg_hi) J.dups[s] = d;
        my_dups += d;
    }
    if (J.total_dups) {
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) my_dups += __shfl_xor_sync(0xffffffffu, my_dups, dd);
        if (lane == 0 && my_dups) atomicAdd(J.total_dups, my_dups);
    }
}

// J: in-place job over `reads` (a = reads, b = scratch, off = reads_off, dups per node); work: the lists of the
// segments left to k_segsort_warp / k_segsort_radix
__global__ void __launch_bounds__(INC_THREADS, 2)
k_unit_lists(const uint2 *__restrict__ rec, const int *__restrict__ unit_lo, const UnitPlan plan, const SegJob J,
             const SegWork work) {
    extern __shared__ __align__(16) unsigned char u_smem[];
    uint32_t *vals0 = reinterpret_cast<uint32_t *>(u_smem);
    uint16_t *sub_off = reinterpret_cast<uint16_t *>(vals0 + INC_CAPV + 4);
    uint8_t *shf = reinterpret_cast<uint8_t *>(sub_off + INC_S + 8);
    uint32_t *cur = reinterpret_cast<uint32_t *>(shf + INC_S + 8);
    __shared__ uint32_t s_warp[INC_THREADS / 32];
    __shared__ uint32_t s_total;
    const int q = blockIdx.x;
    if (q >= plan.n_units) return;
    const int unit_node_lo = unit_lo[q], unit_node_hi = unit_lo[q + 1];
    if (unit_node_hi <= unit_node_lo) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t *const roff = J.off;
    // A unit that outgrows shared memory does so because of its LAST list (the others start inside one INC_C cell):
    // it is done in two parts, the last list on its own.
    const bool two_parts = roff[unit_node_hi] - roff[unit_node_lo] > INC_CAPV && unit_node_hi - unit_node_lo > 1;
  for (int part = 0; part < (two_parts ? 2 : 1); ++part) {
    const int node_lo = (two_parts && part == 1) ? unit_node_hi - 1 : unit_node_lo;
    const int node_hi = (two_parts && part == 0) ? unit_node_hi - 1 : unit_node_hi;
    const int nn = node_hi - node_lo;
    const long long base = roff[node_lo];
    const long long total = roff[node_hi] - base;
    __syncthreads();  // the previous part is done with shared memory
    // the bucket's region of the record array
    const int b = q >> plan.g;
    const long long rb = roff[unit_lo[b << plan.g]], re = roff[unit_lo[(b + 1) << plan.g]];
    const uint32_t nlo = (uint32_t)node_lo, unn = (uint32_t)nn;
    uint32_t *const out = J.a + base;
    const int pad = (int)(base & 3);  // entry j of the piece lives at vals0[pad + j]: same alignment as reads[base + j]
    uint32_t *const vals = vals0 + pad;

    if (total > INC_CAPV) {
        // ---- fallback: per-node cursors in shared memory, reads placed in global memory, segments sorted there
        uint32_t *c32 = vals0;
        for (int i = tid; i < nn; i += INC_THREADS) c32[i] = (uint32_t)(roff[node_lo + i] - base);
        __syncthreads();
        for (long long j = rb + tid; j < re; j += INC_THREADS) {
            const uint2 r = rec[j];
            const uint32_t i = r.x - nlo;
            if (i < unn) out[atomicAdd(&c32[i], 1u)] = r.y;
        }
        __syncthreads();
        segsort_range(J, work, node_lo, node_hi, warp, INC_THREADS / 32, lane);
        continue;
    }

    // ---- sub-bucket table: sub_off[i] = first sub-bucket of node i of the unit, shf[i] = read bits dropped to
    // get the sub-bucket of a read of node i (NPT nodes per thread)
    {
        constexpr int NPT = INC_S / INC_THREADS;
        int ns[NPT];
        uint32_t sum = 0;
#pragma unroll
        for (int u = 0; u < NPT; ++u) {
            const int i = tid * NPT + u;
            ns[u] = 0;
            if (i < nn) {
                ns[u] = inc_nsub(roff[node_lo + i + 1] - roff[node_lo + i]);
                if (J.dups) J.dups[node_lo + i] = 0;
            }
            sum += ns[u];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
        for (int ww = 0; ww < warp; ++ww) run += s_warp[ww];
#pragma unroll
        for (int u = 0; u < NPT; ++u) {
            const int i = tid * NPT + u;
            if (i < nn) {
                sub_off[i] = (uint16_t)run;
                // ns = 2^l sub-buckets: the top l bits of the scaled read (a funnel shift by 32 gives 0)
                shf[i] = (uint8_t)(32 - (31 - __clz(ns[u])));
            }
            run += ns[u];
        }
        if (tid == INC_THREADS - 1) {
            sub_off[nn] = (uint16_t)run;
            s_total = run;
        }
    }
    for (int i = tid; i < INC_CUR_WORDS; i += INC_THREADS) cur[i] = 0;
    __syncthreads();
    const int n_sub = (int)s_total;
    const uint32_t read_lo = plan.read_lo, rscale = plan.rscale;
    const uint2 *const recb = rec + rb;
    const int n_rec = (int)(re - rb);

    // ---- sweep 1: entries per sub-bucket (INC_BATCH records in flight per thread, the next batch loaded before the
    // current one is used)
#define INC_LOAD(dst, j0)                                                          \
    _Pragma("unroll") for (int u = 0; u < INC_BATCH; ++u) {                        \
        const int j_ = (j0) + u * INC_THREADS + tid;                               \
        dst[u] = j_ < n_rec ? recb[j_] : make_uint2(0xFFFFFFFFu, 0u);              \
    }
    {
        uint2 ra[INC_BATCH], rn[INC_BATCH];
        INC_LOAD(ra, 0);
        for (int j0 = 0; j0 < n_rec; j0 += INC_BATCH * INC_THREADS) {
            INC_LOAD(rn, j0 + INC_BATCH * INC_THREADS);
#pragma unroll
            for (int u = 0; u < INC_BATCH; ++u) {
                const uint32_t i = ra[u].x - nlo;
                if (i < unn) {
                    const uint32_t sb = (uint32_t)sub_off[i] + __funnelshift_rc((ra[u].y - read_lo) * rscale, 0u, shf[i]);
                    atomicAdd(&cur[sb >> 1], 1u << ((sb & 1u) << 4));
                }
            }
#pragma unroll
            for (int u = 0; u < INC_BATCH; ++u) ra[u] = rn[u];
        }
    }
    __syncthreads();
    // ---- exclusive scan of the counts -> first entry of every sub-bucket (whole words per thread)
    {
        const int n_words = (n_sub + 1) >> 1;
        const int per = (n_words + INC_THREADS - 1) / INC_THREADS;
        const int w_lo = min(n_words, tid * per), w_hi = min(n_words, w_lo + per);
        uint32_t sum = 0;
        for (int w = w_lo; w < w_hi; ++w) {
            const uint32_t x = cur[w];
            sum += (x & 0xFFFFu) + (x >> 16);
        }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        __syncthreads();  // s_warp is reused
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
        for (int ww = 0; ww < warp; ++ww) run += s_warp[ww];
        for (int w = w_lo; w < w_hi; ++w) {
            const uint32_t x = cur[w];
            const uint32_t c0 = x & 0xFFFFu, c1 = x >> 16;
            cur[w] = run | ((run + c0) << 16);
            run += c0 + c1;
        }
    }
    __syncthreads();
    // ---- sweep 2: place the reads (the cursors end up at the END of their sub-buckets)
    {
        uint2 ra[INC_BATCH], rn[INC_BATCH];
        INC_LOAD(ra, 0);
        for (int j0 = 0; j0 < n_rec; j0 += INC_BATCH * INC_THREADS) {
            INC_LOAD(rn, j0 + INC_BATCH * INC_THREADS);
#pragma unroll
            for (int u = 0; u < INC_BATCH; ++u) {
                const uint32_t i = ra[u].x - nlo;
                if (i < unn) {
                    const uint32_t sb = (uint32_t)sub_off[i] + __funnelshift_rc((ra[u].y - read_lo) * rscale, 0u, shf[i]);
                    const uint32_t sh = (sb & 1u) << 4;
                    const uint32_t old = atomicAdd(&cur[sb >> 1], 1u << sh);
                    vals[(old >> sh) & 0xFFFFu] = ra[u].y;
                }
            }
#pragma unroll
            for (int u = 0; u < INC_BATCH; ++u) ra[u] = rn[u];
        }
    }
#undef INC_LOAD
    __syncthreads();
    // ---- sort the sub-buckets: lane <-> sub-bucket, up to 8 entries in registers, larger ones by the warp.  Equal
    // neighbours (windows of one read) always share a sub-bucket: counted here, per node.
    unsigned int my_dups = 0;
    for (int s0 = warp * 32; s0 < n_sub; s0 += (INC_THREADS / 32) * 32) {
        const int sidx = s0 + lane;
        int lo = 0, m = 0;
        if (sidx < n_sub) {
            lo = sidx ? (int)cur_get(cur, sidx - 1) : 0;
            m = (int)cur_get(cur, sidx) - lo;
        }
        if (!__any_sync(0xffffffffu, m > 1)) continue;
        unsigned int d = 0;
        if (m > 1 && m <= 8) {
            uint32_t v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = i < m ? vals[lo + i] : SEG_PAD;
            sort8(v);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < m) {
                    vals[lo + i] = v[i];
                    if (i > 0 && v[i] == v[i - 1]) ++d;
                }
        }
        unsigned int big = __ballot_sync(0xffffffffu, m > 8);
        while (big) {
            const int l = __ffs(big) - 1;
            big &= big - 1;
            const int blo = __shfl_sync(0xffffffffu, lo, l), bm = __shfl_sync(0xffffffffu, m, l);
            uint32_t *x = vals + blo;
            unsigned int sd;
            if (bm <= 32) sd = warp_sort_segment<1>(x, x, bm, lane);
            else if (bm <= 64) sd = warp_sort_segment<2>(x, x, bm, lane);
            else if (bm <= 128) sd = warp_sort_segment<4>(x, x, bm, lane);
            else if (bm <= 256) sd = warp_sort_segment<8>(x, x, bm, lane);
            else {
                warp_sort_any(x, bm, lane);
                sd = 0;
                for (int i = 1 + lane; i < bm; i += 32) sd += x[i] == x[i - 1];
#pragma unroll
                for (int dd = 16; dd > 0; dd >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, dd);
            }
            if (lane == l) d = sd;
            __syncwarp();
        }
        if (d && J.dups) {
            // rare: the node of this sub-bucket = the last one whose first sub-bucket is <= sidx
            int a = 0, z = nn - 1;
            while (a < z) {
                const int mid = (a + z + 1) >> 1;
                if ((int)sub_off[mid] <= sidx) a = mid;
                else z = mid - 1;
            }
            atomicAdd(&J.dups[node_lo + a], d);
            my_dups += d;
        }
    }
    if (my_dups && J.total_dups) atomicAdd(J.total_dups, (unsigned long long)my_dups);
    __syncthreads();
    // ---- the finished piece leaves coalesced: vals is shifted so that 16-byte alignment in shared memory and in
    // `reads` coincide
    {
        const int n = (int)total;
        const int first = (4 - pad) & 3;  // first entry whose global address is 16-byte aligned
        for (int j = tid; j < min(first, n); j += INC_THREADS) out[j] = vals[j];
        const int n4 = n > first ? (n - first) >> 2 : 0;
        for (int q4 = tid; q4 < n4; q4 += INC_THREADS) {
            const int j = first + 4 * q4;
            *reinterpret_cast<uint4 *>(out + j) = *reinterpret_cast<const uint4 *>(vals + j);
        }
        for (int j = first + 4 * n4 + tid; j < n; j += INC_THREADS) out[j] = vals[j];
    }
  }
}

}  // namespace amira
