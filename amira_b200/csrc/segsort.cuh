// segsort.cuh -- ascending sort of every segment of a CSR value array.
//
// Two users in the GeneMerGraph build, both the second half of a counting-sort "transpose":
//   node -> edges    forward / backward edge lists in edge creation order (construct_graph.py:287-298):
//                    k_segsort_main over the 2N adjacency segments.
//   node -> reads    Node.listOfReads is the ascending list of the reads that touch the node
//                    (construct_node.py:64-67).  incidence.cuh sorts these lists in shared memory; a list too
//                    long for that is placed in global memory and handed to the work-list kernels here
//                    (k_segsort_warp, k_segsort_radix), and a whole unit that outgrows shared memory runs the
//                    loop of k_segsort_main over its own segments.  Equal neighbours after the sort are
//                    windows of one read (counted per node, removed lazily -- they are rare).
//
// Segment sizes span five orders of magnitude (coverage-1 error nodes .. nodes on every read), so:
//   n <= 1          copy
//   n <= 8          one THREAD: 19-comparator network in registers (32 consecutive segments per warp)
//   n <= 256        one WARP, inline: bitonic network, 1..8 keys per lane in registers, shuffles for the
//                   lane-crossing stages
//   n <= 4096       one WARP, from a work list (k_segsort_warp): bucket sort -- the keys are dealt into ~n/4
//                   equal-width value buckets, every lane sorts whole buckets with the 8-key network
//   larger          one CTA (k_segsort_radix): LSD radix sort, 8 warps on contiguous chunks, `match.any`
//                   ranking, ping-pong between the two global buffers A and B (any size)
// Source and destination may differ (out of place: the source segments may even be laid out in a different
// order, `a_start`), which decides the parity of the radix pass count: the last pass must land in the
// destination.  Values are < 0xFFFFFFFF (read / edge indices are < 2^31); 0xFFFFFFFF pads.
#pragma once

#include "common.cuh"

namespace amira {

constexpr int SEG_BITONIC_MAX = 256;
constexpr int SEG_WARP_MAX = 4096;   // largest segment one warp bucket-sorts
constexpr int SEG_STAGE_MAX = 8192;  // largest segment the radix kernel stages in shared memory
constexpr int SEG_RADIX_THREADS = 256;
constexpr int SEG_RADIX_WARPS = SEG_RADIX_THREADS / 32;
constexpr int SEG_MAX_DIGIT_BITS = 11;
constexpr uint32_t SEG_PAD = 0xFFFFFFFFu;

// What to sort.  Segment s has n = off[s+1] - off[s] keys; it is read from a + (a_start ? a_start[s] : off[s])
// and must end up at dst + off[s], where dst = (passes odd) ? b : a.
struct SegJob {
    uint32_t *a;
    uint32_t *b;
    const int64_t *off;
    const uint32_t *a_start;  // nullable
    const long long *n_seg_ptr;
    int seg_mul;
    int passes;      // radix passes (parity: see above)
    int digit_bits;  // bits per radix pass
    uint32_t *dups;                  // nullable: equal neighbours per segment after the sort
    unsigned long long *total_dups;  // nullable
};

// work lists of the segments the main kernel defers to the radix kernel: [0] = number of warp-sized
// segments, [1] = number of larger ones, then the segment ids (warp-sized from the front, larger from the back)
struct SegWork {
    unsigned int *counters;
    long long *list;
    long long cap;
};

__device__ __forceinline__ void ce(uint32_t &a, uint32_t &b) {
    const uint32_t lo = min(a, b), hi = max(a, b);
    a = lo;
    b = hi;
}

// optimal 19-comparator network for 8 keys
__device__ __forceinline__ void sort8(uint32_t (&v)[8]) {
    ce(v[0], v[1]); ce(v[2], v[3]); ce(v[4], v[5]); ce(v[6], v[7]);
    ce(v[0], v[2]); ce(v[1], v[3]); ce(v[4], v[6]); ce(v[5], v[7]);
    ce(v[1], v[2]); ce(v[5], v[6]); ce(v[0], v[4]); ce(v[3], v[7]);
    ce(v[1], v[5]); ce(v[2], v[6]);
    ce(v[1], v[4]); ce(v[3], v[6]);
    ce(v[2], v[4]); ce(v[3], v[5]);
    ce(v[3], v[4]);
}

// bitonic sort of 32 * IPL keys held striped over a warp: element e = item * 32 + lane
template <int IPL>
__device__ __forceinline__ void warp_bitonic(uint32_t (&v)[IPL], const int lane) {
#pragma unroll
    for (int k = 2; k <= 32 * IPL; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int dj = j >> 5;
#pragma unroll
                for (int it = 0; it < IPL; ++it) {
                    if ((it & dj) == 0) {
                        const bool asc = (it & (k >> 5)) == 0;
                        const uint32_t a = v[it], b = v[it | dj];
                        const bool sw = asc ? (a > b) : (a < b);
                        v[it] = sw ? b : a;
                        v[it | dj] = sw ? a : b;
                    }
                }
            } else {
#pragma unroll
                for (int it = 0; it < IPL; ++it) {
                    const uint32_t o = __shfl_xor_sync(0xffffffffu, v[it], j);
                    const bool asc = (k >= 32) ? ((it & (k >> 5)) == 0) : ((lane & k) == 0);
                    const bool lower = (lane & j) == 0;
                    v[it] = (lower == asc) ? min(v[it], o) : max(v[it], o);
                }
            }
        }
    }
}

template <int IPL>
__device__ __forceinline__ unsigned int warp_sort_segment(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst,
                                                          const int n, const int lane) {
    uint32_t v[IPL];
#pragma unroll
    for (int it = 0; it < IPL; ++it) {
        const int e = it * 32 + lane;
        v[it] = e < n ? src[e] : SEG_PAD;
    }
    warp_bitonic<IPL>(v, lane);
    unsigned int dup = 0;
#pragma unroll
    for (int it = 0; it < IPL; ++it) {
        const int e = it * 32 + lane;
        // predecessor of element e: lane - 1 of the same item, or lane 31 of the previous item
        uint32_t prev = __shfl_up_sync(0xffffffffu, v[it], 1);
        if (it > 0) {
            const uint32_t last = __shfl_sync(0xffffffffu, v[it - 1], 31);
            if (lane == 0) prev = last;
        }
        if (e < n) {
            dst[e] = v[it];
            if (e > 0 && prev == v[it]) ++dup;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) dup += __shfl_xor_sync(0xffffffffu, dup, d);
    return dup;
}

// Main pass: one warp per 32 consecutive segments; segments above SEG_BITONIC_MAX go to the work lists.
__global__ void __launch_bounds__(256) k_segsort_main(const SegJob J, const SegWork work) {
    const long long n_seg = *J.n_seg_ptr * J.seg_mul;
    const int lane = threadIdx.x & 31;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    uint32_t *const dst_base = (J.passes & 1) ? J.b : J.a;
    unsigned long long my_dups = 0;
    for (long long base = ((((long long)blockIdx.x * blockDim.x) + threadIdx.x) >> 5) * 32; base < n_seg; base += n_warps * 32) {
        const long long s = base + lane;
        long long o = 0, n = 0, sa = 0;
        if (s < n_seg) {
            o = J.off[s];
            n = J.off[s + 1] - o;
            sa = J.a_start ? (long long)J.a_start[s] : o;
        }
        unsigned int d = 0;
        if (n >= 1 && n <= 2) {
            // most adjacency lists: one or two edges
            const uint32_t *src = J.a + sa;
            uint32_t *dst = dst_base + o;
            uint32_t x = src[0], y = n == 2 ? src[1] : SEG_PAD;
            ce(x, y);
            if (n == 2) {
                dst[0] = x;
                dst[1] = y;
                d = x == y;
            } else if (dst != src) {
                dst[0] = x;
            }
        } else if (n >= 3 && n <= 8) {
            uint32_t v[8];
            const uint32_t *src = J.a + sa;
            uint32_t *dst = dst_base + o;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = i < n ? src[i] : SEG_PAD;
            sort8(v);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < n) {
                    dst[i] = v[i];
                    if (i > 0 && v[i] == v[i - 1]) ++d;
                }
        } else if (n > SEG_BITONIC_MAX) {
            const bool big = n > SEG_WARP_MAX;
            const unsigned int pos = atomicAdd(&work.counters[big ? 1 : 0], 1u);
            if ((long long)pos < work.cap) work.list[big ? work.cap - 1 - pos : pos] = s;
        }
        // segments of 9..256 keys: the whole warp sorts them one after the other
        unsigned int mid = __ballot_sync(0xffffffffu, n > 8 && n <= SEG_BITONIC_MAX);
        while (mid) {
            const int l = __ffs(mid) - 1;
            mid &= mid - 1;
            const long long so = __shfl_sync(0xffffffffu, o, l), ssa = __shfl_sync(0xffffffffu, sa, l);
            const int sn = (int)__shfl_sync(0xffffffffu, n, l);
            const uint32_t *src = J.a + ssa;
            uint32_t *dst = dst_base + so;
            unsigned int sd;
            if (sn <= 32) sd = warp_sort_segment<1>(src, dst, sn, lane);
            else if (sn <= 64) sd = warp_sort_segment<2>(src, dst, sn, lane);
            else if (sn <= 128) sd = warp_sort_segment<4>(src, dst, sn, lane);
            else sd = warp_sort_segment<8>(src, dst, sn, lane);
            if (lane == l) d = sd;
        }
        if (J.dups && s < n_seg) J.dups[s] = d;  // deferred segments are overwritten by the radix kernel
        my_dups += d;
    }
    if (J.total_dups) {
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) my_dups += __shfl_xor_sync(0xffffffffu, my_dups, dd);
        if (lane == 0 && my_dups) atomicAdd(J.total_dups, my_dups);
    }
}

// Segments of 257 .. SEG_WARP_MAX keys from the first work list, one warp each: BUCKET sort.  A node's reads are
// spread evenly over its value range, so the keys are dealt (shared-memory counters, two passes) into
// K ~ n/4 equal-width value buckets of ~2-4 keys, and every lane then sorts whole buckets on its own with the
// 8-key register network: ~2 warp instructions per key, against ~20 for a bitonic network over the whole
// segment (measured: 330M warp instructions, 0.7 ms for the 16M keys of this class on the C5 shard).  Buckets
// above 8 keys (a few per cent) go through the warp networks; a segment so skewed that a bucket exceeds 256 keys
// is handed to the radix kernel.
constexpr int SEG_BUCKETS = 1024;
__global__ void __launch_bounds__(128) k_segsort_warp(const SegJob J, const SegWork work) {
    __shared__ unsigned int s_cnt[4][SEG_BUCKETS + 32];
    const int lane = threadIdx.x & 31;
    unsigned int *c = s_cnt[threadIdx.x >> 5];
    const bool dst_is_b = (J.passes & 1) != 0;
    const long long n_mid = min((long long)work.counters[0], work.cap);
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long my_dups = 0;
    for (long long w = (((long long)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; w < n_mid; w += n_warps) {
        const long long s = work.list[w];
        const long long o = J.off[s];
        const int n = (int)(J.off[s + 1] - o);
        const uint32_t *const ga = J.a + (J.a_start ? (long long)J.a_start[s] : o);
        uint32_t *const gb = J.b + o;
        uint32_t *const dst = dst_is_b ? gb : J.a + o;  // in place: a_start is null, the source is the destination
        // value range of the segment
        uint32_t mn = 0xFFFFFFFFu, mx = 0u;
        for (int i = lane; i < n; i += 32) {
            const uint32_t x = ga[i];
            mn = min(mn, x);
            mx = max(mx, x);
        }
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
        // K = 2^lgK buckets with n/4 <= K < n/2 (at most SEG_BUCKETS); bucket(x) = (x - mn) >> shift with the smallest
        // shift that stays below K
        int lgK = 6;
        while ((4 << lgK) < n && (1 << lgK) < SEG_BUCKETS) ++lgK;
        const int K = 1 << lgK;
        int shift = 0;
        while (((mx - mn) >> shift) >= (uint32_t)K) ++shift;
        for (int i = lane; i <= K; i += 32) c[i] = 0;
        __syncwarp();
        for (int i = lane; i < n; i += 32) atomicAdd(&c[1 + ((ga[i] - mn) >> shift)], 1u);
        __syncwarp();
        // inclusive scan of c[0..K]: c[b] becomes the start of bucket b (c[0] = 0); each lane owns K/32 consecutive
        // counters (+ the last lane the extra one)
        unsigned int biggest = 0;
        {
            const int per = K >> 5;
            unsigned int sum = 0;
            for (int i = 0; i < per; ++i) {
                const unsigned int v = c[1 + lane * per + i];
                biggest = max(biggest, v);
                sum += v;
            }
            unsigned int incl = sum;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, incl, dd);
                if (lane >= dd) incl += t;
            }
            unsigned int run = incl - sum;
            for (int i = 0; i < per; ++i) {
                run += c[1 + lane * per + i];
                c[1 + lane * per + i] = run;
            }
        }
        if (__any_sync(0xffffffffu, biggest > 256u)) {
            if (lane == 0) {
                const unsigned int pos = atomicAdd(&work.counters[1], 1u);
                if ((long long)pos < work.cap) work.list[work.cap - 1 - pos] = s;
            }
            continue;
        }
        __syncwarp();
        // deal the keys: the running start of bucket b is c[b]; afterwards c[b] is the END of bucket b
        for (int i = lane; i < n; i += 32) {
            const uint32_t x = ga[i];
            gb[atomicAdd(&c[(x - mn) >> shift], 1u)] = x;
        }
        __syncwarp();
        unsigned int d = 0;
        for (int b0 = 0; b0 < K; b0 += 32) {
            const int b = b0 + lane;
            const int start = b ? (int)c[b - 1] : 0;
            const int m = (int)c[b] - start;
            if (m >= 1 && m <= 8) {
                uint32_t v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = i < m ? gb[start + i] : SEG_PAD;
                sort8(v);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i < m) {
                        dst[start + i] = v[i];
                        if (i > 0 && v[i] == v[i - 1]) ++d;
                    }
            }
            unsigned int big = __ballot_sync(0xffffffffu, m > 8);
            while (big) {
                const int l = __ffs(big) - 1;
                big &= big - 1;
                const int bs = __shfl_sync(0xffffffffu, start, l), bm = __shfl_sync(0xffffffffu, m, l);
                unsigned int sd;
                if (bm <= 32) sd = warp_sort_segment<1>(gb + bs, dst + bs, bm, lane);
                else if (bm <= 64) sd = warp_sort_segment<2>(gb + bs, dst + bs, bm, lane);
                else if (bm <= 128) sd = warp_sort_segment<4>(gb + bs, dst + bs, bm, lane);
                else sd = warp_sort_segment<8>(gb + bs, dst + bs, bm, lane);
                if (lane == 0) d += sd;
            }
        }
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) d += __shfl_xor_sync(0xffffffffu, d, dd);
        if (J.dups && lane == 0) J.dups[s] = d;
        if (lane == 0) my_dups += d;
        __syncwarp();
    }
    if (J.total_dups && lane == 0 && my_dups) atomicAdd(J.total_dups, my_dups);
}

// ---- LSD radix sort ----------------------------------------------------------------------------------
// One pass of one warp over keys src[0 .. n) (in order): stable scatter into dst by digit, with the
// warp's running digit offsets in cnt[] (shared memory; on entry the exclusive start of every digit).
__device__ __forceinline__ void radix_scatter(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, const long long lo,
                                              const long long hi, const int shift, const uint32_t mask,
                                              unsigned int *cnt, const int lane) {
    const unsigned int lt = (1u << lane) - 1u;
    for (long long i0 = lo; i0 < hi; i0 += 32) {
        const long long i = i0 + lane;
        const bool valid = i < hi;
        const uint32_t x = valid ? src[i] : 0u;
        const uint32_t d = valid ? ((x >> shift) & mask) : SEG_PAD;
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        unsigned int start = 0;
        if (valid) start = cnt[d];
        __syncwarp();
        if (valid) {
            dst[start + __popc(peers & lt)] = x;
            if ((peers & lt) == 0) cnt[d] = start + __popc(peers);  // first lane of the group moves the offset on
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void radix_count(const uint32_t *__restrict__ src, const long long lo, const long long hi,
                                            const int shift, const uint32_t mask, unsigned int *cnt, const int lane) {
    const unsigned int lt = (1u << lane) - 1u;
    for (long long i0 = lo; i0 < hi; i0 += 32) {
        const long long i = i0 + lane;
        const bool valid = i < hi;
        const uint32_t d = valid ? ((src[i] >> shift) & mask) : SEG_PAD;
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        if (valid && (peers & lt) == 0) cnt[d] += __popc(peers);
        __syncwarp();
    }
}

__device__ __forceinline__ unsigned int count_dups(const uint32_t *__restrict__ x, const long long lo, const long long hi,
                                                   const int lane_or_tid, const int stride) {
    unsigned int d = 0;
    for (long long i = lo + lane_or_tid; i < hi; i += stride)
        if (i > 0 && x[i] == x[i - 1]) ++d;
    return d;
}

// Shared memory: [SEG_RADIX_WARPS][D] digit counters, then two key buffers of kcap keys each (kcap may be 0).
// A segment of n <= kcap keys is staged there and its passes run at shared-memory latency; larger segments
// ping-pong between the two global buffers.
__global__ void __launch_bounds__(SEG_RADIX_THREADS) k_segsort_radix(const SegJob J, const SegWork work, const int kcap) {
    extern __shared__ unsigned int s_mem[];
    __shared__ unsigned int s_warp[SEG_RADIX_WARPS];
    __shared__ unsigned int s_dup;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = 1 << J.digit_bits;
    const uint32_t mask = (uint32_t)D - 1u;
    unsigned int *s_cnt = s_mem;
    unsigned int *cnt = s_cnt + warp * D;
    uint32_t *kb0 = s_mem + SEG_RADIX_WARPS * D, *kb1 = kb0 + kcap;
    uint32_t *const final_base = (J.passes & 1) ? J.b : J.a;
    const long long n_mid = min((long long)work.counters[0], work.cap);
    const long long n_big = min((long long)work.counters[1], work.cap - n_mid);

    // ---- larger segments: the whole CTA, warp w on the w-th contiguous chunk
    for (long long w = blockIdx.x; w < n_big; w += gridDim.x) {
        const long long s = work.list[work.cap - 1 - w];
        const long long o = J.off[s], n = J.off[s + 1] - o;
        uint32_t *ga = J.a + (J.a_start ? (long long)J.a_start[s] : o), *gb = J.b + o;
        const bool staged = n <= kcap;
        uint32_t *x = staged ? kb0 : ga, *y = staged ? kb1 : gb;
        if (staged) {
            for (long long i = threadIdx.x; i < n; i += SEG_RADIX_THREADS) x[i] = ga[i];
            __syncthreads();
        }
        const long long chunk = ((n + SEG_RADIX_WARPS - 1) / SEG_RADIX_WARPS + 31) & ~31ll;
        const long long lo = min(n, warp * chunk), hi = min(n, lo + chunk);
        for (int pass = 0; pass < J.passes; ++pass) {
            const int shift = pass * J.digit_bits;
            for (int i = lane; i < D; i += 32) cnt[i] = 0;
            __syncwarp();
            radix_count(x, lo, hi, shift, mask, cnt, lane);
            __syncthreads();
            // digit d, warp w starts at (keys with smaller digits) + (keys with digit d in earlier warps):
            // thread t owns the digits t * per .. + per - 1
            {
                const int per = max(1, D / SEG_RADIX_THREADS);
                const int d0 = threadIdx.x * per;
                unsigned int sum = 0;
                if (d0 < D)
                    for (int i = 0; i < per; ++i)
                        for (int ww = 0; ww < SEG_RADIX_WARPS; ++ww) sum += s_cnt[ww * D + d0 + i];
                unsigned int incl = sum;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, dd);
                    if (lane >= dd) incl += t;
                }
                if (lane == 31) s_warp[warp] = incl;
                __syncthreads();
                unsigned int run = incl - sum;
                for (int ww = 0; ww < warp; ++ww) run += s_warp[ww];
                if (d0 < D)
                    for (int i = 0; i < per; ++i)
                        for (int ww = 0; ww < SEG_RADIX_WARPS; ++ww) {
                            const unsigned int c = s_cnt[ww * D + d0 + i];
                            s_cnt[ww * D + d0 + i] = run;
                            run += c;
                        }
            }
            __syncthreads();
            radix_scatter(x, y, lo, hi, shift, mask, cnt, lane);
            __syncthreads();
            uint32_t *t = x;
            x = y;
            y = t;
        }
        // x holds the sorted keys
        if (threadIdx.x == 0) s_dup = 0;
        __syncthreads();
        unsigned int d = 0;
        if (staged) {
            uint32_t *dst = final_base + o;
            for (long long i = threadIdx.x; i < n; i += SEG_RADIX_THREADS) {
                const uint32_t v = x[i];
                dst[i] = v;
                if (i > 0 && x[i - 1] == v) ++d;
            }
        } else if (J.dups) {
            d = count_dups(x, 0, n, threadIdx.x, SEG_RADIX_THREADS);
        }
        if (J.dups) {
            if (d) atomicAdd(&s_dup, d);
            __syncthreads();
            if (threadIdx.x == 0) {
                J.dups[s] = s_dup;
                if (s_dup && J.total_dups) atomicAdd(J.total_dups, (unsigned long long)s_dup);
            }
        }
        __syncthreads();
    }
}

}  // namespace amira
