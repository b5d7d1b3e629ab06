// scan.cuh -- device-wide exclusive prefix sums in ONE pass (decoupled look-back), hand-written for
// the GeneMerGraph build.  Every scan of the build is small next to the window arrays (reads, bitmap
// words, nodes, adjacency slots), so what matters is (a) one launch instead of three, (b) the element
// count may live in DEVICE memory (the host never learns the node / edge counts during a build, which
// keeps the whole build free of host synchronisation and capturable in a CUDA graph) and (c) the
// producer / consumer of the scanned values is fused in through the Load / Store functors (popcounts of
// the first-seen bitmaps, cursor copies, "last element = total" outputs).
//
// Tile status word: bits 63..62 = flag (0 = nothing yet, 1 = tile aggregate, 2 = inclusive prefix),
// bits 61..0 = value; flag and value travel in one 64-bit store, so no fence is needed between them.
// Tiles are taken in blockIdx order (the hardware dispatches CTAs in that order, which is what makes the
// look-back deadlock-free); the status words are cleared by a cudaMemsetAsync before every launch.
#pragma once

#include "common.cuh"

namespace amira {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr unsigned long long SCAN_VAL_MASK = (1ull << 62) - 1ull;

__device__ __forceinline__ unsigned long long scan_ld_state(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void scan_st_state(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Exclusive scan over items 0 .. n (n + 1 items: item n has value 0, so store(n, total, 0) hands out the
// total).  n = n_ptr ? *n_ptr * n_mul : n_imm.  load(i) -> unsigned long long (< 2^62 in total); store(i, excl, val).
template <typename Load, typename Store>
__global__ void __launch_bounds__(SCAN_THREADS) k_exscan(Load load, Store store, const long long *__restrict__ n_ptr,
                                                         long long n_mul, long long n_imm,
                                                         unsigned long long *__restrict__ state) {
    const long long n = (n_ptr ? *n_ptr * n_mul : n_imm) + 1;
    const long long tile = blockIdx.x;
    const long long base = tile * SCAN_TILE;
    if (base >= n) return;
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    __shared__ unsigned long long s_excl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // blocked arrangement: thread t owns items base + t*ITEMS .. +ITEMS-1
    unsigned long long v[SCAN_ITEMS];
    unsigned long long sum = 0;
    const long long i0 = base + (long long)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        const long long i = i0 + j;
        v[j] = (i < n - 1) ? load(i) : 0ull;
        sum += v[j];
    }
    // block-wide exclusive scan of the thread sums
    unsigned long long incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long warp_off = 0, block_sum = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        const unsigned long long x = s_warp[w];
        if (w < warp) warp_off += x;
        block_sum += x;
    }
    // publish the aggregate, look back for the exclusive prefix of this tile (warp 0)
    if (warp == 0) {
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) scan_st_state(&state[0], (2ull << 62) | block_sum);
        } else {
            if (lane == 0) scan_st_state(&state[tile], (1ull << 62) | block_sum);
            long long idx = tile - 1 - lane;  // lane 0 looks at the nearest predecessor
            while (true) {
                unsigned long long st = (idx >= 0) ? scan_ld_state(&state[idx]) : (2ull << 62);
                while (__any_sync(0xffffffffu, (st >> 62) == 0ull)) {
                    if ((st >> 62) == 0ull) st = scan_ld_state(&state[idx]);
                }
                const unsigned int incl_mask = __ballot_sync(0xffffffffu, (st >> 62) == 2ull);
                // sum the values of the lanes up to and including the first inclusive prefix
                const int stop = incl_mask ? (__ffs(incl_mask) - 1) : 31;
                unsigned long long x = (lane <= stop) ? (st & SCAN_VAL_MASK) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
                excl += x;
                if (incl_mask) break;
                idx -= 32;
            }
            if (lane == 0) scan_st_state(&state[tile], (2ull << 62) | ((excl + block_sum) & SCAN_VAL_MASK));
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    unsigned long long run = s_excl + warp_off + (incl - sum);
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        const long long i = i0 + j;
        if (i < n) store(i, run, v[j]);
        run += v[j];
    }
}

}  // namespace amira
