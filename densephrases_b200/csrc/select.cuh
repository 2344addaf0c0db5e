// select.cuh -- block-wide radix select and bitonic sort on distinct u64 keys (larger key = better).
#pragma once
#include "common.cuh"

struct SelectScratch {     // lives in shared memory
    unsigned hist[256];
    unsigned digit, remaining, bincount, pad;
};

// Returns pivot P such that exactly `need` of the n keys are >= P.  Requires distinct keys, 1 <= need <= n.
// get(i) must be callable by every thread for i in [0,n).  All threads of the block must call this.
template <class Get>
__device__ __forceinline__ unsigned long long block_radix_select(Get get, int n, int need, SelectScratch* sc) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long prefix = 0, mask = 0;
    unsigned remaining = (unsigned)need;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += nt) sc->hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            unsigned long long k = get(i);
            if ((k & mask) == prefix) atomicAdd(&sc->hist[(unsigned)(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (warp == 0) {
            unsigned local[8], s = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) { local[j] = sc->hist[lane * 8 + j]; s += local[j]; }
            unsigned incl = s;   // -> sum over lanes >= lane
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                unsigned v = __shfl_down_sync(0xffffffffu, incl, off);
                if (lane + off < 32) incl += v;
            }
            unsigned above = incl - s;
            if (above < remaining && remaining <= incl) {
                unsigned c = above;
#pragma unroll
                for (int j = 7; j >= 0; j--) {
                    if (c + local[j] >= remaining) {
                        sc->digit = lane * 8 + j; sc->remaining = remaining - c; sc->bincount = local[j];
                        break;
                    }
                    c += local[j];
                }
            }
        }
        __syncthreads();
        prefix |= (unsigned long long)sc->digit << shift;
        mask |= 0xFFull << shift;
        remaining = sc->remaining;
        bool done = (sc->bincount == remaining);
        __syncthreads();
        if (done) break;
    }
    return prefix;
}

// In-place bitonic sort, DESCENDING, of n (power of two) u64 keys in shared memory. All threads call.
__device__ __forceinline__ void block_bitonic_sort_desc(unsigned long long* a, int n) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = tid; t < (n >> 1); t += nt) {
                int lo = 2 * t - (t & (stride - 1));   // index with bit `stride` cleared
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                unsigned long long x = a[lo], y = a[hi];
                if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
            }
        }
    }
    __syncthreads();
}
__device__ __forceinline__ int dph_next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
