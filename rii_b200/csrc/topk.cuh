// Block-level top-k selection under the total order (distance asc, id asc).
//
// Replaces the reference's std::partial_sort call sites (src/rii.h:234-235, :279-280, :312-313), which
// compare distances only and leave exact ties to libstdc++ heap order; here a candidate is the 64-bit key
// (float bits of the distance << 32 | 32-bit local id).  ADC distances are sums of squares (>= +0), so
// the IEEE bit pattern is order preserving and one unsigned compare implements (dist, id).
//
// Scheme (per CTA): candidates that beat the current k-th key are appended to a shared-memory buffer with
// one shared atomic; when the buffer could overflow in the next round the CTA sorts it (bitonic, padded
// to the next power of two of the live count), keeps the k smallest and tightens the threshold.  After
// the first compaction the pass rate is ~k/n, so the steady-state cost per candidate is one FSETP.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
#define RII_KEY_MAX 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ u64 pack_key(float d, uint32_t id)
{
    return ((u64)__float_as_uint(d) << 32) | (u64)id;
}
__device__ __forceinline__ float key_dist(u64 k) { return __uint_as_float((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_id(u64 k) { return (uint32_t)(k & 0xFFFFFFFFull); }

__host__ __device__ inline int next_pow2(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

struct BlockTopk {
    u64 *keys;   // shared, capacity `cap` (power of two)
    int *count;  // shared
    u64 *thr;    // shared: current k-th best key (RII_KEY_MAX until k candidates were seen)
    int cap;
    int k;

    __device__ __forceinline__ void init()
    {
        if (threadIdx.x == 0) { *count = 0; *thr = RII_KEY_MAX; }
        __syncthreads();
    }
    // Upper 32 bits of the threshold = float bits of the k-th distance (0xFFFFFFFF while open): compare
    // as unsigned integers, never as floats (the open threshold is a NaN pattern).
    __device__ __forceinline__ uint32_t thr_hi() const { return (uint32_t)(*thr >> 32); }
    __device__ __forceinline__ u64 thr_key() const { return *thr; }

    // Any thread, any time between two barriers; the caller guarantees count + (pushes this round) <= cap.
    __device__ __forceinline__ void push(u64 key)
    {
        int pos = atomicAdd(count, 1);
        keys[pos] = key;
    }

    // All threads.  Sort the live keys, keep the k smallest, refresh the threshold.
    __device__ void compact()
    {
        __syncthreads();
        int n = *count;
        int P = next_pow2(n < 2 ? 2 : n);
        for (int i = n + threadIdx.x; i < P; i += blockDim.x) keys[i] = RII_KEY_MAX;
        __syncthreads();
        for (int kk = 2; kk <= P; kk <<= 1) {
            for (int j = kk >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < P; i += blockDim.x) {
                    int ixj = i ^ j;
                    if (ixj > i) {
                        u64 a = keys[i], b = keys[ixj];
                        bool up = (i & kk) == 0;
                        if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        }
        if (threadIdx.x == 0) {
            int c = n < k ? n : k;
            *count = c;
            *thr = (c == k) ? keys[k - 1] : RII_KEY_MAX;
        }
        __syncthreads();
    }

    // All threads, at a barrier-aligned point: make room for `incoming` more pushes.
    __device__ __forceinline__ void reserve(int incoming)
    {
        __syncthreads();
        int c = *count;
        __syncthreads();  // every thread has read `count` before anyone pushes again
        if (c + incoming > cap) compact();
    }
};
