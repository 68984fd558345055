// Small index-build kernels of the device-side posting-list construction (src/rii.h:335-359: ids are appended to the
// list of their nearest center in ascending order).  Included by the host translation unit only.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// vals[i] = start + i
__global__ void k_iota_u32(uint32_t *vals, long long n, uint32_t start)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vals[i] = start + (uint32_t)i;
}

// bounds[i] = first position of the ascending `keys` (n) whose key is >= i, for i = 0..nlist  (per-list segments of the
// sorted (list, id) pairs; keys >= nlist -- invalid assignments -- end up beyond bounds[nlist])
__global__ void k_list_bounds(const uint32_t *__restrict__ keys, long long n, int nlist, long long *__restrict__ bounds)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nlist) return;
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (keys[mid] < (uint32_t)i) lo = mid + 1; else hi = mid;
    }
    bounds[i] = lo;
}

// new list i = old list i followed by the sorted new ids of list i.  grid (nlist)
__global__ void k_merge_lists(const long long *__restrict__ old_off, const int *__restrict__ old_ids, const long long *__restrict__ add_bounds,
                              const uint32_t *__restrict__ add_ids, const long long *__restrict__ new_off, int *__restrict__ new_ids)
{
    const int i = blockIdx.x;
    const long long o0 = old_off ? old_off[i] : 0, o1 = old_off ? old_off[i + 1] : 0, a0 = add_bounds[i], a1 = add_bounds[i + 1];
    int *dst = new_ids + new_off[i];
    for (long long j = threadIdx.x; j < o1 - o0; j += blockDim.x) dst[j] = old_ids[o0 + j];
    dst += o1 - o0;
    for (long long j = threadIdx.x; j < a1 - a0; j += blockDim.x) dst[j] = (int)add_ids[a0 + j];
}

// assign[ids[p]] = list holding position p (state import: rebuild the row -> list map from the CSR)
__global__ void k_assign_from_csr(const long long *__restrict__ off, const int *__restrict__ ids, int nlist, long long total, int *__restrict__ assign)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    int lo = 0, hi = nlist - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= p) lo = mid; else hi = mid - 1;
    }
    assign[ids[p]] = lo;
}

// max over the bit patterns of n non-negative floats (NaN / inf patterns compare above every finite value)
__global__ void k_max_bits(const float *__restrict__ x, long long n, unsigned int *out)
{
    unsigned int m = 0u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned int v = __float_as_uint(x[i]) & 0x7fffffffu;
        m = v > m ? v : m;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned int y = __shfl_xor_sync(0xffffffffu, m, o);
        m = y > m ? y : m;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}
