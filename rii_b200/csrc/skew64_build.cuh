// Index-build kernel of the skew64 layout (included by the host translation unit only).
#pragma once
#include "skew64.cuh"

// Build (a range of) skew64 segments.  codes: (rows, M) by id, padded with zeros to RB = 32 or 64 bytes per row; ids / offsets: CSR of the segments (null: ONE segment
// = rows [0, n_single) in id order); skew_off: (nseg + 1) first physical row (32-byte unit) of every segment, a
// multiple of 64.  One thread per 16-byte chunk.
__global__ void k_skew64_build(const uint8_t *__restrict__ codes, const int *__restrict__ ids, const long long *__restrict__ offsets,
                               const long long *__restrict__ skew_off, int nseg, long long n_single, long long prow0,
                               long long prow1, uint8_t *__restrict__ out, int M, int RB)
{
    const long long i = prow0 * 2 + (long long)blockIdx.x * blockDim.x + threadIdx.x;  // 16-byte chunk of the table
    const long long prow = i >> 1;
    if (prow >= prow1 || prow >= skew_off[nseg]) return;  // (prow1 may be an upper bound when the offsets only exist on the device)
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (skew_off[mid] <= prow) lo = mid; else hi = mid - 1;
    }
    const long long c16 = i - skew_off[lo] * 2;  // chunk within the segment: block b, quarter q, lane l
    const long long b = c16 >> 7;
    const int q = (int)(c16 >> 5) & 3, l = (int)(c16 & 31);
    const int s = (q >> 1) * 32 + l, half = q & 1, lag = l;
    const long long len = offsets ? offsets[lo + 1] - offsets[lo] : n_single;
    const long long ioff = offsets ? offsets[lo] : 0;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const long long x = 32 * b - lag + 16 * half + j;  // byte of stream s
        if (x >= 0) {
            const long long r = 64 * (x / RB) + s;         // row of the segment
            const int bb = (int)(x % RB);                  // byte of the padded row
            if (r < len && bb < M) {
                const long long id = ids ? (long long)ids[ioff + r] : r;
                w[j >> 2] |= (uint32_t)__ldg(codes + id * M + bb) << (8 * (j & 3));
            }
        }
    }
    reinterpret_cast<uint4 *>(out)[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

