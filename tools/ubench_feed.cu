// Micro-benchmark (tools only): how should the code bytes be fed to the lookup loop?
// The lookup loop itself (2 streams per lane, PRMT + LDS + FFMA2 pairs, conflict-free 64 KB table) runs at the
// shared-memory limit when its code words come from registers (tools/ubench_step.cu: 24.8 lookups/clk/SM).  The real
// scan kernels (v2: LDGSTS staging, v3: 3-stage LDGSTS ring, v4: LDG.256 into registers) all stop at ~16
// lookups/clk/SM.  This benchmark adds the feed of 1 code byte per lookup in different ways and measures the
// lookup rate:
//   0  no feed (code words synthesised in registers)
//   1  ld.global.nc.L1::no_allocate.v8.b32 into registers, 3 blocks ahead          (the v4 engine)
//   2  ld.global.nc.v8.b32 (L1 allocating) into registers, 3 blocks ahead
//   3  cp.async.bulk (TMA, 2 KB per warp and block) into a 4-stage shared ring + mbarrier, lanes read LDS.128
//   4  cp.async 16 B (LDGSTS) into a 4-stage shared ring, lanes read LDS.128
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_feed.bin tools/ubench_feed.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define NW 8
#define STEP(WX, WY, BYTE, T)                                                                                 \
    {                                                                                                         \
        const uint32_t ax_ = __byte_perm(WX, colreg, 0x7604 | ((BYTE) << 4));                                 \
        const uint32_t ay_ = __byte_perm(WY, colreg, 0x7604 | ((BYTE) << 4));                                 \
        float vx_, vy_;                                                                                       \
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(vx_) : "r"(ax_), "n"(4 * (T)));                     \
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(vy_) : "r"(ay_), "n"(4 * (T)));                     \
        asm("{.reg .b64 v, kk, ss; mov.b64 v, {%4, %5}; mov.b64 kk, {%2, %2}; mov.b64 ss, {%3, %3};"         \
            " fma.rn.f32x2 %1, %0, ss, %1; fma.rn.f32x2 %0, %0, kk, v;}"                                     \
            : "+l"(acc2), "+l"(out2)                                                                          \
            : "f"(keep[T]), "f"(sel[T]), "f"(vx_), "f"(vy_));                                                 \
    }
#define WORD(BX, BY, Q) STEP(BX[Q], BY[Q], 0, 4 * (Q)) STEP(BX[Q], BY[Q], 1, 4 * (Q) + 1) STEP(BX[Q], BY[Q], 2, 4 * (Q) + 2) STEP(BX[Q], BY[Q], 3, 4 * (Q) + 3)
#define BLOCK(BX, BY) { WORD(BX, BY, 0) WORD(BX, BY, 1) WORD(BX, BY, 2) WORD(BX, BY, 3) WORD(BX, BY, 4) WORD(BX, BY, 5) WORD(BX, BY, 6) WORD(BX, BY, 7) }
#define LDG256(W, P, CACHE)                                                                                   \
    asm volatile("ld.global.nc" CACHE ".v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                              \
                 : "=r"(W[0]), "=r"(W[1]), "=r"(W[2]), "=r"(W[3]), "=r"(W[4]), "=r"(W[5]), "=r"(W[6]), "=r"(W[7]) \
                 : "l"(P))
#define LDS128(W, A)                                                                                          \
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(W[0]), "=r"(W[1]), "=r"(W[2]), "=r"(W[3]) : "r"(A))

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{ .reg .pred p; WAIT_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra WAIT_%=; }" ::"r"(bar), "r"(parity) : "memory");
}

// every warp streams its own contiguous slice of `src` (bytes), 2 KB per block, nblk blocks
template <int MODE>
__global__ void __launch_bounds__(NW * 32, 1) k(const uint8_t *__restrict__ src, int wrap, int nblk, float *out, long long *clk)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t lut_off = 0x10000u - smem_base;
    float *lut2 = reinterpret_cast<float *>(smem + lut_off);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < 256 * 64; e += NW * 32) lut2[e] = (float)((e * 2654435761u) >> 20) * 1e-3f;
    // ring: 4 stages x 2 KB per warp above the table; mbarriers below it
    const uint32_t ring = smem_base + lut_off + 65536 + wid * 8192;
    const uint32_t bars = smem_base + 64 + wid * 32;  // 4 x 8 B
    if (MODE == 3 && lane == 0)
        for (int s = 0; s < 4; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * s));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const uint32_t colreg = 0x00010000u | (uint32_t)((32 - lane) * 4);
    float keep[32], sel[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) { keep[t] = lane == t ? 0.f : 1.f; sel[t] = lane == t ? 1.f : 0.f; }
    const uint8_t *p = src + ((long long)blockIdx.x * NW + wid) * (long long)wrap * 2048;  // the warp's slice: `wrap` blocks, re-read cyclically
    uint32_t bx[4][8], by[4][8];
    unsigned long long acc2 = 0ull, out2 = 0ull;
    float sum = 0.f;
    uint32_t seed = threadIdx.x * 2654435761u + 12345u;

#define FEED_ISSUE(S, B)                                                                                      \
    if (MODE == 1) { LDG256(bx[S], p + (long long)((B) & (wrap - 1)) * 2048 + lane * 32, ".L1::no_allocate"); LDG256(by[S], p + (long long)((B) & (wrap - 1)) * 2048 + 1024 + lane * 32, ".L1::no_allocate"); } \
    else if (MODE == 2) { LDG256(bx[S], p + (long long)((B) & (wrap - 1)) * 2048 + lane * 32, ""); LDG256(by[S], p + (long long)((B) & (wrap - 1)) * 2048 + 1024 + lane * 32, ""); } \
    else if (MODE == 3) {                                                                                     \
        if (lane == 0) {                                                                                      \
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                      \
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 2048;" ::"r"(bars + 8 * (S)) : "memory"); \
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 2048, [%2];" \
                         ::"r"(ring + (S) * 2048), "l"(p + (long long)((B) & (wrap - 1)) * 2048), "r"(bars + 8 * (S)) : "memory"); \
        }                                                                                                     \
    } else if (MODE == 4) {                                                                                   \
        _Pragma("unroll") for (int c = 0; c < 4; ++c)                                                         \
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + (S) * 2048 + c * 512 + lane * 16), "l"(p + (long long)((B) & (wrap - 1)) * 2048 + c * 512 + lane * 16)); \
        asm volatile("cp.async.commit_group;");                                                               \
    }
#define FEED_TAKE(S, B)                                                                                       \
    if (MODE == 0) {                                                                                          \
        _Pragma("unroll") for (int q = 0; q < 8; ++q) { seed = seed * 1664525u + 1013904223u; bx[S][q] = seed; by[S][q] = seed ^ 0x5bd1e995u; } \
    } else if (MODE == 3 || MODE == 4) {                                                                      \
        if (MODE == 3) mbar_wait(bars + 8 * (S), ((B) >> 2) & 1);                                             \
        else { asm volatile("cp.async.wait_group 3;" ::: "memory"); __syncwarp(); }                           \
        LDS128((&bx[S][0]), ring + (S) * 2048 + lane * 16);                                                   \
        LDS128((&bx[S][4]), ring + (S) * 2048 + 512 + lane * 16);                                             \
        LDS128((&by[S][0]), ring + (S) * 2048 + 1024 + lane * 16);                                            \
        LDS128((&by[S][4]), ring + (S) * 2048 + 1536 + lane * 16);                                            \
    }
#define STAGE(S)                                                                                              \
    if (m + (S) < nblk) {                                                                                     \
        FEED_TAKE(S, m + (S))                                                                                 \
        if (MODE == 3 || MODE == 4) __syncwarp(); /* every lane has read stage S: it may be refilled */       \
        if (m + (S) + 4 < nblk && (MODE == 3 || MODE == 4)) { FEED_ISSUE(S, m + (S) + 4) }                    \
        else if (MODE == 4) asm volatile("cp.async.commit_group;");                                           \
        if (m + (S) + 3 < nblk && (MODE == 1 || MODE == 2)) { FEED_ISSUE(((S) + 3) & 3, m + (S) + 3) }        \
        BLOCK(bx[S], by[S])                                                                                   \
        float a0, a1;                                                                                         \
        asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(out2));                                         \
        sum += a0; sum += a1; out2 = 0ull;                                                                    \
    }

    const long long t0 = clock64();
    if (MODE == 1 || MODE == 2) { FEED_ISSUE(0, 0) FEED_ISSUE(1, 1) FEED_ISSUE(2, 2) }
    if (MODE == 3 || MODE == 4) { FEED_ISSUE(0, 0) FEED_ISSUE(1, 1) FEED_ISSUE(2, 2) FEED_ISSUE(3, 3) }
#pragma unroll 1
    for (int m = 0; m < nblk; m += 4) { STAGE(0) STAGE(1) STAGE(2) STAGE(3) }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const uint8_t *src, int wrap, int nblk, const char *what)
{
    float *out;
    long long *clk;
    cudaMalloc(&out, 148 * NW * 32 * 4);
    cudaMalloc(&clk, 148 * 8);
    const int smem = 227 * 1024 - 64;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, NW * 32, smem>>>(src, wrap, nblk, out, clk);
    cudaEventRecord(e0);
    k<MODE><<<148, NW * 32, smem>>>(src, wrap, nblk, out, clk);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
    double cyc = 0;
    for (int i = 0; i < 148; ++i) cyc += (double)h[i];
    cyc /= 148;
    const double lookups_per_cta = (double)nblk * 64 * 32 * NW;
    printf("{\"feed\": %d, \"src\": \"%s\", \"err\": \"%s\", \"warps\": %d, \"nblk\": %d, \"ms\": %.4f, \"mean_cta_cycles\": %.0f, "
           "\"lookups_per_clk_per_sm\": %.2f, \"code_TBps\": %.3f}\n",
           MODE, what, cudaGetErrorString(err), NW, nblk, ms, cyc, lookups_per_cta / cyc, 148.0 * lookups_per_cta / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(clk);
}

int main()
{
    uint8_t *big;
    const long long BIG = 2ll << 30, SMALL = 48ll << 20;
    cudaMalloc(&big, BIG);
    cudaMemset(big, 0x5a, BIG);
    (void)SMALL;
    // L2-resident source: 32 KB per warp (37 MB in all), re-read 64 times
    run<0>(big, 16, 1024, "none");
    run<1>(big, 16, 1024, "L2 37MB"); run<2>(big, 16, 1024, "L2 37MB"); run<3>(big, 16, 1024, "L2 37MB"); run<4>(big, 16, 1024, "L2 37MB");
    // HBM source: 1 MB per warp (1.2 GB in all), read once
    run<1>(big, 512, 512, "HBM 1.2GB"); run<2>(big, 512, 512, "HBM 1.2GB"); run<3>(big, 512, 512, "HBM 1.2GB"); run<4>(big, 512, 512, "HBM 1.2GB");
    return 0;
}
