// Micro-benchmark of the skewed lookup step (tools only; not part of the library).
// Question: how many issue slots does one lookup cost in the variants below, measured on a B200, with the table
// and the code words already in shared memory (no HBM traffic)?
//   V0  one stream per lane: PRMT, LDS, ISETP, 2 predicated FADD            (the r01 engine)
//   V1  two streams per lane sharing the phase: 2 PRMT, 2 LDS, 1 ISETP, 2 predicated FADD2 (add.rn.f32x2)
//   V2  two streams per lane, scalar adds: 2 PRMT, 2 LDS, 1 ISETP, 4 predicated FADD
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_step tools/ubench_step.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define REGION_WORDS 40  // per stream: carry row + 4 rows
#define STEP1(W, BYTE, T)                                                                                     \
    {                                                                                                         \
        const uint32_t ad_ = __byte_perm(W, colreg, 0x7604 | ((BYTE) << 4));                                  \
        float v_;                                                                                             \
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v_) : "r"(ad_), "n"(4 * (T)));                      \
        asm("{.reg .pred p; setp.gt.s32 p, %2, %3; @p add.rn.f32 %0, %0, %4; @!p add.rn.f32 %1, %1, %4;}"   \
            : "+f"(accA), "+f"(accB)                                                                          \
            : "r"(lane), "n"(T), "f"(v_));                                                                    \
    }
#define WORD1(OFF, Q)                                                                                         \
    {                                                                                                         \
        const uint32_t x_ = *reinterpret_cast<const uint32_t *>(smem + (OFF) + 4 * (Q));                      \
        const uint32_t wd_ = __funnelshift_rc(xprev, x_, shift);                                              \
        xprev = x_;                                                                                           \
        STEP1(wd_, 0, 4 * (Q) + 0) STEP1(wd_, 1, 4 * (Q) + 1) STEP1(wd_, 2, 4 * (Q) + 2) STEP1(wd_, 3, 4 * (Q) + 3) \
    }
#define BLOCK1(OFF) { WORD1(OFF, 0) WORD1(OFF, 1) WORD1(OFF, 2) WORD1(OFF, 3) WORD1(OFF, 4) WORD1(OFF, 5) WORD1(OFF, 6) WORD1(OFF, 7) }

#define STEP2(WX, WY, BYTE, T)                                                                                \
    {                                                                                                         \
        const uint32_t ax_ = __byte_perm(WX, colreg, 0x7604 | ((BYTE) << 4));                                 \
        const uint32_t ay_ = __byte_perm(WY, colreg, 0x7604 | ((BYTE) << 4));                                 \
        float vx_, vy_;                                                                                       \
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(vx_) : "r"(ax_), "n"(4 * (T)));                     \
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(vy_) : "r"(ay_), "n"(4 * (T)));                     \
        asm("{.reg .pred p; .reg .b64 v; mov.b64 v, {%4, %5}; setp.gt.s32 p, %2, %3;"                        \
            " @p add.rn.f32x2 %0, %0, v; @!p add.rn.f32x2 %1, %1, v;}"                                       \
            : "+l"(accA2), "+l"(accB2)                                                                        \
            : "r"(lane), "n"(T), "f"(vx_), "f"(vy_));                                                         \
    }
#define STEP2S(WX, WY, BYTE, T)                                                                               \
    {                                                                                                         \
        const uint32_t ax_ = __byte_perm(WX, colreg, 0x7604 | ((BYTE) << 4));                                 \
        const uint32_t ay_ = __byte_perm(WY, colreg, 0x7604 | ((BYTE) << 4));                                 \
        float vx_, vy_;                                                                                       \
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(vx_) : "r"(ax_), "n"(4 * (T)));                     \
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(vy_) : "r"(ay_), "n"(4 * (T)));                     \
        asm("{.reg .pred p; setp.gt.s32 p, %4, %5; @p add.rn.f32 %0, %0, %6; @!p add.rn.f32 %1, %1, %6;"     \
            " @p add.rn.f32 %2, %2, %7; @!p add.rn.f32 %3, %3, %7;}"                                         \
            : "+f"(accA), "+f"(accB), "+f"(accC), "+f"(accD)                                                  \
            : "r"(lane), "n"(T), "f"(vx_), "f"(vy_));                                                         \
    }
#define WORD2(ST, OFFX, OFFY, Q)                                                                              \
    {                                                                                                         \
        const uint32_t x_ = *reinterpret_cast<const uint32_t *>(smem + (OFFX) + 4 * (Q));                     \
        const uint32_t y_ = *reinterpret_cast<const uint32_t *>(smem + (OFFY) + 4 * (Q));                     \
        const uint32_t wx_ = __funnelshift_rc(xprev, x_, shift);                                              \
        const uint32_t wy_ = __funnelshift_rc(yprev, y_, shift);                                              \
        xprev = x_;                                                                                           \
        yprev = y_;                                                                                           \
        ST(wx_, wy_, 0, 4 * (Q) + 0) ST(wx_, wy_, 1, 4 * (Q) + 1) ST(wx_, wy_, 2, 4 * (Q) + 2) ST(wx_, wy_, 3, 4 * (Q) + 3) \
    }
#define BLOCK2(ST, OX, OY) { WORD2(ST, OX, OY, 0) WORD2(ST, OX, OY, 1) WORD2(ST, OX, OY, 2) WORD2(ST, OX, OY, 3) \
                             WORD2(ST, OX, OY, 4) WORD2(ST, OX, OY, 5) WORD2(ST, OX, OY, 6) WORD2(ST, OX, OY, 7) }

// V3: two streams per lane, no predicates: acc = acc * keep_t + v (keep_t = lane == t ? 0 : 1 restarts the sum at the
// lane's row boundary), out = acc * sel_t + out (sel_t = 1 - keep_t captures the finished sum); both are ONE FFMA2 on
// the (stream x, stream y) register pair with the per-lane constant broadcast from a 32-bit register.
#define STEP3(WX, WY, BYTE, T)                                                                                \
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

template <int V, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k(int reps, float *out, long long *clk)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t lut_off = 0x10000u - smem_base;
    float *lut2 = reinterpret_cast<float *>(smem + lut_off);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < 256 * 64; e += NW * 32) lut2[e] = (float)((e * 2654435761u) >> 20) * 1e-3f;
    // regions: below the table for the first warps, above for the rest (2 streams per lane, 160 B each)
    const uint32_t warp_bytes = 2 * 32 * REGION_WORDS * 4;
    const int n_lo = (int)((lut_off - 1024) / warp_bytes);
    const int n_hi = (int)((227 * 1024 - 64 - (lut_off + 65536)) / warp_bytes);
    const int rid = wid % (n_lo + n_hi);  // (16 warps do not fit: the surplus warps share a region, read-only)
    const uint32_t region = rid < n_lo ? 1024 + rid * warp_bytes : lut_off + 65536 + (rid - n_lo) * warp_bytes;
    uint32_t *rw = reinterpret_cast<uint32_t *>(smem + region);
    for (int i = lane; i < 2 * 32 * REGION_WORDS; i += 32) rw[i] = (i + rid * 977) * 2246822519u;
    __syncthreads();
    const uint32_t sx = region + lane * REGION_WORDS * 4, sy = sx + 32 * REGION_WORDS * 4;
    const uint32_t rbx = sx + 4 * (8 - (lane >> 2)), rby = sy + 4 * (8 - (lane >> 2));
    const uint32_t shift = 8 * (4 - (lane & 3));
    const uint32_t colreg = 0x00010000u | (uint32_t)((32 - lane) * 4);
    float accA = 0.f, accB = 0.f, accC = 0.f, accD = 0.f, sum = 0.f;
    unsigned long long accA2 = 0ull, accB2 = 0ull;
    uint32_t xprev = 0, yprev = 0;
    unsigned long long acc2 = 0ull, out2 = 0ull;
    float keep[32], sel[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) { keep[t] = lane == t ? 0.f : 1.f; sel[t] = lane == t ? 1.f : 0.f; }
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
        if (V == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                BLOCK1(rbx + 32 * i)
                sum += accA; accA = accB; accB = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                BLOCK1(rby + 32 * i)
                sum += accA; accA = accB; accB = 0.f;
            }
        } else if (V == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                BLOCK2(STEP2, rbx + 32 * i, rby + 32 * i)
                float a0, a1;
                asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(accA2));
                sum += a0; sum += a1;
                accA2 = accB2; accB2 = 0ull;
            }
        } else if (V == 3) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                BLOCK2(STEP3, rbx + 32 * i, rby + 32 * i)
                float a0, a1;
                asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(out2));
                sum += a0; sum += a1;
                out2 = 0ull;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                BLOCK2(STEP2S, rbx + 32 * i, rby + 32 * i)
                sum += accA; sum += accC;
                accA = accB; accB = 0.f; accC = accD; accD = 0.f;
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int V, int NW>
static void run(int reps)
{
    float *out;
    long long *clk;
    cudaMalloc(&out, 148 * NW * 32 * 4);
    cudaMalloc(&clk, 148 * 8);
    const int smem = 227 * 1024 - 64;
    cudaFuncSetAttribute(k<V, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V, NW><<<148, NW * 32, smem>>>(reps, out, clk);
    cudaEventRecord(e0);
    k<V, NW><<<148, NW * 32, smem>>>(reps, out, clk);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
    const double lookups_per_cta = (double)reps * 8 * 32 * 32 * NW;  // 8 row-blocks x 32 steps x 32 lanes per warp
    printf("{\"variant\": %d, \"warps\": %d, \"err\": \"%s\", \"ms\": %.4f, \"cycles\": %lld, \"lookups_per_clk_per_sm\": %.3f, "
           "\"issue_slots_per_warp_step_at_4ipc\": %.3f, \"T_lookups_per_s\": %.3f}\n",
           V, NW, cudaGetErrorString(err), ms, h[0], lookups_per_cta / (double)h[0],
           4.0 * (double)h[0] / (lookups_per_cta / 32.0), 148.0 * lookups_per_cta / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(clk);
}

int main(int argc, char **argv)
{
    const int reps = argc > 1 ? atoi(argv[1]) : 2000;
    run<0, 16>(reps); run<0, 12>(reps); run<0, 8>(reps);
    run<1, 16>(reps); run<1, 12>(reps); run<1, 8>(reps);
    run<2, 16>(reps); run<2, 12>(reps); run<2, 8>(reps);
    run<3, 16>(reps); run<3, 14>(reps); run<3, 12>(reps); run<3, 8>(reps);
    return 0;
}
