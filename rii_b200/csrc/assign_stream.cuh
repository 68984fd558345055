// ===================================================================================================
// K6 on the streaming engine: nearest coarse center of every code row under the symmetric distance.
//
// Reference: PQKMeans::FindNearetCenterLinear (src/pqkmeans.cpp:193-218) with SymmetricDistance (:152-162), called for
// every row by RiiCpp::UpdatePostingLists (src/rii.h:335-359) and by the k-means assignment step (src/pqkmeans.cpp:88-94):
//     dist(row, k) = (((0 + Dm[0][c_k[0]][row[0]]) + Dm[1][c_k[1]][row[1]]) + ...)        sequential fp32 adds in m order
//     assign(row)  = argmin_k dist(row, k), strict '<' from FLT_MAX, k ascending            -> first minimum wins
//
// A center is a "query" whose distance table is T_k[m][j] = Dm[m][c_k[m]][j] (Dm is bit-symmetric, src/pqkmeans.cpp:23-34),
// so the sweep over one center is exactly the linear scan of scan_stream.cuh over the skew64 copy of the rows: same
// table arrangement lut2[ks][64] (bank = sub-space: conflict-free lookups), same private cp.async rings, same packed
// FFMA2 accumulation (bit-identical sequential sums).  What differs is the loop nest and the emission:
//   * CTA (p, g) owns row part p (a contiguous range of 64-row groups) and center group g (a contiguous range of centers,
//     visited in ascending order).  Per center it rebuilds the table from 32 H rows of Dm (L2-resident: M x Ks x Ks floats),
//     then streams its rows once.
//   * the running minimum per row lives in global memory, private to the CTA: best[g][row] / arg[g][row].  The 64 best
//     values a block's emission needs travel through the ring with the block (2 x 4 bytes per lane, cp.async.ca), so the
//     compare never waits on a load; stores happen only on improvement (~ln K times per row).
//   * k_assign_reduce folds the center groups in ascending order with the same strict '<'.
// Algorithmic bytes: n * (32 H + 4) per center swept; lookups n * K * M.
// ===================================================================================================
#pragma once
#include "scan_stream.cuh"

#define AS_STAGE_BYTES (ST_BLOCK_BYTES + 256)  // one block of code windows + the 64 best values of the group it completes

struct AssignArgs {
    const float *Dm;         // (M, Ks, Ks) codeword distance matrices
    const uint8_t *centers;  // (K, M)
    int K, M, Ks;            // M = real sub-spaces (<= 32 H; the rest of a padded row looks up zeros)
    const uint8_t *skew;     // skew64 table of the rows: ONE segment, rows of 32 H bytes
    long long n;             // rows
    int kc;                  // centers per center group (gridDim.y groups)
    float *best;             // (gridDim.y, n_pad)
    int *arg;                // (gridDim.y, n_pad)
    long long n_pad;         // 64 * groups
    int *bad;                // set when a table entry is too large for the packed accumulation (host falls back)
    uint32_t smem_bytes;
};

// issue block I of this warp's walk into ring stage S: the four 16-byte chunks this lane consumes, and -- for the first
// block of a unit after the first -- the best values of the two rows whose distances complete in it
#define AS_ISSUE(S, I)                                                                                        \
    {                                                                                                         \
        const int i_ = (I);                                                                                   \
        if (i_ < nblk) {                                                                                      \
            const int u_ = i_ / H, hb_ = i_ % H;                                                              \
            const uint8_t *p_ = a.skew + ((size_t)(f0 + u_) * H + hb_) * ST_BLOCK_BYTES + lane * 16;          \
            const uint32_t dst_ = ringb + (S) * AS_STAGE_BYTES;                                               \
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_ + lane * 16), "l"(p_));       \
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_ + lane * 16 + 512), "l"(p_ + 512)); \
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_ + lane * 16 + 1024), "l"(p_ + 1024)); \
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_ + lane * 16 + 1536), "l"(p_ + 1536)); \
            if (hb_ == 0 && u_ > 0 && !first) {                                                               \
                const float *b_ = best + (size_t)(f0 + u_ - 1) * 64 + lane;                                   \
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_ + ST_BLOCK_BYTES + lane * 4), "l"(b_)); \
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_ + ST_BLOCK_BYTES + 128 + lane * 4), "l"(b_ + 32)); \
            }                                                                                                 \
        }                                                                                                     \
        asm volatile("cp.async.commit_group;");                                                               \
    }

#define AS_STAGE(S)                                                                                           \
    if (m + (S) < nblk) {                                                                                     \
        constexpr int h_ = (S) % H;                                                                           \
        AS_ISSUE(((S) + ST_D) % ST_R, m + (S) + ST_D)                                                         \
        if constexpr (ST_D == 3) asm volatile("cp.async.wait_group 3;" ::: "memory");                         \
        else if constexpr (ST_D == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");                    \
        else asm volatile("cp.async.wait_group 1;" ::: "memory");                                             \
        uint32_t wx_[8], wy_[8];                                                                              \
        const uint32_t sb_ = ringb + (S) * AS_STAGE_BYTES;                                                    \
        ST_LDS128(wx_, sb_ + lane * 16);                                                                      \
        ST_LDS128(wx_ + 4, sb_ + lane * 16 + 512);                                                            \
        ST_LDS128(wy_, sb_ + lane * 16 + 1024);                                                               \
        ST_LDS128(wy_ + 4, sb_ + lane * 16 + 1536);                                                           \
        ST_BLOCK(wx_, wy_)                                                                                    \
        if constexpr (h_ == 0) {                                                                              \
            float dx_, dy_;                                                                                   \
            asm("mov.b64 {%0, %1}, %2;" : "=f"(dx_), "=f"(dy_) : "l"(out2));                                  \
            out2 = 0ull;                                                                                      \
            const int u_ = (m + (S)) / H;                                                                     \
            if (u_ > 0) {                                                                                     \
                float bx_ = 3.402823466e+38f, by_ = 3.402823466e+38f;                                         \
                if (!first) {                                                                                 \
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bx_) : "r"(sb_ + ST_BLOCK_BYTES + lane * 4)); \
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(by_) : "r"(sb_ + ST_BLOCK_BYTES + 128 + lane * 4)); \
                }                                                                                             \
                const long long rx_ = (long long)(f0 + u_ - 1) * 64 + lane;                                   \
                if (rx_ < a.n && (first || dx_ < bx_)) {                                                      \
                    best[rx_] = dx_ < bx_ ? dx_ : bx_;                                                        \
                    arg[rx_] = dx_ < bx_ ? k : -1;                                                            \
                }                                                                                             \
                if (rx_ + 32 < a.n && (first || dy_ < by_)) {                                                 \
                    best[rx_ + 32] = dy_ < by_ ? dy_ : by_;                                                   \
                    arg[rx_ + 32] = dy_ < by_ ? k : -1;                                                       \
                }                                                                                             \
            }                                                                                                 \
        }                                                                                                     \
    }

// grid (parts, center groups); NW warps; dynamic shared memory: [pad to TB][H tables of 64 KB][NW rings of ST_R stages]
template <int NW, int ST_R, int MINB, uint32_t TB, int H>
__global__ void __launch_bounds__(NW * 32, MINB) k_assign_stream(AssignArgs a)
{
    static_assert(H == 1 || H == 2, "rows of 32 or 64 bytes");
    static_assert(ST_R % H == 0, "a block's half-row index must be a constant of its pipeline stage");
    constexpr int ST_D = ST_R - 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t lut_off = TB - smem_base;
    float *lut2 = reinterpret_cast<float *>(smem_raw + lut_off);
    float *lutB = lut2 + (H - 1) * (SK_LUT_BYTES / 4);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t hi0 = lut_off + H * SK_LUT_BYTES;
    if (hi0 + (size_t)NW * ST_R * AS_STAGE_BYTES > a.smem_bytes) __trap();  // host sized the launch wrongly
    const uint32_t ringb = smem_base + hi0 + wid * (ST_R * AS_STAGE_BYTES);

    // this CTA's rows (64-row groups) and centers; this warp's slice of the groups
    const int G = (int)((a.n + 63) >> 6);
    const int gp = (G + gridDim.x - 1) / gridDim.x;
    const int pg0 = min(G, (int)blockIdx.x * gp), pg1 = min(G, pg0 + gp);
    const int per = (pg1 - pg0 + NW - 1) / NW;
    const int f0 = min(pg1, pg0 + wid * per), f_end = min(pg1, f0 + per);
    const int nblk = f_end > f0 ? (f_end - f0 + 1) * H : 0;  // + the unit that drains the lagging bytes of the last rows
    const int k0 = blockIdx.y * a.kc, k1 = min(a.K, k0 + a.kc);
    float *best = a.best + (size_t)blockIdx.y * a.n_pad;
    int *arg = a.arg + (size_t)blockIdx.y * a.n_pad;

    const uint32_t colreg = (uint32_t)((32 - lane) * 4);
    float keep[32], sel[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) {
        keep[t] = lane == t ? 0.f : 1.f;
        sel[t] = lane == t ? 1.f : 0.f;
    }
    const int rot = (int)((blockIdx.y * gridDim.x + blockIdx.x) * 53u) & 63;

#pragma unroll 1
    for (int k = k0; k < k1; ++k) {
        const bool first = k == k0;
        // the first blocks of the sweep go out before the table is rebuilt (the rings are idle, the table is not read)
        AS_ISSUE(0, 0)
        if constexpr (ST_D >= 2) AS_ISSUE(1 % ST_R, 1)
        if constexpr (ST_D == 3) AS_ISSUE(2 % ST_R, 2)
        __syncthreads();  // every warp finished the previous center's sweep: the table may be overwritten
        int bad = 0;
#pragma unroll
        for (int mh = 0; mh < H; ++mh) {
            const int mm = mh * 32 + lane;
            const int c0 = (mm + 32) & 63;  // column in table 0; column mm in table H - 1 (scan_stream.cuh)
            if (mm < a.M) {
                const float *row = a.Dm + ((size_t)mm * a.Ks + __ldg(a.centers + (size_t)k * a.M + mm)) * a.Ks;
                if (a.Ks == 256) {
#pragma unroll 6
                    for (int j = wid; j < 64; j += NW) {
                        const int ks = ((j + rot) & 63) * 4;
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(row + ks));
                        bad |= !(v.x <= ST_TABLE_LIMIT) | !(v.y <= ST_TABLE_LIMIT) | !(v.z <= ST_TABLE_LIMIT) | !(v.w <= ST_TABLE_LIMIT);
                        lut2[(ks + 0) * 64 + c0] = v.x; lutB[(ks + 0) * 64 + mm] = v.x;
                        lut2[(ks + 1) * 64 + c0] = v.y; lutB[(ks + 1) * 64 + mm] = v.y;
                        lut2[(ks + 2) * 64 + c0] = v.z; lutB[(ks + 2) * 64 + mm] = v.z;
                        lut2[(ks + 3) * 64 + c0] = v.w; lutB[(ks + 3) * 64 + mm] = v.w;
                    }
                } else {
                    for (int ks = wid; ks < 256; ks += NW) {
                        const float v = ks < a.Ks ? __ldg(row + ks) : 0.f;
                        bad |= !(v <= ST_TABLE_LIMIT);
                        lut2[ks * 64 + c0] = v;
                        lutB[ks * 64 + mm] = v;
                    }
                }
            } else if (first) {  // padded sub-spaces: zeros, written once
                for (int ks = wid; ks < 256; ks += NW) {
                    lut2[ks * 64 + c0] = 0.f;
                    lutB[ks * 64 + mm] = 0.f;
                }
            }
        }
        if (__syncthreads_or(bad)) {  // (also: the table is visible)
            if (threadIdx.x == 0) *a.bad = 1;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            return;
        }
        unsigned long long acc2 = 0ull, out2 = 0ull;
#pragma unroll 1
        for (int m = 0; m < nblk; m += ST_R) {
            AS_STAGE(0)
            AS_STAGE(1)
            if constexpr (ST_R >= 3) { AS_STAGE(2 % ST_R) }
            if constexpr (ST_R == 4) { AS_STAGE(3 % ST_R) }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
}

// fold the center groups (ascending, strict '<': the first minimum wins) -> assignment (and distance) per row
__global__ void k_assign_reduce(const float *__restrict__ best, const int *__restrict__ arg, int groups, long long n_pad, long long n,
                                int *__restrict__ assign, float *__restrict__ dist)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float b = 3.402823466e+38f;
    int a = -1;
    for (int g = 0; g < groups; ++g) {
        const float v = best[(size_t)g * n_pad + r];
        if (v < b) { b = v; a = arg[(size_t)g * n_pad + r]; }
    }
    assign[r] = a;
    if (dist) dist[r] = b;
}

static inline size_t assign_smem_bytes(int nw, int ring_stages, uint32_t tb, int H)
{
    return (size_t)tb - 1024 + (size_t)H * SK_LUT_BYTES + (size_t)nw * ring_stages * AS_STAGE_BYTES;
}
