// sm_100a kernels of the ADC hot path.  Reference semantics are cited per kernel (file:line into
// matsui528/rii v0.2.12); arithmetic is fp32 with explicit round-to-nearest sub/mul/add intrinsics so that
// nvcc never contracts to FMA and never reassociates: results are bit-identical to the reference code as
// written (oracle/_ref/strict_*, oracle/rii_oracle.cpp).
#pragma once
#include "topk.cuh"

#define RII_THREADS 256
#define RII_ROWS_PER_THREAD 4

// ---------------------------------------------------------------------------------------------------
// K1  distance table.  src/rii.h:361-373 (DTable) + src/distance.h:117-252 (fvec_L2sqr).
// `variant` = accumulator width of the reference build being mirrored: 16 (AVX-512), 8 (AVX), 4 (SSE).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sqdiff(float a, float b)
{
    float t = __fsub_rn(a, b);
    return __fmul_rn(t, t);
}

__device__ __noinline__ float l2sqr_lanes(const float *__restrict__ x, const float *__restrict__ y, int d, int variant)
{
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    if (d >= 8) {
        float a8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a8[i] = 0.f;
        if (variant == 16) {
            float a16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) a16[i] = 0.f;
            while (d >= 16) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a16[i] = __fadd_rn(a16[i], sqdiff(x[i], y[i]));
                x += 16; y += 16; d -= 16;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) a8[i] = __fadd_rn(a16[8 + i], a16[i]);
        }
        if (variant >= 8) {
            while (d >= 8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a8[i] = __fadd_rn(a8[i], sqdiff(x[i], y[i]));
                x += 8; y += 8; d -= 8;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) a4[i] = __fadd_rn(a8[4 + i], a8[i]);
        } else {
            while (d >= 8) {  // SSE build: 4-lane accumulator over every 4-block
#pragma unroll
                for (int i = 0; i < 4; ++i) a4[i] = __fadd_rn(a4[i], sqdiff(x[i], y[i]));
                x += 4; y += 4; d -= 4;
            }
        }
    }
    if (d >= 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) a4[i] = __fadd_rn(a4[i], sqdiff(x[i], y[i]));
        x += 4; y += 4; d -= 4;
    }
    // masked tail (src/distance.h:44-65): absent lanes contribute (0-0)^2 = +0, and a + 0 == a
    if (d > 0) a4[0] = __fadd_rn(a4[0], sqdiff(x[0], y[0]));
    if (d > 1) a4[1] = __fadd_rn(a4[1], sqdiff(x[1], y[1]));
    if (d > 2) a4[2] = __fadd_rn(a4[2], sqdiff(x[2], y[2]));
    return __fadd_rn(__fadd_rn(a4[0], a4[1]), __fadd_rn(a4[2], a4[3]));
}

// One table entry with the query sub-vector already in registers when Ds <= 4 (every BASELINE shape):
// (s0 + s1) + (s2 + s3) with absent lanes contributing +0 (src/distance.h:148-169).
__device__ __forceinline__ float l2sqr_small(const float (&q)[4], const float *__restrict__ c, int Ds)
{
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (Ds == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(c));
        s0 = sqdiff(q[0], v.x); s1 = sqdiff(q[1], v.y); s2 = sqdiff(q[2], v.z); s3 = sqdiff(q[3], v.w);
    } else {
        s0 = sqdiff(q[0], __ldg(c));
        if (Ds > 1) s1 = sqdiff(q[1], __ldg(c + 1));
        if (Ds > 2) s2 = sqdiff(q[2], __ldg(c + 2));
    }
    return __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3));
}

// grid (ceil(M*Ks/256), B).  Q: (B, M*Ds), cw: (M, Ks, Ds), T: (B, M*Ks)
__global__ void __launch_bounds__(RII_THREADS) k_dtable(const float *__restrict__ Q, const float *__restrict__ cw,
                                                        float *__restrict__ T, int M, int Ks, int Ds, int variant)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * Ks) return;
    int m = e / Ks;
    size_t b = blockIdx.y;
    T[b * (size_t)(M * Ks) + e] = l2sqr_lanes(Q + b * (size_t)(M * Ds) + (size_t)m * Ds, cw + (size_t)e * Ds, Ds, variant);
}

// ---------------------------------------------------------------------------------------------------
// ADC of one code row.  src/rii.h:375-394 (ADist): dist = 0; for m: dist += T[m][code[m]] -- sequential
// fp32 adds in m order.  0 + x == x exactly for x >= +0, so the chain starts at the first lookup.
// Row loads: a 32-byte row is one LDG.256 (sm_100a); rows that are multiples of 16/4 bytes use
// 128-/32-bit loads; anything else falls back to byte loads.
// ---------------------------------------------------------------------------------------------------
template <int M, bool STREAM>
__device__ __forceinline__ void load_row(const uint8_t *__restrict__ p, uint32_t (&w)[(M + 3) / 4])
{
    if constexpr (M % 32 == 0) {
#pragma unroll
        for (int i = 0; i < M / 32; ++i) {
            if constexpr (STREAM)
                asm("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(w[8 * i]), "=r"(w[8 * i + 1]), "=r"(w[8 * i + 2]), "=r"(w[8 * i + 3]),
                               "=r"(w[8 * i + 4]), "=r"(w[8 * i + 5]), "=r"(w[8 * i + 6]), "=r"(w[8 * i + 7])
                             : "l"(p + 32 * i));
            else
                asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(w[8 * i]), "=r"(w[8 * i + 1]), "=r"(w[8 * i + 2]), "=r"(w[8 * i + 3]),
                               "=r"(w[8 * i + 4]), "=r"(w[8 * i + 5]), "=r"(w[8 * i + 6]), "=r"(w[8 * i + 7])
                             : "l"(p + 32 * i));
        }
    } else if constexpr (M % 16 == 0) {
#pragma unroll
        for (int i = 0; i < M / 16; ++i) {
            uint4 v = __ldg(reinterpret_cast<const uint4 *>(p) + i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    } else if constexpr (M % 4 == 0) {
#pragma unroll
        for (int i = 0; i < M / 4; ++i) w[i] = __ldg(reinterpret_cast<const uint32_t *>(p) + i);
    } else {
#pragma unroll
        for (int i = 0; i < (M + 3) / 4; ++i) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (4 * i + j < M) v |= (uint32_t)__ldg(p + 4 * i + j) << (8 * j);
            w[i] = v;
        }
    }
}

template <int M>
__device__ __forceinline__ float adc_regs(const float *lut, int Ks, const uint32_t (&w)[(M + 3) / 4])
{
    float acc = lut[w[0] & 0xff];
#pragma unroll
    for (int m = 1; m < M; ++m) acc = __fadd_rn(acc, lut[m * Ks + ((w[m >> 2] >> (8 * (m & 3))) & 0xff)]);
    return acc;
}

// runtime-M fallback (byte loads)
__device__ __forceinline__ float adc_bytes(const float *lut, int Ks, int M, const uint8_t *__restrict__ row)
{
    float acc = lut[__ldg(row)];
    for (int m = 1; m < M; ++m) acc = __fadd_rn(acc, lut[m * Ks + __ldg(row + m)]);
    return acc;
}

template <int M_T, bool STREAM>
__device__ __forceinline__ float adc_row(const float *lut, int Ks, int M, const uint8_t *__restrict__ row)
{
    if constexpr (M_T > 0) {
        uint32_t w[(M_T + 3) / 4];
        load_row<M_T, STREAM>(row, w);
        return adc_regs<M_T>(lut, Ks, w);
    } else {
        return adc_bytes(lut, Ks, M, row);
    }
}

__device__ __forceinline__ void load_lut(float *lut, const float *__restrict__ T, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) lut[i] = __ldg(T + i);
}

// Dynamic shared memory layout shared by the scan kernels:
//   float lut[M*Ks] | u64 keys[cap] | int count | u64 thr | (kernel specific tail)
struct ScanSmem {
    float *lut;
    BlockTopk tk;
    unsigned char *tail;
};
__device__ __forceinline__ ScanSmem carve_smem(unsigned char *base, int lut_floats, int cap, int k)
{
    ScanSmem s;
    s.lut = reinterpret_cast<float *>(base);
    size_t off = ((size_t)lut_floats * 4 + 15) & ~(size_t)15;
    s.tk.keys = reinterpret_cast<u64 *>(base + off);
    off += (size_t)cap * 8;
    s.tk.thr = reinterpret_cast<u64 *>(base + off);
    off += 8;
    s.tk.count = reinterpret_cast<int *>(base + off);
    off += 8;
    s.tk.cap = cap;
    s.tk.k = k;
    s.tail = base + off;
    return s;
}
static inline size_t scan_smem_bytes(int lut_floats, int cap, size_t tail)
{
    return (((size_t)lut_floats * 4 + 15) & ~(size_t)15) + (size_t)cap * 8 + 16 + tail;
}

// Emit the CTA's sorted top-k: either final (ids/dists/count) or a partial key list for k_merge.
struct TopkOut {
    u64 *partial;        // (B, parts, k) keys, RII_KEY_MAX padded   (when !final)
    long long *out_ids;  // (B, k) global ids                         (when final)
    float *out_dists;    // (B, k)
    int *out_counts;     // (B)
    long long id_base;
    int final;
};
__device__ __forceinline__ void emit_topk(BlockTopk &tk, const TopkOut &o, int b, int part, int parts)
{
    tk.compact();
    int n = *tk.count;
    int k = tk.k;
    if (o.final) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            u64 key = tk.keys[i];
            o.out_ids[(size_t)b * k + i] = o.id_base + (long long)key_id(key);
            o.out_dists[(size_t)b * k + i] = key_dist(key);
        }
        if (threadIdx.x == 0) o.out_counts[b] = n;
    } else {
        u64 *dst = o.partial + ((size_t)b * parts + part) * k;
        for (int i = threadIdx.x; i < k; i += blockDim.x) dst[i] = i < n ? tk.keys[i] : RII_KEY_MAX;
    }
}

// ---------------------------------------------------------------------------------------------------
// K2/K3  linear scan + top-k.  src/rii.h:195-242 (QueryLinear): all rows (S == 0) or exactly the given
// target ids in the given order (S != 0), then top-k.  grid (parts, B).
// ---------------------------------------------------------------------------------------------------
struct LinearArgs {
    const float *T;            // (B, M*Ks) distance tables
    const uint8_t *codes;      // (N, M) local shard
    const long long *tids;     // (S) global ids or null
    long long S;
    long long N;               // local rows
    long long id_base;         // global id of local row 0
    int M, Ks, k, cap;
    TopkOut out;
};

template <int M_T>
__global__ void __launch_bounds__(RII_THREADS) k_scan_linear(LinearArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScanSmem s = carve_smem(smem_raw, a.M * a.Ks, a.cap, a.k);
    const int b = blockIdx.y;
    load_lut(s.lut, a.T + (size_t)b * a.M * a.Ks, a.M * a.Ks);
    s.tk.init();

    const long long ncand = a.S ? a.S : a.N;
    long long chunk = (ncand + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + RII_THREADS - 1) / RII_THREADS * RII_THREADS;
    long long pos = (long long)blockIdx.x * chunk;
    long long end = pos + chunk < ncand ? pos + chunk : ncand;
    bool first = true;
    while (pos < end) {
        const int R = first ? 1 : RII_ROWS_PER_THREAD;  // small first round: cheap first threshold
        first = false;
        s.tk.reserve(R * RII_THREADS);
        const uint32_t thr_hi = s.tk.thr_hi();
        const u64 thr_key = s.tk.thr_key();
        long long row[RII_ROWS_PER_THREAD];
#pragma unroll
        for (int r = 0; r < RII_ROWS_PER_THREAD; ++r) {
            row[r] = -1;
            long long idx = pos + (long long)r * RII_THREADS + threadIdx.x;
            if (r < R && idx < end) {
                long long rr = a.S ? (a.tids[idx] - a.id_base) : idx;
                if (rr >= 0 && rr < a.N) row[r] = rr;
            }
        }
        float d[RII_ROWS_PER_THREAD];
        if constexpr (M_T > 0) {
            uint32_t w[RII_ROWS_PER_THREAD][(M_T + 3) / 4];
#pragma unroll
            for (int r = 0; r < RII_ROWS_PER_THREAD; ++r)
                if (row[r] >= 0) {
                    if (a.S) load_row<M_T, false>(a.codes + row[r] * M_T, w[r]);
                    else load_row<M_T, true>(a.codes + row[r] * M_T, w[r]);
                }
#pragma unroll
            for (int r = 0; r < RII_ROWS_PER_THREAD; ++r)
                if (row[r] >= 0) d[r] = adc_regs<M_T>(s.lut, a.Ks, w[r]);
        } else {
#pragma unroll
            for (int r = 0; r < RII_ROWS_PER_THREAD; ++r)
                if (row[r] >= 0) d[r] = adc_bytes(s.lut, a.Ks, a.M, a.codes + row[r] * a.M);
        }
#pragma unroll
        for (int r = 0; r < RII_ROWS_PER_THREAD; ++r) {
            if (row[r] >= 0 && __float_as_uint(d[r]) <= thr_hi) {
                u64 key = pack_key(d[r], (uint32_t)row[r]);
                if (key < thr_key) s.tk.push(key);
            }
        }
        pos += (long long)R * RII_THREADS;
    }
    emit_topk(s.tk, a.out, b, blockIdx.x, gridDim.x);
}

// Merge `parts` partial key lists per query into the final top-k.  grid (B).
__global__ void __launch_bounds__(RII_THREADS) k_merge(const u64 *__restrict__ partial, int parts, int k, int cap,
                                                       TopkOut out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScanSmem s = carve_smem(smem_raw, 0, cap, k);
    s.tk.init();
    const int b = blockIdx.x;
    const u64 *src = partial + (size_t)b * parts * k;
    const long long n = (long long)parts * k;
    for (long long pos = 0; pos < n; pos += RII_THREADS) {
        s.tk.reserve(RII_THREADS);
        u64 thr_key = s.tk.thr_key();
        long long i = pos + threadIdx.x;
        if (i < n) {
            u64 key = src[i];
            if (key != RII_KEY_MAX && key < thr_key) s.tk.push(key);
        }
    }
    TopkOut o = out;
    o.final = 1;
    emit_topk(s.tk, o, b, 0, 1);
}

// Cross-shard merge: per query, the G per-shard top-k lists (global 64-bit ids, ascending
// (dist, id) each) gathered over NVLink are merged into the global top-k.  grid (B); P = pow2 >= G*k.
__global__ void __launch_bounds__(RII_THREADS) k_merge_shards(const long long *__restrict__ ids, const float *__restrict__ dists,
                                                              const int *__restrict__ counts, int G, int B, int k, int P,
                                                              long long *out_ids, float *out_dists, int *out_counts)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    long long *s_id = reinterpret_cast<long long *>(smem_raw);
    uint32_t *s_d = reinterpret_cast<uint32_t *>(smem_raw + (size_t)P * 8);
    const int b = blockIdx.x;
    int total = 0;
    for (int g = 0; g < G; ++g) total += counts[(size_t)g * B + b];
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        int g = i / k, j = i % k;
        bool valid = g < G && j < counts[(size_t)g * B + b];
        s_id[i] = valid ? ids[((size_t)g * B + b) * k + j] : 0x7fffffffffffffffll;
        s_d[i] = valid ? __float_as_uint(dists[((size_t)g * B + b) * k + j]) : 0xffffffffu;
    }
    __syncthreads();
    for (int kk = 2; kk <= P; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    uint32_t da = s_d[i], db = s_d[ixj];
                    long long ia = s_id[i], ib = s_id[ixj];
                    bool gt = da > db || (da == db && ia > ib);
                    bool up = (i & kk) == 0;
                    if (gt == up) { s_d[i] = db; s_d[ixj] = da; s_id[i] = ib; s_id[ixj] = ia; }
                }
            }
            __syncthreads();
        }
    int n = total < k ? total : k;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        out_ids[(size_t)b * k + i] = s_id[i];
        out_dists[(size_t)b * k + i] = __uint_as_float(s_d[i]);
    }
    if (threadIdx.x == 0) out_counts[b] = n;
}

// ---------------------------------------------------------------------------------------------------
// K4  coarse ranking + candidate plan.  src/rii.h:259-280: ADist to every coarse center, w = number of
// lists to consider, partial_sort of the first w.  We rank by (coarse dist, list id).  grid (B).
// The plan (prefix-sum form) turns the reference's sequential posting-list walk
// (src/rii.h:286-322) into per-list take counts.
// ---------------------------------------------------------------------------------------------------
struct PlanArgs {
    // inputs
    const int *glob_len;   // (nlist) global (all-shard) list lengths       [no subset]
    const int *pre_len;    // (nlist) sum of lengths on lower ranks, or null (single shard)
    const int *loc_len;    // (nlist) local list lengths
    const int *filt_cnt;   // (B, w_eff) filtered counts per ranked list, or null  [subset]: all shards together
    const int *filt_pre;   // (B, w_eff) [subset, sharded] members held by lower ranks, or null
    const int *filt_loc;   // (B, w_eff) [subset, sharded] members held locally, or null
    long long L;
    int topk;
    int w;                 // the reference's w (src/rii.h:267-277)
    int w_eff;             // ranked lists available (== w, or nlist on the full re-run)
    int nlist;
    // outputs
    int *ranked;           // (B, w_eff) list ids in rank order
    int *cum;              // (B, w_eff) inclusive prefix of local take counts
    int *take_last;        // (B) global take count from the last segment (subset truncation)
    int *J;                // (B) number of segments
    int *flags;            // (B) bit0: needs full ranking (walk beyond w), bit1: empty result
};

// f / pre / loc are indexed by RANK j (0..w_eff): (filtered or global) length of the j-th ranked list, the part
// of it held by lower ranks (null: 0) and the part held locally (null: subset mode, counts are local already).
__device__ void make_plan(const PlanArgs &p, int b, const int *f_by_rank, const int *pre_by_rank, const int *loc_by_rank,
                          int *cum_out = nullptr)
{
    // single thread over <= w_eff entries that the caller staged (shared memory in the fused kernels)
    int *cum = cum_out ? cum_out : p.cum + (size_t)b * p.w_eff;
    long long P = 0;
    int J = 0, flag = 0, local = 0;
    bool done = false;
    int take_last = 0;
    for (int j = 0; j < p.w_eff; ++j) {
        long long f = f_by_rank[j];
        long long take = f;
        if (P + f >= p.L) { take = p.L - P; done = true; }            // src/rii.h:302-304
        P += take;
        long long lt = take;
        if (loc_by_rank) {
            lt = take - (pre_by_rank ? pre_by_rank[j] : 0);
            if (lt < 0) lt = 0;
            if (lt > loc_by_rank[j]) lt = loc_by_rank[j];
        }
        local += (int)lt;
        cum[j] = local;
        take_last = (int)take;
        J = j + 1;
        if (done) break;
        if (j == p.w - 1 && P >= p.topk) { done = true; break; }       // src/rii.h:309
    }
    if (!done) {
        if (p.w_eff >= p.nlist) flag |= 2;   // src/rii.h:325: nothing (enough) found -> empty result
        else flag |= 1;                      // walk continues beyond w: host re-runs with the full ranking
    }
    p.J[b] = J;
    p.take_last[b] = take_last;
    p.flags[b] = flag;
}

struct CoarseArgs {
    const float *T;          // (B, M*Ks), or null: build the table in-kernel from Q / cw (K1 fused)
    const float *Q;          // (B, M*Ds)
    const float *cw;         // (M, Ks, Ds)
    int Ds, variant;
    const uint8_t *centers;  // (nlist, M)
    int M, Ks, nlist, cap;
    int do_plan;
    PlanArgs plan;
};

template <int M_T>
__global__ void __launch_bounds__(RII_THREADS) k_coarse_rank(CoarseArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScanSmem s = carve_smem(smem_raw, a.M * a.Ks, a.cap, a.plan.w_eff);
    const int b = blockIdx.x;
    if (a.T) {
        load_lut(s.lut, a.T + (size_t)b * a.M * a.Ks, a.M * a.Ks);
    } else {  // K1 fused: T[m][ks] straight into shared memory (codewords are read coalesced)
        const float *q = a.Q + (size_t)b * a.M * a.Ds;
        if (a.Ds <= 4) {  // every BASELINE shape: loads of 8 entries in flight per thread
#pragma unroll 8
            for (int e = threadIdx.x; e < a.M * a.Ks; e += RII_THREADS) {
                const float *qm = q + (size_t)(e / a.Ks) * a.Ds;
                float qv[4] = {0.f, 0.f, 0.f, 0.f};
                for (int i = 0; i < a.Ds; ++i) qv[i] = __ldg(qm + i);
                s.lut[e] = l2sqr_small(qv, a.cw + (size_t)e * a.Ds, a.Ds);
            }
        } else {
            for (int e = threadIdx.x; e < a.M * a.Ks; e += RII_THREADS)
                s.lut[e] = l2sqr_lanes(q + (size_t)(e / a.Ks) * a.Ds, a.cw + (size_t)e * a.Ds, a.Ds, a.variant);
        }
    }
    s.tk.init();
    for (int pos = 0; pos < a.nlist; pos += RII_THREADS) {
        s.tk.reserve(RII_THREADS);
        const uint32_t thr_hi = s.tk.thr_hi();
        const u64 thr_key = s.tk.thr_key();
        int no = pos + threadIdx.x;
        if (no < a.nlist) {
            float d = adc_row<M_T, false>(s.lut, a.Ks, a.M, a.centers + (size_t)no * a.M);
            if (__float_as_uint(d) <= thr_hi) {
                u64 key = pack_key(d, (uint32_t)no);
                if (key < thr_key) s.tk.push(key);
            }
        }
    }
    s.tk.compact();
    int *ranked = a.plan.ranked + (size_t)b * a.plan.w_eff;
    int *s_f = reinterpret_cast<int *>(s.tail), *s_pre = s_f + a.plan.w_eff, *s_loc = s_pre + a.plan.w_eff;
    for (int i = threadIdx.x; i < a.plan.w_eff; i += blockDim.x) {
        const int no = (int)key_id(s.tk.keys[i]);
        ranked[i] = no;
        if (a.do_plan) {  // stage the list lengths in rank order (parallel loads; the plan itself is a short serial scan)
            s_f[i] = a.plan.glob_len[no];
            s_pre[i] = a.plan.pre_len ? a.plan.pre_len[no] : 0;
            s_loc[i] = a.plan.loc_len[no];
        }
    }
    __syncthreads();
    if (a.do_plan && threadIdx.x == 0) make_plan(a.plan, b, s_f, s_pre, s_loc);
}

__global__ void k_plan(PlanArgs p, int B)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const size_t o = (size_t)b * p.w_eff;
    make_plan(p, b, p.filt_cnt + o, p.filt_pre ? p.filt_pre + o : nullptr, p.filt_loc ? p.filt_loc + o : nullptr);
    // sharded subset: the cut of the last segment counts members in global id order; lower ranks hold the first ones
    if (p.filt_pre && p.J[b] > 0) {
        const int t = p.take_last[b] - p.filt_pre[o + p.J[b] - 1];
        p.take_last[b] = t < 0 ? 0 : t;
    }
}

// ---------------------------------------------------------------------------------------------------
// K5  posting-list scan.  src/rii.h:283-320: walk the ranked lists, ADist every visited id, top-k.
// No subset: the candidate set is the concatenation of the first cum[j]-cum[j-1] ids of each planned
// segment; grid (parts, B), each CTA takes a contiguous slice of the flattened candidate space.
// ---------------------------------------------------------------------------------------------------
struct IvfArgs {
    const float *T;
    const uint8_t *codes;
    const long long *offsets;  // (nlist+1) CSR of local posting lists
    const int *ids;            // local ids, ascending per list
    const int *ranked;         // (B, w_eff)
    const int *cum;            // (B, w_eff)
    const int *J;              // (B)
    const int *flags;          // (B)
    const int *take_last;      // (B)
    const uint32_t *bitmap;    // subset membership (bit per local id) or null
    int w_eff;
    int M, Ks, k, cap;
    TopkOut out;
};

template <int M_T>
__global__ void __launch_bounds__(RII_THREADS) k_scan_ivf(IvfArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScanSmem s = carve_smem(smem_raw, a.M * a.Ks, a.cap, a.k);
    const int b = blockIdx.y;
    const int J = (a.flags[b] != 0) ? 0 : a.J[b];
    int *s_cum = reinterpret_cast<int *>(s.tail);
    int *s_list = s_cum + a.w_eff;
    load_lut(s.lut, a.T + (size_t)b * a.M * a.Ks, a.M * a.Ks);
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
        s_cum[j] = a.cum[(size_t)b * a.w_eff + j];
        s_list[j] = a.ranked[(size_t)b * a.w_eff + j];
    }
    s.tk.init();
    const int total = J ? s_cum[J - 1] : 0;
    int chunk = (total + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + RII_THREADS - 1) / RII_THREADS * RII_THREADS;
    int pos = blockIdx.x * chunk;
    int end = pos + chunk < total ? pos + chunk : total;
    bool first = true;
    while (pos < end) {
        const int R = first ? 1 : RII_ROWS_PER_THREAD;
        first = false;
        s.tk.reserve(R * RII_THREADS);
        const uint32_t thr_hi = s.tk.thr_hi();
        const u64 thr_key = s.tk.thr_key();
        int row[RII_ROWS_PER_THREAD];
#pragma unroll
        for (int r = 0; r < RII_ROWS_PER_THREAD; ++r) {
            row[r] = -1;
            if (r >= R) continue;
            int idx = pos + r * RII_THREADS + threadIdx.x;
            if (idx < end) {
                int lo = 0, hi = J - 1;  // first segment with cum > idx
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (s_cum[mid] > idx) hi = mid; else lo = mid + 1;
                }
                int p = idx - (lo ? s_cum[lo - 1] : 0);
                row[r] = __ldg(a.ids + a.offsets[s_list[lo]] + p);
            }
        }
        float d[RII_ROWS_PER_THREAD];
        if constexpr (M_T > 0) {
            uint32_t w[RII_ROWS_PER_THREAD][(M_T + 3) / 4];
#pragma unroll
            for (int r = 0; r < RII_ROWS_PER_THREAD; ++r)
                if (row[r] >= 0) load_row<M_T, false>(a.codes + (size_t)row[r] * M_T, w[r]);
#pragma unroll
            for (int r = 0; r < RII_ROWS_PER_THREAD; ++r)
                if (row[r] >= 0) d[r] = adc_regs<M_T>(s.lut, a.Ks, w[r]);
        } else {
#pragma unroll
            for (int r = 0; r < RII_ROWS_PER_THREAD; ++r)
                if (row[r] >= 0) d[r] = adc_bytes(s.lut, a.Ks, a.M, a.codes + (size_t)row[r] * a.M);
        }
#pragma unroll
        for (int r = 0; r < RII_ROWS_PER_THREAD; ++r) {
            if (row[r] >= 0 && __float_as_uint(d[r]) <= thr_hi) {
                u64 key = pack_key(d[r], (uint32_t)row[r]);
                if (key < thr_key) s.tk.push(key);
            }
        }
        pos += R * RII_THREADS;
    }
    emit_topk(s.tk, a.out, b, blockIdx.x, gridDim.x);
}

// Subset variant (src/rii.h:294 binary_search filter): membership comes from a bitmap over local ids.
// One CTA walks whole segments (segment j -> CTA j % parts) in stored order; the last segment is cut
// after its first take_last[b] *member* ids, which needs the in-order rank of every member.
template <int M_T>
__global__ void __launch_bounds__(RII_THREADS) k_scan_ivf_subset(IvfArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScanSmem s = carve_smem(smem_raw, a.M * a.Ks, a.cap, a.k);
    const int b = blockIdx.y;
    const int J = (a.flags[b] != 0) ? 0 : a.J[b];
    int *s_warp = reinterpret_cast<int *>(s.tail);  // 8 warp counts + running base
    load_lut(s.lut, a.T + (size_t)b * a.M * a.Ks, a.M * a.Ks);
    s.tk.init();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int j = blockIdx.x; j < J; j += gridDim.x) {
        const int no = a.ranked[(size_t)b * a.w_eff + j];
        const long long beg = a.offsets[no];
        const int len = (int)(a.offsets[no + 1] - beg);
        const int limit = (j == J - 1) ? a.take_last[b] : 0x7fffffff;
        int base = 0;  // members seen so far in this list (uniform across the CTA)
        for (int pos = 0; pos < len && base < limit; pos += RII_THREADS) {
            s.tk.reserve(RII_THREADS);
            const uint32_t thr_hi = s.tk.thr_hi();
            const u64 thr_key = s.tk.thr_key();
            int i = pos + threadIdx.x;
            int id = -1;
            bool member = false;
            if (i < len) {
                id = __ldg(a.ids + beg + i);
                member = (__ldg(a.bitmap + (id >> 5)) >> (id & 31)) & 1u;
            }
            unsigned bal = __ballot_sync(0xffffffffu, member);
            if (lane == 0) s_warp[wid] = __popc(bal);
            __syncthreads();
            int before = base + __popc(bal & ((1u << lane) - 1));
            int tot = 0;
            for (int w2 = 0; w2 < RII_THREADS / 32; ++w2) {
                int c = s_warp[w2];
                if (w2 < wid) before += c;
                tot += c;
            }
            if (member && before < limit) {
                float d = adc_row<M_T, false>(s.lut, a.Ks, a.M, a.codes + (size_t)id * a.M);
                if (__float_as_uint(d) <= thr_hi) {
                    u64 key = pack_key(d, (uint32_t)id);
                    if (key < thr_key) s.tk.push(key);
                }
            }
            base += tot;
            __syncthreads();  // s_warp reused next iteration
        }
    }
    emit_topk(s.tk, a.out, b, blockIdx.x, gridDim.x);
}

// membership bitmap over local ids from (sorted or unsorted) global target ids
__global__ void k_bitmap_set(const long long *__restrict__ tids, long long S, long long id_base, long long N,
                             uint32_t *bitmap)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    long long id = tids[i] - id_base;
    if (id >= 0 && id < N) atomicOr(bitmap + (id >> 5), 1u << (id & 31));
}

// filtered length of every ranked list.  grid (w_eff, B)
__global__ void __launch_bounds__(RII_THREADS) k_count_members(const long long *__restrict__ offsets,
                                                               const int *__restrict__ ids,
                                                               const int *__restrict__ ranked, int w_eff,
                                                               const uint32_t *__restrict__ bitmap, int *filt_cnt)
{
    const int b = blockIdx.y, j = blockIdx.x;
    const int no = ranked[(size_t)b * w_eff + j];
    const long long beg = offsets[no];
    const int len = (int)(offsets[no + 1] - beg);
    int c = 0;
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        int id = __ldg(ids + beg + i);
        c += (__ldg(bitmap + (id >> 5)) >> (id & 31)) & 1u;
    }
    __shared__ int red[RII_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < RII_THREADS / 32; ++i) t += red[i];
        filt_cnt[(size_t)b * w_eff + j] = t;
    }
}

// ---------------------------------------------------------------------------------------------------
// K6  symmetric-distance assignment.  src/pqkmeans.cpp:23-34,164-173 (codeword distance matrices),
// :152-162 (SymmetricDistance), :193-218 (FindNearetCenterLinear, first minimum wins), called from
// src/rii.h:350-354 and src/pqkmeans.cpp:88-94.
// ---------------------------------------------------------------------------------------------------
// Dm[m][k1][k2] = sum_i (c1[i]-c2[i])^2, scalar, sequential in i.  grid (ceil(Ks*Ks/256), M)
__global__ void __launch_bounds__(RII_THREADS) k_symmat(const float *__restrict__ cw, float *__restrict__ Dm, int Ks,
                                                        int Ds)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Ks * Ks) return;
    int m = blockIdx.y, k1 = e / Ks, k2 = e % Ks;
    const float *a = cw + ((size_t)m * Ks + k1) * Ds, *bb = cw + ((size_t)m * Ks + k2) * Ds;
    float dist = 0.f;
    for (int i = 0; i < Ds; ++i) dist = __fadd_rn(dist, sqdiff(a[i], bb[i]));
    Dm[(size_t)m * Ks * Ks + e] = dist;
}

// Each CTA owns a tile of TILE codes (staged once in shared memory) and sweeps all K centers in groups
// of G; the group's tables T_g[m][a] = Dm[m][center_g[m]][a] (Dm is bit-symmetric) are packed G-wide so
// one LDS.(32*G) serves G centers.  Running (min dist, argmin) per code lives in registers; centers are
// visited in ascending order with strict '<', i.e. the first minimum wins exactly as in the reference.
template <int G, int CPT>
__global__ void __launch_bounds__(RII_THREADS) k_assign(const float *__restrict__ Dm, const uint8_t *__restrict__ codes,
                                                        long long N, const uint8_t *__restrict__ centers, int K, int M,
                                                        int Ks, int *__restrict__ assign, float *__restrict__ best_dist)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *lut = reinterpret_cast<float *>(smem_raw);                     // [M][Ks][G]
    uint8_t *tile = smem_raw + (size_t)M * Ks * G * sizeof(float);        // [TILE][M]
    const int TILE = RII_THREADS * CPT;
    const long long row0 = (long long)blockIdx.x * TILE;
    const long long rows = (N - row0) < TILE ? (N - row0) : TILE;
    {   // stage the code tile (bytes are contiguous in global memory)
        const uint8_t *src = codes + row0 * M;
        const long long nbytes = rows * M;
        for (long long i = threadIdx.x; i < nbytes; i += blockDim.x) tile[i] = __ldg(src + i);
    }
    float best[CPT];
    int arg[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) { best[c] = 3.402823466e+38f; arg[c] = -1; }

    for (int k0 = 0; k0 < K; k0 += G) {
        __syncthreads();
        for (int e = threadIdx.x; e < M * Ks; e += blockDim.x) {
            int m = e / Ks, a = e % Ks;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                int k = k0 + g < K ? k0 + g : K - 1;
                lut[(size_t)e * G + g] = __ldg(Dm + ((size_t)m * Ks + centers[(size_t)k * M + m]) * Ks + a);
            }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            int r = c * RII_THREADS + threadIdx.x;
            if (r >= rows) continue;
            const uint8_t *code = tile + (size_t)r * M;
            float acc[G];
#pragma unroll
            for (int g = 0; g < G; ++g) acc[g] = 0.f;
            for (int m = 0; m < M; ++m) {
                const float *e = lut + ((size_t)m * Ks + code[m]) * G;
                if constexpr (G == 4) {
                    float4 v = *reinterpret_cast<const float4 *>(e);
                    acc[0] = __fadd_rn(acc[0], v.x); acc[1] = __fadd_rn(acc[1], v.y);
                    acc[2] = __fadd_rn(acc[2], v.z); acc[3] = __fadd_rn(acc[3], v.w);
                } else if constexpr (G == 2) {
                    float2 v = *reinterpret_cast<const float2 *>(e);
                    acc[0] = __fadd_rn(acc[0], v.x); acc[1] = __fadd_rn(acc[1], v.y);
                } else {
                    acc[0] = __fadd_rn(acc[0], e[0]);
                }
            }
#pragma unroll
            for (int g = 0; g < G; ++g)
                if (k0 + g < K && acc[g] < best[c]) { best[c] = acc[g]; arg[c] = k0 + g; }
        }
    }
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        int r = c * RII_THREADS + threadIdx.x;
        if (r < rows) {
            assign[row0 + r] = arg[c];
            if (best_dist) best_dist[row0 + r] = best[c];
        }
    }
}

// gather rows by id (sampling for PQk-means, src/rii.h:126-132)
__global__ void k_gather_rows(const uint8_t *__restrict__ codes, const long long *__restrict__ pick, long long n, int M,
                              uint8_t *__restrict__ out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * M) return;
    long long r = i / M;
    int m = (int)(i % M);
    out[i] = codes[pick[r] * M + m];
}

// list-ordered copy of the codes: out[p] = codes[ids[p]] (32-byte rows, 16 bytes per thread)
__global__ void k_gather_rows32_by_list(const uint8_t *__restrict__ codes, const int *__restrict__ ids, long long n,
                                        uint8_t *__restrict__ out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // 16-byte chunk index
    if (i >= n * 2) return;
    const long long p = i >> 1;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(codes + (size_t)ids[p] * 32) + (i & 1));
    reinterpret_cast<uint4 *>(out)[i] = v;
}

// histogram of code bytes per (cluster, subspace): hist[k][m][ks].  src/pqkmeans.cpp:229-233
__global__ void k_vote_hist(const uint8_t *__restrict__ codes, const int *__restrict__ assign, long long n, int M,
                            int Ks, int *hist)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * M) return;
    long long r = i / M;
    int m = (int)(i % M);
    atomicAdd(hist + ((size_t)assign[r] * M + m) * Ks + codes[i], 1);
}

// sparse voting, src/pqkmeans.cpp:235-258: vote[k2] = sum over k1 (ascending, freq != 0) of
// (float)freq * Dm[m][k1][k2] (mul then add), argmin with strict '<' from FLT_MAX.
// grid (M, K), Ks <= 256 threads.  Empty clusters keep their center (src/pqkmeans.cpp:115-120).
__global__ void __launch_bounds__(256) k_vote_centers(const float *__restrict__ Dm, const int *__restrict__ hist,
                                                      int M, int Ks, uint8_t *centers)
{
    const int m = blockIdx.x, k = blockIdx.y, k2 = threadIdx.x;
    __shared__ int s_hist[256];
    __shared__ float s_vote[256];
    __shared__ int s_total;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    int h = 0;
    if (k2 < Ks) h = hist[((size_t)k * M + m) * Ks + k2];
    s_hist[k2] = h;
    if (h) atomicAdd(&s_total, h);
    __syncthreads();
    if (s_total == 0) return;
    float vote = 0.f;
    if (k2 < Ks) {
        for (int k1 = 0; k1 < Ks; ++k1) {
            int freq = s_hist[k1];
            if (freq == 0) continue;
            vote = __fadd_rn(vote, __fmul_rn((float)freq, __ldg(Dm + ((size_t)m * Ks + k1) * Ks + k2)));
        }
    }
    s_vote[k2] = vote;
    __syncthreads();
    if (threadIdx.x == 0) {
        float min_dist = 3.402823466e+38f;
        int min_ks = -1;
        for (int ks = 0; ks < Ks; ++ks)
            if (s_vote[ks] < min_dist) { min_ks = ks; min_dist = s_vote[ks]; }
        centers[(size_t)k * M + m] = (uint8_t)min_ks;
    }
}

// ---------------------------------------------------------------------------------------------------
// PQ encoder (the step before the path, call site rii/rii.py:185 fine_quantizer.encode):
// codes[n][m] = argmin_ks sum_i (x[n][m*Ds+i] - C[m][ks][i])^2, fp32 sequential in i, first minimum wins.
// grid (ceil(n/256), M); the subspace's codebook (Ks*Ds floats) is broadcast-read from shared memory.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RII_THREADS) k_pq_encode(const float *__restrict__ X, long long n, const float *__restrict__ cw,
                                                           int M, int Ks, int Ds, uint8_t *__restrict__ codes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *cb = reinterpret_cast<float *>(smem_raw);
    const int m = blockIdx.y;
    for (int i = threadIdx.x; i < Ks * Ds; i += blockDim.x) cb[i] = __ldg(cw + (size_t)m * Ks * Ds + i);
    __syncthreads();
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const float *x = X + row * (size_t)(M * Ds) + (size_t)m * Ds;
    float best = 3.402823466e+38f;
    int arg = 0;
    if (Ds <= 8) {
        float xv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) xv[i] = i < Ds ? __ldg(x + i) : 0.f;
        for (int ks = 0; ks < Ks; ++ks) {
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < Ds) d = __fadd_rn(d, sqdiff(xv[i], cb[ks * Ds + i]));
            if (d < best) { best = d; arg = ks; }
        }
    } else {
        for (int ks = 0; ks < Ks; ++ks) {
            float d = 0.f;
            for (int i = 0; i < Ds; ++i) d = __fadd_rn(d, sqdiff(__ldg(x + i), cb[ks * Ds + i]));
            if (d < best) { best = d; arg = ks; }
        }
    }
    codes[row * M + m] = (uint8_t)arg;
}

// all ADC distances of a row range (diagnostics / K-parity tests): out[b][n]
template <int M_T>
__global__ void __launch_bounds__(RII_THREADS) k_adc_all(const float *__restrict__ T, const uint8_t *__restrict__ codes,
                                                         long long N, int M, int Ks, float *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *lut = reinterpret_cast<float *>(smem_raw);
    const int b = blockIdx.y;
    load_lut(lut, T + (size_t)b * M * Ks, M * Ks);
    __syncthreads();
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x)
        out[(size_t)b * N + n] = adc_row<M_T, true>(lut, Ks, M, codes + n * M);
}

// ===================================================================================================
// K2/K5 v2: bank-conflict-free scan for M = 32 ("skewed" schedule), linear and posting-list (IVF) flavours.
//
// Why: with the natural layout lut[m][ks] every lane of a warp looks up the same m at the same time, the
// bank is ks % 32 -- random -- and a warp-wide LDS costs ~2.6-3.5 crossbar cycles (ncu: profiles/r01_*):
// the scan is bound by the shared-memory crossbar at ~2.1-2.4 T lookups/s, a third of what HBM can feed.
// Here lane l runs l bytes behind lane 0 through its own stream of code rows, so at any instant the 32
// lanes work on 32 different sub-spaces m = (t - l) mod 32, and with the table stored transposed,
// lut2[ks][m], the bank is m: every LDS is conflict free.  Each lane still adds its candidate's 32 table
// entries in m = 0..31 order, so distances stay bit-identical to the reference's sequential sum
// (src/rii.h:386-394).
//
// Mechanics (per warp; no CTA barrier in the main loop):
//  * lut2 is [256][64] floats: column c holds sub-space c % 32, so the lane's column t + 32 - l never wraps
//    and the lookup address is ONE byte-permute: (ks << 8) | ((32 - l) * 4), plus the immediate 4 * t.
//  * code rows are staged global -> shared with cp.async into per-lane regions
//    [carry row | half A: 4 rows | half B: 4 rows] (stride 72 words = 8 mod 32, which makes the lanes' 4-byte
//    code-word reads conflict free as well).  A lane's region is its byte stream; reading it at word
//    (q - l/4) and funnel-shifting by l % 4 bytes yields the l-byte lag for free.  Linear: 16-byte chunks
//    dealt round-robin (512 contiguous bytes per warp instruction).  IVF: the scan reads a LIST-ORDERED copy of
//    the codes (row p of the copy = code of ids[p], rebuilt with the posting lists), so a planned segment is a
//    contiguous run of rows and is staged exactly like the linear scan; ids are only looked up for survivors.
//    (A per-lane id-indirected gather was measured first: it halves the issue rate -- 2 x 32 sector requests
//    per 32 rows congest the LSU queue; profiles/r01_ncu_k_scan_skew32_ivf_v2b.txt.)
//  * a lane finishes one candidate per 32 steps at its own phase: steps t < l still belong to the previous
//    row (accumulator A), steps t >= l to the new one (B); at the block end A is complete in every lane.
//  * top-k per warp (ballot-compacted pushes into a small shared buffer, warp-level bitonic compaction),
//    with a CTA-shared threshold tightened by atomicMin; the CTA merges its warps' lists at the end.
// ===================================================================================================
#define SK_J 4                                   // rows per lane per tile
#define SK_TILE_ROWS (32 * SK_J)                 // 128 rows = 4 KB of codes
#define SK_HALF_WORDS (SK_J * 8)                 // 32
#define SK_REGION_WORDS (8 + 2 * SK_HALF_WORDS)  // 72
#define SK_REGION_BYTES (SK_REGION_WORDS * 4)    // 288
#define SK_WARP_BYTES (32 * SK_REGION_BYTES)     // 9216
#define SK_LUT_BYTES 65536
#define SK_MAX_K 224

struct SkewArgs {
    const float *T;            // (B, 32*Ks), or null: build the table in-kernel from Q / cw (K1 fused)
    const float *Q;            // (B, 32*Ds)
    const float *cw;           // (32, Ks, Ds)
    const float *cw_t;         // (Ks, 32, Ds): the same codewords, sub-space fastest (coalesced in-kernel table build)
    int Ds, variant;
    const uint8_t *codes;      // linear: (N, 32) by id.  IVF: (N, 32) list-ordered copy (row p <-> ids[p])
    long long N;               // linear: rows of the shard
    const long long *offsets;  // IVF: CSR
    const int *ids;
    const long long *skew_off; // v4 (scan_stream.cuh): first physical row of every posting list in the skew64 table
    const int *ranked, *cum, *J, *flags;  // IVF plan
    int w_eff;
    int Ks, k, cap;            // cap = per-warp key capacity (power of two >= max(k, w_eff) + 32)
    uint32_t smem_bytes;       // dynamic shared memory of the launch (the kernel lays its regions out around the table)
    const uint8_t *centers;    // IVF fused: (nlist, 32) coarse centers, or null (plan comes from a separate k_coarse_rank)
    int nlist;
    int coarse_lists;          // v4 fused: rank the centers with the warps' top-k lists (nlist > 1024) instead of keeping every distance
    PlanArgs plan;             // IVF fused: plan inputs (lengths, L, topk, w) and its global outputs (ranked, J, flags)
    TopkOut out;
    long long *dbg;            // optional: per-CTA clock64() at [start, table ready, scan done, end] (tools/microbench.py)
};

struct WarpTopk {
    u64 *keys;   // shared, this warp's buffer (>= cap keys)
    int cap, k;  // cap = next_pow2(k + 32): compaction threshold of the current pass
    int count;   // warp-uniform
    u64 *thr_w;  // shared [nw]: every warp's ceil(k/nw)-th smallest key (RII_KEY_MAX until it has that many)
    int nw, wid;
};

// Bitonic sort of 32*R keys held in registers (element e = r*32 + lane), ascending.  Exchanges at distance >= 32
// are register-to-register inside a lane, smaller distances are warp shuffles: no shared memory, no barriers.
// (The shared-memory version took ~145 cycles per compare-exchange round: 16 K cycles for 128 keys, measured with
// the phase clocks -- profiles/r01_micro_ivf_phase_clocks_*.jsonl.)
template <int R>
__device__ __forceinline__ void warp_sort_regs(u64 (&v)[R], int lane)
{
#pragma unroll
    for (int kk = 2; kk <= 32 * R; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & jr) == 0) {
                        const bool up = ((r * 32) & kk) == 0;  // kk > 32 here: bit of the register index
                        const u64 x = v[r], y = v[r | jr];
                        const bool sw = (x > y) == up;
                        v[r] = sw ? y : x;
                        v[r | jr] = sw ? x : y;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const u64 x = v[r];
                    const u64 y = __shfl_xor_sync(0xffffffffu, x, j);
                    const bool up = (((r * 32 + lane) & kk) == 0);
                    const bool lower = (lane & j) == 0;
                    const bool take_min = lower == up;
                    v[r] = take_min ? (x < y ? x : y) : (x < y ? y : x);
                }
            }
        }
    }
}

// sort the first n (<= 32*R) keys of a shared-memory buffer in place (one warp), pad with RII_KEY_MAX
template <int R>
__device__ __forceinline__ void warp_sort_buf(u64 *keys, int n, int lane)
{
    u64 v[R];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = (r * 32 + lane) < n ? keys[r * 32 + lane] : RII_KEY_MAX;
    warp_sort_regs<R>(v, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) keys[r * 32 + lane] = v[r];
    __syncwarp();
}

__device__ __noinline__ void warp_sort_any(u64 *keys, int n, int lane)  // n <= 256; buffer holds >= next_pow2-ish 32*R slots
{
    if (n <= 64) warp_sort_buf<2>(keys, n, lane);
    else if (n <= 128) warp_sort_buf<4>(keys, n, lane);
    else warp_sort_buf<8>(keys, n, lane);
}

__device__ __noinline__ void warp_compact(WarpTopk &w, u64 *cta_thr, int lane)
{
    const int n = w.count;  // <= cap <= 256
    warp_sort_any(w.keys, n, lane);
    w.count = n < w.k ? n : w.k;
    // Two valid upper bounds of the CTA's k-th key tighten the shared threshold:
    //  (1) this warp's own k-th key;
    //  (2) the LARGEST, over all warps, of the warps' ceil(k/nw)-th keys: at least nw * ceil(k/nw) >= k keys lie below
    //      it.  With the candidates spread evenly over the warps (2) is ~nw times tighter than (1) for k >= nw.
    const int kq = (w.k + w.nw - 1) / w.nw;
    if (lane == 0) {
        if (w.count == w.k) atomicMin(cta_thr, w.keys[w.k - 1]);
        if (w.count >= kq) atomicMin(w.thr_w + w.wid, w.keys[kq - 1]);  // (atomics: other warps read this slot concurrently)
    }
    __syncwarp();
    u64 t = lane < w.nw ? atomicMin(w.thr_w + lane, RII_KEY_MAX) : 0ull;   // atomic read
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const u64 y = __shfl_xor_sync(0xffffffffu, t, o);
        t = t > y ? t : y;
    }
    if (lane == 0 && t != RII_KEY_MAX) atomicMin(cta_thr, t);
    __syncwarp();
}

// slow path of an emission: some lane beat the cached distance threshold.  Re-test against the exact,
// current (distance, id) threshold, append the survivors (ballot-compacted), compact when nearly full.
__device__ __noinline__ void warp_push(WarpTopk &w, u64 *cta_thr, int lane, float dist, uint32_t id, bool pre)
{
    const u64 thr = *reinterpret_cast<volatile u64 *>(cta_thr);
    const u64 key = pack_key(dist, id);
    const bool pass = pre && key < thr;
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if (!bal) return;
    if (pass) w.keys[w.count + __popc(bal & ((1u << lane) - 1u))] = key;
    w.count += __popc(bal);
    if (w.count + 32 > w.cap) warp_compact(w, cta_thr, lane);
}

// one lookup step.  The table sits at the 64 KB-aligned ABSOLUTE shared address 0x10000, so the byte-permute of
// (code word, colreg = 0x00010000 | lane column offset) IS the lookup address: 0x10000 | ks << 8 | col; the step
// index goes into the load's immediate.  (t < l ? A : B) += v with a predicated add pair.
#define SK_STEP(W, BYTE, T)                                                                                   \
    {                                                                                                         \
        const uint32_t ad_ = __byte_perm(W, colreg, 0x7604 | ((BYTE) << 4));                                  \
        float v_;                                                                                             \
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v_) : "r"(ad_), "n"(4 * (T)));                      \
        asm("{.reg .pred p; setp.gt.s32 p, %2, %3; @p add.rn.f32 %0, %0, %4; @!p add.rn.f32 %1, %1, %4;}"   \
            : "+f"(accA), "+f"(accB)                                                                          \
            : "r"(lane), "n"(T), "f"(v_));                                                                    \
    }

// 4 steps = one code word of the lane's (lagged) stream
#define SK_WORD(WOFF, Q)                                                                                      \
    {                                                                                                         \
        const uint32_t x_ = *reinterpret_cast<const uint32_t *>(smem_raw + (WOFF) + 4 * (Q));                 \
        const uint32_t wd_ = __funnelshift_rc(xprev, x_, shift);                                              \
        xprev = x_;                                                                                           \
        SK_STEP(wd_, 0, 4 * (Q) + 0)                                                                          \
        SK_STEP(wd_, 1, 4 * (Q) + 1)                                                                          \
        SK_STEP(wd_, 2, 4 * (Q) + 2)                                                                          \
        SK_STEP(wd_, 3, 4 * (Q) + 3)                                                                          \
    }

// drain variant: the 32 steps after a lane's last row.  Only the lagging steps (t < l) carry real data (the tail of the
// last row); the rest would read past the lane's region -- into the next lane's / warp's carry row, harmless for the
// results but a data race (racecheck) -- so the word address is clamped to the lane's own last word.
#define SK_WORD_CLAMP(WOFF, Q, LIM)                                                                           \
    {                                                                                                         \
        const uint32_t wa_ = (WOFF) + 4 * (Q) < (LIM) ? (WOFF) + 4 * (Q) : (LIM);                             \
        const uint32_t x_ = *reinterpret_cast<const uint32_t *>(smem_raw + wa_);                              \
        const uint32_t wd_ = __funnelshift_rc(xprev, x_, shift);                                              \
        xprev = x_;                                                                                           \
        SK_STEP(wd_, 0, 4 * (Q) + 0)                                                                          \
        SK_STEP(wd_, 1, 4 * (Q) + 1)                                                                          \
        SK_STEP(wd_, 2, 4 * (Q) + 2)                                                                          \
        SK_STEP(wd_, 3, 4 * (Q) + 3)                                                                          \
    }
#define SK_BLOCK_CLAMP(WOFF, LIM)                                                                             \
    {                                                                                                         \
        SK_WORD_CLAMP(WOFF, 0, LIM) SK_WORD_CLAMP(WOFF, 1, LIM) SK_WORD_CLAMP(WOFF, 2, LIM) SK_WORD_CLAMP(WOFF, 3, LIM) \
        SK_WORD_CLAMP(WOFF, 4, LIM) SK_WORD_CLAMP(WOFF, 5, LIM) SK_WORD_CLAMP(WOFF, 6, LIM) SK_WORD_CLAMP(WOFF, 7, LIM) \
    }

// one block = 32 steps = 8 code words starting at byte offset WOFF of the dynamic shared memory
#define SK_BLOCK(WOFF)                                                                                        \
    {                                                                                                         \
        SK_WORD(WOFF, 0) SK_WORD(WOFF, 1) SK_WORD(WOFF, 2) SK_WORD(WOFF, 3)                                   \
        SK_WORD(WOFF, 4) SK_WORD(WOFF, 5) SK_WORD(WOFF, 6) SK_WORD(WOFF, 7)                                   \
    }

// end of a block: accumulator A holds the finished distance of local candidate `eloc` (id ID) in every lane
#define SK_EMIT(ID)                                                                                           \
    if (IVF && direct) { /* coarse pass of the fused kernel: keep every distance */                           \
        if (eloc < (uint32_t)cnt) pool_d[(ID)] = __float_as_uint(accA);                                       \
        accA = accB;                                                                                          \
        accB = 0.f;                                                                                           \
    } else {                                                                                                  \
        if constexpr (IVF) thr_hi = reinterpret_cast<volatile uint32_t *>(cta_thr)[1];                        \
        const bool pre_ = eloc < (uint32_t)cnt && __float_as_uint(accA) <= thr_hi;                            \
        if (__any_sync(0xffffffffu, pre_)) {                                                                  \
            const uint32_t id_ = segm ? (pre_ ? cand_id(ID) : 0u) : (ID);                                     \
            warp_push(wt, cta_thr, lane, accA, id_, pre_);                                                    \
            thr_hi = reinterpret_cast<volatile uint32_t *>(cta_thr)[1];                                       \
        }                                                                                                     \
        accA = accB;                                                                                          \
        accB = 0.f;                                                                                           \
    }

// bitonic sort of P (power of two) keys in shared memory by ONE warp (warp barriers only)
__device__ __forceinline__ void warp_sort_smem(u64 *k, int P, int lane)
{
    __syncwarp();
    for (int kk = 2; kk <= P; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < P; i += 32) {
                int ixj = i ^ j;
                if (ixj > i) {
                    u64 x = k[i], y = k[ixj];
                    bool up = (i & kk) == 0;
                    if ((x > y) == up) { k[i] = y; k[ixj] = x; }
                }
            }
            __syncwarp();
        }
}

// Coarse selection (fused kernel): the w smallest of np (distance bits, index) pairs, distances in shared memory.
// One CTA-wide histogram pass: 256 equal-width buckets over [min, max] of the (non-negative float) distance bits, a
// redundant per-warp scan finds the bucket b* holding the w-th smallest; everything in buckets <= b* (w keys plus the
// few extra of bucket b*) is gathered as (dist, index) keys by one warp and sorted in registers.
// (Measured alternatives, phase clocks: per-warp top-w lists + pool sort 27 K cycles; single-warp bisection 60 K; four
// 8-bit radix passes 13 K.)
// Returns (in every thread) the number of keys in `out` (>= w), or -1 if more than 256 qualify (heavily tied
// distances: the caller falls back to a full sort).
// mm_ready: the caller already zeroed hist[0..255], accumulated min / max of d[] into hist[256] / hist[257] and
// passed a CTA barrier (the v4 engine does that while it emits the distances).
template <int NT>
__device__ __forceinline__ int cta_select_smallest(const uint32_t *d, int np, int w, u64 *out, int *hist /* 256 + 4 ints */,
                                                   bool mm_ready = false)
{
    const int lane = threadIdx.x & 31;
    uint32_t *mm = reinterpret_cast<uint32_t *>(hist + 256);  // [0] min, [1] max, [2] result count
    if (!mm_ready) {
    for (int i = threadIdx.x; i < 256; i += NT) hist[i] = 0;
    if (threadIdx.x == 0) { mm[0] = 0xffffffffu; mm[1] = 0u; }
    __syncthreads();
    {
        uint32_t lo = 0xffffffffu, hi = 0u;
        for (int i = threadIdx.x; i < np; i += NT) {
            const uint32_t v = d[i];
            lo = v < lo ? v : lo;
            hi = v > hi ? v : hi;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t a = __shfl_xor_sync(0xffffffffu, lo, o), c = __shfl_xor_sync(0xffffffffu, hi, o);
            lo = a < lo ? a : lo;
            hi = c > hi ? c : hi;
        }
        if (lane == 0) { atomicMin(&mm[0], lo); atomicMax(&mm[1], hi); }
    }
    __syncthreads();
    }
    const uint32_t mn = mm[0], range = mm[1] - mn;
    const int sh = range >= 256u ? (32 - __clz(range)) - 8 : 0;  // (v - mn) >> sh is in [0, 255]
    for (int i = threadIdx.x; i < np; i += NT) atomicAdd(&hist[(d[i] - mn) >> sh], 1);
    __syncthreads();
    // bins 8*lane .. 8*lane+7 -> inclusive prefix over lanes -> the bin holding the w-th smallest value
    int c[8], tot = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = hist[8 * lane + j]; tot += c[j]; }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    const int excl = incl - tot;
    const bool mine = excl < w && w <= incl;  // exactly one lane (w <= np)
    int bin = 0;
    if (mine) {
        int run = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (run + c[j] >= w) { bin = 8 * lane + j; break; }
            run += c[j];
        }
    }
    bin = __shfl_sync(0xffffffffu, bin, __ffs(__ballot_sync(0xffffffffu, mine)) - 1);
    if (threadIdx.x < 32) {
        int n = 0;
        for (int i0 = 0; i0 < np; i0 += 32) {
            const int i = i0 + lane;
            const bool ok = i < np && (int)((d[i] - mn) >> sh) <= bin;
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (n + __popc(bal) > 256) { n = -1; break; }
            if (ok) out[n + __popc(bal & ((1u << lane) - 1u))] = ((u64)d[i] << 32) | (u64)(uint32_t)i;
            n += __popc(bal);
        }
        if (n > 0) warp_sort_any(out, n, lane);
        if (lane == 0) mm[2] = (uint32_t)n;
    }
    __syncthreads();
    return (int)mm[2];
}

// Phases: IVF launches with a.centers != null run TWO passes of the same engine in one CTA -- pass 0 ranks the
// coarse centers (a plain linear scan over the (nlist, 32) center table with k = w_eff; K4, src/rii.h:259-280),
// the plan is made in shared memory (make_plan), pass 1 scans the planned posting-list segments (K5).  One table
// build, one launch, no round trip of ranked lists / plans through HBM.
template <int NW, bool IVF>
__global__ void __launch_bounds__(NW * 32, 1) k_scan_skew32(SkewArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout (dynamic shared memory, window starts at absolute shared address sbase ~ 1 KB):
    //   [NW key buffers][cta_thr][IVF: s_off i64[w] | s_cum, s_f, s_pre, s_loc i32[w] | s_plan i32[4]][n_lo regions] ...
    //   lut2 (64 KB) at ABSOLUTE shared address 0x10000 ... [n_hi regions]
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t lut_off = 0x10000u - smem_base;
    float *lut2 = reinterpret_cast<float *>(smem_raw + lut_off);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int capw = a.cap;
    long long *dbg = a.dbg ? a.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();
    const uint32_t keys_off = 0;
    u64 *wkeys = reinterpret_cast<u64 *>(smem_raw + keys_off) + (size_t)wid * capw;
    u64 *cta_thr = reinterpret_cast<u64 *>(smem_raw + keys_off) + (size_t)NW * capw;
    u64 *thr_w = cta_thr + 1;  // [NW]
    long long *s_off = reinterpret_cast<long long *>(smem_raw + keys_off + (size_t)NW * capw * 8 + 8 + NW * 8);
    const int wq = IVF ? a.w_eff : 0;
    int *s_cum = reinterpret_cast<int *>(s_off + wq);
    int *s_f = s_cum + wq, *s_pre = s_f + wq, *s_loc = s_pre + wq, *s_plan = s_loc + wq;  // s_plan: [J, flags]
    const int b = blockIdx.y;
    const bool fused = IVF && a.centers != nullptr;
    int J = 0;
    if constexpr (IVF) {
        if (!fused) {
            J = (a.flags[b] != 0) ? 0 : a.J[b];
            for (int j = threadIdx.x; j < J; j += blockDim.x) {
                s_cum[j] = a.cum[(size_t)b * a.w_eff + j];
                s_off[j] = a.offsets[a.ranked[(size_t)b * a.w_eff + j]];
            }
        }
    }
    if (threadIdx.x == 0) *cta_thr = RII_KEY_MAX;
    if (threadIdx.x < NW) thr_w[threadIdx.x] = RII_KEY_MAX;
    __syncthreads();  // plan (unfused IVF) and threshold are visible; the table is built after the first tile is in flight

    const uint32_t lo_reg0 = (uint32_t)(((size_t)NW * capw * 8 + 16 + NW * 8 + (IVF ? (size_t)a.w_eff * 24 + 32 : 0) + 15) & ~(size_t)15);
    const uint32_t hi_reg0 = lut_off + SK_LUT_BYTES;
    const int n_lo = (int)((lut_off - lo_reg0) / SK_WARP_BYTES);
    if (n_lo + (int)((a.smem_bytes - hi_reg0) / SK_WARP_BYTES) < NW) __trap();  // host sized the launch wrongly
    const uint32_t region = wid < n_lo ? lo_reg0 + wid * SK_WARP_BYTES : hi_reg0 + (wid - n_lo) * SK_WARP_BYTES;
    const uint32_t myreg = region + lane * SK_REGION_BYTES;
    const uint32_t rb = myreg + 4 * (8 - (lane >> 2));  // lane stream base with the word part of the lag folded in
    const uint32_t shift = 8 * (4 - (lane & 3));         // funnel shift (32 == no byte lag)
    const uint32_t colreg = 0x00010000u | (uint32_t)((32 - lane) * 4);  // table address 0x10000 | column byte offset
    // linear: destination of 16-byte chunk (it, lane): rows are dealt SK_J per lane
    const uint32_t cp_dst = region + (lane >> 3) * SK_REGION_BYTES + 32 + (lane & 7) * 16;

    // per-pass state
    const uint8_t *pc = a.codes;   // row table of the pass
    long long total = 0, base = 0, end = 0;
    int cnt = 0, ntiles = 0;
    int segw = 0;
    WarpTopk wt;
    wt.keys = wkeys;
    wt.cap = capw;
    wt.k = a.k;
    wt.count = 0;
    wt.thr_w = thr_w;
    wt.nw = NW;
    wt.wid = wid;

    auto set_range = [&](long long tot, int nsplit, int split) {  // this warp's slice [base, base + cnt) of [0, tot)
        total = tot;
        const long long per_cta = ((tot + nsplit - 1) / nsplit + NW * SK_TILE_ROWS - 1) / (NW * SK_TILE_ROWS) * (NW * SK_TILE_ROWS);
        base = (long long)split * per_cta + (long long)wid * (per_cta / NW);
        end = base + per_cta / NW;
        if (end > tot) end = tot;
        cnt = end > base ? (int)(end - base) : 0;
        ntiles = (cnt + SK_TILE_ROWS - 1) / SK_TILE_ROWS;
    };
    auto issue_tile = [&](int n) {  // rows [base + 128 n, +128) of the contiguous table pc
        const long long r0 = base + (long long)n * SK_TILE_ROWS;
        const uint8_t *g = pc + r0 * 32 + lane * 16;
        const uint32_t dst = smem_base + cp_dst + ((n & 1) ? SK_J * 32 : 0);
        if (r0 + SK_TILE_ROWS <= end) {  // full tile: 8 x 512 contiguous bytes per warp, immediates only
#pragma unroll
            for (int it = 0; it < SK_TILE_ROWS / 16; ++it)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + it * 4 * SK_REGION_BYTES), "l"(g + it * 512));
        } else {                          // last tile: rows past the end are zero filled, their results masked
#pragma unroll
            for (int it = 0; it < SK_TILE_ROWS / 16; ++it) {
                const int nbytes = r0 + it * 16 + (lane >> 1) < end ? 16 : 0;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + it * 4 * SK_REGION_BYTES),
                             "l"(nbytes ? g + it * 512 : pc), "r"(nbytes));
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    // IVF pass 1: tile n = flattened candidates [c0, c0 + 128) of the plan; segment j covers [cum[j-1], cum[j]) and
    // starts at row s_off[j] of the list-ordered code copy.  segw = segment of c0 (warp-uniform, carried along).
    auto issue_tile_seg = [&](int n) {
        const int c0 = (int)base + n * SK_TILE_ROWS;
        const int cend = (int)end;
        while (segw < J - 1 && s_cum[segw] <= c0) ++segw;
        const int seg_lo = segw ? s_cum[segw - 1] : 0;
        const uint32_t dst = smem_base + cp_dst + ((n & 1) ? SK_J * 32 : 0);
        if (c0 + SK_TILE_ROWS <= cend && c0 + SK_TILE_ROWS <= s_cum[segw]) {  // one segment, full tile: pure stream
            const uint8_t *g = pc + (size_t)(s_off[segw] + (c0 - seg_lo)) * 32 + lane * 16;
#pragma unroll
            for (int it = 0; it < SK_TILE_ROWS / 16; ++it)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + it * 4 * SK_REGION_BYTES), "l"(g + it * 512));
        } else {                                                             // crosses a segment boundary / tail
#pragma unroll
            for (int it = 0; it < SK_TILE_ROWS / 16; ++it) {
                const int c = c0 + it * 16 + (lane >> 1);
                const bool ok = c < cend;
                int seg = segw;
                if (ok) while (s_cum[seg] <= c) ++seg;
                const uint8_t *g = pc + (size_t)(s_off[seg] + (c - (seg ? s_cum[seg - 1] : 0))) * 32 + (lane & 1) * 16;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + it * 4 * SK_REGION_BYTES),
                             "l"(ok ? g : pc), "r"(ok ? 16 : 0));
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    // IVF pass 1: posting-list id of flattened candidate c (survivors only)
    auto cand_id = [&](uint32_t c) -> uint32_t {
        if (c >= (uint32_t)total) return 0u;
        int lo = 0, hi = J - 1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (s_cum[mid] > (int)c) hi = mid; else lo = mid + 1;
        }
        return (uint32_t)__ldg(a.ids + s_off[lo] + ((int)c - (lo ? s_cum[lo - 1] : 0)));
    };

    // ---- pass setup: the first tile goes out before the table is built ---------------------------------
    bool segm = IVF && !fused;  // true: pass over planned posting-list segments; false: plain row range
    bool direct = fused;         // coarse pass: every distance goes to pool_d[center]
    // the warps' key buffers are idle during the coarse pass: they hold the nlist distances (host checks the size)
    uint32_t *pool_d = reinterpret_cast<uint32_t *>(smem_raw + keys_off);
    wt.cap = next_pow2(wt.k + 32) < 64 ? 64 : next_pow2(wt.k + 32);
    if (fused) {
        pc = a.centers;
        set_range(a.nlist, 1, 0);
    } else if (IVF) {
        set_range(J ? (long long)s_cum[J - 1] : 0, gridDim.x, blockIdx.x);
    } else {
        set_range(a.N, gridDim.x, blockIdx.x);
    }
    // zero the carry row of this lane (read by the lagging steps of the very first block)
    *reinterpret_cast<uint4 *>(smem_raw + myreg) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4 *>(smem_raw + myreg + 16) = make_uint4(0, 0, 0, 0);
    if (ntiles > 0) {
        if (segm) issue_tile_seg(0);
        else issue_tile(0);
    }
    {   // lut2[ks][c] = T[c % 32][ks]; rows >= Ks are zero (zero-filled padding rows index row 0 only)
        if (a.T) {
            const float *T = a.T + (size_t)b * 32 * a.Ks;
#pragma unroll 8
            for (int e = threadIdx.x; e < 256 * 64; e += NW * 32) {
                int ks = e >> 6, c = e & 63;
                lut2[e] = ks < a.Ks ? __ldg(T + (c & 31) * a.Ks + ks) : 0.f;
            }
        } else {
            // K1 fused (src/rii.h:361-373): entry (m = lane, ks) -> both columns m and m + 32 of row ks; the
            // lane's query sub-vector stays in registers, codewords come from the sub-space-fastest copy (one
            // contiguous 32*Ds-float row per ks), stores are bank-conflict free.
            const float *qm = a.Q + (size_t)b * 32 * a.Ds + (size_t)lane * a.Ds;
            if (a.Ds <= 4) {
                float qv[4] = {0.f, 0.f, 0.f, 0.f};
                for (int i = 0; i < a.Ds; ++i) qv[i] = __ldg(qm + i);
#pragma unroll 8
                for (int ks = wid; ks < 256; ks += NW) {
                    float v = 0.f;
                    if (ks < a.Ks) v = l2sqr_small(qv, a.cw_t + ((size_t)ks * 32 + lane) * a.Ds, a.Ds);
                    lut2[ks * 64 + lane] = v;
                    lut2[ks * 64 + lane + 32] = v;
                }
            } else {
#pragma unroll 1
                for (int ks = wid; ks < 256; ks += NW) {
                    float v = 0.f;
                    if (ks < a.Ks) v = l2sqr_lanes(qm, a.cw_t + ((size_t)ks * 32 + lane) * a.Ds, a.Ds, a.variant);
                    lut2[ks * 64 + lane] = v;
                    lut2[ks * 64 + lane + 32] = v;
                }
            }
        }
    }
    __syncthreads();
    if (dbg && threadIdx.x == 0 && !fused) dbg[1] = clock64();
    if (dbg && threadIdx.x == 0) dbg[4] = clock64();  // table ready

    const int npass = fused ? 2 : 1;
#pragma unroll 1
    for (int pass = 0; pass < npass; ++pass) {
        if (pass == 1) {
            // ---- between the passes: select + rank the w_eff nearest centers, plan (all in shared memory) ----------
            if (dbg && threadIdx.x == 0) dbg[5] = clock64();  // coarse pass done (the pass loop ended with a barrier)
            u64 *sel = reinterpret_cast<u64 *>(smem_raw + hi_reg0);          // the regions are idle now
            int *hist = reinterpret_cast<int *>(smem_raw + hi_reg0 + 4096);  // 256 keys above `sel`
            int np = cta_select_smallest<NW * 32>(pool_d, a.nlist, a.w_eff, sel, hist);
            if (wid == 0) {
                if (np < 0) {  // > 256 exact ties at the w-th distance: full sort of all (dist, index) keys
                    const int P = next_pow2(a.nlist);
                    for (int i = lane; i < P; i += 32) sel[i] = i < a.nlist ? (((u64)pool_d[i] << 32) | (u64)(uint32_t)i) : RII_KEY_MAX;
                    warp_sort_smem(sel, P, lane);
                    np = a.nlist;
                }
                if (dbg && lane == 0) { dbg[6] = clock64(); dbg[7] = np; }
                int *ranked_g = a.plan.ranked + (size_t)b * a.w_eff;
                for (int j = lane; j < a.w_eff; j += 32) {  // w_eff <= nlist, np >= w_eff
                    const int no = (int)key_id(sel[j]);
                    ranked_g[j] = no;
                    s_f[j] = a.plan.glob_len[no];
                    s_pre[j] = a.plan.pre_len ? a.plan.pre_len[no] : 0;
                    s_loc[j] = a.plan.loc_len[no];
                    s_off[j] = a.offsets[no];
                }
                __syncwarp();
                if (lane == 0) {
                    make_plan(a.plan, b, s_f, s_pre, s_loc, s_cum);
                    s_plan[0] = a.plan.flags[b] != 0 ? 0 : a.plan.J[b];
                    *cta_thr = RII_KEY_MAX;  // (thr_w is still all-MAX: the coarse pass does not use the warp lists)
                }
            }
            __syncthreads();
            if (dbg && threadIdx.x == 0) dbg[1] = clock64();
            J = s_plan[0];
            pc = a.codes;
            segm = true;
            direct = false;
            segw = 0;
            wt.k = a.k;
            wt.cap = next_pow2(wt.k + 32) < 64 ? 64 : next_pow2(wt.k + 32);
            wt.count = 0;
            set_range(J ? (long long)s_cum[J - 1] : 0, 1, 0);
            *reinterpret_cast<uint4 *>(smem_raw + myreg) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4 *>(smem_raw + myreg + 16) = make_uint4(0, 0, 0, 0);
            if (ntiles > 0) issue_tile_seg(0);
        }

        float accA = 0.f, accB = 0.f;
        uint32_t xprev = 0;
        // distance part of the CTA threshold: long linear scans re-read it once per tile and after every push (a
        // stale value is merely less strict); the short per-query IVF passes re-read it at every emission
        uint32_t thr_hi = 0xffffffffu;
        // local index of the candidate whose distance completes at the end of the current block: it started one
        // block earlier, so the first block completes nothing (index "-1" of the previous tile: fails eloc < cnt)
        uint32_t eloc = (uint32_t)(SK_J * lane + SK_J - 1 - SK_TILE_ROWS);
#pragma unroll 1
        for (int n = 0; n < ntiles; ++n) {
            if ((n & 1) == 0 && n > 0) {  // entering half A again: the stream continues from B's last row via the carry row
                unsigned char *reg = smem_raw + myreg;
                uint4 x0 = *reinterpret_cast<uint4 *>(reg + 32 + (2 * SK_J - 1) * 32);
                uint4 x1 = *reinterpret_cast<uint4 *>(reg + 32 + (2 * SK_J - 1) * 32 + 16);
                *reinterpret_cast<uint4 *>(reg) = x0;
                *reinterpret_cast<uint4 *>(reg + 16) = x1;
            }
            asm volatile("cp.async.wait_group 0;");  // tile n has landed
            __syncwarp();                            // rows were written by other lanes of the warp
            if constexpr (!IVF) thr_hi = reinterpret_cast<volatile uint32_t *>(cta_thr)[1];
            const uint32_t rbw = rb + 4 * ((n & 1) * SK_HALF_WORDS);
            SK_BLOCK(rbw)
            SK_EMIT((uint32_t)(base + eloc))
            eloc += SK_TILE_ROWS - SK_J + 1;
            __syncwarp();  // the other half's last reader finished with this block
            if (n + 1 < ntiles) {
                if (segm) issue_tile_seg(n + 1);
                else issue_tile(n + 1);
            }
#pragma unroll
            for (int i = 1; i < SK_J; ++i) {
                SK_BLOCK(rbw + 32 * i)
                SK_EMIT((uint32_t)(base + eloc))
                eloc += 1;
            }
        }
        if (ntiles > 0) {  // drain: 32 more steps complete the last row of every lane
            const uint32_t rbw = rb + 4 * (((ntiles - 1) & 1) * SK_HALF_WORDS + SK_HALF_WORDS);
            SK_BLOCK_CLAMP(rbw, myreg + SK_REGION_BYTES - 4)
            SK_EMIT((uint32_t)(base + eloc))
        }
        if (!(IVF && direct)) warp_compact(wt, cta_thr, lane);
        __syncthreads();
    }
    if (dbg && threadIdx.x == 0) dbg[2] = clock64();
    {   // CTA merge of the (sorted) warp lists, reusing the lut2 area for the keys
        __shared__ int s_cnt[NW];
        if (lane == 0) s_cnt[wid] = wt.count;
        __syncthreads();
        int tot = 0;
        for (int w2 = 0; w2 < NW; ++w2) tot += s_cnt[w2];
        const u64 *allkeys = reinterpret_cast<const u64 *>(smem_raw + keys_off);
        if (tot <= 256) {
            // small (the usual topk <= 16 case): one warp gathers and bitonic-sorts <= 256 keys with warp barriers only
            if (wid == 0) {
                u64 *mk = reinterpret_cast<u64 *>(smem_raw + lut_off);
                int o = 0;
                for (int w2 = 0; w2 < NW; ++w2) {
                    for (int i = lane; i < s_cnt[w2]; i += 32) mk[o + i] = allkeys[(size_t)w2 * capw + i];
                    o += s_cnt[w2];
                }
                warp_sort_any(mk, tot, lane);
                const int n = tot < a.k ? tot : a.k;
                if (a.out.final) {
                    for (int i = lane; i < n; i += 32) {
                        a.out.out_ids[(size_t)b * a.k + i] = a.out.id_base + (long long)key_id(mk[i]);
                        a.out.out_dists[(size_t)b * a.k + i] = key_dist(mk[i]);
                    }
                    if (lane == 0) a.out.out_counts[b] = n;
                } else {
                    u64 *dst = a.out.partial + ((size_t)b * gridDim.x + blockIdx.x) * a.k;
                    for (int i = lane; i < a.k; i += 32) dst[i] = i < n ? mk[i] : RII_KEY_MAX;
                }
            }
        } else {
            BlockTopk tk;
            const int mcap = next_pow2(NW * a.k + 1);
            tk.keys = reinterpret_cast<u64 *>(smem_raw + lut_off);
            tk.count = reinterpret_cast<int *>(smem_raw + lut_off + (size_t)mcap * 8 + 8);
            tk.thr = reinterpret_cast<u64 *>(smem_raw + lut_off + (size_t)mcap * 8);
            tk.cap = mcap;
            tk.k = a.k;
            tk.init();
            for (int w2 = 0; w2 < NW; ++w2)
                for (int i = threadIdx.x; i < s_cnt[w2]; i += blockDim.x) tk.push(allkeys[(size_t)w2 * capw + i]);
            emit_topk(tk, a.out, b, blockIdx.x, gridDim.x);
        }
    }
    if (dbg && threadIdx.x == 0) dbg[3] = clock64();
}

// Host-side sizing.  The kernel always gets the maximum dynamic shared memory; what varies is how many warp regions
// fit below and above the 64 KB table pinned at absolute shared address 0x10000 (the window itself starts at
// ~1 KB: reserved + static shared memory; both extremes are allowed for).
#define SK_DYN_SMEM (227 * 1024 - 64)
static inline int skew_regions_fit(bool ivf, int nw, int capw, int w_eff)
{
    const size_t meta = (((size_t)nw * capw * 8 + 16 + (size_t)nw * 8 + (ivf ? (size_t)w_eff * 24 + 32 : 0)) + 15) & ~(size_t)15;
    const long long lut_off_min = 0x10000 - 2048, lut_off_max = 0x10000 - 1024;
    const long long n_lo = (lut_off_min - (long long)meta) / SK_WARP_BYTES;
    const long long n_hi = ((long long)SK_DYN_SMEM - (lut_off_max + SK_LUT_BYTES)) / SK_WARP_BYTES;
    return (int)((n_lo < 0 ? 0 : n_lo) + (n_hi < 0 ? 0 : n_hi));
}
static inline size_t skew_smem_bytes(int nw, bool ivf, int capw, int w_eff)
{
    return (size_t)SK_LUT_BYTES + (size_t)nw * SK_WARP_BYTES + (size_t)nw * capw * 8 + 16 + 64 +
           (ivf ? (size_t)w_eff * 24 + 32 : 0);
}

#include "scan_dual.cuh"
#include "scan_stream.cuh"
