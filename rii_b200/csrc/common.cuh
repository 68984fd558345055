// Shared device helpers of the ADC path: the reference's fp32 arithmetic (explicit round-to-nearest sub / mul / add so
// that nvcc never contracts to FMA and never reassociates: bit-identical to the reference code as written,
// oracle/_ref/strict_*, oracle/rii_oracle.cpp), the top-k output descriptor, the posting-list walk plan
// and the argument block of the streaming scan engine.
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

// One table entry for Ds <= 4 from the zero-padded float4 copies of the query sub-vector and the codeword:
// (s0 + s1) + (s2 + s3), s_i = (q_i - c_i)^2, absent lanes contribute +0 (src/distance.h:148-169).  The subtracts and the
// multiplies are packed (FADD2 / FMUL2: two lanes per instruction, each rounded like the scalar op); the adds stay scalar --
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 despite the explicit rounding modifiers, which would change
// the bits.
__device__ __forceinline__ float sqdist4(float2 qa, float2 qb, float4 c)
{
    float2 ta = __fadd2_rn(qa, make_float2(-c.x, -c.y)), tb = __fadd2_rn(qb, make_float2(-c.z, -c.w));
    ta = __fmul2_rn(ta, ta);
    tb = __fmul2_rn(tb, tb);
    return __fadd_rn(__fadd_rn(ta.x, ta.y), __fadd_rn(tb.x, tb.y));
}

static __device__ __noinline__ float l2sqr_lanes(const float *__restrict__ x, const float *__restrict__ y, int d, int variant)
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


// Emit the CTA's sorted top-k: either final (ids/dists/count) or a partial key list for k_merge.
struct TopkOut {
    u64 *partial;        // (B, parts, k) keys, RII_KEY_MAX padded   (when !final)
    long long *out_ids;  // (B, k) global ids                         (when final)
    float *out_dists;    // (B, k)
    int *out_counts;     // (B)
    long long id_base;
    const long long *id_map;  // when set: key id -> global id (subset scans over a compact copy of the target rows)
    int final;
    int *merge_cnt;      // (B) zero on entry, when !final: the LAST CTA of a query to deliver its partial list merges all of
                         // them into out_ids / out_dists / out_counts itself (and resets the counter) -- no separate k_merge launch
};
__device__ __forceinline__ void emit_topk(BlockTopk &tk, const TopkOut &o, int b, int part, int parts)
{
    tk.compact();
    int n = *tk.count;
    int k = tk.k;
    if (o.final) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            u64 key = tk.keys[i];
            o.out_ids[(size_t)b * k + i] = o.id_map ? o.id_map[key_id(key)] : o.id_base + (long long)key_id(key);
            o.out_dists[(size_t)b * k + i] = key_dist(key);
        }
        if (threadIdx.x == 0) o.out_counts[b] = n;
    } else {
        u64 *dst = o.partial + ((size_t)b * parts + part) * k;
        for (int i = threadIdx.x; i < k; i += blockDim.x) dst[i] = i < n ? tk.keys[i] : RII_KEY_MAX;
    }
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
static __device__ void make_plan(const PlanArgs &p, int b, const int *f_by_rank, const int *pre_by_rank, const int *loc_by_rank,
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

struct SkewArgs {
    const float *T;            // (B, 32*Ks), or null: build the table in-kernel from Q / cw (K1 fused)
    const float *Q;            // (B, 32*Ds)
    const float *cw;           // (32, Ks, Ds)
    const float *cw_t;         // (Ks, 32, Ds): the same codewords, sub-space fastest (coalesced in-kernel table build)
    int Ds, variant;
    int M;                     // sub-spaces (the engine pads rows to 32 or 64 bytes)
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
    int coarse_mode;           // IVF: 0 = one launch does everything; 1 = coarse pass only (write plan.ranked, no scan);
                               //      2 = no coarse pass: the ranking is read from plan.ranked, the plan is made in-kernel
    int coarse_lists;          // v4 fused: rank the centers with the warps' top-k lists (nlist > 1024) instead of keeping every distance
    PlanArgs plan;             // IVF fused: plan inputs (lengths, L, topk, w) and its global outputs (ranked, J, flags)
    TopkOut out;
    long long *dbg;            // optional: per-CTA clock64() at [start, table ready, scan done, end] (tools/microbench.py)
    int q_inline;              // single host call with D <= 128, Ds <= 4: the query vector travels in the kernel parameters (qv) --
    float qv[128];             //   the launch carries it, no read of the pinned host buffer over PCIe before the table build can start
};
