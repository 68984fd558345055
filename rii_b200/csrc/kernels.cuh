// sm_100a kernels of the ADC hot path in the NATURAL code layout (one candidate per lane, table lut[m][ks]): every M,
// unsorted target_ids, large top-k.  The streaming engine (scan_stream.cuh) serves the shapes that matter for speed.
// Reference semantics are cited per kernel (file:line into matsui528/rii v0.2.12).
#pragma once
#include "common.cuh"

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
                                                              const int *__restrict__ counts, long long stride_bytes, int G, int B, int k,
                                                              int P, long long *out_ids, float *out_dists, int *out_counts)
{
    // stride_bytes == 0: three contiguous arrays (G, B, k) / (G, B, k) / (G, B).  Otherwise shard g's [ids | dists | counts]
    // block starts g * stride_bytes after each base pointer (one packed all-gather buffer).
    extern __shared__ __align__(16) unsigned char smem_raw[];
    long long *s_id = reinterpret_cast<long long *>(smem_raw);
    uint32_t *s_d = reinterpret_cast<uint32_t *>(smem_raw + (size_t)P * 8);
    const int b = blockIdx.x;
    const size_t s_ids = stride_bytes ? (size_t)stride_bytes / 8 : (size_t)B * k, s_dst = stride_bytes ? (size_t)stride_bytes / 4 : (size_t)B * k,
                 s_cnt = stride_bytes ? (size_t)stride_bytes / 4 : (size_t)B;
    int total = 0;
    for (int g = 0; g < G; ++g) total += counts[g * s_cnt + b];
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        int g = i / k, j = i % k;
        bool valid = g < G && j < counts[g * s_cnt + b];
        s_id[i] = valid ? ids[g * s_ids + (size_t)b * k + j] : 0x7fffffffffffffffll;
        s_d[i] = valid ? __float_as_uint(dists[g * s_dst + (size_t)b * k + j]) : 0xffffffffu;
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

// ---------------------------------------------------------------------------------------------------
// Subset searches (target_ids).  src/rii.h:218-228 (linear: exactly the given ids) and :294 (IVF: binary_search filter
// while walking the lists).  Both run as ordinary scans over a temporary SUB-INDEX built from the target ids:
//   linear: a compact copy of the target rows (one segment, in the given order), ids mapped back through the target ids;
//   IVF   : the posting lists restricted to the members (a stable sort of the (list, id) pairs of the targets keeps every
//           list ascending in id), so "skip non-members, stop after L members" IS the plain walk over the sub-lists.
// ---------------------------------------------------------------------------------------------------
// flags[0] |= 1 if not ascending; flags[1] = ids < lo; flags[2] = ids >= hi  (the in-range run of a sorted array)
__global__ void k_tids_scan(const long long *__restrict__ tids, long long S, long long lo, long long hi, int *flags)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const long long t = tids[i];
    if (i > 0 && tids[i - 1] > t) atomicOr(flags, 1);
    if (t < lo) atomicAdd(flags + 1, 1);
    if (t >= hi) atomicAdd(flags + 2, 1);
}

// rows[i] = tids[i] - id_base (the caller passes the in-range run)
__global__ void k_tids_to_rows(const long long *__restrict__ tids, long long n, long long id_base, int *__restrict__ rows)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rows[i] = (int)(tids[i] - id_base);
}

// (list, row) pair of every target id: list = assign[row]; ids outside this shard, repeated ids (the reference tests
// membership, so a posting-list id is a candidate once) and rows that are in no list get the key ~0 and sort to the end
__global__ void k_sub_keys(const long long *__restrict__ tids, long long S, long long id_base, long long N, const int *__restrict__ assign,
                           uint32_t *__restrict__ keys, uint32_t *__restrict__ rows)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const long long t = tids[i], r = t - id_base;
    uint32_t key = 0xffffffffu;
    if (r >= 0 && r < N && !(i > 0 && tids[i - 1] == t)) key = (uint32_t)assign[r];
    keys[i] = key;
    rows[i] = (uint32_t)(r >= 0 && r < N ? r : 0);
}

// per-list lengths and skew64 segment offsets from the CSR bounds of the sorted pairs.  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024) k_sub_layout(const long long *__restrict__ bounds, int nlist, int H, int *__restrict__ len,
                                                      long long *__restrict__ skew_off)
{
    __shared__ long long s_part[1024];
    const int per = (nlist + 1023) / 1024;
    const int i0 = threadIdx.x * per, i1 = min(nlist, i0 + per);
    long long sum = 0;
    for (int i = i0; i < i1; ++i) {
        const long long l = bounds[i + 1] - bounds[i];
        len[i] = (int)l;
        sum += 64 * ((l + 63) / 64 + 1) * H;
    }
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const long long y = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += y;
        __syncthreads();
    }
    long long run = threadIdx.x ? s_part[threadIdx.x - 1] : 0;
    for (int i = i0; i < i1; ++i) {
        skew_off[i] = run;
        const long long l = bounds[i + 1] - bounds[i];
        run += 64 * ((l + 63) / 64 + 1) * H;
    }
    if (threadIdx.x == 1023) skew_off[nlist] = s_part[1023];
}

// ---------------------------------------------------------------------------------------------------
// General (global-memory) path: every M, any topk (the reference allows topk = N, rii/rii.py:280-281) and any number
// of ranked lists.  Distances go to HBM as (dist, id) keys, a segmented radix sort (device_sort.cu) orders them.
// ---------------------------------------------------------------------------------------------------
// keys[b][i] for the ncand candidates of a linear scan (all rows, or the given target ids).  grid (blocks, B)
template <int M_T>
__global__ void __launch_bounds__(RII_THREADS) k_keys_linear(LinearArgs a, u64 *__restrict__ keys, long long stride)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *lut = reinterpret_cast<float *>(smem_raw);
    const int b = blockIdx.y;
    load_lut(lut, a.T + (size_t)b * a.M * a.Ks, a.M * a.Ks);
    __syncthreads();
    const long long ncand = a.S ? a.S : a.N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ncand; i += (long long)gridDim.x * blockDim.x) {
        const long long r = a.S ? a.tids[i] - a.id_base : i;
        u64 key = RII_KEY_MAX;
        if (r >= 0 && r < a.N) key = pack_key(adc_row<M_T, false>(lut, a.Ks, a.M, a.codes + r * a.M), (uint32_t)r);
        keys[(size_t)b * stride + i] = key;
    }
}

// coarse keys[b][no] = (ADC distance to center no, no).  grid (blocks, B)
template <int M_T>
__global__ void __launch_bounds__(RII_THREADS) k_coarse_keys(const float *__restrict__ T, const uint8_t *__restrict__ centers, int nlist, int M,
                                                             int Ks, u64 *__restrict__ keys)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *lut = reinterpret_cast<float *>(smem_raw);
    const int b = blockIdx.y;
    load_lut(lut, T + (size_t)b * M * Ks, M * Ks);
    __syncthreads();
    for (int no = blockIdx.x * blockDim.x + threadIdx.x; no < nlist; no += gridDim.x * blockDim.x)
        keys[(size_t)b * nlist + no] = pack_key(adc_row<M_T, false>(lut, Ks, M, centers + (size_t)no * M), (uint32_t)no);
}

// ranked[b][j] and the list lengths by rank (what k_plan consumes) from the sorted coarse keys.  grid (blocks, B)
__global__ void k_rank_gather(const u64 *__restrict__ sorted, int nlist, int w_eff, const int *__restrict__ glob_len, const int *__restrict__ pre_len,
                              const int *__restrict__ loc_len, int *__restrict__ ranked, int *__restrict__ f, int *__restrict__ pre, int *__restrict__ loc)
{
    const int b = blockIdx.y;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < w_eff; j += gridDim.x * blockDim.x) {
        const int no = (int)key_id(sorted[(size_t)b * nlist + j]);
        const size_t o = (size_t)b * w_eff + j;
        ranked[o] = no;
        f[o] = glob_len[no];
        pre[o] = pre_len ? pre_len[no] : 0;
        loc[o] = loc_len[no];
    }
}

// keys of the planned candidates of every query: position p < cum[J - 1] -> (segment, offset) -> id -> ADC.  grid (blocks, B)
template <int M_T>
__global__ void __launch_bounds__(RII_THREADS) k_ivf_keys(IvfArgs a, u64 *__restrict__ keys, long long stride, int *__restrict__ ncand)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *lut = reinterpret_cast<float *>(smem_raw);
    const int b = blockIdx.y;
    load_lut(lut, a.T + (size_t)b * a.M * a.Ks, a.M * a.Ks);
    __syncthreads();
    const int J = (a.flags[b] != 0) ? 0 : a.J[b];
    const int *cum = a.cum + (size_t)b * a.w_eff, *ranked = a.ranked + (size_t)b * a.w_eff;
    const int total = J ? cum[J - 1] : 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) ncand[b] = total;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = J - 1;  // first segment with cum > p
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cum[mid] > p) hi = mid; else lo = mid + 1;
        }
        const int id = __ldg(a.ids + a.offsets[ranked[lo]] + (p - (lo ? cum[lo - 1] : 0)));
        keys[(size_t)b * stride + p] = pack_key(adc_row<M_T, false>(lut, a.Ks, a.M, a.codes + (size_t)id * a.M), (uint32_t)id);
    }
}

// OPQ rotation of the queries (rii/rii.py:305-306: fine_quantizer.rotate(q) = q @ R): out[b][j] = sum_i Q[b][i] R[i][j],
// fp32 FMA chain in i order.  (A D x D contraction per query; tensor cores are deliberately not used: bf16 / tf32 inputs
// would move the rotated coordinates by ~1e-3 relative, far beyond the 1e-5 contract on the distances.)  grid (ceil(D/128), B)
__global__ void k_rotate(const float *__restrict__ Q, const float *__restrict__ R, int D, float *__restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (j >= D) return;
    const float *q = Q + (size_t)b * D;
    float acc = 0.f;
    for (int i = 0; i < D; ++i) acc = fmaf(__ldg(q + i), __ldg(R + (size_t)i * D + j), acc);
    out[(size_t)b * D + j] = acc;
}

// out[i] = Q[idx[i]] (rows of D floats); and the way back for results
__global__ void k_gather_queries(const float *__restrict__ Q, const int *__restrict__ idx, int n, int D, float *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * D) return;
    out[i] = Q[(size_t)idx[i / D] * D + i % D];
}
__global__ void k_scatter_results(const int *__restrict__ idx, int n, int k, const long long *__restrict__ ids, const float *__restrict__ d,
                                  const int *__restrict__ c, long long *__restrict__ out_ids, float *__restrict__ out_d, int *__restrict__ out_c)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * k) return;
    const int r = (int)(i / k), j = (int)(i % k), b = idx[r];
    if (j < c[r]) {
        out_ids[(size_t)b * k + j] = ids[i];
        out_d[(size_t)b * k + j] = d[i];
    }
    if (j == 0) out_c[b] = c[r];
}

// seg[b] = b * stride, seg_end[b] = b * stride + count[b] (count null: n)
__global__ void k_seg_bounds(int B, long long stride, const int *__restrict__ count, long long n, long long *__restrict__ beg, long long *__restrict__ end)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    beg[b] = (long long)b * stride;
    end[b] = (long long)b * stride + (count ? count[b] : n);
}

// the first min(k, valid) keys of every sorted segment -> outputs.  grid (blocks, B)
__global__ void k_take_sorted(const u64 *__restrict__ sorted, long long stride, const int *__restrict__ count, long long n, int k, TopkOut out)
{
    const int b = blockIdx.y;
    const long long c = count ? count[b] : n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k && i < c; i += gridDim.x * blockDim.x) {
        const u64 key = sorted[(size_t)b * stride + i];
        if (key != RII_KEY_MAX) {
            out.out_ids[(size_t)b * k + i] = out.id_map ? out.id_map[key_id(key)] : out.id_base + (long long)key_id(key);
            out.out_dists[(size_t)b * k + i] = key_dist(key);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // keys are ascending and RII_KEY_MAX pads: count = index of the first pad
        long long lo = 0, hi = c < k ? c : k;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (sorted[(size_t)b * stride + mid] == RII_KEY_MAX) hi = mid; else lo = mid + 1;
        }
        out.out_counts[b] = (int)lo;
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
// grid (K, M), Ks <= 256 threads.  Empty clusters keep their center (src/pqkmeans.cpp:115-120).
__global__ void __launch_bounds__(256) k_vote_centers(const float *__restrict__ Dm, const int *__restrict__ hist,
                                                      int M, int Ks, uint8_t *centers)
{
    const int k = blockIdx.x, m = blockIdx.y, k2 = threadIdx.x;
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


