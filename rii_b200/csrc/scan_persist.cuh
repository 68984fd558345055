// ===================================================================================================
// K1 + K4 + K5 for per-query IVF batches as a PERSISTENT, WARP-SPECIALISED kernel (rows of 32 bytes).
//
// What the phase clocks of k_scan_stream32 show for a batch of independent queries (profiles/r01_micro_ivf_*, r02_phase_*):
// a query's CTA spends 20-25 % of its cycles outside the scan -- table build (K1), coarse pass + selection + plan (K4),
// final merge -- and during those phases its warps stream nothing; two co-resident CTAs overlap that only partly, and a
// grid of one CTA per query pays wave quantisation on top (1024 queries over 296 slots = 4 waves for 3.46 waves of work).
//
// Here one CTA per SM stays resident and walks its queries (b = blockIdx.x, += gridDim.x).  11 CONSUMER warps do nothing
// but stream the planned posting-list segments of query i through their private cp.async rings (the engine of
// scan_stream.cuh: skew64 layout, conflict-free table lut2[ks][64], packed FFMA2 accumulation -- bit-identical sequential
// fp32 sums, src/rii.h:386-394).  One PRODUCER warp prepares query i + 1 meanwhile, in the other half of double-buffered
// shared memory: distance table (src/rii.h:361-373), coarse pass over the skew64 centers with that table, selection of the
// w nearest lists (histogram select), the candidate plan in prefix-sum form (plan_warp) -- or, when the
// ranking comes from a separate coarse launch (sharded batches), just table + plan -- and it merges the consumers' top-k
// lists of query i - 1 into the output.  Hand-over by four mbarriers in shared memory: FULL[p] (one arrival, the producer's:
// table, segments, key buffers of parity p are ready) and DONE[p] (one arrival per consumer warp: its result of parity p is
// final).  A consumer warp waits for FULL on its own -- the warps of a CTA are NOT synchronised with each other at a query
// boundary (the named-barrier version of this kernel lost 6-9 % of the consumers' time there, profiles/r02_ncu_*_persist_v1*):
// a warp that finishes its slice early starts on the next query while the others are still scanning.
//
// The walk of a warp is a list of RUNS: consecutive 2 KB blocks of one posting list (its groups inside the warp's slice
// plus the drain block behind them, which is simply the next block in memory).  Inside a run the issue logic is a pointer
// increment and a counter; segment tables are only read when a run starts (~5 times per query and warp).  (Bulk L2 prefetches
// -- cp.async.bulk.prefetch.L2, UBLKPF -- running 8 blocks ahead of the cp.async front were measured on HBM-resident lists:
// 0.463 vs 0.465 ms for the C5-shaped batch, i.e. nothing: the scan is bound on the SM side, not by memory latency.  Removed.
// Also measured and removed: OR-merging the drain block of a list with the first block of the next one (complementary zero
// padding; consumed together, the second block's pipeline instance only refills the ring).  It saves 34 of a C2 query's 606
// blocks, but the two extra warp-uniform branches per block and the shorter prefetch distance of merged blocks cost more:
// 4.42 -> 4.24 M queries/s (tools/gpu_r2_x.sh).)
//
// topk == 1 (the recall@1 operating point of every BASELINE config) has its own instantiation: a warp keeps its best
// (distance, position) in two warp-uniform registers -- no key buffers, no compaction, no shared-memory traffic; a group
// is looked at more closely only when some lane is at or below the warp's best distance.  Exact ties across posting lists
// are resolved by id (one global load each, rare); the id of a warp's winner is looked up once per query.
//
// Shared memory (227 KB, absolute addresses; the dynamic window starts at 0x400):
//   0x00400  segments[2] | key buffers[2] | plan inputs | histogram | coarse distances (nlist <= 1024) | selection keys
//            | producer ring | 5 consumer rings
//   0x10000  table of parity 0 (64 KB)        0x20000  table of parity 1 (64 KB)
//   0x30000  6 consumer rings (3 stages x 2 KB each)                                                   0x39000 end
// The table's 64 KB-aligned base travels in the upper half of the lane's column register, so the lookup address is still
// one PRMT (ks into byte 1) + the immediate 4 t.
// Limits: 12 <= M <= 32, topk <= 16, w_eff <= 64, nlist <= 1024 when the coarse pass is fused.  Everything else runs on
// k_scan_stream32.
// ===================================================================================================
#pragma once
#include "scan_stream.cuh"

#define PS_NC 11
#define PS_R 3
#define PS_WMAX 64
#define PS_CAPW 64
#define PS_MAXK 16
#define PS_NLIST_MAX 1024
#define PS_T0 0x10000u
#define PS_RING_BYTES (PS_R * ST_BLOCK_BYTES)
#define PS_SMEM_BYTES (0x39000 - 0x400)

struct PsSeg {  // per-parity query state: written by the producer, read by the consumers
    long long off[PS_WMAX], prow[PS_WMAX];
    int gcum[PS_WMAX], take[PS_WMAX];
    int J, b, plain, pad;
};
struct PsKeys {  // per-parity top-k state of the consumers
    u64 keys[PS_NC * PS_CAPW];
    u64 cta_thr;
    u64 thr_w[PS_NC + 1];
    int cnt[PS_NC + 1];
};
// offsets from the start of the dynamic window (absolute 0x400)
#define PS_OFF_SEG 0
#define PS_OFF_KEYS (PS_OFF_SEG + 2 * (int)sizeof(PsSeg))
#define PS_OFF_BAR (PS_OFF_KEYS + 2 * (int)sizeof(PsKeys))           /* mbarriers: FULL[2], DONE[2] */
#define PS_OFF_PLAN (PS_OFF_BAR + 32)                                 /* s_f, s_pre, s_loc: 3 x PS_WMAX ints */
#define PS_OFF_HIST (PS_OFF_PLAN + 3 * PS_WMAX * 4)                   /* 256 + 4 ints */
#define PS_OFF_POOL (PS_OFF_HIST + 260 * 4)                           /* PS_NLIST_MAX words */
#define PS_OFF_SELK (((PS_OFF_POOL + PS_NLIST_MAX * 4) + 15) & ~15)   /* 256 keys, followed by the producer ring: 1024 keys for the full-sort fallback */
#define PS_OFF_PRING (PS_OFF_SELK + 256 * 8)
#define PS_OFF_RINGS_A (PS_OFF_PRING + PS_RING_BYTES)
#define PS_RINGS_A 5
#define PS_OFF_RINGS_B (0x30000 - 0x400)
static_assert(PS_OFF_RINGS_A + PS_RINGS_A * PS_RING_BYTES <= 0x10000 - 0x400, "the low region overflows into the tables");
static_assert(PS_OFF_RINGS_B + (PS_NC - PS_RINGS_A) * PS_RING_BYTES <= PS_SMEM_BYTES, "the high region overflows");
static_assert(sizeof(PsSeg) % 16 == 0 && sizeof(PsKeys) % 8 == 0, "alignment");

__device__ __forceinline__ void ps_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// one arrival (release at CTA scope: everything this thread -- and, after a __syncwarp, its warp -- wrote before is visible to
// a thread that has seen the phase complete)
__device__ __forceinline__ void ps_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// wait until the phase with the given parity has completed (acquire); returns the cycles spent waiting when `timed`.
// backoff_ns > 0: sleep between polls -- the producer's waits are long (it is idle 70 % of the time when the ranking is given;
// its try_wait loop was 8 % of all executed instructions, profiles/r02_ncu_c5shape_persist_final_sass_hot.txt).  Measured: the
// scan time does not change (0.4356 vs 0.4360 ms) -- try_wait already suspends the thread -- so this only saves issue slots.
__device__ __forceinline__ long long ps_mbar_wait(uint32_t bar, uint32_t parity, bool timed, unsigned backoff_ns = 0)
{
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return 0;
    const long long t0 = timed ? clock64() : 0;
    do {
        if (backoff_ns) __nanosleep(backoff_ns);
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
    return timed ? clock64() - t0 : 0;
}

// the next block of the warp's walk goes into ring stage S; D describes it (scan_stream.cuh ST_ISSUE): flattened group << 2
// | y row valid << 1 | x row valid, 0 for a drain block.  Warp-uniform state: run_left (blocks left in the run, drain block
// included), bp (this lane's 16-byte column of the run's next block), cur_f, last_bits (valid bits of the run's last group).
#define PS_ISSUE(S, D, ACTIVE)                                                                                \
    {                                                                                                         \
        if (ACTIVE) {                                                                                         \
            if (run_left == 0) next_run();                                                                    \
            if (run_left == 1) D = 0u;                                                                        \
            else { D = ((uint32_t)cur_f << 2) | (run_left == 2 ? last_bits : 3u); ++cur_f; }                  \
            const uint32_t dst_ = ring + (S) * ST_BLOCK_BYTES;                                                \
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_), "l"(bp));                   \
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_ + 512), "l"(bp + 512));       \
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_ + 1024), "l"(bp + 1024));     \
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_ + 1536), "l"(bp + 1536));     \
            bp += ST_BLOCK_BYTES;                                                                             \
            --run_left;                                                                                       \
        }                                                                                                     \
        asm volatile("cp.async.commit_group;");                                                               \
    }
// one pipeline stage (rows of 32 bytes: every block completes the rows of the previous group; scan_stream.cuh ST_STAGE)
#define PS_STAGE(S)                                                                                           \
    if (m + (S) < nblk) {                                                                                     \
        constexpr int h_ = 0;                                                                                 \
        const uint32_t dsel_ = dsc[((S) + ST_D) % ST_R];                                                      \
        PS_ISSUE(((S) + ST_D) % ST_R, dsc[((S) + ST_D) % ST_R], m + (S) + ST_D < nblk)                        \
        asm volatile("cp.async.wait_group 2;" ::: "memory");                                                  \
        uint32_t wx_[8], wy_[8];                                                                              \
        ST_LDS128(wx_, ring + (S) * ST_BLOCK_BYTES);                                                          \
        ST_LDS128(wx_ + 4, ring + (S) * ST_BLOCK_BYTES + 512);                                                \
        ST_LDS128(wy_, ring + (S) * ST_BLOCK_BYTES + 1024);                                                   \
        ST_LDS128(wy_ + 4, ring + (S) * ST_BLOCK_BYTES + 1536);                                               \
        ST_BLOCK(wx_, wy_)                                                                                    \
        float dx_, dy_;                                                                                       \
        asm("mov.b64 {%0, %1}, %2;" : "=f"(dx_), "=f"(dy_) : "l"(out2));                                      \
        out2 = 0ull;                                                                                          \
        emit2(dx_, dy_, dsel_);                                                                               \
    }

// the w smallest of np (distance bits, index) pairs by ONE warp: 256-bin histogram over [mn, mx], the bin holding the
// w-th smallest, gather of everything up to that bin as (dist, index) keys, register sort.  Returns the number of keys in
// `out` (>= w, <= 256) or -1 (more than 256 qualify: heavily tied distances).  (CTA-wide form: cta_select_smallest.)
__device__ __forceinline__ int warp_select_smallest(const uint32_t *d, int np, int w, u64 *out, int *hist, uint32_t mn, uint32_t mx, int lane)
{
    // np <= PS_NLIST_MAX = 1024: the lane's 32 values stay in registers for both passes (every load is independent: one
    // shared-memory latency for the whole read instead of one per dependent iteration)
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = j * 32 + lane < np ? d[j * 32 + lane] : 0xffffffffu;
#pragma unroll
    for (int j = 0; j < 8; ++j) hist[8 * lane + j] = 0;
    __syncwarp();
    const uint32_t range = mx - mn;
    const int sh = range >= 256u ? (32 - __clz(range)) - 8 : 0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
        if (j * 32 + lane < np) atomicAdd(&hist[(v[j] - mn) >> sh], 1);
    __syncwarp();
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
    int n = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const bool ok = j * 32 + lane < np && (int)((v[j] - mn) >> sh) <= bin;
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (bal) {
            if (n + __popc(bal) > 256) return -1;
            if (ok) out[n + __popc(bal & ((1u << lane) - 1u))] = ((u64)v[j] << 32) | (u64)(uint32_t)(j * 32 + lane);
            n += __popc(bal);
        }
    }
    warp_sort_any(out, n, lane);
    return n;
}

// grid (min(B, SMs)); 12 warps; dynamic shared memory PS_SMEM_BYTES (and no static shared memory: the window must start
// at absolute shared address 0x400).  a.coarse_mode: 0 = coarse pass fused (a.centers = skew64 of the centers), 2 = the
// ranking is read from a.plan.ranked.  nq = number of queries.  K1: the topk == 1 instantiation.
template <bool K1>
__global__ void __launch_bounds__((PS_NC + 1) * 32, 1) k_scan_persist32(SkewArgs a, int nq)
{
    constexpr int ST_R = PS_R, ST_D = PS_R - 1;  // (rows of 32 bytes: H = 1)
    constexpr uint32_t TB = 0;  // the table base is in colreg
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    if (smem_base != 0x400u) __trap();  // the layout above is in absolute shared addresses
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    PsSeg *segs = reinterpret_cast<PsSeg *>(smem_raw + PS_OFF_SEG);
    PsKeys *keyb = reinterpret_cast<PsKeys *>(smem_raw + PS_OFF_KEYS);
    const uint32_t bar_full = smem_base + PS_OFF_BAR, bar_done = bar_full + 16;  // [parity] x 8 bytes
    const int Mr = a.M;
    float keep[32], sel[32];  // the per-lane row-boundary constants of the accumulation (scan_stream.cuh)
#pragma unroll
    for (int t = 0; t < 32; ++t) {
        keep[t] = lane == t ? 0.f : 1.f;
        sel[t] = lane == t ? 1.f : 0.f;
    }
    long long *dbg = a.dbg ? a.dbg + (size_t)blockIdx.x * 16 : nullptr;  // optional cycle counters per CTA
    const bool timed = dbg != nullptr;
    long long c_wait = 0, c_scan = 0;
    const int n_my = blockIdx.x < nq ? (nq - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (threadIdx.x == 0) {
        ps_mbar_init(bar_full, 1);
        ps_mbar_init(bar_full + 8, 1);
        ps_mbar_init(bar_done, PS_NC);
        ps_mbar_init(bar_done + 8, PS_NC);
    }
    // The table of the CTA's FIRST query is built by all 12 warps (rows ks = wid, wid + 12, ...: 22 rows per warp instead
    // of 256 for the producer alone) -- otherwise the consumers idle for a whole single-warp table build (~25 K cycles) at
    // the start, 4 % of a batch of 1024 queries (7 per CTA).  Ds <= 4 only; other shapes leave it to the producer.
    const bool first_coop = n_my > 0 && a.Ds <= 4;
    int bad0 = 0;
    if (first_coop) {
        float *lut2 = reinterpret_cast<float *>(smem_raw + (PS_T0 - 0x400u));
        const int m = lane;
        if (m < Mr) {
            const float *qm = a.Q + (size_t)blockIdx.x * Mr * a.Ds + (size_t)m * a.Ds;
            float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
            q4.x = __ldg(qm);
            if (a.Ds > 1) q4.y = __ldg(qm + 1);
            if (a.Ds > 2) q4.z = __ldg(qm + 2);
            if (a.Ds > 3) q4.w = __ldg(qm + 3);
            const float2 qa = make_float2(q4.x, q4.y), qb = make_float2(q4.z, q4.w);
            const float4 *cw4 = reinterpret_cast<const float4 *>(a.cw_t) + m;
#pragma unroll 8
            for (int ks = wid; ks < a.Ks; ks += PS_NC + 1) {
                const float v = sqdist4(qa, qb, __ldg(cw4 + ks * Mr));
                bad0 |= !(v <= ST_TABLE_LIMIT);
                lut2[ks * 64 + m + 32] = v;
                lut2[ks * 64 + m] = v;
            }
        } else {
            for (int ks = wid; ks < a.Ks; ks += PS_NC + 1) {
                lut2[ks * 64 + m + 32] = 0.f;
                lut2[ks * 64 + m] = 0.f;
            }
        }
    }
    bad0 = __syncthreads_or(bad0);  // (the only CTA-wide barrier of the kernel; also publishes the mbarrier inits)

    if (wid < PS_NC) {
        // =========================================== consumers ===========================================
        const uint32_t ring = smem_base + (wid < PS_RINGS_A ? PS_OFF_RINGS_A + wid * PS_RING_BYTES : PS_OFF_RINGS_B + (wid - PS_RINGS_A) * PS_RING_BYTES) +
                              lane * 16;
        const uint8_t *pc = a.codes;
#pragma unroll 1
        for (int qi = 0; qi < n_my; ++qi) {
            const int p = qi & 1;
            c_wait += ps_mbar_wait(bar_full + 8 * p, (uint32_t)(qi >> 1) & 1u, timed);
            long long t0_ = 0;
            if (timed) t0_ = clock64();
            PsSeg &sg = segs[p];
            PsKeys &kb = keyb[p];
            const int J = sg.J;
            const int *s_gcum = sg.gcum, *s_take = sg.take;
            const long long *s_off = sg.off, *s_prow = sg.prow;
            const float *lut2 = reinterpret_cast<const float *>(smem_raw + (PS_T0 - 0x400u) + p * SK_LUT_BYTES);
            const uint32_t colreg = (PS_T0 + (uint32_t)p * SK_LUT_BYTES) | (uint32_t)((32 - lane) * 4);
            auto seg_of = [&](int f) -> int {  // segment holding flattened group f
                int lo = 0, hi = J - 1;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (s_gcum[mid] > f) hi = mid; else lo = mid + 1;
                }
                return lo;
            };
            // id of the candidate at position pos = flattened group << 6 | row of the group (lazy: only winners / ties are looked up)
            auto id_of = [&](uint32_t pos) -> uint32_t {
                const int f = (int)(pos >> 6), j = seg_of(f);
                const int r = (f - (j ? s_gcum[j - 1] : 0)) * 64 + (int)(pos & 63u);
                return (uint32_t)__ldg(a.ids + s_off[j] + r);
            };
            // ---- this warp's slice [f0, f_end) of the query's flattened 64-row groups, as runs of consecutive blocks ----
            int f0 = 0, f_end = 0, cur_f = 0, nblk = 0, run_left = 0;
            int seg = 0, seg_g0 = 0, seg_gend = 0;
            uint32_t last_bits = 3u;
            const uint8_t *bp = pc;
            auto next_run = [&]() {  // (warp-uniform; cur_f < f_end)
                if (cur_f == seg_gend) {
                    ++seg;
                    seg_g0 = seg_gend;
                    seg_gend = s_gcum[seg];
                }
                const int end = seg_gend < f_end ? seg_gend : f_end;
                run_left = end - cur_f + 1;  // + the drain block: the next block in memory
                bp = pc + (size_t)s_prow[seg] * 32 + (size_t)(cur_f - seg_g0) * ST_BLOCK_BYTES + lane * 16;
                int tail = 64;
                if (end == seg_gend) tail = s_take[seg] - (seg_gend - seg_g0 - 1) * 64;  // rows of the list's last group
                last_bits = (lane < tail ? 1u : 0u) | (lane + 32 < tail ? 2u : 0u);
            };
            {
                const int G = J ? s_gcum[J - 1] : 0;
                const int per = (G + PS_NC - 1) / PS_NC;
                f0 = wid * per;
                if (f0 > G) f0 = G;
                f_end = f0 + per < G ? f0 + per : G;
                cur_f = f0;
                if (f_end > f0) {
                    const int sa_ = seg_of(f0), sb_ = seg_of(f_end - 1);
                    nblk = (f_end - f0) + (sb_ - sa_ + 1);
                    seg = sa_;
                    seg_g0 = sa_ ? s_gcum[sa_ - 1] : 0;
                    seg_gend = s_gcum[sa_];
                }
            }
            uint32_t dsc[ST_R];
#pragma unroll
            for (int s = 0; s < ST_R; ++s) dsc[s] = 0u;
            PS_ISSUE(0, dsc[0], 0 < nblk)
            PS_ISSUE(1, dsc[1], 1 < nblk)

            // ---- top-k state ----
            // K1: the warp's best so far, warp-uniform: distance bits, position, id (or -1: not looked up yet), segment
            uint32_t b_thr = 0xffffffffu, b_pos = 0u;
            int b_id = -1, b_seg = -1;
            bool b_have = false;
            WarpTopk wt;
            u64 *cta_thr = &kb.cta_thr;
            uint32_t thr_hi = 0xffffffffu;
            if constexpr (!K1) {
                wt.keys = kb.keys + wid * PS_CAPW;
                wt.cap = PS_CAPW;
                wt.k = a.k;
                wt.count = 0;
                wt.thr_w = kb.thr_w;
                wt.nw = PS_NC;
                wt.wid = wid;
                wt.ids = a.ids;  // lazy ids: pushes carry positions, warp_compact looks the ids up
                wt.s_off = s_off;
                wt.s_gcum = s_gcum;
                wt.J = J;
                wt.nres = 0;
            }
            auto emit2 = [&](float dx, float dy, uint32_t d) {
                const uint32_t ux = __float_as_uint(dx), uy = __float_as_uint(dy);
                if constexpr (K1) {
                    const bool px = (d & 1u) && ux <= b_thr, py = (d & 2u) && uy <= b_thr;
                    if (__any_sync(0xffffffffu, px || py)) {
                        // some row of this group is at or below the warp's best distance (rare after the first groups)
                        const uint32_t mine = umin(px ? ux : 0xffffffffu, py ? uy : 0xffffffffu);
                        const uint32_t mn = __reduce_min_sync(0xffffffffu, mine);
                        // the lowest position at that distance: x rows (row = lane) lie before y rows (row = 32 + lane)
                        const unsigned bx = __ballot_sync(0xffffffffu, px && ux == mn), by = __ballot_sync(0xffffffffu, py && uy == mn);
                        const uint32_t pos = ((d >> 2) << 6) | (bx ? (uint32_t)(__ffs(bx) - 1) : 32u + (uint32_t)(__ffs(by) - 1));
                        if (!b_have || mn < b_thr) {
                            b_thr = mn; b_pos = pos; b_id = -1; b_seg = seg_of((int)(d >> 2)); b_have = true;
                        } else {  // an exact tie with the best so far (which lies at a lower position)
                            const int sn = seg_of((int)(d >> 2));
                            if (sn != b_seg) {  // another posting list: positions do not order ids across lists (src/rii.h:312: we rank by (dist, id))
                                const int idn = (int)id_of(pos);
                                if (b_id < 0) b_id = (int)id_of(b_pos);
                                if (idn < b_id) { b_pos = pos; b_id = idn; b_seg = sn; }
                            }
                        }
                    }
                } else {
                    const int f = (int)(d >> 2);
                    thr_hi = reinterpret_cast<volatile uint32_t *>(cta_thr)[1];
                    const bool px = (d & 1u) && ux <= thr_hi;
                    const bool py = (d & 2u) && uy <= thr_hi;
                    if (__any_sync(0xffffffffu, px || py)) {
                        warp_push(wt, cta_thr, lane, dx, px ? (uint32_t)(f * 64 + lane) : 0u, px);
                        warp_push(wt, cta_thr, lane, dy, py ? (uint32_t)(f * 64 + 32 + lane) : 0u, py);
                    }
                }
            };
            if (sg.plain) {  // a table with huge / inf / NaN entries: exact per-candidate sums (scan_stream.cuh plain_slice)
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                for (int f = f0; f < f_end; ++f) {
                    const int j = seg_of(f);
                    const int g = f - (j ? s_gcum[j - 1] : 0);
                    float dd[2] = {0.f, 0.f};
                    uint32_t dsc_ = (uint32_t)f << 2;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int s = half * 32 + lane;
                        if (g * 64 + s < s_take[j]) {
                            dsc_ |= 1u << half;
                            const uint8_t *pp = pc + (size_t)s_prow[j] * 32 + (size_t)g * ST_BLOCK_BYTES + half * 1024 + lane * 16;
                            float d = 0.f;
                            for (int m = 0; m < 32; ++m) {
                                const int x = lane + m;
                                const uint32_t ks = __ldg(pp + (x >> 5) * ST_BLOCK_BYTES + ((x >> 4) & 1) * 512 + (x & 15));
                                const float v = lut2[ks * 64 + ((m + 32) & 63)];
                                d = m ? __fadd_rn(d, v) : v;
                            }
                            dd[half] = d;
                        }
                    }
                    emit2(dd[0], dd[1], dsc_);
                }
            } else {
                unsigned long long acc2 = 0ull, out2 = 0ull;
#pragma unroll 1
                for (int m = 0; m < nblk; m += ST_R) {
                    PS_STAGE(0)
                    PS_STAGE(1)
                    PS_STAGE(2)
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            if constexpr (K1) {
                if (b_have && b_id < 0) b_id = (int)id_of(b_pos);
                if (lane == 0) {
                    kb.keys[wid * PS_CAPW] = ((u64)b_thr << 32) | (u64)(uint32_t)b_id;
                    kb.cnt[wid] = b_have ? 1 : 0;
                }
            } else {
                warp_compact(wt, cta_thr, lane);
                if (lane == 0) kb.cnt[wid] = wt.count;
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) ps_mbar_arrive(bar_done + 8 * p);
            if (timed) c_scan += clock64() - t0_;
        }
        if (dbg && lane == 0 && (wid == 0 || wid == PS_NC - 1)) { dbg[wid == 0 ? 0 : 2] = c_wait; dbg[wid == 0 ? 1 : 3] = c_scan; }
    } else {
        // =========================================== producer ============================================
        int *s_f = reinterpret_cast<int *>(smem_raw + PS_OFF_PLAN), *s_pre = s_f + PS_WMAX, *s_loc = s_pre + PS_WMAX;
        int *hist = reinterpret_cast<int *>(smem_raw + PS_OFF_HIST);
        uint32_t *pool_d = reinterpret_cast<uint32_t *>(smem_raw + PS_OFF_POOL);
        u64 *selk = reinterpret_cast<u64 *>(smem_raw + PS_OFF_SELK);
        const uint32_t ring = smem_base + PS_OFF_PRING + lane * 16;
        const bool fused = a.coarse_mode == 0;
        long long p_wait = 0, p_merge = 0, p_table = 0, p_coarse = 0, p_select = 0, p_plan = 0;

        // final merge of the consumers' results of query qi -> output
        auto merge = [&](int qi) {
            const int p = qi & 1, b = (int)blockIdx.x + qi * (int)gridDim.x;
            const PsKeys &kb = keyb[p];
            if constexpr (K1) {  // <= one key per consumer warp: a warp minimum under (distance, id)
                u64 key = lane < PS_NC && kb.cnt[lane] ? kb.keys[lane * PS_CAPW] : RII_KEY_MAX;
                const bool any = __any_sync(0xffffffffu, key != RII_KEY_MAX);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const u64 y = __shfl_xor_sync(0xffffffffu, key, o);
                    key = y < key ? y : key;
                }
                if (lane == 0) {
                    if (any) {
                        a.out.out_ids[b] = a.out.id_base + (long long)key_id(key);
                        a.out.out_dists[b] = key_dist(key);
                    }
                    a.out.out_counts[b] = any ? 1 : 0;
                }
            } else {  // the consumers' sorted lists (<= PS_NC * PS_MAXK keys)
                int o = 0;
                for (int w2 = 0; w2 < PS_NC; ++w2) {
                    const int c = kb.cnt[w2];
                    for (int i = lane; i < c; i += 32) selk[o + i] = kb.keys[w2 * PS_CAPW + i];
                    o += c;
                }
                warp_sort_any(selk, o, lane);
                const int n = o < a.k ? o : a.k;
                for (int i = lane; i < n; i += 32) {
                    a.out.out_ids[(size_t)b * a.k + i] = a.out.id_base + (long long)key_id(selk[i]);
                    a.out.out_dists[(size_t)b * a.k + i] = key_dist(selk[i]);
                }
                if (lane == 0) a.out.out_counts[b] = n;
            }
            __syncwarp();
        };
        // table rows >= Ks are never written by the fast table build: zero them once in both tables (zero-padded code bytes and
        // invalid codes must look up finite values)
        for (int e = a.Ks * 64 + lane; e < 2 * 256 * 64; e += 32) {
            const int t = e >> 14, r = e & 16383;
            if (r >= a.Ks * 64) reinterpret_cast<float *>(smem_raw + (PS_T0 - 0x400u) + t * SK_LUT_BYTES)[r] = 0.f;
        }
        __syncwarp();
#pragma unroll 1
        for (int qi = 0; qi < n_my + 2; ++qi) {
            long long tp_ = 0;
            if (qi >= 2) {  // the consumers' results of query qi - 2 (same parity) are final: merge them, freeing the buffers
                p_wait += ps_mbar_wait(bar_done + 8 * (qi & 1), (uint32_t)((qi - 2) >> 1) & 1u, timed, 400u);
                if (timed) tp_ = clock64();
                merge(qi - 2);
                if (timed) p_merge += clock64() - tp_;
            }
            if (qi >= n_my) continue;
            if (timed) tp_ = clock64();
            // ------------------------------------ prepare query qi ------------------------------------
            const int p = qi & 1, b = (int)blockIdx.x + qi * (int)gridDim.x;
            PsSeg &sg = segs[p];
            PsKeys &kb = keyb[p];
            float *lut2 = reinterpret_cast<float *>(smem_raw + (PS_T0 - 0x400u) + p * SK_LUT_BYTES);
            // ---- K1 (src/rii.h:361-373): lane = sub-space, every ks; column c of the table holds sub-space c mod 32 ----
            int bad = qi == 0 && first_coop ? bad0 : 0;
            if (!(qi == 0 && first_coop)) {
                const int m = lane;
                const float *qm = a.Q + (size_t)b * Mr * a.Ds + (size_t)m * a.Ds;
                if (m >= Mr) {
                    for (int ks = 0; ks < 256; ++ks) {
                        lut2[ks * 64 + m + 32] = 0.f;
                        lut2[ks * 64 + m] = 0.f;
                    }
                } else if (a.Ds <= 4) {  // the codeword copy is padded to 4 floats per sub-vector (zeros add +0: same sum)
                    float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    q4.x = __ldg(qm);
                    if (a.Ds > 1) q4.y = __ldg(qm + 1);
                    if (a.Ds > 2) q4.z = __ldg(qm + 2);
                    if (a.Ds > 3) q4.w = __ldg(qm + 3);
                    const float4 *cw4 = reinterpret_cast<const float4 *>(a.cw_t) + m;
                    // rows [0, Ks) in steps of 8; the 8 codeword rows of the next step are in flight (L2) while a step is computed.
                    // Rows >= Ks were zeroed when the kernel started and are never written.
                    const float2 qa = make_float2(q4.x, q4.y), qb = make_float2(q4.z, q4.w);
                    uint32_t vmax = 0u;  // distances are >= +0: unsigned order of the bits == order of the values, NaNs above everything
                    float4 ca[8], cb[8];
                    auto load8 = [&](float4 (&c)[8], int k0) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) c[u] = __ldg(cw4 + (k0 + u < 256 ? k0 + u : 255) * Mr);
                    };
                    auto comp8 = [&](const float4 (&c)[8], int k0) {
                        if (k0 + 8 <= a.Ks) {
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const float v = sqdist4(qa, qb, c[u]);
                                vmax = umax(vmax, __float_as_uint(v));
                                lut2[(k0 + u) * 64 + m + 32] = v;
                                lut2[(k0 + u) * 64 + m] = v;
                            }
                        } else {
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                if (k0 + u < a.Ks) {
                                    const float v = sqdist4(qa, qb, c[u]);
                                    vmax = umax(vmax, __float_as_uint(v));
                                    lut2[(k0 + u) * 64 + m + 32] = v;
                                    lut2[(k0 + u) * 64 + m] = v;
                                }
                            }
                        }
                    };
                    load8(ca, 0);
#pragma unroll 1
                    for (int k0 = 0; k0 < a.Ks; k0 += 16) {
                        load8(cb, k0 + 8);
                        comp8(ca, k0);
                        load8(ca, k0 + 16);
                        if (k0 + 8 < a.Ks) comp8(cb, k0 + 8);
                    }
                    bad |= vmax > __float_as_uint(ST_TABLE_LIMIT);
                } else {
#pragma unroll 1
                    for (int ks = 0; ks < 256; ++ks) {
                        float v = 0.f;
                        if (ks < a.Ks) v = l2sqr_lanes(qm, a.cw_t + ((size_t)ks * Mr + m) * a.Ds, a.Ds, a.variant);
                        bad |= !(v <= ST_TABLE_LIMIT);
                        lut2[ks * 64 + m + 32] = v;
                        lut2[ks * 64 + m] = v;
                    }
                }
            }
            const bool plain = __any_sync(0xffffffffu, bad) != 0;  // (also orders the table writes before the lookups below)
            __syncwarp();
            if (timed) { const long long t_ = clock64(); p_table += t_ - tp_; tp_ = t_; }
            int *ranked_g = a.plan.ranked + (size_t)b * a.w_eff;
            if (fused) {
                // ---- K4 coarse pass (src/rii.h:259-265): this warp scans the skew64 centers with the new table ----
                uint32_t d_lo = 0xffffffffu, d_hi = 0u;
                const int Gc = (a.nlist + 63) >> 6;
                if (plain) {
                    for (int i = lane; i < a.nlist; i += 32) {
                        const uint8_t *pp = a.centers + (size_t)(i >> 6) * ST_BLOCK_BYTES + ((i >> 5) & 1) * 1024 + (i & 31) * 16;
                        float d = 0.f;
                        for (int m = 0; m < 32; ++m) {
                            const int x = (i & 31) + m;
                            const uint32_t ks = __ldg(pp + (x >> 5) * ST_BLOCK_BYTES + ((x >> 4) & 1) * 512 + (x & 15));
                            const float v = lut2[ks * 64 + ((m + 32) & 63)];
                            d = m ? __fadd_rn(d, v) : v;
                        }
                        const uint32_t u = __float_as_uint(d);
                        pool_d[i] = u;
                        d_lo = u < d_lo ? u : d_lo;
                        d_hi = u > d_hi ? u : d_hi;
                    }
                } else {
                    const uint32_t colreg = (PS_T0 + (uint32_t)p * SK_LUT_BYTES) | (uint32_t)((32 - lane) * 4);
                    // one run: the Gc groups of the centers + the drain block
                    int cur_f = 0, run_left = Gc + 1;
                    const int nblk = Gc + 1;
                    const int tail = a.nlist - (Gc - 1) * 64;
                    const uint32_t last_bits = (lane < tail ? 1u : 0u) | (lane + 32 < tail ? 2u : 0u);
                    const uint8_t *bp = a.centers + lane * 16;
                    auto next_run = [&]() {};
                    uint32_t dsc[ST_R];
#pragma unroll
                    for (int s = 0; s < ST_R; ++s) dsc[s] = 0u;
                    auto emit2 = [&](float dx, float dy, uint32_t d) {
                        const int f = (int)(d >> 2);
                        const uint32_t ux = __float_as_uint(dx), uy = __float_as_uint(dy);
                        if (d & 1u) { pool_d[f * 64 + lane] = ux; d_lo = ux < d_lo ? ux : d_lo; d_hi = ux > d_hi ? ux : d_hi; }
                        if (d & 2u) { pool_d[f * 64 + 32 + lane] = uy; d_lo = uy < d_lo ? uy : d_lo; d_hi = uy > d_hi ? uy : d_hi; }
                    };
                    PS_ISSUE(0, dsc[0], 0 < nblk)
                    PS_ISSUE(1, dsc[1], 1 < nblk)
                    unsigned long long acc2 = 0ull, out2 = 0ull;
#pragma unroll 1
                    for (int m = 0; m < nblk; m += ST_R) {
                        PS_STAGE(0)
                        PS_STAGE(1)
                        PS_STAGE(2)
                    }
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
                d_lo = __reduce_min_sync(0xffffffffu, d_lo);
                d_hi = __reduce_max_sync(0xffffffffu, d_hi);
                __syncwarp();
                if (timed) { const long long t_ = clock64(); p_coarse += t_ - tp_; tp_ = t_; }
                // ---- selection of the w nearest lists under (distance, list id): src/rii.h:279-280 ----
                int np = warp_select_smallest(pool_d, a.nlist, a.w_eff, selk, hist, d_lo, d_hi, lane);
                if (np < 0) {  // > 256 exact ties at the w-th distance: sort every (dist, index) key (selk + the idle ring: 1024 keys)
                    const int P = next_pow2(a.nlist);
                    for (int i = lane; i < P; i += 32) selk[i] = i < a.nlist ? (((u64)pool_d[i] << 32) | (u64)(uint32_t)i) : RII_KEY_MAX;
                    warp_sort_smem(selk, P, lane);
                }
                __syncwarp();
                for (int j = lane; j < a.w_eff; j += 32) ranked_g[j] = (int)key_id(selk[j]);
                __syncwarp();
                if (timed) { const long long t_ = clock64(); p_select += t_ - tp_; tp_ = t_; }
            }
            // ---- the plan from the ranking (own, or given) ----
            for (int j = lane; j < a.w_eff; j += 32) {
                const int no = fused ? (int)key_id(selk[j]) : ranked_g[j];
                s_f[j] = a.plan.glob_len[no];
                s_pre[j] = a.plan.pre_len ? a.plan.pre_len[no] : 0;
                s_loc[j] = a.plan.loc_len[no];
                sg.off[j] = a.offsets[no];
                sg.prow[j] = a.skew_off[no];
            }
            __syncwarp();
            const int jc = plan_warp(a.plan, b, lane, s_f, s_pre, s_loc, sg.off, sg.prow, sg.gcum, sg.take);
            if (lane == 0) {
                sg.J = jc;
                sg.b = b;
                sg.plain = plain ? 1 : 0;
                kb.cta_thr = RII_KEY_MAX;
            }
            if (lane < PS_NC) kb.thr_w[lane] = RII_KEY_MAX;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) ps_mbar_arrive(bar_full + 8 * p);
            if (timed) p_plan += clock64() - tp_;
        }
        if (dbg && lane == 0) { dbg[4] = p_wait; dbg[5] = p_merge; dbg[6] = p_table; dbg[7] = p_coarse; dbg[8] = p_select; dbg[9] = p_plan; dbg[10] = n_my; }
    }
}
