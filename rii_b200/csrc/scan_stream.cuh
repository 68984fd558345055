// ===================================================================================================
// K2/K5 v4 ("stream"): the skewed scan engine over a PRE-SKEWED copy of the codes.
//
// What v2 / v3 taught (profiles/r01_*): the lookup loop is bound by the shared-memory pipe, not by issue slots.  Every
// code byte went global -> shared (LDGSTS into per-lane padded regions) -> register (one LDS.32 + funnel shift per
// 4 bytes) just so that lane l could read its stream l bytes late, with warp barriers around every tile.
//
// Here the lag is applied ONCE, when the index is built: the "skew64" layout stores, for every 64 consecutive rows of
// a segment (the whole table for the linear scan, one posting list for the IVF scan), 64 byte streams -- stream s =
// rows s, 64 + s, 128 + s, ... back to back -- each delayed by (s & 31) bytes and cut into 32-byte windows:
//        window[b][s][i] = stream_s[32 b - (s & 31) + i]          (zeros before the first / after the last row)
// Same size as the codes (+ one block of 64 rows per segment).  Lane l of a warp owns streams l (x) and 32 + l (y); a
// 2 KB block b holds the 64 windows as four 512-byte quarters [x bytes 0-15 | x 16-31 | y 0-15 | y 16-31] x 32 lanes,
// so a lane copies exactly the four 16-byte chunks it will consume itself (cp.async, 512 contiguous bytes per warp
// instruction, conflict free) into a private 4-stage ring and reads them back with four LDS.128: no cross-lane
// traffic, hence NO warp barrier and no funnel shifts -- the lookup address is one PRMT of a register.  Per pair of
// lookups:       2 PRMT + 2 LDS + 2 FFMA2            (+ 4 LDGSTS + 4 LDS.128 per 64 lookups)
// with the predicate-free accumulation of scan_dual.cuh (acc = acc * keep_t + v, out = acc * sel_t + out).  Distances
// are still the sequential fp32 sums of src/rii.h:386-394, bit for bit.  Three blocks (6 KB per warp, 72 KB per SM)
// are in flight while one is scanned.  Feeding variants measured in isolation (tools/ubench_feed.cu,
// profiles/r01_ubench_feed.jsonl; lookups/clk/SM, HBM source): this ring 20.9, LDG.256 into registers 19.3 (and the
// real kernel lost much more to scoreboard sharing between its 8 outstanding LDG.256 and the table LDS), TMA bulk 19.7.
//
// Work is split by GROUPS of 64 rows.  A warp walks a contiguous range of the pass's flattened group list; at the end
// of a segment (or of its range) one extra "drain" block finishes the lagging rows.  Rows past a segment's take
// count are masked when their distance completes.
// ===================================================================================================
#pragma once
#include "common.cuh"
#include "warp_topk.cuh"

#define ST_BLOCK_BYTES 2048         // 64 windows of 32 bytes
// Launch shapes (template parameters NW warps, R ring stages = R - 1 blocks in flight, MINB CTAs per SM, TB = absolute
// shared address of the table):
//   one CTA per SM:  NW = 12, R = 4, TB = 0x10000 -- long scans (linear), large topk / many lists
//   two CTAs per SM: NW = 6,  R = 3, TB = 0x3000  -- per-query IVF batches: the serial phases of one query (table build,
//                    coarse selection, plan, final merge) overlap the scan of the other CTA's query
//   M = 64:          NW = 8,  R = 4, TB = 0x6000, two tables (128 KB), one CTA per SM; NW = 10, TB = 0x2400 when the keys fit
#define ST_TB1 0x10000u
#define ST_TB2 0x3000u
#define ST_TB3 0x6000u
#define ST_TB4 0x2400u      // rows of 64 bytes, 10 warps: 8 KB of keys / segments below the tables
#define ST_TABLE_LIMIT 5e36f        // 64 table entries below this cannot overflow fp32 (acc * 0 needs finite acc)

#include "skew64.cuh"

// The candidate plan of make_plan (kernels.cuh; SURVEY Appendix A.3, src/rii.h:286-322) computed by ONE WARP with prefix
// scans instead of a serial walk, fused with the compaction into the stream engine's segment list.  Inputs by rank j <
// w_eff in shared memory: s_f (global list length), s_pre (part held by lower ranks), s_loc (part held locally), s_off
// (CSR offset), s_prow (first skew64 row).  Outputs: the non-empty segments, compacted in place -- s_gcum (inclusive
// prefix of 64-row groups), s_take (local take), s_off, s_prow -- and J / take_last / flags of the query in global
// memory exactly as make_plan writes them.  Returns the number of segments (0 when the plan is flagged).
__device__ __forceinline__ int plan_warp(const PlanArgs &p, int b, int lane, const int *s_f, const int *s_pre, const int *s_loc,
                                         long long *s_off, long long *s_prow, int *s_gcum, int *s_take, bool publish = true)
{
    const int W = p.w_eff;
    // pass 1: F_j = inclusive prefix of the list lengths; jL = first rank with F_j >= L; F at rank w - 1
    long long carry = 0, F_before = 0, F_w = -1;
    int jL = W;
    for (int base = 0; base < W; base += 32) {
        const int j = base + lane;
        const long long f = j < W ? (long long)s_f[j] : 0;
        long long x = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        x += carry;
        const unsigned hit = __ballot_sync(0xffffffffu, j < W && x >= p.L);
        if (hit && jL == W) {
            const int src = __ffs(hit) - 1;
            jL = base + src;
            F_before = __shfl_sync(0xffffffffu, x - f, src);
        }
        if (p.w - 1 >= base && p.w - 1 < base + 32 && p.w - 1 < W) F_w = __shfl_sync(0xffffffffu, x, p.w - 1 - base);
        carry = __shfl_sync(0xffffffffu, x, 31);
    }
    // where the walk stops: at L (src/rii.h:302-304) or after the w-th list with >= topk candidates (src/rii.h:309)
    int jstop = -1;
    bool by_L = false;
    if (jL < W && jL <= p.w - 1) { jstop = jL; by_L = true; }
    else if (p.w - 1 < W && F_w >= p.topk) jstop = p.w - 1;
    else if (jL < W) { jstop = jL; by_L = true; }
    const int J = jstop >= 0 ? jstop + 1 : W;
    const int flag = jstop >= 0 ? 0 : (W >= p.nlist ? 2 : 1);  // 2: empty result (src/rii.h:325); 1: walk beyond w (host re-runs)
    // pass 2: local takes, compaction of the non-empty segments, group prefix
    int cnt = 0, gcarry = 0, take_last = 0;
    for (int base = 0; base < J; base += 32) {
        const int j = base + lane;
        long long take = 0;
        int lt = 0;
        long long off = 0, prow = 0;
        if (j < J) {
            take = (by_L && j == jstop) ? p.L - F_before : (long long)s_f[j];
            long long l2 = take - s_pre[j];
            l2 = l2 < 0 ? 0 : l2;
            l2 = l2 > s_loc[j] ? s_loc[j] : l2;
            lt = (int)l2;
            off = s_off[j];
            prow = s_prow[j];
        }
        if (J - 1 >= base && J - 1 < base + 32) take_last = (int)__shfl_sync(0xffffffffu, take, J - 1 - base);
        const bool nz = flag == 0 && lt > 0;
        const unsigned bal = __ballot_sync(0xffffffffu, nz);
        int g = nz ? (lt + 63) >> 6 : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, g, o);
            if (lane >= o) g += y;
        }
        g += gcarry;
        __syncwarp();  // every lane has read its rank's inputs: compacted slots (<= j) may be overwritten
        if (nz) {
            const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
            s_gcum[pos] = g;
            s_take[pos] = lt;
            s_off[pos] = off;
            s_prow[pos] = prow;
        }
        cnt += __popc(bal);
        gcarry = __shfl_sync(0xffffffffu, g, 31);
        __syncwarp();
    }
    if (lane == 0 && publish) {
        p.J[b] = J;
        p.take_last[b] = take_last;
        p.flags[b] = flag;
    }
    return cnt;
}

#define ST_LDS128(W, A)                                                                                       \
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"((W)[0]), "=r"((W)[1]), "=r"((W)[2]), "=r"((W)[3]) : "r"(A))
// issue the copy of the next block of this warp's walk into ring stage S and describe it: D = flattened group << 2 | y row
// valid << 1 | x row valid for the first block of a group, 0 for its other half (H = 2) and for drain blocks.  The walk is a
// list of RUNS -- the consecutive blocks of one segment inside the warp's slice plus the H drain blocks behind them (the next
// blocks in memory): inside a run the logic is a pointer increment and a counter, segment tables are read by next_run() only.
// Warp-uniform state: run_left (blocks left in the run, drain blocks included), bp (this lane's 16-byte column of the next
// block), cur_f, hb, last_bits (valid bits of the run's last group).  Past the end of the walk only an (empty) group is
// committed, so that "wait_group ST_D" always means "the block of this stage has landed".
#define ST_ISSUE(S, D, ACTIVE)                                                                                \
    {                                                                                                         \
        if (ACTIVE) {                                                                                         \
            if (run_left == 0) next_run();                                                                    \
            if (run_left <= H) D = 0u; /* drain */                                                            \
            else if (H == 2 && hb) { D = 0u; hb = 0; ++cur_f; }                                               \
            else {                                                                                            \
                D = ((uint32_t)cur_f << 2) | (run_left <= 2 * H ? last_bits : 3u);                            \
                if (H == 1) ++cur_f; else hb = 1;                                                             \
            }                                                                                                 \
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
// one lookup step of both streams: PRMT builds ks << 8 | column byte offset, the load's immediate adds the (compile-time)
// absolute shared address of the block's table and the step.  In a block that holds the first 32 bytes of the rows
// (h_ == 0) the sums restart / are captured at the lane's row boundary (acc / out updates of DU_STEP, scan_dual.cuh); in
// the other blocks of a row (M = 64) the lookups are simply added.
#define ST_STEP(WX, WY, BYTE, T)                                                                              \
    {                                                                                                         \
        const uint32_t ax_ = __byte_perm(WX, colreg, 0x7604 | ((BYTE) << 4));                                 \
        const uint32_t ay_ = __byte_perm(WY, colreg, 0x7604 | ((BYTE) << 4));                                 \
        float vx_, vy_;                                                                                       \
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(vx_) : "r"(ax_), "n"(TB + h_ * SK_LUT_BYTES + 4 * (T))); \
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(vy_) : "r"(ay_), "n"(TB + h_ * SK_LUT_BYTES + 4 * (T))); \
        if constexpr (h_ == 0)                                                                                \
            asm("{.reg .b64 v, kk, ss; mov.b64 v, {%4, %5}; mov.b64 kk, {%2, %2}; mov.b64 ss, {%3, %3};"     \
                " fma.rn.f32x2 %1, %0, ss, %1; fma.rn.f32x2 %0, %0, kk, v;}"                                 \
                : "+l"(acc2), "+l"(out2)                                                                      \
                : "f"(keep[T]), "f"(sel[T]), "f"(vx_), "f"(vy_));                                             \
        else                                                                                                  \
            asm("{.reg .b64 v; mov.b64 v, {%1, %2}; add.rn.f32x2 %0, %0, v;}" : "+l"(acc2) : "f"(vx_), "f"(vy_)); \
    }
#define ST_WORD(BX, BY, Q)                                                                                    \
    ST_STEP(BX[Q], BY[Q], 0, 4 * (Q) + 0)                                                                     \
    ST_STEP(BX[Q], BY[Q], 1, 4 * (Q) + 1)                                                                     \
    ST_STEP(BX[Q], BY[Q], 2, 4 * (Q) + 2)                                                                     \
    ST_STEP(BX[Q], BY[Q], 3, 4 * (Q) + 3)
#define ST_BLOCK(BX, BY)                                                                                      \
    {                                                                                                         \
        ST_WORD(BX, BY, 0) ST_WORD(BX, BY, 1) ST_WORD(BX, BY, 2) ST_WORD(BX, BY, 3)                           \
        ST_WORD(BX, BY, 4) ST_WORD(BX, BY, 5) ST_WORD(BX, BY, 6) ST_WORD(BX, BY, 7)                           \
    }
// one pipeline stage: block m + S lives in ring stage S (its half-row index is S % H: every walk is a whole number of
// H-block units); the stage of block m + S - 1 takes block m + S + ST_D.  After a block with h_ == 0, out2 holds the
// finished distances of the PREVIOUS group (descriptor d_last0).
#define ST_STAGE(S)                                                                                           \
    if (m + (S) < nblk) {                                                                                     \
        constexpr int h_ = (S) % H;                                                                           \
        /* descriptor of the group that completes in this block: H = 1: the previous block's (its slot is refilled */ \
        /* below); H = 2: kept in d_last0 */                                                                  \
        const uint32_t dsel_ = H == 1 ? dsc[((S) + ST_D) % ST_R] : dsc[S];                                    \
        ST_ISSUE(((S) + ST_D) % ST_R, dsc[((S) + ST_D) % ST_R], m + (S) + ST_D < nblk)                        \
        if constexpr (ST_D == 3) asm volatile("cp.async.wait_group 3;" ::: "memory");                         \
        else if constexpr (ST_D == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");                    \
        else asm volatile("cp.async.wait_group 1;" ::: "memory");                                             \
        uint32_t wx_[8], wy_[8];                                                                              \
        ST_LDS128(wx_, ring + (S) * ST_BLOCK_BYTES);                                                          \
        ST_LDS128(wx_ + 4, ring + (S) * ST_BLOCK_BYTES + 512);                                                \
        ST_LDS128(wy_, ring + (S) * ST_BLOCK_BYTES + 1024);                                                   \
        ST_LDS128(wy_ + 4, ring + (S) * ST_BLOCK_BYTES + 1536);                                               \
        ST_BLOCK(wx_, wy_)                                                                                    \
        if constexpr (h_ == 0) {                                                                              \
            float dx_, dy_;                                                                                   \
            asm("mov.b64 {%0, %1}, %2;" : "=f"(dx_), "=f"(dy_) : "l"(out2));                                  \
            out2 = 0ull;                                                                                      \
            if constexpr (H == 1) {                                                                           \
                emit2(dx_, dy_, dsel_);                                                                       \
            } else {                                                                                          \
                emit2(dx_, dy_, d_last0);                                                                     \
                d_last0 = dsel_;                                                                              \
            }                                                                                                 \
        }                                                                                                     \
    }

// Args: SkewArgs (kernels.cuh) with `codes` = skew64 table of the pass-1 rows (linear: the codes by id, one segment;
// IVF: every local posting list, segment i at physical row skew_off[i]) and `centers` = skew64 of the coarse centers.
// K1: the topk == 1 instantiation -- in the final pass a warp keeps its best (distance, position) in warp-uniform registers
// instead of pushing keys into its shared-memory list (no pushes, no compaction, ids looked up for the winner and for exact
// cross-list ties only; same scheme as k_scan_persist32<true>).
template <int NW, bool IVF, int ST_R, int MINB, uint32_t TB, int H, bool K1>
__global__ void __launch_bounds__(NW * 32, MINB) k_scan_stream32(SkewArgs a)
{
    static_assert(H == 1 || H == 2, "M = 32 or 64");
    static_assert(ST_R % H == 0, "a block's half-row index must be a constant of its pipeline stage");
    constexpr int M = 32 * H;                                // bytes per (zero-padded) code row; H tables of 64 KB
    const int Mr = a.M;                                      // real sub-spaces (<= M): columns beyond look up +0.0f, and x + 0 == x
    constexpr int ST_D = ST_R - 1;                           // blocks in flight ahead of the one being scanned
    constexpr uint32_t ST_RING_BYTES = ST_R * ST_BLOCK_BYTES;  // per warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout (dynamic shared memory; the window starts at absolute shared address ~1 KB):
    //   [NW key buffers][cta_thr][thr_w][segments: s_off, s_prow i64[wq] | s_gcum, s_take, s_cum, s_f, s_pre, s_loc i32[wq] | s_plan]
    //   [fused IVF: nlist coarse distances]
    //   ... lut2 (64 KB) at ABSOLUTE shared address TB ... [NW rings of ST_R x 2 KB; between the passes: selection scratch]
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t lut_off = TB - smem_base;
    float *lut2 = reinterpret_cast<float *>(smem_raw + lut_off);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int capw = a.cap;
    long long *dbg = a.dbg ? a.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();
    u64 *wkeys = reinterpret_cast<u64 *>(smem_raw) + (size_t)wid * capw;
    u64 *cta_thr = reinterpret_cast<u64 *>(smem_raw) + (size_t)NW * capw;
    u64 *thr_w = cta_thr + 1;  // [NW]
    const int wq = IVF ? a.w_eff : 1;
    long long *s_off = reinterpret_cast<long long *>(smem_raw + (size_t)NW * capw * 8 + 8 + NW * 8);
    long long *s_prow = s_off + wq;
    int *s_gcum = reinterpret_cast<int *>(s_prow + wq);
    int *s_take = s_gcum + wq, *s_cum = s_take + wq, *s_f = s_cum + wq, *s_pre = s_f + wq, *s_loc = s_pre + wq, *s_plan = s_loc + wq;
    const uint32_t hi0 = lut_off + H * SK_LUT_BYTES;                  // rings / scratch above the table(s)
    // coarse distances of the fused coarse pass: nlist words after the segment tables
    const size_t meta_end = ((size_t)NW * capw * 8 + 8 + NW * 8 + (size_t)wq * 40 + 16 + 15) & ~(size_t)15;
    uint32_t *pool_d = reinterpret_cast<uint32_t *>(smem_raw + meta_end);
    // ... followed by the selection histogram: 256 bins + [min, max, count, -] (cta_select_smallest)
    const bool use_pool = IVF && a.centers && !a.coarse_lists;
    int *hist = reinterpret_cast<int *>(pool_d + ((use_pool ? a.nlist : 0) + 3) / 4 * 4);
    if (meta_end + (use_pool ? (size_t)(a.nlist + 3) / 4 * 16 + 1040 : 0) > lut_off || hi0 + (size_t)NW * ST_RING_BYTES > a.smem_bytes)
        __trap();  // host sized the launch wrongly
    const uint32_t ring = smem_base + hi0 + wid * ST_RING_BYTES + lane * 16;  // this lane's chunk column of the warp's ring
    const int b = blockIdx.y;
    const bool fused = IVF && a.centers != nullptr;
    int J = 0;  // segments of the current pass (after dropping empty ones)

    if constexpr (IVF) {
        if (!fused && a.coarse_mode == 2) {
            // the ranking comes from a separate coarse-only launch (possibly of another GPU: the coarse phase of a sharded
            // batch is split over the ranks); the plan is made here, identically by every CTA of the query
            if (wid == 0) {
                const int *ranked_g = a.plan.ranked + (size_t)b * a.w_eff;
                for (int j = lane; j < a.w_eff; j += 32) {
                    const int no = ranked_g[j];
                    s_f[j] = a.plan.glob_len[no];
                    s_pre[j] = a.plan.pre_len ? a.plan.pre_len[no] : 0;
                    s_loc[j] = a.plan.loc_len[no];
                    s_off[j] = a.offsets[no];
                    s_prow[j] = a.skew_off[no];
                }
                __syncwarp();
                const int jc = plan_warp(a.plan, b, lane, s_f, s_pre, s_loc, s_off, s_prow, s_gcum, s_take, blockIdx.x == 0);
                if (lane == 0) s_plan[0] = jc;
            }
        } else if (use_pool) {
            for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        }
        if (fused && threadIdx.x == 0) {  // coarse pass: one segment, the skew64 copy of the centers
            if (use_pool) {
                hist[256] = -1;  // min (as unsigned)
                hist[257] = 0;   // max
            }
            s_gcum[0] = (a.nlist + 63) >> 6;
            s_take[0] = a.nlist;
            s_off[0] = 0;
            s_prow[0] = 0;
            s_plan[0] = 1;
        }
    } else if (threadIdx.x == 0) {
        s_gcum[0] = (int)((a.N + 63) >> 6);
        s_take[0] = (int)a.N;
        s_off[0] = 0;
        s_prow[0] = 0;
        s_plan[0] = a.N > 0 ? 1 : 0;
    }
    if (threadIdx.x == 0) *cta_thr = RII_KEY_MAX;
    if (threadIdx.x < NW) thr_w[threadIdx.x] = RII_KEY_MAX;
    __syncthreads();
    J = s_plan[0];

    const uint32_t colreg = (uint32_t)((32 - lane) * 4);  // column byte offset of the lane (the table base is in the load's immediate)

    // per-pass state (warp-uniform)
    const uint8_t *pc = fused ? a.centers : a.codes;  // skew64 table of the pass
    int f0 = 0, f_end = 0, cur_f = 0, nblk = 0, hb = 0, run_left = 0;
    uint32_t last_bits = 3u;
    const uint8_t *bp = pc;
    uint32_t d_last0 = 0u;  // descriptor of the group whose rows complete in the next first-half block
    int seg = 0, seg_g0 = 0, seg_gend = 0;
    WarpTopk wt;
    wt.keys = wkeys;
    wt.cap = next_pow2(a.k + 32) < 64 ? 64 : next_pow2(a.k + 32);
    wt.k = a.k;
    wt.count = 0;
    wt.thr_w = thr_w;
    wt.nw = NW;
    wt.wid = wid;
    wt.ids = nullptr;  // set for the posting-list pass below
    wt.s_off = s_off;
    wt.s_gcum = s_gcum;
    wt.J = 0;
    wt.nres = 0;
    if constexpr (IVF) {
        if (!fused) {  // scan-only launch (the ranking was given): the one pass walks posting lists
            wt.ids = a.ids;
            wt.J = J;
        }
    }

    auto seg_of = [&](int f) -> int {  // segment holding flattened group f
        int lo = 0, hi = J - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_gcum[mid] > f) hi = mid; else lo = mid + 1;
        }
        return lo;
    };
    auto load_seg = [&](int j) {
        seg = j;
        seg_g0 = j ? s_gcum[j - 1] : 0;
        seg_gend = s_gcum[j];
    };
    auto next_run = [&]() {  // (warp-uniform; cur_f < f_end)
        if (cur_f == seg_gend) load_seg(seg + 1);
        const int end = seg_gend < f_end ? seg_gend : f_end;
        run_left = (end - cur_f + 1) * H;  // + the drain blocks: the next blocks in memory
        bp = pc + (size_t)s_prow[seg] * 32 + (size_t)(cur_f - seg_g0) * H * ST_BLOCK_BYTES + lane * 16;
        int tail = 64;
        if (end == seg_gend) tail = s_take[seg] - (seg_gend - seg_g0 - 1) * 64;  // rows of the segment's last group
        last_bits = (lane < tail ? 1u : 0u) | (lane + 32 < tail ? 2u : 0u);
    };
    auto set_range = [&](int nsplit, int split) {  // this warp's slice [f0, f_end) of the pass's groups
        const int G = J ? s_gcum[J - 1] : 0;
        const int nvs = nsplit * NW;
        const int per = (G + nvs - 1) / nvs;
        f0 = (split * NW + wid) * per;
        if (f0 > G) f0 = G;
        f_end = f0 + per < G ? f0 + per : G;
        cur_f = f0;
        hb = 0;
        run_left = 0;
        d_last0 = 0u;
        nblk = 0;
        if (f_end > f0) {
            const int sa_ = seg_of(f0), sb_ = seg_of(f_end - 1);
            nblk = ((f_end - f0) + (sb_ - sa_ + 1)) * H;
            load_seg(sa_);
        }
    };
    uint32_t dsc[ST_R];
#pragma unroll
    for (int s = 0; s < ST_R; ++s) dsc[s] = 0u;
    // coarse pass of the fused kernel.  nlist <= 1024 ("direct"): every distance goes to pool_d[center] and a histogram
    // select ranks the w nearest.  Larger nlist (a.coarse_lists): the pass keeps the w_eff best (distance, center) keys in
    // the warps' top-k lists like any scan, and the CTA merges them.
    bool pass0 = fused;
    bool direct = fused && !a.coarse_lists;
    if (fused && !direct) {
        wt.k = a.w_eff;
        wt.cap = next_pow2(a.w_eff + 32) < 64 ? 64 : next_pow2(a.w_eff + 32);
    }
    __shared__ int s_cnt[NW];
    if (fused) set_range(1, 0);
    else set_range(gridDim.x, blockIdx.x);
    // the first ST_D blocks go out before the table is built
    ST_ISSUE(0, dsc[0], 0 < nblk)
    if constexpr (ST_D >= 2) ST_ISSUE(1 % ST_R, dsc[1 % ST_R], 1 < nblk)
    if constexpr (ST_D == 3) ST_ISSUE(2 % ST_R, dsc[2 % ST_R], 2 < nblk)

    int bad = 0;
    {   // table h (h < H), column c < 64 holds sub-space (32 h + c - 32) mod M: for M = 32 that is c mod 32 (every
        // sub-space twice, so that the lane's column t + 32 - l never wraps); for M = 64 each table holds all 64
        // sub-spaces once, table 0 arranged for the first halves of the rows and table 1 for the second halves.  Entry
        // (m, ks) therefore goes to column (m + 32) & 63 of table 0 and to column m of table H - 1.  Rows >= Ks are zero.
        const int rot = (int)((blockIdx.y * gridDim.x + blockIdx.x) * 53u) & 255;
        float *lutB = lut2 + (H - 1) * (SK_LUT_BYTES / 4);
        if (a.T) {
            const float *T = a.T + (size_t)b * Mr * a.Ks;
            for (int e = threadIdx.x; e < 256 * M; e += NW * 32) {
                const int ks = e / M, m = e % M;
                const float v = ks < a.Ks && m < Mr ? __ldg(T + m * a.Ks + ks) : 0.f;
                bad |= !(v <= ST_TABLE_LIMIT);
                lut2[ks * 64 + ((m + 32) & 63)] = v;
                lutB[ks * 64 + (m & 63)] = v;
            }
        } else {
            // K1 fused (src/rii.h:361-373): the lane's query sub-vector(s) stay in registers, codewords come from the
            // sub-space-fastest copy (one contiguous M*Ds-float row per ks), stores are bank-conflict free.  Every CTA
            // of the grid reads the same M*Ks*Ds floats at about the same time: each starts at a different codeword
            // row, which spreads the requests over the L2 slices.
#pragma unroll
            for (int mh = 0; mh < H; ++mh) {
                const int m = mh * 32 + lane;
                const float *qm = a.Q + (size_t)b * Mr * a.Ds + (size_t)m * a.Ds;
                if (m >= Mr) {  // padded sub-space: a column of zeros
                    for (int ks = wid; ks < 256; ks += NW) {
                        lut2[ks * 64 + ((m + 32) & 63)] = 0.f;
                        lutB[ks * 64 + m] = 0.f;
                    }
                } else if (a.Ds <= 4) {  // every BASELINE shape: the codeword copy is padded to 4 floats (zeros add +0: same sum), 16
                                         // independent 16-byte loads in flight per lane
                    float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (a.q_inline) {  // (one query, b == 0)
                        q4.x = a.qv[m * a.Ds];
                        if (a.Ds > 1) q4.y = a.qv[m * a.Ds + 1];
                        if (a.Ds > 2) q4.z = a.qv[m * a.Ds + 2];
                        if (a.Ds > 3) q4.w = a.qv[m * a.Ds + 3];
                    } else {
                        q4.x = __ldg(qm);
                        if (a.Ds > 1) q4.y = __ldg(qm + 1);
                        if (a.Ds > 2) q4.z = __ldg(qm + 2);
                        if (a.Ds > 3) q4.w = __ldg(qm + 3);
                    }
                    const float4 *cw4 = reinterpret_cast<const float4 *>(a.cw_t) + m;
#pragma unroll 16
                    for (int i = wid; i < 256; i += NW) {
                        const int ks = (i + rot) & 255;
                        float v = sqdist4(make_float2(q4.x, q4.y), make_float2(q4.z, q4.w), __ldg(cw4 + ks * Mr));
                        v = ks < a.Ks ? v : 0.f;
                        bad |= !(v <= ST_TABLE_LIMIT);
                        lut2[ks * 64 + ((m + 32) & 63)] = v;
                        lutB[ks * 64 + m] = v;
                    }
                } else {
#pragma unroll 1
                    for (int ks = wid; ks < 256; ks += NW) {
                        float v = 0.f;
                        if (ks < a.Ks) v = l2sqr_lanes(qm, a.cw_t + ((size_t)ks * Mr + m) * a.Ds, a.Ds, a.variant);
                        bad |= !(v <= ST_TABLE_LIMIT);
                        lut2[ks * 64 + ((m + 32) & 63)] = v;
                        lutB[ks * 64 + m] = v;
                    }
                }
            }
        }
    }
    const bool plain = __syncthreads_or(bad) != 0;  // (also: the table is visible)
    if (dbg && threadIdx.x == 0 && !fused) dbg[1] = clock64();
    if (dbg && threadIdx.x == 0) dbg[4] = clock64();  // table ready

    float keep[32], sel[32];  // the per-lane row-boundary constants of the accumulation (scan_dual.cuh)
#pragma unroll
    for (int t = 0; t < 32; ++t) {
        keep[t] = lane == t ? 0.f : 1.f;
        sel[t] = lane == t ? 1.f : 0.f;
    }
    uint32_t thr_hi = 0xffffffffu;
    uint32_t d_lo = 0xffffffffu, d_hi = 0u;  // coarse pass: range of the distances this lane emitted
    // position of the row `half` (0: x, 1: y) of flattened group f in this lane: the center's index (coarse pass), the id
    // (linear scan), or what warp_compact turns into an id (posting lists: WarpTopk lazy ids)
    auto row_id = [&](int f, int half) -> uint32_t { return (uint32_t)(f * 64 + half * 32 + lane); };
    // K1: the warp's best so far (warp-uniform): distance bits, position, id (-1: not looked up yet), segment
    uint32_t b_thr = 0xffffffffu, b_pos = 0u;
    int b_id = -1, b_seg = -1;
    bool b_have = false;
    auto id_of = [&](uint32_t pos) -> uint32_t {  // position (flattened group << 6 | row) -> key id
        if (!IVF) return pos;                      // linear scans: rows are ids
        const int f = (int)(pos >> 6), j = seg_of(f);
        const int r = (f - (j ? s_gcum[j - 1] : 0)) * 64 + (int)(pos & 63u);
        return (uint32_t)__ldg(a.ids + s_off[j] + r);
    };
    auto emit2 = [&](float dx, float dy, uint32_t d) {
        const int f = (int)(d >> 2);
        if constexpr (K1) {
            if (!pass0) {
                const uint32_t ux = __float_as_uint(dx), uy = __float_as_uint(dy);
                const bool px = (d & 1u) && ux <= b_thr, py = (d & 2u) && uy <= b_thr;
                if (__any_sync(0xffffffffu, px || py)) {
                    const uint32_t mine = umin(px ? ux : 0xffffffffu, py ? uy : 0xffffffffu);
                    const uint32_t mn = __reduce_min_sync(0xffffffffu, mine);
                    // the lowest position at that distance: x rows (row = lane) lie before y rows (row = 32 + lane)
                    const unsigned bx = __ballot_sync(0xffffffffu, px && ux == mn), by = __ballot_sync(0xffffffffu, py && uy == mn);
                    const uint32_t pos = ((uint32_t)f << 6) | (bx ? (uint32_t)(__ffs(bx) - 1) : 32u + (uint32_t)(__ffs(by) - 1));
                    if (!b_have || mn < b_thr) {
                        b_thr = mn; b_pos = pos; b_id = -1; b_seg = IVF ? seg_of(f) : 0; b_have = true;
                    } else if (IVF) {  // an exact tie with the best so far (at a lower position): ids decide across posting lists
                        const int sn = seg_of(f);
                        if (sn != b_seg) {
                            const int idn = (int)id_of(pos);
                            if (b_id < 0) b_id = (int)id_of(b_pos);
                            if (idn < b_id) { b_pos = pos; b_id = idn; b_seg = sn; }
                        }
                    }
                }
                return;
            }
        }
        if (IVF && direct) {  // coarse pass of the fused kernel: keep every distance (and their range, for the selection)
            const uint32_t ux = __float_as_uint(dx), uy = __float_as_uint(dy);
            if (d & 1u) { pool_d[f * 64 + lane] = ux; d_lo = ux < d_lo ? ux : d_lo; d_hi = ux > d_hi ? ux : d_hi; }
            if (d & 2u) { pool_d[f * 64 + 32 + lane] = uy; d_lo = uy < d_lo ? uy : d_lo; d_hi = uy > d_hi ? uy : d_hi; }
            return;
        }
        // distance part of the CTA threshold: long linear scans keep a cached copy that is refreshed after every push
        // and every few blocks (a stale value is merely less strict); the short per-query IVF passes re-read it at
        // every emission
        if constexpr (IVF) thr_hi = reinterpret_cast<volatile uint32_t *>(cta_thr)[1];
        const bool px = (d & 1u) && __float_as_uint(dx) <= thr_hi;
        const bool py = (d & 2u) && __float_as_uint(dy) <= thr_hi;
        if (__any_sync(0xffffffffu, px || py)) {
            warp_push(wt, cta_thr, lane, dx, px ? row_id(f, 0) : 0u, px);
            warp_push(wt, cta_thr, lane, dy, py ? row_id(f, 1) : 0u, py);
            thr_hi = reinterpret_cast<volatile uint32_t *>(cta_thr)[1];
        }
    };
    // exact for ANY table (inf / NaN / huge entries): one candidate per lane and half group, natural order of
    // additions; the lane un-skews its own stream (byte m of row b of stream s = stream byte 32 b + m)
    auto plain_slice = [&]() {
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
                    const uint8_t *p = pc + (size_t)s_prow[j] * 32 + (size_t)g * H * ST_BLOCK_BYTES + half * 1024 + lane * 16;
                    float d = 0.f;
                    for (int m = 0; m < M; ++m) {
                        const int x = lane + m;  // byte of the span [window b | window b + 1 | ...] of stream s
                        const uint32_t ks = __ldg(p + (x >> 5) * ST_BLOCK_BYTES + ((x >> 4) & 1) * 512 + (x & 15));
                        const float v = lut2[ks * 64 + ((m + 32) & 63)];  // table 0 holds every sub-space
                        d = m ? __fadd_rn(d, v) : v;
                    }
                    dd[half] = d;
                }
            }
            emit2(dd[0], dd[1], dsc_);
        }
    };

    const int npass = fused ? 2 : 1;
#pragma unroll 1
    for (int pass = 0; pass < npass; ++pass) {
        if (pass == 1) {
            // ---- between the passes: select + rank the w_eff nearest centers, plan (all in shared memory) ----------
            if (dbg && threadIdx.x == 0) dbg[5] = clock64();  // coarse pass done (the pass loop ended with a barrier)
            u64 *selk = reinterpret_cast<u64 *>(smem_raw + hi0);          // (rings idle) <= 256 selected keys, the full sort (nlist <= 1024), or the merged warp lists
            int np = 0;
            if (direct) {
                np = a.w_eff <= 224 ? cta_select_smallest<NW * 32>(pool_d, a.nlist, a.w_eff, selk, hist, true) : -1;
                if (np < 0) {  // many lists to rank (subset searches), or > 256 exact ties at the w-th distance: sort all (dist, index) keys
                    const int P = next_pow2(a.nlist);
                    for (int i = threadIdx.x; i < P; i += NW * 32) selk[i] = i < a.nlist ? (((u64)pool_d[i] << 32) | (u64)(uint32_t)i) : RII_KEY_MAX;
                    cta_sort_smem<NW * 32>(selk, P);
                    np = a.nlist;
                }
            } else {  // merge the warps' sorted lists (each <= w_eff keys) into the ranked list
                if (lane == 0) s_cnt[wid] = wt.count;
                __syncthreads();
                int tot = 0;
                for (int w2 = 0; w2 < NW; ++w2) tot += s_cnt[w2];
                const u64 *allkeys = reinterpret_cast<const u64 *>(smem_raw);
                if (tot <= 256) {
                    if (wid == 0) {
                        int o = 0;
                        for (int w2 = 0; w2 < NW; ++w2) {
                            for (int i = lane; i < s_cnt[w2]; i += 32) selk[o + i] = allkeys[(size_t)w2 * capw + i];
                            o += s_cnt[w2];
                        }
                        warp_sort_any(selk, tot, lane);
                    }
                } else {
                    BlockTopk tk;
                    const int mcap = next_pow2(NW * a.w_eff + 1);
                    tk.keys = selk;
                    tk.thr = selk + mcap;
                    tk.count = reinterpret_cast<int *>(selk + mcap + 1);
                    tk.cap = mcap;
                    tk.k = a.w_eff;
                    tk.init();
                    for (int w2 = 0; w2 < NW; ++w2)
                        for (int i = threadIdx.x; i < s_cnt[w2]; i += blockDim.x) tk.push(allkeys[(size_t)w2 * capw + i]);
                    tk.compact();
                }
                __syncthreads();
                np = tot;
                if (threadIdx.x < NW) thr_w[threadIdx.x] = RII_KEY_MAX;  // (cta_thr is reset with the plan below)
            }
            if (wid == 0) {
                if (dbg && lane == 0) { dbg[6] = clock64(); dbg[7] = np; }
                int *ranked_g = a.plan.ranked + (size_t)b * a.w_eff;
                for (int j = lane; j < a.w_eff; j += 32) {  // w_eff <= nlist, np >= w_eff
                    const int no = (int)key_id(selk[j]);
                    if (blockIdx.x == 0) ranked_g[j] = no;  // (every CTA of the query computes the same ranking)
                    s_f[j] = a.plan.glob_len[no];
                    s_pre[j] = a.plan.pre_len ? a.plan.pre_len[no] : 0;
                    s_loc[j] = a.plan.loc_len[no];
                    s_off[j] = a.offsets[no];
                    s_prow[j] = a.skew_off[no];
                }
                __syncwarp();
                const int jc = a.coarse_mode == 1 ? 0 : plan_warp(a.plan, b, lane, s_f, s_pre, s_loc, s_off, s_prow, s_gcum, s_take, blockIdx.x == 0);
                if (lane == 0) {
                    s_plan[0] = jc;
                    *cta_thr = RII_KEY_MAX;
                }
            }
            if (a.coarse_mode == 1) return;  // coarse-only launch: the ranking is all that was asked for (CTA-uniform)
            __syncthreads();
            if (dbg && threadIdx.x == 0) dbg[1] = clock64();
            J = s_plan[0];
            pc = a.codes;
            direct = false;
            pass0 = false;
            wt.k = a.k;
            wt.cap = next_pow2(a.k + 32) < 64 ? 64 : next_pow2(a.k + 32);
            wt.count = 0;
            wt.nres = 0;
            wt.ids = a.ids;
            wt.J = J;
            thr_hi = 0xffffffffu;
            set_range(gridDim.x, blockIdx.x);  // this CTA's share of the planned groups (also resets the walk state: hb, drain_left, d_last0)
#pragma unroll
            for (int s = 0; s < ST_R; ++s) dsc[s] = 0u;
            ST_ISSUE(0, dsc[0], 0 < nblk)
            if constexpr (ST_D >= 2) ST_ISSUE(1 % ST_R, dsc[1 % ST_R], 1 < nblk)
            if constexpr (ST_D == 3) ST_ISSUE(2 % ST_R, dsc[2 % ST_R], 2 < nblk)
        }

        if (plain) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            plain_slice();
        } else {
            unsigned long long acc2 = 0ull, out2 = 0ull;
#pragma unroll 1
            for (int m = 0; m < nblk; m += ST_R) {
                if constexpr (!IVF) thr_hi = reinterpret_cast<volatile uint32_t *>(cta_thr)[1];
                ST_STAGE(0)
                ST_STAGE(1)
                if constexpr (ST_R >= 3) { ST_STAGE(2 % ST_R) }
                if constexpr (ST_R == 4) { ST_STAGE(3 % ST_R) }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");  // (only empty groups can be pending; the ring area is reused below)
        if (IVF && direct) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const uint32_t x = __shfl_xor_sync(0xffffffffu, d_lo, o), y = __shfl_xor_sync(0xffffffffu, d_hi, o);
                d_lo = x < d_lo ? x : d_lo;
                d_hi = y > d_hi ? y : d_hi;
            }
            if (lane == 0) {
                atomicMin(reinterpret_cast<uint32_t *>(hist) + 256, d_lo);
                atomicMax(reinterpret_cast<uint32_t *>(hist) + 257, d_hi);
            }
        } else if (K1 && !pass0) {
            if (b_have && b_id < 0) b_id = (int)id_of(b_pos);
            if (lane == 0) wkeys[0] = ((u64)b_thr << 32) | (u64)(uint32_t)b_id;
            wt.count = b_have ? 1 : 0;
            __syncwarp();
        } else {
            warp_compact(wt, cta_thr, lane);
        }
        __syncthreads();
    }
    if (dbg && threadIdx.x == 0) dbg[2] = clock64();
    {   // CTA merge of the (sorted) warp lists, reusing the lut2 area for the keys
        if (lane == 0) s_cnt[wid] = wt.count;
        __syncthreads();
        int tot = 0;
        for (int w2 = 0; w2 < NW; ++w2) tot += s_cnt[w2];
        const u64 *allkeys = reinterpret_cast<const u64 *>(smem_raw);
        if (K1 && NW <= 32) {
            // topk == 1: every warp holds at most one key -- a warp minimum under (distance, id), no sort
            if (wid == 0) {
                u64 key = lane < NW && s_cnt[lane] ? allkeys[(size_t)lane * capw] : RII_KEY_MAX;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const u64 y = __shfl_xor_sync(0xffffffffu, key, o);
                    key = y < key ? y : key;
                }
                if (lane == 0) {
                    if (a.out.final) {
                        if (key != RII_KEY_MAX) {
                            a.out.out_ids[b] = a.out.id_map ? a.out.id_map[key_id(key)] : a.out.id_base + (long long)key_id(key);
                            a.out.out_dists[b] = key_dist(key);
                        }
                        a.out.out_counts[b] = key != RII_KEY_MAX ? 1 : 0;
                    } else {
                        a.out.partial[(size_t)b * gridDim.x + blockIdx.x] = key;
                    }
                }
            }
        } else if (tot <= 256) {
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
                        a.out.out_ids[(size_t)b * a.k + i] = a.out.id_map ? a.out.id_map[key_id(mk[i])] : a.out.id_base + (long long)key_id(mk[i]);
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
    if (!a.out.final && a.out.merge_cnt) {
        // several CTAs served this query: the last one to deliver its partial list merges them all (gridDim.x * k <= 256 keys;
        // the host falls back to k_merge otherwise).  Classic last-block pattern: fence, count, the last arriver reads.
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(a.out.merge_cnt + b, 1) == (int)gridDim.x - 1 ? 1 : 0;
        __syncthreads();
        if (s_last && wid == 0) {
            __threadfence();
            u64 *mk = reinterpret_cast<u64 *>(smem_raw + lut_off);
            const int tot = (int)gridDim.x * a.k;
            const u64 *src = a.out.partial + (size_t)b * gridDim.x * a.k;
            if (K1) {  // one key per CTA: a warp minimum
                u64 key = RII_KEY_MAX;
                for (int i = lane; i < tot; i += 32) {
                    const u64 y = __ldcg(src + i);
                    key = y < key ? y : key;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const u64 y = __shfl_xor_sync(0xffffffffu, key, o);
                    key = y < key ? y : key;
                }
                if (lane == 0) mk[0] = key;
                __syncwarp();
            } else {
                for (int i = lane; i < tot; i += 32) mk[i] = __ldcg(src + i);
                warp_sort_any(mk, tot, lane);
            }
            int n = 0;
            for (int i = lane; i < a.k; i += 32) {
                const u64 key = mk[i];
                if (key != RII_KEY_MAX) {
                    a.out.out_ids[(size_t)b * a.k + i] = a.out.id_map ? a.out.id_map[key_id(key)] : a.out.id_base + (long long)key_id(key);
                    a.out.out_dists[(size_t)b * a.k + i] = key_dist(key);
                    ++n;
                }
            }
            n = __reduce_add_sync(0xffffffffu, n);
            if (lane == 0) {
                a.out.out_counts[b] = n;
                a.out.merge_cnt[b] = 0;
            }
        }
    }
    if (dbg && threadIdx.x == 0) dbg[3] = clock64();
}

// dynamic shared memory of a launch shape, or 0 if keys + thresholds + segment tables do not fit below the table
static inline size_t stream_smem_bytes(bool ivf, int nw, int ring_stages, uint32_t tb, int capw, int w_eff, size_t pool_bytes, int H = 1)
{
    const size_t meta = (((size_t)nw * capw * 8 + 8 + (size_t)nw * 8 + (size_t)(ivf ? w_eff : 1) * 40 + 16 + 15) & ~(size_t)15) +
                        (pool_bytes ? (pool_bytes + 15) / 16 * 16 + 1040 : 0);  // coarse distances + selection histogram
    if (meta > (size_t)tb - 1280) return 0;  // the window starts at 1 KB + static shared memory (<= 256 B allowed for)
    return (size_t)tb - 1024 + (size_t)H * SK_LUT_BYTES + (size_t)nw * ring_stages * ST_BLOCK_BYTES;
}
