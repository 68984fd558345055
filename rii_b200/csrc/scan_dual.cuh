// ===================================================================================================
// K2/K5 v3 ("dual"): the skewed, bank-conflict-free scan engine of kernels.cuh (k_scan_skew32) with
//   (1) TWO code streams per lane that share the lane's phase, and
//   (2) predicate-free accumulation with the packed fp32x2 FMA of sm_100 (FFMA2).
//
// Why.  The v2 engine is issue bound: per lookup PRMT + LDS + ISETP + 2 predicated FADD (+ 1/2 for the code word)
// = 6.3 issue slots with everything else, 70 % of the HBM roofline on the linear scan (profiles/r01_*).  The
// "which row does this lookup belong to" decision (lane l switches rows at step l of every 32-step block) is
// moved from predicates into DATA:
//        acc = acc * keep_t + v          keep_t = (lane == t) ? 0 : 1      restarts the sum at the row boundary
//        out = acc * sel_t  + out        sel_t  = (lane == t) ? 1 : 0      captures the finished sum (out = 0 per block)
// keep_t / sel_t are 64 per-lane constants in registers.  x * 1 + y rounds once, exactly like x + y; x * 0 + y = y
// exactly for finite x >= 0: every candidate's distance is still the sequential fp32 sum T[0][c0] + T[1][c1] + ...
// of src/rii.h:386-394, bit for bit.  With two streams per lane both updates are ONE FFMA2 each on the
// (stream x, stream y) register pair, the constant broadcast from a 32-bit register:
//        per 2 lookups: 2 PRMT + 2 LDS + 2 FFMA2          (3 + 1/2 slots per lookup instead of 5 + 1/2)
// Measured in isolation (tools/ubench_step.cu, profiles/r01_ubench_step.jsonl): 24.8 lookups/clk/SM = the
// shared-memory limit of 1.25 wavefronts per warp-lookup, 7.2 T lookups/s per B200 -- above what HBM can feed.
// The trick needs finite partial sums (inf * 0 = NaN): the table build checks every entry against 1e37 and a
// table that fails (queries ~1e18 away from the codewords, inf, NaN) takes `plain_slice`, a plain per-candidate
// loop over the same table -- slow, exact for anything.
//
// Staging.  Per stream a region [pad 32 B | stage 0 | stage 1 | stage 2], a stage = DU_J = 2 rows; 224 B = 56
// words = 24 mod 32, which keeps the lanes' code-word reads conflict free.  A warp's tile = 64 streams x 2 rows
// = 128 rows = 4 KB, cp.async'ed as 512 contiguous bytes per warp instruction; tiles n+1 and n+2 are in flight
// while tile n is scanned (3-stage ring: ~2 tile times of latency tolerance, 8 KB per warp in flight).  The ring
// wraps after stage 2: the lagging reads of the first block of stage 0 are redirected to the tail of stage 2
// (DU_WORD_WRAP), so no carry row has to be copied.
// ===================================================================================================
#pragma once

#define DU_J 2
#define DU_STAGES 3
#define DU_TILE_ROWS (64 * DU_J)                           // 128 rows per warp per stage
#define DU_STAGE_BYTES (DU_J * 32)                         // 64
#define DU_RING_BYTES (DU_STAGES * DU_STAGE_BYTES)         // 192
#define DU_REGION_BYTES (32 + DU_RING_BYTES)               // 224
#define DU_Y_OFF (32 * DU_REGION_BYTES)                    // the y stream of lane l is stream 32 + l
#define DU_WARP_BYTES (64 * DU_REGION_BYTES)               // 14336
#define DU_TABLE_LIMIT 1e37f                               // 32 entries below this cannot overflow fp32

// one lookup step of both streams.  Table at absolute shared address 0x10000 (see SK_STEP in kernels.cuh).
#define DU_STEP(WX, WY, BYTE, T)                                                                              \
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
#define DU_WORD_AT(OX, OY, Q)                                                                                 \
    {                                                                                                         \
        const uint32_t x_ = *reinterpret_cast<const uint32_t *>(smem_raw + (OX));                             \
        const uint32_t y_ = *reinterpret_cast<const uint32_t *>(smem_raw + (OY));                             \
        const uint32_t wx_ = __funnelshift_rc(xprev, x_, shift);                                              \
        const uint32_t wy_ = __funnelshift_rc(yprev, y_, shift);                                              \
        xprev = x_;                                                                                           \
        yprev = y_;                                                                                           \
        DU_STEP(wx_, wy_, 0, 4 * (Q) + 0)                                                                     \
        DU_STEP(wx_, wy_, 1, 4 * (Q) + 1)                                                                     \
        DU_STEP(wx_, wy_, 2, 4 * (Q) + 2)                                                                     \
        DU_STEP(wx_, wy_, 3, 4 * (Q) + 3)                                                                     \
    }
// 4 steps = one code word of each of the lane's two (lagged) streams; OFF = byte offset of the block's word 0 in x
#define DU_WORD(OFF, Q) DU_WORD_AT((OFF) + 4 * (Q), (OFF) + DU_Y_OFF + 4 * (Q), Q)
// first block of stage 0: words that still belong to the previous row (Q < lane / 4) sit at the tail of stage 2
#define DU_WORD_WRAP(OFF, Q)                                                                                  \
    {                                                                                                         \
        const uint32_t wo_ = (OFF) + 4 * (Q) + ((Q) < lagw ? (uint32_t)DU_RING_BYTES : 0u);                   \
        DU_WORD_AT(wo_, wo_ + DU_Y_OFF, Q)                                                                    \
    }
#define DU_BLOCK(OFF)                                                                                         \
    {                                                                                                         \
        DU_WORD(OFF, 0) DU_WORD(OFF, 1) DU_WORD(OFF, 2) DU_WORD(OFF, 3)                                       \
        DU_WORD(OFF, 4) DU_WORD(OFF, 5) DU_WORD(OFF, 6) DU_WORD(OFF, 7)                                       \
    }
#define DU_BLOCK_WRAP(OFF)                                                                                    \
    {                                                                                                         \
        DU_WORD_WRAP(OFF, 0) DU_WORD_WRAP(OFF, 1) DU_WORD_WRAP(OFF, 2) DU_WORD_WRAP(OFF, 3)                   \
        DU_WORD_WRAP(OFF, 4) DU_WORD_WRAP(OFF, 5) DU_WORD_WRAP(OFF, 6) DU_WORD_WRAP(OFF, 7)                   \
    }
// end of a block: out2 holds the finished distances of local candidates eloc (x) and eloc + 64 (y)
#define DU_EMIT()                                                                                             \
    {                                                                                                         \
        float dx_, dy_;                                                                                       \
        asm("mov.b64 {%0, %1}, %2;" : "=f"(dx_), "=f"(dy_) : "l"(out2));                                      \
        out2 = 0ull;                                                                                          \
        emit2(dx_, dy_, eloc);                                                                                \
    }

// Args / phases / top-k exactly as k_scan_skew32 (SkewArgs; IVF with a.centers != null = fused coarse + plan + scan).
template <int NW, bool IVF>
__global__ void __launch_bounds__(NW * 32, 1) k_scan_dual32(SkewArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout (dynamic shared memory; the window starts at absolute shared address ~1 KB):
    //   [NW key buffers][cta_thr][thr_w][IVF: s_off i64[w] | s_cum, s_f, s_pre, s_loc i32[w] | s_plan i32[4]][n_lo warp regions]
    //   ... lut2 (64 KB) at ABSOLUTE shared address 0x10000 ... [n_hi warp regions]
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t lut_off = 0x10000u - smem_base;
    float *lut2 = reinterpret_cast<float *>(smem_raw + lut_off);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int capw = a.cap;
    long long *dbg = a.dbg ? a.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();
    u64 *wkeys = reinterpret_cast<u64 *>(smem_raw) + (size_t)wid * capw;
    u64 *cta_thr = reinterpret_cast<u64 *>(smem_raw) + (size_t)NW * capw;
    u64 *thr_w = cta_thr + 1;  // [NW]
    long long *s_off = reinterpret_cast<long long *>(smem_raw + (size_t)NW * capw * 8 + 8 + NW * 8);
    const int wq = IVF ? a.w_eff : 0;
    int *s_cum = reinterpret_cast<int *>(s_off + wq);
    int *s_f = s_cum + wq, *s_pre = s_f + wq, *s_loc = s_pre + wq, *s_plan = s_loc + wq;  // s_plan: [J]
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
    __syncthreads();

    const uint32_t lo_reg0 = (uint32_t)(((size_t)NW * capw * 8 + 16 + NW * 8 + (IVF ? (size_t)a.w_eff * 24 + 32 : 0) + 15) & ~(size_t)15);
    const uint32_t hi_reg0 = lut_off + SK_LUT_BYTES;
    const int n_lo = lut_off > lo_reg0 ? (int)((lut_off - lo_reg0) / DU_WARP_BYTES) : 0;
    if (n_lo + (int)((a.smem_bytes - hi_reg0) / DU_WARP_BYTES) < NW) __trap();  // host sized the launch wrongly
    const uint32_t region = wid < n_lo ? lo_reg0 + wid * DU_WARP_BYTES : hi_reg0 + (wid - n_lo) * DU_WARP_BYTES;
    const uint32_t lagw = (uint32_t)lane >> 2;                               // whole words of the lane's lag
    const uint32_t rb = region + lane * DU_REGION_BYTES + 32 - 4 * lagw;     // x stream, word part of the lag folded in
    const uint32_t shift = 8 * (4 - (lane & 3));                             // funnel shift (32 == no byte lag)
    const uint32_t colreg = 0x00010000u | (uint32_t)((32 - lane) * 4);       // table address 0x10000 | column byte offset
    // destination of 16-byte chunk (it, lane) of a tile: row = 16 it + lane / 2 -> stream row / 2, slot row % 2
    const uint32_t cp_dst = region + (lane >> 2) * DU_REGION_BYTES + 32 + (lane & 3) * 16;
    float keep[32], sel[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) {
        keep[t] = lane == t ? 0.f : 1.f;
        sel[t] = lane == t ? 1.f : 0.f;
    }

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
        const long long per_cta = ((tot + nsplit - 1) / nsplit + NW * DU_TILE_ROWS - 1) / (NW * DU_TILE_ROWS) * (NW * DU_TILE_ROWS);
        base = (long long)split * per_cta + (long long)wid * (per_cta / NW);
        end = base + per_cta / NW;
        if (end > tot) end = tot;
        cnt = end > base ? (int)(end - base) : 0;
        ntiles = (cnt + DU_TILE_ROWS - 1) / DU_TILE_ROWS;
    };
    // every call commits exactly one cp.async group (an empty one past the last tile): "wait_group 1" at the top
    // of tile n then always means "tile n has landed"
    auto issue_tile = [&](int n, uint32_t st_off) {  // rows [base + 128 n, +128) of the contiguous table pc
        if (n < ntiles) {
            const long long r0 = base + (long long)n * DU_TILE_ROWS;
            const uint8_t *g = pc + r0 * 32 + lane * 16;
            const uint32_t dst = smem_base + cp_dst + st_off;
            if (r0 + DU_TILE_ROWS <= end) {  // full tile: 8 x 512 contiguous bytes per warp, immediates only
#pragma unroll
                for (int it = 0; it < DU_TILE_ROWS / 16; ++it)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + it * 8 * DU_REGION_BYTES), "l"(g + it * 512));
            } else {                          // last tile: rows past the end are zero filled, their results masked
#pragma unroll
                for (int it = 0; it < DU_TILE_ROWS / 16; ++it) {
                    const int nbytes = r0 + it * 16 + (lane >> 1) < end ? 16 : 0;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + it * 8 * DU_REGION_BYTES),
                                 "l"(nbytes ? g + it * 512 : pc), "r"(nbytes));
                }
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    // IVF pass 1: tile n = flattened candidates [c0, c0 + 128) of the plan; segment j covers [cum[j-1], cum[j]) and
    // starts at row s_off[j] of the list-ordered code copy.  segw = segment of c0 (warp-uniform, carried along).
    auto issue_tile_seg = [&](int n, uint32_t st_off) {
        if (n < ntiles) {
            const int c0 = (int)base + n * DU_TILE_ROWS;
            const int cend = (int)end;
            while (segw < J - 1 && s_cum[segw] <= c0) ++segw;
            const int seg_lo = segw ? s_cum[segw - 1] : 0;
            const uint32_t dst = smem_base + cp_dst + st_off;
            if (c0 + DU_TILE_ROWS <= cend && c0 + DU_TILE_ROWS <= s_cum[segw]) {  // one segment, full tile: pure stream
                const uint8_t *g = pc + (size_t)(s_off[segw] + (c0 - seg_lo)) * 32 + lane * 16;
#pragma unroll
                for (int it = 0; it < DU_TILE_ROWS / 16; ++it)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + it * 8 * DU_REGION_BYTES), "l"(g + it * 512));
            } else {                                                             // crosses a segment boundary / tail
#pragma unroll
                for (int it = 0; it < DU_TILE_ROWS / 16; ++it) {
                    const int c = c0 + it * 16 + (lane >> 1);
                    const bool ok = c < cend;
                    int seg = segw;
                    if (ok) while (s_cum[seg] <= c) ++seg;
                    const uint8_t *g = pc + (size_t)(s_off[seg] + (c - (seg ? s_cum[seg - 1] : 0))) * 32 + (lane & 1) * 16;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + it * 8 * DU_REGION_BYTES),
                                 "l"(ok ? g : pc), "r"(ok ? 16 : 0));
                }
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    // IVF pass 1: segment and row (of the list-ordered copy == position in a.ids) of flattened candidate c < total
    auto cand_pos = [&](int c) -> long long {
        int lo = 0, hi = J - 1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (s_cum[mid] > c) hi = mid; else lo = mid + 1;
        }
        return s_off[lo] + (c - (lo ? s_cum[lo - 1] : 0));
    };

    // ---- pass setup: the first two tiles go out before the table is built ----------------------------------
    bool segm = IVF && !fused;  // true: pass over planned posting-list segments; false: plain row range
    bool direct = fused;         // coarse pass: every distance goes to pool_d[center]
    // the warps' key buffers are idle during the coarse pass: they hold the nlist distances (host checks the size)
    uint32_t *pool_d = reinterpret_cast<uint32_t *>(smem_raw);
    wt.cap = next_pow2(wt.k + 32) < 64 ? 64 : next_pow2(wt.k + 32);
    if (fused) {
        pc = a.centers;
        set_range(a.nlist, 1, 0);
    } else if (IVF) {
        set_range(J ? (long long)s_cum[J - 1] : 0, gridDim.x, blockIdx.x);
    } else {
        set_range(a.N, gridDim.x, blockIdx.x);
    }
    if (segm) { issue_tile_seg(0, 0); issue_tile_seg(1, DU_STAGE_BYTES); }
    else { issue_tile(0, 0); issue_tile(1, DU_STAGE_BYTES); }
    int bad = 0;
    {   // lut2[ks][c] = T[c % 32][ks]; rows >= Ks are zero (zero-filled padding rows index row 0 only)
        if (a.T) {
            const float *T = a.T + (size_t)b * 32 * a.Ks;
#pragma unroll 8
            for (int e = threadIdx.x; e < 256 * 64; e += NW * 32) {
                int ks = e >> 6, c = e & 63;
                const float v = ks < a.Ks ? __ldg(T + (c & 31) * a.Ks + ks) : 0.f;
                bad |= !(v <= DU_TABLE_LIMIT);
                lut2[e] = v;
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
                    bad |= !(v <= DU_TABLE_LIMIT);
                    lut2[ks * 64 + lane] = v;
                    lut2[ks * 64 + lane + 32] = v;
                }
            } else {
#pragma unroll 1
                for (int ks = wid; ks < 256; ks += NW) {
                    float v = 0.f;
                    if (ks < a.Ks) v = l2sqr_lanes(qm, a.cw_t + ((size_t)ks * 32 + lane) * a.Ds, a.Ds, a.variant);
                    bad |= !(v <= DU_TABLE_LIMIT);
                    lut2[ks * 64 + lane] = v;
                    lut2[ks * 64 + lane + 32] = v;
                }
            }
        }
    }
    const bool plain = __syncthreads_or(bad) != 0;  // (also: the table is visible)
    if (dbg && threadIdx.x == 0 && !fused) dbg[1] = clock64();
    if (dbg && threadIdx.x == 0) dbg[4] = clock64();  // table ready

    uint32_t thr_hi = 0xffffffffu;
    // one finished candidate per lane: local index loc of this warp's slice, distance d
    auto emit1 = [&](float d, uint32_t loc, bool pre) {
        const uint32_t id = (IVF && segm) ? (pre ? (uint32_t)__ldg(a.ids + cand_pos((int)(base + loc))) : 0u) : (uint32_t)(base + loc);
        warp_push(wt, cta_thr, lane, d, id, pre);
    };
    auto emit2 = [&](float dx, float dy, uint32_t ex) {
        const uint32_t ey = ex + 64;
        if (IVF && direct) {  // coarse pass of the fused kernel: keep every distance
            if (ex < (uint32_t)cnt) pool_d[(uint32_t)base + ex] = __float_as_uint(dx);
            if (ey < (uint32_t)cnt) pool_d[(uint32_t)base + ey] = __float_as_uint(dy);
            return;
        }
        // distance part of the CTA threshold: long linear scans re-read it once per tile and after every push (a
        // stale value is merely less strict); the short per-query IVF passes re-read it at every emission
        if constexpr (IVF) thr_hi = reinterpret_cast<volatile uint32_t *>(cta_thr)[1];
        const bool px = ex < (uint32_t)cnt && __float_as_uint(dx) <= thr_hi;
        const bool py = ey < (uint32_t)cnt && __float_as_uint(dy) <= thr_hi;
        if (__any_sync(0xffffffffu, px || py)) {
            emit1(dx, ex, px);
            emit1(dy, ey, py);
            thr_hi = reinterpret_cast<volatile uint32_t *>(cta_thr)[1];
        }
    };
    // exact for ANY table (inf / NaN / huge entries): one candidate per lane, natural order of additions, no tricks
    auto plain_slice = [&]() {
        asm volatile("cp.async.wait_group 0;");
        for (int c0 = 0; c0 < cnt; c0 += 32) {
            const int c = c0 + lane;
            float d = 0.f;
            if (c < cnt) {
                const long long row = (IVF && segm) ? cand_pos((int)(base + c)) : base + c;
                const uint4 *g = reinterpret_cast<const uint4 *>(pc + (size_t)row * 32);
                const uint4 r0 = __ldg(g), r1 = __ldg(g + 1);
                const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                d = lut2[(w[0] & 0xff) * 64];
#pragma unroll
                for (int m = 1; m < 32; ++m) d = __fadd_rn(d, lut2[((w[m >> 2] >> (8 * (m & 3))) & 0xff) * 64 + m]);
            }
            if (IVF && direct) {
                if (c < cnt) pool_d[(uint32_t)base + c] = __float_as_uint(d);
            } else {
                const u64 thr = *reinterpret_cast<volatile u64 *>(cta_thr);
                // (NaN distances: the reference's comparator never ranks them first either; they are dropped here)
                const bool pre = c < cnt && pack_key(d, 0u) <= (thr | 0xffffffffull) && d == d;
                if (__any_sync(0xffffffffu, pre)) emit1(d, (uint32_t)c, pre);
            }
        }
    };

    const int npass = fused ? 2 : 1;
#pragma unroll 1
    for (int pass = 0; pass < npass; ++pass) {
        if (pass == 1) {
            // ---- between the passes: select + rank the w_eff nearest centers, plan (all in shared memory) ----------
            if (dbg && threadIdx.x == 0) dbg[5] = clock64();  // coarse pass done (the pass loop ended with a barrier)
            u64 *selk = reinterpret_cast<u64 *>(smem_raw + hi_reg0);         // the regions are idle now
            int *hist = reinterpret_cast<int *>(smem_raw + hi_reg0 + 4096);  // 256 keys above `selk`
            int np = cta_select_smallest<NW * 32>(pool_d, a.nlist, a.w_eff, selk, hist);
            if (wid == 0) {
                if (np < 0) {  // > 256 exact ties at the w-th distance: full sort of all (dist, index) keys
                    const int P = next_pow2(a.nlist);
                    for (int i = lane; i < P; i += 32) selk[i] = i < a.nlist ? (((u64)pool_d[i] << 32) | (u64)(uint32_t)i) : RII_KEY_MAX;
                    warp_sort_smem(selk, P, lane);
                    np = a.nlist;
                }
                if (dbg && lane == 0) { dbg[6] = clock64(); dbg[7] = np; }
                int *ranked_g = a.plan.ranked + (size_t)b * a.w_eff;
                for (int j = lane; j < a.w_eff; j += 32) {  // w_eff <= nlist, np >= w_eff
                    const int no = (int)key_id(selk[j]);
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
            thr_hi = 0xffffffffu;
            set_range(J ? (long long)s_cum[J - 1] : 0, 1, 0);
            issue_tile_seg(0, 0);
            issue_tile_seg(1, DU_STAGE_BYTES);
        }

        if (plain) {
            plain_slice();
        } else {
            unsigned long long acc2 = 0ull, out2 = 0ull;
            uint32_t xprev = 0, yprev = 0;
            // local index of the x candidate whose distance completes at the end of the current block: it started one
            // block earlier, so the first block completes nothing (index "-1" of the previous tile: fails eloc < cnt)
            uint32_t eloc = (uint32_t)(DU_J * lane + DU_J - 1 - DU_TILE_ROWS);
            uint32_t st_off = 0;  // stage of tile n, in bytes
#pragma unroll 1
            for (int n = 0; n < ntiles; ++n) {
                asm volatile("cp.async.wait_group 1;");  // tile n has landed (tile n + 1 may still be in flight)
                __syncwarp();                            // rows were written by other lanes of the warp
                if constexpr (!IVF) thr_hi = reinterpret_cast<volatile uint32_t *>(cta_thr)[1];
                const uint32_t rbw = rb + st_off;
                if (st_off == 0) DU_BLOCK_WRAP(rbw)
                else DU_BLOCK(rbw)
                DU_EMIT()
                eloc += DU_TILE_ROWS - DU_J + 1;
                __syncwarp();  // every lane is done with the tail of tile n - 1: its stage takes tile n + 2
                {
                    const uint32_t st2 = st_off == 0 ? 2 * DU_STAGE_BYTES : st_off - DU_STAGE_BYTES;
                    if (segm) issue_tile_seg(n + 2, st2);
                    else issue_tile(n + 2, st2);
                }
                DU_BLOCK(rbw + 32)
                DU_EMIT()
                eloc += 1;
                st_off = st_off == 2 * DU_STAGE_BYTES ? 0 : st_off + DU_STAGE_BYTES;
            }
            if (ntiles > 0) {  // drain: 32 more steps complete the last row of every stream (the rest of the block
                               // reads stale rows of the lane's own ring; nothing of it is ever captured)
                const uint32_t rbw = rb + st_off;
                if (st_off == 0) DU_BLOCK_WRAP(rbw)
                else DU_BLOCK(rbw)
                DU_EMIT()
            }
            asm volatile("cp.async.wait_group 0;");
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
        const u64 *allkeys = reinterpret_cast<const u64 *>(smem_raw);
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

// Host-side sizing: how many warp regions fit below and above the 64 KB table pinned at absolute shared address
// 0x10000 (the window starts at ~1 KB: reserved + static shared memory; both extremes are allowed for).
static inline int dual_regions_fit(bool ivf, int nw, int capw, int w_eff)
{
    const size_t meta = (((size_t)nw * capw * 8 + 16 + (size_t)nw * 8 + (ivf ? (size_t)w_eff * 24 + 32 : 0)) + 15) & ~(size_t)15;
    const long long lut_off_min = 0x10000 - 1280, lut_off_max = 0x10000 - 1024;
    const long long n_lo = (lut_off_min - (long long)meta) / DU_WARP_BYTES;
    const long long n_hi = ((long long)SK_DYN_SMEM - (lut_off_max + SK_LUT_BYTES)) / DU_WARP_BYTES;
    return (int)((n_lo < 0 ? 0 : n_lo) + (n_hi < 0 ? 0 : n_hi));
}
