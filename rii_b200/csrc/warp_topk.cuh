// Warp-level top-k lists, register bitonic sorts and the CTA-wide histogram select used by the streaming scan engine
// (scan_stream.cuh) and the assignment engine (assign_stream.cuh).
#pragma once
#include "topk.cuh"

#define SK_TILE_ROWS 128                         // rows per warp used by the host when it sizes grids
#define SK_LUT_BYTES 65536                       // one table: [256 ks][64 columns] float
#define SK_MAX_K 224                             // keys a warp list can return (topk, or ranked lists for large nlist)
#define SK_DYN_SMEM (227 * 1024 - 64)

struct WarpTopk {
    u64 *keys;   // shared, this warp's buffer (>= cap keys)
    int cap, k;  // cap = next_pow2(k + 32): compaction threshold of the current pass
    int count;   // warp-uniform
    u64 *thr_w;  // shared [nw]: every warp's ceil(k/nw)-th smallest key (RII_KEY_MAX until it has that many)
    int nw, wid;
    // Lazy ids (posting-list scans).  A push carries the candidate's POSITION in the pass (flattened 64-row group << 6 | row
    // of the group); the id is looked up when the list is compacted -- all new entries at once, one memory latency per
    // compaction instead of one per push on the scanning warp's critical path (ncu: stall_long_sb at every warp_push call
    // site, profiles/r02_ncu_c5shape_*).  ids == nullptr: positions are ids already (linear scans, center ranking).
    const int *ids;
    const long long *s_off;  // CSR offset of every planned segment
    const int *s_gcum;       // inclusive prefix of the segments' 64-row groups
    int J;                   // planned segments
    int nres;                // entries [0, nres) carry ids (kept by the last compaction)
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

static __device__ __noinline__ void warp_sort_any(u64 *keys, int n, int lane)  // n <= 256; buffer holds >= next_pow2-ish 32*R slots
{
    if (n <= 64) warp_sort_buf<2>(keys, n, lane);
    else if (n <= 128) warp_sort_buf<4>(keys, n, lane);
    else warp_sort_buf<8>(keys, n, lane);
}

static __device__ __noinline__ void warp_compact(WarpTopk &w, u64 *cta_thr, int lane)
{
    const int n = w.count;  // <= cap <= 256
    __syncwarp();  // the keys were written by other lanes of this warp (warp_push): order them before the reads below
    if (w.ids) {  // positions -> ids for everything pushed since the last compaction
        for (int e = w.nres + lane; e < n; e += 32) {
            const u64 key = w.keys[e];
            const uint32_t pos = key_id(key);
            const int f = (int)(pos >> 6);
            int lo = 0, hi = w.J - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (w.s_gcum[mid] > f) hi = mid; else lo = mid + 1;
            }
            const int r = (f - (lo ? w.s_gcum[lo - 1] : 0)) * 64 + (int)(pos & 63u);
            w.keys[e] = (key & 0xffffffff00000000ull) | (u64)(uint32_t)__ldg(w.ids + w.s_off[lo] + r);
        }
        __syncwarp();
    }
    warp_sort_any(w.keys, n, lane);
    w.count = n < w.k ? n : w.k;
    w.nres = w.count;
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
static __device__ __noinline__ void warp_push(WarpTopk &w, u64 *cta_thr, int lane, float dist, uint32_t id, bool pre)
{
    const u64 thr = *reinterpret_cast<volatile u64 *>(cta_thr);
    const u64 key = pack_key(dist, id);
    // (with lazy ids the candidate's id is not known yet: everything up to and including the threshold's distance passes)
    const bool pass = pre && (w.ids ? (uint32_t)(key >> 32) <= (uint32_t)(thr >> 32) : key < thr);
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if (!bal) return;
    if (pass) w.keys[w.count + __popc(bal & ((1u << lane) - 1u))] = key;
    w.count += __popc(bal);
    if (w.count + 32 > w.cap) warp_compact(w, cta_thr, lane);
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

// bitonic sort of P (power of two) keys in shared memory by the whole CTA of NT threads
template <int NT>
__device__ __forceinline__ void cta_sort_smem(u64 *k, int P)
{
    __syncthreads();
    for (int kk = 2; kk <= P; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += NT) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const u64 x = k[i], y = k[ixj];
                    const bool up = (i & kk) == 0;
                    if ((x > y) == up) { k[i] = y; k[ixj] = x; }
                }
            }
            __syncthreads();
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
