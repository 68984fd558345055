// Host-side launch entry points of the separately compiled kernel translation units (scan_stream.cu, assign_stream.cu,
// device_sort.cu).  Every function returns 0 or a negative RII_ERR_* code with the message left in rii_last_error().
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <string>

struct SkewArgs;

int rii_fail(int code, const std::string &msg);   // records the calling thread's error message; returns `code`
void rii_count_launch();                          // rii_launch_count() bookkeeping

// ---- scan_stream.cu: the skew64 streaming scan engine (k_scan_stream32) ------------------------------------------------
// shapes: 1 = one CTA per SM (12 warps, 4-stage rings, table at 0x10000); 2 = two CTAs per SM (6 warps, 3-stage rings,
// table at 0x3000); 3 = rows of 64 bytes (8 warps, 4-stage rings, two tables at 0x6000, one CTA per SM); 4 = the same
// with 10 warps and the tables at 0x2400 (small top-k).
// Returns the shape that fits (0: none) and its warps / dynamic shared memory.
int stream_pick(int row_bytes, bool ivf, bool two_ctas, int capw, int w_eff, size_t pool_bytes, int *nw, size_t *smem);
int launch_stream(int shape, bool ivf, const SkewArgs &a, int parts, int B, size_t smem, cudaStream_t st);

// ---- assign_stream.cu: K6 (nearest coarse center under the symmetric distance) on the streaming engine -----------------
struct AssignPlan {
    int shape;        // 1: 12 warps, one CTA per SM; 2: two CTAs of 6 warps per SM; 3: rows of 64 bytes (8 warps)
    int parts;        // row parts (gridDim.x)
    int groups, kc;   // center groups (gridDim.y) of kc centers each
    long long n_pad;  // rows of the per-group workspaces: best float (groups, n_pad), arg int (groups, n_pad)
};
AssignPlan assign_stream_plan(long long n, int K, int row_bytes, int shape_opt);
// skew: skew64 table (one segment) of the n rows, rows of 32 (M <= 32) or 64 (M <= 64) bytes, zero padded beyond M.
// d_bad: device int, must be 0 on entry; set to 1 when a Dm entry is too large for the packed accumulation (the results
// are then invalid and the caller falls back to the natural-layout kernel).
int launch_assign_stream(const AssignPlan &p, const float *Dm, const uint8_t *d_centers, int K, int M, int Ks, const uint8_t *skew,
                         long long n, float *ws_best, int *ws_arg, int *d_bad, int *d_assign, float *d_dist, cudaStream_t st);

// ---- device_sort.cu: CUB radix sorts (index build and the global-memory top-k fallbacks; not on the scan hot path) -----
struct SortTmp { void *p = nullptr; size_t cap = 0; };   // grow-only temporary storage, owned by the caller
// stable sort of (key, value) pairs by the low `end_bit` bits of the key
int dev_sort_pairs_u32(const uint32_t *keys_in, uint32_t *keys_out, const uint32_t *vals_in, uint32_t *vals_out, long long n,
                       int end_bit, SortTmp *tmp, cudaStream_t st);
int dev_sort_keys_u64(const unsigned long long *in, unsigned long long *out, long long n, SortTmp *tmp, cudaStream_t st);
// nseg segments [seg_beg[i], seg_end[i]) of u64 keys, each sorted ascending
int dev_segsort_keys_u64(const unsigned long long *in, unsigned long long *out, long long n, int nseg, const long long *d_seg_beg,
                         const long long *d_seg_end, SortTmp *tmp, cudaStream_t st);

// ---- scan_persist.cu: persistent warp-specialised IVF batch kernel (k_scan_persist32) ---------------------------------
bool persist_fits(int row_bytes, int topk, int w_eff, int nlist, bool fused_coarse);
int launch_persist(const SkewArgs &a, int B, cudaStream_t st);
