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
// table at 0x3000); 3 = rows of 64 bytes (8 warps, 4-stage rings, two tables at 0x6000, one CTA per SM).
// Returns the shape that fits (0: none) and its warps / dynamic shared memory.
int stream_pick(int row_bytes, bool ivf, bool two_ctas, int capw, int w_eff, size_t pool_bytes, int *nw, size_t *smem);
int launch_stream(int shape, bool ivf, const SkewArgs &a, int parts, int B, size_t smem, cudaStream_t st);
