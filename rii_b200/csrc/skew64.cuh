// The skew64 layout of code rows (see scan_stream.cuh): sizing helper shared by the host and the engines.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// 32-byte windows of a skew64 segment holding `len` code rows of M = 32 H bytes: H blocks of 64 windows per group of 64
// rows, plus H blocks after the last group (the first one holds the lagging tail of the last rows; with H = 2 the
// second keeps every segment an even number of blocks, so that a block's half-row index stays a compile-time constant
// of the pipeline stage)
static __host__ __device__ inline long long skew64_rows(long long len, int H = 1) { return 64 * ((len + 63) / 64 + 1) * H; }

