// Radix sorts used OUTSIDE the scan hot path, through CUB (part of the CUDA toolkit): the stable (list, id) sort behind the
// device-side posting-list build, and the global-memory top-k / ranking fallbacks for shapes whose keys do not fit shared
// memory (large topk, many ranked lists).  Isolated in its own translation unit: CUB is slow to compile.
#include "launch.h"
#include "../../include/rii_b200.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_radix_sort.cuh>

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return rii_fail(RII_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + \
                                              ":" + std::to_string(__LINE__) + ")");                         \
    } while (0)

static int ensure_tmp(SortTmp *t, size_t bytes)
{
    if (bytes <= t->cap) return 0;
    if (t->p) cudaFree(t->p);
    t->p = nullptr;
    t->cap = 0;
    CK(cudaMalloc(&t->p, bytes + bytes / 4 + 256));
    t->cap = bytes + bytes / 4 + 256;
    return 0;
}

int dev_sort_pairs_u32(const uint32_t *keys_in, uint32_t *keys_out, const uint32_t *vals_in, uint32_t *vals_out, long long n,
                       int end_bit, SortTmp *tmp, cudaStream_t st)
{
    if (n <= 0) return 0;
    if (n >= (1ll << 31)) return rii_fail(RII_ERR_LIMIT, "sort of >= 2^31 items");
    size_t need = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit, st));
    if (int rc = ensure_tmp(tmp, need)) return rc;
    CK(cub::DeviceRadixSort::SortPairs(tmp->p, need, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit, st));
    rii_count_launch();
    return 0;
}

int dev_sort_keys_u64(const unsigned long long *in, unsigned long long *out, long long n, SortTmp *tmp, cudaStream_t st)
{
    if (n <= 0) return 0;
    if (n >= (1ll << 31)) return rii_fail(RII_ERR_LIMIT, "sort of >= 2^31 items");
    size_t need = 0;
    CK(cub::DeviceRadixSort::SortKeys(nullptr, need, in, out, (int)n, 0, 64, st));
    if (int rc = ensure_tmp(tmp, need)) return rc;
    CK(cub::DeviceRadixSort::SortKeys(tmp->p, need, in, out, (int)n, 0, 64, st));
    rii_count_launch();
    return 0;
}

int dev_segsort_keys_u64(const unsigned long long *in, unsigned long long *out, long long n, int nseg, const long long *d_seg_beg,
                         const long long *d_seg_end, SortTmp *tmp, cudaStream_t st)
{
    if (n <= 0 || nseg <= 0) return 0;
    if (n >= (1ll << 31)) return rii_fail(RII_ERR_LIMIT, "sort of >= 2^31 items");
    size_t need = 0;
    CK(cub::DeviceSegmentedRadixSort::SortKeys(nullptr, need, in, out, (int)n, nseg, d_seg_beg, d_seg_end, 0, 64, st));
    if (int rc = ensure_tmp(tmp, need)) return rc;
    CK(cub::DeviceSegmentedRadixSort::SortKeys(tmp->p, need, in, out, (int)n, nseg, d_seg_beg, d_seg_end, 0, 64, st));
    rii_count_launch();
    return 0;
}
