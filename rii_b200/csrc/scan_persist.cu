// Translation unit of the persistent, warp-specialised IVF batch kernel (scan_persist.cuh).
#include "launch.h"
#include "scan_persist.cuh"
#include "../../include/rii_b200.h"

#include <algorithm>

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return rii_fail(RII_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + \
                                              ":" + std::to_string(__LINE__) + ")");                         \
    } while (0)

bool persist_fits(int row_bytes, int topk, int w_eff, int nlist, bool fused_coarse)
{
    return row_bytes == 32 && topk <= PS_MAXK && w_eff <= PS_WMAX && (!fused_coarse || nlist <= PS_NLIST_MAX);
}

int launch_persist(const SkewArgs &a, int B, cudaStream_t st)
{
    static bool configured[64] = {false};
    static int sms[64] = {0};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return rii_fail(RII_ERR_LIMIT, "device ordinal >= 64");
    if (!configured[dev]) {
        CK(cudaFuncSetAttribute(k_scan_persist32<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS_SMEM_BYTES));
        CK(cudaFuncSetAttribute(k_scan_persist32<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CK(cudaFuncSetAttribute(k_scan_persist32<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS_SMEM_BYTES));
        CK(cudaFuncSetAttribute(k_scan_persist32<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CK(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev));
        configured[dev] = true;
    }
    const int grid = std::min(B, sms[dev]);
    if (a.k == 1) k_scan_persist32<true><<<grid, (PS_NC + 1) * 32, PS_SMEM_BYTES, st>>>(a, B);
    else k_scan_persist32<false><<<grid, (PS_NC + 1) * 32, PS_SMEM_BYTES, st>>>(a, B);
    rii_count_launch();
    CK(cudaGetLastError());
    return 0;
}
