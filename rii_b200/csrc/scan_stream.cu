// Translation unit of the streaming scan engine: instantiations of k_scan_stream32 and their launcher.
#include "launch.h"
#include "scan_stream.cuh"
#include "../../include/rii_b200.h"

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return rii_fail(RII_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + \
                                              ":" + std::to_string(__LINE__) + ")");                         \
    } while (0)

int stream_pick(int row_bytes, bool ivf, bool two_ctas, int capw, int w_eff, size_t pool_bytes, int *nw, size_t *smem)
{
    if (row_bytes == 64) {
        {   // small top-k / few lists: 10 warps with the tables low in shared memory
            const size_t b = stream_smem_bytes(ivf, 10, 4, ST_TB4, capw, w_eff, pool_bytes, 2);
            if (b && b <= SK_DYN_SMEM) { *nw = 10; *smem = b; return 4; }
        }
        const size_t b = stream_smem_bytes(ivf, 8, 4, ST_TB3, capw, w_eff, pool_bytes, 2);
        if (b && b <= SK_DYN_SMEM) { *nw = 8; *smem = b; return 3; }
        return 0;
    }
    if (ivf && two_ctas) {
        const size_t b = stream_smem_bytes(true, 6, 3, ST_TB2, capw, w_eff, pool_bytes);
        if (b && b <= 113 * 1024) { *nw = 6; *smem = b; return 2; }
    }
    const size_t b = stream_smem_bytes(ivf, 12, 4, ST_TB1, capw, w_eff, pool_bytes);
    if (b && b <= SK_DYN_SMEM) { *nw = 12; *smem = b; return 1; }
    return 0;
}

template <int NW, bool IVF, int R, int MINB, uint32_t TB, int H, bool K1>
static int launch_stream_k(SkewArgs a, int parts, int B, size_t smem, cudaStream_t st)
{
    auto kern = k_scan_stream32<NW, IVF, R, MINB, TB, H, K1>;
    static bool configured[64] = {false};  // function attributes are per device: once per (instantiation, device)
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_DYN_SMEM));
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    a.smem_bytes = (uint32_t)smem;
    kern<<<dim3(parts, B), NW * 32, smem, st>>>(a);
    rii_count_launch();
    CK(cudaGetLastError());
    return 0;
}

template <int NW, bool IVF, int R, int MINB, uint32_t TB, int H>
static int launch_stream_t(const SkewArgs &a, int parts, int B, size_t smem, cudaStream_t st)
{
    // topk == 1 in a launch that produces results (not the coarse-only ranking launch, whose lists hold w_eff keys)
    if (a.k == 1 && !(IVF && a.coarse_mode == 1)) return launch_stream_k<NW, IVF, R, MINB, TB, H, true>(a, parts, B, smem, st);
    return launch_stream_k<NW, IVF, R, MINB, TB, H, false>(a, parts, B, smem, st);
}

int launch_stream(int shape, bool ivf, const SkewArgs &a, int parts, int B, size_t smem, cudaStream_t st)
{
    if (shape == 4) return ivf ? launch_stream_t<10, true, 4, 1, ST_TB4, 2>(a, parts, B, smem, st)
                               : launch_stream_t<10, false, 4, 1, ST_TB4, 2>(a, parts, B, smem, st);
    if (shape == 3) return ivf ? launch_stream_t<8, true, 4, 1, ST_TB3, 2>(a, parts, B, smem, st)
                               : launch_stream_t<8, false, 4, 1, ST_TB3, 2>(a, parts, B, smem, st);
    if (shape == 2) return launch_stream_t<6, true, 3, 2, ST_TB2, 1>(a, parts, B, smem, st);
    if (ivf) return launch_stream_t<12, true, 4, 1, ST_TB1, 1>(a, parts, B, smem, st);
    return launch_stream_t<12, false, 4, 1, ST_TB1, 1>(a, parts, B, smem, st);
}
