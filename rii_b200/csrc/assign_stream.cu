// Translation unit of the streaming assignment engine (K6): instantiations of k_assign_stream and their launcher.
#include "launch.h"
#include "assign_stream.cuh"
#include "../../include/rii_b200.h"

#include <algorithm>

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return rii_fail(RII_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + \
                                              ":" + std::to_string(__LINE__) + ")");                         \
    } while (0)

#define AS_TB 0x800u  // nothing lives below the table in this kernel

static int sm_count()
{
    static int n[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!n[dev] && cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n[dev] = 148;
    return n[dev];
}

// Split of the sweep over CTAs: `parts` row parts x `groups` center groups.  A table rebuild costs about as much as
// streaming 1500 rows, so a part should hold >= 16 K rows; the fewer center groups, the less work for the final fold.
AssignPlan assign_stream_plan(long long n, int K, int row_bytes, int shape_opt)
{
    AssignPlan p{};
    p.shape = row_bytes == 64 ? 3 : (shape_opt == 1 ? 1 : 2);
    const int ctas = sm_count() * (p.shape == 2 ? 2 : 1);
    const long long G = (n + 63) / 64;
    int groups = std::min(K, ctas), parts = 1;
    for (int g = 1; g <= std::min(K, ctas); ++g) {
        const int pr = (int)std::min<long long>(std::max<long long>(1, G), ctas / g);
        if (n / pr >= 16384 || pr == 1) { groups = g; parts = pr; break; }
    }
    if (parts == 1) groups = std::min(K, ctas);
    p.parts = parts;
    p.kc = (K + groups - 1) / groups;
    p.groups = (K + p.kc - 1) / p.kc;
    p.n_pad = G * 64;
    return p;
}

template <int NW, int R, int MINB, int H>
static int launch_t(AssignArgs a, const AssignPlan &p, cudaStream_t st)
{
    auto kern = k_assign_stream<NW, R, MINB, AS_TB, H>;
    const size_t smem = assign_smem_bytes(NW, R, AS_TB, H);
    static bool configured[64] = {false};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    a.smem_bytes = (uint32_t)smem;
    kern<<<dim3(p.parts, p.groups), NW * 32, smem, st>>>(a);
    rii_count_launch();
    CK(cudaGetLastError());
    return 0;
}

int launch_assign_stream(const AssignPlan &p, const float *Dm, const uint8_t *d_centers, int K, int M, int Ks, const uint8_t *skew,
                         long long n, float *ws_best, int *ws_arg, int *d_bad, int *d_assign, float *d_dist, cudaStream_t st)
{
    AssignArgs a{};
    a.Dm = Dm; a.centers = d_centers; a.K = K; a.M = M; a.Ks = Ks; a.skew = skew; a.n = n; a.kc = p.kc;
    a.best = ws_best; a.arg = ws_arg; a.n_pad = p.n_pad; a.bad = d_bad;
    int rc;
    if (p.shape == 3) rc = launch_t<8, 4, 1, 2>(a, p, st);
    else if (p.shape == 2) rc = launch_t<6, 3, 2, 1>(a, p, st);
    else rc = launch_t<12, 4, 1, 1>(a, p, st);
    if (rc < 0) return rc;
    k_assign_reduce<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws_best, ws_arg, p.groups, p.n_pad, n, d_assign, d_dist);
    rii_count_launch();
    CK(cudaGetLastError());
    return 0;
}
