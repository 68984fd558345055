// librii_b200.so -- host side of the B200 ADC path + the C ABI declared in include/rii_b200.h.
//
// Mirrors the state and the entry points of the reference's `RiiCpp` (src/rii.h:39-83, src/main.cpp:12-54)
// with the data resident in HBM:
//   codes      (N, M)  uint8, row-major, rows M bytes apart (32-byte rows for M=32 -> one LDG.256 / sector)
//   codewords  (M, Ks, Ds) float32             coarse centers (nlist, M) uint8
//   posting lists as CSR: offsets int64 (nlist+1), ids int32 (N)   (the reference stores ids only, too)
//   Dm         (M, Ks, Ks) float32 codeword distance matrices (built on first use)
// The host keeps only what the reference's host glue needs (posting lists / centers mirrors and the two
// libstdc++ shuffles that define the reconfigure sampling, src/rii.h:120-124, src/pqkmeans.cpp:177-191).
#include "../../include/rii_b200.h"
#include "kernels.cuh"
#include "launch.h"
#include "skew64_build.cuh"
#include "build_kernels.cuh"
#include "warp_topk.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <random>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(RII_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ +  \
                                          ":" + std::to_string(__LINE__) + ")");                             \
    } while (0)
#define CKR(call)                     \
    do {                              \
        int r_ = (call);              \
        if (r_ < 0) return r_;        \
    } while (0)
#define LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

#define PS_CAPW_HOST 64
const size_t SMEM_MAX = 227 * 1024;  // opt-in shared memory per CTA on sm_100

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return fail(RII_ERR_CUDA, std::string("cudaMalloc scratch: ") + cudaGetErrorString(e));
        cap = want;
        return 0;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

int host_l2_variant()
{
    FILE *f = fopen("/proc/cpuinfo", "r");
    if (!f) return 16;
    std::string s;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0 && s.size() < (1u << 20)) s.append(buf, n);
    fclose(f);
    if (s.find(" avx512f") != std::string::npos) return 16;
    if (s.find(" avx ") != std::string::npos || s.find(" avx2") != std::string::npos) return 8;
    return 4;
}

}  // namespace

int rii_fail(int code, const std::string &msg) { return fail(code, msg); }
void rii_count_launch() { LAUNCHED(); }

enum { PK_DTABLE = 0, PK_SCAN_LINEAR, PK_MERGE, PK_COARSE, PK_SUBSET, PK_PLAN, PK_SCAN_IVF, PK_ASSIGN, PK_SORT, PK_N };
static const char *PK_NAMES[PK_N] = {"dtable", "scan_linear", "merge", "coarse_rank", "subset_build", "plan",
                                      "scan_ivf", "assign", "sort"};
struct ProfRec { int kind; cudaEvent_t a, b; };

struct rii_index {
    // optional per-kernel CUDA-event timing (rii_profile_*): events are recorded around each launch on the
    // launching stream and only read back (after a sync) when the totals are asked for.
    bool prof = false;
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[PK_N] = {0};
    long long prof_n[PK_N] = {0};
    int M = 0, Ks = 0, Ds = 0, variant = 16, verbose = 0, device = 0;
    int rb = 0;  // bytes of a (zero-padded) code row on the streaming engine: 32 (12 <= M <= 32), 64 (32 < M <= 64), 0: natural-layout kernels only
    long long N = 0, cap_rows = 0;      // local rows
    long long id_base = 0, N_total = -1;  // sharding (N_total < 0: single shard, N_total = N)
    int nlist = 0;
    cudaStream_t stream = nullptr;

    float *d_cw = nullptr;
    float *d_cw_t = nullptr;  // (256, M, Dp) copy of the codewords, sub-space fastest, Dp = 4 for Ds <= 4 (coalesced in-kernel table build)
    float *d_Dm = nullptr;
    uint8_t *d_codes = nullptr;
    DevBuf centers, offsets, ids, loc_len, glob_len, pre_len;
    // skew64 copies (scan_stream.cuh; M == 32, built lazily on the query stream): the codes by id, every local posting
    // list (segment i at physical row skew_off[i]) and the coarse centers
    DevBuf skew_lin, skew_lists, skew_off, centers_skew, skew_misc_off;
    long long skew_lin_rows = -1;  // rows covered by skew_lin (-1: stale)
    bool skew_lists_valid = false, centers_skew_valid = false;
    bool has_global = false;

    std::vector<uint8_t> h_centers;      // (nlist, M)
    std::vector<long long> h_offsets;    // (nlist+1)
    DevBuf assign;                       // (N) list of every row (kept for subset searches: sub-index build)
    long long N_assigned = 0;            // rows of `assign` that are valid
    DevBuf ws_best, ws_arg, d_flag;      // assignment engine workspaces; d_flag: [0] Dm max bits, [2] engine flag
    bool dm_ok = true;                   // every Dm entry is small enough for the packed accumulation
    SortTmp sort_tmp;
    bool shard_stale = false;            // rows were added to / removed from a shard: rii_set_shard must be called again
    std::vector<long long> len_sorted_prefix;  // prefix sums of ascending *global* list lengths

    // scratch (grow only)
    // derived layouts (skew64 copies) are built lazily on the stream of the query that first needs them; a later query on
    // ANOTHER stream must not read them before that build has finished: the build records this event, other streams wait on it
    cudaEvent_t built_ev = nullptr;
    cudaStream_t built_on = nullptr;
    bool built_pending = false;
    const float *q_host = nullptr;  // set by query_host around a single zero-copy call: the query also travels in the kernel parameters
    DevBuf merge_cnt;   // (B) per-query arrival counters of the in-kernel merge (zero between launches)
    DevBuf T, partial, ranked, cum, take_last, J, flags, filt, bitmap, q, tids, o_ids, o_dists, o_counts, tmp0, tmp1,
        tmp2, tmp3;
    // sub-index of a subset search (target_ids): (list, row) pairs, their sorted form = CSR of the members, skew64 copy
    DevBuf sub_keys, sub_rows, sub_keys_s, sub_rows_s, sub_bounds, sub_len, sub_off, sub_skew, sub_glob, sub_pre;
    long long sub_S = -1;     // two-phase sharded subset search: target count of the prepared sub-index (-1: none)
    // general (global-memory) path: candidate keys and their sorted copy, segment bounds, candidate counts
    DevBuf gen_keys, gen_sorted, gen_seg, gen_cnt;
    // batched re-run of flagged queries
    DevBuf redo_idx, redo_q, redo_ids, redo_d, redo_c;
    // small calls (the reference's one-query-per-call shape, src/main.cpp:17-27): queries and results travel through one
    // pinned, device-mapped host buffer -- the kernels read the query and write the results over PCIe themselves, so a call
    // is launches + one stream synchronisation, no cudaMemcpy
    void *pin = nullptr;
    size_t pin_cap = 0;
    // large host batches: chunked H2D copies on their own stream, overlapped with the search of the previous chunk
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> copy_events;
    // OPQ rotation applied on the device before the table build (rii/rii.py:305-306): R (D, D) row-major, or null
    float *d_R = nullptr;
    DevBuf qrot;

    DevBuf dbg;               // optional phase clocks of the v2 scan kernel ("debug_clocks" option)
    int opt_debug_clocks = 0;
    int opt_stream_ctas = 0;  // v4 engine: 1 = always one CTA per SM; otherwise per-query IVF batches run two CTAs per SM
    int opt_fuse_coarse = 1;  // fuse coarse ranking + plan into the v2 posting-list scan when one CTA serves a query
    int opt_zero_copy = 1;      // small host calls go through mapped pinned memory (no cudaMemcpy)
    int opt_persist = 1;        // 0 = never, 1 = auto (batches of >= 296 queries), 2 = whenever the shape fits (tests)
    int opt_assign_kernel = 0;  // 0 auto (streaming engine, two CTAs per SM), 1 natural-layout k_assign, 3 streaming engine with one CTA per SM
    int opt_scan_kernel = 0;  // 0 auto, 1 natural-layout kernels (v1), 2 skewed conflict-free kernel (v2), 3 dual-stream FFMA2
                              // skewed kernel (v3), 4 register-streaming kernel over the skew64 layout (v4; what auto picks
                              // when it applies).  v2 / v3 / v4: M == 32 only

    long long n_total() const { return N_total >= 0 ? N_total : N; }
};

namespace {

struct Prof {  // RAII: time one launch when profiling is on
    rii_index *h; cudaStream_t st; ProfRec r; bool on;
    Prof(rii_index *h_, cudaStream_t st_, int kind) : h(h_), st(st_), on(h_->prof)
    {
        if (!on) return;
        r.kind = kind;
        for (cudaEvent_t *e : {&r.a, &r.b}) {
            if (!h->prof_pool.empty()) { *e = h->prof_pool.back(); h->prof_pool.pop_back(); }
            else cudaEventCreate(e);
        }
        cudaEventRecord(r.a, st);
    }
    ~Prof()
    {
        if (!on) return;
        cudaEventRecord(r.b, st);
        h->prof_pending.push_back(r);
    }
};

void prof_collect(rii_index *h)
{
    for (auto &r : h->prof_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            h->prof_ms[r.kind] += ms;
            h->prof_n[r.kind] += 1;
        }
        h->prof_pool.push_back(r.a);
        h->prof_pool.push_back(r.b);
    }
    h->prof_pending.clear();
}

template <class F> int set_smem(F *kernel, size_t bytes)
{
    if (bytes > SMEM_MAX) return fail(RII_ERR_LIMIT, "shape needs " + std::to_string(bytes) + " B of shared memory per CTA (> 227 KB)");
    if (bytes > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

// ---- per-M dispatch (register-resident rows for the common sizes, byte loads otherwise) -------------
#define DISPATCH_M(M, ...)                                  \
    switch (M) {                                            \
    case 4: { constexpr int MT = 4; __VA_ARGS__; } break;   \
    case 8: { constexpr int MT = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int MT = 16; __VA_ARGS__; } break; \
    case 32: { constexpr int MT = 32; __VA_ARGS__; } break; \
    case 64: { constexpr int MT = 64; __VA_ARGS__; } break; \
    default: { constexpr int MT = 0; __VA_ARGS__; } break;  \
    }

int ensure_Dm(rii_index *h)
{
    if (h->d_Dm) return 0;
    const size_t nDm = (size_t)h->M * h->Ks * h->Ks;
    CK(cudaMalloc(&h->d_Dm, nDm * sizeof(float)));
    dim3 grid((h->Ks * h->Ks + RII_THREADS - 1) / RII_THREADS, h->M);
    k_symmat<<<grid, RII_THREADS, 0, h->stream>>>(h->d_cw, h->d_Dm, h->Ks, h->Ds);
    LAUNCHED();
    CK(cudaGetLastError());
    // the packed accumulation of the streaming engine needs finite partial sums (scan_stream.cuh ST_TABLE_LIMIT)
    CKR(h->d_flag.ensure(16));
    CK(cudaMemsetAsync(h->d_flag.p, 0, 16, h->stream));
    k_max_bits<<<296, 256, 0, h->stream>>>(h->d_Dm, (long long)nDm, h->d_flag.as<unsigned int>());
    LAUNCHED();
    unsigned int mx = 0;
    CK(cudaMemcpyAsync(&mx, h->d_flag.p, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float lim = 5e36f;
    unsigned int lim_bits;
    std::memcpy(&lim_bits, &lim, 4);
    h->dm_ok = mx <= lim_bits;
    return 0;
}

// Rows to assign: natural layout (n, M) and, for the streaming engine, their skew64 copy (one segment).
struct AssignSrc {
    const uint8_t *natural = nullptr;
    const uint8_t *skew = nullptr;  // null: natural-layout kernel only
    long long n = 0;
};

int note_build(rii_index *h, cudaStream_t st)  // a derived layout was (re)built on `st`
{
    if (!h->built_ev) CK(cudaEventCreateWithFlags(&h->built_ev, cudaEventDisableTiming));
    CK(cudaEventRecord(h->built_ev, st));
    h->built_on = st;
    h->built_pending = true;
    return 0;
}

int ensure_skew_lin(rii_index *h, cudaStream_t st);
int skew_build(const uint8_t *codes, const int *ids, const long long *offsets, const long long *d_skew_off, int nseg, long long n_single,
               long long prows, uint8_t *out, int M, int RB, cudaStream_t st);

// skew64 copy of n natural-layout rows into `buf` (reused across the iterations of a fit)
int make_assign_src(rii_index *h, const uint8_t *d_rows, long long n, DevBuf *buf, AssignSrc *out)
{
    out->natural = d_rows;
    out->n = n;
    out->skew = nullptr;
    CKR(ensure_Dm(h));
    if (!h->rb || !h->dm_ok || h->opt_assign_kernel == 1 || n <= 0) return 0;
    if (d_rows == h->d_codes && n == h->N) {  // the whole index: the linear scan's copy
        CKR(ensure_skew_lin(h, h->stream));
        out->skew = h->skew_lin.as<uint8_t>();
        return 0;
    }
    const long long prows = skew64_rows(n, h->rb / 32);
    CKR(buf->ensure((size_t)prows * 32 + 32));
    long long *off = reinterpret_cast<long long *>(buf->as<uint8_t>() + (size_t)prows * 32);
    const long long hoff[2] = {0, prows};
    CK(cudaMemcpyAsync(off, hoff, 16, cudaMemcpyHostToDevice, h->stream));
    CKR(skew_build(d_rows, nullptr, nullptr, off, 1, n, prows, buf->as<uint8_t>(), h->M, h->rb, h->stream));
    out->skew = buf->as<uint8_t>();
    return 0;
}

// K6 launcher: rows `src`, d_centers (K, M) device -> d_assign (n) [, d_dist (n)]
int launch_assign(rii_index *h, const AssignSrc &src, const uint8_t *d_centers, int K, int *d_assign, float *d_dist)
{
    const long long n = src.n;
    if (n == 0) return 0;
    CKR(ensure_Dm(h));
    Prof pr(h, h->stream, PK_ASSIGN);
    if (src.skew) {  // streaming engine (assign_stream.cuh)
        const AssignPlan pl = assign_stream_plan(n, K, h->rb, h->opt_assign_kernel == 3 ? 1 : 0);
        CKR(h->ws_best.ensure((size_t)pl.groups * pl.n_pad * 4));
        CKR(h->ws_arg.ensure((size_t)pl.groups * pl.n_pad * 4));
        CKR(h->d_flag.ensure(16));
        return launch_assign_stream(pl, h->d_Dm, d_centers, K, h->M, h->Ks, src.skew, n, h->ws_best.as<float>(), h->ws_arg.as<int>(),
                                    h->d_flag.as<int>() + 2, d_assign, d_dist, h->stream);
    }
    const uint8_t *d_codes = src.natural;
    const size_t lut1 = (size_t)h->M * h->Ks * sizeof(float);
    int G = 0, CPT = 4;
    for (int g : {4, 2, 1}) {
        if (lut1 * g + (size_t)RII_THREADS * 4 * h->M <= 200 * 1024) { G = g; CPT = 4; break; }
    }
    if (!G) {
        if (lut1 + (size_t)RII_THREADS * h->M <= 220 * 1024) { G = 1; CPT = 1; }
        else return fail(RII_ERR_LIMIT, "assignment kernel: M*Ks too large for shared memory");
    }
    const int tile = RII_THREADS * CPT;
    const size_t smem = lut1 * G + (size_t)tile * h->M;
    const unsigned grid = (unsigned)((n + tile - 1) / tile);
#define LAUNCH_ASSIGN(GG, CC)                                                                                        \
    do {                                                                                                             \
        CKR(set_smem(k_assign<GG, CC>, smem));                                                                       \
        k_assign<GG, CC><<<grid, RII_THREADS, smem, h->stream>>>(h->d_Dm, d_codes, n, d_centers, K, h->M, h->Ks,    \
                                                                  d_assign, d_dist);                                 \
    } while (0)
    if (G == 4) LAUNCH_ASSIGN(4, 4);
    else if (G == 2) LAUNCH_ASSIGN(2, 4);
    else if (CPT == 4) LAUNCH_ASSIGN(1, 4);
    else LAUNCH_ASSIGN(1, 1);
#undef LAUNCH_ASSIGN
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

// Publish the host-side offsets: device copies of the offsets and lengths, derived layouts invalidated.
int finish_lists(rii_index *h)
{
    const int nlist = h->nlist;
    CKR(h->offsets.ensure((size_t)(nlist + 1) * 8));
    CKR(h->loc_len.ensure((size_t)std::max(1, nlist) * 4));
    CK(cudaMemcpyAsync(h->offsets.p, h->h_offsets.data(), (size_t)(nlist + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    std::vector<int> len(nlist);
    for (int i = 0; i < nlist; ++i) len[i] = (int)(h->h_offsets[i + 1] - h->h_offsets[i]);
    if (nlist) CK(cudaMemcpyAsync(h->loc_len.p, len.data(), (size_t)nlist * 4, cudaMemcpyHostToDevice, h->stream));
    h->skew_lists_valid = false;  // derived copies are rebuilt lazily by the first query that needs them
    CK(cudaStreamSynchronize(h->stream));
    h->has_global = false;  // a shard must exchange its lengths again (rii_set_global_lengths)
    std::sort(len.begin(), len.end());
    h->len_sorted_prefix.assign(nlist + 1, 0);
    for (int i = 0; i < nlist; ++i) h->len_sorted_prefix[i + 1] = h->len_sorted_prefix[i] + len[i];
    return 0;
}

int grow_assign(rii_index *h, long long rows)
{
    if ((size_t)rows * 4 <= h->assign.cap) return 0;
    DevBuf nb;
    CKR(nb.ensure((size_t)std::max(rows, h->cap_rows) * 4));
    if (h->assign.p && h->N_assigned > 0)
        CK(cudaMemcpyAsync(nb.p, h->assign.p, (size_t)h->N_assigned * 4, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->assign.release();
    h->assign = nb;
    return 0;
}

// src/rii.h:335-359 UpdatePostingLists(start, num): assign rows [start, start + num) on the GPU and append their ids to
// the lists, all on the device: a stable radix sort of the (list, id) pairs keeps every list ascending in id (new ids
// are larger than every id already listed), the per-list segments are merged behind the old lists.
int update_posting_lists(rii_index *h, long long start, long long num, const int *d_given_assign = nullptr)
{
    if (num <= 0) return finish_lists(h);
    const int nlist = h->nlist;
    CKR(grow_assign(h, start + num));
    int *d_new = h->assign.as<int>() + start;
    if (d_given_assign) {  // lists from an external clustering: no K6
        CK(cudaMemcpyAsync(d_new, d_given_assign, (size_t)num * 4, cudaMemcpyDeviceToDevice, h->stream));
    } else {
        AssignSrc src;
        DevBuf tmp_skew;
        int rc = make_assign_src(h, h->d_codes + start * h->M, num, &tmp_skew, &src);
        if (rc == 0) rc = launch_assign(h, src, h->centers.as<uint8_t>(), nlist, d_new, nullptr);
        if (rc == 0 && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(RII_ERR_CUDA, "assignment failed");
        tmp_skew.release();
        CKR(rc);
    }
    h->N_assigned = start + num;
    // (list, id) pairs sorted by list
    CKR(h->tmp0.ensure((size_t)num * 4));  // ids in
    CKR(h->tmp1.ensure((size_t)num * 4));  // lists out
    CKR(h->tmp2.ensure((size_t)num * 4));  // ids out
    CKR(h->tmp3.ensure((size_t)(nlist + 2) * 8 * 2));  // bounds of the new ids per list | new offsets
    k_iota_u32<<<(unsigned)((num + 255) / 256), 256, 0, h->stream>>>(h->tmp0.as<uint32_t>(), num, (uint32_t)start);
    LAUNCHED();
    int bits = 1;
    while ((1ll << bits) <= nlist) ++bits;  // keys 0..nlist-1 and the invalid marker -1 (all ones) sort correctly on `bits` bits
    CKR(dev_sort_pairs_u32(reinterpret_cast<const uint32_t *>(d_new), h->tmp1.as<uint32_t>(), h->tmp0.as<uint32_t>(), h->tmp2.as<uint32_t>(),
                           num, bits, &h->sort_tmp, h->stream));
    long long *d_bounds = h->tmp3.as<long long>(), *d_noff = d_bounds + nlist + 2;
    k_list_bounds<<<(nlist + 1 + 255) / 256, 256, 0, h->stream>>>(h->tmp1.as<uint32_t>(), num, nlist, d_bounds);
    LAUNCHED();
    std::vector<long long> bounds(nlist + 1);
    CK(cudaMemcpyAsync(bounds.data(), d_bounds, (size_t)(nlist + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (bounds[nlist] != num) return fail(RII_ERR_CUDA, "assignment kernel produced an invalid list id");
    std::vector<long long> noff(nlist + 1, 0);
    for (int i = 0; i < nlist; ++i) noff[i + 1] = noff[i] + (h->h_offsets[i + 1] - h->h_offsets[i]) + (bounds[i + 1] - bounds[i]);
    const long long old_total = h->h_offsets[nlist];
    DevBuf nids;
    CKR(nids.ensure(std::max<size_t>(4, (size_t)noff[nlist] * 4)));
    CK(cudaMemcpyAsync(d_noff, noff.data(), (size_t)(nlist + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    k_merge_lists<<<nlist, 256, 0, h->stream>>>(old_total ? h->offsets.as<long long>() : nullptr, h->ids.as<int>(), d_bounds,
                                               h->tmp2.as<uint32_t>(), d_noff, nids.as<int>());
    LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    h->ids.release();
    h->ids = nids;
    h->h_offsets.swap(noff);
    return finish_lists(h);
}

int set_centers(rii_index *h, const uint8_t *centers_host, int nlist)
{
    h->nlist = nlist;
    h->centers_skew_valid = false;
    h->h_centers.assign(centers_host, centers_host + (size_t)nlist * h->M);
    CKR(h->centers.ensure((size_t)nlist * h->M));
    CK(cudaMemcpyAsync(h->centers.p, h->h_centers.data(), (size_t)nlist * h->M, cudaMemcpyHostToDevice, h->stream));
    h->h_offsets.assign(nlist + 1, 0);
    h->N_assigned = 0;
    h->has_global = false;
    return 0;
}

// src/pqkmeans.cpp:46-133 on a device-resident sample (already in the reference's shuffled order).
// Leaves the final centers in tmp2 (device) and copies them to centers_out (host).
int fit_coarse(rii_index *h, const uint8_t *d_sample, long long ns, int nlist, int iter, uint8_t *centers_out)
{
    const int M = h->M, Ks = h->Ks;
    if (nlist <= 0 || (long long)nlist > ns) return fail(RII_ERR_ARG, "fit_coarse: need 0 < nlist <= number of sample codes");
    CKR(ensure_Dm(h));
    // InitializeCentersByRandomPicking, src/pqkmeans.cpp:177-191
    std::vector<int> ids((size_t)ns);
    std::iota(ids.begin(), ids.end(), 0);
    std::mt19937 random_engine(0);
    std::shuffle(ids.begin(), ids.end(), random_engine);
    std::vector<long long> pick(nlist);
    for (int k = 0; k < nlist; ++k) pick[k] = ids[k];
    CKR(h->tmp1.ensure((size_t)nlist * 8));
    CKR(h->tmp2.ensure((size_t)nlist * M));       // centers_new
    CKR(h->tmp3.ensure((size_t)nlist * M));       // centers_old
    CK(cudaMemcpyAsync(h->tmp1.p, pick.data(), (size_t)nlist * 8, cudaMemcpyHostToDevice, h->stream));
    {
        long long tot = (long long)nlist * M;
        k_gather_rows<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(d_sample, h->tmp1.as<long long>(), nlist, M,
                                                                          h->tmp2.as<uint8_t>());
        LAUNCHED();
        CK(cudaGetLastError());
    }
    DevBuf assign, hist, skew_tmp;
    AssignSrc src;
    int rc = 0;
    do {
        if ((rc = make_assign_src(h, d_sample, ns, &skew_tmp, &src)) < 0) break;
        if ((rc = assign.ensure((size_t)ns * 4)) < 0) break;
        if (iter > 1 && (rc = hist.ensure((size_t)nlist * M * Ks * 4)) < 0) break;
        for (int itr = 0; itr < iter; ++itr) {
            if (h->verbose) printf("Iteration start: %d / %d\n", itr, iter);
            cudaError_t ce = cudaMemcpyAsync(h->tmp3.p, h->tmp2.p, (size_t)nlist * M, cudaMemcpyDeviceToDevice, h->stream);
            if (ce != cudaSuccess) { rc = fail(RII_ERR_CUDA, std::string("fit_coarse: ") + cudaGetErrorString(ce)); break; }
            if ((rc = launch_assign(h, src, h->tmp3.as<uint8_t>(), nlist, assign.as<int>(), nullptr)) < 0) break;
            if (itr != iter - 1) {  // src/pqkmeans.cpp:110
                ce = cudaMemsetAsync(hist.p, 0, (size_t)nlist * M * Ks * 4, h->stream);
                long long tot = ns * M;
                k_vote_hist<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(d_sample, assign.as<int>(), ns, M, Ks,
                                                                                hist.as<int>());
                LAUNCHED();
                if (ce == cudaSuccess) ce = cudaGetLastError();
                k_vote_centers<<<dim3(nlist, M), 256, 0, h->stream>>>(h->d_Dm, hist.as<int>(), M, Ks, h->tmp2.as<uint8_t>());
                LAUNCHED();
                if (ce == cudaSuccess) ce = cudaGetLastError();
                if (ce != cudaSuccess) { rc = fail(RII_ERR_CUDA, std::string("fit_coarse: ") + cudaGetErrorString(ce)); break; }
            }
        }
        if (rc < 0) break;
        cudaError_t e = cudaMemcpyAsync(centers_out, h->tmp2.p, (size_t)nlist * M, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(RII_ERR_CUDA, std::string("fit_coarse: ") + cudaGetErrorString(e));
    } while (0);
    if (rc < 0) cudaStreamSynchronize(h->stream);
    assign.release();
    hist.release();
    skew_tmp.release();
    return rc;
}

int grow_codes(rii_index *h, long long rows)
{
    if (rows <= h->cap_rows) return 0;
    long long ncap = std::max(rows, h->cap_rows + h->cap_rows / 2);
    uint8_t *nb = nullptr;
    CK(cudaMalloc(&nb, (size_t)ncap * h->M + 64));
    if (h->N) CK(cudaMemcpyAsync(nb, h->d_codes, (size_t)h->N * h->M, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->d_codes) cudaFree(h->d_codes);
    h->d_codes = nb;
    h->cap_rows = ncap;
    return 0;
}

// ---- derived code layouts (M == 32), built lazily on the query stream ------------------------------------
int skew_build(const uint8_t *codes, const int *ids, const long long *offsets, const long long *d_skew_off, int nseg, long long n_single,
               long long prows, uint8_t *out, int M, int RB, cudaStream_t st)
{
    if (prows <= 0) return 0;
    const long long thr = prows * 2;
    k_skew64_build<<<(unsigned)((thr + 255) / 256), 256, 0, st>>>(codes, ids, offsets, d_skew_off, nseg, n_single, 0, prows, out, M, RB);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

int ensure_skew_lin(rii_index *h, cudaStream_t st)  // skew64 of the codes by id (linear scan, scan_stream.cuh)
{
    if (h->skew_lin_rows == h->N) return 0;
    const int H = h->rb / 32;
    const long long prows = skew64_rows(h->N, H);
    // rows appended since the last build only change the windows from the first incomplete group on: keep the rest
    long long keep = h->skew_lin_rows > 0 && h->skew_lin_rows < h->N ? (h->skew_lin_rows / 64) * H * 64 : 0;
    if ((size_t)prows * 32 > h->skew_lin.cap) {
        DevBuf nb;
        CKR(nb.ensure((size_t)prows * 32));  // (DevBuf over-allocates by 25 %: amortises a stream of adds)
        if (keep) CK(cudaMemcpyAsync(nb.p, h->skew_lin.p, (size_t)keep * 32, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        h->skew_lin.release();
        h->skew_lin = nb;
    }
    CKR(h->skew_misc_off.ensure(32));
    const long long off[2] = {0, prows};
    CK(cudaMemcpyAsync(h->skew_misc_off.p, off, 16, cudaMemcpyHostToDevice, st));
    const long long thr = (prows - keep) * 2;
    k_skew64_build<<<(unsigned)((thr + 255) / 256), 256, 0, st>>>(h->d_codes, nullptr, nullptr, h->skew_misc_off.as<long long>(), 1, h->N,
                                                                  keep, prows, h->skew_lin.as<uint8_t>(), h->M, h->rb);
    LAUNCHED();
    CK(cudaGetLastError());
    h->skew_lin_rows = h->N;
    return note_build(h, st);
}

int ensure_centers_skew(rii_index *h, cudaStream_t st)  // skew64 of the coarse centers (fused coarse pass)
{
    if (h->centers_skew_valid) return 0;
    const long long prows = skew64_rows(h->nlist, h->rb / 32);
    CKR(h->centers_skew.ensure((size_t)prows * 32));
    CKR(h->skew_misc_off.ensure(32));
    const long long off[2] = {0, prows};
    CK(cudaMemcpyAsync(h->skew_misc_off.as<long long>() + 2, off, 16, cudaMemcpyHostToDevice, st));
    CKR(skew_build(h->centers.as<uint8_t>(), nullptr, nullptr, h->skew_misc_off.as<long long>() + 2, 1, h->nlist, prows,
                   h->centers_skew.as<uint8_t>(), h->M, h->rb, st));
    h->centers_skew_valid = true;
    return note_build(h, st);
}

int ensure_skew_lists(rii_index *h, cudaStream_t st)  // skew64 of every local posting list
{
    if (h->skew_lists_valid) return 0;
    const int nlist = h->nlist;
    std::vector<long long> off((size_t)nlist + 1, 0);
    for (int i = 0; i < nlist; ++i) off[i + 1] = off[i] + skew64_rows(h->h_offsets[i + 1] - h->h_offsets[i], h->rb / 32);
    CKR(h->skew_off.ensure((size_t)(nlist + 1) * 8));
    CKR(h->skew_lists.ensure((size_t)std::max<long long>(1, off[nlist]) * 32));
    CK(cudaMemcpyAsync(h->skew_off.p, off.data(), (size_t)(nlist + 1) * 8, cudaMemcpyHostToDevice, st));
    CKR(skew_build(h->d_codes, h->ids.as<int>(), h->offsets.as<long long>(), h->skew_off.as<long long>(), nlist, 0, off[nlist],
                   h->skew_lists.as<uint8_t>(), h->M, h->rb, st));
    h->skew_lists_valid = true;
    return note_build(h, st);
}

// ---- the query pipeline on device buffers -----------------------------------------------------------
// What a posting-list scan walks: the index's own lists, or the temporary sub-index of a subset search.
struct ListsView {
    const long long *offsets = nullptr;  // device CSR (nlist + 1)
    const int *ids = nullptr;            // device: local row ids grouped by list, ascending per list
    const int *loc_len = nullptr, *glob_len = nullptr, *pre_len = nullptr;  // device (nlist); pre_len null: single shard
    const uint8_t *skew = nullptr;       // skew64 segment per list (streaming engine), or null
    const long long *skew_off = nullptr;
    long long cap_local = 0;             // upper bound of the ids listed locally
    long long skew_bytes = 0;            // size of the skew64 table (decides whether it can be L2-resident)
};

int make_tables(rii_index *h, const float *d_Q, int B, cudaStream_t st)  // K1 for the natural-layout / general kernels
{
    const int lutf = h->M * h->Ks;
    CKR(h->T.ensure((size_t)B * lutf * 4));
    Prof pr(h, st, PK_DTABLE);
    dim3 grid((lutf + RII_THREADS - 1) / RII_THREADS, B);
    k_dtable<<<grid, RII_THREADS, 0, st>>>(d_Q, h->d_cw, h->T.as<float>(), h->M, h->Ks, h->Ds, h->variant);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

// Partial lists of `parts` CTAs per query: merged by the last CTA itself when parts * topk keys fit one warp sort (returns
// true: no merge launch needed), else by k_merge.
int prepare_partial(rii_index *h, int B, int parts, int topk, TopkOut *out, cudaStream_t st, bool *in_kernel)
{
    CKR(h->partial.ensure((size_t)B * parts * topk * 8));
    out->partial = h->partial.as<u64>();
    out->merge_cnt = nullptr;
    *in_kernel = (long long)parts * topk <= 256;
    if (*in_kernel) {
        if ((size_t)B * 4 > h->merge_cnt.cap) {
            CKR(h->merge_cnt.ensure((size_t)std::max(B, 4096) * 4));
            CK(cudaMemsetAsync(h->merge_cnt.p, 0, h->merge_cnt.cap, st));  // (the kernels leave the counters at zero)
        }
        out->merge_cnt = h->merge_cnt.as<int>();
    }
    return 0;
}

int launch_merge(rii_index *h, int B, int parts, int topk, TopkOut out, cudaStream_t st)
{
    const int mcap = next_pow2(topk + RII_THREADS);
    const size_t msmem = scan_smem_bytes(0, mcap, 0);
    CKR(set_smem(k_merge, msmem));
    Prof pr(h, st, PK_MERGE);
    k_merge<<<B, RII_THREADS, msmem, st>>>(h->partial.as<u64>(), parts, topk, mcap, out);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

// sort nseg segments of `stride` slots each (the first count[b] / n of them are keys): long segments one by one with
// the device-wide sort, short ones with the segmented sort (one CTA per segment)
int sort_segments(rii_index *h, int nseg, long long stride, const int *d_count, long long n, cudaStream_t st)
{
    Prof pr(h, st, PK_SORT);
    if (stride >= 32768 || nseg == 1) {
        if (d_count) {  // sort the whole slot range: unused slots hold RII_KEY_MAX and stay at the end
            for (int b = 0; b < nseg; ++b)
                CKR(dev_sort_keys_u64(h->gen_keys.as<u64>() + (size_t)b * stride, h->gen_sorted.as<u64>() + (size_t)b * stride, stride, &h->sort_tmp, st));
        } else {
            for (int b = 0; b < nseg; ++b)
                CKR(dev_sort_keys_u64(h->gen_keys.as<u64>() + (size_t)b * stride, h->gen_sorted.as<u64>() + (size_t)b * stride, n, &h->sort_tmp, st));
        }
        return 0;
    }
    CKR(h->gen_seg.ensure((size_t)nseg * 16));
    long long *beg = h->gen_seg.as<long long>(), *end = beg + nseg;
    k_seg_bounds<<<(nseg + 255) / 256, 256, 0, st>>>(nseg, stride, d_count, n, beg, end);
    LAUNCHED();
    return dev_segsort_keys_u64(h->gen_keys.as<u64>(), h->gen_sorted.as<u64>(), (long long)nseg * stride, nseg, beg, end, &h->sort_tmp, st);
}

// General linear scan: every candidate's (dist, id) key goes to HBM, a radix sort orders them (any M, any topk).
int general_linear(rii_index *h, const float *d_Q, int B, int topk, const long long *d_tids, long long S, long long *d_out_ids,
                   float *d_out_dists, int *d_out_counts, cudaStream_t st)
{
    const int M = h->M, Ks = h->Ks, lutf = M * Ks;
    const long long ncand = S ? S : h->N;
    if ((size_t)lutf * 4 > SMEM_MAX) return fail(RII_ERR_LIMIT, "M * Ks * 4 bytes exceed the shared memory of a CTA");
    const int Bc = (int)std::max<long long>(1, std::min<long long>(B, (1ll << 27) / std::max<long long>(1, ncand)));
    CKR(h->gen_keys.ensure((size_t)Bc * ncand * 8));
    CKR(h->gen_sorted.ensure((size_t)Bc * ncand * 8));
    for (int b0 = 0; b0 < B; b0 += Bc) {
        const int bc = std::min(Bc, B - b0);
        CKR(make_tables(h, d_Q + (size_t)b0 * M * h->Ds, bc, st));
        LinearArgs a{};
        a.T = h->T.as<float>(); a.codes = h->d_codes; a.tids = S ? d_tids : nullptr; a.S = S; a.N = h->N; a.id_base = h->id_base;
        a.M = M; a.Ks = Ks;
        const unsigned gx = (unsigned)std::min<long long>(1184, (ncand + RII_THREADS - 1) / RII_THREADS);
        {
            Prof pr(h, st, PK_SCAN_LINEAR);
            DISPATCH_M(M, {
                CKR(set_smem(k_keys_linear<MT>, (size_t)lutf * 4));
                k_keys_linear<MT><<<dim3(gx, bc), RII_THREADS, (size_t)lutf * 4, st>>>(a, h->gen_keys.as<u64>(), ncand);
            });
            LAUNCHED();
            CK(cudaGetLastError());
        }
        CKR(sort_segments(h, bc, ncand, nullptr, ncand, st));
        TopkOut out{};
        out.out_ids = d_out_ids + (size_t)b0 * topk; out.out_dists = d_out_dists + (size_t)b0 * topk; out.out_counts = d_out_counts + b0;
        out.id_base = h->id_base; out.final = 1;
        k_take_sorted<<<dim3((topk + 255) / 256, bc), 256, 0, st>>>(h->gen_sorted.as<u64>(), ncand, nullptr, ncand, topk, out);
        LAUNCHED();
        CK(cudaGetLastError());
    }
    return 0;
}

// K2/K3 (src/rii.h:195-242).  tids_state: 1 = ascending and inside this index, 0 = not, -1 = unknown (checked on the device)
int run_linear(rii_index *h, const float *d_Q, int B, int topk, const long long *d_tids, long long S, int tids_state,
               long long *d_out_ids, float *d_out_dists, int *d_out_counts, cudaStream_t st)
{
    const int M = h->M, Ks = h->Ks, lutf = M * Ks;
    const long long ncand = S ? S : h->N;
    TopkOut out{};
    out.out_ids = d_out_ids; out.out_dists = d_out_dists; out.out_counts = d_out_counts; out.id_base = h->id_base;
    // ---- streaming engine: the skew64 copy of the codes, or of the target rows (subset) ----
    const int capw = std::max(64, next_pow2(topk + 32));
    int nw = 0, shape = 0;
    size_t smem4 = 0;
    if (h->rb && topk <= SK_MAX_K && h->opt_scan_kernel != 1) shape = stream_pick(h->rb, false, false, capw, 0, 0, &nw, &smem4);
    bool use4 = shape > 0 && (h->opt_scan_kernel == 4 || (S == 0 ? h->N >= 32768 : ncand * B >= (1ll << 22)));
    long long i0 = 0, cnt = S;
    if (use4 && S) {
        if (tids_state < 0 || h->N_total >= 0) {  // ascending?  which run of it lies in this shard?
            CKR(h->d_flag.ensure(32));
            int *fl = h->d_flag.as<int>() + 4;
            CK(cudaMemsetAsync(fl, 0, 12, st));
            k_tids_scan<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(d_tids, S, h->id_base, h->id_base + h->N, fl);
            LAUNCHED();
            int hf[3];
            CK(cudaMemcpyAsync(hf, fl, 12, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (tids_state < 0) tids_state = (hf[0] == 0 && (h->N_total >= 0 || hf[1] + hf[2] == 0)) ? 1 : 0;
            i0 = hf[1];
            cnt = S - hf[1] - hf[2];
        }
        if (tids_state != 1) use4 = false;
    }
    if (h->opt_scan_kernel == 4 && !use4) return fail(RII_ERR_LIMIT, "scan_kernel=4 needs 12 <= M <= 64, topk <= 224 and ascending target_ids");
    if (use4) {
        SkewArgs sa{};
        sa.Q = d_Q; sa.cw = h->d_cw; sa.cw_t = h->d_cw_t; sa.Ds = h->Ds; sa.variant = h->variant; sa.M = M;
        sa.Ks = Ks; sa.k = topk; sa.cap = capw;
        if (h->q_host && B == 1 && M * h->Ds <= 128 && h->Ds <= 4 && !h->d_R) {
            sa.q_inline = 1;
            std::memcpy(sa.qv, h->q_host, (size_t)M * h->Ds * 4);
        }
        if (S) {  // compact skew64 copy of the target rows (in the given order; repeated ids stay repeated, src/rii.h:222-227)
            if (cnt <= 0) {  // nothing of the subset lives in this shard
                CK(cudaMemsetAsync(d_out_counts, 0, (size_t)B * 4, st));
                return 0;
            }
            Prof pr(h, st, PK_SUBSET);
            const long long prows = skew64_rows(cnt, h->rb / 32);
            CKR(h->sub_rows.ensure((size_t)cnt * 4));
            CKR(h->sub_skew.ensure((size_t)prows * 32));
            CKR(h->sub_off.ensure(32));
            k_tids_to_rows<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(d_tids + i0, cnt, h->id_base, h->sub_rows.as<int>());
            LAUNCHED();
            const long long hoff[2] = {0, prows};
            CK(cudaMemcpyAsync(h->sub_off.p, hoff, 16, cudaMemcpyHostToDevice, st));
            CKR(skew_build(h->d_codes, h->sub_rows.as<int>(), nullptr, h->sub_off.as<long long>(), 1, cnt, prows, h->sub_skew.as<uint8_t>(),
                           M, h->rb, st));
            sa.codes = h->sub_skew.as<uint8_t>();
            sa.N = cnt;
            out.id_map = d_tids + i0;
        } else {
            CKR(ensure_skew_lin(h, st));
            sa.codes = h->skew_lin.as<uint8_t>();
            sa.N = h->N;
        }
        // CTAs per query: fewer queries than SMs -> each is spread over 148 / B CTAs.  (For B > 148 a wave-aware split -- 256 queries
        // on 148 SMs are 2 waves for 1.73 waves of work -- was measured: 4 CTAs per query made the C3 linear leg 9 % SLOWER, the
        // per-CTA table build, pipeline fill and partial-list merge cost more than the emptier last wave.)
        const int parts = (int)std::min<long long>(std::max(1, 148 / std::min(B, 148)), std::max<long long>(1, sa.N / (nw * SK_TILE_ROWS * 4)));
        out.final = parts == 1;
        bool in_kernel = false;
        if (!out.final) CKR(prepare_partial(h, B, parts, topk, &out, st, &in_kernel));
        sa.out = out;
        if (h->opt_debug_clocks) { CKR(h->dbg.ensure((size_t)parts * B * 64)); sa.dbg = h->dbg.as<long long>(); }
        {
            Prof pr(h, st, PK_SCAN_LINEAR);
            CKR(launch_stream(shape, false, sa, parts, B, smem4, st));
        }
        if (!out.final && !in_kernel) CKR(launch_merge(h, B, parts, topk, out, st));
        return 0;
    }
    // ---- natural-layout kernel: keys of the CTA in shared memory ----
    const int cap = next_pow2(topk + RII_THREADS * RII_ROWS_PER_THREAD);
    const size_t smem = scan_smem_bytes(lutf, cap, 0);
    if (smem > SMEM_MAX) return general_linear(h, d_Q, B, topk, d_tids, S, d_out_ids, d_out_dists, d_out_counts, st);
    const int max_parts = std::max(1, 592 / B);
    const long long per_cta = B >= 148 ? 8192 : 1024;
    const int parts = (int)std::min<long long>(max_parts, std::max<long long>(1, (ncand + per_cta - 1) / per_cta));
    out.final = parts == 1;
    if (!out.final) {
        CKR(h->partial.ensure((size_t)B * parts * topk * 8));
        out.partial = h->partial.as<u64>();
    }
    CKR(make_tables(h, d_Q, B, st));
    LinearArgs a{};
    a.T = h->T.as<float>(); a.codes = h->d_codes; a.tids = S ? d_tids : nullptr; a.S = S; a.N = h->N; a.id_base = h->id_base;
    a.M = M; a.Ks = Ks; a.k = topk; a.cap = cap; a.out = out;
    {
        Prof pr(h, st, PK_SCAN_LINEAR);
        DISPATCH_M(M, {
            CKR(set_smem(k_scan_linear<MT>, smem));
            k_scan_linear<MT><<<dim3(parts, B), RII_THREADS, smem, st>>>(a);
        });
        LAUNCHED();
        CK(cudaGetLastError());
    }
    if (!out.final) CKR(launch_merge(h, B, parts, topk, out, st));
    return 0;
}

int ensure_centers_skew(rii_index *h, cudaStream_t st);
int ensure_skew_lists(rii_index *h, cudaStream_t st);

int main_view(rii_index *h, cudaStream_t st, ListsView *v)
{
    if (h->rb && h->opt_scan_kernel != 1 && h->h_offsets.back() > 0) {
        CKR(ensure_skew_lists(h, st));
        v->skew = h->skew_lists.as<uint8_t>();
        v->skew_off = h->skew_off.as<long long>();
        v->skew_bytes = (h->h_offsets.back() + 64ll * h->nlist) * h->rb;
    }
    v->offsets = h->offsets.as<long long>();
    v->ids = h->ids.as<int>();
    v->loc_len = h->loc_len.as<int>();
    v->glob_len = h->has_global ? h->glob_len.as<int>() : h->loc_len.as<int>();
    v->pre_len = h->has_global ? h->pre_len.as<int>() : nullptr;
    v->cap_local = h->h_offsets.back();
    return 0;
}

// Sub-index of a subset search (src/rii.h:294): the members of every posting list, ascending in id.  Device only, no
// host round trip: the sizes are bounded by S.
int build_subview(rii_index *h, const long long *d_tids, long long S, cudaStream_t st, ListsView *v)
{
    if (h->N_assigned != h->N)
        return fail(RII_ERR_STATE, "subset search: codes were added without updating the posting lists (reconfigure or add with update)");
    Prof pr(h, st, PK_SUBSET);
    const int nlist = h->nlist, H = std::max(1, h->rb / 32);
    CKR(h->sub_keys.ensure((size_t)S * 4));
    CKR(h->sub_rows.ensure((size_t)S * 4));
    CKR(h->sub_keys_s.ensure((size_t)S * 4));
    CKR(h->sub_rows_s.ensure((size_t)S * 4));
    CKR(h->sub_bounds.ensure((size_t)(nlist + 2) * 8));
    CKR(h->sub_len.ensure((size_t)nlist * 4));
    CKR(h->sub_off.ensure((size_t)(nlist + 2) * 8));
    k_sub_keys<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(d_tids, S, h->id_base, h->N, h->assign.as<int>(), h->sub_keys.as<uint32_t>(),
                                                            h->sub_rows.as<uint32_t>());
    LAUNCHED();
    int bits = 1;
    while ((1ll << bits) <= nlist) ++bits;
    CKR(dev_sort_pairs_u32(h->sub_keys.as<uint32_t>(), h->sub_keys_s.as<uint32_t>(), h->sub_rows.as<uint32_t>(), h->sub_rows_s.as<uint32_t>(), S,
                           bits, &h->sort_tmp, st));
    k_list_bounds<<<(nlist + 1 + 255) / 256, 256, 0, st>>>(h->sub_keys_s.as<uint32_t>(), S, nlist, h->sub_bounds.as<long long>());
    LAUNCHED();
    k_sub_layout<<<1, 1024, 0, st>>>(h->sub_bounds.as<long long>(), nlist, H, h->sub_len.as<int>(), h->sub_off.as<long long>());
    LAUNCHED();
    CK(cudaGetLastError());
    v->offsets = h->sub_bounds.as<long long>();
    v->ids = h->sub_rows_s.as<int>();
    v->loc_len = v->glob_len = h->sub_len.as<int>();
    v->pre_len = nullptr;
    v->cap_local = std::min<long long>(S, h->N);
    if (h->rb && h->opt_scan_kernel != 1) {
        const long long cap_rows = (v->cap_local + 128ll * nlist) * H;  // every list: its groups + one drain group + rounding
        CKR(h->sub_skew.ensure((size_t)cap_rows * 32));
        CKR(skew_build(h->d_codes, v->ids, v->offsets, h->sub_off.as<long long>(), nlist, 0, cap_rows, h->sub_skew.as<uint8_t>(), h->M, h->rb, st));
        v->skew = h->sub_skew.as<uint8_t>();
        v->skew_off = h->sub_off.as<long long>();
        v->skew_bytes = cap_rows * 32;
    }
    return 0;
}

// General posting-list search through HBM: coarse keys -> sort -> plan -> candidate keys -> sort -> top-k.
int general_ivf(rii_index *h, const float *d_Q, int B, int topk, long long L, const ListsView &v, int w, int w_eff, long long *d_out_ids,
                float *d_out_dists, int *d_out_counts, cudaStream_t st)
{
    const int M = h->M, Ks = h->Ks, lutf = M * Ks, nlist = h->nlist;
    if ((size_t)lutf * 4 > SMEM_MAX) return fail(RII_ERR_LIMIT, "M * Ks * 4 bytes exceed the shared memory of a CTA");
    const long long lcap = std::max<long long>(1, std::min<long long>(L, v.cap_local));
    const long long stride = std::max<long long>(nlist, lcap);
    const int Bc = (int)std::max<long long>(1, std::min<long long>(B, (1ll << 27) / stride));
    CKR(h->gen_keys.ensure((size_t)Bc * stride * 8));
    CKR(h->gen_sorted.ensure((size_t)Bc * stride * 8));
    CKR(h->gen_cnt.ensure((size_t)Bc * 4));
    CKR(h->filt.ensure((size_t)Bc * w_eff * 4 * 3));
    int *f = h->filt.as<int>(), *pre = f + (size_t)Bc * w_eff, *loc = pre + (size_t)Bc * w_eff;
    for (int b0 = 0; b0 < B; b0 += Bc) {
        const int bc = std::min(Bc, B - b0);
        CKR(make_tables(h, d_Q + (size_t)b0 * M * h->Ds, bc, st));
        {
            Prof pr(h, st, PK_COARSE);
            const unsigned gx = (unsigned)std::min(296, (nlist + RII_THREADS - 1) / RII_THREADS);
            DISPATCH_M(M, {
                CKR(set_smem(k_coarse_keys<MT>, (size_t)lutf * 4));
                k_coarse_keys<MT><<<dim3(gx, bc), RII_THREADS, (size_t)lutf * 4, st>>>(h->T.as<float>(), h->centers.as<uint8_t>(), nlist, M, Ks,
                                                                                     h->gen_keys.as<u64>());
            });
            LAUNCHED();
            CK(cudaGetLastError());
        }
        CKR(sort_segments(h, bc, nlist, nullptr, nlist, st));
        PlanArgs p{};
        p.glob_len = v.glob_len; p.pre_len = v.pre_len; p.loc_len = v.loc_len;
        p.filt_cnt = f; p.filt_pre = v.pre_len ? pre : nullptr; p.filt_loc = loc;
        p.L = L; p.topk = topk; p.w = w; p.w_eff = w_eff; p.nlist = nlist;
        p.ranked = h->ranked.as<int>() + (size_t)b0 * w_eff; p.cum = h->cum.as<int>() + (size_t)b0 * w_eff;
        p.take_last = h->take_last.as<int>() + b0; p.J = h->J.as<int>() + b0; p.flags = h->flags.as<int>() + b0;
        {
            Prof pr(h, st, PK_PLAN);
            k_rank_gather<<<dim3((w_eff + 255) / 256, bc), 256, 0, st>>>(h->gen_sorted.as<u64>(), nlist, w_eff, v.glob_len, v.pre_len, v.loc_len, p.ranked,
                                                                        f, pre, loc);
            LAUNCHED();
            k_plan<<<(bc + 127) / 128, 128, 0, st>>>(p, bc);
            LAUNCHED();
            CK(cudaGetLastError());
        }
        IvfArgs a{};
        a.T = h->T.as<float>(); a.codes = h->d_codes; a.offsets = v.offsets; a.ids = v.ids;
        a.ranked = p.ranked; a.cum = p.cum; a.J = p.J; a.flags = p.flags; a.take_last = p.take_last; a.w_eff = w_eff;
        a.M = M; a.Ks = Ks; a.k = topk;
        {
            Prof pr(h, st, PK_SCAN_IVF);
            CK(cudaMemsetAsync(h->gen_keys.p, 0xff, (size_t)bc * lcap * 8, st));  // unused slots = RII_KEY_MAX
            const unsigned gx = (unsigned)std::min<long long>(1184, (lcap + RII_THREADS - 1) / RII_THREADS);
            DISPATCH_M(M, {
                CKR(set_smem(k_ivf_keys<MT>, (size_t)lutf * 4));
                k_ivf_keys<MT><<<dim3(gx, bc), RII_THREADS, (size_t)lutf * 4, st>>>(a, h->gen_keys.as<u64>(), lcap, h->gen_cnt.as<int>());
            });
            LAUNCHED();
            CK(cudaGetLastError());
        }
        CKR(sort_segments(h, bc, lcap, h->gen_cnt.as<int>(), lcap, st));
        TopkOut out{};
        out.out_ids = d_out_ids + (size_t)b0 * topk; out.out_dists = d_out_dists + (size_t)b0 * topk; out.out_counts = d_out_counts + b0;
        out.id_base = h->id_base; out.final = 1;
        k_take_sorted<<<dim3((topk + 255) / 256, bc), 256, 0, st>>>(h->gen_sorted.as<u64>(), lcap, h->gen_cnt.as<int>(), lcap, topk, out);
        LAUNCHED();
        CK(cudaGetLastError());
    }
    return 0;
}

// K4 + K5 (src/rii.h:244-326) over `v`.  mode 0: the whole search; 1: coarse ranking only -> d_ranked (B, w_eff);
// 2: scan with the given ranking d_ranked.  The per-query plan flags are left in h->flags (B).
int run_ivf(rii_index *h, const float *d_Q, int B, int topk, long long L, const ListsView &v, int w, int w_eff, long long *d_out_ids,
            float *d_out_dists, int *d_out_counts, cudaStream_t st, int mode = 0, int *d_ranked = nullptr)
{
    const int M = h->M, Ks = h->Ks, lutf = M * Ks, nlist = h->nlist;
    CKR(h->ranked.ensure((size_t)B * w_eff * 4));
    CKR(h->cum.ensure((size_t)B * w_eff * 4));
    CKR(h->take_last.ensure((size_t)B * 4));
    CKR(h->J.ensure((size_t)B * 4));
    CKR(h->flags.ensure((size_t)B * 4));
    PlanArgs p{};
    p.glob_len = v.glob_len; p.pre_len = v.pre_len; p.loc_len = v.loc_len;
    p.L = L; p.topk = topk; p.w = w; p.w_eff = w_eff; p.nlist = nlist;
    p.ranked = d_ranked ? d_ranked : h->ranked.as<int>();
    p.cum = h->cum.as<int>(); p.take_last = h->take_last.as<int>(); p.J = h->J.as<int>(); p.flags = h->flags.as<int>();
    TopkOut out{};
    out.out_ids = d_out_ids; out.out_dists = d_out_dists; out.out_counts = d_out_counts; out.id_base = h->id_base;

    // ---- streaming engine.  nlist <= 1024: every coarse distance stays in shared memory (any w_eff); larger nlist: the
    // warps' top-k lists rank the centers and must hold max(topk, w_eff) keys
    const bool big_nlist = nlist > 1024;
    const int capw = std::max(64, next_pow2((big_nlist ? std::max(topk, w_eff) : topk) + 32));
    const size_t pool = !big_nlist ? (size_t)nlist * 4 : 0;
    bool use4 = h->rb && v.skew && topk <= SK_MAX_K && (!big_nlist || w_eff <= SK_MAX_K) && h->opt_scan_kernel != 1;
    int nw = 0, shape = 0, nw_c = 0, shape_c = 0;
    size_t smem4 = 0, smem_c = 0;
    const bool two = B >= 148 && h->opt_stream_ctas != 1;
    if (use4) {
        shape_c = stream_pick(h->rb, true, two, capw, w_eff, pool, &nw_c, &smem_c);  // with the coarse pass
        shape = stream_pick(h->rb, true, two, std::max(64, next_pow2(topk + 32)), w_eff, 0, &nw, &smem4);  // scan only
        use4 = shape > 0 && shape_c > 0;
    }
    if (h->opt_scan_kernel == 4 && !use4) return fail(RII_ERR_LIMIT, "scan_kernel=4 (ivf) needs 12 <= M <= 64, topk <= 224 and a list plan that fits shared memory");
    if (!use4 && mode != 0) return fail(RII_ERR_LIMIT, "the split coarse / scan phases need the streaming engine (12 <= M <= 64, topk <= 224)");
    if (use4) {
        int parts = mode == 1 ? 1 : (int)std::min<long long>(std::max(1, 148 / std::min(B, 148)),
                                                            std::max<long long>(1, (L + nw * SK_TILE_ROWS - 1) / (nw * SK_TILE_ROWS)));
        // one launch does everything when a CTA serves a whole query -- or, with nlist <= 1024 (every CTA of a query repeats the
        // cheap coarse pass + selection + plan and scans its share), for any number of CTAs per query: a single query is then
        // ONE kernel launch (table, ranking, plan, scan, merge by the last CTA)
        const bool fuse = mode == 0 && h->opt_fuse_coarse && (parts == 1 || (!big_nlist && (long long)parts * topk <= 256));
        SkewArgs sa{};
        sa.Q = d_Q; sa.cw = h->d_cw; sa.cw_t = h->d_cw_t; sa.Ds = h->Ds; sa.variant = h->variant; sa.M = M;
        sa.codes = v.skew; sa.offsets = v.offsets; sa.ids = v.ids; sa.skew_off = v.skew_off;
        sa.w_eff = w_eff; sa.Ks = Ks; sa.k = topk; sa.nlist = nlist; sa.plan = p;
        if (h->q_host && B == 1 && M * h->Ds <= 128 && h->Ds <= 4 && !h->d_R) {
            sa.q_inline = 1;
            std::memcpy(sa.qv, h->q_host, (size_t)M * h->Ds * 4);
        }
        if (fuse || mode == 1 || mode == 0) CKR(ensure_centers_skew(h, st));
        // batches of independent queries: the persistent warp-specialised kernel (scan_persist.cuh) when the shape fits
        const bool pers_shape = persist_fits(h->rb, topk, w_eff, nlist, mode == 0) && (mode == 2 || (mode == 0 && h->opt_fuse_coarse));
        if (mode != 1 && pers_shape && (h->opt_persist == 2 || (h->opt_persist == 1 && B >= 2 * 148))) {
            SkewArgs sp_ = sa;
            if (h->opt_debug_clocks) { CKR(h->dbg.ensure((size_t)std::max(B, 148) * 128)); CK(cudaMemsetAsync(h->dbg.p, 0, (size_t)148 * 128, st)); sp_.dbg = h->dbg.as<long long>(); }
            sp_.centers = mode == 0 ? h->centers_skew.as<uint8_t>() : nullptr;
            sp_.coarse_mode = mode == 0 ? 0 : 2;
            sp_.cap = PS_CAPW_HOST;
            out.final = 1;
            sp_.out = out;
            Prof pr(h, st, PK_SCAN_IVF);
            return launch_persist(sp_, B, st);
        }
        if (fuse) {
            sa.centers = h->centers_skew.as<uint8_t>(); sa.coarse_lists = big_nlist ? 1 : 0; sa.coarse_mode = 0;
            sa.cap = capw; out.final = parts == 1;
            bool in_kernel = true;
            if (!out.final) CKR(prepare_partial(h, B, parts, topk, &out, st, &in_kernel));
            sa.out = out;
            if (h->opt_debug_clocks) { CKR(h->dbg.ensure((size_t)B * parts * 64)); sa.dbg = h->dbg.as<long long>(); }
            Prof pr(h, st, PK_SCAN_IVF);
            return launch_stream(shape_c, true, sa, parts, B, smem_c, st);
        }
        if (mode != 2) {  // coarse pass on its own: one CTA per query ranks the lists
            SkewArgs sc = sa;
            sc.centers = h->centers_skew.as<uint8_t>(); sc.coarse_lists = big_nlist ? 1 : 0; sc.coarse_mode = 1; sc.cap = capw;
            Prof pr(h, st, PK_COARSE);
            CKR(launch_stream(shape_c, true, sc, 1, B, smem_c, st));
            if (mode == 1) return 0;
        }
        sa.centers = nullptr; sa.coarse_mode = 2; sa.cap = std::max(64, next_pow2(topk + 32));
        out.final = parts == 1;
        bool in_kernel = false;
        if (!out.final) CKR(prepare_partial(h, B, parts, topk, &out, st, &in_kernel));
        sa.out = out;
        if (h->opt_debug_clocks) { CKR(h->dbg.ensure((size_t)parts * B * 64)); sa.dbg = h->dbg.as<long long>(); }
        {
            Prof pr(h, st, PK_SCAN_IVF);
            CKR(launch_stream(shape, true, sa, parts, B, smem4, st));
        }
        if (!out.final && !in_kernel) CKR(launch_merge(h, B, parts, topk, out, st));
        return 0;
    }
    // ---- natural-layout kernels: coarse keys and candidate keys of a CTA in shared memory ----
    const int cap = next_pow2(topk + RII_THREADS * RII_ROWS_PER_THREAD);
    const int ccap = next_pow2(w_eff + RII_THREADS);
    const size_t smem_coarse = scan_smem_bytes(lutf, ccap, (size_t)w_eff * 12), smem_scan = scan_smem_bytes(lutf, cap, (size_t)w_eff * 8);
    if (smem_coarse > SMEM_MAX || smem_scan > SMEM_MAX)
        return general_ivf(h, d_Q, B, topk, L, v, w, w_eff, d_out_ids, d_out_dists, d_out_counts, st);
    {
        CoarseArgs a{};
        a.T = nullptr; a.Q = d_Q; a.cw = h->d_cw; a.Ds = h->Ds; a.variant = h->variant;
        a.centers = h->centers.as<uint8_t>();
        a.M = M; a.Ks = Ks; a.nlist = nlist; a.cap = ccap; a.do_plan = 1; a.plan = p;
        Prof pr(h, st, PK_COARSE);
        DISPATCH_M(M, {
            CKR(set_smem(k_coarse_rank<MT>, smem_coarse));
            k_coarse_rank<MT><<<B, RII_THREADS, smem_coarse, st>>>(a);
        });
        LAUNCHED();
        CK(cudaGetLastError());
    }
    const int max_parts = std::max(1, 592 / B);
    const long long per_cta = B >= 148 ? 8192 : 1024;
    const int parts = (int)std::min<long long>(max_parts, std::max<long long>(1, (L + per_cta - 1) / per_cta));
    out.final = parts == 1;
    if (!out.final) {
        CKR(h->partial.ensure((size_t)B * parts * topk * 8));
        out.partial = h->partial.as<u64>();
    }
    CKR(make_tables(h, d_Q, B, st));
    IvfArgs a{};
    a.T = h->T.as<float>(); a.codes = h->d_codes; a.offsets = v.offsets; a.ids = v.ids;
    a.ranked = p.ranked; a.cum = p.cum; a.J = p.J; a.flags = p.flags; a.take_last = p.take_last; a.w_eff = w_eff;
    a.M = M; a.Ks = Ks; a.k = topk; a.cap = cap; a.out = out;
    {
        Prof pr(h, st, PK_SCAN_IVF);
        DISPATCH_M(M, {
            CKR(set_smem(k_scan_ivf<MT>, smem_scan));
            k_scan_ivf<MT><<<dim3(parts, B), RII_THREADS, smem_scan, st>>>(a);
        });
        LAUNCHED();
        CK(cudaGetLastError());
    }
    if (!out.final) CKR(launch_merge(h, B, parts, topk, out, st));
    return 0;
}

// src/rii.h:267-277
int ivf_width(const rii_index *h, long long S, long long L)
{
    size_t ww = (size_t)std::round((double)L * h->nlist / (double)(S == 0 ? h->n_total() : S));
    ww += 3;
    if ((size_t)h->nlist < ww) ww = h->nlist;
    return (int)ww;
}

// IVF over `v` for B queries, in chunks; queries whose plan comes back flagged 1 (fewer than topk candidates in the
// first w lists: the reference walks on through ALL lists, src/rii.h:309-322) are gathered and re-run as one
// batch with the full ranking.  d_flags_out (optional): the final flags (bit 1: empty result).
int ivf_batches(rii_index *h, const float *d_Q, int B, int topk, long long L, const ListsView &v, int w, bool may_flag,
                long long *d_out_ids, float *d_out_dists, int *d_out_counts, cudaStream_t st)
{
    const int D = h->M * h->Ds;
    const int CH = 32768;  // queries per launch: long grids keep the tail (last partial wave of CTAs) small
    for (int b0 = 0; b0 < B; b0 += CH) {
        const int bc = std::min(CH, B - b0);
        CKR(run_ivf(h, d_Q + (size_t)b0 * D, bc, topk, L, v, w, w, d_out_ids + (size_t)b0 * topk, d_out_dists + (size_t)b0 * topk,
                    d_out_counts + b0, st));
        if (!may_flag || w >= h->nlist) continue;
        std::vector<int> flags(bc);
        CK(cudaMemcpyAsync(flags.data(), h->flags.p, (size_t)bc * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        std::vector<int> redo;
        for (int i = 0; i < bc; ++i)
            if (flags[i] & 1) redo.push_back(b0 + i);
        if (redo.empty()) continue;
        const int nr = (int)redo.size();
        CKR(h->redo_idx.ensure((size_t)nr * 4));
        CKR(h->redo_q.ensure((size_t)nr * D * 4));
        CKR(h->redo_ids.ensure((size_t)nr * topk * 8));
        CKR(h->redo_d.ensure((size_t)nr * topk * 4));
        CKR(h->redo_c.ensure((size_t)nr * 4));
        CK(cudaMemcpyAsync(h->redo_idx.p, redo.data(), (size_t)nr * 4, cudaMemcpyHostToDevice, st));
        k_gather_queries<<<(unsigned)(((long long)nr * D + 255) / 256), 256, 0, st>>>(d_Q, h->redo_idx.as<int>(), nr, D, h->redo_q.as<float>());
        LAUNCHED();
        CK(cudaStreamSynchronize(st));  // (redo is a host vector read by the async copy above)
        const int RC = 2048;
        for (int r0 = 0; r0 < nr; r0 += RC) {
            const int rc = std::min(RC, nr - r0);
            CKR(run_ivf(h, h->redo_q.as<float>() + (size_t)r0 * D, rc, topk, L, v, w, h->nlist, h->redo_ids.as<long long>() + (size_t)r0 * topk,
                        h->redo_d.as<float>() + (size_t)r0 * topk, h->redo_c.as<int>() + r0, st));
        }
        k_scatter_results<<<(unsigned)(((long long)nr * topk + 255) / 256), 256, 0, st>>>(h->redo_idx.as<int>(), nr, topk, h->redo_ids.as<long long>(),
                                                                                          h->redo_d.as<float>(), h->redo_c.as<int>(), d_out_ids,
                                                                                          d_out_dists, d_out_counts);
        LAUNCHED();
        CK(cudaGetLastError());
    }
    return 0;
}

int query_dev(rii_index *h, const float *d_Q, int B, int topk, const long long *d_tids, long long S, long long L,
              int method, long long *d_out_ids, float *d_out_dists, int *d_out_counts, cudaStream_t st, int tids_state = -1)
{
    if (B <= 0) return 0;
    if (h->N <= 0 && h->n_total() <= 0) return fail(RII_ERR_STATE, "query on an empty index");
    if (topk < 1) return fail(RII_ERR_ARG, "topk must be >= 1");
    if (h->shard_stale) return fail(RII_ERR_STATE, "codes were added to / cleared from this shard: call rii_set_shard (and rii_set_global_lengths) again");
    if (h->built_pending && h->built_on != st) CK(cudaStreamWaitEvent(st, h->built_ev, 0));  // (lazy builds of an earlier query on another stream)
    const long long Ntot = h->n_total();
    if (S < 0 || S > Ntot) return fail(RII_ERR_ARG, "need 0 <= len(target_ids) <= N");            // src/rii.h:220
    if ((long long)topk > (S ? S : Ntot)) return fail(RII_ERR_ARG, "need topk <= N (and topk <= len(target_ids))");  // :200,:219
    if (S > 0 && !d_tids) return fail(RII_ERR_ARG, "target_ids is null but S > 0");
    const int D = h->M * h->Ds;
    if (h->d_R) {  // OPQ: rotate the queries first (rii/rii.py:305-306, fine_quantizer.rotate)
        CKR(h->qrot.ensure((size_t)B * D * 4));
        k_rotate<<<dim3((D + 127) / 128, B), 128, 0, st>>>(d_Q, h->d_R, D, h->qrot.as<float>());
        LAUNCHED();
        CK(cudaGetLastError());
        d_Q = h->qrot.as<float>();
    }
    if (method == RII_METHOD_LINEAR) {
        if (h->N <= 0) {  // an empty shard of a larger index
            CK(cudaMemsetAsync(d_out_counts, 0, (size_t)B * 4, st));
            return 0;
        }
        const int CH = 32768;
        for (int b0 = 0; b0 < B; b0 += CH) {
            const int bc = std::min(CH, B - b0);
            CKR(run_linear(h, d_Q + (size_t)b0 * D, bc, topk, d_tids, S, tids_state, d_out_ids + (size_t)b0 * topk, d_out_dists + (size_t)b0 * topk,
                           d_out_counts + b0, st));
        }
        return 0;
    }
    if (method != RII_METHOD_IVF) return fail(RII_ERR_ARG, "unknown method");
    if (h->nlist <= 0) return fail(RII_ERR_STATE, "query_ivf before reconfigure(): no posting lists");
    if (h->N_total >= 0 && h->N_total != h->N && !h->has_global)
        return fail(RII_ERR_STATE, "IVF query on a shard whose list lengths were not exchanged: call rii_set_global_lengths");
    if (!(topk <= L && L <= Ntot)) return fail(RII_ERR_ARG, "need topk <= L <= N");           // src/rii.h:251
    const int w = ivf_width(h, S, L);
    ListsView v;
    bool may_flag;
    if (S != 0) {
        if (h->has_global)
            return fail(RII_ERR_STATE, "IVF + target_ids on a shard needs the member counts of the other shards: use rii_subset_begin_dev / "
                                       "rii_subset_set_global_dev / rii_subset_query_dev (rii_b200.sharded.sharded_query_subset)");
        CKR(build_subview(h, d_tids, S, st, &v));
        may_flag = true;
    } else {
        CKR(main_view(h, st, &v));
        // can the first w ranked lists hold fewer than topk candidates?  (only then the walk continues beyond w)
        may_flag = w < h->nlist && h->len_sorted_prefix.size() > (size_t)w && h->len_sorted_prefix[w] < topk && h->len_sorted_prefix[w] < L;
    }
    return ivf_batches(h, d_Q, B, topk, L, v, w, may_flag, d_out_ids, d_out_dists, d_out_counts, st);
}

int query_host(rii_index *h, const float *Q, int B, int topk, const int64_t *tids, int64_t S, int64_t L, int method,
               int64_t *out_ids, float *out_dists, int32_t *out_counts)
{
    if (B <= 0) return 0;
    if (!Q || !out_ids || !out_dists || !out_counts) return fail(RII_ERR_ARG, "null buffer");
    if (S > 0 && !tids) return fail(RII_ERR_ARG, "target_ids is null but S > 0");
    CK(cudaSetDevice(h->device));
    const int D = h->M * h->Ds;
    cudaStream_t st = h->stream;
    if (B <= 16 && S == 0 && topk <= 1024 && h->opt_zero_copy) {  // low-latency path: zero-copy through mapped pinned memory
        const size_t oq = 0, oi = ((size_t)B * D * 4 + 15) & ~(size_t)15, od = oi + (size_t)B * topk * 8, oc = od + (size_t)B * topk * 4;
        const size_t need = oc + (size_t)B * 4;
        if (need > h->pin_cap) {
            if (h->pin) cudaFreeHost(h->pin);
            h->pin = nullptr;
            h->pin_cap = 0;
            CK(cudaHostAlloc(&h->pin, need + 4096, cudaHostAllocMapped));
            h->pin_cap = need + 4096;
        }
        char *hp = (char *)h->pin, *dp = nullptr;
        CK(cudaHostGetDevicePointer((void **)&dp, h->pin, 0));
        std::memcpy(hp + oq, Q, (size_t)B * D * 4);
        std::memset(hp + oc, 0, (size_t)B * 4);
        h->q_host = B == 1 ? Q : nullptr;
        const int rc_ = query_dev(h, (const float *)(dp + oq), B, topk, nullptr, 0, L, method, (long long *)(dp + oi), (float *)(dp + od), (int *)(dp + oc), st, -1);
        h->q_host = nullptr;
        CKR(rc_);
        CK(cudaStreamSynchronize(st));
        std::memcpy(out_ids, hp + oi, (size_t)B * topk * 8);
        std::memcpy(out_dists, hp + od, (size_t)B * topk * 4);
        std::memcpy(out_counts, hp + oc, (size_t)B * 4);
        return 0;
    }
    CKR(h->q.ensure((size_t)B * D * 4));
    CKR(h->o_ids.ensure((size_t)B * topk * 8));
    CKR(h->o_dists.ensure((size_t)B * topk * 4));
    CKR(h->o_counts.ensure((size_t)B * 4));
    const int HC = 8192;  // queries per pipelined chunk of a large batch
    if (S == 0 && B >= 2 * HC) {
        // large batch: the queries travel in chunks on a copy stream while the previous chunk is searched (the H2D copy of a
        // 32768-query step is 4.5 % of its search time when it is not overlapped); results come back in one copy at the end
        if (!h->copy_stream) CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        // a small first chunk so that the search starts early, the rest in one piece (copied while the first chunk is searched)
        const int c0 = 2048;
        const int nch = 2;
        while ((int)h->copy_events.size() < nch) {
            cudaEvent_t ev;
            CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            h->copy_events.push_back(ev);
        }
        for (int c = 0; c < nch; ++c) {
            const size_t o = c ? (size_t)c0 : 0, n = c ? (size_t)(B - c0) : (size_t)c0;
            CK(cudaMemcpyAsync(h->q.as<float>() + o * D, Q + o * D, n * D * 4, cudaMemcpyHostToDevice, h->copy_stream));
            CK(cudaEventRecord(h->copy_events[c], h->copy_stream));
        }
        int rc = 0;
        for (int c = 0; c < nch && rc == 0; ++c) {
            const size_t o = c ? (size_t)c0 : 0;
            const int n = c ? B - c0 : c0;
            if (cudaStreamWaitEvent(st, h->copy_events[c], 0) != cudaSuccess) { rc = fail(RII_ERR_CUDA, "cudaStreamWaitEvent failed"); break; }
            rc = query_dev(h, h->q.as<float>() + o * D, n, topk, nullptr, 0, L, method, h->o_ids.as<long long>() + o * topk,
                           h->o_dists.as<float>() + o * topk, h->o_counts.as<int>() + o, st, -1);
        }
        if (rc < 0) {  // the copy stream may still be writing h->q: drain it before anyone reuses the buffer
            cudaStreamSynchronize(h->copy_stream);
            return rc;
        }
        CK(cudaMemcpyAsync(out_ids, h->o_ids.p, (size_t)B * topk * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(out_dists, h->o_dists.p, (size_t)B * topk * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(out_counts, h->o_counts.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return 0;
    }
    CK(cudaMemcpyAsync(h->q.p, Q, (size_t)B * D * 4, cudaMemcpyHostToDevice, st));
    int tids_state = -1;
    if (S > 0) {
        // one pass over the host ids: the reference indexes codes[tid] unchecked (src/rii.h:225), we refuse ids outside the
        // index; ascending ids (what Rii.query passes, rii/rii.py:296) take the streaming engine
        bool asc = true;
        for (int64_t i = 0; i < S; ++i) {
            if (h->N_total < 0 && (tids[i] < 0 || tids[i] >= h->N)) return fail(RII_ERR_ARG, "target_ids contains an id outside [0, N)");
            if (i && tids[i - 1] > tids[i]) asc = false;
        }
        tids_state = asc ? 1 : 0;
        CKR(h->tids.ensure((size_t)S * 8));
        CK(cudaMemcpyAsync(h->tids.p, tids, (size_t)S * 8, cudaMemcpyHostToDevice, st));
    }
    CKR(query_dev(h, h->q.as<float>(), B, topk, h->tids.as<long long>(), S, L, method, h->o_ids.as<long long>(),
                  h->o_dists.as<float>(), h->o_counts.as<int>(), st, tids_state));
    CK(cudaMemcpyAsync(out_ids, h->o_ids.p, (size_t)B * topk * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_dists, h->o_dists.p, (size_t)B * topk * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_counts, h->o_counts.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace

// =====================================================================================================
extern "C" {

const char *rii_last_error(void) { return g_err.c_str(); }
const char *rii_version(void) { return "0.2.12+b200.1"; }
int64_t rii_launch_count(void) { return g_launches.load(); }

int rii_create(const float *codewords, int M, int Ks, int Ds, int verbose, int device, int l2_variant, rii_index_t **out)
{
    if (!codewords || !out) return fail(RII_ERR_ARG, "null argument");
    if (M <= 0 || Ks <= 0 || Ds <= 0) return fail(RII_ERR_ARG, "codewords must have shape (M, Ks, Ds) with positive sizes");
    if (Ks > 256) return fail(RII_ERR_ARG, "Ks must be <= 256 so that each code is one uint8 (rii/rii.py:35)");
    if (l2_variant == 0) l2_variant = host_l2_variant();
    if (l2_variant != 16 && l2_variant != 8 && l2_variant != 4) return fail(RII_ERR_ARG, "l2_variant must be 0, 4, 8 or 16");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(RII_ERR_ARG, "no such CUDA device");
    CK(cudaSetDevice(device));
    rii_index *h = new rii_index();
    h->M = M; h->Ks = Ks; h->Ds = Ds; h->rb = M >= 12 && M <= 32 ? 32 : (M > 32 && M <= 64 ? 64 : 0); h->verbose = verbose; h->device = device; h->variant = l2_variant;
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_cw, (size_t)M * Ks * Ds * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_cw, codewords, (size_t)M * Ks * Ds * sizeof(float), cudaMemcpyHostToDevice);
    {   // sub-space-fastest copy for the in-kernel table build: (256, M, Dp), Dp = 4 for Ds <= 4 (zero padded: 16-byte loads), rows >= Ks zero
        const int Dp = Ds <= 4 ? 4 : Ds;
        std::vector<float> t((size_t)256 * M * Dp, 0.f);
        for (int m = 0; m < M; ++m)
            for (int ks = 0; ks < Ks; ++ks)
                for (int i = 0; i < Ds; ++i) t[((size_t)ks * M + m) * Dp + i] = codewords[((size_t)m * Ks + ks) * Ds + i];
        if (e == cudaSuccess) e = cudaMalloc(&h->d_cw_t, t.size() * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(h->d_cw_t, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        delete h;
        return fail(RII_ERR_CUDA, std::string("rii_create: ") + cudaGetErrorString(e));
    }
    h->h_offsets.assign(1, 0);
    if (verbose) printf("rii_b200: sm_100a ADC path on device %d (fvec_L2sqr lane width %d)\n", device, l2_variant);
    *out = h;
    return 0;
}

int rii_destroy(rii_index_t *h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (DevBuf *b : {&h->skew_lin, &h->skew_lists, &h->skew_off, &h->centers_skew, &h->skew_misc_off, &h->dbg, &h->centers, &h->offsets, &h->ids, &h->loc_len, &h->glob_len, &h->pre_len, &h->T, &h->partial, &h->ranked,
                      &h->merge_cnt, &h->cum, &h->take_last, &h->J, &h->flags, &h->filt, &h->bitmap, &h->q, &h->tids, &h->o_ids, &h->o_dists,
                      &h->o_counts, &h->tmp0, &h->tmp1, &h->tmp2, &h->tmp3, &h->assign, &h->ws_best, &h->ws_arg, &h->d_flag,
                      &h->sub_keys, &h->sub_rows, &h->sub_keys_s, &h->sub_rows_s, &h->sub_bounds, &h->sub_len, &h->sub_off, &h->sub_skew,
                      &h->sub_glob, &h->sub_pre, &h->gen_keys, &h->gen_sorted, &h->gen_seg, &h->gen_cnt, &h->redo_idx, &h->redo_q,
                      &h->redo_ids, &h->redo_d, &h->redo_c})
        b->release();
    if (h->sort_tmp.p) cudaFree(h->sort_tmp.p);
    if (h->pin) cudaFreeHost(h->pin);
    for (cudaEvent_t ev : h->copy_events) cudaEventDestroy(ev);
    if (h->built_ev) cudaEventDestroy(h->built_ev);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->d_R) cudaFree(h->d_R);
    h->qrot.release();
    if (h->d_cw) cudaFree(h->d_cw);
    if (h->d_cw_t) cudaFree(h->d_cw_t);
    if (h->d_Dm) cudaFree(h->d_Dm);
    if (h->d_codes) cudaFree(h->d_codes);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

static int add_codes_impl(rii_index_t *h, const uint8_t *codes, int64_t n, int update_flag, cudaMemcpyKind kind)
{
    if (!h || n < 0 || (n > 0 && !codes)) return fail(RII_ERR_ARG, "bad arguments");
    if (update_flag && h->nlist == 0)  // src/rii.h:166-170
        return fail(RII_ERR_STATE, "reconfigure() must be called before running add(vecs=X, update_posting_lists=True). "
                                   "If this is the first addition, please call add_configure(vecs=X)");
    if (h->N + n >= (1ll << 31)) return fail(RII_ERR_LIMIT, "a shard holds at most 2^31-1 codes (posting lists store int32 ids, src/rii.h:82)");
    CK(cudaSetDevice(h->device));
    if (kind == cudaMemcpyDeviceToDevice) CK(cudaDeviceSynchronize());  // the producer of d_codes may run on any stream of the caller
    const long long N0 = h->N;
    CKR(grow_codes(h, N0 + n));
    if (n) CK(cudaMemcpyAsync(h->d_codes + N0 * h->M, codes, (size_t)n * h->M, kind, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->N = N0 + n;  // (skew_lin_rows keeps the row count of the last build: the next linear query extends the copy)
    if (n && h->N_total >= 0) { h->shard_stale = true; h->has_global = false; }  // lengths / N_total of a shard are stale now
    if (h->verbose) printf("%lld new vectors are added.\nTotal number of codes is %lld\n", (long long)n, h->N);
    if (update_flag) {
        if (h->verbose) printf("Start to update posting lists\n");
        CKR(update_posting_lists(h, N0, n));
    }
    return 0;
}

int rii_add_codes(rii_index_t *h, const uint8_t *codes, int64_t n, int update_flag)
{
    return add_codes_impl(h, codes, n, update_flag, cudaMemcpyHostToDevice);
}

int rii_add_codes_dev(rii_index_t *h, const uint8_t *d_codes, int64_t n, int update_flag)
{
    return add_codes_impl(h, d_codes, n, update_flag, cudaMemcpyDeviceToDevice);
}

int rii_reconfigure(rii_index_t *h, int nlist, int iter)
{
    if (!h) return fail(RII_ERR_ARG, "null index");
    if (!(0 < nlist)) return fail(RII_ERR_ARG, "need 0 < nlist");                  // src/rii.h:110
    if ((long long)nlist > h->N) return fail(RII_ERR_ARG, "need nlist <= N");     // src/rii.h:111
    if (iter < 0) return fail(RII_ERR_ARG, "iter must be >= 0");
    if (h->N_total >= 0 && h->N_total != h->N) return fail(RII_ERR_STATE, "rii_reconfigure on a shard: use rii_fit_coarse + rii_set_coarse_centers");
    CK(cudaSetDevice(h->device));
    // (1) sampling, src/rii.h:115-124 (libstdc++ shuffle defines which codes are sampled)
    const long long N = h->N;
    const long long ns = std::min<long long>(N, (long long)nlist * 100);
    if (h->verbose) printf("The number of vectors used for training of coarse centers: %lld\n", ns);
    std::vector<size_t> pick((size_t)N);
    std::iota(pick.begin(), pick.end(), 0);
    std::shuffle(pick.begin(), pick.end(), std::default_random_engine(123));
    pick.resize((size_t)ns);
    DevBuf d_pick, d_sample;
    int rc = 0;
    std::vector<uint8_t> centers((size_t)nlist * h->M);
    do {
        if ((rc = d_pick.ensure((size_t)ns * 8)) < 0) break;
        if ((rc = d_sample.ensure((size_t)ns * h->M)) < 0) break;
        cudaError_t ce = cudaMemcpyAsync(d_pick.p, pick.data(), (size_t)ns * 8, cudaMemcpyHostToDevice, h->stream);
        long long tot = ns * h->M;
        k_gather_rows<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(h->d_codes, d_pick.as<long long>(), ns, h->M,
                                                                          d_sample.as<uint8_t>());
        LAUNCHED();
        if (ce == cudaSuccess) ce = cudaGetLastError();
        if (ce != cudaSuccess) { rc = fail(RII_ERR_CUDA, std::string("rii_reconfigure: ") + cudaGetErrorString(ce)); break; }
        // (2)+(3) PQk-means, src/rii.h:136-146
        if (h->verbose) printf("Start to run PQk-means\n");
        rc = fit_coarse(h, d_sample.as<uint8_t>(), ns, nlist, iter, centers.data());
    } while (0);
    if (rc < 0) cudaStreamSynchronize(h->stream);
    d_pick.release();
    d_sample.release();
    if (rc < 0) return rc;
    // (4) posting lists, src/rii.h:148-155
    if (h->verbose) printf("Start to update posting lists\n");
    CKR(set_centers(h, centers.data(), nlist));
    return update_posting_lists(h, 0, N);
}

int rii_clear(rii_index_t *h)
{
    if (!h) return fail(RII_ERR_ARG, "null index");
    h->N = 0;
    h->nlist = 0;
    h->skew_lin_rows = -1;
    h->skew_lists_valid = h->centers_skew_valid = false;
    h->h_centers.clear();
    h->h_offsets.assign(1, 0);
    h->N_assigned = 0;
    h->has_global = false;
    h->len_sorted_prefix.clear();
    if (h->N_total >= 0) h->shard_stale = true;
    // the reference's clear() releases its vectors (src/rii.h:328-333): give the code table and every derived layout back
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->d_codes) cudaFree(h->d_codes);
    h->d_codes = nullptr;
    h->cap_rows = 0;
    for (DevBuf *b : {&h->skew_lin, &h->skew_lists, &h->skew_off, &h->centers_skew, &h->ids, &h->assign, &h->ws_best, &h->ws_arg,
                      &h->partial, &h->T, &h->bitmap, &h->tmp0, &h->tmp1, &h->tmp2, &h->tmp3})
        b->release();
    return 0;
}

int64_t rii_query_linear(rii_index_t *h, const float *query, int topk, const int64_t *target_ids, int64_t S, int64_t *out_ids,
                         float *out_dists)
{
    if (!h) return fail(RII_ERR_ARG, "null index");
    int32_t cnt = 0;
    int rc = query_host(h, query, 1, topk, target_ids, S, 0, RII_METHOD_LINEAR, out_ids, out_dists, &cnt);
    return rc < 0 ? rc : cnt;
}

int64_t rii_query_ivf(rii_index_t *h, const float *query, int topk, const int64_t *target_ids, int64_t S, int64_t L,
                      int64_t *out_ids, float *out_dists)
{
    if (!h) return fail(RII_ERR_ARG, "null index");
    int32_t cnt = 0;
    int rc = query_host(h, query, 1, topk, target_ids, S, L, RII_METHOD_IVF, out_ids, out_dists, &cnt);
    return rc < 0 ? rc : cnt;
}

int rii_query_batch(rii_index_t *h, const float *queries, int B, int topk, const int64_t *target_ids, int64_t S, int64_t L,
                    int method, int64_t *out_ids, float *out_dists, int32_t *out_counts)
{
    if (!h) return fail(RII_ERR_ARG, "null index");
    return query_host(h, queries, B, topk, target_ids, S, L, method, out_ids, out_dists, out_counts);
}

int rii_query_batch_dev(rii_index_t *h, const float *d_queries, int B, int topk, const int64_t *d_target_ids, int64_t S,
                        int64_t L, int method, int64_t *d_out_ids, float *d_out_dists, int32_t *d_out_counts, void *stream)
{
    if (!h) return fail(RII_ERR_ARG, "null index");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;  // NULL == the CUDA (legacy) default stream, as everywhere in CUDA
    return query_dev(h, d_queries, B, topk, (const long long *)d_target_ids, S, L, method, (long long *)d_out_ids, d_out_dists,
                     d_out_counts, st);
}

int rii_set_option(rii_index_t *h, const char *name, int64_t value)
{
    if (!h || !name) return fail(RII_ERR_ARG, "bad arguments");
    if (!strcmp(name, "stream_ctas")) {
        if (value < 0 || value > 2) return fail(RII_ERR_ARG, "stream_ctas must be 0 (auto), 1 or 2");
        h->opt_stream_ctas = (int)value;
        return 0;
    }
    if (!strcmp(name, "scan_kernel")) {
        if (value != 0 && value != 1 && value != 4)
            return fail(RII_ERR_ARG, "scan_kernel must be 0 (auto), 1 (natural-layout kernels) or 4 (skew64 streaming engine)");
        h->opt_scan_kernel = (int)value;
        return 0;
    }
    if (!strcmp(name, "zero_copy")) {
        h->opt_zero_copy = value != 0;
        return 0;
    }
    if (!strcmp(name, "persist")) {
        if (value < 0 || value > 2) return fail(RII_ERR_ARG, "persist must be 0 (off), 1 (auto) or 2 (whenever the shape fits)");
        h->opt_persist = (int)value;
        return 0;
    }
    if (!strcmp(name, "assign_kernel")) {
        if (value != 0 && value != 1 && value != 3) return fail(RII_ERR_ARG, "assign_kernel must be 0 (auto), 1 (natural layout) or 3 (streaming, one CTA per SM)");
        h->opt_assign_kernel = (int)value;
        return 0;
    }
    if (!strcmp(name, "fuse_coarse")) {
        h->opt_fuse_coarse = value != 0;
        return 0;
    }
    if (!strcmp(name, "debug_clocks")) {
        h->opt_debug_clocks = value != 0;
        return 0;
    }
    return fail(RII_ERR_ARG, std::string("unknown option: ") + name);
}

int rii_debug_clocks(rii_index_t *h, int64_t n_ctas, int64_t *out)
{
    // out: (n_ctas, 8) clock64() values recorded by the last v2 scan launch (option "debug_clocks" = 1)
    if (!h || !out || n_ctas <= 0) return fail(RII_ERR_ARG, "bad arguments");
    if ((size_t)n_ctas * 64 > h->dbg.cap) return fail(RII_ERR_ARG, "no debug clocks recorded for that many CTAs");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, h->dbg.p, (size_t)n_ctas * 64, cudaMemcpyDeviceToHost));
    return 0;
}

int rii_profile_enable(rii_index_t *h, int on)
{
    if (!h) return fail(RII_ERR_ARG, "null index");
    h->prof = on != 0;
    return 0;
}
int rii_profile_reset(rii_index_t *h)
{
    if (!h) return fail(RII_ERR_ARG, "null index");
    prof_collect(h);
    for (int i = 0; i < PK_N; ++i) { h->prof_ms[i] = 0; h->prof_n[i] = 0; }
    return 0;
}
int rii_profile_get(rii_index_t *h, const char *kernel, double *ms_total, int64_t *launches)
{
    if (!h || !kernel) return fail(RII_ERR_ARG, "bad arguments");
    prof_collect(h);
    for (int i = 0; i < PK_N; ++i)
        if (!strcmp(kernel, PK_NAMES[i])) {
            if (ms_total) *ms_total = h->prof_ms[i];
            if (launches) *launches = h->prof_n[i];
            return 0;
        }
    return fail(RII_ERR_ARG, std::string("unknown kernel name: ") + kernel);
}

int rii_sample_ids(int64_t N_total, int nlist, int64_t *out_ids, int64_t *out_n)
{
    // src/rii.h:115-124: first min(N, 100*nlist) entries of std::shuffle(iota(N), default_random_engine(123))
    if (N_total <= 0 || nlist <= 0 || !out_n) return fail(RII_ERR_ARG, "bad arguments");
    const long long ns = std::min<long long>(N_total, (long long)nlist * 100);
    *out_n = ns;
    if (!out_ids) return 0;
    std::vector<size_t> pick((size_t)N_total);
    std::iota(pick.begin(), pick.end(), 0);
    std::shuffle(pick.begin(), pick.end(), std::default_random_engine(123));
    for (long long i = 0; i < ns; ++i) out_ids[i] = (int64_t)pick[i];
    return 0;
}

// ---- IVF + target_ids on an id-range shard (the one exchange step) -------------------------------
// The cut after L *member* candidates and the topk test at the w-th list are global, so every shard needs the member
// counts of the others -- per LIST, not per query: each shard builds the sub-index of its own members once per target
// set, the per-list member counts are all-gathered (nlist ints per shard), and the queries then run as ordinary sharded
// IVF searches over the sub-indexes.
int rii_subset_begin_dev(rii_index_t *h, const int64_t *d_target_ids, int64_t S, int32_t *d_list_counts, void *stream)
{
    if (!h || !d_target_ids || S <= 0) return fail(RII_ERR_ARG, "bad arguments");
    if (h->nlist <= 0) return fail(RII_ERR_STATE, "query_ivf before reconfigure(): no posting lists");
    if (S > h->n_total()) return fail(RII_ERR_ARG, "need len(target_ids) <= N");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    ListsView v;
    h->sub_S = -1;
    CKR(build_subview(h, (const long long *)d_target_ids, S, st, &v));
    if (d_list_counts) CK(cudaMemcpyAsync(d_list_counts, h->sub_len.p, (size_t)h->nlist * 4, cudaMemcpyDeviceToDevice, st));
    CKR(h->sub_glob.ensure((size_t)h->nlist * 4));
    CKR(h->sub_pre.ensure((size_t)h->nlist * 4));
    CK(cudaMemcpyAsync(h->sub_glob.p, h->sub_len.p, (size_t)h->nlist * 4, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemsetAsync(h->sub_pre.p, 0, (size_t)h->nlist * 4, st));
    h->sub_S = S;
    return 0;
}

int rii_subset_set_global_dev(rii_index_t *h, const int32_t *d_counts_all, const int32_t *d_counts_lower, void *stream)
{
    if (!h || !d_counts_all || !d_counts_lower) return fail(RII_ERR_ARG, "bad arguments");
    if (h->sub_S < 0) return fail(RII_ERR_STATE, "rii_subset_begin_dev first");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(h->sub_glob.p, d_counts_all, (size_t)h->nlist * 4, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(h->sub_pre.p, d_counts_lower, (size_t)h->nlist * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int rii_subset_query_dev(rii_index_t *h, const float *d_queries, int B, int topk, int64_t L, int64_t *d_out_ids, float *d_out_dists,
                         int32_t *d_out_counts, void *stream)
{
    if (!h || !d_queries || !d_out_ids || !d_out_dists || !d_out_counts || B < 1) return fail(RII_ERR_ARG, "bad arguments");
    if (h->sub_S < 0) return fail(RII_ERR_STATE, "rii_subset_begin_dev first");
    if (h->shard_stale) return fail(RII_ERR_STATE, "codes were added to / cleared from this shard: call rii_set_shard again");
    const long long S = h->sub_S, Ntot = h->n_total();
    if (topk < 1 || topk > S) return fail(RII_ERR_ARG, "need 1 <= topk <= len(target_ids)");
    if (!(topk <= L && L <= Ntot)) return fail(RII_ERR_ARG, "need topk <= L <= N");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    ListsView v;
    v.offsets = h->sub_bounds.as<long long>();
    v.ids = h->sub_rows_s.as<int>();
    v.loc_len = h->sub_len.as<int>();
    v.glob_len = h->sub_glob.as<int>();
    v.pre_len = h->sub_pre.as<int>();
    v.cap_local = std::min<long long>(S, h->N);
    if (h->rb && h->opt_scan_kernel != 1) { v.skew = h->sub_skew.as<uint8_t>(); v.skew_off = h->sub_off.as<long long>(); }
    return ivf_batches(h, d_queries, B, topk, L, v, ivf_width(h, S, L), true, (long long *)d_out_ids, d_out_dists, d_out_counts, st);
}

// ---- coarse phase split from the scan (sharded batches: every rank ranks the lists for its share of the queries, the
// rankings are all-gathered, every rank scans its shard for all queries) ----------------------------------------
int rii_coarse_width(rii_index_t *h, int64_t L)
{
    if (!h) return fail(RII_ERR_ARG, "null index");
    if (h->nlist <= 0) return fail(RII_ERR_STATE, "no posting lists");
    return ivf_width(h, 0, L);
}

int rii_coarse_rank_dev(rii_index_t *h, const float *d_queries, int B, int topk, int64_t L, int32_t *d_ranked, void *stream)
{
    if (!h || !d_queries || !d_ranked || B < 1) return fail(RII_ERR_ARG, "bad arguments");
    if (h->nlist <= 0) return fail(RII_ERR_STATE, "no posting lists");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    ListsView v;
    CKR(main_view(h, st, &v));
    const int w = ivf_width(h, 0, L);
    const int D = h->M * h->Ds, CH = 32768;
    for (int b0 = 0; b0 < B; b0 += CH)
        CKR(run_ivf(h, d_queries + (size_t)b0 * D, std::min(CH, B - b0), topk, L, v, w, w, nullptr, nullptr, nullptr, st, 1, d_ranked + (size_t)b0 * w));
    return 0;
}

int rii_query_ranked_dev(rii_index_t *h, const float *d_queries, int B, int topk, int64_t L, const int32_t *d_ranked, int64_t *d_out_ids,
                         float *d_out_dists, int32_t *d_out_counts, int32_t *d_flags, void *stream)
{
    if (!h || !d_queries || !d_ranked || !d_out_ids || !d_out_dists || !d_out_counts || B < 1) return fail(RII_ERR_ARG, "bad arguments");
    if (h->nlist <= 0) return fail(RII_ERR_STATE, "no posting lists");
    if (h->shard_stale) return fail(RII_ERR_STATE, "codes were added to / cleared from this shard: call rii_set_shard again");
    if (h->N_total >= 0 && h->N_total != h->N && !h->has_global)
        return fail(RII_ERR_STATE, "IVF query on a shard whose list lengths were not exchanged: call rii_set_global_lengths");
    const long long Ntot = h->n_total();
    if (topk < 1 || !(topk <= L && L <= Ntot)) return fail(RII_ERR_ARG, "need 1 <= topk <= L <= N");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    ListsView v;
    CKR(main_view(h, st, &v));
    const int w = ivf_width(h, 0, L);
    const int D = h->M * h->Ds, CH = 32768;
    for (int b0 = 0; b0 < B; b0 += CH) {
        const int bc = std::min(CH, B - b0);
        CKR(run_ivf(h, d_queries + (size_t)b0 * D, bc, topk, L, v, w, w, (long long *)d_out_ids + (size_t)b0 * topk, d_out_dists + (size_t)b0 * topk,
                    d_out_counts + b0, st, 2, const_cast<int *>(d_ranked) + (size_t)b0 * w));
        if (d_flags) CK(cudaMemcpyAsync(d_flags + b0, h->flags.p, (size_t)bc * 4, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

int rii_merge_shards_dev(rii_index_t *h, const int64_t *d_ids, const float *d_dists, const int32_t *d_counts, int G, int B,
                         int k, int64_t *d_out_ids, float *d_out_dists, int32_t *d_out_counts, void *stream)
{
    if (!h || !d_ids || !d_dists || !d_counts || G <= 0 || B <= 0 || k <= 0) return fail(RII_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int P = next_pow2(G * k < 2 ? 2 : G * k);
    const size_t smem = (size_t)P * 12;
    CKR(set_smem(k_merge_shards, smem));
    Prof pr(h, st, PK_MERGE);
    k_merge_shards<<<B, RII_THREADS, smem, st>>>((const long long *)d_ids, d_dists, d_counts, 0, G, B, k, P, (long long *)d_out_ids,
                                                 d_out_dists, d_out_counts);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

int64_t rii_get_N(const rii_index_t *h) { return h ? h->N : 0; }
int rii_get_nlist(const rii_index_t *h) { return h ? h->nlist : 0; }
int rii_get_verbose(const rii_index_t *h) { return h ? h->verbose : 0; }
int rii_set_verbose(rii_index_t *h, int verbose)
{
    if (!h) return fail(RII_ERR_ARG, "null index");
    h->verbose = verbose;
    return 0;
}
int rii_get_dims(const rii_index_t *h, int *M, int *Ks, int *Ds)
{
    if (!h) return fail(RII_ERR_ARG, "null index");
    if (M) *M = h->M;
    if (Ks) *Ks = h->Ks;
    if (Ds) *Ds = h->Ds;
    return 0;
}

int rii_copy_codes(const rii_index_t *h, uint8_t *out)
{
    if (!h || !out) return fail(RII_ERR_ARG, "null argument");
    CK(cudaSetDevice(h->device));
    if (h->N) CK(cudaMemcpy(out, h->d_codes, (size_t)h->N * h->M, cudaMemcpyDeviceToHost));
    return 0;
}
int rii_copy_coarse_centers(const rii_index_t *h, uint8_t *out)
{
    if (!h || !out) return fail(RII_ERR_ARG, "null argument");
    if (!h->h_centers.empty()) std::memcpy(out, h->h_centers.data(), h->h_centers.size());
    return 0;
}
int rii_copy_posting_lists(const rii_index_t *h, int64_t *offsets, int32_t *ids)
{
    if (!h || !offsets) return fail(RII_ERR_ARG, "null argument");
    for (int i = 0; i <= h->nlist; ++i) offsets[i] = h->h_offsets[i];
    const long long tot = h->h_offsets[h->nlist];
    if (ids && tot > 0) {
        CK(cudaSetDevice(h->device));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpy(ids, h->ids.p, (size_t)tot * 4, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int rii_set_state(rii_index_t *h, const uint8_t *coarse_centers, int nlist, const uint8_t *codes, int64_t N,
                  const int64_t *offsets, const int32_t *ids)
{
    if (!h || nlist < 0 || N < 0) return fail(RII_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(h->device));
    CKR(rii_clear(h));
    if (N) {
        if (!codes) return fail(RII_ERR_ARG, "codes is null");
        CKR(rii_add_codes(h, codes, N, 0));
    }
    if (nlist) {
        if (!coarse_centers || !offsets) return fail(RII_ERR_ARG, "coarse centers / offsets are null");
        CKR(set_centers(h, coarse_centers, nlist));
        const long long tot = offsets[nlist];
        if (tot > 0 && !ids) return fail(RII_ERR_ARG, "ids is null");
        for (int i = 0; i < nlist; ++i)
            if (offsets[i + 1] < offsets[i]) return fail(RII_ERR_ARG, "posting list offsets must be non-decreasing");
        for (long long i = 0; i < tot; ++i)
            if (ids[i] < 0 || ids[i] >= N) return fail(RII_ERR_ARG, "posting list id outside [0, N)");
        if (offsets[0] != 0) return fail(RII_ERR_ARG, "posting list offsets must start at 0");
        h->h_offsets.assign(offsets, offsets + nlist + 1);
        CKR(h->ids.ensure(std::max<size_t>(4, (size_t)tot * 4)));
        if (tot) CK(cudaMemcpyAsync(h->ids.p, ids, (size_t)tot * 4, cudaMemcpyHostToDevice, h->stream));
        CKR(finish_lists(h));
        if (tot) {  // row -> list map (rows that are in no list keep -1)
            CKR(grow_assign(h, N));
            CK(cudaMemsetAsync(h->assign.p, 0xff, (size_t)N * 4, h->stream));
            k_assign_from_csr<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(h->offsets.as<long long>(), h->ids.as<int>(), nlist, tot,
                                                                                    h->assign.as<int>());
            LAUNCHED();
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(h->stream));
            h->N_assigned = N;
        }
    }
    return 0;
}

int rii_dtable(rii_index_t *h, const float *queries, int B, float *out)
{
    if (!h || !queries || !out || B <= 0) return fail(RII_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(h->device));
    const int D = h->M * h->Ds, lutf = h->M * h->Ks;
    CKR(h->q.ensure((size_t)B * D * 4));
    CKR(h->T.ensure((size_t)B * lutf * 4));
    CK(cudaMemcpyAsync(h->q.p, queries, (size_t)B * D * 4, cudaMemcpyHostToDevice, h->stream));
    dim3 grid((lutf + RII_THREADS - 1) / RII_THREADS, B);
    k_dtable<<<grid, RII_THREADS, 0, h->stream>>>(h->q.as<float>(), h->d_cw, h->T.as<float>(), h->M, h->Ks, h->Ds, h->variant);
    LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, h->T.p, (size_t)B * lutf * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int rii_adist_all(rii_index_t *h, const float *query, float *out)
{
    if (!h || !query || !out) return fail(RII_ERR_ARG, "bad arguments");
    if (h->N == 0) return 0;
    CK(cudaSetDevice(h->device));
    const int D = h->M * h->Ds, lutf = h->M * h->Ks;
    CKR(h->q.ensure((size_t)D * 4));
    CKR(h->T.ensure((size_t)lutf * 4));
    CKR(h->tmp0.ensure((size_t)h->N * 4));
    CK(cudaMemcpyAsync(h->q.p, query, (size_t)D * 4, cudaMemcpyHostToDevice, h->stream));
    k_dtable<<<dim3((lutf + RII_THREADS - 1) / RII_THREADS, 1), RII_THREADS, 0, h->stream>>>(h->q.as<float>(), h->d_cw, h->T.as<float>(),
                                                                                           h->M, h->Ks, h->Ds, h->variant);
    LAUNCHED();
    const size_t smem = (size_t)lutf * 4;
    const unsigned grid = (unsigned)std::min<long long>(1184, (h->N + RII_THREADS - 1) / RII_THREADS);
    DISPATCH_M(h->M, {
        CKR(set_smem(k_adc_all<MT>, smem));
        k_adc_all<MT><<<dim3(grid, 1), RII_THREADS, smem, h->stream>>>(h->T.as<float>(), h->d_codes, h->N, h->M, h->Ks, h->tmp0.as<float>());
    });
    LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, h->tmp0.p, (size_t)h->N * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int rii_assign(rii_index_t *h, const uint8_t *codes, int64_t n, const uint8_t *centers, int K, int32_t *out_assign, float *out_dist)
{
    if (!h || !codes || !centers || !out_assign || n < 0 || K <= 0) return fail(RII_ERR_ARG, "bad arguments");
    if (n == 0) return 0;
    CK(cudaSetDevice(h->device));
    DevBuf dc, dk, da, dd, dsk;
    int rc = 0;
    do {
        if ((rc = dc.ensure((size_t)n * h->M)) < 0) break;
        if ((rc = dk.ensure((size_t)K * h->M)) < 0) break;
        if ((rc = da.ensure((size_t)n * 4)) < 0) break;
        if (out_dist && (rc = dd.ensure((size_t)n * 4)) < 0) break;
        cudaMemcpyAsync(dc.p, codes, (size_t)n * h->M, cudaMemcpyHostToDevice, h->stream);
        cudaMemcpyAsync(dk.p, centers, (size_t)K * h->M, cudaMemcpyHostToDevice, h->stream);
        AssignSrc src;
        if ((rc = make_assign_src(h, dc.as<uint8_t>(), n, &dsk, &src)) < 0) break;
        if ((rc = launch_assign(h, src, dk.as<uint8_t>(), K, da.as<int>(), out_dist ? dd.as<float>() : nullptr)) < 0) break;
        cudaMemcpyAsync(out_assign, da.p, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream);
        if (out_dist) cudaMemcpyAsync(out_dist, dd.p, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream);
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(RII_ERR_CUDA, std::string("rii_assign: ") + cudaGetErrorString(e));
    } while (0);
    dc.release(); dk.release(); da.release(); dd.release(); dsk.release();
    return rc;
}

int rii_sym_matrices(rii_index_t *h, float *out)
{
    if (!h || !out) return fail(RII_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(h->device));
    CKR(ensure_Dm(h));
    CK(cudaMemcpyAsync(out, h->d_Dm, (size_t)h->M * h->Ks * h->Ks * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int rii_encode(rii_index_t *h, const float *vecs, int64_t n, uint8_t *out_codes)
{
    if (!h || !vecs || !out_codes || n < 0) return fail(RII_ERR_ARG, "bad arguments");
    if (n == 0) return 0;
    CK(cudaSetDevice(h->device));
    const int D = h->M * h->Ds;
    const size_t smem = (size_t)h->Ks * h->Ds * 4;
    CKR(set_smem(k_pq_encode, smem));
    DevBuf dx, dc;
    int rc = 0;
    const long long CH = 4ll << 20;  // rows per pass
    do {
        if ((rc = dx.ensure((size_t)std::min<long long>(n, CH) * D * 4)) < 0) break;
        if ((rc = dc.ensure((size_t)std::min<long long>(n, CH) * h->M)) < 0) break;
        for (long long s0 = 0; s0 < n && rc == 0; s0 += CH) {
            const long long c = std::min<long long>(CH, n - s0);
            cudaMemcpyAsync(dx.p, vecs + (size_t)s0 * D, (size_t)c * D * 4, cudaMemcpyHostToDevice, h->stream);
            k_pq_encode<<<dim3((unsigned)((c + RII_THREADS - 1) / RII_THREADS), h->M), RII_THREADS, smem, h->stream>>>(
                dx.as<float>(), c, h->d_cw, h->M, h->Ks, h->Ds, dc.as<uint8_t>());
            LAUNCHED();
            cudaMemcpyAsync(out_codes + (size_t)s0 * h->M, dc.p, (size_t)c * h->M, cudaMemcpyDeviceToHost, h->stream);
            cudaError_t e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) rc = fail(RII_ERR_CUDA, std::string("rii_encode: ") + cudaGetErrorString(e));
        }
    } while (0);
    dx.release();
    dc.release();
    return rc;
}

int rii_set_rotation(rii_index_t *h, const float *R)
{
    if (!h) return fail(RII_ERR_ARG, "null index");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    if (h->d_R) cudaFree(h->d_R);
    h->d_R = nullptr;
    if (!R) return 0;
    const size_t D = (size_t)h->M * h->Ds;
    CK(cudaMalloc(&h->d_R, D * D * 4));
    CK(cudaMemcpy(h->d_R, R, D * D * 4, cudaMemcpyHostToDevice));
    return 0;
}

int rii_set_shard(rii_index_t *h, int64_t id_base, int64_t N_total)
{
    if (!h || id_base < 0 || N_total < 0) return fail(RII_ERR_ARG, "bad arguments");
    if (N_total < h->N) return fail(RII_ERR_ARG, "N_total is smaller than the number of local codes");
    h->id_base = id_base;
    h->N_total = N_total;
    h->shard_stale = false;
    return 0;
}

int rii_set_lists_dev(rii_index_t *h, const uint8_t *centers, int nlist, const int32_t *d_assign)
{
    if (!h || !centers || nlist <= 0 || !d_assign) return fail(RII_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());  // the producer of d_assign may run on any stream of the caller
    CKR(set_centers(h, centers, nlist));
    return update_posting_lists(h, 0, h->N, d_assign);
}

int rii_reserve(rii_index_t *h, int64_t rows)
{
    if (!h || rows < 0) return fail(RII_ERR_ARG, "bad arguments");
    if (rows >= (1ll << 31)) return fail(RII_ERR_LIMIT, "a shard holds at most 2^31-1 codes");
    CK(cudaSetDevice(h->device));
    return grow_codes(h, rows);
}

int rii_merge_shards_packed_dev(rii_index_t *h, const void *d_packed, int64_t stride_bytes, int G, int B, int k, int64_t *d_out_ids,
                                float *d_out_dists, int32_t *d_out_counts, void *stream)
{
    if (!h || !d_packed || G <= 0 || B <= 0 || k <= 0 || stride_bytes < (int64_t)B * k * 12 + (int64_t)B * 4)
        return fail(RII_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int P = next_pow2(G * k < 2 ? 2 : G * k);
    const size_t smem = (size_t)P * 12;
    CKR(set_smem(k_merge_shards, smem));
    Prof pr(h, st, PK_MERGE);
    const char *base = (const char *)d_packed;
    k_merge_shards<<<B, RII_THREADS, smem, st>>>((const long long *)base, (const float *)(base + (size_t)B * k * 8),
                                                 (const int *)(base + (size_t)B * k * 12), stride_bytes, G, B, k, P, (long long *)d_out_ids,
                                                 d_out_dists, d_out_counts);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

int rii_set_coarse_centers(rii_index_t *h, const uint8_t *centers, int nlist)
{
    if (!h || !centers || nlist <= 0) return fail(RII_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(h->device));
    CKR(set_centers(h, centers, nlist));
    return update_posting_lists(h, 0, h->N);
}

int rii_fit_coarse(rii_index_t *h, const uint8_t *sample, int64_t ns, int nlist, int iter, uint8_t *centers_out)
{
    if (!h || !sample || !centers_out || ns <= 0) return fail(RII_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(h->device));
    DevBuf ds;
    int rc = ds.ensure((size_t)ns * h->M);
    if (rc == 0) {
        cudaMemcpyAsync(ds.p, sample, (size_t)ns * h->M, cudaMemcpyHostToDevice, h->stream);
        rc = fit_coarse(h, ds.as<uint8_t>(), ns, nlist, iter, centers_out);
    }
    ds.release();
    return rc;
}

int rii_copy_list_lengths(const rii_index_t *h, int32_t *out)
{
    if (!h || !out) return fail(RII_ERR_ARG, "bad arguments");
    for (int i = 0; i < h->nlist; ++i) out[i] = (int32_t)(h->h_offsets[i + 1] - h->h_offsets[i]);
    return 0;
}

int rii_set_global_lengths(rii_index_t *h, const int32_t *glob_len, const int32_t *pre_len)
{
    if (!h || !glob_len || !pre_len || h->nlist <= 0) return fail(RII_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(h->device));
    CKR(h->glob_len.ensure((size_t)h->nlist * 4));
    CKR(h->pre_len.ensure((size_t)h->nlist * 4));
    CK(cudaMemcpy(h->glob_len.p, glob_len, (size_t)h->nlist * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->pre_len.p, pre_len, (size_t)h->nlist * 4, cudaMemcpyHostToDevice));
    std::vector<int> len(glob_len, glob_len + h->nlist);
    std::sort(len.begin(), len.end());
    h->len_sorted_prefix.assign(h->nlist + 1, 0);
    for (int i = 0; i < h->nlist; ++i) h->len_sorted_prefix[i + 1] = h->len_sorted_prefix[i] + len[i];
    h->has_global = true;
    return 0;
}

}  // extern "C"
