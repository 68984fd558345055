// oracle/rii_oracle.cpp -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference's ADC hot path).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library, and only as the *checker* (or the reported CPU baseline) -- never as the product path.
// The product (rii_b200/csrc) does not link, import or call anything in oracle/.
//
// Parity status: PINNED.  Every function below is checked bit-for-bit against the UNMODIFIED reference
// compiled from /root/reference/src with "-O2 -ffp-contract=off" (oracle/_ref/strict_*; see
// oracle/Makefile; tests/golden/make_golden.py drives it) and against golden vectors generated from that build
// (tests/golden/*.npz, checked by tests/test_oracle_golden.py).
//
// All file:line citations are into /root/reference/ (matsui528/rii v0.2.12).  The arithmetic is the
// reference's arithmetic *as written*: fp32, separate sub/mul/add (no FMA contraction, no
// reassociation) -- compile this file with -ffp-contract=off and without -ffast-math.
//
// Tie-breaking: the reference's std::partial_sort compares distances only (src/rii.h:234-235,279-280,
// 312-313), which leaves the order of exact ties to libstdc++'s heap internals.  The oracle (and the
// CUDA path) use the total order (distance ascending, id ascending) everywhere.

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <random>
#include <utility>
#include <vector>

namespace {

// ---- src/distance.h:117-170 (AVX-512), :177-217 (AVX), :225-252 (SSE), masked_read :44-65 ----------
// Lane-structured emulation of fvec_L2sqr.  `variant` is the widest accumulator the reference build
// uses: 16 (__AVX512F__), 8 (__AVX__) or 4 (SSE).  For d < 8 the three variants are identical.
float l2sqr_lanes(const float *x, const float *y, int d, int variant)
{
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    if (variant == 16) {
        float a16[16];
        for (int i = 0; i < 16; ++i) a16[i] = 0.f;
        while (d >= 16) {                                        // :120-127
            for (int i = 0; i < 16; ++i) { float t = x[i] - y[i]; a16[i] = a16[i] + t * t; }
            x += 16; y += 16; d -= 16;
        }
        float a8[8];
        for (int i = 0; i < 8; ++i) a8[i] = a16[8 + i] + a16[i];  // :129-131  hi + lo
        while (d >= 8) {                                         // :133-141
            for (int i = 0; i < 8; ++i) { float t = x[i] - y[i]; a8[i] = a8[i] + t * t; }
            x += 8; y += 8; d -= 8;
        }
        for (int i = 0; i < 4; ++i) a4[i] = a8[4 + i] + a8[i];    // :143-145
    } else if (variant == 8) {
        float a8[8];
        for (int i = 0; i < 8; ++i) a8[i] = 0.f;
        while (d >= 8) {                                         // :181-189
            for (int i = 0; i < 8; ++i) { float t = x[i] - y[i]; a8[i] = a8[i] + t * t; }
            x += 8; y += 8; d -= 8;
        }
        for (int i = 0; i < 4; ++i) a4[i] = a8[4 + i] + a8[i];    // :191-193
    } else {                                                     // SSE :229-237 (loops over 4-blocks)
        while (d >= 8) {
            for (int i = 0; i < 4; ++i) { float t = x[i] - y[i]; a4[i] = a4[i] + t * t; }
            x += 4; y += 4; d -= 4;
        }
    }
    if (d >= 4) {                                                // :147-155 / :195-203 / :229-237
        for (int i = 0; i < 4; ++i) { float t = x[i] - y[i]; a4[i] = a4[i] + t * t; }
        x += 4; y += 4; d -= 4;
    }
    if (d > 0) {                                                 // masked tail :157-164 (zero padded lanes)
        for (int i = 0; i < 4; ++i) {
            float xi = i < d ? x[i] : 0.f, yi = i < d ? y[i] : 0.f;
            float t = xi - yi;
            a4[i] = a4[i] + t * t;
        }
    }
    float h0 = a4[0] + a4[1];                                    // two _mm_hadd_ps :166-168
    float h1 = a4[2] + a4[3];
    return h0 + h1;
}

// ---- src/rii.h:375-394  ADist: dist = 0; for m: dist += T[m][code[m]]  (sequential fp32) -----------
inline float adist(const float *T, const uint8_t *code, int M, int Ks)
{
    float dist = 0.f;
    for (int m = 0; m < M; ++m) dist += T[(size_t)m * Ks + code[m]];
    return dist;
}

struct Cand { float dist; int64_t id; };
inline bool cand_less(const Cand &a, const Cand &b)
{
    return a.dist < b.dist || (a.dist == b.dist && a.id < b.id);
}

// ---- src/pqkmeans.cpp:152-162 SymmetricDistance ----------------------------------------------------
inline float symdist(const float *Dm, const uint8_t *a, const uint8_t *b, int M, int Ks)
{
    float dist = 0.f;
    for (int m = 0; m < M; ++m) dist += Dm[((size_t)m * Ks + a[m]) * Ks + b[m]];
    return dist;
}

// ---- src/pqkmeans.cpp:193-218 FindNearetCenterLinear: first minimum wins (strict <) ---------------
inline std::pair<int, float> nearest_center(const float *Dm, const uint8_t *code, const uint8_t *centers,
                                            int K, int M, int Ks)
{
    float min_dist = FLT_MAX;
    int min_i = -1;
    for (int i = 0; i < K; ++i) {
        float d = symdist(Dm, code, centers + (size_t)i * M, M, Ks);
        if (d < min_dist) { min_i = i; min_dist = d; }
    }
    return {min_i, min_dist};
}

}  // namespace

extern "C" {

float orc_l2sqr(const float *x, const float *y, int d, int variant) { return l2sqr_lanes(x, y, d, variant); }

// src/rii.h:361-373 DTable: T[m][ks] = fvec_L2sqr(q + m*Ds, codewords[m][ks], Ds)
void orc_dtable(const float *q, const float *cw, int M, int Ks, int Ds, int variant, float *T)
{
    for (int m = 0; m < M; ++m)
        for (int ks = 0; ks < Ks; ++ks)
            T[(size_t)m * Ks + ks] = l2sqr_lanes(q + (size_t)m * Ds, cw + ((size_t)m * Ks + ks) * Ds, Ds, variant);
}

// src/rii.h:386-394 for every row of `codes`
void orc_adist_all(const float *T, const uint8_t *codes, int64_t N, int M, int Ks, float *out)
{
#pragma omp parallel for
    for (int64_t n = 0; n < N; ++n) out[n] = adist(T, codes + n * M, M, Ks);
}

// src/rii.h:195-242 QueryLinear.  tids: candidates in the given order (S==0 -> all ids).
// Returns the number of results written (== topk).
int64_t orc_query_linear(const float *T, const uint8_t *codes, int64_t N, int M, int Ks, int topk,
                         const int64_t *tids, int64_t S, int64_t *out_ids, float *out_dists)
{
    std::vector<Cand> sc;
    if (S == 0) {
        sc.resize(N);
#pragma omp parallel for
        for (int64_t n = 0; n < N; ++n) sc[n] = {adist(T, codes + n * M, M, Ks), n};
    } else {
        sc.resize(S);
#pragma omp parallel for
        for (int64_t s = 0; s < S; ++s) sc[s] = {adist(T, codes + tids[s] * M, M, Ks), tids[s]};
    }
    if ((size_t)topk > sc.size()) topk = (int)sc.size();
    std::partial_sort(sc.begin(), sc.begin() + topk, sc.end(), cand_less);
    for (int i = 0; i < topk; ++i) { out_ids[i] = sc[i].id; out_dists[i] = sc[i].dist; }
    return topk;
}

// src/rii.h:244-326 QueryIvf, sequential form.  CSR posting lists
// (offsets[nlist+1], ids int32 ascending per list).  tids must be sorted when S != 0 (:294).
// Lists are ranked by (coarse dist, list id) -- the reference ranks only the first w and leaves the
// rest in heap-remnant order (:279-280).  Returns the number of results (0 == the empty case :325).
// If `out_cand` != null it receives the number of candidates collected (for tests).
int64_t orc_query_ivf(const float *T, const uint8_t *codes, int64_t N, int M, int Ks,
                      const uint8_t *centers, int nlist, const int64_t *offsets, const int32_t *ids,
                      int topk, const int64_t *tids, int64_t S, int64_t L,
                      int64_t *out_ids, float *out_dists, int64_t *out_cand)
{
    std::vector<Cand> coarse(nlist);
    for (int no = 0; no < nlist; ++no) coarse[no] = {adist(T, centers + (size_t)no * M, M, Ks), no};   // :262-264
    size_t w = (size_t)std::round((double)L * nlist / (S == 0 ? N : S));                                   // :267-272
    w += 3;                                                                                                 // :273
    if ((size_t)nlist < w) w = nlist;                                                                       // :274-276
    std::sort(coarse.begin(), coarse.end(), cand_less);                                                     // :279-280 (total order)

    std::vector<Cand> sc;
    sc.reserve(L);
    size_t coarse_cnt = 0;
    bool finished = false;
    for (const auto &c : coarse) {                                                                          // :286
        int no = (int)c.id;
        coarse_cnt++;
        for (int64_t p = offsets[no]; p < offsets[no + 1]; ++p) {                                           // :291
            int64_t n = ids[p];
            if (S != 0 && !std::binary_search(tids, tids + S, n)) continue;                                 // :294
            sc.push_back({adist(T, codes + n * M, M, Ks), n});                                              // :299
            if ((int64_t)sc.size() == L) { finished = true; break; }                                        // :302-304
        }
        if (finished) break;
        if (coarse_cnt == w && sc.size() >= (size_t)topk) { finished = true; break; }                       // :309
    }
    if (out_cand) *out_cand = (int64_t)sc.size();
    if (!finished) return 0;                                                                                // :325
    std::partial_sort(sc.begin(), sc.begin() + topk, sc.end(), cand_less);                                  // :312-313
    for (int i = 0; i < topk; ++i) { out_ids[i] = sc[i].id; out_dists[i] = sc[i].dist; }
    return topk;
}

// src/pqkmeans.cpp:23-34 + L2SquaredDistance :164-173 (scalar, sequential in i)
void orc_sym_matrices(const float *cw, int M, int Ks, int Ds, float *Dm)
{
    for (int m = 0; m < M; ++m)
        for (int k1 = 0; k1 < Ks; ++k1)
            for (int k2 = 0; k2 < Ks; ++k2) {
                const float *a = cw + ((size_t)m * Ks + k1) * Ds, *b = cw + ((size_t)m * Ks + k2) * Ds;
                float dist = 0.f;
                for (int i = 0; i < Ds; ++i) dist += (a[i] - b[i]) * (a[i] - b[i]);
                Dm[((size_t)m * Ks + k1) * Ks + k2] = dist;
            }
}

// src/rii.h:350-354 / src/pqkmeans.cpp:88-94: assign[n] = argmin_k SD(code_n, center_k), first min wins
void orc_assign(const float *Dm, const uint8_t *codes, int64_t N, const uint8_t *centers, int K, int M, int Ks,
                int32_t *assign, float *dist)
{
#pragma omp parallel for
    for (int64_t n = 0; n < N; ++n) {
        auto r = nearest_center(Dm, codes + n * M, centers, K, M, Ks);
        assign[n] = r.first;
        if (dist) dist[n] = r.second;
    }
}

// src/rii.h:108-156 Reconfigure + src/pqkmeans.cpp:46-133 fit + :177-191 init + :223-260 sparse voting +
// src/rii.h:335-359 UpdatePostingLists.  Outputs: centers (nlist, M) and the assignment of every code
// (posting list `no` = ascending ids with assign == no).
void orc_reconfigure(const float *cw, int M, int Ks, int Ds, const uint8_t *codes, int64_t N, int nlist, int iter,
                     uint8_t *centers_out, int32_t *assign_out)
{
    std::vector<float> Dm((size_t)M * Ks * Ks);
    orc_sym_matrices(cw, M, Ks, Ds, Dm.data());

    // (1) sampling, src/rii.h:115-124
    size_t Ns = std::min((size_t)N, (size_t)nlist * 100);
    std::vector<size_t> pick(N);
    std::iota(pick.begin(), pick.end(), 0);
    std::shuffle(pick.begin(), pick.end(), std::default_random_engine(123));
    pick.resize(Ns);
    std::vector<uint8_t> sample(Ns * M);
    for (size_t i = 0; i < Ns; ++i) std::memcpy(&sample[i * M], codes + pick[i] * M, M);

    // (2) PQk-means, src/pqkmeans.cpp:46-133
    std::vector<uint8_t> centers_new((size_t)nlist * M), centers_old;
    {   // InitializeCentersByRandomPicking :177-191
        std::vector<int> ids(Ns);
        std::iota(ids.begin(), ids.end(), 0);
        std::mt19937 random_engine(0);
        std::shuffle(ids.begin(), ids.end(), random_engine);
        for (int k = 0; k < nlist; ++k) std::memcpy(&centers_new[(size_t)k * M], &sample[(size_t)ids[k] * M], M);
    }
    std::vector<int32_t> assign(Ns);
    for (int itr = 0; itr < iter; ++itr) {
        centers_old = centers_new;
        orc_assign(Dm.data(), sample.data(), (int64_t)Ns, centers_old.data(), nlist, M, Ks, assign.data(), nullptr);
        if (itr != iter - 1) {                                           // :110
            std::vector<std::vector<size_t>> members(nlist);
            for (size_t n = 0; n < Ns; ++n) members[assign[n]].push_back(n);
            for (int k = 0; k < nlist; ++k) {
                if (members[k].empty()) continue;                        // :115-120 keep previous center
                for (int m = 0; m < M; ++m) {                            // ComputeCenterBySparseVoting :223-260
                    std::vector<int> hist(Ks, 0);
                    for (size_t id : members[k]) ++hist[sample[id * M + m]];
                    std::vector<float> vote(Ks, 0.f);
                    for (int k1 = 0; k1 < Ks; ++k1) {
                        int freq = hist[k1];
                        if (freq == 0) continue;
                        for (int k2 = 0; k2 < Ks; ++k2)
                            vote[k2] += (float)freq * Dm[((size_t)m * Ks + k1) * Ks + k2];
                    }
                    float min_dist = FLT_MAX;
                    int min_ks = -1;
                    for (int ks = 0; ks < Ks; ++ks)
                        if (vote[ks] < min_dist) { min_ks = ks; min_dist = vote[ks]; }
                    centers_new[(size_t)k * M + m] = (uint8_t)min_ks;
                }
            }
        }
    }
    std::memcpy(centers_out, centers_new.data(), (size_t)nlist * M);

    // (4) UpdatePostingLists(0, N), src/rii.h:335-359
    orc_assign(Dm.data(), codes, N, centers_out, nlist, M, Ks, assign_out, nullptr);
}

}  // extern "C"
