// synth_host.cpp -- the same family of synthetic rating matrices as synth.cu, for shapes whose user x item grid is far too
// large to visit pair by pair (the 10M x 1M, 1B-rating configuration of BASELINE.json: 1e13 pairs).  Host threads, plain C++.
//
// Model (identical to synth.cu): pair (i, j) is rated with probability p_ij = min(1, c * a_i * b_j),
// a_i = (rank_u(i) + 1)^-s_user, b_j = (rank_v(j) + 1)^-s_item over random permutations of the ids, c solved so that the
// expected number of ratings is n_ratings; ratings from a planted rank-16 model rounded to half stars in [0.5, 5]; output
// sorted by (user, item); each pair goes to the test set with probability test_frac.
//
// Sampling is sparse: for one user, item ranks t with c*a_i*b_t >= 1 are all taken; the remaining ranks are cut into bands
// [lo, 2*lo) inside which p varies by at most 2^s_item, candidates are drawn by geometric skips at the band's largest p and
// thinned by p_t / p_max.  Cost is O(ratings), independent of users x items.  Every user has its own counter-based random
// stream, so the matrix does not depend on the number of threads.  Bench / test input only (the reference ships no generator).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sbmf_cuda.h"

namespace {

constexpr int RANK = 16;
enum Site : uint64_t { S_PERM_U = 1, S_PERM_V = 2, S_PAIR = 3, S_SPLIT = 4, S_FACT_U = 5, S_FACT_V = 6, S_BIAS_U = 7, S_BIAS_V = 8, S_NOISE = 9 };

inline uint64_t mix64(uint64_t x)   // splitmix64 finaliser
{
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
inline uint64_t key(uint64_t seed, uint64_t site, uint64_t a, uint64_t b) { return mix64(mix64(mix64(seed ^ (site << 56)) ^ a) ^ (b * 0xd1342543de82ef95ull)); }
inline double u01(uint64_t x) { return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0); }   // (0, 1)
inline float normal_from(uint64_t k)   // Box-Muller on two hashes of k
{
    const double u1 = u01(mix64(k)), u2 = u01(mix64(k ^ 0xa5a5a5a5a5a5a5a5ull));
    return (float)(sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2));
}

struct Stream {   // per-user sequence
    uint64_t s;
    double next() { return u01(mix64(s++)); }
};

struct Gen {
    uint32_t I, J;
    double su, sv, c, test_frac;
    uint64_t seed;
    std::vector<uint32_t> rank_u;        // user id -> rank
    std::vector<uint32_t> item_of_rank;  // rank -> item id
    std::vector<double> bs, bp;          // b by rank, prefix sums (bp[t] = sum_{t' < t} bs[t'])
    std::vector<float> Q, vb;            // planted item factors / biases by item id

    // expected number of rated items of a user with scaled weight x = c * a_i
    double expect_user(double x) const
    {
        uint32_t t0 = 0;
        if (x >= 1.0) {
            const double r = pow(x, 1.0 / sv);   // ranks t with (t+1)^-sv >= 1/x  <=>  t+1 <= x^(1/sv)
            t0 = r >= (double)J ? J : (uint32_t)r;
            while (t0 < J && x * bs[t0] >= 1.0) ++t0;
            while (t0 > 0 && x * bs[t0 - 1] < 1.0) --t0;
        }
        return (double)t0 + x * (bp[J] - bp[t0]);
    }

    // item ranks rated by user i, in increasing rank order, appended to out
    template <class F>
    void sample_user(uint32_t i, F&& emit) const
    {
        const double x = c * pow((double)rank_u[i] + 1.0, -su);
        Stream st{key(seed, S_PAIR, i, 0)};
        uint32_t t = 0;
        while (t < J && x * bs[t] >= 1.0) emit(t++);   // saturated head (popular items of a heavy user)
        while (t < J) {
            const uint32_t lo = t;
            const uint64_t hi64 = std::min<uint64_t>((uint64_t)J, 2ull * ((uint64_t)lo + 1));
            const uint32_t hi = (uint32_t)hi64;
            const double pmax = x * bs[lo];   // < 1
            const double lq = log1p(-pmax);
            double pos = (double)lo;
            for (;;) {
                pos += floor(log(st.next()) / lq);   // failures before the next candidate
                if (pos >= (double)hi) break;
                const uint32_t tc = (uint32_t)pos;
                if (st.next() * bs[lo] < bs[tc]) emit(tc);   // thinning by p_t / p_max
                pos += 1.0;
            }
            t = hi;
        }
    }
};

std::string g_err;

// dynamic chunks of [0, n) over `threads` host threads; body(begin, end, thread id)
void parallel_for(uint32_t n, uint32_t chunk, int threads, const std::function<void(uint32_t, uint32_t, int)>& body)
{
    std::atomic<uint64_t> next{0};
    auto work = [&](int tid) {
        for (;;) {
            const uint64_t b = next.fetch_add(chunk);
            if (b >= n) return;
            body((uint32_t)b, (uint32_t)std::min<uint64_t>(n, b + chunk), tid);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
}
}  // namespace

extern "C" const char* sbmf_cuda_synth_host_last_error(void) { return g_err.c_str(); }

extern "C" void sbmf_cuda_synth_host_free(void* p) { free(p); }

extern "C" int sbmf_cuda_synth_host_generate(const sbmf_synth_spec* spec, int threads, uint64_t* n_train, uint64_t* n_test, uint32_t** train_user,
                                        uint32_t** train_item, float** train_rating, uint32_t** test_user, uint32_t** test_item,
                                        float** test_rating)
{
    if (!spec || !n_train || !n_test || !train_user || !train_item || !train_rating || !test_user || !test_item || !test_rating ||
        spec->num_users == 0 || spec->num_items == 0 || !(spec->s_item > 0.0) || !(spec->s_user >= 0.0) || spec->test_frac < 0.0 ||
        spec->test_frac >= 1.0 || (double)spec->n_ratings > 0.5 * (double)spec->num_users * (double)spec->num_items) {
        g_err = "synth_host_generate: invalid spec (needs s_item > 0 and n_ratings <= half of users x items)";
        return SBMF_ERR_INVALID;
    }
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    Gen g;
    g.I = spec->num_users; g.J = spec->num_items; g.su = spec->s_user; g.sv = spec->s_item; g.seed = spec->seed; g.test_frac = spec->test_frac;
    const uint32_t I = g.I, J = g.J;
    try {
        // random permutations of the ids: sort by hashed key (ties are impossible to matter: the key includes the id)
        {
            std::vector<std::pair<uint64_t, uint32_t>> ku(I);
            for (uint32_t i = 0; i < I; ++i) ku[i] = {key(g.seed, S_PERM_U, i, 0), i};
            std::sort(ku.begin(), ku.end());
            g.rank_u.resize(I);
            for (uint32_t r = 0; r < I; ++r) g.rank_u[ku[r].second] = r;
        }
        {
            std::vector<std::pair<uint64_t, uint32_t>> kv(J);
            for (uint32_t j = 0; j < J; ++j) kv[j] = {key(g.seed, S_PERM_V, j, 0), j};
            std::sort(kv.begin(), kv.end());
            g.item_of_rank.resize(J);
            for (uint32_t r = 0; r < J; ++r) g.item_of_rank[r] = kv[r].second;
        }
        g.bs.resize(J);
        g.bp.resize((size_t)J + 1);
        g.bp[0] = 0.0;
        for (uint32_t t = 0; t < J; ++t) {
            g.bs[t] = pow((double)t + 1.0, -g.sv);
            g.bp[t + 1] = g.bp[t] + g.bs[t];
        }
        // c: bisection (in log space) on the exact expectation sum_i E[deg_i]
        {
            std::vector<double> part(threads);
            auto expected = [&](double c) {
                std::fill(part.begin(), part.end(), 0.0);
                parallel_for(I, 65536, threads, [&](uint32_t b, uint32_t e, int tid) {
                    double s = 0.0;
                    for (uint32_t r = b; r < e; ++r) s += g.expect_user(c * pow((double)r + 1.0, -g.su));
                    part[tid] += s;
                });
                return std::accumulate(part.begin(), part.end(), 0.0);
            };
            double lo = -60.0, hi = 60.0;   // log c
            for (int it = 0; it < 60; ++it) {
                const double mid = 0.5 * (lo + hi);
                if (expected(exp(mid)) < (double)spec->n_ratings) lo = mid;
                else hi = mid;
            }
            g.c = exp(0.5 * (lo + hi));
        }
        // planted model
        g.Q.resize((size_t)J * RANK);
        g.vb.resize(J);
        parallel_for(J, 4096, threads, [&](uint32_t b, uint32_t e, int) {
            for (uint32_t j = b; j < e; ++j) {
                for (int k = 0; k < RANK; ++k) g.Q[(size_t)j * RANK + k] = 0.3f * normal_from(key(g.seed, S_FACT_V, j, k));
                g.vb[j] = 0.3f * normal_from(key(g.seed, S_BIAS_V, j, 0));
            }
        });
        const uint64_t test_thr = (uint64_t)(g.test_frac * 18446744073709551615.0);
        auto is_test = [&](uint32_t i, uint32_t j) { return g.test_frac > 0.0 && key(g.seed, S_SPLIT, i, j) < test_thr; };

        // pass 1: per-user (train, test) counts
        std::vector<uint64_t> otr((size_t)I + 1), ote((size_t)I + 1);
        parallel_for(I, 256, threads, [&](uint32_t b, uint32_t e, int) {
            for (uint32_t i = b; i < e; ++i) {
                uint64_t tr = 0, te = 0;
                g.sample_user(i, [&](uint32_t t) {
                    if (is_test(i, g.item_of_rank[t])) ++te;
                    else ++tr;
                });
                otr[i + 1] = tr;
                ote[i + 1] = te;
            }
        });
        otr[0] = ote[0] = 0;
        for (uint32_t i = 0; i < I; ++i) {
            otr[i + 1] += otr[i];
            ote[i + 1] += ote[i];
        }
        const uint64_t ntr = otr[I], nte = ote[I];
        uint32_t *tu = (uint32_t*)malloc((ntr + 1) * 4), *ti = (uint32_t*)malloc((ntr + 1) * 4), *su = (uint32_t*)malloc((nte + 1) * 4),
                 *si = (uint32_t*)malloc((nte + 1) * 4);
        float *trr = (float*)malloc((ntr + 1) * 4), *sr = (float*)malloc((nte + 1) * 4);
        if (!tu || !ti || !trr || !su || !si || !sr) {
            free(tu); free(ti); free(trr); free(su); free(si); free(sr);
            g_err = "synth_host_generate: out of host memory";
            return SBMF_ERR_NOMEM;
        }
        // pass 2: the same streams again, now sorted by item id and written at the user's offsets
        parallel_for(I, 256, threads, [&](uint32_t b, uint32_t e, int) {
            std::vector<uint32_t> items;
            for (uint32_t i = b; i < e; ++i) {
                items.clear();
                g.sample_user(i, [&](uint32_t t) { items.push_back(g.item_of_rank[t]); });
                std::sort(items.begin(), items.end());
                float P[RANK];
                for (int k = 0; k < RANK; ++k) P[k] = 0.3f * normal_from(key(g.seed, S_FACT_U, i, k));
                const float ub = 0.3f * normal_from(key(g.seed, S_BIAS_U, i, 0));
                uint64_t ptr = otr[i], pte = ote[i];
                for (uint32_t j : items) {
                    float x = 3.5f + ub + g.vb[j] + 0.8f * normal_from(key(g.seed, S_NOISE, i, j));
                    const float* q = &g.Q[(size_t)j * RANK];
                    for (int k = 0; k < RANK; ++k) x += P[k] * q[k];
                    float r = roundf(2.0f * x) * 0.5f;
                    r = r < 0.5f ? 0.5f : (r > 5.0f ? 5.0f : r);
                    if (is_test(i, j)) {
                        su[pte] = i; si[pte] = j; sr[pte] = r; ++pte;
                    } else {
                        tu[ptr] = i; ti[ptr] = j; trr[ptr] = r; ++ptr;
                    }
                }
            }
        });
        *n_train = ntr; *n_test = nte;
        *train_user = tu; *train_item = ti; *train_rating = trr;
        *test_user = su; *test_item = si; *test_rating = sr;
    } catch (const std::bad_alloc&) {
        g_err = "synth_host_generate: out of host memory";
        return SBMF_ERR_NOMEM;
    }
    return SBMF_OK;
}
