// tools/fm_emulate.cu -- host emulation of the general FM Gibbs schedule of csrc/fm.cu (zero noise): the SAME formulas
// (csrc/fm_math.cuh, shared with the kernels), the same precisions (fp32 caches and parameters, fp64 hyper-parameters), the same
// run partition (plan_runs) and the same two-phase treatment of a run -- every column of a run is reduced and drawn from the
// caches as they stand when the run starts, then all updates are applied -- executed sequentially on the CPU.  The CPU test
// suite compares it with the fp64 checker, which walks the attributes strictly one after the other like libFM: that validates
// the algebra and the claim that a run's columns commute, without a GPU.  Build: nvcc -shared (host code only).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "fm_math.cuh"

using namespace sbmf_fm;

namespace {
constexpr double ALPHA_0 = 1.0, GAMMA_0 = 1.0, BETA_0 = 1.0, MU_0 = 0.0;

struct Cols {
    std::vector<int64_t> ptr;
    std::vector<uint32_t> cs;
    std::vector<float> x;
};

void predict(uint32_t n, const int64_t* rp, const uint32_t* at, const float* x, uint32_t p, uint32_t K, int k0, int k1, const float* w, const float* V,
             double w0, std::vector<double>& out)
{
    out.assign(n, 0.0);
    std::vector<float> s(K), ss(K);
    for (uint32_t r = 0; r < n; ++r) {
        float acc = 0.f;
        std::fill(s.begin(), s.end(), 0.f);
        std::fill(ss.begin(), ss.end(), 0.f);
        for (int64_t k = rp[r]; k < rp[r + 1]; ++k) {
            if (k1) acc = fmaf(w[at[k]], x[k], acc);
            for (uint32_t f = 0; f < K; ++f) {
                const float d = V[(size_t)at[k] * K + f] * x[k];
                s[f] += d;
                ss[f] = fmaf(d, d, ss[f]);
            }
        }
        for (uint32_t f = 0; f < K; ++f) acc += 0.5f * (s[f] * s[f] - ss[f]);
        out[r] = (double)acc + (k0 ? w0 : 0.0);
    }
}

template <int COORD>
void do_run(uint32_t b, uint32_t en, const Cols& c, float* e, float* q, float* theta, uint32_t stride, uint32_t f, const uint32_t* group,
            const double* mu, const double* lambda, uint32_t hstride, double alpha)
{
    std::vector<float> told(en - b), delta(en - b);
    for (uint32_t j = b; j < en; ++j) {          // phase 1: all reductions and draws of the run
        const float t0 = theta[(size_t)j * stride + f];
        ColSums s{0.0, 0.0};
        for (int64_t k = c.ptr[j]; k < c.ptr[j + 1]; ++k) accumulate_entry<COORD>(s, c.x[k], e[c.cs[k]], COORD == COORD_V ? q[c.cs[k]] : 0.f, t0);
        const uint32_t g = group[j];
        const Posterior post = posterior<COORD>(s.hh, s.he, (double)t0, alpha, mu[(size_t)g * hstride + f], lambda[(size_t)g * hstride + f]);
        const float t1 = (float)settle(post, 0.0, (double)t0);
        theta[(size_t)j * stride + f] = t1;
        told[j - b] = t0;
        delta[j - b] = t0 - t1;
    }
    for (uint32_t j = b; j < en; ++j) {          // phase 2: all cache updates
        if (delta[j - b] == 0.f) continue;
        for (int64_t k = c.ptr[j]; k < c.ptr[j + 1]; ++k) {
            float ev = e[c.cs[k]], qv = COORD == COORD_V ? q[c.cs[k]] : 0.f;
            apply_entry<COORD>(c.x[k], ev, qv, told[j - b], delta[j - b]);
            e[c.cs[k]] = ev;
            if (COORD == COORD_V) q[c.cs[k]] = qv;
        }
    }
}

void hypers(const float* theta, uint32_t stride, uint32_t F, uint32_t p, uint32_t G, const uint32_t* group, const std::vector<uint32_t>& npg, double* mu,
            double* lambda, int do_multilevel)
{
    for (uint32_t f = 0; f < F; ++f) {
        std::vector<double> S1(G, 0.0), S2(G, 0.0);
        for (uint32_t j = 0; j < p; ++j) {
            const double t = theta[(size_t)j * stride + f], m = mu[(size_t)group[j] * F + f];
            S1[group[j]] += t;
            S2[group[j]] += (t - m) * (t - m);
        }
        for (uint32_t g = 0; g < G; ++g) {
            double& M = mu[(size_t)g * F + f];
            double& L = lambda[(size_t)g * F + f];
            if (!do_multilevel) {
                M = MU_0;
                continue;
            }
            const GroupPosterior gp = group_posterior(S1[g], S2[g], (double)npg[g], M, ALPHA_0, BETA_0, GAMMA_0, MU_0);
            const double l = gp.lambda_shape / gp.lambda_rate;
            if (!isnan(l) && !isinf(l)) L = l;
            const double nm = gp.mu_mean;
            if (!isnan(nm) && !isinf(nm)) M = nm;
        }
    }
}
}  // namespace

// v_io: [K][p] (libFM layout) in and out.  Returns the number of runs.
extern "C" int fm_emulate(uint32_t n, const int64_t* rp_in, const uint32_t* at_in, const float* x_in, const float* y, uint32_t nt, const int64_t* trp_in,
                          const uint32_t* tat, const float* tx, const float* ty, uint32_t p, uint32_t G, const uint32_t* group_in, uint32_t K, int k0, int k1,
                          int do_multilevel, double reg0, double regw, double regv, uint32_t iters, float* w_io, float* v_io, double* hyp_out /* w_mu[G],
                          w_lambda[G], v_mu[G*K], v_lambda[G*K], w0, alpha */, float* e_out, double* rmse_train, double* rmse_test, uint32_t* run_begin_out)
{
    std::vector<uint32_t> group(p, 0);
    if (group_in) group.assign(group_in, group_in + p);
    std::vector<uint32_t> npg(G, 0);
    for (uint32_t j = 0; j < p; ++j) npg[group[j]]++;
    // row form with ascending attributes + column form with ascending cases (what set_train builds on the device)
    std::vector<int64_t> rp(rp_in, rp_in + n + 1);
    std::vector<uint32_t> at(at_in, at_in + rp[n]);
    std::vector<float> x(x_in, x_in + rp[n]);
    for (uint32_t r = 0; r < n; ++r) {
        std::vector<std::pair<uint32_t, float>> row;
        for (int64_t k = rp[r]; k < rp[r + 1]; ++k) row.push_back({at[k], x[k]});
        std::stable_sort(row.begin(), row.end(), [](const std::pair<uint32_t, float>& a, const std::pair<uint32_t, float>& b) { return a.first < b.first; });
        for (int64_t k = rp[r]; k < rp[r + 1]; ++k) {
            at[k] = row[k - rp[r]].first;
            x[k] = row[k - rp[r]].second;
        }
    }
    Cols c;
    c.ptr.assign((size_t)p + 1, 0);
    for (int64_t k = 0; k < rp[n]; ++k) c.ptr[at[k] + 1]++;
    for (uint32_t j = 0; j < p; ++j) c.ptr[j + 1] += c.ptr[j];
    c.cs.resize(rp[n]);
    c.x.resize(rp[n]);
    {
        std::vector<int64_t> fill(c.ptr.begin(), c.ptr.end() - 1);
        for (uint32_t r = 0; r < n; ++r)
            for (int64_t k = rp[r]; k < rp[r + 1]; ++k) {
                c.cs[fill[at[k]]] = r;
                c.x[fill[at[k]]++] = x[k];
            }
    }
    std::vector<uint32_t> next_attr(p, UINT32_MAX), run_begin((size_t)p + 1);
    for (uint32_t r = 0; r < n; ++r)
        for (int64_t k = rp[r]; k + 1 < rp[r + 1]; ++k) next_attr[at[k]] = std::min(next_attr[at[k]], at[k + 1]);
    const uint32_t nruns = plan_runs(p, next_attr.data(), run_begin.data());
    if (run_begin_out) memcpy(run_begin_out, run_begin.data(), ((size_t)nruns + 1) * 4);

    const uint32_t F = std::max<uint32_t>(K, 1);
    std::vector<float> w(w_io, w_io + p), V((size_t)p * F, 0.f), e(n), q(n, 0.f);
    for (uint32_t f = 0; f < K; ++f)
        for (uint32_t j = 0; j < p; ++j) V[(size_t)j * K + f] = v_io[(size_t)f * p + j];
    std::vector<double> w_mu(G, 0.0), w_lambda(G, regw), v_mu((size_t)G * F, 0.0), v_lambda((size_t)G * F, regv), pred, pred_sum(nt, 0.0);
    double w0 = 0.0, alpha = 1.0;
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
    for (uint32_t r = 0; r < n; ++r) {
        mn = std::min(mn, y[r]);
        mx = std::max(mx, y[r]);
    }
    predict(n, rp.data(), at.data(), x.data(), p, K, k0, k1, w.data(), V.data(), w0, pred);
    for (uint32_t r = 0; r < n; ++r) e[r] = (float)(pred[r] - (double)y[r]);

    for (uint32_t it = 0; it < iters; ++it) {
        double S1 = 0.0, S2 = 0.0;
        for (uint32_t r = 0; r < n; ++r) {
            S1 += (double)e[r];
            S2 += (double)e[r] * (double)e[r];
        }
        alpha = do_multilevel ? ((ALPHA_0 + (double)n) / 2.0) / ((GAMMA_0 + S2) / 2.0) : ALPHA_0;
        if (k0) {
            const double var = 1.0 / (reg0 + alpha * (double)n);
            const double nw = -var * (alpha * (S1 - (double)n * w0));
            const float d = (float)(w0 - nw);
            w0 = nw;
            for (uint32_t r = 0; r < n; ++r) e[r] -= d;
        }
        if (k1) {
            hypers(w.data(), 1, 1, p, G, group.data(), npg, w_mu.data(), w_lambda.data(), do_multilevel);
            for (uint32_t r = 0; r < nruns; ++r)
                do_run<COORD_W>(run_begin[r], run_begin[r + 1], c, e.data(), q.data(), w.data(), 1, 0, group.data(), w_mu.data(), w_lambda.data(), 1, alpha);
        }
        if (K) {
            hypers(V.data(), K, K, p, G, group.data(), npg, v_mu.data(), v_lambda.data(), do_multilevel);
            for (uint32_t f = 0; f < K; ++f) {
                for (uint32_t r = 0; r < n; ++r) {
                    float s = 0.f;
                    for (int64_t k = rp[r]; k < rp[r + 1]; ++k) s = fmaf(V[(size_t)at[k] * K + f], x[k], s);
                    q[r] = s;
                }
                for (uint32_t r = 0; r < nruns; ++r)
                    do_run<COORD_V>(run_begin[r], run_begin[r + 1], c, e.data(), q.data(), V.data(), K, f, group.data(), v_mu.data(), v_lambda.data(), K, alpha);
            }
        }
        predict(n, rp.data(), at.data(), x.data(), p, K, k0, k1, w.data(), V.data(), w0, pred);
        double a = 0.0;
        for (uint32_t r = 0; r < n; ++r) {
            const double err = fmax((double)mn, fmin((double)mx, pred[r])) - (double)y[r];
            a += err * err;
            e[r] = (float)(pred[r] - (double)y[r]);
        }
        rmse_train[it] = sqrt(a / n);
        predict(nt, trp_in, tat, tx, p, K, k0, k1, w.data(), V.data(), w0, pred);
        double b = 0.0;
        for (uint32_t r = 0; r < nt; ++r) {
            pred_sum[r] += fmax((double)mn, fmin((double)mx, pred[r]));
            const double err = fmax((double)mn, fmin((double)mx, pred_sum[r] * (1.0 / ((double)it + 1.0)))) - (double)ty[r];
            b += err * err;
        }
        rmse_test[it] = sqrt(b / nt);
    }
    memcpy(w_io, w.data(), (size_t)p * 4);
    for (uint32_t f = 0; f < K; ++f)
        for (uint32_t j = 0; j < p; ++j) v_io[(size_t)f * p + j] = V[(size_t)j * K + f];
    double* h = hyp_out;
    memcpy(h, w_mu.data(), G * 8); h += G;
    memcpy(h, w_lambda.data(), G * 8); h += G;
    memcpy(h, v_mu.data(), (size_t)G * K * 8); h += (size_t)G * K;
    memcpy(h, v_lambda.data(), (size_t)G * K * 8); h += (size_t)G * K;
    h[0] = w0;
    h[1] = alpha;
    memcpy(e_out, e.data(), (size_t)n * 4);
    return (int)nruns;
}
