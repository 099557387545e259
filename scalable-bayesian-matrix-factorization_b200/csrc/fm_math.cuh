// fm_math.cuh -- the arithmetic of one coordinate draw of the general FM Gibbs sampler (fm.cu), SURVEY.md 8(f)-4.
// Restates libFM's fm_learn_mcmc.h ("[G]"): draw_w [G]:671-719, draw_v [G]:780-836, draw_w0 [G]:628-668, the group hyper
// draws [G]:931-1088.  Everything here is __host__ __device__ so that tools/fm_emulate.cu can run the kernels' exact formulas
// and schedule on the CPU (fp32, zero noise) against the fp64 checker in the CPU test suite.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace sbmf_fm {

enum Coord : int { COORD_W = 0, COORD_V = 1 };

// sums of one attribute column: hh = sum h^2, he = sum h * (e - theta * h) in the form libFM accumulates them.  The terms are
// formed from the fp32 caches; the SUMS are fp64: he cancels heavily on long columns (tens of thousands of cases), where an
// fp32 running sum loses the 1e-4 parity bar (measured with tools/fm_emulate.cu: 1e-2 on a 54,000-entry column).
struct ColSums {
    double hh, he;
};

// one entry (case c of the column): x = design value, e / q = the case's caches, theta = current w_j or v_jf
template <int COORD>
__host__ __device__ __forceinline__ void accumulate_entry(ColSums& s, float x, float e, float q, float theta)
{
    if (COORD == COORD_W) {          // [G]:675-680: w_mean += x (e - w x); w_sigma_sqr += x^2
        s.he += (double)x * (double)(e - theta * x);
        s.hh += (double)x * (double)x;
    } else {                         // [G]:785-793: h = x (q - x v); v_mean += h e; v_sigma_sqr += h^2
        const float h = x * (q - x * theta);
        s.he += (double)h * (double)e;
        s.hh += (double)h * (double)h;
    }
}

struct Posterior {
    double mean, var;
    bool degenerate;                 // [G]:686, 800: sigma_sqr NaN or inf => the coordinate is set to 0
};

// [G]:681-682 / 794-796: theta | rest ~ N(mean, var)
template <int COORD>
__host__ __device__ __forceinline__ Posterior posterior(double hh, double he, double theta, double alpha, double mu, double lambda)
{
    if (COORD == COORD_V) he -= theta * hh;                 // [G]:794
    Posterior p;
    p.var = 1.0 / (lambda + alpha * hh);
    p.mean = -p.var * (alpha * he - mu * lambda);
    p.degenerate = isnan(p.var) || isinf(p.var);
    return p;
}

// theta_new from the posterior and a standard normal z (z = 0: conditional mean); libFM's out-of-bounds rule [G]:697-710
__host__ __device__ __forceinline__ double settle(const Posterior& p, double z, double theta_old)
{
    double t = p.degenerate ? 0.0 : p.mean + sqrt(p.var) * z;
    if (isnan(t) || isinf(t)) t = theta_old;
    return t;
}

// caches of one case after theta_old -> theta_old - delta
template <int COORD>
__host__ __device__ __forceinline__ void apply_entry(float x, float& e, float& q, float theta_old, float delta)
{
    if (COORD == COORD_W) {          // [G]:712-717
        e -= x * delta;
    } else {                         // [G]:826-833
        const float h = x * (q - x * theta_old);
        q -= x * delta;
        e -= h * delta;
    }
}

// group hyper-parameters ([G]:931-1088) from S1 = sum theta, S2 = sum (theta - mu_old)^2 over the n attributes of the group
struct GroupPosterior {
    double lambda_shape, lambda_rate;      // lambda ~ Gamma(shape, rate)            [G]:984-988, 1066-1070
    double n_beta;                         // n + beta_0
    double mu_mean;                        // mu ~ N(mu_mean, 1 / (n_beta * lambda)) [G]:943-944, 1025-1026
};
__host__ __device__ __forceinline__ GroupPosterior group_posterior(double S1, double S2, double n, double mu_old, double alpha_0,
                                                                    double beta_0, double gamma_0, double mu_0)
{
    GroupPosterior g;
    g.lambda_shape = (alpha_0 + n + 1.0) / 2.0;
    g.lambda_rate = (beta_0 * (mu_old - mu_0) * (mu_old - mu_0) + gamma_0 + S2) / 2.0;
    g.n_beta = n + beta_0;
    g.mu_mean = (S1 + beta_0 * mu_0) / g.n_beta;
    return g;
}

// Maximal conflict-free runs of the attribute sequence (host; see sbmf_fm_plan_runs in sbmf_fm_cuda.h)
inline uint32_t plan_runs(uint32_t p, const uint32_t* next_attr, uint32_t* run_begin)
{
    uint32_t n = 0, limit = UINT32_MAX;
    for (uint32_t j = 0; j < p; ++j) {
        if (j == 0 || j >= limit) {
            run_begin[n++] = j;
            limit = UINT32_MAX;
        }
        if (next_attr[j] < limit) limit = next_attr[j];
    }
    run_begin[n] = p;
    return n;
}

}  // namespace sbmf_fm
