/*
 * shim_random.h -- drop-in replacement for the reference's src/util/random.h.  TEST INFRASTRUCTURE ONLY.
 *
 * The UNMODIFIED reference gibbs_sbpmf2.cpp is compiled against this header with
 *     g++ -DRANDOM_H_ -include oracle/shim_random.h -I/root/reference/src/libfm /root/reference/gibbs_sbpmf2.cpp
 * (-DRANDOM_H_ disables the body of the reference's own random.h; no reference source is edited or copied).
 * It gives the reference two things it does not have:
 *   SBMF_SHIM_LOG=<path>    append {tag, a, b} (3 doubles; tag 0 = ran_gaussian(mean,stdev), 1 = ran_gamma(alpha,beta))
 *                           for every sampler call, in call order -- the full posterior-parameter trajectory.
 *   SBMF_SHIM_MODE=zero     the zero-noise definition of SURVEY.md 8(c): ran_gaussian(m,s) -> m, ran_gamma(a,b) -> a/b,
 *                           except the first SBMF_SHIM_LIVE_INIT two-argument gaussian calls (factor initialisation,
 *                           [T]:239-250), which stay live so the factors are not identically zero.
 *   SBMF_SHIM_SEED=<n>      srand(n) before the first sampler call.  [T] and [S] never seed (glibc seed 1, the default here);
 *                           libFM's main() calls srand(time(NULL)) ([L]:124-125), which this overrides so its runs repeat.
 * In the default mode (rand) the draws are the reference's own algorithms (rand_samplers.h restates [R]).
 */
#ifndef SBMF_SHIM_RANDOM_H_
#define SBMF_SHIM_RANDOM_H_
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <assert.h>
#include "rand_samplers.h"

struct sbmf_shim_state {
    int zero;
    long live_init;
    long n_gauss2;
    FILE* log;
    int seed_pending;
    unsigned seed;
    sbmf_shim_state() : zero(0), live_init(0), n_gauss2(0), log(NULL), seed_pending(0), seed(1)
    {
        const char* sd = getenv("SBMF_SHIM_SEED");
        if (sd && sd[0]) { seed_pending = 1; seed = (unsigned)strtoul(sd, NULL, 10); }
        const char* m = getenv("SBMF_SHIM_MODE");
        zero = (m && !strcmp(m, "zero"));
        const char* li = getenv("SBMF_SHIM_LIVE_INIT");
        live_init = li ? atol(li) : 0;
        const char* lp = getenv("SBMF_SHIM_LOG");
        if (lp && lp[0]) log = fopen(lp, "wb");
    }
    ~sbmf_shim_state() { if (log) fclose(log); }
    void first_use() { if (seed_pending) { srand(seed); seed_pending = 0; } }
    void rec(double tag, double a, double b)
    {
        /* flushed per record: the reference can die in its own cleanup ([T]:648-666 delete[]s never-allocated rows
           once operator[] at [T]:520/568 has inserted their ids), and static destructors do not run then */
        if (log) { double r[3] = {tag, a, b}; fwrite(r, sizeof(double), 3, log); fflush(log); }
    }
};
static sbmf_shim_state sbmf_shim;

inline double ran_uniform() { return ran_uniform_rand(); }
inline double ran_gaussian() { return ran_gaussian_leva(); }
inline double ran_gaussian(double mean, double stdev)
{
    sbmf_shim.first_use();
    sbmf_shim.rec(0.0, mean, stdev);
    long idx = sbmf_shim.n_gauss2++;
    if (sbmf_shim.zero && idx >= sbmf_shim.live_init) return mean;
    if ((stdev == 0.0) || (std::isnan(stdev))) return mean;
    return mean + stdev * ran_gaussian_leva();
}
inline double ran_gamma(double alpha) { return ran_gamma_mt_rand(alpha); }
inline double ran_gamma(double alpha, double beta)
{
    sbmf_shim.first_use();
    sbmf_shim.rec(1.0, alpha, beta);
    if (sbmf_shim.zero) return alpha / beta;
    return ran_gamma_mt_rand(alpha) / beta;
}
inline double ran_exp() { return -std::log(1 - ran_uniform()); }
inline bool ran_bernoulli(double p) { return (ran_uniform() < p); }
#endif
