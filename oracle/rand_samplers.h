/*
 * rand_samplers.h -- restatement of the reference samplers (src/util/random.h, "[R]") on glibc rand().
 * TEST INFRASTRUCTURE ONLY.  Shared by sbmf_oracle.c and shim_random.h so both consume the rand()
 * stream identically to the reference: uniform [R]:174-176, Leva normal [R]:150-164, Marsaglia-Tsang
 * gamma [R]:118-144.
 */
#ifndef SBMF_RAND_SAMPLERS_H_
#define SBMF_RAND_SAMPLERS_H_
#include <math.h>
#include <stdlib.h>

static inline double ran_uniform_rand(void) { return rand() / ((double)RAND_MAX + 1); }   /* [R]:174-176 */

static inline double ran_gaussian_leva(void)                                               /* [R]:150-164 */
{
    double u, v, x, y, Q;
    do {
        do {
            u = ran_uniform_rand();
        } while (u == 0.0);
        v = 1.7156 * (ran_uniform_rand() - 0.5);
        x = u - 0.449871;
        y = fabs(v) + 0.386595;
        Q = x * x + y * (0.19600 * y - 0.25472 * x);
        if (Q < 0.27597) break;
    } while ((Q > 0.27846) || ((v * v) > (-4.0 * u * u * log(u))));
    return v / u;
}

static inline double ran_gamma_mt_rand(double alpha)                                       /* [R]:118-144 */
{
    if (alpha < 1.0) {
        double u;
        do {
            u = ran_uniform_rand();
        } while (u == 0.0);
        return ran_gamma_mt_rand(alpha + 1.0) * pow(u, 1.0 / alpha);
    } else {
        double d, c, x, v, u;
        d = alpha - 1.0 / 3.0;
        c = 1.0 / sqrt(9.0 * d);
        do {
            do {
                x = ran_gaussian_leva();
                v = 1.0 + c * x;
            } while (v <= 0.0);
            v = v * v * v;
            u = ran_uniform_rand();
        } while ((u >= (1.0 - 0.0331 * (x * x) * (x * x))) && (log(u) >= (0.5 * x * x + d * (1.0 - v + log(v)))));
        return d * v;
    }
}

#endif
