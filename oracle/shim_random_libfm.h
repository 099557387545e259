/*
 * shim_random_libfm.h -- shim_random.h plus the three sampler-header functions only libFM's classification branch names
 * (fm_learn_mcmc_simultaneous.h:165-205; never reached with -task r).  TEST INFRASTRUCTURE ONLY.
 *
 * The UNMODIFIED reference libFM (src/libfm/libfm.cpp, "[L]") is compiled against it with
 *     g++ -DRANDOM_H_ -include oracle/shim_random_libfm.h -I/root/reference/src/libfm /root/reference/src/libfm/libfm.cpp
 * so that its `-method mcmc` run logs every sampler argument, can run in the zero-noise mode and repeats under
 * SBMF_SHIM_SEED (see shim_random.h).
 */
#ifndef SBMF_SHIM_RANDOM_LIBFM_H_
#define SBMF_SHIM_RANDOM_LIBFM_H_
#include "shim_random.h"
inline double cdf_gaussian(double) { fprintf(stderr, "shim: classification sampler called\n"); abort(); }
inline double cdf_gaussian(double, double, double) { fprintf(stderr, "shim: classification sampler called\n"); abort(); }
inline double ran_left_tgaussian(double, double, double) { fprintf(stderr, "shim: classification sampler called\n"); abort(); }
inline double ran_right_tgaussian(double, double, double) { fprintf(stderr, "shim: classification sampler called\n"); abort(); }
#endif
