/*
 * sbmf_oracle.h -- CPU restatement of the SBMF Gibbs sweep.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle for the CUDA path.  It restates, in plain C and fp64, the
 * algorithm of the reference's top-level gibbs_sbpmf2.cpp ("[T]" in SURVEY.md) and the
 * samplers of src/util/random.h ("[R]").  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product path
 * (libsbmf_cuda.so) never links, loads or calls anything in oracle/.
 *
 * Pinning (see oracle/README.md): with noise mode SBMF_ORACLE_NOISE_RAND the restatement
 * reproduces the 100 "rmse is" values printed by the UNMODIFIED reference on ML-100K
 * (tests/golden/ref_ml100k_K20_T100_rmse.txt) and the full (mean, stdev) / (shape, rate)
 * argument stream of every sampler call made by the reference (captured by compiling the
 * unmodified [T] against oracle/shim_random.h).
 */
#ifndef SBMF_ORACLE_H_
#define SBMF_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sbmf_oracle sbmf_oracle;

enum {
    SBMF_ORACLE_NOISE_RAND = 0,   /* glibc rand() + Leva / Marsaglia-Tsang exactly as [R]:118-176 */
    SBMF_ORACLE_NOISE_ZERO = 1,   /* ran_gaussian(m,s) -> m ; ran_gamma(a,b) -> a/b   */
    SBMF_ORACLE_NOISE_PHILOX = 2  /* counter-based Philox4x32-10 streams shared with the device */
};

enum {
    SBMF_ORACLE_STDEV_REF = 0,    /* x = m + (1/lambda) z : the reference passes a variance as stdev */
    SBMF_ORACLE_STDEV_SQRT = 1    /* x = m + sqrt(1/lambda) z */
};

/* Philox stream ("site") ids.  Shared with csrc/philox.cuh -- keep in sync. */
enum {
    SBMF_SITE_INIT_U = 0, SBMF_SITE_INIT_V = 1,
    SBMF_SITE_U = 2, SBMF_SITE_V = 3,
    SBMF_SITE_BI = 4, SBMF_SITE_BJ = 5,
    SBMF_SITE_MU_BI = 6, SBMF_SITE_MU_BJ = 7,
    SBMF_SITE_SIGMA_BI = 8, SBMF_SITE_SIGMA_BJ = 9,
    SBMF_SITE_SIGMA_U = 10, SBMF_SITE_MU_U = 11,
    SBMF_SITE_SIGMA_V = 12, SBMF_SITE_MU_V = 13,
    SBMF_SITE_ALPHA = 14, SBMF_SITE_SIGMA_B0 = 15, SBMF_SITE_MU_B0 = 16, SBMF_SITE_B0 = 17
};

/* COO in file order; ids 0-based.  num_users/num_items = 1+max id over train U test ([T]:151-153). */
sbmf_oracle* sbmf_oracle_create(uint64_t n, const uint32_t* user, const uint32_t* item, const double* rating,
                                uint64_t nt, const uint32_t* tuser, const uint32_t* titem, const double* trating,
                                uint32_t num_users, uint32_t num_items, uint32_t K,
                                int noise_mode, int stdev_mode, uint64_t seed);
void sbmf_oracle_destroy(sbmf_oracle*);

/* 0 (default) = the top-level gibbs_sbpmf2.cpp ([T]); 1 = src/libfm/gibbs_sbpmf2.cpp ([S]) as committed: no biases / global
   mean, Normal-Gamma factor hyper-prior, tau ~ G(a0 + N/2, b0 + sum e^2 / 2), including the slip at [S]:412; 2 = [S] with the
   slip corrected.  Call before init_factors. */
void sbmf_oracle_set_variant(sbmf_oracle*, int variant);

/* glibc srand(); the reference never calls it (seed 1). */
void sbmf_oracle_srand(unsigned seed);

/* U0 is [I][K] row-major, V0 is [K][J] (dimension-major) like [T]:229-250.  NULL => draw
   0.1*N(0,1) in [T]'s order (all of U, i outer/k inner; then V, k outer/j inner). */
void sbmf_oracle_init_factors(sbmf_oracle*, const double* U0, const double* V0, double init_stdev);

/* Run n sweeps of [T]:335-637; rmse_out[n] receives the running-posterior-mean test RMSE
   printed by [T]:635; rmse_sweep_out (may be NULL) the RMSE of that sweep's own prediction. */
void sbmf_oracle_sweep(sbmf_oracle*, uint32_t n, double* rmse_out, double* rmse_sweep_out);

/* Optional call log: every sampler call appends {tag(0=gauss,1=gamma), a, b} as 3 doubles. */
int sbmf_oracle_set_log(sbmf_oracle*, const char* path);

/* State getters (copy out). */
void sbmf_oracle_get_U(const sbmf_oracle*, double* U /*[I][K]*/);
void sbmf_oracle_get_V(const sbmf_oracle*, double* V /*[K][J]*/);
void sbmf_oracle_get_bias(const sbmf_oracle*, double* b_i, double* b_j);
void sbmf_oracle_get_bias_hypers(const sbmf_oracle*, double* mu_b_i, double* sigma_b_i, double* mu_b_j, double* sigma_b_j);
void sbmf_oracle_get_dim_hypers(const sbmf_oracle*, double* sigma_u, double* mu_u, double* sigma_v, double* mu_v);
/* scalars[4] = {b_0, alpha, mu_b_0, sigma_b_0} */
void sbmf_oracle_get_scalars(const sbmf_oracle*, double* scalars);
void sbmf_oracle_get_E(const sbmf_oracle*, double* E /*[N] file order*/);
void sbmf_oracle_get_pred_mean(const sbmf_oracle*, double* pred /*[Nt]*/);

/* Integer layout restatement ([T]:156-221): stable-by-file-order jagged rows flattened.
   row_ptr[I+1], col[N] (item of each CSR slot), csr_id[N] (rating index n of each CSR slot),
   col_ptr[J+1], row[N] (user of each CSC slot), csc_id[N] (rating index of each CSC slot). */
void sbmf_oracle_get_layout(const sbmf_oracle*, int64_t* row_ptr, uint32_t* col, uint64_t* csr_id,
                            int64_t* col_ptr, uint32_t* row, uint64_t* csc_id);

/* Triple-file reader with [T]:35-73 semantics (sscanf "%u%c%u%c%lf", line counts iff >= 5
   conversions).  Two-call protocol: pass NULL arrays to count. Returns number of ratings, or -1. */
int64_t sbmf_oracle_read_triples(const char* path, uint32_t* user, uint32_t* item, double* rating,
                                 uint32_t* user_max, uint32_t* item_max);

/* Philox4x32-10 restatement (Salmon et al. 2011), exposed for known-answer tests. */
void sbmf_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* The device's standard-normal for (site,row,c1,sweep); fp32 Box-Muller. */
float sbmf_oracle_philox_normal_f32(uint64_t seed, uint32_t site, uint32_t row, uint32_t c1, uint32_t sweep);
double sbmf_oracle_philox_normal_f64(uint64_t seed, uint32_t site, uint32_t row, uint32_t c1, uint32_t sweep);

#ifdef __cplusplus
}
#endif
#endif
