/*
 * fm_oracle.h -- CPU restatement of libFM's MCMC learner for regression.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates, in plain C and fp64 (design-matrix values and targets in fp32 like libFM's FM_FLOAT / DATA_FLOAT), the path
 *   libfm.cpp ("[L]") : model + learner set-up                       [L]:371-388, 411-419, 481-520
 *   fm_core/fm_model.h: fm_model::init (v ~ N(init_mean, init_stdev)) fm_model.h:92-102
 *   src/fm_learn_mcmc.h ("[G]")             : predict_data_and_write_to_eterms 117-349, add_main_q 384-409, draw_all 411-626,
 *                                             draw_w0 628-668, draw_w 671-719, draw_v 780-836, draw_alpha 901-929,
 *                                             draw_w_mu 931-968, draw_w_lambda 970-1007, draw_v_mu 1011-1049,
 *                                             draw_v_lambda 1051-1088, init 1092-1113
 *   src/fm_learn_mcmc_simultaneous.h ("[GS]"): _learn 50-262 (regression branch), _evaluate 307-326
 * without relations (`--relation` blocks, [G]:57-64, are not restated) and without the classification task.
 * The samplers are rand_samplers.h (= src/util/random.h on glibc rand()).
 *
 * Pinning (tests/test_fm_oracle.py): the restatement reproduces the "#Iter= i Train= Test=" values printed by the UNMODIFIED
 * libFM built from /root/reference (oracle/_ref/libFM_shim, `make -C oracle ref`) and the sha256 of its complete sampler-
 * argument stream (shim_random.h log format), live under SBMF_SHIM_SEED and in zero-noise mode, on a matrix-factorisation
 * fixture and on a general fixture with real-valued, multi-hot and grouped attributes (tests/golden/make_fm_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may load it; the product never does.
 */
#ifndef SBMF_FM_ORACLE_H_
#define SBMF_FM_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fm_oracle fm_oracle;

enum {
    FM_ORACLE_NOISE_RAND = 0,   /* glibc rand() + Leva / Marsaglia-Tsang exactly as random.h */
    FM_ORACLE_NOISE_ZERO = 1    /* ran_gaussian(m,s) -> m ; ran_gamma(a,b) -> a/b; a rand-drawn init stays live */
};

typedef struct {
    uint32_t num_attr;        /* [L]:326: max(train.num_feature, test.num_feature) + 1 = largest id + 2 in this fork */
    uint32_t num_groups;      /* meta->num_attr_groups; 1 without -meta */
    uint32_t K;               /* -dim k0,k1,K */
    int32_t k0, k1;
    int32_t do_sample;        /* -do_sampling ([L]:418) */
    int32_t do_multilevel;    /* -do_multilevel ([L]:419) */
    int32_t noise;            /* FM_ORACLE_NOISE_* */
    double init_stdev;        /* -init_stdev, default 0.1 ([L]:127) */
    double reg0, regw, regv;  /* -regular ([L]:485-505); w_lambda / v_lambda start at regw / regv */
} fm_oracle_config;

/* Design matrices in row form (CSR: row_ptr[n+1], attr[nnz], x[nnz]) + targets.  attr_group: [num_attr] or NULL (all 0). */
fm_oracle* fm_oracle_create(const fm_oracle_config* cfg, uint32_t n_train, const int64_t* row_ptr, const uint32_t* attr, const float* x,
                            const float* y, uint32_t n_test, const int64_t* t_row_ptr, const uint32_t* t_attr, const float* t_x,
                            const float* t_y, const uint32_t* attr_group);
void fm_oracle_destroy(fm_oracle*);
void fm_oracle_srand(unsigned seed);
/* append {tag, a, b} (3 doubles) per two-argument sampler call, the format of shim_random.h; NULL closes */
int fm_oracle_set_log(fm_oracle*, const char* path);
/* w_init [num_attr], v_init [K][num_attr], or NULL for libFM's own order of draws: all of v (f outer, attribute inner,
   fm_model.h:96), then all of w ([L]:412).  Also runs the first prediction pass ([GS]:73-78). */
void fm_oracle_init(fm_oracle*, const double* w_init, const double* v_init);
/* iterations continue the chain; rmse_train[i] / rmse_test[i] are the values of the "#Iter=" line ([GS]:244) */
void fm_oracle_learn(fm_oracle*, uint32_t iters, double* rmse_train, double* rmse_test);

/* state (any pointer may be NULL): w [num_attr], v [K][num_attr], w_mu/w_lambda [G], v_mu/v_lambda [G][K], e [n_train],
   pred_sum [n_test] (clamped predictions summed over the iterations), scal = {w0, alpha} */
void fm_oracle_get_state(const fm_oracle*, double* w, double* v, double* w_mu, double* w_lambda, double* v_mu, double* v_lambda,
                         double* e, double* pred_sum, double* scal);
/* train design matrix in column form as libFM's create_data_t builds it (Data.h:472-528): col_ptr [num_attr+1], case_id / x [nnz] */
void fm_oracle_get_columns(const fm_oracle*, int64_t* col_ptr, uint32_t* case_id, float* x);

#ifdef __cplusplus
}
#endif
#endif
