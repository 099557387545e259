/*
 * sbmf_fm_cuda.h -- C ABI of the B200-native general FM Gibbs sampler (libsbmf_cuda.so), SURVEY.md 8(f)-4: libFM's
 * `-method mcmc` / `-method als` for regression on an arbitrary sparse design matrix with attribute groups.
 *
 * The reference's operator interface for this path is the learner class of libFM: fm_learn::init / learn / predict
 * (src/libfm/src/fm_learn.h:80, 150, 191) as implemented by fm_learn_mcmc ("[G]", src/fm_learn_mcmc.h) and
 * fm_learn_mcmc_simultaneous ("[GS]"), configured by libfm.cpp ("[L]").  One entry point per virtual; each comment cites
 * the reference lines it replaces.  Not covered: relation blocks (`--relation`, [G]:57-64) and the classification task.
 *
 * Conventions are those of sbmf_cuda.h (status codes, caller-owned host buffers, no CPU fallback, one handle per GPU).
 */
#ifndef SBMF_FM_CUDA_H_
#define SBMF_FM_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#include "sbmf_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sbmf_fm_handle sbmf_fm_handle;

typedef enum sbmf_fm_sample_mode {
    SBMF_FM_SAMPLE_LIVE = 0,         /* ran_gaussian / ran_gamma draws (counter-based Philox streams) */
    SBMF_FM_SAMPLE_ZERO_NOISE = 1    /* ran_gaussian(m, s) -> m, ran_gamma(a, b) -> a / b: the parity mode of SURVEY.md 8(c) */
} sbmf_fm_sample_mode;

typedef struct sbmf_fm_config {
    uint32_t struct_size;     /* = sizeof(sbmf_fm_config); set by sbmf_fm_config_default */
    uint32_t num_attr;        /* number of attributes p.  libFM: max(train.num_feature, test.num_feature) + 1, [L]:326 */
    uint32_t num_groups;      /* meta->num_attr_groups, Data.h:49-61; 1 without -meta */
    uint32_t K;               /* -dim 'k0,k1,K': number of factors, [L]:377-383 */
    int32_t k0, k1;           /* use the global bias / the one-way weights */
    int32_t do_sample;        /* [L]:418; 0 = conditional means (libFM's -method als) */
    int32_t do_multilevel;    /* [L]:419; 0 = fixed hyper-parameters (alpha = 1, lambda = -regular, mu = 0) */
    int32_t sample_mode;      /* sbmf_fm_sample_mode */
    int32_t device;           /* CUDA device ordinal */
    uint64_t seed;            /* Philox key ([L]:124 seeds rand() with time(NULL)) */
    double init_stdev;        /* -init_stdev, default 0.1 ([L]:127) */
    double reg0, regw, regv;  /* -regular 'r0,r1,r2' ([L]:485-505): prior precision of w0; start values of w_lambda / v_lambda */
} sbmf_fm_config;

/* Host views for sbmf_fm_get_state; any pointer may be NULL.  v is [K][num_attr] like fm_model::v (fm_model.h:41). */
typedef struct sbmf_fm_state {
    float* w;                 /* [num_attr] */
    float* v;                 /* [K][num_attr] */
    double* w_mu;             /* [num_groups] */
    double* w_lambda;         /* [num_groups] */
    double* v_mu;             /* [num_groups][K] */
    double* v_lambda;         /* [num_groups][K] */
    float* e;                 /* [n_train]: prediction - target, the e-cache of [G]:52-55 */
    double* pred_sum;         /* [n_test]: pred_sum_all of [GS]:154-158 */
    double w0, alpha;
    uint32_t iterations;      /* completed iterations */
} sbmf_fm_state;

int sbmf_fm_config_default(sbmf_fm_config* cfg);                       /* [L]:125-128 + [G]:1099-1106 defaults: K = 8, stdev 0.1 */
int sbmf_fm_create(const sbmf_fm_config* cfg, sbmf_fm_handle** out);
int sbmf_fm_destroy(sbmf_fm_handle* h);
const char* sbmf_fm_last_error(const sbmf_fm_handle* h);

/* meta->attr_group ([L]:333-335, Data.h:49-61): group id of every attribute, [num_attr], ids < num_groups.  Before set_train. */
int sbmf_fm_set_groups(sbmf_fm_handle* h, const uint32_t* attr_group);

/* Data::load + create_data_t (Data.h:106-283, 472-528): the design matrix in row form (row_ptr [n+1], attr / x [nnz]) and the
   targets; the transposed (column) form the sampler walks is built on the device, stable in case order.  An attribute may be
   listed at most once per case.  min_target / max_target come from the TRAIN targets ([L]:466-467). */
int sbmf_fm_set_train(sbmf_fm_handle* h, uint32_t n, const int64_t* row_ptr, const uint32_t* attr, const float* x, const float* y);
int sbmf_fm_set_test(sbmf_fm_handle* h, uint32_t n, const int64_t* row_ptr, const uint32_t* attr, const float* x, const float* y);

/* fm_model::init + [L]:412 + fm_learn_mcmc::init + the first prediction pass of [GS]:73-78.  w_init [num_attr] and
   v_init [K][num_attr] are uploaded when given, else drawn as init_stdev * N(0,1) (Philox). */
int sbmf_fm_init(sbmf_fm_handle* h, const float* w_init, const float* v_init);

/* fm_learn::learn: `iters` further iterations of [GS]:97-262 = draw_all ([G]:411-626) + the full re-prediction of train and
   test + the running test prediction.  Enqueue only; the next call that returns values synchronises. */
int sbmf_fm_learn(sbmf_fm_handle* h, uint32_t iters);

/* the "#Iter= i Train= Test=" values of [GS]:244 for iterations [first, first + count) */
int sbmf_fm_rmse_history(sbmf_fm_handle* h, uint32_t first, uint32_t count, double* rmse_train, double* rmse_test);

/* fm_learn_mcmc::predict ([G]:355-379) for the test set: pred_sum_all / iterations (do_sample) or the last prediction, clamped */
int sbmf_fm_predict(sbmf_fm_handle* h, float* pred);

int sbmf_fm_get_state(sbmf_fm_handle* h, sbmf_fm_state* out);

/* The transposed design matrix as built on the device (bit-exact test against Data.h:472-528): col_ptr [num_attr + 1],
   case_id / x [nnz]; and the conflict-free runs of the attribute sequence the sampler draws in parallel: run_begin
   [*n_runs + 1] (capacity num_attr + 1).  Attributes of one run share no case, so drawing them concurrently IS libFM's
   sequential scan ([G]:441-457, 552-565). */
int sbmf_fm_get_columns(sbmf_fm_handle* h, int64_t* col_ptr, uint32_t* case_id, float* x);
int sbmf_fm_get_runs(sbmf_fm_handle* h, uint32_t* n_runs, uint32_t* run_begin);

/* Host-only planner behind set_train (no GPU needed; covered by the CPU tests): next_attr[a] = the smallest attribute id
   greater than a that occurs together with a in some case (UINT32_MAX if none).  Writes the maximal conflict-free runs of
   0..num_attr-1 in scan order to run_begin [<= num_attr + 1] and returns their number. */
uint32_t sbmf_fm_plan_runs(uint32_t num_attr, const uint32_t* next_attr, uint32_t* run_begin);

#ifdef __cplusplus
}
#endif
#endif
