/*
 * sbmf_oracle.c -- CPU restatement of the SBMF Gibbs sweep.  TEST INFRASTRUCTURE ONLY
 * (see sbmf_oracle.h for who may load it and how it is pinned).
 *
 * Citations: [T] = /root/reference/gibbs_sbpmf2.cpp, [R] = /root/reference/src/util/random.h.
 * The arithmetic below keeps [T]'s operation order in fp64 so that, fed by glibc rand() through
 * [R]'s samplers, it reproduces the reference's printed RMSE trajectory and sampler-argument stream.
 * Storage differs on purpose (flat CSR/CSC instead of one new[] per row); values do not.
 */
#include "sbmf_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct sbmf_oracle {
    uint32_t I, J, K;
    uint64_t N, Nt;
    uint32_t *tu, *ti;       /* train pairs, file order                 [T]:78-97  */
    double* target;          /*                                           [T]:89     */
    uint32_t *su, *si;       /* test pairs                               [T]:130-149 */
    double *ttarget, *sum;   /* test target, running prediction sum      [T]:145     */
    /* R / R_t of [T]:156-221 flattened: slot p of row i is r_id[rptr[i]+p], r_val[...] */
    int64_t *rptr, *cptr;
    uint64_t *r_id, *c_id;
    uint32_t *r_val, *c_val;
    double *U, *V;           /* U[i*K+k] ([T]:229-232), V[k*J+j] ([T]:234-237) */
    double *sigma_u, *mu_u, *sigma_v, *mu_v;
    double *mu_b_i, *sigma_b_i, *b_i, *mu_b_j, *sigma_b_j, *b_j;
    double mu_b_0, sigma_b_0, b_0, alpha;
    double* E;
    double last_rmse_sweep;
    uint32_t iter;
    int noise, stdev_mode;
    uint64_t seed;
    FILE* log;
    /* priors [T]:284-313 -- all 1/1/0/1 in the reference */
    double alpha_0, alpha_1, alpha_2, alpha_4, alpha_5, alpha_0_dash;
    double beta_0, beta_1, beta_2, beta_4, beta_5, beta_0_dash;
    double mu_0, mu_1, mu_2, mu_4, mu_5;
    double sigma_0, sigma_1, sigma_2, sigma_4, sigma_5;
    double clamp_lo, clamp_hi;
    /* variant 0 = [T] (top-level gibbs_sbpmf2.cpp); 1 = [S] (src/libfm/gibbs_sbpmf2.cpp as committed: no biases, no global
       mean, Normal-Gamma factor hyper-prior, including the slip at [S]:412); 2 = [S] with that slip corrected */
    int variant;
    double ng_a_0, ng_b_0, nu_0, ng_mu_0, ng_alpha_0, ng_beta_0;   /* [S]:260-269 */
};

#include "rand_samplers.h"

/* ------------------------------------------------------------------ Philox4x32-10 streams */

void sbmf_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void philox_site(uint64_t seed, uint32_t site, uint32_t row, uint32_t c1, uint32_t sweep, uint32_t out[4])
{
    uint32_t ctr[4] = {row, c1, site, sweep};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    sbmf_oracle_philox4x32_10(ctr, key, out);
}

float sbmf_oracle_philox_normal_f32(uint64_t seed, uint32_t site, uint32_t row, uint32_t c1, uint32_t sweep)
{
    uint32_t x[4];
    philox_site(seed, site, row, c1, sweep, x);
    float u1 = ((float)(x[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float u2 = ((float)(x[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float r = sqrtf(-2.0f * logf(u1));
    return r * cosf(6.28318530717958647692f * u2);
}

static double philox_normal_from(const uint32_t x[4])
{
    double u1 = ((double)x[0] + 0.5) * (1.0 / 4294967296.0);
    double u2 = ((double)x[1] + 0.5) * (1.0 / 4294967296.0);
    return sqrt(-2.0 * log(u1)) * cos(6.28318530717958647692 * u2);
}

double sbmf_oracle_philox_normal_f64(uint64_t seed, uint32_t site, uint32_t row, uint32_t c1, uint32_t sweep)
{
    uint32_t x[4];
    philox_site(seed, site, row, c1, sweep, x);
    return philox_normal_from(x);
}

static double philox_gamma(uint64_t seed, uint32_t site, uint32_t row, uint32_t sweep, double shape)
{
    /* Marsaglia-Tsang as [R]:129-143; attempt a uses counter (row, a, site, sweep): normal from words 0/1,
       uniform from word 2.  shape < 1 ([R]:120-125): Gamma(shape + 1) * u^(1/shape), u from word 3 of counter
       (row, 0xffffffff, site, sweep) -- the same streams as the device's draw_gamma_f64 (csrc/common.cuh). */
    double boost = 1.0;
    if (shape < 1.0) {
        uint32_t xb[4];
        philox_site(seed, site, row, 0xffffffffu, sweep, xb);
        boost = pow(((double)xb[3] + 0.5) * (1.0 / 4294967296.0), 1.0 / shape);
        shape += 1.0;
    }
    double d = shape - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    for (uint32_t a = 0;; ++a) {
        uint32_t x[4];
        philox_site(seed, site, row, a, sweep, x);
        double z = philox_normal_from(x);
        double v = 1.0 + c * z;
        if (v <= 0.0) continue;
        v = v * v * v;
        double u = ((double)x[2] + 0.5) * (1.0 / 4294967296.0);
        if (u < 1.0 - 0.0331 * (z * z) * (z * z)) return d * v * boost;
        if (log(u) < 0.5 * z * z + d * (1.0 - v + log(v))) return d * v * boost;
    }
}

/* ------------------------------------------------------------------ sampler front-ends */

static int site_is_f32(uint32_t site)
{
    return site == SBMF_SITE_U || site == SBMF_SITE_V || site == SBMF_SITE_BI || site == SBMF_SITE_BJ ||
           site == SBMF_SITE_INIT_U || site == SBMF_SITE_INIT_V;
}

/* ran_gaussian(mean, stdev) of [R]:166-172 incl. the "stdev==0 or NaN => mean" guard. */
static double draw_gauss(sbmf_oracle* h, uint32_t site, uint32_t row, uint32_t c1, double mean, double stdev)
{
    if (h->log) {
        double rec[3] = {0.0, mean, stdev};
        fwrite(rec, sizeof(double), 3, h->log);
    }
    if (h->noise == SBMF_ORACLE_NOISE_ZERO) return mean;
    if ((stdev == 0.0) || isnan(stdev)) return mean;
    if (h->noise == SBMF_ORACLE_NOISE_RAND) return mean + stdev * ran_gaussian_leva();
    if (site_is_f32(site)) return mean + stdev * (double)sbmf_oracle_philox_normal_f32(h->seed, site, row, c1, h->iter);
    return mean + stdev * sbmf_oracle_philox_normal_f64(h->seed, site, row, c1, h->iter);
}

/* ran_gamma(alpha, beta) = ran_gamma(alpha)/beta of [R]:146-148. */
static double draw_gamma(sbmf_oracle* h, uint32_t site, uint32_t row, double a, double b)
{
    if (h->log) {
        double rec[3] = {1.0, a, b};
        fwrite(rec, sizeof(double), 3, h->log);
    }
    if (h->noise == SBMF_ORACLE_NOISE_ZERO) return a / b;
    if (h->noise == SBMF_ORACLE_NOISE_RAND) return ran_gamma_mt_rand(a) / b;
    return philox_gamma(h->seed, site, row, h->iter, a) / b;
}

/* The reference hands the posterior variance 1/lambda to ran_gaussian as its stdev. */
static double post_stdev(const sbmf_oracle* h, double var) { return h->stdev_mode == SBMF_ORACLE_STDEV_SQRT ? sqrt(var) : var; }

/* ------------------------------------------------------------------ construction */

static void* xcalloc(size_t n, size_t sz)
{
    void* p = calloc(n ? n : 1, sz);
    if (!p) {
        fprintf(stderr, "sbmf_oracle: out of memory\n");
        abort();
    }
    return p;
}

sbmf_oracle* sbmf_oracle_create(uint64_t n, const uint32_t* user, const uint32_t* item, const double* rating,
                                uint64_t nt, const uint32_t* tuser, const uint32_t* titem, const double* trating,
                                uint32_t num_users, uint32_t num_items, uint32_t K,
                                int noise_mode, int stdev_mode, uint64_t seed)
{
    sbmf_oracle* h = (sbmf_oracle*)xcalloc(1, sizeof(*h));
    h->I = num_users; h->J = num_items; h->K = K; h->N = n; h->Nt = nt;
    h->noise = noise_mode; h->stdev_mode = stdev_mode; h->seed = seed;
    h->tu = (uint32_t*)xcalloc(n, 4); h->ti = (uint32_t*)xcalloc(n, 4); h->target = (double*)xcalloc(n, 8);
    h->su = (uint32_t*)xcalloc(nt, 4); h->si = (uint32_t*)xcalloc(nt, 4);
    h->ttarget = (double*)xcalloc(nt, 8); h->sum = (double*)xcalloc(nt, 8);
    memcpy(h->tu, user, n * 4); memcpy(h->ti, item, n * 4); memcpy(h->target, rating, n * 8);
    if (nt) { memcpy(h->su, tuser, nt * 4); memcpy(h->si, titem, nt * 4); memcpy(h->ttarget, trating, nt * 8); }

    /* R / R_t: append in file order to the user's row and the item's row ([T]:209-214) */
    h->rptr = (int64_t*)xcalloc((size_t)h->I + 1, 8); h->cptr = (int64_t*)xcalloc((size_t)h->J + 1, 8);
    h->r_id = (uint64_t*)xcalloc(n, 8); h->c_id = (uint64_t*)xcalloc(n, 8);
    h->r_val = (uint32_t*)xcalloc(n, 4); h->c_val = (uint32_t*)xcalloc(n, 4);
    for (uint64_t t = 0; t < n; ++t) { h->rptr[user[t] + 1]++; h->cptr[item[t] + 1]++; }
    for (uint32_t i = 0; i < h->I; ++i) h->rptr[i + 1] += h->rptr[i];
    for (uint32_t j = 0; j < h->J; ++j) h->cptr[j + 1] += h->cptr[j];
    int64_t* ui = (int64_t*)xcalloc(h->I, 8); int64_t* ii = (int64_t*)xcalloc(h->J, 8);
    for (uint64_t t = 0; t < n; ++t) {
        int64_t p = h->rptr[user[t]] + ui[user[t]]++;
        h->r_id[p] = t; h->r_val[p] = item[t];
        int64_t q = h->cptr[item[t]] + ii[item[t]]++;
        h->c_id[q] = t; h->c_val[q] = user[t];
    }
    free(ui); free(ii);

    h->U = (double*)xcalloc((size_t)h->I * K, 8); h->V = (double*)xcalloc((size_t)K * h->J, 8);
    h->sigma_u = (double*)xcalloc(K, 8); h->mu_u = (double*)xcalloc(K, 8);
    h->sigma_v = (double*)xcalloc(K, 8); h->mu_v = (double*)xcalloc(K, 8);
    h->mu_b_i = (double*)xcalloc(h->I, 8); h->sigma_b_i = (double*)xcalloc(h->I, 8); h->b_i = (double*)xcalloc(h->I, 8);
    h->mu_b_j = (double*)xcalloc(h->J, 8); h->sigma_b_j = (double*)xcalloc(h->J, 8); h->b_j = (double*)xcalloc(h->J, 8);
    h->E = (double*)xcalloc(n, 8);
    /* [T]:284-318 */
    h->alpha_0 = h->alpha_1 = h->alpha_2 = h->alpha_4 = h->alpha_5 = h->alpha_0_dash = 1.0;
    h->beta_0 = h->beta_1 = h->beta_2 = h->beta_4 = h->beta_5 = h->beta_0_dash = 1.0;
    h->mu_0 = h->mu_1 = h->mu_2 = h->mu_4 = h->mu_5 = 0.0;
    h->sigma_0 = h->sigma_1 = h->sigma_2 = h->sigma_4 = h->sigma_5 = 1.0;
    h->mu_b_0 = 0.0; h->sigma_b_0 = 0.0; h->b_0 = 0.0; h->alpha = 0.0;
    h->clamp_lo = 0.5; h->clamp_hi = 5.0;                                    /* [T]:627-628 */
    h->variant = 0;
    h->ng_a_0 = 1; h->ng_b_0 = 1; h->nu_0 = 1; h->ng_mu_0 = 0.0; h->ng_alpha_0 = 1; h->ng_beta_0 = 1;   /* [S]:260-269 */
    return h;
}

void sbmf_oracle_destroy(sbmf_oracle* h)
{
    if (!h) return;
    if (h->log) fclose(h->log);
    free(h->tu); free(h->ti); free(h->target); free(h->su); free(h->si); free(h->ttarget); free(h->sum);
    free(h->rptr); free(h->cptr); free(h->r_id); free(h->c_id); free(h->r_val); free(h->c_val);
    free(h->U); free(h->V); free(h->sigma_u); free(h->mu_u); free(h->sigma_v); free(h->mu_v);
    free(h->mu_b_i); free(h->sigma_b_i); free(h->b_i); free(h->mu_b_j); free(h->sigma_b_j); free(h->b_j);
    free(h->E);
    free(h);
}

void sbmf_oracle_srand(unsigned seed) { srand(seed); }
void sbmf_oracle_set_variant(sbmf_oracle* h, int variant) { h->variant = variant; }

int sbmf_oracle_set_log(sbmf_oracle* h, const char* path)
{
    if (h->log) { fclose(h->log); h->log = NULL; }
    if (!path) return 0;
    h->log = fopen(path, "wb");
    return h->log ? 0 : -1;
}

static void log_init(sbmf_oracle* h)
{
    if (h->log) {
        /* [T]:242, 248 call ran_gaussian(0.0,1.0) and scale by 0.1; [S]:240, 248 call ran_gaussian(0.0,0.1) */
        double rec[3] = {0.0, 0.0, h->variant != 0 ? 0.1 : 1.0};
        fwrite(rec, sizeof(double), 3, h->log);
    }
}

void sbmf_oracle_init_factors(sbmf_oracle* h, const double* U0, const double* V0, double init_stdev)
{
    /* [T]:239-250: U first (i outer, k inner), then V (k outer, j inner); 0.1*ran_gaussian(0,1).
       Initial draws are always live (a zero-noise run still needs non-zero factors): RAND and
       ZERO modes use rand()+Leva, PHILOX uses the INIT_U / INIT_V streams at sweep 0. */
    const uint32_t K = h->K;
    if (U0) memcpy(h->U, U0, (size_t)h->I * K * 8);
    else
        for (uint32_t i = 0; i < h->I; ++i)
            for (uint32_t k = 0; k < K; ++k) {
                log_init(h);
                double z = h->noise == SBMF_ORACLE_NOISE_PHILOX
                               ? (double)sbmf_oracle_philox_normal_f32(h->seed, SBMF_SITE_INIT_U, i, k, 0)
                               : (0.0 + 1.0 * ran_gaussian_leva());
                h->U[(size_t)i * K + k] = init_stdev * z;
            }
    if (V0) memcpy(h->V, V0, (size_t)K * h->J * 8);
    else
        for (uint32_t k = 0; k < K; ++k)
            for (uint32_t j = 0; j < h->J; ++j) {
                log_init(h);
                double z = h->noise == SBMF_ORACLE_NOISE_PHILOX
                               ? (double)sbmf_oracle_philox_normal_f32(h->seed, SBMF_SITE_INIT_V, j, k, 0)
                               : (0.0 + 1.0 * ran_gaussian_leva());
                h->V[(size_t)k * h->J + j] = init_stdev * z;
            }
}

/* ------------------------------------------------------------------ one sweep, [T]:335-637 */

static void one_sweep(sbmf_oracle* h, double* rmse_out, double* rmse_sweep_out)
{
    const uint32_t I = h->I, J = h->J, D = h->K;
    const uint64_t N = h->N;
    double* U = h->U; double* V = h->V; double* E = h->E;
    const double num_rows = (double)N;   /* [T] uses uint num_rows in double expressions */

    /* k1: residual rebuild + stats, [T]:342-359 */
    double E_sum = 0.0, E_sq = 0.0;
    for (uint64_t n = 0; n < N; ++n) {
        uint32_t user = h->tu[n], item = h->ti[n];
        double temp = 0.0;
        for (uint32_t k = 0; k < D; ++k) temp += U[(size_t)user * D + k] * V[(size_t)k * J + item];
        E[n] = h->target[n] - (h->b_0 + h->b_i[user] + h->b_j[item] + temp);
        E_sum += E[n];
        E_sq += (E[n] * E[n]);
    }

    /* alpha, [T]:366-372 */
    {
        double a = h->alpha_0_dash + num_rows;
        double b = h->beta_0_dash + E_sq;
        h->alpha = draw_gamma(h, SBMF_SITE_ALPHA, 0, a, b);
    }
    /* sigma_b_0, [T]:378-383 */
    {
        double a = h->alpha_0 + 1;
        double b = h->beta_0 + (0.5 * (h->b_0 - h->mu_b_0) * (h->b_0 - h->mu_b_0));
        h->sigma_b_0 = draw_gamma(h, SBMF_SITE_SIGMA_B0, 0, a, b);
    }
    /* mu_b_0, [T]:388-393 */
    {
        double s = 1.0 / (h->sigma_0 + h->sigma_b_0);
        double m = s * ((h->sigma_0 * h->mu_0) + h->b_0 * h->sigma_b_0);
        h->mu_b_0 = draw_gauss(h, SBMF_SITE_MU_B0, 0, 0, m, post_stdev(h, s));
    }
    /* b_0 and the global shift, [T]:398-410 */
    {
        double s = 1 / (h->sigma_b_0 + h->alpha * num_rows);
        double m = s * (h->sigma_b_0 * h->mu_b_0 + h->alpha * (E_sum + num_rows * h->b_0));
        double old = h->b_0;
        h->b_0 = draw_gauss(h, SBMF_SITE_B0, 0, 0, m, post_stdev(h, s));
        for (uint64_t n = 0; n < N; ++n) E[n] += (old - h->b_0);
    }

    /* k3: per-dimension hypers, [T]:415-467 (uses the previous sweep's mu_u / mu_v in the sums) */
    for (uint32_t k = 0; k < D; ++k) {
        double temp = 0.0, temp2 = 0.0;
        for (uint32_t i = 0; i < I; ++i) {
            double u = U[(size_t)i * D + k];
            temp += (u - h->mu_u[k]) * (u - h->mu_u[k]);
            temp2 += u;
        }
        h->sigma_u[k] = draw_gamma(h, SBMF_SITE_SIGMA_U, k, h->alpha_2 + (double)I, h->beta_2 + (0.5) * temp);
        {
            double s = 1 / (h->sigma_2 + h->sigma_u[k] * (double)I);
            double m = s * (h->sigma_2 * h->mu_2 + h->sigma_u[k] * temp2);
            h->mu_u[k] = draw_gauss(h, SBMF_SITE_MU_U, k, 0, m, post_stdev(h, s));
        }
        temp = 0.0; temp2 = 0.0;
        for (uint32_t j = 0; j < J; ++j) {
            double v = V[(size_t)k * J + j];
            temp += (v - h->mu_v[k]) * (v - h->mu_v[k]);
            temp2 += v;
        }
        h->sigma_v[k] = draw_gamma(h, SBMF_SITE_SIGMA_V, k, h->alpha_1 + (double)J, h->beta_1 + (0.5) * temp);
        {
            double s = 1 / (h->sigma_1 + h->sigma_v[k] * (double)J);
            double m = s * (h->sigma_1 * h->mu_1 + h->sigma_v[k] * temp2);
            h->mu_v[k] = draw_gauss(h, SBMF_SITE_MU_V, k, 0, m, post_stdev(h, s));
        }
    }

    /* k4: user-bias hypers, [T]:469-489 */
    for (uint32_t i = 0; i < I; ++i) {
        double a = h->alpha_4 + 1;
        double b = h->beta_4 + (0.5 * (h->b_i[i] - h->mu_b_i[i]) * (h->b_i[i] - h->mu_b_i[i]));
        h->sigma_b_i[i] = draw_gamma(h, SBMF_SITE_SIGMA_BI, i, a, b);
        double s = 1.0 / (h->sigma_4 + h->sigma_b_i[i]);
        double m = s * ((h->sigma_4 * h->mu_4) + h->b_i[i] * h->sigma_b_i[i]);
        h->mu_b_i[i] = draw_gauss(h, SBMF_SITE_MU_BI, i, 0, m, post_stdev(h, s));
    }
    /* k5: item-bias hypers, [T]:491-511 */
    for (uint32_t j = 0; j < J; ++j) {
        double a = h->alpha_5 + 1;
        double b = h->beta_5 + (0.5 * (h->b_j[j] - h->mu_b_j[j]) * (h->b_j[j] - h->mu_b_j[j]));
        h->sigma_b_j[j] = draw_gamma(h, SBMF_SITE_SIGMA_BJ, j, a, b);
        double s = 1.0 / (h->sigma_5 + h->sigma_b_j[j]);
        double m = s * ((h->sigma_5 * h->mu_5) + h->b_j[j] * h->sigma_b_j[j]);
        h->mu_b_j[j] = draw_gauss(h, SBMF_SITE_MU_BJ, j, 0, m, post_stdev(h, s));
    }

    /* k6 + k7: user phase, [T]:514-558 */
    for (uint32_t i = 0; i < I; ++i) {
        const int64_t beg = h->rptr[i];
        const uint32_t c = (uint32_t)(h->rptr[i + 1] - beg);
        const uint64_t* id = h->r_id + beg;
        const uint32_t* val = h->r_val + beg;
        {
            double s = 1 / (h->sigma_b_i[i] + (h->alpha * c));
            double temp = 0.0;
            for (uint32_t p = 0; p < c; ++p) temp += (E[id[p]] + h->b_i[i]);
            double m = s * ((h->sigma_b_i[i] * h->mu_b_i[i]) + h->alpha * temp);
            double old = h->b_i[i];
            h->b_i[i] = draw_gauss(h, SBMF_SITE_BI, i, 0, m, post_stdev(h, s));
            for (uint32_t p = 0; p < c; ++p) E[id[p]] += (old - h->b_i[i]);
        }
        for (uint32_t k = 0; k < D; ++k) {
            const double* Vk = V + (size_t)k * J;
            double* u = &U[(size_t)i * D + k];
            double temp = 0.0, temp2 = 0.0;
            for (uint32_t p = 0; p < c; ++p) {
                temp += (Vk[val[p]] * Vk[val[p]]);
                temp2 += (Vk[val[p]] * (E[id[p]] + Vk[val[p]] * (*u)));
            }
            double s = 1 / (h->sigma_u[k] + (h->alpha * temp));
            double m = s * (h->alpha * temp2 + h->sigma_u[k] * h->mu_u[k]);
            double old = *u;
            *u = draw_gauss(h, SBMF_SITE_U, i, k, m, post_stdev(h, s));
            for (uint32_t p = 0; p < c; ++p) E[id[p]] += Vk[val[p]] * (old - *u);
        }
    }

    /* k8 + k9: item phase, [T]:563-606 */
    for (uint32_t j = 0; j < J; ++j) {
        const int64_t beg = h->cptr[j];
        const uint32_t c = (uint32_t)(h->cptr[j + 1] - beg);
        const uint64_t* id = h->c_id + beg;
        const uint32_t* val = h->c_val + beg;
        {
            double s = 1 / (h->sigma_b_j[j] + (h->alpha * c));
            double temp = 0.0;
            for (uint32_t p = 0; p < c; ++p) temp += (E[id[p]] + h->b_j[j]);
            double m = s * ((h->sigma_b_j[j] * h->mu_b_j[j]) + h->alpha * temp);
            double old = h->b_j[j];
            h->b_j[j] = draw_gauss(h, SBMF_SITE_BJ, j, 0, m, post_stdev(h, s));
            for (uint32_t p = 0; p < c; ++p) E[id[p]] += (old - h->b_j[j]);
        }
        for (uint32_t k = 0; k < D; ++k) {
            double* v = &V[(size_t)k * J + j];
            double temp = 0.0, temp2 = 0.0;
            for (uint32_t p = 0; p < c; ++p) {
                double u = U[(size_t)val[p] * D + k];
                temp += (u * u);
                temp2 += (u * (E[id[p]] + (*v) * u));
            }
            double s = 1 / (h->sigma_v[k] + (h->alpha * temp));
            double m = s * (h->alpha * temp2 + h->sigma_v[k] * h->mu_v[k]);
            double old = *v;
            *v = draw_gauss(h, SBMF_SITE_V, j, k, m, post_stdev(h, s));
            for (uint32_t p = 0; p < c; ++p) E[id[p]] += U[(size_t)val[p] * D + k] * (old - *v);
        }
    }

    /* k10: test prediction + running-mean RMSE, [T]:610-636 */
    {
        double diff_sqr_sum = 0.0, diff_sweep = 0.0;
        for (uint64_t t = 0; t < h->Nt; ++t) {
            uint32_t user = h->su[t], item = h->si[t];
            double temp = h->b_0 + h->b_i[user] + h->b_j[item];
            for (uint32_t k = 0; k < D; ++k) temp += U[(size_t)user * D + k] * V[(size_t)k * J + item];
            temp = temp < h->clamp_hi ? temp : h->clamp_hi;   /* std::min(5.0, temp) */
            temp = temp > h->clamp_lo ? temp : h->clamp_lo;   /* std::max(0.5, temp) */
            h->sum[t] += temp;
            double d = h->ttarget[t] - ((double)h->sum[t] / (h->iter + 1));
            diff_sqr_sum += d * d;
            diff_sweep += (h->ttarget[t] - temp) * (h->ttarget[t] - temp);
        }
        double rmse = sqrt(diff_sqr_sum / (double)h->Nt);
        h->last_rmse_sweep = sqrt(diff_sweep / (double)h->Nt);
        if (rmse_out) *rmse_out = rmse;
        if (rmse_sweep_out) *rmse_sweep_out = h->last_rmse_sweep;
    }
    h->iter++;
}

/* ------------------------------------------------------------------ one sweep of [S] = src/libfm/gibbs_sbpmf2.cpp:308-563
 * (the active code only: everything about b_0 / b_i / b_j sits inside comments there) */
static void one_sweep_S(sbmf_oracle* h, double* rmse_out, double* rmse_sweep_out)
{
    const uint32_t I = h->I, J = h->J, D = h->K;
    const uint64_t N = h->N;
    double* U = h->U; double* V = h->V; double* E = h->E;
    const double num_rows = (double)N;

    /* [S]:317-334 */
    double E_sum = 0.0, E_sq = 0.0;
    for (uint64_t n = 0; n < N; ++n) {
        uint32_t user = h->tu[n], item = h->ti[n];
        double temp = 0.0;
        for (uint32_t k = 0; k < D; ++k) temp += U[(size_t)user * D + k] * V[(size_t)k * J + item];
        E[n] = h->target[n] - (0.0 + 0.0 + 0.0 + temp);
        E_sum += E[n];
        E_sq += (E[n] * E[n]);
    }
    /* tau, [S]:339-342 */
    h->alpha = draw_gamma(h, SBMF_SITE_ALPHA, 0, h->ng_a_0 + 0.5 * num_rows, h->ng_b_0 + 0.5 * E_sq);

    /* Normal-Gamma factor hypers, [S]:375-414 */
    for (uint32_t k = 0; k < D; ++k) {
        double temp = 0.0, temp2 = 0.0;
        for (uint32_t i = 0; i < I; ++i) {
            double u = U[(size_t)i * D + k];
            temp += (u - h->mu_u[k]) * (u - h->mu_u[k]);
            temp2 += u;
        }
        double a = h->ng_alpha_0 + 0.5 * (double)(I + 1);
        double b = h->ng_beta_0 + h->nu_0 * (h->mu_u[k] - h->ng_mu_0) * (h->mu_u[k] - h->ng_mu_0) + (0.5) * temp;
        h->sigma_u[k] = draw_gamma(h, SBMF_SITE_SIGMA_U, k, a, b);
        double sigma_u_k_star = (double)1.0 / (h->nu_0 * h->sigma_u[k] + h->sigma_u[k] * (double)I);
        double mu_u_k_star = sigma_u_k_star * (h->nu_0 * h->ng_mu_0 * h->sigma_u[k] + h->sigma_u[k] * temp2);
        h->mu_u[k] = draw_gauss(h, SBMF_SITE_MU_U, k, 0, mu_u_k_star, post_stdev(h, sigma_u_k_star));

        temp = 0.0; temp2 = 0.0;
        for (uint32_t j = 0; j < J; ++j) {
            double v = V[(size_t)k * J + j];
            temp += (v - h->mu_v[k]) * (v - h->mu_v[k]);
            temp2 += v;
        }
        a = h->ng_alpha_0 + 0.5 * (double)(J + 1);
        b = h->ng_beta_0 + h->nu_0 * (h->mu_v[k] - h->ng_mu_0) * (h->mu_v[k] - h->ng_mu_0) + (0.5) * temp;
        h->sigma_v[k] = draw_gamma(h, SBMF_SITE_SIGMA_V, k, a, b);
        double sigma_v_k_star = (double)1.0 / (h->nu_0 * h->sigma_v[k] + h->sigma_v[k] * (double)J);
        /* [S]:412 multiplies by sigma_u_k_star (a copy-paste slip); variant 2 uses sigma_v_k_star */
        double lead = (h->variant == 1) ? sigma_u_k_star : sigma_v_k_star;
        double mu_v_k_star = lead * (h->nu_0 * h->ng_mu_0 * h->sigma_v[k] + h->sigma_v[k] * temp2);
        h->mu_v[k] = draw_gauss(h, SBMF_SITE_MU_V, k, 0, mu_v_k_star, post_stdev(h, sigma_v_k_star));
    }

    /* users, [S]:452-489 (bias step commented out there) */
    for (uint32_t i = 0; i < I; ++i) {
        const int64_t beg = h->rptr[i];
        const uint32_t c = (uint32_t)(h->rptr[i + 1] - beg);
        const uint64_t* id = h->r_id + beg;
        const uint32_t* val = h->r_val + beg;
        for (uint32_t k = 0; k < D; ++k) {
            const double* Vk = V + (size_t)k * J;
            double* u = &U[(size_t)i * D + k];
            double temp = 0.0, temp2 = 0.0;
            for (uint32_t p = 0; p < c; ++p) {
                temp += (Vk[val[p]] * Vk[val[p]]);
                temp2 += (Vk[val[p]] * (E[id[p]] + Vk[val[p]] * (*u)));
            }
            double s = (double)1.0 / (h->sigma_u[k] + (h->alpha * temp));
            double m = s * (h->alpha * temp2 + h->sigma_u[k] * h->mu_u[k]);
            double old = *u;
            *u = draw_gauss(h, SBMF_SITE_U, i, k, m, post_stdev(h, s));
            for (uint32_t p = 0; p < c; ++p) E[id[p]] += Vk[val[p]] * (old - *u);
        }
    }
    /* items, [S]:494-534 */
    for (uint32_t j = 0; j < J; ++j) {
        const int64_t beg = h->cptr[j];
        const uint32_t c = (uint32_t)(h->cptr[j + 1] - beg);
        const uint64_t* id = h->c_id + beg;
        const uint32_t* val = h->c_val + beg;
        for (uint32_t k = 0; k < D; ++k) {
            double* v = &V[(size_t)k * J + j];
            double temp = 0.0, temp2 = 0.0;
            for (uint32_t p = 0; p < c; ++p) {
                double u = U[(size_t)val[p] * D + k];
                temp += (u * u);
                temp2 += (u * (E[id[p]] + (*v) * u));
            }
            double s = (double)1.0 / (h->sigma_v[k] + (h->alpha * temp));
            double m = s * (h->alpha * temp2 + h->sigma_v[k] * h->mu_v[k]);
            double old = *v;
            *v = draw_gauss(h, SBMF_SITE_V, j, k, m, post_stdev(h, s));
            for (uint32_t p = 0; p < c; ++p) E[id[p]] += U[(size_t)val[p] * D + k] * (old - *v);
        }
    }
    /* test RMSE, [S]:538-562 */
    {
        double diff_sqr_sum = 0.0, diff_sweep = 0.0;
        for (uint64_t t = 0; t < h->Nt; ++t) {
            uint32_t user = h->su[t], item = h->si[t];
            double temp = 0.0;
            for (uint32_t k = 0; k < D; ++k) temp += U[(size_t)user * D + k] * V[(size_t)k * J + item];
            temp = temp < h->clamp_hi ? temp : h->clamp_hi;
            temp = temp > h->clamp_lo ? temp : h->clamp_lo;
            h->sum[t] += temp;
            double d = h->ttarget[t] - ((double)h->sum[t] / (h->iter + 1));
            diff_sqr_sum += d * d;
            diff_sweep += (h->ttarget[t] - temp) * (h->ttarget[t] - temp);
        }
        double rmse = sqrt(diff_sqr_sum / (double)h->Nt);
        h->last_rmse_sweep = sqrt(diff_sweep / (double)h->Nt);
        if (rmse_out) *rmse_out = rmse;
        if (rmse_sweep_out) *rmse_sweep_out = h->last_rmse_sweep;
    }
    h->iter++;
    (void)E_sum;
}

void sbmf_oracle_sweep(sbmf_oracle* h, uint32_t n, double* rmse_out, double* rmse_sweep_out)
{
    if (h->variant != 0) {
        for (uint32_t s = 0; s < n; ++s) one_sweep_S(h, rmse_out ? rmse_out + s : NULL, rmse_sweep_out ? rmse_sweep_out + s : NULL);
        if (h->log) fflush(h->log);
        return;
    }
    for (uint32_t s = 0; s < n; ++s) one_sweep(h, rmse_out ? rmse_out + s : NULL, rmse_sweep_out ? rmse_sweep_out + s : NULL);
    if (h->log) fflush(h->log);
}

/* ------------------------------------------------------------------ getters */

void sbmf_oracle_get_U(const sbmf_oracle* h, double* U) { memcpy(U, h->U, (size_t)h->I * h->K * 8); }
void sbmf_oracle_get_V(const sbmf_oracle* h, double* V) { memcpy(V, h->V, (size_t)h->K * h->J * 8); }
void sbmf_oracle_get_bias(const sbmf_oracle* h, double* b_i, double* b_j)
{
    memcpy(b_i, h->b_i, (size_t)h->I * 8);
    memcpy(b_j, h->b_j, (size_t)h->J * 8);
}
void sbmf_oracle_get_bias_hypers(const sbmf_oracle* h, double* mu_b_i, double* sigma_b_i, double* mu_b_j, double* sigma_b_j)
{
    memcpy(mu_b_i, h->mu_b_i, (size_t)h->I * 8); memcpy(sigma_b_i, h->sigma_b_i, (size_t)h->I * 8);
    memcpy(mu_b_j, h->mu_b_j, (size_t)h->J * 8); memcpy(sigma_b_j, h->sigma_b_j, (size_t)h->J * 8);
}
void sbmf_oracle_get_dim_hypers(const sbmf_oracle* h, double* sigma_u, double* mu_u, double* sigma_v, double* mu_v)
{
    memcpy(sigma_u, h->sigma_u, (size_t)h->K * 8); memcpy(mu_u, h->mu_u, (size_t)h->K * 8);
    memcpy(sigma_v, h->sigma_v, (size_t)h->K * 8); memcpy(mu_v, h->mu_v, (size_t)h->K * 8);
}
void sbmf_oracle_get_scalars(const sbmf_oracle* h, double* s)
{
    s[0] = h->b_0; s[1] = h->alpha; s[2] = h->mu_b_0; s[3] = h->sigma_b_0;
}
void sbmf_oracle_get_E(const sbmf_oracle* h, double* E) { memcpy(E, h->E, (size_t)h->N * 8); }
void sbmf_oracle_get_pred_mean(const sbmf_oracle* h, double* pred)
{
    for (uint64_t t = 0; t < h->Nt; ++t) pred[t] = h->iter ? h->sum[t] / h->iter : 0.0;
}
void sbmf_oracle_get_layout(const sbmf_oracle* h, int64_t* row_ptr, uint32_t* col, uint64_t* csr_id,
                            int64_t* col_ptr, uint32_t* row, uint64_t* csc_id)
{
    memcpy(row_ptr, h->rptr, ((size_t)h->I + 1) * 8); memcpy(col, h->r_val, (size_t)h->N * 4);
    memcpy(csr_id, h->r_id, (size_t)h->N * 8);
    memcpy(col_ptr, h->cptr, ((size_t)h->J + 1) * 8); memcpy(row, h->c_val, (size_t)h->N * 4);
    memcpy(csc_id, h->c_id, (size_t)h->N * 8);
}

/* ------------------------------------------------------------------ triple reader, [T]:35-73 */

int64_t sbmf_oracle_read_triples(const char* path, uint32_t* user, uint32_t* item, double* rating,
                                 uint32_t* user_max, uint32_t* item_max)
{
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    char* line = NULL;
    size_t cap = 0;
    int64_t n = 0;
    uint32_t umax = 0, imax = 0;
    while (getline(&line, &cap, f) >= 0) {
        unsigned u, m;
        double r;
        char ws1, ws2;
        if (sscanf(line, "%u%c%u%c%lf", &u, &ws1, &m, &ws2, &r) >= 5) {
            if (user) { user[n] = u; item[n] = m; rating[n] = r; }
            if (umax < u) umax = u;
            if (imax < m) imax = m;
            n++;
        }
    }
    free(line);
    fclose(f);
    if (user_max) *user_max = umax;
    if (item_max) *item_max = imax;
    return n;
}

/* ------------------------------------------------------------------ stand-alone runner
 * sbmf_oracle_cli <train> <test> <K> <T> [noise=0|1|2] [stdev_mode=0|1] [calllog]
 * prints the same three header lines and per-sweep "rmse is" line as [T]:225-227, 635. */
#ifdef SBMF_ORACLE_MAIN
int main(int argc, char** argv)
{
    if (argc < 5) {
        fprintf(stderr, "usage: %s train test K T [noise] [stdev_mode] [calllog] [live_init_only]\n", argv[0]);
        return 2;
    }
    uint32_t K = (uint32_t)atoi(argv[3]), T = (uint32_t)atoi(argv[4]);
    int noise = argc > 5 ? atoi(argv[5]) : 0, sm = argc > 6 ? atoi(argv[6]) : 0;
    uint32_t um = 0, im = 0, um2 = 0, im2 = 0;
    int64_t n = sbmf_oracle_read_triples(argv[1], NULL, NULL, NULL, &um, &im);
    int64_t nt = sbmf_oracle_read_triples(argv[2], NULL, NULL, NULL, &um2, &im2);
    if (n < 0 || nt < 0) { fprintf(stderr, "cannot read input\n"); return 1; }
    if (um2 > um) um = um2;
    if (im2 > im) im = im2;
    uint32_t* u = (uint32_t*)xcalloc(n, 4); uint32_t* it = (uint32_t*)xcalloc(n, 4); double* r = (double*)xcalloc(n, 8);
    uint32_t* tu = (uint32_t*)xcalloc(nt, 4); uint32_t* ti = (uint32_t*)xcalloc(nt, 4); double* tr = (double*)xcalloc(nt, 8);
    sbmf_oracle_read_triples(argv[1], u, it, r, NULL, NULL);
    sbmf_oracle_read_triples(argv[2], tu, ti, tr, NULL, NULL);
    sbmf_oracle* h = sbmf_oracle_create((uint64_t)n, u, it, r, (uint64_t)nt, tu, ti, tr, um + 1, im + 1, K, noise, sm, 1);
    if (argc > 7 && argv[7][0]) sbmf_oracle_set_log(h, argv[7]);
    printf("number rows =%lld\nnumber of user =%u\nnumber of items =%u\n", (long long)n, um + 1, im + 1);
    sbmf_oracle_init_factors(h, NULL, NULL, 0.1);
    for (uint32_t s = 0; s < T; ++s) {
        double rmse;
        sbmf_oracle_sweep(h, 1, &rmse, NULL);
        printf("rmse is %g\n", rmse);   /* std::cout default precision = 6 significant digits = %g */
    }
    sbmf_oracle_destroy(h);
    free(u); free(it); free(r); free(tu); free(ti); free(tr);
    return 0;
}
#endif
