/*
 * fm_oracle.c -- CPU restatement of libFM's MCMC learner for regression, no relations (see fm_oracle.h for the file:line map).
 * TEST INFRASTRUCTURE ONLY.  Every expression keeps libFM's operand types and association (x and y are float, everything else
 * double, products written in libFM's order) so that, on the same rand() stream, every sampler argument is bit-identical.
 */
#include "fm_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rand_samplers.h"

typedef struct {
    uint32_t n;          /* cases */
    uint32_t ncol;       /* this file's own feature count: its transposed matrix has this many rows (Data.h:222, 483) */
    int64_t* col_ptr;    /* [num_attr + 1] */
    uint32_t* case_id;   /* ascending within a column (Data.h:512-524) */
    float* x;
    float* y;
    double* e;           /* e_q_term.e / .q, [G]:52-55 */
    double* q;
} fm_data;

struct fm_oracle {
    fm_oracle_config c;
    uint32_t* group;          /* [num_attr] */
    uint32_t* n_per_group;    /* [G] */
    fm_data tr, te;
    double w0, *w, *v;        /* v[f * p + i] */
    double alpha, alpha_0, gamma_0, beta_0, mu_0, w0_mean_0;   /* [G]:1099-1106 */
    double *w_mu, *w_lambda, *v_mu, *v_lambda;                 /* [g], [g * K + f] */
    double* group_cache;      /* cache_for_group_values */
    double* pred_sum;         /* pred_sum_all, [GS]:154-158 */
    double min_target, max_target;
    uint32_t iter;
    int in_init;
    FILE* log;
};

/* ---- samplers (shim_random.h semantics) ---------------------------------------------------------------------------- */
static void rec(fm_oracle* o, double tag, double a, double b)
{
    if (o->log) {
        double r[3] = {tag, a, b};
        fwrite(r, sizeof(double), 3, o->log);
    }
}
static double gauss2(fm_oracle* o, double mean, double stdev)
{
    rec(o, 0.0, mean, stdev);
    if (o->c.noise == FM_ORACLE_NOISE_ZERO && !o->in_init) return mean;
    if (stdev == 0.0 || isnan(stdev)) return mean;
    return mean + stdev * ran_gaussian_leva();
}
static double gamma2(fm_oracle* o, double a, double b)
{
    rec(o, 1.0, a, b);
    if (o->c.noise == FM_ORACLE_NOISE_ZERO) return a / b;
    return ran_gamma_mt_rand(a) / b;
}

/* ---- data ----------------------------------------------------------------------------------------------------------- */
static int build_columns(fm_data* d, uint32_t n, const int64_t* row_ptr, const uint32_t* attr, const float* x, const float* y, uint32_t p)
{
    const int64_t nnz = row_ptr[n];
    d->n = n;
    d->ncol = 0;
    d->col_ptr = (int64_t*)calloc((size_t)p + 1, sizeof(int64_t));
    d->case_id = (uint32_t*)malloc((size_t)(nnz ? nnz : 1) * sizeof(uint32_t));
    d->x = (float*)malloc((size_t)(nnz ? nnz : 1) * sizeof(float));
    d->y = (float*)malloc((size_t)(n ? n : 1) * sizeof(float));
    d->e = (double*)calloc(n ? n : 1, sizeof(double));
    d->q = (double*)calloc(n ? n : 1, sizeof(double));
    if (!d->col_ptr || !d->case_id || !d->x || !d->y || !d->e || !d->q) return -1;
    memcpy(d->y, y, (size_t)n * sizeof(float));
    for (int64_t k = 0; k < nnz; ++k) {
        if (attr[k] >= p) return -1;
        d->col_ptr[attr[k] + 1]++;
        if (attr[k] + 1 > d->ncol) d->ncol = attr[k] + 1;
    }
    for (uint32_t i = 0; i < p; ++i) d->col_ptr[i + 1] += d->col_ptr[i];
    int64_t* fill = (int64_t*)malloc(((size_t)p + 1) * sizeof(int64_t));
    if (!fill) return -1;
    memcpy(fill, d->col_ptr, ((size_t)p + 1) * sizeof(int64_t));
    for (uint32_t r = 0; r < n; ++r)                     /* Data.h:512-524: rows in order => ascending case ids per column */
        for (int64_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
            const int64_t at = fill[attr[k]]++;
            d->case_id[at] = r;
            d->x[at] = x[k];
        }
    free(fill);
    return 0;
}
static void free_data(fm_data* d)
{
    free(d->col_ptr); free(d->case_id); free(d->x); free(d->y); free(d->e); free(d->q);
}

fm_oracle* fm_oracle_create(const fm_oracle_config* cfg, uint32_t n_train, const int64_t* row_ptr, const uint32_t* attr, const float* x,
                            const float* y, uint32_t n_test, const int64_t* t_row_ptr, const uint32_t* t_attr, const float* t_x,
                            const float* t_y, const uint32_t* attr_group)
{
    if (!cfg || cfg->num_attr == 0 || cfg->num_groups == 0) return NULL;
    fm_oracle* o = (fm_oracle*)calloc(1, sizeof(fm_oracle));
    if (!o) return NULL;
    o->c = *cfg;
    const uint32_t p = cfg->num_attr, G = cfg->num_groups, K = cfg->K;
    o->group = (uint32_t*)calloc(p, sizeof(uint32_t));
    o->n_per_group = (uint32_t*)calloc(G, sizeof(uint32_t));
    o->w = (double*)calloc(p, sizeof(double));
    o->v = (double*)calloc((size_t)(K ? K : 1) * p, sizeof(double));
    o->w_mu = (double*)calloc(G, sizeof(double));
    o->w_lambda = (double*)calloc(G, sizeof(double));
    o->v_mu = (double*)calloc((size_t)G * (K ? K : 1), sizeof(double));
    o->v_lambda = (double*)calloc((size_t)G * (K ? K : 1), sizeof(double));
    o->group_cache = (double*)calloc(G, sizeof(double));
    o->pred_sum = (double*)calloc(n_test ? n_test : 1, sizeof(double));
    int bad = !o->group || !o->n_per_group || !o->w || !o->v || !o->w_mu || !o->w_lambda || !o->v_mu || !o->v_lambda || !o->group_cache ||
              !o->pred_sum;
    if (!bad) bad = build_columns(&o->tr, n_train, row_ptr, attr, x, y, p) || build_columns(&o->te, n_test, t_row_ptr, t_attr, t_x, t_y, p);
    for (uint32_t i = 0; i < p && !bad; ++i) {
        o->group[i] = attr_group ? attr_group[i] : 0;
        if (o->group[i] >= G) bad = 1;
        else o->n_per_group[o->group[i]]++;
    }
    if (bad) {
        fm_oracle_destroy(o);
        return NULL;
    }
    /* Data.h:193-201: min / max over the TRAIN targets (float) */
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
    for (uint32_t c = 0; c < n_train; ++c) {
        if (y[c] < mn) mn = y[c];
        if (y[c] > mx) mx = y[c];
    }
    o->min_target = mn;
    o->max_target = mx;
    return o;
}

void fm_oracle_destroy(fm_oracle* o)
{
    if (!o) return;
    if (o->log) fclose(o->log);
    free_data(&o->tr);
    free_data(&o->te);
    free(o->group); free(o->n_per_group); free(o->w); free(o->v); free(o->w_mu); free(o->w_lambda); free(o->v_mu); free(o->v_lambda);
    free(o->group_cache); free(o->pred_sum);
    free(o);
}

void fm_oracle_srand(unsigned seed) { srand(seed); }

int fm_oracle_set_log(fm_oracle* o, const char* path)
{
    if (o->log) fclose(o->log);
    o->log = NULL;
    if (path && path[0]) {
        o->log = fopen(path, "wb");
        if (!o->log) return -1;
    }
    return 0;
}

/* ---- [G]:117-349 without relations: e := prediction, q := 0 ---------------------------------------------------------- */
static void predict_to_eterms(fm_oracle* o, fm_data* d)
{
    const uint32_t p = o->c.num_attr;
    for (uint32_t c = 0; c < d->n; ++c) d->e[c] = d->q[c] = 0.0;
    for (uint32_t f = 0; f < o->c.K; ++f) {              /* (1) e = 1/2 sum_f (sum_i v_if x_i)^2 */
        const double* v = o->v + (size_t)f * p;
        for (uint32_t i = 0; i < d->ncol; ++i) {
            const double v_if = v[i];
            for (int64_t k = d->col_ptr[i]; k < d->col_ptr[i + 1]; ++k) d->q[d->case_id[k]] += v_if * d->x[k];
        }
        for (uint32_t c = 0; c < d->n; ++c) {
            const double q_all = d->q[c];
            d->e[c] += 0.5 * q_all * q_all;
            d->q[c] = 0.0;
        }
    }
    for (uint32_t f = 0; f < o->c.K; ++f) {              /* (2) q = -1/2 sum_f sum_i v_if^2 x_i^2 */
        const double* v = o->v + (size_t)f * p;
        for (uint32_t i = 0; i < d->ncol; ++i) {
            const double v_if = v[i];
            for (int64_t k = d->col_ptr[i]; k < d->col_ptr[i + 1]; ++k) {
                const float x_li = d->x[k];
                d->q[d->case_id[k]] -= 0.5 * v_if * v_if * x_li * x_li;
            }
        }
    }
    if (o->c.k1)                                          /* (3) q += sum_i w_i x_i */
        for (uint32_t i = 0; i < d->ncol; ++i) {
            const double w_i = o->w[i];
            for (int64_t k = d->col_ptr[i]; k < d->col_ptr[i + 1]; ++k) d->q[d->case_id[k]] += w_i * d->x[k];
        }
    for (uint32_t c = 0; c < d->n; ++c) {
        const double q_all = d->q[c];
        d->e[c] = d->e[c] + q_all;
        if (o->c.k0) d->e[c] += o->w0;
        d->q[c] = 0.0;
    }
}

void fm_oracle_init(fm_oracle* o, const double* w_init, const double* v_init)
{
    const uint32_t p = o->c.num_attr, G = o->c.num_groups, K = o->c.K;
    o->in_init = 1;
    o->w0 = 0.0;
    for (uint32_t f = 0; f < K; ++f)                      /* fm_model.h:96: v.init(init_mean = 0, init_stdev) */
        for (uint32_t i = 0; i < p; ++i) o->v[(size_t)f * p + i] = v_init ? v_init[(size_t)f * p + i] : gauss2(o, 0.0, o->c.init_stdev);
    for (uint32_t i = 0; i < p; ++i) o->w[i] = w_init ? w_init[i] : gauss2(o, 0.0, o->c.init_stdev);   /* [L]:412 (also when k1 == 0) */
    o->in_init = 0;
    o->alpha_0 = 1.0; o->gamma_0 = 1.0; o->beta_0 = 1.0; o->mu_0 = 0.0;   /* [G]:1099-1106 */
    o->alpha = 1;
    o->w0_mean_0 = 0.0;
    for (uint32_t g = 0; g < G; ++g) {
        o->w_mu[g] = 0.0;
        o->w_lambda[g] = o->c.regw;                        /* [L]:485-505 */
        for (uint32_t f = 0; f < K; ++f) {
            o->v_mu[(size_t)g * K + f] = 0.0;
            o->v_lambda[(size_t)g * K + f] = o->c.regv;
        }
    }
    memset(o->pred_sum, 0, (size_t)(o->te.n ? o->te.n : 1) * sizeof(double));
    o->iter = 0;
    predict_to_eterms(o, &o->tr);                          /* [GS]:73-78 */
    predict_to_eterms(o, &o->te);
    for (uint32_t c = 0; c < o->tr.n; ++c) o->tr.e[c] = o->tr.e[c] - o->tr.y[c];
}

/* ---- the draws ------------------------------------------------------------------------------------------------------- */
static void draw_alpha(fm_oracle* o)                      /* [G]:901-929 */
{
    if (!o->c.do_multilevel) {
        o->alpha = o->alpha_0;
        return;
    }
    const uint32_t n = o->tr.n;
    const double alpha_n = o->alpha_0 + n;
    double gamma_n = o->gamma_0;
    for (uint32_t i = 0; i < n; ++i) gamma_n += o->tr.e[i] * o->tr.e[i];
    const double alpha_old = o->alpha;
    o->alpha = gamma2(o, alpha_n / 2.0, gamma_n / 2.0);
    if (isnan(o->alpha) || isinf(o->alpha)) o->alpha = alpha_old;
}

static void draw_w0(fm_oracle* o)                         /* [G]:628-668 */
{
    const uint32_t n = o->tr.n;
    double w0_mean = 0;
    for (uint32_t i = 0; i < n; ++i) w0_mean += o->tr.e[i] - o->w0;
    const double w0_sigma_sqr = (double)1.0 / (o->c.reg0 + o->alpha * n);
    w0_mean = -w0_sigma_sqr * (o->alpha * w0_mean - o->w0_mean_0 * o->c.reg0);
    const double w0_old = o->w0;
    o->w0 = o->c.do_sample ? gauss2(o, w0_mean, sqrt(w0_sigma_sqr)) : w0_mean;
    if (isnan(o->w0) || isinf(o->w0)) {
        o->w0 = w0_old;
        return;
    }
    for (uint32_t i = 0; i < n; ++i) o->tr.e[i] -= (w0_old - o->w0);
}

static int draw_w_lambda(fm_oracle* o)                    /* [G]:970-1007 */
{
    if (!o->c.do_multilevel) return 0;
    const uint32_t p = o->c.num_attr, G = o->c.num_groups;
    double* lg = o->group_cache;
    for (uint32_t g = 0; g < G; ++g) lg[g] = o->beta_0 * (o->w_mu[g] - o->mu_0) * (o->w_mu[g] - o->mu_0) + o->gamma_0;
    for (uint32_t i = 0; i < p; ++i) {
        const uint32_t g = o->group[i];
        lg[g] += (o->w[i] - o->w_mu[g]) * (o->w[i] - o->w_mu[g]);
    }
    for (uint32_t g = 0; g < G; ++g) {
        const double a = o->alpha_0 + o->n_per_group[g] + 1;
        const double old = o->w_lambda[g];
        o->w_lambda[g] = o->c.do_sample ? gamma2(o, a / 2.0, lg[g] / 2.0) : a / lg[g];
        if (isnan(o->w_lambda[g]) || isinf(o->w_lambda[g])) {
            o->w_lambda[g] = old;
            return 1;                                      /* libFM returns from the whole function here */
        }
    }
    return 0;
}

static int draw_w_mu(fm_oracle* o)                        /* [G]:931-968 */
{
    const uint32_t p = o->c.num_attr, G = o->c.num_groups;
    if (!o->c.do_multilevel) {
        for (uint32_t g = 0; g < G; ++g) o->w_mu[g] = o->mu_0;
        return 0;
    }
    double* mm = o->group_cache;
    for (uint32_t g = 0; g < G; ++g) mm[g] = 0.0;
    for (uint32_t i = 0; i < p; ++i) mm[o->group[i]] += o->w[i];
    for (uint32_t g = 0; g < G; ++g) {
        mm[g] = (mm[g] + o->beta_0 * o->mu_0) / (o->n_per_group[g] + o->beta_0);
        const double s2 = (double)1.0 / ((o->n_per_group[g] + o->beta_0) * o->w_lambda[g]);
        const double old = o->w_mu[g];
        o->w_mu[g] = o->c.do_sample ? gauss2(o, mm[g], sqrt(s2)) : mm[g];
        if (isnan(o->w_mu[g]) || isinf(o->w_mu[g])) {
            o->w_mu[g] = old;
            return 1;
        }
    }
    return 0;
}

static void draw_w(fm_oracle* o, uint32_t i)              /* [G]:671-719 */
{
    fm_data* d = &o->tr;
    const uint32_t g = o->group[i];
    double* w = &o->w[i];
    const double w_mu = o->w_mu[g], w_lambda = o->w_lambda[g];
    double w_sigma_sqr = 0, w_mean = 0;
    for (int64_t k = d->col_ptr[i]; k < d->col_ptr[i + 1]; ++k) {
        const float x_li = d->x[k];
        w_mean += x_li * (d->e[d->case_id[k]] - *w * x_li);
        w_sigma_sqr += x_li * x_li;                        /* float product, as in libFM (FM_FLOAT x_li) */
    }
    w_sigma_sqr = (double)1.0 / (w_lambda + o->alpha * w_sigma_sqr);
    w_mean = -w_sigma_sqr * (o->alpha * w_mean - w_mu * w_lambda);
    const double w_old = *w;
    if (isnan(w_sigma_sqr) || isinf(w_sigma_sqr)) *w = 0.0;
    else *w = o->c.do_sample ? gauss2(o, w_mean, sqrt(w_sigma_sqr)) : w_mean;
    if (isnan(*w) || isinf(*w)) {
        *w = w_old;
        return;
    }
    for (int64_t k = d->col_ptr[i]; k < d->col_ptr[i + 1]; ++k) {
        const double h = d->x[k];
        d->e[d->case_id[k]] -= h * (w_old - *w);
    }
}

static void draw_v_lambda(fm_oracle* o)                   /* [G]:1051-1088 */
{
    if (!o->c.do_multilevel) return;
    const uint32_t p = o->c.num_attr, G = o->c.num_groups, K = o->c.K;
    double* lg = o->group_cache;
    for (uint32_t f = 0; f < K; ++f) {
        const double* v = o->v + (size_t)f * p;
        for (uint32_t g = 0; g < G; ++g) {
            const double m = o->v_mu[(size_t)g * K + f];
            lg[g] = o->beta_0 * (m - o->mu_0) * (m - o->mu_0) + o->gamma_0;
        }
        for (uint32_t i = 0; i < p; ++i) {
            const uint32_t g = o->group[i];
            const double m = o->v_mu[(size_t)g * K + f];
            lg[g] += (v[i] - m) * (v[i] - m);
        }
        for (uint32_t g = 0; g < G; ++g) {
            const double a = o->alpha_0 + o->n_per_group[g] + 1;
            double* lam = &o->v_lambda[(size_t)g * K + f];
            const double old = *lam;
            *lam = o->c.do_sample ? gamma2(o, a / 2.0, lg[g] / 2.0) : a / lg[g];
            if (isnan(*lam) || isinf(*lam)) {
                *lam = old;
                return;
            }
        }
    }
}

static void draw_v_mu(fm_oracle* o)                       /* [G]:1011-1049 */
{
    const uint32_t p = o->c.num_attr, G = o->c.num_groups, K = o->c.K;
    if (!o->c.do_multilevel) {
        for (size_t t = 0; t < (size_t)G * K; ++t) o->v_mu[t] = o->mu_0;
        return;
    }
    double* mm = o->group_cache;
    for (uint32_t f = 0; f < K; ++f) {
        const double* v = o->v + (size_t)f * p;
        for (uint32_t g = 0; g < G; ++g) mm[g] = 0.0;
        for (uint32_t i = 0; i < p; ++i) mm[o->group[i]] += v[i];
        for (uint32_t g = 0; g < G; ++g) {
            mm[g] = (mm[g] + o->beta_0 * o->mu_0) / (o->n_per_group[g] + o->beta_0);
            const double s2 = (double)1.0 / ((o->n_per_group[g] + o->beta_0) * o->v_lambda[(size_t)g * K + f]);
            double* mu = &o->v_mu[(size_t)g * K + f];
            const double old = *mu;
            *mu = o->c.do_sample ? gauss2(o, mm[g], sqrt(s2)) : mm[g];
            if (isnan(*mu) || isinf(*mu)) {
                *mu = old;
                return;
            }
        }
    }
}

static void draw_v(fm_oracle* o, uint32_t f, uint32_t i)  /* [G]:780-836 */
{
    fm_data* d = &o->tr;
    const uint32_t g = o->group[i], K = o->c.K;
    double* v = &o->v[(size_t)f * o->c.num_attr + i];
    const double v_mu = o->v_mu[(size_t)g * K + f], v_lambda = o->v_lambda[(size_t)g * K + f];
    double v_sigma_sqr = 0, v_mean = 0;
    for (int64_t k = d->col_ptr[i]; k < d->col_ptr[i + 1]; ++k) {
        const float x_li = d->x[k];
        const uint32_t c = d->case_id[k];
        const double h = x_li * (d->q[c] - x_li * *v);
        v_mean += h * d->e[c];
        v_sigma_sqr += h * h;
    }
    v_mean -= *v * v_sigma_sqr;
    v_sigma_sqr = (double)1.0 / (v_lambda + o->alpha * v_sigma_sqr);
    v_mean = -v_sigma_sqr * (o->alpha * v_mean - v_mu * v_lambda);
    const double v_old = *v;
    if (isnan(v_sigma_sqr) || isinf(v_sigma_sqr)) *v = 0.0;
    else *v = o->c.do_sample ? gauss2(o, v_mean, sqrt(v_sigma_sqr)) : v_mean;
    if (isnan(*v) || isinf(*v)) {
        *v = v_old;
        return;
    }
    for (int64_t k = d->col_ptr[i]; k < d->col_ptr[i + 1]; ++k) {
        const float x_li = d->x[k];
        const uint32_t c = d->case_id[k];
        const double h = x_li * (d->q[c] - x_li * v_old);
        d->q[c] -= x_li * (v_old - *v);
        d->e[c] -= h * (v_old - *v);
    }
}

static void draw_all(fm_oracle* o)                        /* [G]:411-626 */
{
    fm_data* d = &o->tr;
    const uint32_t p = o->c.num_attr;
    draw_alpha(o);
    if (o->c.k0) draw_w0(o);
    if (o->c.k1) {
        draw_w_lambda(o);
        draw_w_mu(o);
        for (uint32_t i = 0; i < p; ++i) draw_w(o, i);    /* columns of the train file, then the attributes it never mentions */
    }
    if (o->c.K > 0) {
        draw_v_lambda(o);
        draw_v_mu(o);
    }
    for (uint32_t f = 0; f < o->c.K; ++f) {
        const double* v = o->v + (size_t)f * p;
        for (uint32_t c = 0; c < d->n; ++c) d->q[c] = 0.0;
        for (uint32_t i = 0; i < d->ncol; ++i) {          /* add_main_q, [G]:384-409 */
            const double v_if = v[i];
            for (int64_t k = d->col_ptr[i]; k < d->col_ptr[i + 1]; ++k) d->q[d->case_id[k]] += v_if * d->x[k];
        }
        for (uint32_t i = 0; i < p; ++i) draw_v(o, f, i);
    }
}

void fm_oracle_learn(fm_oracle* o, uint32_t iters, double* rmse_train_out, double* rmse_test_out)   /* [GS]:97-262 */
{
    fm_data *tr = &o->tr, *te = &o->te;
    for (uint32_t it = 0; it < iters; ++it, ++o->iter) {
        draw_all(o);
        predict_to_eterms(o, tr);
        predict_to_eterms(o, te);
        for (uint32_t c = 0; c < te->n; ++c) {
            double pr = te->e[c];
            pr = fmin(o->max_target, pr);
            pr = fmax(o->min_target, pr);
            o->pred_sum[c] += pr;
        }
        double rmse_train = 0.0;
        for (uint32_t c = 0; c < tr->n; ++c) {
            double pr = tr->e[c];
            pr = fmin(o->max_target, pr);
            pr = fmax(o->min_target, pr);
            const double err = pr - tr->y[c];
            rmse_train += err * err;
            tr->e[c] = tr->e[c] - tr->y[c];
        }
        rmse_train = sqrt(rmse_train / tr->n);
        const double normalizer = 1.0 / (o->iter + 1);     /* [GS]:241, 307-326 */
        double r = 0;
        uint32_t cnt = 0;
        for (uint32_t c = 0; c < te->n; ++c) {
            double pr = o->pred_sum[c] * normalizer;
            pr = fmin(o->max_target, pr);
            pr = fmax(o->min_target, pr);
            const double err = pr - te->y[c];
            r += err * err;
            cnt++;
        }
        if (rmse_train_out) rmse_train_out[it] = rmse_train;
        if (rmse_test_out) rmse_test_out[it] = sqrt(r / cnt);
    }
    if (o->log) fflush(o->log);
}

void fm_oracle_get_state(const fm_oracle* o, double* w, double* v, double* w_mu, double* w_lambda, double* v_mu, double* v_lambda,
                         double* e, double* pred_sum, double* scal)
{
    const uint32_t p = o->c.num_attr, G = o->c.num_groups, K = o->c.K;
    if (w) memcpy(w, o->w, (size_t)p * 8);
    if (v) memcpy(v, o->v, (size_t)K * p * 8);
    if (w_mu) memcpy(w_mu, o->w_mu, (size_t)G * 8);
    if (w_lambda) memcpy(w_lambda, o->w_lambda, (size_t)G * 8);
    if (v_mu) memcpy(v_mu, o->v_mu, (size_t)G * K * 8);
    if (v_lambda) memcpy(v_lambda, o->v_lambda, (size_t)G * K * 8);
    if (e) memcpy(e, o->tr.e, (size_t)o->tr.n * 8);
    if (pred_sum) memcpy(pred_sum, o->pred_sum, (size_t)o->te.n * 8);
    if (scal) {
        scal[0] = o->w0;
        scal[1] = o->alpha;
    }
}

void fm_oracle_get_columns(const fm_oracle* o, int64_t* col_ptr, uint32_t* case_id, float* x)
{
    const int64_t nnz = o->tr.col_ptr[o->c.num_attr];
    if (col_ptr) memcpy(col_ptr, o->tr.col_ptr, ((size_t)o->c.num_attr + 1) * 8);
    if (case_id) memcpy(case_id, o->tr.case_id, (size_t)nnz * 4);
    if (x) memcpy(x, o->tr.x, (size_t)nnz * 4);
}
