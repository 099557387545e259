// api.cu -- the extern "C" sbmf_cuda_* boundary (include/sbmf_cuda.h) over the device model.
// Nothing here computes on the CPU: every phase of the sweep is a kernel in kernels.cu, and there is no
// fallback -- sbmf_cuda_create fails if no CUDA device of compute capability 10.x is usable.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "model.h"

using namespace sbmf;

static thread_local std::string g_create_err;

#define API_CK(call)                                                                               \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            m.err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            return (e_ == cudaErrorMemoryAllocation) ? SBMF_ERR_NOMEM : SBMF_ERR_CUDA;             \
        }                                                                                          \
    } while (0)

// frees a temporary device buffer on every exit path (API_CK returns early)
struct DeviceTmp {
    float* p = nullptr;
    ~DeviceTmp() { if (p) cudaFree(p); }
};

static cudaError_t copy_out_widen(uint64_t* dst, const uint32_t* d_src, size_t n, std::vector<uint32_t>& tmp)
{
    tmp.resize(n ? n : 1);
    cudaError_t e = cudaMemcpy(tmp.data(), d_src, n * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return e;
    for (size_t i = 0; i < n; ++i) dst[i] = tmp[i];
    return cudaSuccess;
}

struct OptionDesc {
    const char* name;
    int64_t Options::*field;
    int64_t lo, hi;
    bool before_train;   // shapes the layout / work lists: only accepted before set_train
};
static const OptionDesc kOptions[] = {
    {"l2_budget_mb", &Options::l2_budget_mb, 1, 1 << 20, false},
    {"max_blocks_per_launch", &Options::max_blocks_per_launch, 0, 1 << 20, false},
    {"resident_max", &Options::resident_max, 1, RESIDENT_MAX, true},
    {"resident_max_user", &Options::resident_max_user, 1, RESIDENT_MAX, true},
    {"resident_max_item", &Options::resident_max_item, 1, RESIDENT_MAX, true},
    {"slice_len", &Options::slice_len, 0, 1 << 24, true},
    {"relabel", &Options::relabel, 0, 1, true},
    {"group_rows", &Options::group_rows, 0, 1, false},
    {"row_kernels", &Options::row_kernels, 1, 3, false},
    {"alt_bins", &Options::alt_bins, 0, 1, false},
    {"pair_gather", &Options::pair_gather, 0, 1, true},
    {"fuse_solve", &Options::fuse_solve, 0, 1, false},
    {"heavy_chains", &Options::heavy_chains, 1, 2, false},
    {"fold_user", &Options::fold_user, 0, 1, false},
    {"fold_item", &Options::fold_item, 0, 1, false},
    {"graph", &Options::graph, 0, 1, false},
    {"device_plan", &Options::device_plan, 0, 1, true},
    {"mgpu_pool", &Options::mgpu_pool, 0, 1, true},
    {"fuse_exchange", &Options::fuse_exchange, 0, 1, true},
    {"peer", &Options::peer, 0, 1, true},
    {"trace", &Options::trace, 0, 2, false},
};

static const OptionDesc* find_option(const char* name)
{
    if (!name) return nullptr;
    for (const OptionDesc& d : kOptions)
        if (!strcmp(d.name, name)) return &d;
    return nullptr;
}

extern "C" {

int sbmf_cuda_abi_version(void) { return SBMF_CUDA_ABI_VERSION; }

int sbmf_cuda_set_option(sbmf_handle* h, const char* name, int64_t value)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    const OptionDesc* d = find_option(name);
    if (!d) {
        m.err = std::string("set_option: unknown option '") + (name ? name : "(null)") + "'";
        return SBMF_ERR_INVALID;
    }
    if (value < d->lo || value > d->hi) {
        m.err = std::string("set_option: ") + name + " must be in [" + std::to_string(d->lo) + ", " + std::to_string(d->hi) + "]";
        return SBMF_ERR_INVALID;
    }
    if (d->before_train && m.have_train) {
        m.err = std::string("set_option: ") + name + " shapes the layout and must be set before set_train";
        return SBMF_ERR_STATE;
    }
    m.opt.*(d->field) = value;
    if (d->field == &Options::resident_max) m.opt.resident_max_user = m.opt.resident_max_item = value;
    if (m.graph_exec) {   // the captured sweep may embed the old choice
        cudaSetDevice(m.device);
        cudaStreamSynchronize(m.s_main);
        cudaGraphExecDestroy(m.graph_exec);
        m.graph_exec = nullptr;
    }
    m.graph_failed = false;
    return SBMF_OK;
}

int sbmf_cuda_get_option(sbmf_handle* h, const char* name, int64_t* value)
{
    if (!h || !value) return SBMF_ERR_INVALID;
    const OptionDesc* d = find_option(name);
    if (!d) {
        h->m.err = std::string("get_option: unknown option '") + (name ? name : "(null)") + "'";
        return SBMF_ERR_INVALID;
    }
    *value = h->m.opt.*(d->field);
    return SBMF_OK;
}

// SBMF_OPTIONS="name=value,name=value": the ONE environment variable of the library, read once per handle in
// sbmf_cuda_create and applied through sbmf_cuda_set_option (A/B scripts); a malformed entry fails the create.
static int apply_env_options(sbmf_handle* h)
{
    const char* ev = getenv("SBMF_OPTIONS");
    if (!ev || !*ev) return SBMF_OK;
    std::string s(ev);
    size_t pos = 0;
    while (pos < s.size()) {
        size_t end = s.find(',', pos);
        if (end == std::string::npos) end = s.size();
        const std::string item = s.substr(pos, end - pos);
        pos = end + 1;
        if (item.empty()) continue;
        const size_t eq = item.find('=');
        if (eq == std::string::npos) {
            g_create_err = "create: SBMF_OPTIONS entry '" + item + "' is not name=value";
            return SBMF_ERR_INVALID;
        }
        char* endp = nullptr;
        const long long v = strtoll(item.c_str() + eq + 1, &endp, 10);
        if (!endp || *endp) {
            g_create_err = "create: SBMF_OPTIONS entry '" + item + "': value is not an integer";
            return SBMF_ERR_INVALID;
        }
        const int rc = sbmf_cuda_set_option(h, item.substr(0, eq).c_str(), (int64_t)v);
        if (rc != SBMF_OK) {
            g_create_err = "create: SBMF_OPTIONS: " + h->m.err;
            return rc;
        }
    }
    return SBMF_OK;
}

int sbmf_cuda_config_default(sbmf_config* cfg)
{
    if (!cfg) return SBMF_ERR_INVALID;
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = (uint32_t)sizeof(sbmf_config);
    cfg->K = 20;                                   // [T]:224
    cfg->device = 0;
    cfg->sample_mode = SBMF_SAMPLE_REF_VAR_AS_STDEV;
    cfg->hyper_mode = SBMF_HYPER_REF_T;
    cfg->rebuild_every = 1;                        // [T]:342-359 rebuilds E every sweep
    cfg->residual_mode = 0;                        // ... fused into the user phase; 1 = stand-alone rebuild kernel
    cfg->burn_in = 0;                              // [T]:323
    cfg->seed = 1;
    cfg->init_stdev = 0.1;                         // [T]:242, 248
    cfg->clamp_lo = 0.5;                           // [T]:627-628
    cfg->clamp_hi = 5.0;
    for (int i = 0; i < 6; ++i) {                  // [T]:284-313
        cfg->priors.alpha[i] = 1.0;
        cfg->priors.beta[i] = 1.0;
        cfg->priors.mu[i] = 0.0;
        cfg->priors.sigma[i] = 1.0;
    }
    cfg->priors.alpha_dash = 1.0;
    cfg->priors.beta_dash = 1.0;
    cfg->priors.ng_a_0 = cfg->priors.ng_b_0 = cfg->priors.ng_alpha_0 = cfg->priors.ng_beta_0 = cfg->priors.ng_nu_0 = 1.0;   // [S]:260-269
    cfg->priors.ng_mu_0 = 0.0;
    cfg->rank = 0;
    cfg->world_size = 1;
    return SBMF_OK;
}

int sbmf_cuda_create(const sbmf_config* cfg, sbmf_handle** out)
{
    if (!cfg || !out) {
        g_create_err = "create: null argument";
        return SBMF_ERR_INVALID;
    }
    *out = nullptr;
    if (cfg->struct_size != sizeof(sbmf_config)) {
        g_create_err = "create: sbmf_config.struct_size mismatch (use sbmf_cuda_config_default)";
        return SBMF_ERR_INVALID;
    }
    if (cfg->K == 0 || cfg->K > SBMF_MAX_K) {
        g_create_err = "create: K must be in [1, SBMF_MAX_K]";
        return SBMF_ERR_INVALID;
    }
    if (cfg->sample_mode < 0 || cfg->sample_mode > 2 || cfg->hyper_mode < 0 || cfg->hyper_mode > 2) {
        g_create_err = "create: unknown sample_mode / hyper_mode";
        return SBMF_ERR_INVALID;
    }
    {
        // Every Gamma draw of the sweep has shape = prior + a non-negative count and rate = prior + a non-negative sum of
        // squares ([T]:366-511, [S]:339-413): positive, finite priors keep every shape and rate positive whatever the data
        // (including an empty training set).  Anything else would hand ran_gamma a non-positive argument ([R]:119 asserts).
        const sbmf_priors& p = cfg->priors;
        bool ok = true;
        auto pos = [&](double v) { ok = ok && v > 0.0 && v < 1e300; };
        auto fin = [&](double v) { ok = ok && v > -1e300 && v < 1e300; };
        for (int i = 0; i < 6; ++i) {
            pos(p.alpha[i]); pos(p.beta[i]); pos(p.sigma[i]); fin(p.mu[i]);
        }
        pos(p.alpha_dash); pos(p.beta_dash);
        pos(p.ng_a_0); pos(p.ng_b_0); pos(p.ng_alpha_0); pos(p.ng_beta_0); pos(p.ng_nu_0); fin(p.ng_mu_0);
        if (!ok) {
            g_create_err = "create: priors must be finite, and every shape / rate / precision prior (alpha, beta, sigma, *_dash, ng_*) > 0";
            return SBMF_ERR_INVALID;
        }
        if (!(cfg->init_stdev >= 0.0) || !(cfg->clamp_lo <= cfg->clamp_hi)) {
            g_create_err = "create: need init_stdev >= 0 and clamp_lo <= clamp_hi";
            return SBMF_ERR_INVALID;
        }
    }
    if (cfg->world_size < 1 || cfg->world_size > 64 || cfg->rank < 0 || cfg->rank >= cfg->world_size) {
        g_create_err = "create: need 0 <= rank < world_size <= 64";
        return SBMF_ERR_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
        return SBMF_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) {
        g_create_err = "create: device ordinal out of range";
        return SBMF_ERR_INVALID;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) {
        g_create_err = std::string("create: cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        return SBMF_ERR_CUDA;
    }
    if (prop.major != 10) {
        g_create_err = "create: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                       ", this library is built for sm_100a (B200) only";
        return SBMF_ERR_UNSUPPORTED;
    }
    sbmf_handle* h = new (std::nothrow) sbmf_handle();
    if (!h) {
        g_create_err = "create: out of host memory";
        return SBMF_ERR_NOMEM;
    }
    Model& m = h->m;
    m.cfg = *cfg;
    if (m.cfg.rebuild_every == 0) m.cfg.rebuild_every = 1;
    m.device = cfg->device;
    m.sm_count = prop.multiProcessorCount;
    m.rank = cfg->rank;
    m.world = cfg->world_size;
    m.K = cfg->K;
    m.KB = (cfg->K + KBLK - 1) / KBLK;
    m.KP = m.KB * KBLK;
    m.us.site_f = SITE_U; m.us.site_b = SITE_BI; m.us.site_sigma_k = SITE_SIGMA_U; m.us.site_mu_k = SITE_MU_U;
    m.us.site_sigma_b = SITE_SIGMA_BI; m.us.site_mu_b = SITE_MU_BI; m.us.prior = 2; m.us.prior_b = 4;
    m.it.site_f = SITE_V; m.it.site_b = SITE_BJ; m.it.site_sigma_k = SITE_SIGMA_V; m.it.site_mu_k = SITE_MU_V;
    m.it.site_sigma_b = SITE_SIGMA_BJ; m.it.site_mu_b = SITE_MU_BJ; m.it.prior = 1; m.it.prior_b = 5;
    bool ok = cudaSetDevice(m.device) == cudaSuccess;
    if (ok) init_constant_tables();
    if (ok) {   // keep freed stream-ordered allocations (set_train temporaries) in the pool instead of returning them to the OS
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, m.device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    ok = ok && cudaStreamCreateWithFlags(&m.s_main, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&m.s_aux, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&m.s_aux2, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&m.ev_join2, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) {
        ok = cudaStreamCreateWithFlags(&m.s_res[i], cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&m.ev_join_res[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&m.ev_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&m.ev_join, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 8 && ok; ++i) ok = cudaEventCreate(&m.ev_t[i]) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) ok = cudaEventCreate(&m.ev_call[i]) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) ok = cudaEventCreate(&m.ev_c[i]) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&m.sc, sizeof(Scalars)) == cudaSuccess;
    ok = ok && cudaMemset(m.sc, 0, sizeof(Scalars)) == cudaSuccess;
    m.hist_cap = 4096;
    ok = ok && cudaMalloc((void**)&m.rmse_hist, (size_t)m.hist_cap * 2 * sizeof(double)) == cudaSuccess;
    if (!ok) {
        g_create_err = std::string("create: CUDA resource setup failed: ") + cudaGetErrorString(cudaGetLastError());
        sbmf_cuda_destroy(h);
        return SBMF_ERR_CUDA;
    }
    if (m.world > 1) {   // one communicator per rank; the id comes from sbmf_cuda_nccl_unique_id on rank 0
        std::string err;
        if (comm_init(m.comm, cfg->nccl_id, m.rank, m.world, err) != 0) {
            g_create_err = "create: NCCL: " + err;
            sbmf_cuda_destroy(h);
            return SBMF_ERR_NCCL;
        }
        // NCCL connects peers lazily on the first collective that needs them; do that here, once, with one small instance of
        // every pattern the sweep uses (all-reduce, broadcast from every root, all-to-all), not inside the first sweep
        {
            float* wf = nullptr;
            double* wd = nullptr;
            const int G = m.world;
            bool wok = cudaMalloc((void**)&wf, (size_t)G * 2 * 256 * 4) == cudaSuccess && cudaMalloc((void**)&wd, 16) == cudaSuccess;
            wok = wok && cudaMemset(wf, 0, (size_t)G * 2 * 256 * 4) == cudaSuccess && cudaMemset(wd, 0, 16) == cudaSuccess;
            if (wok) {
                std::vector<size_t> off(G), cnt(G, 256);
                for (int q = 0; q < G; ++q) off[q] = (size_t)q * 256;
                wok = comm_allreduce_sum_f64(m.comm, wd, 2, m.s_main, err) == 0;
                wok = wok && comm_allgatherv_f32(m.comm, wf, off.data(), cnt.data(), m.s_main, err) == 0;
                wok = wok && comm_alltoallv_f32(m.comm, wf, off.data(), cnt.data(), wf + (size_t)G * 256, off.data(), cnt.data(), m.s_main, err) == 0;
                wok = wok && cudaStreamSynchronize(m.s_main) == cudaSuccess;
            }
            cudaFree(wf);
            cudaFree(wd);
            if (!wok) {
                g_create_err = "create: NCCL warm-up failed: " + err;
                sbmf_cuda_destroy(h);
                return SBMF_ERR_NCCL;
            }
        }
    }
    {
        const int orc = apply_env_options(h);
        if (orc != SBMF_OK) {
            sbmf_cuda_destroy(h);
            return orc;
        }
    }
    *out = h;
    return SBMF_OK;
}

int sbmf_cuda_destroy(sbmf_handle* h)
{
    if (!h) return SBMF_OK;
    Model& m = h->m;
    cudaSetDevice(m.device);
    if (m.s_main) cudaStreamSynchronize(m.s_main);
    if (m.s_aux) cudaStreamSynchronize(m.s_aux);
    if (m.s_aux2) cudaStreamSynchronize(m.s_aux2);
    free_storage(m);
    free_test(m);
    comm_destroy(m.comm);
    cudaFree(m.sc);
    cudaFree(m.rmse_hist);
    for (int i = 0; i < 8; ++i)
        if (m.ev_t[i]) cudaEventDestroy(m.ev_t[i]);
    for (int i = 0; i < 2; ++i)
        if (m.ev_call[i]) cudaEventDestroy(m.ev_call[i]);
    for (int i = 0; i < 2; ++i)
        if (m.ev_c[i]) cudaEventDestroy(m.ev_c[i]);
    for (cudaEvent_t ev : m.ev_top) cudaEventDestroy(ev);
    if (m.graph_exec) cudaGraphExecDestroy(m.graph_exec);
    if (m.ev_fork) cudaEventDestroy(m.ev_fork);
    if (m.ev_join) cudaEventDestroy(m.ev_join);
    if (m.ev_join2) cudaEventDestroy(m.ev_join2);
    if (m.s_main) cudaStreamDestroy(m.s_main);
    if (m.s_aux) cudaStreamDestroy(m.s_aux);
    if (m.s_aux2) cudaStreamDestroy(m.s_aux2);
    for (int i = 0; i < 2; ++i) {
        if (m.ev_join_res[i]) cudaEventDestroy(m.ev_join_res[i]);
        if (m.s_res[i]) cudaStreamDestroy(m.s_res[i]);
    }
    delete h;
    return SBMF_OK;
}

const char* sbmf_cuda_last_error(const sbmf_handle* h) { return h ? h->m.err.c_str() : g_create_err.c_str(); }

int sbmf_cuda_nccl_unique_id(uint8_t out[128])
{
    if (!out) return SBMF_ERR_INVALID;
    std::string err;
    if (comm_unique_id(out, err) != 0) {
        g_create_err = "nccl_unique_id: " + err;
        return SBMF_ERR_NCCL;
    }
    return SBMF_OK;
}

int sbmf_cuda_set_train(sbmf_handle* h, uint64_t n, const uint32_t* user, const uint32_t* item, const float* rating, uint32_t num_users,
                        uint32_t num_items)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if ((n && (!user || !item || !rating)) || num_users == 0 || num_items == 0) {
        m.err = "set_train: null array or empty id space";
        return SBMF_ERR_INVALID;
    }
    API_CK(cudaSetDevice(m.device));
    free_test(m);
    return build_storage(m, n, user, item, rating, num_users, num_items);
}

int sbmf_cuda_set_test(sbmf_handle* h, uint64_t nt, const uint32_t* user, const uint32_t* item, const float* rating)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.have_train) {
        m.err = "set_test: call set_train first (it fixes num_users / num_items)";
        return SBMF_ERR_STATE;
    }
    if (nt && (!user || !item || !rating)) {
        m.err = "set_test: null array";
        return SBMF_ERR_INVALID;
    }
    for (uint64_t t = 0; t < nt; ++t)
        if (user[t] >= m.I || item[t] >= m.J) {
            m.err = "set_test: id out of range at test rating " + std::to_string(t) +
                    " (num_users / num_items must cover train and test, [T]:151-153)";
            return SBMF_ERR_INVALID;
        }
    API_CK(cudaSetDevice(m.device));
    return build_test(m, nt, user, item, rating);
}

static int layout_preamble(Model& m, const char* who)
{
    if (!m.have_train) {
        m.err = std::string(who) + ": no training set";
        return SBMF_ERR_STATE;
    }
    if (m.world > 1) {
        m.err = std::string(who) + ": only the single-GPU handle keeps the global layout";
        return SBMF_ERR_UNSUPPORTED;
    }
    API_CK(cudaSetDevice(m.device));
    API_CK(cudaStreamSynchronize(m.s_main));
    return SBMF_OK;
}

int sbmf_cuda_get_layout(sbmf_handle* h, int64_t* row_ptr, uint32_t* col, uint64_t* csr_id, int64_t* col_ptr, uint32_t* row, uint64_t* csc_id,
                         uint64_t* perm)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    const int rc = layout_preamble(m, "get_layout");
    if (rc != SBMF_OK) return rc;
    // relabelled model (option relabel): the stored arrays are in position space and sorted within rows; [T]'s layout of the
    // caller's ids is built from the restored COO by the same device sorts (storage.cu)
    if (m.us.id_at) return export_reference_layout(m, row_ptr, col, csr_id, col_ptr, row, csc_id, perm);
    return sbmf_cuda_get_storage_layout(h, row_ptr, col, csr_id, col_ptr, row, csc_id, perm);
}

int sbmf_cuda_get_row_positions(sbmf_handle* h, uint32_t* user_pos, uint32_t* item_pos)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.have_train) {
        m.err = "get_row_positions: no training set";
        return SBMF_ERR_STATE;
    }
    API_CK(cudaSetDevice(m.device));
    uint32_t* out[2] = {user_pos, item_pos};
    const Side* sd[2] = {&m.us, &m.it};
    for (int q = 0; q < 2; ++q) {
        if (!out[q]) continue;
        if (sd[q]->pos_of) API_CK(cudaMemcpy(out[q], sd[q]->pos_of, (size_t)sd[q]->n * 4, cudaMemcpyDeviceToHost));
        else
            for (uint32_t r = 0; r < sd[q]->n; ++r) out[q][r] = r;
    }
    return SBMF_OK;
}

int sbmf_cuda_get_storage_layout(sbmf_handle* h, int64_t* row_ptr, uint32_t* col, uint64_t* csr_id, int64_t* col_ptr, uint32_t* row,
                                 uint64_t* csc_id, uint64_t* perm)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    const int rc = layout_preamble(m, "get_storage_layout");
    if (rc != SBMF_OK) return rc;
    std::vector<uint32_t> tmp;
    if (row_ptr) API_CK(cudaMemcpy(row_ptr, m.us.ptr, ((size_t)m.I + 1) * 8, cudaMemcpyDeviceToHost));
    if (col_ptr) API_CK(cudaMemcpy(col_ptr, m.it.ptr, ((size_t)m.J + 1) * 8, cudaMemcpyDeviceToHost));
    if (col) API_CK(cudaMemcpy(col, m.us.idx, m.N * 4, cudaMemcpyDeviceToHost));
    if (row) API_CK(cudaMemcpy(row, m.it.idx, m.N * 4, cudaMemcpyDeviceToHost));
    if (csr_id) API_CK(copy_out_widen(csr_id, m.csr_id, m.N, tmp));
    if (csc_id) API_CK(copy_out_widen(csc_id, m.csc_id, m.N, tmp));
    if (perm) API_CK(copy_out_widen(perm, m.perm, m.N, tmp));
    return SBMF_OK;
}

int sbmf_cuda_init_factors(sbmf_handle* h, const float* U0, const float* V0)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.have_train) {
        m.err = "init_factors: call set_train first";
        return SBMF_ERR_STATE;
    }
    API_CK(cudaSetDevice(m.device));
    cudaStream_t st = m.s_main;
    DeviceTmp tmp;
    float*& d_tmp = tmp.p;
    const size_t nu = (size_t)m.I * m.K, nv = (size_t)m.K * m.J;
    if (U0 || V0) API_CK(cudaMalloc((void**)&d_tmp, std::max(nu, nv) * 4));
    if (U0) {
        API_CK(cudaMemcpyAsync(d_tmp, U0, nu * 4, cudaMemcpyHostToDevice, st));
        launch_load_factors(m, m.us, d_tmp, false, st);
        API_CK(cudaStreamSynchronize(st));
    } else {
        launch_init_factors(m, m.us, SITE_INIT_U, st);
    }
    if (V0) {
        API_CK(cudaMemcpyAsync(d_tmp, V0, nv * 4, cudaMemcpyHostToDevice, st));
        launch_load_factors(m, m.it, d_tmp, true, st);
        API_CK(cudaStreamSynchronize(st));
    } else {
        launch_init_factors(m, m.it, SITE_INIT_V, st);
    }
    // [T]:268-281, 315-318: biases, bias hypers, per-dimension hypers and the scalars start at 0
    for (Side* s : {&m.us, &m.it}) {
        API_CK(cudaMemsetAsync(s->bias, 0, (size_t)s->n * 4, st));
        API_CK(cudaMemsetAsync(s->mu_b, 0, (size_t)s->n * 4, st));
        API_CK(cudaMemsetAsync(s->sigma_b, 0, (size_t)s->n * 4, st));
        API_CK(cudaMemsetAsync(s->sigma_k, 0, (size_t)m.KP * 8, st));
        API_CK(cudaMemsetAsync(s->mu_k, 0, (size_t)m.KP * 8, st));
        API_CK(cudaMemsetAsync(s->sigma_kf, 0, (size_t)m.KP * 4, st));
        API_CK(cudaMemsetAsync(s->mu_kf, 0, (size_t)m.KP * 4, st));
    }
    API_CK(cudaMemsetAsync(m.sc, 0, sizeof(Scalars), st));
    API_CK(cudaMemsetAsync(m.rmse_hist, 0, (size_t)m.hist_cap * 16, st));
    if (m.have_test) API_CK(cudaMemsetAsync(m.t_sum, 0, (m.Nt ? m.Nt : 1) * 8, st));
    API_CK(cudaGetLastError());
    API_CK(cudaStreamSynchronize(st));
    m.sweeps_done = 0;
    m.sweeps_since_init = 0;
    m.e_in_csc = false;
    m.need_rebuild = false;
    m.have_factors = true;
    return SBMF_OK;
}

// One body of the loop [T]:335-637: enqueue only (no host synchronisation), so it can also be captured into a CUDA graph.
static int enqueue_sweep(Model& m, bool timing)
{
    cudaStream_t st = m.s_main;
    // Residual hygiene ([T]:342-359 rebuilds E from scratch every sweep).  residual_mode 0 (default): the same rebuild, but
    // fused into the user phase (its kernels already gather every factor the prediction needs), so the statistics of this
    // sweep come from the incrementally updated residual of the previous one; residual_mode 1: the stand-alone rebuild kernel
    // at the start of the sweep, literally as in [T].  Sweep 0 always rebuilds stand-alone (nothing to update yet).
    const bool due = (m.sweeps_done % m.cfg.rebuild_every) == 0;
    // need_rebuild: state restored without a residual (sbmf_cuda_set_state with E == NULL) -- there is nothing to update yet
    const bool standalone = m.need_rebuild || (due && (m.sweeps_done == 0 || m.cfg.residual_mode == 1));
    const bool fused = due && !standalone;
    if (timing) cudaEventRecord(m.ev_t[0], st);
    int crc = 0;
    // One GPU: the residual changes slot order (CSR <-> CSC) between the phases.  CSC -> CSR (before the user phase) is not
    // a pass of its own: the first touch of e in the user phase reads it through the inverse permutation from the item side's
    // array (PhaseArgs::e_map), where the gather hides behind the factor gathers (measured on the Netflix-shaped matrix:
    // -0.9 ms for the pass, +0.4 ms in the phase).  The other direction was a stand-alone pass in round 1 (the first streaming pass
    // turns HBM-sector-bound with the gather: -0.5 ms, +0.7 ms); on the relabelled layout of round 2 it is folded as well
    // (-0.61 ms for the pass, +0.41 ms in the phase).  Options fold_user / fold_item switch either (A/B measurements, tests).
    const bool fold_user = m.world == 1 && m.opt.fold_user;
    const bool fold = m.world == 1 && m.opt.fold_item;
    bool user_mapped = false;
    if (standalone) {
        launch_rebuild(m, st);                       // [T]:342-359
    } else {
        if (m.e_in_csc) {
            if (fold_user && ensure_perm_inverse(m, st)) user_mapped = true;
            else crc |= launch_permute(m, false, nullptr, st);   // (single GPU, or first use; multi-GPU sweeps do it at their end)
        }
        launch_stats(m, st, user_mapped);
    }
    crc |= launch_reduce_pair(m, st);                // sum e, sum e^2 (all ranks)
    if (timing) cudaEventRecord(m.ev_t[1], st);
    launch_global_hypers(m, st);                     // [T]:366-410
    launch_dim_hypers(m, st);                        // [T]:415-467
    if (m.cfg.hyper_mode == SBMF_HYPER_REF_T) launch_bias_hypers(m, st);   // [T]:469-511 ([S] has no biases)
    // peer-mapped replicas: the user phase writes U rows into every replica, so every rank must be done reading U first
    if (m.peer_ok) crc |= launch_barrier(m, st);
    if (timing) cudaEventRecord(m.ev_t[2], st);
    // G GPUs with peer pushes (option fuse_exchange): the user phase stores its final residual straight into the receive buffer of
    // the rank that owns the rating's item (PhaseArgs::xmap), overlapped with the phase; what is left between the phases is the
    // barrier (every peer's residuals and U rows have landed), and the item phase reads through recv_pos like a folded hand-over
    const bool fx = m.world > 1 && m.peer_ok && m.opt.fuse_exchange && m.xmap_fwd;
    launch_phase(m, m.us, m.it, true, fused, user_mapped ? m.it.e : nullptr, user_mapped ? m.perm_inv : nullptr, fx);   // [T]:514-558 (+ the fused residual refresh)
    if (timing) cudaEventRecord(m.ev_t[3], st);
    // residual CSR order -> CSC order; multi-GPU: all-to-all grouped with the all-gather of the updated U rows and user biases
    if (fx) crc |= launch_barrier(m, st);
    else if (!fold) crc |= launch_permute(m, true, &m.us, st);
    if (timing) cudaEventRecord(m.ev_t[4], st);
    launch_phase(m, m.it, m.us, false, false, fx ? m.recvbuf : (fold ? m.us.e : nullptr), fx ? m.recv_pos : (fold ? m.perm : nullptr));   // [T]:563-606
    m.e_in_csc = true;
    if (timing) cudaEventRecord(m.ev_t[5], st);
    if (m.world > 1) {
        // replicate the updated V rows / item biases; if the next sweep starts from the incremental residual, its CSC -> CSR
        // all-to-all rides in the same grouped launch
        const bool next_standalone = ((m.sweeps_done + 1) % m.cfg.rebuild_every) == 0 && m.cfg.residual_mode == 1;
        if (!next_standalone) {
            crc |= launch_permute(m, false, &m.it, st);
            m.e_in_csc = false;
        } else {
            crc |= launch_allgather_side(m, m.it, st);
        }
    }
    if (timing) cudaEventRecord(m.ev_c[0], st);
    launch_eval(m, st);                              // [T]:610-636
    crc |= launch_reduce_pair(m, st);
    launch_eval_final(m, st);
    if (timing) cudaEventRecord(m.ev_t[6], st);
    if (crc) return SBMF_ERR_NCCL;
    return SBMF_OK;
}

struct SweepKey {   // everything host-side that shapes the launch sequence of a sweep
    bool standalone, fused, next_standalone, e_in_csc;
    bool operator==(const SweepKey& o) const
    {
        return standalone == o.standalone && fused == o.fused && next_standalone == o.next_standalone && e_in_csc == o.e_in_csc;
    }
};

static SweepKey sweep_key(const Model& m)
{
    const bool due = (m.sweeps_done % m.cfg.rebuild_every) == 0;
    const bool standalone = m.need_rebuild || (due && (m.sweeps_done == 0 || m.cfg.residual_mode == 1));
    const bool next_standalone = ((m.sweeps_done + 1) % m.cfg.rebuild_every) == 0 && m.cfg.residual_mode == 1;
    return SweepKey{standalone, due && !standalone, next_standalone, m.e_in_csc};
}

// With per-phase timing off, a steady-state sweep (~100-250 dependent launches on five streams, plus the NCCL calls) is captured
// once into a CUDA graph and replayed: the launch sequence only depends on SweepKey, every per-sweep value (alpha, b_0, sweep
// counter, ...) lives in device memory.  Matters where a sweep is short: small matrices, or 1/8 of a big one per GPU.
static int graph_sweep(Model& m, bool& done)
{
    done = false;
    const bool disabled = !m.opt.graph;
    // the first two sweeps after init_factors / set_state run as direct launches: they build the lazily allocated inverse
    // maps (ensure_inverse), which must not happen inside a capture
    if (disabled || m.timing_enabled || m.sweeps_done < 2 || m.sweeps_since_init < 2 || m.need_rebuild) return SBMF_OK;
    const SweepKey key = sweep_key(m);
    if (m.graph_exec && !(key == SweepKey{m.gk[0], m.gk[1], m.gk[2], m.gk[3]})) {
        cudaGraphExecDestroy(m.graph_exec);
        m.graph_exec = nullptr;
    }
    if (!m.graph_exec) {
        if (m.graph_failed) return SBMF_OK;
        const uint64_t l0 = m.launches;
        cudaGraph_t g = nullptr;
        if (cudaStreamBeginCapture(m.s_main, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            m.graph_failed = true;
            return SBMF_OK;
        }
        const int rc = enqueue_sweep(m, false);
        const cudaError_t ce = cudaStreamEndCapture(m.s_main, &g);
        const bool periodic = m.e_in_csc == key.e_in_csc;
        if (rc != SBMF_OK || ce != cudaSuccess || !g || !periodic || cudaGraphInstantiate(&m.graph_exec, g, 0) != cudaSuccess) {
            cudaGetLastError();
            if (g) cudaGraphDestroy(g);
            m.graph_exec = nullptr;
            m.graph_failed = true;
            m.launches = l0;
            m.e_in_csc = key.e_in_csc;
            return rc != SBMF_OK ? rc : SBMF_OK;   // nothing was executed: the caller falls back to direct launches
        }
        cudaGraphDestroy(g);
        m.graph_launches = m.launches - l0;
        m.launches = l0;
        m.gk[0] = key.standalone; m.gk[1] = key.fused; m.gk[2] = key.next_standalone; m.gk[3] = key.e_in_csc;
    }
    if (cudaGraphLaunch(m.graph_exec, m.s_main) != cudaSuccess) {
        m.err = std::string("sweep: cudaGraphLaunch: ") + cudaGetErrorString(cudaGetLastError());
        return SBMF_ERR_CUDA;
    }
    m.launches += m.graph_launches;
    m.sweeps_done++;
    m.sweeps_since_init++;
    done = true;
    return SBMF_OK;
}

static int one_sweep(Model& m)
{
    bool done = false;
    int grc = graph_sweep(m, done);
    if (grc != SBMF_OK || done) return grc;
    const bool timing = m.timing_enabled;
    grc = enqueue_sweep(m, timing);
    if (grc != SBMF_OK) return grc;
    m.need_rebuild = false;
    m.sweeps_done++;
    m.sweeps_since_init++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        m.err = std::string("sweep: kernel launch failed: ") + cudaGetErrorString(e);
        return SBMF_ERR_CUDA;
    }
    if (timing) {
        // per-phase timing needs the events of this sweep: it serialises host and device once per sweep
        if ((e = cudaEventSynchronize(m.ev_t[6])) != cudaSuccess) {
            m.err = std::string("sweep: ") + cudaGetErrorString(e);
            return SBMF_ERR_CUDA;
        }
        float ms[6];
        for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&ms[i], m.ev_t[i], m.ev_t[i + 1]);
        float ag_pre = 0.f;
        cudaEventElapsedTime(&ag_pre, m.ev_t[5], m.ev_c[0]);
        m.timing.ms_rebuild += ms[0];
        m.timing.ms_hypers += ms[1];
        m.timing.ms_user_phase += ms[2];
        m.timing.ms_exchange += ms[3];
        m.timing.ms_item_phase += ms[4];
        m.timing.ms_eval += ms[5] - ag_pre;
        float tot = 0.f;
        cudaEventElapsedTime(&tot, m.ev_t[0], m.ev_t[6]);
        m.timing.ms_total += tot;
        m.timing.sweeps++;
        float ag = 0.f;   // multi-GPU: the grouped V all-gather (+ reverse residual all-to-all) between item phase and evaluation
        cudaEventElapsedTime(&ag, m.ev_t[5], m.ev_c[0]);
        m.timing.ms_allgather += ag;
        if (m.timing_detail && m.ev_top_used) {
            for (uint32_t i = 0; i + 1 < m.ev_top_used; i += 2) {
                float t = 0.f;
                cudaEventElapsedTime(&t, m.ev_top[i], m.ev_top[i + 1]);
                m.timing.ms_top_kernel += t;
                m.timing.top_kernel_launches++;
            }
            m.timing.top_kernel_ratings = m.it.nnz_heavy;
            m.ev_top_used = 0;
        }
    }
    return SBMF_OK;
}

int sbmf_cuda_sweep(sbmf_handle* h, uint32_t n_sweeps)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.have_train || !m.have_factors) {
        m.err = "sweep: call set_train and init_factors first";
        return SBMF_ERR_STATE;
    }
    API_CK(cudaSetDevice(m.device));   // before anything allocates: the caller may have switched devices (several handles per thread)
    if (!m.have_test) {
        int rc = build_test(m, 0, nullptr, nullptr, nullptr);
        if (rc != SBMF_OK) return rc;
    }
    if ((uint64_t)m.sweeps_done + n_sweeps > m.hist_cap) {   // the RMSE history grows with the chain (it used to stop at 4096 sweeps)
        uint64_t cap = m.hist_cap;
        while (cap < (uint64_t)m.sweeps_done + n_sweeps) cap *= 2;
        if (cap > 0xffffffffull) cap = 0xffffffffull;
        double* nh = nullptr;
        API_CK(cudaStreamSynchronize(m.s_main));
        API_CK(cudaMalloc((void**)&nh, (size_t)cap * 16));
        API_CK(cudaMemset(nh, 0, (size_t)cap * 16));
        API_CK(cudaMemcpy(nh, m.rmse_hist, (size_t)m.hist_cap * 16, cudaMemcpyDeviceToDevice));
        cudaFree(m.rmse_hist);
        m.rmse_hist = nh;
        m.hist_cap = (uint32_t)cap;
        if (m.graph_exec) {   // the captured sweep holds the old pointer
            cudaGraphExecDestroy(m.graph_exec);
            m.graph_exec = nullptr;
        }
    }
    API_CK(cudaEventRecord(m.ev_call[0], m.s_main));
    for (uint32_t s = 0; s < n_sweeps; ++s) {
        int rc = one_sweep(m);
        if (rc != SBMF_OK) return rc;
    }
    API_CK(cudaEventRecord(m.ev_call[1], m.s_main));
    return SBMF_OK;
}

int sbmf_cuda_last_sweep_call_ms(sbmf_handle* h, double* ms)
{
    if (!h || !ms) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (m.sweeps_done == 0) {
        m.err = "last_sweep_call_ms: no sweep has run";
        return SBMF_ERR_STATE;
    }
    API_CK(cudaSetDevice(m.device));
    API_CK(cudaEventSynchronize(m.ev_call[1]));
    float t = 0.f;
    API_CK(cudaEventElapsedTime(&t, m.ev_call[0], m.ev_call[1]));
    *ms = (double)t;
    return SBMF_OK;
}

int sbmf_cuda_host_alloc(void** ptr, size_t bytes)
{
    if (!ptr) return SBMF_ERR_INVALID;
    cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        g_create_err = std::string("host_alloc: ") + cudaGetErrorString(e);
        return SBMF_ERR_NOMEM;
    }
    return SBMF_OK;
}

int sbmf_cuda_host_free(void* ptr)
{
    if (ptr) cudaFreeHost(ptr);
    return SBMF_OK;
}

int sbmf_cuda_synchronize(sbmf_handle* h)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    API_CK(cudaSetDevice(m.device));
    API_CK(cudaStreamSynchronize(m.s_main));
    API_CK(cudaStreamSynchronize(m.s_aux));
    API_CK(cudaStreamSynchronize(m.s_aux2));
    return SBMF_OK;
}

int sbmf_cuda_eval(sbmf_handle* h, double* rmse_running_mean, double* rmse_last_sweep)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (m.sweeps_done == 0) {
        m.err = "eval: no sweep has run";
        return SBMF_ERR_STATE;
    }
    API_CK(cudaSetDevice(m.device));
    double v[2];
    API_CK(cudaMemcpyAsync(v, &m.sc->rmse_mean, 16, cudaMemcpyDeviceToHost, m.s_main));
    API_CK(cudaStreamSynchronize(m.s_main));
    if (rmse_running_mean) *rmse_running_mean = v[0];
    if (rmse_last_sweep) *rmse_last_sweep = v[1];
    return SBMF_OK;
}

int sbmf_cuda_get_rmse_history(sbmf_handle* h, uint32_t first, uint32_t count, double* rmse_running_mean, double* rmse_sweep)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if ((uint64_t)first + count > m.sweeps_done || (uint64_t)first + count > m.hist_cap) {
        m.err = "get_rmse_history: range exceeds the sweeps run";
        return SBMF_ERR_INVALID;
    }
    API_CK(cudaSetDevice(m.device));
    std::vector<double> tmp((size_t)count * 2 + 2);
    API_CK(cudaMemcpyAsync(tmp.data(), m.rmse_hist + (size_t)first * 2, (size_t)count * 16, cudaMemcpyDeviceToHost, m.s_main));
    API_CK(cudaStreamSynchronize(m.s_main));
    for (uint32_t i = 0; i < count; ++i) {
        if (rmse_running_mean) rmse_running_mean[i] = tmp[(size_t)i * 2];
        if (rmse_sweep) rmse_sweep[i] = tmp[(size_t)i * 2 + 1];
    }
    return SBMF_OK;
}

int sbmf_cuda_get_pred(sbmf_handle* h, float* pred)
{
    if (!h || !pred) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.have_test || m.sweeps_done <= m.cfg.burn_in) {
        m.err = "get_pred: no test set or no collected sweep";
        return SBMF_ERR_STATE;
    }
    API_CK(cudaSetDevice(m.device));
    if (m.world > 1) {   // every rank accumulated its own slice of the test set
        std::vector<size_t> off(m.world), cnt(m.world);
        for (int q = 0; q < m.world; ++q) {
            off[q] = m.Nt * (uint64_t)q / (uint64_t)m.world;
            cnt[q] = m.Nt * (uint64_t)(q + 1) / (uint64_t)m.world - off[q];
        }
        if (comm_allgatherv_f64(m.comm, m.t_sum, off.data(), cnt.data(), m.s_main, m.err) != 0) return SBMF_ERR_NCCL;
    }
    const double denom = (double)(m.sweeps_done - m.cfg.burn_in);
    float* d_out = nullptr;
    API_CK(cudaMallocAsync((void**)&d_out, (m.Nt ? m.Nt : 1) * 4, m.s_main));
    launch_pred_mean(m, d_out, denom, m.s_main);
    API_CK(cudaMemcpyAsync(pred, d_out, m.Nt * 4, cudaMemcpyDeviceToHost, m.s_main));
    API_CK(cudaFreeAsync(d_out, m.s_main));
    API_CK(cudaStreamSynchronize(m.s_main));
    return SBMF_OK;
}

int sbmf_cuda_get_state(sbmf_handle* h, sbmf_state* out)
{
    if (!h || !out) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.have_factors) {
        m.err = "get_state: call init_factors first";
        return SBMF_ERR_STATE;
    }
    API_CK(cudaSetDevice(m.device));
    cudaStream_t st = m.s_main;
    API_CK(cudaStreamSynchronize(st));
    if (out->U || out->V) {
        DeviceTmp tmp;
        float*& d_tmp = tmp.p;
        const size_t nu = (size_t)m.I * m.K, nv = (size_t)m.K * m.J;
        API_CK(cudaMalloc((void**)&d_tmp, std::max(nu, nv) * 4));
        if (out->U) {
            launch_export_factors(m, m.us, d_tmp, false, st);
            API_CK(cudaMemcpyAsync(out->U, d_tmp, nu * 4, cudaMemcpyDeviceToHost, st));
            API_CK(cudaStreamSynchronize(st));
        }
        if (out->V) {
            launch_export_factors(m, m.it, d_tmp, true, st);
            API_CK(cudaMemcpyAsync(out->V, d_tmp, nv * 4, cudaMemcpyDeviceToHost, st));
            API_CK(cudaStreamSynchronize(st));
        }
    }
    {   // per-row arrays: stored by position, returned by the caller's row id
        std::vector<float> byp;
        auto row_out = [&](float* dst, const float* d_src, const Side& sd) -> cudaError_t {
            if (!dst) return cudaSuccess;
            if (sd.h_id_at.empty()) return cudaMemcpy(dst, d_src, (size_t)sd.n * 4, cudaMemcpyDeviceToHost);
            byp.resize(sd.n);
            const cudaError_t e = cudaMemcpy(byp.data(), d_src, (size_t)sd.n * 4, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess)
                for (uint32_t p = 0; p < sd.n; ++p) dst[sd.h_id_at[p]] = byp[p];
            return e;
        };
        API_CK(row_out(out->b_i, m.us.bias, m.us));
        API_CK(row_out(out->b_j, m.it.bias, m.it));
        API_CK(row_out(out->mu_b_i, m.us.mu_b, m.us));
        API_CK(row_out(out->sigma_b_i, m.us.sigma_b, m.us));
        API_CK(row_out(out->mu_b_j, m.it.mu_b, m.it));
        API_CK(row_out(out->sigma_b_j, m.it.sigma_b, m.it));
    }
    if (out->sigma_u) API_CK(cudaMemcpy(out->sigma_u, m.us.sigma_k, (size_t)m.K * 8, cudaMemcpyDeviceToHost));
    if (out->mu_u) API_CK(cudaMemcpy(out->mu_u, m.us.mu_k, (size_t)m.K * 8, cudaMemcpyDeviceToHost));
    if (out->sigma_v) API_CK(cudaMemcpy(out->sigma_v, m.it.sigma_k, (size_t)m.K * 8, cudaMemcpyDeviceToHost));
    if (out->mu_v) API_CK(cudaMemcpy(out->mu_v, m.it.mu_k, (size_t)m.K * 8, cudaMemcpyDeviceToHost));
    if (out->E) {
        // residual in rating (file) order, from whichever layout holds the freshest copy
        // (multi-GPU: only the entries of this rank's shard are written)
        const uint64_t nl = m.e_in_csc ? m.n_csc : m.n_csr;
        std::vector<float> e(nl ? nl : 1);
        std::vector<uint32_t> id(nl ? nl : 1);
        API_CK(cudaMemcpy(e.data(), m.e_in_csc ? m.it.e : m.us.e, nl * 4, cudaMemcpyDeviceToHost));
        API_CK(cudaMemcpy(id.data(), m.e_in_csc ? m.csc_id : m.csr_id, nl * 4, cudaMemcpyDeviceToHost));
        for (uint64_t s = 0; s < nl; ++s) out->E[id[s]] = e[s];
    }
    Scalars sc;
    API_CK(cudaMemcpy(&sc, m.sc, sizeof(Scalars), cudaMemcpyDeviceToHost));
    out->b_0 = sc.b_0;
    out->alpha = sc.alpha;
    out->mu_b_0 = sc.mu_b_0;
    out->sigma_b_0 = sc.sigma_b_0;
    out->sum_e = sc.sum_e;
    out->sum_e2 = sc.sum_e2;
    out->sweeps_done = m.sweeps_done;
    return SBMF_OK;
}

// Restore of the complete sampler state at a sweep boundary (checkpoint / resume; the reference has none).
// Built on init_factors (factor upload + the zero state of [T]:268-281, 315-318), then every array the caller provides
// overwrites its zero.  Per-sweep scratch (shift_f, pacc, heavy-row partials, post_var) is recomputed by the next sweep.
int sbmf_cuda_set_state(sbmf_handle* h, const sbmf_state* in)
{
    if (!h || !in) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.have_train) {
        m.err = "set_state: call set_train first";
        return SBMF_ERR_STATE;
    }
    if (!in->U || !in->V) {
        m.err = "set_state: U and V are required";
        return SBMF_ERR_INVALID;
    }
    const int rc = sbmf_cuda_init_factors(h, in->U, in->V);
    if (rc != SBMF_OK) return rc;
    cudaStream_t st = m.s_main;
    struct Up {
        float* dst;
        const float* src;
        const Side* sd;
    };
    const Up ups[] = {{m.us.bias, in->b_i, &m.us},    {m.it.bias, in->b_j, &m.it},        {m.us.mu_b, in->mu_b_i, &m.us},
                      {m.us.sigma_b, in->sigma_b_i, &m.us}, {m.it.mu_b, in->mu_b_j, &m.it}, {m.it.sigma_b, in->sigma_b_j, &m.it}};
    std::vector<float> byp[6];   // per-row arrays arrive by the caller's row id and are stored by position (alive until the sync below)
    for (int q = 0; q < 6; ++q) {
        const Up& u = ups[q];
        if (!u.src) continue;
        const float* src = u.src;
        if (!u.sd->h_id_at.empty()) {
            byp[q].resize(u.sd->n);
            for (uint32_t p = 0; p < u.sd->n; ++p) byp[q][p] = u.src[u.sd->h_id_at[p]];
            src = byp[q].data();
        }
        API_CK(cudaMemcpyAsync(u.dst, src, (size_t)u.sd->n * 4, cudaMemcpyHostToDevice, st));
    }
    // per-dimension hyper-parameters: fp64 masters (the next sweep's Gamma step needs the OLD means, [T]:424, 450) and the fp32
    // mirrors the row kernels read; padding dimensions K..KP-1 get the values dim_hyper_final_kernel gives them (1, 0)
    std::vector<double> hd[4];
    std::vector<float> hf[4];
    {
        const double* src[4] = {in->sigma_u, in->mu_u, in->sigma_v, in->mu_v};
        double* dstd[4] = {m.us.sigma_k, m.us.mu_k, m.it.sigma_k, m.it.mu_k};
        float* dstf[4] = {m.us.sigma_kf, m.us.mu_kf, m.it.sigma_kf, m.it.mu_kf};
        for (int a = 0; a < 4; ++a) {
            if (!src[a]) continue;
            const double padv = (a % 2 == 0) ? 1.0 : 0.0;
            hd[a].assign(m.KP, padv);
            hf[a].assign(m.KP, (float)padv);
            for (uint32_t k = 0; k < m.K; ++k) {
                hd[a][k] = src[a][k];
                hf[a][k] = (float)src[a][k];
            }
            API_CK(cudaMemcpyAsync(dstd[a], hd[a].data(), (size_t)m.KP * 8, cudaMemcpyHostToDevice, st));
            API_CK(cudaMemcpyAsync(dstf[a], hf[a].data(), (size_t)m.KP * 4, cudaMemcpyHostToDevice, st));
        }
    }
    Scalars sc;
    memset(&sc, 0, sizeof(sc));
    sc.b_0 = in->b_0;
    sc.alpha = in->alpha;
    sc.mu_b_0 = in->mu_b_0;
    sc.sigma_b_0 = in->sigma_b_0;
    sc.sum_e = in->sum_e;
    sc.sum_e2 = in->sum_e2;
    sc.alpha_f = (float)in->alpha;
    sc.b_0_f = (float)in->b_0;
    sc.sweep = in->sweeps_done;
    API_CK(cudaMemcpyAsync(m.sc, &sc, sizeof(Scalars), cudaMemcpyHostToDevice, st));
    std::vector<float> e;
    if (in->E) {
        // the residual goes back where the uninterrupted chain keeps it at a sweep boundary: the item side's (CSC) order on one
        // GPU, the user side's (CSR) order on several (enqueue_sweep), so the next sweep's statistics sum it in the same order
        const bool to_csc = (m.world == 1);
        const uint64_t nl = to_csc ? m.n_csc : m.n_csr;
        std::vector<uint32_t> id(nl ? nl : 1);
        e.resize(nl ? nl : 1);
        API_CK(cudaMemcpyAsync(id.data(), to_csc ? m.csc_id : m.csr_id, nl * 4, cudaMemcpyDeviceToHost, st));
        API_CK(cudaStreamSynchronize(st));
        for (uint64_t s = 0; s < nl; ++s) e[s] = in->E[id[s]];
        API_CK(cudaMemcpyAsync(to_csc ? m.it.e : m.us.e, e.data(), nl * 4, cudaMemcpyHostToDevice, st));
        m.e_in_csc = to_csc;
    }
    API_CK(cudaStreamSynchronize(st));
    m.need_rebuild = !in->E && in->sweeps_done > 0;
    m.sweeps_done = in->sweeps_done;
    return SBMF_OK;
}

int sbmf_cuda_get_pred_sum(sbmf_handle* h, double* sum)
{
    if (!h || !sum) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.have_test) {
        m.err = "get_pred_sum: no test set";
        return SBMF_ERR_STATE;
    }
    API_CK(cudaSetDevice(m.device));
    if (m.world > 1) {   // every rank accumulated its own slice of the test set
        std::vector<size_t> off(m.world), cnt(m.world);
        for (int q = 0; q < m.world; ++q) {
            off[q] = m.Nt * (uint64_t)q / (uint64_t)m.world;
            cnt[q] = m.Nt * (uint64_t)(q + 1) / (uint64_t)m.world - off[q];
        }
        if (comm_allgatherv_f64(m.comm, m.t_sum, off.data(), cnt.data(), m.s_main, m.err) != 0) return SBMF_ERR_NCCL;
    }
    API_CK(cudaMemcpyAsync(sum, m.t_sum, m.Nt * 8, cudaMemcpyDeviceToHost, m.s_main));
    API_CK(cudaStreamSynchronize(m.s_main));
    return SBMF_OK;
}

int sbmf_cuda_set_pred_sum(sbmf_handle* h, const double* sum)
{
    if (!h || !sum) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.have_test) {
        m.err = "set_pred_sum: no test set";
        return SBMF_ERR_STATE;
    }
    API_CK(cudaSetDevice(m.device));
    API_CK(cudaMemcpyAsync(m.t_sum, sum, m.Nt * 8, cudaMemcpyHostToDevice, m.s_main));
    API_CK(cudaStreamSynchronize(m.s_main));
    return SBMF_OK;
}

int sbmf_cuda_get_timing(sbmf_handle* h, sbmf_timing* out)
{
    if (!h || !out) return SBMF_ERR_INVALID;
    Model& m = h->m;
    *out = m.timing;
    out->kernel_launches = m.launches;
    out->nnz_light_user = m.us.nnz_resident;
    out->nnz_heavy_user = m.us.nnz_heavy;
    out->nnz_light_item = m.it.nnz_resident;
    out->nnz_heavy_item = m.it.nnz_heavy;
    return SBMF_OK;
}

int sbmf_cuda_reset_timing(sbmf_handle* h)
{
    if (!h) return SBMF_ERR_INVALID;
    h->m.timing = sbmf_timing{};
    h->m.launches = 0;
    return SBMF_OK;
}

int sbmf_cuda_set_timing_enabled(sbmf_handle* h, int enabled)
{
    if (!h) return SBMF_ERR_INVALID;
    h->m.timing_enabled = enabled != 0;
    h->m.timing_detail = enabled == 2;
    return SBMF_OK;
}

}  // extern "C"
