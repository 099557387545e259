// fm.cu -- general FM Gibbs sampler on B200: libFM's `-method mcmc` / `-method als` for regression on an
// arbitrary sparse design matrix, behind the C ABI of include/sbmf_fm_cuda.h.  sm_100a only, no CPU fallback.
//
// Reference: src/libfm/src/fm_learn_mcmc.h ("[G]"), fm_learn_mcmc_simultaneous.h ("[GS]"), libfm.cpp ("[L]").
//
// libFM draws the coordinates of one factor strictly one attribute after the other ([G]:552-565); each draw reads and updates
// the per-case caches e (prediction error) and q (sum_j v_jf x_j).  Two attributes that never occur in the same case touch
// disjoint cache entries, so their draws commute.  set_train therefore cuts the attribute sequence 0..p-1 into maximal RUNS of
// consecutive attributes that share no case (for one-hot blocks -- users, items, ... -- a run is the whole block; a dense
// real-valued attribute is a run of its own) and the sampler processes run after run, all columns of a run concurrently:
// exactly libFM's scan, not an approximation of it.  Within a run a column is owned by a warp (<= 512 entries), a CTA
// (<= 16384) or cut into 8192-entry slices (reduce -> draw -> apply), all with fixed summation trees (run-to-run reproducible).
//
// HBM layout: the design matrix twice -- column form (col_ptr, case_id, x: what the draws stream) and row form with ascending
// attribute ids (row_ptr, attr, x: q rebuild and the full re-prediction of [GS]:134) -- both built on the device from the
// caller's row form by two stable radix sorts; e, q fp32 per case; w [p] and V [p][K] (attribute-major: one attribute's K
// factors are contiguous for the prediction pass) fp32; hyper-parameters fp64.  Every kernel here is a streaming / gather pass:
// 16 B (w) or 24 B (v) of algorithmic traffic per design-matrix entry and draw.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "../../include/sbmf_fm_cuda.h"
#include "common.cuh"
#include "fm_math.cuh"

namespace sbmf_fm {

using sbmf::draw_gamma_f64;
using sbmf::normal_f32;
using sbmf::normal_f64;
using sbmf::philox_site;

// Philox stream ids of this sampler (counter word 2); disjoint from sbmf::Site
enum FmSite : uint32_t {
    SITE_FM_INIT_V = 32, SITE_FM_INIT_W = 33, SITE_FM_ALPHA = 34, SITE_FM_W0 = 35, SITE_FM_W_LAMBDA = 36, SITE_FM_W_MU = 37,
    SITE_FM_W = 38, SITE_FM_V_LAMBDA = 39, SITE_FM_V_MU = 40, SITE_FM_V = 41
};

constexpr double ALPHA_0 = 1.0, GAMMA_0 = 1.0, BETA_0 = 1.0, MU_0 = 0.0, W0_MEAN_0 = 0.0;   // [G]:1099-1106
// tier limits; the CPU execution of these kernels (tools/simt_emu.h, tests/test_fm_simt_emulation.py) shrinks them with -D so that
// the small fixtures reach every tier
#ifndef FM_WARP_COL_MAX
#define FM_WARP_COL_MAX 512
#define FM_BLOCK_COL_MAX 16384
#define FM_SLICE_LEN 8192
#define FM_HYPER_CHUNK 4096
#endif
#ifndef FM_BLOCK_T
#define FM_BLOCK_T 256
#endif
constexpr uint32_t WARP_COL_MAX = FM_WARP_COL_MAX, BLOCK_COL_MAX = FM_BLOCK_COL_MAX, SLICE_LEN = FM_SLICE_LEN, HYPER_CHUNK = FM_HYPER_CHUNK,
                   HIST_CAP = 1u << 16;
constexpr int BLOCK_T = FM_BLOCK_T;   // threads per CTA of the streaming kernels (a multiple of 32)

// kernel launch: <<<grid, block, 0, stream>>> on the GPU; under SBMF_SIMT_EMU (a test-only host build of this file against
// tools/emu_include) the same kernel body is executed block by block on host threads
#ifdef SBMF_SIMT_EMU
#define FM_LAUNCH(kernel, grid, block, stream, ...) simt::launch((grid), (block), [&] { kernel(__VA_ARGS__); })
#else
#define FM_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif

struct Scal {
    double w0, alpha, w0_delta;
    double min_target, max_target;
    uint32_t iter, pad;
};

struct Slice {          // <= SLICE_LEN consecutive entries of one long column
    int64_t begin;
    uint32_t len, gi;   // gi: index of the column in its run's long-column list
};
struct LongCol {
    uint32_t col, slice_begin, slice_end, pad;
};
struct Run {            // offsets into the flattened work lists
    uint32_t w_off, w_cnt, c_off, c_cnt, g_off, g_cnt, s_off, s_cnt;
};

struct RowForm {
    uint32_t n = 0;
    int64_t nnz = 0;
    int64_t* row_ptr = nullptr;
    uint32_t* attr = nullptr;
    float* x = nullptr;
    float* y = nullptr;
};

struct Model {
    sbmf_fm_config cfg;
    std::string err;
    cudaStream_t st = nullptr;
    // the three column tiers of a run (warp / CTA / sliced columns) touch disjoint columns AND disjoint cases -- a run's columns share
    // no case by construction -- so they execute concurrently: fork from st, join before the next run
    cudaStream_t st_tier[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    int sm_count = 148;
    uint32_t p = 0, K = 0, G = 1;
    bool have_train = false, have_test = false, inited = false;
    std::vector<uint32_t> group_h;
    // design matrices
    RowForm tr, te;
    int64_t* col_ptr = nullptr;
    uint32_t* case_id = nullptr;
    float* xc = nullptr;
    // work lists
    std::vector<Run> runs;
    std::vector<uint32_t> run_begin;
    uint32_t *wl_cols = nullptr;
    Slice* wl_slices = nullptr;
    LongCol* wl_long = nullptr;
    double2* slice_part = nullptr;
    float2* long_scratch = nullptr;       // (theta_old, delta) per long column of the current run
    // group-sorted attribute list for the hyper sums
    uint32_t* gs_attr = nullptr;
    uint32_t* chunk_begin = nullptr;      // [nchunks + 1] positions in gs_attr
    uint32_t* gchunk_ptr = nullptr;       // [G + 1] chunk range of every group
    uint32_t* n_per_group = nullptr;      // [G]
    uint32_t nchunks = 0;
    double* hyper_part = nullptr;         // [nchunks][F][2]
    // state
    uint32_t* group = nullptr;
    float *w = nullptr, *V = nullptr, *e = nullptr, *q = nullptr, *pred_this = nullptr;
    double *w_mu = nullptr, *w_lambda = nullptr, *v_mu = nullptr, *v_lambda = nullptr, *pred_sum = nullptr;
    Scal* sc = nullptr;
    double* red_part = nullptr;           // [3][red_blocks][2]: stats, train error, test error
    uint32_t red_blocks = 0;
    double* hist = nullptr;               // [HIST_CAP][2]
    uint32_t iters_done = 0;
    uint64_t launches = 0;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            m.err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            return (e_ == cudaErrorMemoryAllocation) ? SBMF_ERR_NOMEM : SBMF_ERR_CUDA;             \
        }                                                                                          \
    } while (0)

template <typename T>
static cudaError_t dmalloc(T** p, size_t n)
{
    return cudaMalloc((void**)p, (n ? n : 1) * sizeof(T));
}
template <typename T>
static void dfree(T*& p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

// ---------------------------------------------------------------------------------------------------------------------
// reductions with a fixed tree
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// every thread of a BLOCK_T-thread CTA receives the same total
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* [BLOCK_T / 32] */)
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
    __syncthreads();
    T t = smem[0];
#pragma unroll
    for (int i = 1; i < BLOCK_T / 32; ++i) t += smem[i];
    return t;
}

// ---------------------------------------------------------------------------------------------------------------------
// set-up kernels
__global__ void iota_kernel(uint32_t* v, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = (uint32_t)i;
}
// row of every entry of a row-form matrix: the r with row_ptr[r] <= k < row_ptr[r + 1]
__global__ void entry_row_kernel(const int64_t* row_ptr, uint32_t n, int64_t nnz, uint32_t* entry_row)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = n;             // first r with row_ptr[r + 1] > k
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (row_ptr[mid + 1] <= k) lo = mid + 1;
            else hi = mid;
        }
        entry_row[k] = lo;
    }
}
// ptr[r] = first position whose (sorted) key is >= r, r = 0..nkeys
__global__ void seg_ptr_kernel(const uint32_t* sorted_keys, int64_t n, uint32_t nkeys, int64_t* ptr)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nkeys) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < r) lo = mid + 1;
        else hi = mid;
    }
    ptr[r] = lo;
}
__global__ void gather_u32_kernel(const uint32_t* src, const uint32_t* id, uint32_t* dst, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[id[i]];
}
__global__ void gather_f32_kernel(const float* src, const uint32_t* id, float* dst, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[id[i]];
}
// rows hold ascending attribute ids: adjacent entries of a row give next_attr (smallest co-occurring larger id) and expose
// an attribute listed twice in one case
__global__ void next_attr_kernel(const uint32_t* attr, const uint32_t* entry_row, int64_t nnz, uint32_t* next_attr, uint32_t* dup_flag)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k + 1 < nnz; k += (int64_t)gridDim.x * blockDim.x) {
        if (entry_row[k] != entry_row[k + 1]) continue;
        if (attr[k] == attr[k + 1]) atomicExch(dup_flag, 1u);
        else atomicMin(&next_attr[attr[k]], attr[k + 1]);
    }
}
__global__ void fill_u32_kernel(uint32_t* v, uint32_t n, uint32_t val)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = val;
}

__global__ void init_params_kernel(float* w, float* V, uint32_t p, uint32_t K, uint64_t seed, float init_stdev, int draw_w, int draw_v)
{
    const int64_t total = (int64_t)p * (K + 1);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(i / (K + 1)), f = (uint32_t)(i % (K + 1));
        if (f == K) {
            if (draw_w) w[j] = init_stdev * normal_f32(philox_site(seed, SITE_FM_INIT_W, j, 0, 0));     // [L]:412
        } else if (draw_v) {
            V[(size_t)j * K + f] = init_stdev * normal_f32(philox_site(seed, SITE_FM_INIT_V, j, f, 0));  // fm_model.h:96
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// coordinate draws
struct ColArgs {
    const int64_t* col_ptr;
    const uint32_t* case_id;
    const float* xc;
    float *e, *q;
    float* theta;               // w (stride 1) or V (stride K)
    uint32_t stride, f;
    const uint32_t* group;
    const double *mu, *lambda;  // [g * hstride + f]
    uint32_t hstride;
    const Scal* sc;
    uint64_t seed;
    uint32_t site;
    int live;                   // draw noise (do_sample && !zero-noise)
};

template <int COORD>
__device__ __forceinline__ float draw_theta(const ColArgs& a, uint32_t j, double hh, double he, float theta_old)
{
    const uint32_t g = a.group[j];
    const Posterior post = posterior<COORD>(hh, he, (double)theta_old, a.sc->alpha, a.mu[(size_t)g * a.hstride + a.f], a.lambda[(size_t)g * a.hstride + a.f]);
    const double z = a.live ? (double)normal_f32(philox_site(a.seed, a.site, j, a.f, a.sc->iter)) : 0.0;
    return (float)settle(post, z, (double)theta_old);
}

// The two passes over a column's entries [k0, en) with stride STEP (lanes of a warp / threads of a CTA), UNR entries per thread in
// flight: the entry's case id, then e / q of that case, are dependent loads (two DRAM round trips), and with one entry at a time a
// column pass ran at a few percent of the memory system (ncu, col_block_kernel: 103 long-scoreboard stalls per issue, DRAM 10 %).
// A thread still visits its entries in ascending order, so every sum is formed exactly as before.
constexpr int COL_UNR = 4;
template <int COORD, int STEP>
__device__ __forceinline__ void col_accumulate(const ColArgs& a, int64_t k0, int64_t en, float theta_old, ColSums& s)
{
    for (int64_t k = k0; k < en; k += (int64_t)STEP * COL_UNR) {
        uint32_t c[COL_UNR];
        float x[COL_UNR], ev[COL_UNR], qv[COL_UNR];
#pragma unroll
        for (int u = 0; u < COL_UNR; ++u) {
            const int64_t kk = k + (int64_t)u * STEP;
            c[u] = kk < en ? a.case_id[kk] : 0u;
            x[u] = kk < en ? a.xc[kk] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < COL_UNR; ++u) {
            const bool ok = k + (int64_t)u * STEP < en;
            ev[u] = ok ? a.e[c[u]] : 0.f;
            qv[u] = (ok && COORD == COORD_V) ? a.q[c[u]] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < COL_UNR; ++u)
            if (k + (int64_t)u * STEP < en) accumulate_entry<COORD>(s, x[u], ev[u], qv[u], theta_old);
    }
}
template <int COORD, int STEP>
__device__ __forceinline__ void col_apply(const ColArgs& a, int64_t k0, int64_t en, float theta_old, float delta)
{
    for (int64_t k = k0; k < en; k += (int64_t)STEP * COL_UNR) {
        uint32_t c[COL_UNR];
        float x[COL_UNR], ev[COL_UNR], qv[COL_UNR];
#pragma unroll
        for (int u = 0; u < COL_UNR; ++u) {
            const int64_t kk = k + (int64_t)u * STEP;
            c[u] = kk < en ? a.case_id[kk] : 0u;
            x[u] = kk < en ? a.xc[kk] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < COL_UNR; ++u) {
            const bool ok = k + (int64_t)u * STEP < en;
            ev[u] = ok ? a.e[c[u]] : 0.f;
            qv[u] = (ok && COORD == COORD_V) ? a.q[c[u]] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < COL_UNR; ++u) {
            if (k + (int64_t)u * STEP >= en) continue;
            apply_entry<COORD>(x[u], ev[u], qv[u], theta_old, delta);
            a.e[c[u]] = ev[u];
            if (COORD == COORD_V) a.q[c[u]] = qv[u];
        }
    }
}

// one warp per column of <= WARP_COL_MAX entries (also the empty ones: a draw from the prior, [G]:458-466)
template <int COORD>
__global__ void __launch_bounds__(BLOCK_T) col_warp_kernel(ColArgs a, const uint32_t* __restrict__ cols, uint32_t ncols)
{
    const uint32_t gw = (blockIdx.x * BLOCK_T + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= ncols) return;
    const uint32_t j = cols[gw];
    const int64_t b = a.col_ptr[j], en = a.col_ptr[j + 1];
    float* th = a.theta + (size_t)j * a.stride + a.f;
    const float theta_old = *th;
    ColSums s{0.0, 0.0};
    col_accumulate<COORD, 32>(a, b + lane, en, theta_old, s);
    const double hh = warp_sum(s.hh), he = warp_sum(s.he);     // identical on every lane
    const float theta_new = draw_theta<COORD>(a, j, hh, he, theta_old);
    const float delta = theta_old - theta_new;
    if (lane == 0) *th = theta_new;
    if (delta == 0.f) return;
    col_apply<COORD, 32>(a, b + lane, en, theta_old, delta);
}

// one CTA per column of <= BLOCK_COL_MAX entries
template <int COORD>
__global__ void __launch_bounds__(BLOCK_T) col_block_kernel(ColArgs a, const uint32_t* __restrict__ cols)
{
    __shared__ double sm[BLOCK_T / 32];
    const uint32_t j = cols[blockIdx.x];
    const int64_t b = a.col_ptr[j], en = a.col_ptr[j + 1];
    float* th = a.theta + (size_t)j * a.stride + a.f;
    const float theta_old = *th;
    ColSums s{0.0, 0.0};
    col_accumulate<COORD, BLOCK_T>(a, b + threadIdx.x, en, theta_old, s);
    const double hh = block_sum(s.hh, sm);
    const double he = block_sum(s.he, sm);                    // (block_sum's leading barrier also orders the theta_old reads)
    const float theta_new = draw_theta<COORD>(a, j, hh, he, theta_old);
    const float delta = theta_old - theta_new;
    if (threadIdx.x == 0) *th = theta_new;
    if (delta == 0.f) return;
    col_apply<COORD, BLOCK_T>(a, b + threadIdx.x, en, theta_old, delta);
}

// long columns: per-slice partial sums, one draw per column (slices combined in slice order), per-slice apply
template <int COORD>
__global__ void __launch_bounds__(BLOCK_T) slice_reduce_kernel(ColArgs a, const Slice* __restrict__ slices, const LongCol* __restrict__ lcols,
                                                                 double2* __restrict__ part)
{
    __shared__ double sm[BLOCK_T / 32];
    const Slice sl = slices[blockIdx.x];
    const uint32_t j = lcols[sl.gi].col;
    const float theta_old = a.theta[(size_t)j * a.stride + a.f];
    ColSums s{0.0, 0.0};
    col_accumulate<COORD, BLOCK_T>(a, (int64_t)sl.begin + threadIdx.x, (int64_t)sl.begin + sl.len, theta_old, s);
    const double hh = block_sum(s.hh, sm);
    const double he = block_sum(s.he, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = make_double2(hh, he);
}
template <int COORD>
__global__ void slice_draw_kernel(ColArgs a, const LongCol* __restrict__ lcols, uint32_t nlong, uint32_t slice_base, const double2* __restrict__ part,
                                  float2* __restrict__ scratch)
{
    const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= nlong) return;
    const LongCol lc = lcols[gi];
    double hh = 0.0, he = 0.0;
    for (uint32_t s = lc.slice_begin; s < lc.slice_end; ++s) {
        const double2 v = part[s - slice_base];
        hh += v.x;
        he += v.y;
    }
    float* th = a.theta + (size_t)lc.col * a.stride + a.f;
    const float theta_old = *th;
    const float theta_new = draw_theta<COORD>(a, lc.col, hh, he, theta_old);
    *th = theta_new;
    scratch[gi] = make_float2(theta_old, theta_old - theta_new);
}
template <int COORD>
__global__ void __launch_bounds__(BLOCK_T) slice_apply_kernel(ColArgs a, const Slice* __restrict__ slices, const float2* __restrict__ scratch)
{
    const Slice sl = slices[blockIdx.x];
    const float2 od = scratch[sl.gi];
    if (od.y == 0.f) return;
    col_apply<COORD, BLOCK_T>(a, (int64_t)sl.begin + threadIdx.x, (int64_t)sl.begin + sl.len, od.x, od.y);
}

// ---------------------------------------------------------------------------------------------------------------------
// row passes: q rebuild ([G]:384-409) and the full prediction ([G]:117-349, fm_model.h:104-129)
template <int LPR>   // lanes per row
__global__ void __launch_bounds__(BLOCK_T) q_rebuild_kernel(const int64_t* __restrict__ row_ptr, const uint32_t* __restrict__ attr,
                                                             const float* __restrict__ x, uint32_t n, const float* __restrict__ V, uint32_t K,
                                                             uint32_t f, float* __restrict__ q)
{
    const uint32_t gid = blockIdx.x * BLOCK_T + threadIdx.x;
    const uint32_t row = gid / LPR, lg = gid % LPR;
    float s = 0.f;
    if (row < n)
        for (int64_t k = row_ptr[row] + lg; k < row_ptr[row + 1]; k += LPR) s = fmaf(V[(size_t)attr[k] * K + f], x[k], s);
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (row < n && lg == 0) q[row] = s;
}

// G lanes per case, lane lg owns the factors f = lg, lg + G, ...; all lanes of a group walk the case's entries together, so the
// K factors of an attribute are read as one contiguous segment.  TRAIN: e = prediction - y and the clamped squared error
// ([GS]:161-170).  TEST: pred_this, pred_sum += clamp(prediction), squared error of the running mean ([GS]:150-158, 307-326).
template <int G, bool TRAIN>
__global__ void __launch_bounds__(BLOCK_T) predict_kernel(const int64_t* __restrict__ row_ptr, const uint32_t* __restrict__ attr,
                                                           const float* __restrict__ x, const float* __restrict__ y, uint32_t n,
                                                           const float* __restrict__ w, const float* __restrict__ V, uint32_t K, int k0, int k1,
                                                           const Scal* __restrict__ sc, float* __restrict__ e, float* __restrict__ pred_this,
                                                           double* __restrict__ pred_sum, int accumulate, double* __restrict__ part)
{
    constexpr int NF = (G == 32) ? 8 : 2;                      // K <= 256 (G = 32) or K <= 16 (G = 8)
    __shared__ double sm[BLOCK_T / 32];
    const uint32_t gid = blockIdx.x * BLOCK_T + threadIdx.x;
    const uint32_t row = gid / G, lg = gid % G;
    float s[NF], ss[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) s[i] = ss[i] = 0.f;
    float acc = 0.f;
    if (row < n) {
        for (int64_t k = row_ptr[row]; k < row_ptr[row + 1]; ++k) {
            const uint32_t a = attr[k];
            const float xv = x[k];
            if (lg == 0 && k1) acc = fmaf(w[a], xv, acc);
            const float* va = V + (size_t)a * K;
#pragma unroll
            for (int i = 0; i < NF; ++i) {
                const uint32_t f = lg + i * G;
                if (f < K) {
                    const float d = va[f] * xv;
                    s[i] += d;
                    ss[i] = fmaf(d, d, ss[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NF; ++i) acc += 0.5f * (s[i] * s[i] - ss[i]);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    double sq = 0.0;
    if (row < n && lg == 0) {
        const double pred = (double)acc + (k0 ? sc->w0 : 0.0);
        const double lo = sc->min_target, hi = sc->max_target;
        if (TRAIN) {
            const double pc = fmax(lo, fmin(hi, pred));
            const double err = pc - (double)y[row];
            sq = err * err;
            e[row] = (float)(pred - (double)y[row]);
        } else {
            pred_this[row] = (float)pred;
            if (accumulate) {
                const double ps = pred_sum[row] + fmax(lo, fmin(hi, pred));
                pred_sum[row] = ps;
                const double avg = ps * (1.0 / ((double)sc->iter + 1.0));
                const double err = fmax(lo, fmin(hi, avg)) - (double)y[row];
                sq = err * err;
            }
        }
    }
    const double tot = block_sum(sq, sm);
    if (threadIdx.x == 0 && part) part[blockIdx.x] = tot;
}

// ---------------------------------------------------------------------------------------------------------------------
// scalars: sum e, sum e^2 -> alpha ([G]:901-929), w0 ([G]:628-668)
__global__ void __launch_bounds__(BLOCK_T) stats_kernel(const float* __restrict__ e, uint32_t n, double* __restrict__ part /* [grid][2] */)
{
    __shared__ double sm[BLOCK_T / 32];
    double s1 = 0.0, s2 = 0.0;
    for (uint32_t i = blockIdx.x * BLOCK_T + threadIdx.x; i < n; i += gridDim.x * BLOCK_T) {
        const double v = (double)e[i];
        s1 += v;
        s2 += v * v;
    }
    const double t1 = block_sum(s1, sm);
    const double t2 = block_sum(s2, sm);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = t1;
        part[2 * blockIdx.x + 1] = t2;
    }
}
__global__ void global_draw_kernel(Scal* sc, const double* __restrict__ part, uint32_t nblocks, uint32_t n, int k0, int do_sample, int do_multilevel,
                                   int zero, double reg0, uint64_t seed)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double S1 = 0.0, S2 = 0.0;
    for (uint32_t b = 0; b < nblocks; ++b) {
        S1 += part[2 * b];
        S2 += part[2 * b + 1];
    }
    const uint32_t it = sc->iter;
    const int gmode = zero ? sbmf::SAMPLE_ZERO : sbmf::SAMPLE_SQRT;
    double alpha = sc->alpha;
    if (!do_multilevel) {
        alpha = ALPHA_0;
    } else {
        const double a = draw_gamma_f64(gmode, seed, SITE_FM_ALPHA, 0, it, (ALPHA_0 + (double)n) / 2.0, (GAMMA_0 + S2) / 2.0);   // not gated by do_sample
        if (!isnan(a) && !isinf(a)) alpha = a;
    }
    sc->alpha = alpha;
    sc->w0_delta = 0.0;
    if (k0) {
        const double w0 = sc->w0;
        const double sum = S1 - (double)n * w0;                  // sum (e - w0)
        const double var = 1.0 / (reg0 + alpha * (double)n);
        const double mean = -var * (alpha * sum - W0_MEAN_0 * reg0);
        const double z = (do_sample && !zero) ? normal_f64(philox_site(seed, SITE_FM_W0, 0, 0, it)) : 0.0;
        const double nw = mean + sqrt(var) * z;
        if (!isnan(nw) && !isinf(nw)) {
            sc->w0 = nw;
            sc->w0_delta = w0 - nw;
        }
    }
}
__global__ void shift_kernel(float* __restrict__ e, uint32_t n, const Scal* __restrict__ sc)
{
    const float d = (float)sc->w0_delta;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) e[i] -= d;
}

// group hyper-parameters: S1 = sum theta, S2 = sum (theta - mu_old)^2 per (chunk of a group, f), then one thread per (group, f)
__global__ void __launch_bounds__(BLOCK_T) hyper_reduce_kernel(const float* __restrict__ theta, uint32_t stride, const uint32_t* __restrict__ gs_attr,
                                                                const uint32_t* __restrict__ chunk_begin, const uint32_t* __restrict__ group,
                                                                const double* __restrict__ mu, uint32_t F, double* __restrict__ part)
{
    __shared__ double sm[BLOCK_T / 32];
    const uint32_t ch = blockIdx.x, f = blockIdx.y;
    const uint32_t b = chunk_begin[ch], en = chunk_begin[ch + 1];
    const double m = mu[(size_t)group[gs_attr[b]] * F + f];
    double s1 = 0.0, s2 = 0.0;
    for (uint32_t i = b + threadIdx.x; i < en; i += BLOCK_T) {
        const double t = (double)theta[(size_t)gs_attr[i] * stride + f];
        s1 += t;
        s2 += (t - m) * (t - m);
    }
    const double t1 = block_sum(s1, sm);
    const double t2 = block_sum(s2, sm);
    if (threadIdx.x == 0) {
        part[((size_t)ch * F + f) * 2] = t1;
        part[((size_t)ch * F + f) * 2 + 1] = t2;
    }
}
__global__ void hyper_draw_kernel(double* __restrict__ mu, double* __restrict__ lambda, const double* __restrict__ part,
                                  const uint32_t* __restrict__ gchunk_ptr, const uint32_t* __restrict__ n_per_group, uint32_t G, uint32_t F,
                                  const Scal* __restrict__ sc, int do_sample, int do_multilevel, int zero, uint64_t seed, uint32_t site_lambda,
                                  uint32_t site_mu)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G * F) return;
    const uint32_t g = t / F, f = t % F;
    if (!do_multilevel) {          // [G]:932-935, 971-973: lambda keeps its -regular value, mu = mu_0
        mu[t] = MU_0;
        return;
    }
    double S1 = 0.0, S2 = 0.0;
    for (uint32_t ch = gchunk_ptr[g]; ch < gchunk_ptr[g + 1]; ++ch) {
        S1 += part[((size_t)ch * F + f) * 2];
        S2 += part[((size_t)ch * F + f) * 2 + 1];
    }
    const uint32_t it = sc->iter;
    const GroupPosterior gp = group_posterior(S1, S2, (double)n_per_group[g], mu[t], ALPHA_0, BETA_0, GAMMA_0, MU_0);
    double lam = lambda[t];
    {
        const int gmode = (zero || !do_sample) ? sbmf::SAMPLE_ZERO : sbmf::SAMPLE_SQRT;   // not sampled: shape / rate ([G]:989-991)
        const double l = draw_gamma_f64(gmode, seed, site_lambda, t, it, gp.lambda_shape, gp.lambda_rate);
        if (!isnan(l) && !isinf(l)) lam = l;
    }
    lambda[t] = lam;
    const double var = 1.0 / (gp.n_beta * lam);
    const double z = (do_sample && !zero) ? normal_f64(philox_site(seed, site_mu, g, f, it)) : 0.0;
    const double nm = gp.mu_mean + sqrt(var) * z;
    if (!isnan(nm) && !isinf(nm)) mu[t] = nm;
}

constexpr int FIN_T = 256;
__global__ void __launch_bounds__(FIN_T) finish_iteration_kernel(Scal* sc, const double* __restrict__ part_train, uint32_t nb_train, uint32_t n_train,
                                        const double* __restrict__ part_test, uint32_t nb_test, uint32_t n_test, double* __restrict__ hist)
{
    // one CTA of FIN_T threads: strided partial sums, then a fixed tree (the same order whatever the grid of the predict kernels
    // was).  A single thread walking the ~2,400 partials took 1.2 ms -- a quarter of an iteration on the ML-1M-shaped matrix.
    __shared__ double sa[FIN_T], sb[FIN_T];
    if (blockIdx.x != 0) return;
    double a = 0.0, b = 0.0;
    for (uint32_t i = threadIdx.x; i < nb_train; i += FIN_T) a += part_train[i];
    for (uint32_t i = threadIdx.x; i < nb_test; i += FIN_T) b += part_test[i];
    sa[threadIdx.x] = a;
    sb[threadIdx.x] = b;
    __syncthreads();
    for (int o = FIN_T / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sa[threadIdx.x] += sa[threadIdx.x + o];
            sb[threadIdx.x] += sb[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    a = sa[0];
    b = sb[0];
    const uint32_t it = sc->iter;
    if (it < HIST_CAP) {
        hist[2 * it] = sqrt(a / (double)n_train);
        hist[2 * it + 1] = sqrt(b / (double)n_test);       // NaN without a test set, like libFM's 0 / 0
    }
    sc->iter = it + 1;
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
static uint32_t grid_for(const Model& m, int64_t n, int t = BLOCK_T)
{
    return (uint32_t)std::max<int64_t>(1, std::min<int64_t>((n + t - 1) / t, (int64_t)m.sm_count * 16));
}

static void free_rowform(RowForm& r)
{
    dfree(r.row_ptr); dfree(r.attr); dfree(r.x); dfree(r.y);
    r.n = 0;
    r.nnz = 0;
}
static void free_train(Model& m)
{
    free_rowform(m.tr);
    dfree(m.col_ptr); dfree(m.case_id); dfree(m.xc); dfree(m.wl_cols); dfree(m.wl_slices); dfree(m.wl_long); dfree(m.slice_part);
    dfree(m.long_scratch); dfree(m.e); dfree(m.q);
    m.runs.clear();
    m.run_begin.clear();
    m.have_train = false;
    m.inited = false;
}
static void free_test(Model& m)
{
    free_rowform(m.te);
    dfree(m.pred_this); dfree(m.pred_sum);
    m.have_test = false;
}

static int check_rowform(Model& m, const char* what, uint32_t n, const int64_t* row_ptr, const uint32_t* attr, const float* x, const float* y)
{
    if (!row_ptr || (n && !y)) {
        m.err = std::string(what) + ": null array";
        return SBMF_ERR_INVALID;
    }
    if (row_ptr[0] != 0) {
        m.err = std::string(what) + ": row_ptr[0] must be 0";
        return SBMF_ERR_INVALID;
    }
    for (uint32_t r = 0; r < n; ++r)
        if (row_ptr[r + 1] < row_ptr[r]) {
            m.err = std::string(what) + ": row_ptr is not non-decreasing";
            return SBMF_ERR_INVALID;
        }
    const int64_t nnz = row_ptr[n];
    if (nnz >= (1ll << 31)) {
        m.err = std::string(what) + ": more than 2^31-1 design-matrix entries are not supported";
        return SBMF_ERR_UNSUPPORTED;
    }
    if (nnz && (!attr || !x)) {
        m.err = std::string(what) + ": null array";
        return SBMF_ERR_INVALID;
    }
    for (int64_t k = 0; k < nnz; ++k)
        if (attr[k] >= m.p) {
            m.err = std::string(what) + ": attribute id " + std::to_string(attr[k]) + " >= num_attr";
            return SBMF_ERR_INVALID;
        }
    return SBMF_OK;
}

// uploads a row-form matrix and rewrites it with ascending attribute ids inside every row; for the train set also builds the
// column form (Data.h:472-528: cases ascending within a column), next_attr and the duplicate check
static int build_matrix(Model& m, RowForm& r, uint32_t n, const int64_t* row_ptr, const uint32_t* attr, const float* x, const float* y, bool train,
                        std::vector<uint32_t>* next_attr_h)
{
    cudaStream_t st = m.st;
    const int64_t nnz = row_ptr[n];
    r.n = n;
    r.nnz = nnz;
    CK(dmalloc(&r.row_ptr, (size_t)n + 1)); CK(dmalloc(&r.attr, (size_t)nnz)); CK(dmalloc(&r.x, (size_t)nnz)); CK(dmalloc(&r.y, (size_t)n));
    CK(cudaMemcpyAsync(r.row_ptr, row_ptr, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n) CK(cudaMemcpyAsync(r.y, y, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    uint32_t *d_attr = nullptr, *d_row = nullptr, *d_iota = nullptr, *d_key = nullptr, *d_ord = nullptr, *d_case = nullptr, *d_next = nullptr, *d_dup = nullptr;
    float *d_x = nullptr, *d_xc = nullptr;
    void* d_tmp = nullptr;
    auto cleanup = [&]() {
        for (void* q : {(void*)d_attr, (void*)d_row, (void*)d_iota, (void*)d_key, (void*)d_ord, (void*)d_case, (void*)d_next, (void*)d_dup, (void*)d_x,
                        (void*)d_xc, d_tmp})
            if (q) cudaFree(q);
    };
#define CKC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            m.err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            cleanup();                                                                             \
            return (e_ == cudaErrorMemoryAllocation) ? SBMF_ERR_NOMEM : SBMF_ERR_CUDA;             \
        }                                                                                          \
    } while (0)
    const size_t z = (size_t)nnz;
    CKC(dmalloc(&d_attr, z)); CKC(dmalloc(&d_x, z)); CKC(dmalloc(&d_row, z)); CKC(dmalloc(&d_iota, z)); CKC(dmalloc(&d_key, z));
    CKC(dmalloc(&d_ord, z)); CKC(dmalloc(&d_case, z)); CKC(dmalloc(&d_xc, z)); CKC(dmalloc(&d_next, (size_t)m.p)); CKC(dmalloc(&d_dup, 1));
    if (nnz) {
        CKC(cudaMemcpyAsync(d_attr, attr, z * 4, cudaMemcpyHostToDevice, st));
        CKC(cudaMemcpyAsync(d_x, x, z * 4, cudaMemcpyHostToDevice, st));
    }
    size_t tmp_bytes = 0;
    CKC(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_attr, d_key, d_iota, d_ord, (int)nnz, 0, 32, st));
    CKC(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1));
    const uint32_t g = grid_for(m, nnz);
    CKC(cudaMemsetAsync(d_dup, 0, 4, st));
    FM_LAUNCH(fill_u32_kernel, (m.p + 255) / 256, 256, st, d_next, m.p, UINT32_MAX);
    if (nnz) {
        FM_LAUNCH(entry_row_kernel, g, BLOCK_T, st, r.row_ptr, n, nnz, d_row);
        FM_LAUNCH(iota_kernel, g, BLOCK_T, st, d_iota, nnz);
        // column form: stable sort of the entries by attribute (entries are in case order)
        CKC(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_attr, d_key, d_iota, d_ord, (int)nnz, 0, 32, st));
        FM_LAUNCH(gather_u32_kernel, g, BLOCK_T, st, d_row, d_ord, d_case, nnz);
        FM_LAUNCH(gather_f32_kernel, g, BLOCK_T, st, d_x, d_ord, d_xc, nnz);
    }
    if (train) {
        CKC(dmalloc(&m.col_ptr, (size_t)m.p + 1)); CKC(dmalloc(&m.case_id, z)); CKC(dmalloc(&m.xc, z));
        FM_LAUNCH(seg_ptr_kernel, (m.p + 1 + 255) / 256, 256, st, d_key, nnz, m.p, m.col_ptr);
        if (nnz) {
            CKC(cudaMemcpyAsync(m.case_id, d_case, z * 4, cudaMemcpyDeviceToDevice, st));
            CKC(cudaMemcpyAsync(m.xc, d_xc, z * 4, cudaMemcpyDeviceToDevice, st));
        }
    }
    if (nnz) {
        // row form with ascending attributes: stable sort of the column form by case (d_key = attribute of every column-form entry)
        uint32_t* d_case_sorted = d_attr;   // reuse: the caller's attribute order is no longer needed
        CKC(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_case, d_case_sorted, d_iota, d_ord, (int)nnz, 0, 32, st));
        FM_LAUNCH(gather_u32_kernel, g, BLOCK_T, st, d_key, d_ord, r.attr, nnz);
        FM_LAUNCH(gather_f32_kernel, g, BLOCK_T, st, d_xc, d_ord, r.x, nnz);
        // the sorted case keys ARE the row of every entry of the new row form
        FM_LAUNCH(next_attr_kernel, g, BLOCK_T, st, r.attr, d_case_sorted, nnz, d_next, d_dup);
    }
    CKC(cudaGetLastError());
    uint32_t dup = 0;
    CKC(cudaMemcpyAsync(&dup, d_dup, 4, cudaMemcpyDeviceToHost, st));
    if (next_attr_h) {
        next_attr_h->resize(m.p);
        CKC(cudaMemcpyAsync(next_attr_h->data(), d_next, (size_t)m.p * 4, cudaMemcpyDeviceToHost, st));
    }
    CKC(cudaStreamSynchronize(st));
    cleanup();
#undef CKC
    if (dup) {
        m.err = "design matrix: an attribute is listed twice in one case";
        return SBMF_ERR_INVALID;
    }
    return SBMF_OK;
}

static int build_worklists(Model& m, const std::vector<uint32_t>& next_attr)
{
    std::vector<int64_t> cp((size_t)m.p + 1);
    CK(cudaMemcpy(cp.data(), m.col_ptr, cp.size() * 8, cudaMemcpyDeviceToHost));
    m.run_begin.assign((size_t)m.p + 1, 0);
    const uint32_t nruns = plan_runs(m.p, next_attr.data(), m.run_begin.data());
    m.run_begin.resize((size_t)nruns + 1);
    std::vector<uint32_t> wcols, ccols;
    std::vector<LongCol> lcols;
    std::vector<Slice> slices;
    m.runs.assign(nruns, Run());
    uint32_t max_slices = 1, max_long = 1;
    for (uint32_t r = 0; r < nruns; ++r) {
        Run& R = m.runs[r];
        R.w_off = (uint32_t)wcols.size(); R.c_off = (uint32_t)ccols.size(); R.g_off = (uint32_t)lcols.size(); R.s_off = (uint32_t)slices.size();
        for (uint32_t j = m.run_begin[r]; j < m.run_begin[r + 1]; ++j) {
            const int64_t len = cp[j + 1] - cp[j];
            if (len <= WARP_COL_MAX) wcols.push_back(j);
            else if (len <= BLOCK_COL_MAX) ccols.push_back(j);
            else {
                LongCol lc{j, (uint32_t)slices.size(), 0, 0};
                for (int64_t b = cp[j]; b < cp[j + 1]; b += SLICE_LEN)
                    slices.push_back(Slice{b, (uint32_t)std::min<int64_t>(SLICE_LEN, cp[j + 1] - b), (uint32_t)lcols.size() - R.g_off});
                lc.slice_end = (uint32_t)slices.size();
                lcols.push_back(lc);
            }
        }
        R.w_cnt = (uint32_t)wcols.size() - R.w_off; R.c_cnt = (uint32_t)ccols.size() - R.c_off;
        R.g_cnt = (uint32_t)lcols.size() - R.g_off; R.s_cnt = (uint32_t)slices.size() - R.s_off;
        max_slices = std::max(max_slices, R.s_cnt);
        max_long = std::max(max_long, R.g_cnt);
    }
    const size_t nw = wcols.size();
    wcols.insert(wcols.end(), ccols.begin(), ccols.end());      // one array: warp-tier lists first, then the CTA-tier lists
    for (Run& R : m.runs) R.c_off += (uint32_t)nw;
    CK(dmalloc(&m.wl_cols, wcols.size())); CK(dmalloc(&m.wl_slices, slices.size())); CK(dmalloc(&m.wl_long, lcols.size()));
    CK(dmalloc(&m.slice_part, (size_t)max_slices)); CK(dmalloc(&m.long_scratch, (size_t)max_long));
    if (!wcols.empty()) CK(cudaMemcpy(m.wl_cols, wcols.data(), wcols.size() * 4, cudaMemcpyHostToDevice));
    if (!slices.empty()) CK(cudaMemcpy(m.wl_slices, slices.data(), slices.size() * sizeof(Slice), cudaMemcpyHostToDevice));
    if (!lcols.empty()) CK(cudaMemcpy(m.wl_long, lcols.data(), lcols.size() * sizeof(LongCol), cudaMemcpyHostToDevice));
    return SBMF_OK;
}

static int build_groups(Model& m)
{
    dfree(m.group); dfree(m.gs_attr); dfree(m.chunk_begin); dfree(m.gchunk_ptr); dfree(m.n_per_group); dfree(m.hyper_part);
    const uint32_t p = m.p, G = m.G;
    std::vector<uint32_t> npg(G, 0), start(G + 1, 0), gs(p), cb, gcp(G + 1, 0);
    for (uint32_t j = 0; j < p; ++j) npg[m.group_h[j]]++;
    for (uint32_t g = 0; g < G; ++g) start[g + 1] = start[g] + npg[g];
    {
        std::vector<uint32_t> fill(start.begin(), start.end() - 1);
        for (uint32_t j = 0; j < p; ++j) gs[fill[m.group_h[j]]++] = j;
    }
    for (uint32_t g = 0; g < G; ++g) {
        gcp[g] = (uint32_t)cb.size();
        for (uint32_t b = start[g]; b < start[g + 1]; b += HYPER_CHUNK) cb.push_back(b);
    }
    gcp[G] = (uint32_t)cb.size();
    m.nchunks = (uint32_t)cb.size();
    // chunk_begin[ch + 1] must be the END of chunk ch even at a group boundary: chunks are consecutive in gs, so it is
    cb.push_back(p);
    const uint32_t F = std::max<uint32_t>(m.K, 1);
    CK(dmalloc(&m.group, (size_t)p)); CK(dmalloc(&m.gs_attr, (size_t)p)); CK(dmalloc(&m.chunk_begin, cb.size())); CK(dmalloc(&m.gchunk_ptr, (size_t)G + 1));
    CK(dmalloc(&m.n_per_group, (size_t)G)); CK(dmalloc(&m.hyper_part, (size_t)std::max<uint32_t>(m.nchunks, 1) * F * 2));
    CK(cudaMemcpy(m.group, m.group_h.data(), (size_t)p * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.gs_attr, gs.data(), (size_t)p * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.chunk_begin, cb.data(), cb.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.gchunk_ptr, gcp.data(), ((size_t)G + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.n_per_group, npg.data(), (size_t)G * 4, cudaMemcpyHostToDevice));
    return SBMF_OK;
}

static bool live_noise(const Model& m) { return m.cfg.do_sample && m.cfg.sample_mode == SBMF_FM_SAMPLE_LIVE; }

template <int COORD>
static void launch_run(Model& m, const ColArgs& a, const Run& R)
{
    cudaStream_t st = m.st;
    // tiers on their own streams when the run has more than one: the sliced chain (three dependent launches) stays on st, the
    // warp and CTA tiers fork.  slice_part / long_scratch belong to the sliced tier alone, e / q entries are disjoint across tiers.
    const int tiers = (R.w_cnt ? 1 : 0) + (R.c_cnt ? 1 : 0) + (R.g_cnt ? 1 : 0);
    const bool fork = tiers > 1 && m.st_tier[0] && m.st_tier[1];
    cudaStream_t sw = st, sc = st;
    if (fork) {
        cudaEventRecord(m.ev_fork, st);
        if (R.w_cnt && (R.c_cnt || R.g_cnt)) {
            sw = m.st_tier[0];
            cudaStreamWaitEvent(sw, m.ev_fork, 0);
        }
        if (R.c_cnt && R.g_cnt) {
            sc = m.st_tier[1];
            cudaStreamWaitEvent(sc, m.ev_fork, 0);
        }
    }
    if (R.w_cnt) {
        FM_LAUNCH(col_warp_kernel<COORD>, (R.w_cnt + BLOCK_T / 32 - 1) / (BLOCK_T / 32), BLOCK_T, sw, a, m.wl_cols + R.w_off, R.w_cnt);
        m.launches++;
    }
    if (R.c_cnt) {
        FM_LAUNCH(col_block_kernel<COORD>, R.c_cnt, BLOCK_T, sc, a, m.wl_cols + R.c_off);
        m.launches++;
    }
    if (R.g_cnt) {
        FM_LAUNCH(slice_reduce_kernel<COORD>, R.s_cnt, BLOCK_T, st, a, m.wl_slices + R.s_off, m.wl_long + R.g_off, m.slice_part);
        FM_LAUNCH(slice_draw_kernel<COORD>, (R.g_cnt + 127) / 128, 128, st, a, m.wl_long + R.g_off, R.g_cnt, R.s_off, m.slice_part, m.long_scratch);
        FM_LAUNCH(slice_apply_kernel<COORD>, R.s_cnt, BLOCK_T, st, a, m.wl_slices + R.s_off, m.long_scratch);
        m.launches += 3;
    }
    if (sw != st) {
        cudaEventRecord(m.ev_join[0], sw);
        cudaStreamWaitEvent(st, m.ev_join[0], 0);
    }
    if (sc != st) {
        cudaEventRecord(m.ev_join[1], sc);
        cudaStreamWaitEvent(st, m.ev_join[1], 0);
    }
}

static void launch_predict(Model& m, const RowForm& r, bool train, int accumulate, double* part, uint32_t& nblocks)
{
    const uint32_t lanes = m.K <= 16 ? 8 : 32;
    nblocks = (uint32_t)(((uint64_t)r.n * lanes + BLOCK_T - 1) / BLOCK_T);
    if (nblocks == 0) return;
#define PREDICT(Gv, TRv)                                                                                                             \
    {                                                                                                                                \
        auto kfn = predict_kernel<Gv, TRv>;                                                                                          \
        FM_LAUNCH(kfn, nblocks, BLOCK_T, m.st, r.row_ptr, r.attr, r.x, r.y, r.n, m.w, m.V, m.K, m.cfg.k0, m.cfg.k1, m.sc, m.e,       \
                  m.pred_this, m.pred_sum, accumulate, part);                                                                        \
    }
    if (lanes == 8) {
        if (train) PREDICT(8, true)
        else PREDICT(8, false)
    } else {
        if (train) PREDICT(32, true)
        else PREDICT(32, false)
    }
#undef PREDICT
    m.launches++;
}

static uint32_t predict_blocks(const Model& m, uint32_t n)
{
    const uint32_t lanes = m.K <= 16 ? 8 : 32;
    return (uint32_t)(((uint64_t)n * lanes + BLOCK_T - 1) / BLOCK_T);
}

// one iteration of [GS]:97-262: draw_all ([G]:411-626), re-prediction of train and test, running test prediction
static int enqueue_iteration(Model& m)
{
    cudaStream_t st = m.st;
    const sbmf_fm_config& c = m.cfg;
    const int zero = c.sample_mode == SBMF_FM_SAMPLE_ZERO_NOISE;
    const int live = live_noise(m);
    const uint32_t n = m.tr.n;
    double* part_stats = m.red_part;
    FM_LAUNCH(stats_kernel, m.red_blocks, BLOCK_T, st, m.e, n, part_stats);
    FM_LAUNCH(global_draw_kernel, 1, 32, st, m.sc, part_stats, m.red_blocks, n, c.k0, c.do_sample, c.do_multilevel, zero, c.reg0, c.seed);
    m.launches += 2;
    if (c.k0) {
        FM_LAUNCH(shift_kernel, grid_for(m, n), BLOCK_T, st, m.e, n, m.sc);
        m.launches++;
    }
    ColArgs a;
    a.col_ptr = m.col_ptr; a.case_id = m.case_id; a.xc = m.xc; a.e = m.e; a.q = m.q; a.group = m.group; a.sc = m.sc; a.seed = c.seed; a.live = live;
    if (c.k1) {
        if (m.nchunks) FM_LAUNCH(hyper_reduce_kernel, dim3(m.nchunks, 1), BLOCK_T, st, m.w, 1, m.gs_attr, m.chunk_begin, m.group, m.w_mu, 1, m.hyper_part);
        FM_LAUNCH(hyper_draw_kernel, (m.G + 127) / 128, 128, st, m.w_mu, m.w_lambda, m.hyper_part, m.gchunk_ptr, m.n_per_group, m.G, 1, m.sc, c.do_sample,
                                                            c.do_multilevel, zero, c.seed, SITE_FM_W_LAMBDA, SITE_FM_W_MU);
        m.launches += 2;
        a.theta = m.w; a.stride = 1; a.f = 0; a.mu = m.w_mu; a.lambda = m.w_lambda; a.hstride = 1; a.site = SITE_FM_W;
        for (const Run& R : m.runs) launch_run<COORD_W>(m, a, R);
    }
    if (m.K > 0) {
        if (m.nchunks) FM_LAUNCH(hyper_reduce_kernel, dim3(m.nchunks, m.K), BLOCK_T, st, m.V, m.K, m.gs_attr, m.chunk_begin, m.group, m.v_mu, m.K, m.hyper_part);
        FM_LAUNCH(hyper_draw_kernel, (m.G * m.K + 127) / 128, 128, st, m.v_mu, m.v_lambda, m.hyper_part, m.gchunk_ptr, m.n_per_group, m.G, m.K, m.sc,
                                                                  c.do_sample, c.do_multilevel, zero, c.seed, SITE_FM_V_LAMBDA, SITE_FM_V_MU);
        m.launches += 2;
        a.theta = m.V; a.stride = m.K; a.mu = m.v_mu; a.lambda = m.v_lambda; a.hstride = m.K; a.site = SITE_FM_V;
        const double avg_row = n ? (double)m.tr.nnz / n : 0.0;
        for (uint32_t f = 0; f < m.K; ++f) {
            if (avg_row <= 4.0) FM_LAUNCH(q_rebuild_kernel<1>, (n + BLOCK_T - 1) / BLOCK_T, BLOCK_T, st, m.tr.row_ptr, m.tr.attr, m.tr.x, n, m.V, m.K, f, m.q);
            else if (avg_row <= 64.0)
                FM_LAUNCH(q_rebuild_kernel<8>, (uint32_t)(((uint64_t)n * 8 + BLOCK_T - 1) / BLOCK_T), BLOCK_T, st, m.tr.row_ptr, m.tr.attr, m.tr.x, n, m.V, m.K, f, m.q);
            else
                FM_LAUNCH(q_rebuild_kernel<32>, (uint32_t)(((uint64_t)n * 32 + BLOCK_T - 1) / BLOCK_T), BLOCK_T, st, m.tr.row_ptr, m.tr.attr, m.tr.x, n, m.V, m.K, f, m.q);
            m.launches++;
            a.f = f;
            for (const Run& R : m.runs) launch_run<COORD_V>(m, a, R);
        }
    }
    double* part_train = m.red_part + 2 * (size_t)m.red_blocks;
    double* part_test = part_train + predict_blocks(m, m.tr.n);
    uint32_t nb_train = 0, nb_test = 0;
    launch_predict(m, m.tr, true, 0, part_train, nb_train);
    if (m.have_test) launch_predict(m, m.te, false, 1, part_test, nb_test);
    FM_LAUNCH(finish_iteration_kernel, 1, FIN_T, st, m.sc, part_train, nb_train, m.tr.n, part_test, nb_test, m.have_test ? m.te.n : 0, m.hist);
    m.launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        m.err = std::string("learn: kernel launch failed: ") + cudaGetErrorString(e);
        return SBMF_ERR_CUDA;
    }
    return SBMF_OK;
}

}  // namespace sbmf_fm

// =====================================================================================================================
using namespace sbmf_fm;

struct sbmf_fm_handle {
    Model m;
};

static std::string g_fm_create_err;

#define API_CK(call)                                                                               \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            m.err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            return (e_ == cudaErrorMemoryAllocation) ? SBMF_ERR_NOMEM : SBMF_ERR_CUDA;             \
        }                                                                                          \
    } while (0)

extern "C" {

int sbmf_fm_config_default(sbmf_fm_config* cfg)
{
    if (!cfg) return SBMF_ERR_INVALID;
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = (uint32_t)sizeof(sbmf_fm_config);
    cfg->num_groups = 1;
    cfg->K = 8;                // [L]:128 -dim default '1,1,8'
    cfg->k0 = cfg->k1 = 1;
    cfg->do_sample = 1;        // [L]:418
    cfg->do_multilevel = 1;    // [L]:419
    cfg->sample_mode = SBMF_FM_SAMPLE_LIVE;
    cfg->seed = 1;
    cfg->init_stdev = 0.1;     // [L]:127
    return SBMF_OK;
}

int sbmf_fm_create(const sbmf_fm_config* cfg, sbmf_fm_handle** out)
{
    if (!cfg || !out) {
        g_fm_create_err = "fm create: null argument";
        return SBMF_ERR_INVALID;
    }
    *out = nullptr;
    if (cfg->struct_size != sizeof(sbmf_fm_config)) {
        g_fm_create_err = "fm create: sbmf_fm_config.struct_size mismatch (use sbmf_fm_config_default)";
        return SBMF_ERR_INVALID;
    }
    if (cfg->num_attr == 0 || cfg->num_groups == 0 || cfg->K > SBMF_MAX_K || cfg->sample_mode < 0 || cfg->sample_mode > 1 ||
        cfg->num_groups > cfg->num_attr) {
        g_fm_create_err = "fm create: need num_attr >= num_groups >= 1, K <= SBMF_MAX_K and a known sample_mode";
        return SBMF_ERR_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_fm_create_err = std::string("fm create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
        return SBMF_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) {
        g_fm_create_err = "fm create: device ordinal out of range";
        return SBMF_ERR_INVALID;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) {
        g_fm_create_err = std::string("fm create: cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        return SBMF_ERR_CUDA;
    }
    if (prop.major != 10) {
        g_fm_create_err = "fm create: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ", this library is built for sm_100a (B200) only";
        return SBMF_ERR_UNSUPPORTED;
    }
    sbmf_fm_handle* h = new (std::nothrow) sbmf_fm_handle();
    if (!h) {
        g_fm_create_err = "fm create: out of host memory";
        return SBMF_ERR_NOMEM;
    }
    Model& m = h->m;
    m.cfg = *cfg;
    m.p = cfg->num_attr; m.K = cfg->K; m.G = cfg->num_groups;
    m.sm_count = prop.multiProcessorCount;
    m.group_h.assign(m.p, 0);
    const uint32_t F = std::max<uint32_t>(m.K, 1);
    bool ok = cudaSetDevice(cfg->device) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&m.st, cudaStreamNonBlocking) == cudaSuccess;
    for (int q = 0; q < 2 && ok; ++q) {
        ok = cudaStreamCreateWithFlags(&m.st_tier[q], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&m.ev_join[q], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&m.ev_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && dmalloc(&m.sc, 1) == cudaSuccess && cudaMemset(m.sc, 0, sizeof(Scal)) == cudaSuccess;
    ok = ok && dmalloc(&m.w, (size_t)m.p) == cudaSuccess && dmalloc(&m.V, (size_t)m.p * F) == cudaSuccess;
    ok = ok && dmalloc(&m.w_mu, (size_t)m.G) == cudaSuccess && dmalloc(&m.w_lambda, (size_t)m.G) == cudaSuccess;
    ok = ok && dmalloc(&m.v_mu, (size_t)m.G * F) == cudaSuccess && dmalloc(&m.v_lambda, (size_t)m.G * F) == cudaSuccess;
    ok = ok && dmalloc(&m.hist, (size_t)HIST_CAP * 2) == cudaSuccess;
    if (!ok) {
        g_fm_create_err = std::string("fm create: CUDA resource setup failed: ") + cudaGetErrorString(cudaGetLastError());
        sbmf_fm_destroy(h);
        return SBMF_ERR_CUDA;
    }
    if (build_groups(m) != SBMF_OK) {
        g_fm_create_err = "fm create: " + m.err;
        sbmf_fm_destroy(h);
        return SBMF_ERR_CUDA;
    }
    *out = h;
    return SBMF_OK;
}

int sbmf_fm_destroy(sbmf_fm_handle* h)
{
    if (!h) return SBMF_OK;
    Model& m = h->m;
    cudaSetDevice(m.cfg.device);
    if (m.st) cudaStreamSynchronize(m.st);
    free_train(m);
    free_test(m);
    dfree(m.group); dfree(m.gs_attr); dfree(m.chunk_begin); dfree(m.gchunk_ptr); dfree(m.n_per_group); dfree(m.hyper_part);
    dfree(m.w); dfree(m.V); dfree(m.w_mu); dfree(m.w_lambda); dfree(m.v_mu); dfree(m.v_lambda); dfree(m.sc); dfree(m.red_part); dfree(m.hist);
    for (int q = 0; q < 2; ++q) {
        if (m.st_tier[q]) {
            cudaStreamSynchronize(m.st_tier[q]);
            cudaStreamDestroy(m.st_tier[q]);
        }
        if (m.ev_join[q]) cudaEventDestroy(m.ev_join[q]);
    }
    if (m.ev_fork) cudaEventDestroy(m.ev_fork);
    if (m.st) cudaStreamDestroy(m.st);
    delete h;
    return SBMF_OK;
}

const char* sbmf_fm_last_error(const sbmf_fm_handle* h) { return h ? h->m.err.c_str() : g_fm_create_err.c_str(); }

int sbmf_fm_set_groups(sbmf_fm_handle* h, const uint32_t* attr_group)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!attr_group) {
        m.err = "set_groups: null array";
        return SBMF_ERR_INVALID;
    }
    if (m.inited) {
        m.err = "set_groups: call before init";
        return SBMF_ERR_STATE;
    }
    for (uint32_t j = 0; j < m.p; ++j)
        if (attr_group[j] >= m.G) {
            m.err = "set_groups: group id " + std::to_string(attr_group[j]) + " >= num_groups";
            return SBMF_ERR_INVALID;
        }
    API_CK(cudaSetDevice(m.cfg.device));
    m.group_h.assign(attr_group, attr_group + m.p);
    return build_groups(m);
}

int sbmf_fm_set_train(sbmf_fm_handle* h, uint32_t n, const int64_t* row_ptr, const uint32_t* attr, const float* x, const float* y)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (n == 0) {
        m.err = "set_train: no cases";
        return SBMF_ERR_INVALID;
    }
    int rc = check_rowform(m, "set_train", n, row_ptr, attr, x, y);
    if (rc != SBMF_OK) return rc;
    API_CK(cudaSetDevice(m.cfg.device));
    free_train(m);
    std::vector<uint32_t> next_attr;
    if ((rc = build_matrix(m, m.tr, n, row_ptr, attr, x, y, true, &next_attr)) != SBMF_OK || (rc = build_worklists(m, next_attr)) != SBMF_OK) {
        free_train(m);
        return rc;
    }
    API_CK(dmalloc(&m.e, (size_t)n));
    API_CK(dmalloc(&m.q, (size_t)n));
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;          // Data.h:193-201
    for (uint32_t c = 0; c < n; ++c) {
        mn = std::min(mn, y[c]);
        mx = std::max(mx, y[c]);
    }
    Scal s;
    memset(&s, 0, sizeof(s));
    s.min_target = mn;
    s.max_target = mx;
    API_CK(cudaMemcpy(m.sc, &s, sizeof(s), cudaMemcpyHostToDevice));
    m.have_train = true;
    m.inited = false;
    return SBMF_OK;
}

int sbmf_fm_set_test(sbmf_fm_handle* h, uint32_t n, const int64_t* row_ptr, const uint32_t* attr, const float* x, const float* y)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    int rc = check_rowform(m, "set_test", n, row_ptr, attr, x, y);
    if (rc != SBMF_OK) return rc;
    API_CK(cudaSetDevice(m.cfg.device));
    free_test(m);
    if ((rc = build_matrix(m, m.te, n, row_ptr, attr, x, y, false, nullptr)) != SBMF_OK) {
        free_test(m);
        return rc;
    }
    API_CK(dmalloc(&m.pred_this, (size_t)n));
    API_CK(dmalloc(&m.pred_sum, (size_t)n));
    API_CK(cudaMemset(m.pred_sum, 0, (size_t)(n ? n : 1) * 8));
    API_CK(cudaMemset(m.pred_this, 0, (size_t)(n ? n : 1) * 4));
    m.have_test = n > 0;
    m.inited = false;
    return SBMF_OK;
}

int sbmf_fm_init(sbmf_fm_handle* h, const float* w_init, const float* v_init)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.have_train) {
        m.err = "init: call set_train first";
        return SBMF_ERR_STATE;
    }
    API_CK(cudaSetDevice(m.cfg.device));
    cudaStream_t st = m.st;
    const uint32_t p = m.p, K = m.K, G = m.G, F = std::max<uint32_t>(K, 1);
    FM_LAUNCH(init_params_kernel, grid_for(m, (int64_t)p * (K + 1)), BLOCK_T, st, m.w, m.V, p, K, m.cfg.seed, (float)m.cfg.init_stdev, w_init == nullptr,
                                                                             v_init == nullptr);
    if (w_init) API_CK(cudaMemcpyAsync(m.w, w_init, (size_t)p * 4, cudaMemcpyHostToDevice, st));
    std::vector<float> vt;
    if (v_init && K) {                       // [K][p] (fm_model::v) -> attribute-major [p][K]
        vt.resize((size_t)p * K);
        for (uint32_t f = 0; f < K; ++f)
            for (uint32_t j = 0; j < p; ++j) vt[(size_t)j * K + f] = v_init[(size_t)f * p + j];
        API_CK(cudaMemcpyAsync(m.V, vt.data(), vt.size() * 4, cudaMemcpyHostToDevice, st));
    }
    // [G]:1099-1113 and [L]:485-505
    std::vector<double> wl(G, m.cfg.regw), vl((size_t)G * F, m.cfg.regv);
    API_CK(cudaMemsetAsync(m.w_mu, 0, (size_t)G * 8, st));
    API_CK(cudaMemsetAsync(m.v_mu, 0, (size_t)G * F * 8, st));
    API_CK(cudaMemcpyAsync(m.w_lambda, wl.data(), (size_t)G * 8, cudaMemcpyHostToDevice, st));
    API_CK(cudaMemcpyAsync(m.v_lambda, vl.data(), (size_t)G * F * 8, cudaMemcpyHostToDevice, st));
    Scal s;
    API_CK(cudaMemcpyAsync(&s, m.sc, sizeof(s), cudaMemcpyDeviceToHost, st));
    API_CK(cudaStreamSynchronize(st));
    s.w0 = 0.0; s.alpha = 1.0; s.w0_delta = 0.0; s.iter = 0;
    API_CK(cudaMemcpyAsync(m.sc, &s, sizeof(s), cudaMemcpyHostToDevice, st));
    if (m.have_test) API_CK(cudaMemsetAsync(m.pred_sum, 0, (size_t)m.te.n * 8, st));
    API_CK(cudaMemsetAsync(m.hist, 0, (size_t)HIST_CAP * 16, st));
    dfree(m.red_part);
    m.red_blocks = (uint32_t)m.sm_count * 4;
    API_CK(dmalloc(&m.red_part, 2 * (size_t)m.red_blocks + predict_blocks(m, m.tr.n) + predict_blocks(m, m.have_test ? m.te.n : 0) + 2));
    // [GS]:73-78: e = prediction - target
    uint32_t nb = 0;
    launch_predict(m, m.tr, true, 0, nullptr, nb);
    API_CK(cudaGetLastError());
    API_CK(cudaStreamSynchronize(st));
    m.iters_done = 0;
    m.inited = true;
    return SBMF_OK;
}

int sbmf_fm_learn(sbmf_fm_handle* h, uint32_t iters)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.inited) {
        m.err = "learn: call set_train, (set_test,) init first";
        return SBMF_ERR_STATE;
    }
    if ((uint64_t)m.iters_done + iters > HIST_CAP) {
        m.err = "learn: more than " + std::to_string(HIST_CAP) + " iterations per chain are not supported";
        return SBMF_ERR_UNSUPPORTED;
    }
    API_CK(cudaSetDevice(m.cfg.device));
    for (uint32_t i = 0; i < iters; ++i) {
        const int rc = enqueue_iteration(m);
        if (rc != SBMF_OK) return rc;
        m.iters_done++;
    }
    return SBMF_OK;
}

int sbmf_fm_rmse_history(sbmf_fm_handle* h, uint32_t first, uint32_t count, double* rmse_train, double* rmse_test)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if ((uint64_t)first + count > m.iters_done) {
        m.err = "rmse_history: range beyond the completed iterations";
        return SBMF_ERR_INVALID;
    }
    API_CK(cudaSetDevice(m.cfg.device));
    std::vector<double> buf((size_t)count * 2);
    API_CK(cudaMemcpyAsync(buf.data(), m.hist + 2 * (size_t)first, buf.size() * 8, cudaMemcpyDeviceToHost, m.st));
    API_CK(cudaStreamSynchronize(m.st));
    for (uint32_t i = 0; i < count; ++i) {
        if (rmse_train) rmse_train[i] = buf[2 * (size_t)i];
        if (rmse_test) rmse_test[i] = buf[2 * (size_t)i + 1];
    }
    return SBMF_OK;
}

int sbmf_fm_predict(sbmf_fm_handle* h, float* pred)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.have_test || !m.inited || m.iters_done == 0 || !pred) {
        m.err = "predict: needs a test set, a destination and at least one completed iteration";
        return SBMF_ERR_STATE;
    }
    API_CK(cudaSetDevice(m.cfg.device));
    const uint32_t n = m.te.n;
    Scal s;
    API_CK(cudaMemcpyAsync(&s, m.sc, sizeof(s), cudaMemcpyDeviceToHost, m.st));
    if (m.cfg.do_sample) {       // [G]:357-361
        std::vector<double> ps(n);
        API_CK(cudaMemcpyAsync(ps.data(), m.pred_sum, (size_t)n * 8, cudaMemcpyDeviceToHost, m.st));
        API_CK(cudaStreamSynchronize(m.st));
        for (uint32_t c = 0; c < n; ++c) pred[c] = (float)std::max(s.min_target, std::min(s.max_target, ps[c] / (double)m.iters_done));
    } else {                     // [G]:362-366
        API_CK(cudaMemcpyAsync(pred, m.pred_this, (size_t)n * 4, cudaMemcpyDeviceToHost, m.st));
        API_CK(cudaStreamSynchronize(m.st));
        for (uint32_t c = 0; c < n; ++c) pred[c] = (float)std::max(s.min_target, std::min(s.max_target, (double)pred[c]));
    }
    return SBMF_OK;
}

int sbmf_fm_get_state(sbmf_fm_handle* h, sbmf_fm_state* out)
{
    if (!h || !out) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.inited) {
        m.err = "get_state: call init first";
        return SBMF_ERR_STATE;
    }
    API_CK(cudaSetDevice(m.cfg.device));
    API_CK(cudaStreamSynchronize(m.st));
    const uint32_t p = m.p, K = m.K, G = m.G;
    if (out->w) API_CK(cudaMemcpy(out->w, m.w, (size_t)p * 4, cudaMemcpyDeviceToHost));
    if (out->v && K) {
        std::vector<float> vt((size_t)p * K);
        API_CK(cudaMemcpy(vt.data(), m.V, vt.size() * 4, cudaMemcpyDeviceToHost));
        for (uint32_t j = 0; j < p; ++j)
            for (uint32_t f = 0; f < K; ++f) out->v[(size_t)f * p + j] = vt[(size_t)j * K + f];
    }
    if (out->w_mu) API_CK(cudaMemcpy(out->w_mu, m.w_mu, (size_t)G * 8, cudaMemcpyDeviceToHost));
    if (out->w_lambda) API_CK(cudaMemcpy(out->w_lambda, m.w_lambda, (size_t)G * 8, cudaMemcpyDeviceToHost));
    if (out->v_mu && K) API_CK(cudaMemcpy(out->v_mu, m.v_mu, (size_t)G * K * 8, cudaMemcpyDeviceToHost));
    if (out->v_lambda && K) API_CK(cudaMemcpy(out->v_lambda, m.v_lambda, (size_t)G * K * 8, cudaMemcpyDeviceToHost));
    if (out->e) API_CK(cudaMemcpy(out->e, m.e, (size_t)m.tr.n * 4, cudaMemcpyDeviceToHost));
    if (out->pred_sum && m.have_test) API_CK(cudaMemcpy(out->pred_sum, m.pred_sum, (size_t)m.te.n * 8, cudaMemcpyDeviceToHost));
    Scal s;
    API_CK(cudaMemcpy(&s, m.sc, sizeof(s), cudaMemcpyDeviceToHost));
    out->w0 = s.w0;
    out->alpha = s.alpha;
    out->iterations = m.iters_done;
    return SBMF_OK;
}

int sbmf_fm_get_columns(sbmf_fm_handle* h, int64_t* col_ptr, uint32_t* case_id, float* x)
{
    if (!h) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.have_train) {
        m.err = "get_columns: no training set";
        return SBMF_ERR_STATE;
    }
    API_CK(cudaSetDevice(m.cfg.device));
    if (col_ptr) API_CK(cudaMemcpy(col_ptr, m.col_ptr, ((size_t)m.p + 1) * 8, cudaMemcpyDeviceToHost));
    if (case_id && m.tr.nnz) API_CK(cudaMemcpy(case_id, m.case_id, (size_t)m.tr.nnz * 4, cudaMemcpyDeviceToHost));
    if (x && m.tr.nnz) API_CK(cudaMemcpy(x, m.xc, (size_t)m.tr.nnz * 4, cudaMemcpyDeviceToHost));
    return SBMF_OK;
}

int sbmf_fm_get_runs(sbmf_fm_handle* h, uint32_t* n_runs, uint32_t* run_begin)
{
    if (!h || !n_runs) return SBMF_ERR_INVALID;
    Model& m = h->m;
    if (!m.have_train) {
        m.err = "get_runs: no training set";
        return SBMF_ERR_STATE;
    }
    *n_runs = (uint32_t)m.runs.size();
    if (run_begin) memcpy(run_begin, m.run_begin.data(), m.run_begin.size() * 4);
    return SBMF_OK;
}

uint32_t sbmf_fm_plan_runs(uint32_t num_attr, const uint32_t* next_attr, uint32_t* run_begin)
{
    if (!num_attr || !next_attr || !run_begin) return 0;
    return plan_runs(num_attr, next_attr, run_begin);
}

}  // extern "C"
