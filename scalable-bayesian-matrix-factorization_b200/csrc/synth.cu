// synth.cu -- synthetic MovieLens/Netflix-shaped rating matrices, generated on the device.
//
// Bench / test input only; the reference ships no generator (its fixtures are the MovieLens files under
// data/), so this one is defined to look like them:
//   * pair (i,j) is rated with probability p_ij = min(1, c * a_i * b_j), a_i = (rank_u(i)+1)^-s_user,
//     b_j = (rank_v(j)+1)^-s_item over random permutations of the ids; c is solved (bisection on the exact
//     expectation) so that the expected number of ratings is n_ratings.  Pairs are distinct by construction
//     and come out sorted by (user, item) like every shipped fixture.
//   * rating = clamp(round_to_half(3.5 + ub_i + vb_j + <P_i, Q_j> + 0.8 z), 0.5, 5) from a planted rank-16
//     model with N(0, 0.3^2) factors and biases, so RMSE curves are non-trivial.
//   * each pair goes to the test set with probability test_frac.
// Everything is a pure function of (spec.seed, i, j) through Philox4x32-10: the same spec reproduces the same
// matrix on any grid, and the two calls of the count-then-fill protocol agree.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "common.cuh"
#include "model.h"

namespace sbmf {

enum SynthSite : uint32_t { SY_PERM_U = 100, SY_PERM_V = 101, SY_PAIR = 102, SY_SPLIT = 103, SY_FACT_U = 104, SY_FACT_V = 105, SY_BIAS = 106 };
constexpr int SY_RANK = 16;

__global__ void sy_perm_keys(uint32_t* keys, uint32_t* vals, uint32_t n, uint64_t seed, uint32_t site)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = philox_site(seed, site, i, 0, 0).x;
    vals[i] = i;
}
// weight of the id that landed on rank r
__global__ void sy_weights(const uint32_t* order, float* w, uint32_t n, double s)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    w[order[r]] = (float)pow((double)r + 1.0, -s);
}
// sorted (descending) item weights + their inclusive prefix sums are tiny: built on the host.

// expected number of ratings for a given c: sum_i h(c * a_i), h(x) = #{j: x b_j >= 1} + x * sum_{j: x b_j < 1} b_j
__global__ void sy_expect(const float* a, uint32_t I, const double* b_sorted, const double* b_prefix, uint32_t J, double c, double* out)
{
    double s = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < I; i += gridDim.x * blockDim.x) {
        const double x = c * (double)a[i];
        // first index with x * b_sorted[j] < 1 (b_sorted descending)
        uint32_t lo = 0, hi = J;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (x * b_sorted[mid] >= 1.0) lo = mid + 1;
            else hi = mid;
        }
        const double tail = b_prefix[J] - b_prefix[lo];
        s += (double)lo + x * tail;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) out[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = s;   // summed in order on the host
}

__device__ __forceinline__ bool sy_included(uint32_t w, float p)
{
    if (p >= 1.0f) return true;
    return w < (uint32_t)(p * 4294967296.0f);
}

__device__ __forceinline__ bool sy_is_test(uint64_t seed, uint32_t i, uint32_t j, uint32_t test_thr)
{
    return philox_site(seed, SY_SPLIT, i, j, 0).x < test_thr;
}

// pass 1: per-user (train, test) counts
__global__ void sy_count(const float* a, const float* b, uint32_t I, uint32_t J, float c, uint64_t seed, uint32_t test_thr, uint32_t* n_tr,
                         uint32_t* n_te)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= I) return;
    const float ca = c * a[i];
    uint32_t tr = 0, te = 0;
    for (uint32_t j0 = 0; j0 < J; j0 += 4) {
        const uint4 w = philox_site(seed, SY_PAIR, i, j0 >> 2, 0);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t j = j0 + q;
            if (j < J && sy_included(ww[q], ca * b[j])) {
                if (sy_is_test(seed, i, j, test_thr)) ++te;
                else ++tr;
            }
        }
    }
    n_tr[i] = tr;
    n_te[i] = te;
}

__device__ __forceinline__ float sy_rating(uint64_t seed, uint32_t i, uint32_t j, const float* P, const float* Q, const float* ub, const float* vb)
{
    float r = 3.5f + ub[i] + vb[j];
#pragma unroll
    for (int d = 0; d < SY_RANK; ++d) r = fmaf(P[(size_t)i * SY_RANK + d], Q[(size_t)j * SY_RANK + d], r);
    const uint4 w = philox_site(seed, SY_SPLIT, i, j, 1);
    r += 0.8f * normal_f32(w);
    r = rintf(r * 2.0f) * 0.5f;
    return fminf(5.0f, fmaxf(0.5f, r));
}

// pass 2: fill, in (user, item) order
__global__ void sy_fill(const float* a, const float* b, uint32_t I, uint32_t J, float c, uint64_t seed, uint32_t test_thr, const uint64_t* off_tr,
                        const uint64_t* off_te, const float* P, const float* Q, const float* ub, const float* vb, uint32_t* tr_u, uint32_t* tr_i,
                        float* tr_r, uint32_t* te_u, uint32_t* te_i, float* te_r)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= I) return;
    const float ca = c * a[i];
    uint64_t ptr = off_tr[i], pte = off_te[i];
    for (uint32_t j0 = 0; j0 < J; j0 += 4) {
        const uint4 w = philox_site(seed, SY_PAIR, i, j0 >> 2, 0);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t j = j0 + q;
            if (j < J && sy_included(ww[q], ca * b[j])) {
                const float r = sy_rating(seed, i, j, P, Q, ub, vb);
                if (sy_is_test(seed, i, j, test_thr)) {
                    te_u[pte] = i; te_i[pte] = j; te_r[pte] = r; ++pte;
                } else {
                    tr_u[ptr] = i; tr_i[ptr] = j; tr_r[ptr] = r; ++ptr;
                }
            }
        }
    }
}

__global__ void sy_normals(float* out, uint64_t n, uint32_t width, uint64_t seed, uint32_t site, float stdev)
{
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x)
        out[t] = stdev * normal_f32(philox_site(seed, site, (uint32_t)(t / width), (uint32_t)(t % width), 0));
}

__global__ void sy_widen(const uint32_t* in, uint64_t* out, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

}  // namespace sbmf

using namespace sbmf;

static thread_local std::string g_synth_err;
extern "C" const char* sbmf_cuda_synth_last_error(void) { return g_synth_err.c_str(); }

#define SY_CK(call)                                                                                \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g_synth_err = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
            rc = (e_ == cudaErrorMemoryAllocation) ? SBMF_ERR_NOMEM : SBMF_ERR_CUDA;               \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)

extern "C" int sbmf_cuda_synth_generate(const sbmf_synth_spec* spec, uint64_t* n_train, uint64_t* n_test, uint32_t* train_user,
                                        uint32_t* train_item, float* train_rating, uint32_t* test_user, uint32_t* test_item, float* test_rating)
{
    if (!spec || !n_train || !n_test || spec->num_users == 0 || spec->num_items == 0 || spec->test_frac < 0.0 || spec->test_frac >= 1.0 ||
        (double)spec->n_ratings > 0.9 * (double)spec->num_users * (double)spec->num_items) {
        g_synth_err = "synth_generate: invalid spec";
        return SBMF_ERR_INVALID;
    }
    const bool fill = train_user != nullptr;
    if (fill && (!train_item || !train_rating || !test_user || !test_item || !test_rating)) {
        g_synth_err = "synth_generate: all six output arrays must be given together";
        return SBMF_ERR_INVALID;
    }
    const uint32_t I = spec->num_users, J = spec->num_items;
    const uint64_t seed = spec->seed;
    const int T = 256;
    int rc = SBMF_OK;
    uint32_t *d_keys = nullptr, *d_vals = nullptr, *d_keys2 = nullptr, *d_order = nullptr, *d_ntr = nullptr, *d_nte = nullptr;
    float *d_a = nullptr, *d_b = nullptr, *d_P = nullptr, *d_Q = nullptr, *d_ub = nullptr, *d_vb = nullptr;
    double *d_bs = nullptr, *d_bp = nullptr, *d_out = nullptr;
    uint64_t *d_wtr = nullptr, *d_wte = nullptr, *d_otr = nullptr, *d_ote = nullptr;
    uint32_t *d_tru = nullptr, *d_tri = nullptr, *d_teu = nullptr, *d_tei = nullptr;
    float *d_trr = nullptr, *d_ter = nullptr;
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0, tb2 = 0;
    const uint32_t nmax = I > J ? I : J;
    std::vector<double> bs((size_t)J), bp((size_t)J + 1);
    double c_lo = 0.0, c_hi = 1.0, c = 0.0;
    uint64_t tot_tr = 0, tot_te = 0;
    const uint32_t test_thr = (uint32_t)(spec->test_frac * 4294967296.0);

    SY_CK(cudaSetDevice(spec->device));
    SY_CK(cudaMalloc(&d_keys, (size_t)nmax * 4)); SY_CK(cudaMalloc(&d_vals, (size_t)nmax * 4));
    SY_CK(cudaMalloc(&d_keys2, (size_t)nmax * 4)); SY_CK(cudaMalloc(&d_order, (size_t)nmax * 4));
    SY_CK(cudaMalloc(&d_a, (size_t)I * 4)); SY_CK(cudaMalloc(&d_b, (size_t)J * 4));
    SY_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_order, (int)nmax));
    SY_CK(cub::DeviceScan::ExclusiveSum(nullptr, tb2, d_wtr, d_otr, (int)I));
    if (tb2 > tmp_bytes) tmp_bytes = tb2;
    SY_CK(cudaMalloc(&d_tmp, tmp_bytes));
    // random permutation of the ids -> Zipf weights
    sy_perm_keys<<<(I + T - 1) / T, T>>>(d_keys, d_vals, I, seed, SY_PERM_U);
    SY_CK(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_order, (int)I));
    sy_weights<<<(I + T - 1) / T, T>>>(d_order, d_a, I, spec->s_user);
    sy_perm_keys<<<(J + T - 1) / T, T>>>(d_keys, d_vals, J, seed, SY_PERM_V);
    SY_CK(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_order, (int)J));
    sy_weights<<<(J + T - 1) / T, T>>>(d_order, d_b, J, spec->s_item);

    // solve c: expected count == n_ratings
    bp[0] = 0.0;
    for (uint32_t j = 0; j < J; ++j) {
        bs[j] = (double)(float)pow((double)j + 1.0, -spec->s_item);
        bp[j + 1] = bp[j] + bs[j];
    }
    SY_CK(cudaMalloc(&d_bs, (size_t)J * 8)); SY_CK(cudaMalloc(&d_bp, ((size_t)J + 1) * 8)); SY_CK(cudaMalloc(&d_out, 592 * 8 * 8));
    SY_CK(cudaMemcpy(d_bs, bs.data(), (size_t)J * 8, cudaMemcpyHostToDevice));
    SY_CK(cudaMemcpy(d_bp, bp.data(), ((size_t)J + 1) * 8, cudaMemcpyHostToDevice));
    {
        auto expect = [&](double cc, double& val) -> cudaError_t {
            double part[592 * 8];
            sy_expect<<<592, T>>>(d_a, I, d_bs, d_bp, J, cc, d_out);
            cudaError_t e = cudaMemcpy(part, d_out, sizeof(part), cudaMemcpyDeviceToHost);
            val = 0.0;
            for (int q = 0; q < 592 * 8; ++q) val += part[q];
            return e;
        };
        double v = 0.0;
        for (int it = 0; it < 80; ++it) {
            SY_CK(expect(c_hi, v));
            if (v >= (double)spec->n_ratings) break;
            c_hi *= 4.0;
        }
        for (int it = 0; it < 60; ++it) {
            c = 0.5 * (c_lo + c_hi);
            SY_CK(expect(c, v));
            if (v < (double)spec->n_ratings) c_lo = c;
            else c_hi = c;
        }
        c = 0.5 * (c_lo + c_hi);
    }

    SY_CK(cudaMalloc(&d_ntr, (size_t)I * 4)); SY_CK(cudaMalloc(&d_nte, (size_t)I * 4));
    sy_count<<<(I + 127) / 128, 128>>>(d_a, d_b, I, J, (float)c, seed, test_thr, d_ntr, d_nte);
    SY_CK(cudaGetLastError());
    SY_CK(cudaMalloc(&d_wtr, (size_t)I * 8)); SY_CK(cudaMalloc(&d_wte, (size_t)I * 8));
    SY_CK(cudaMalloc(&d_otr, (size_t)I * 8)); SY_CK(cudaMalloc(&d_ote, (size_t)I * 8));
    sy_widen<<<(I + T - 1) / T, T>>>(d_ntr, d_wtr, I);
    sy_widen<<<(I + T - 1) / T, T>>>(d_nte, d_wte, I);
    SY_CK(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_wtr, d_otr, (int)I));
    SY_CK(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_wte, d_ote, (int)I));
    {
        uint64_t last_o[2], last_n[2];
        SY_CK(cudaMemcpy(&last_o[0], d_otr + (I - 1), 8, cudaMemcpyDeviceToHost));
        SY_CK(cudaMemcpy(&last_o[1], d_ote + (I - 1), 8, cudaMemcpyDeviceToHost));
        SY_CK(cudaMemcpy(&last_n[0], d_wtr + (I - 1), 8, cudaMemcpyDeviceToHost));
        SY_CK(cudaMemcpy(&last_n[1], d_wte + (I - 1), 8, cudaMemcpyDeviceToHost));
        tot_tr = last_o[0] + last_n[0];
        tot_te = last_o[1] + last_n[1];
    }
    if (fill && (*n_train != tot_tr || *n_test != tot_te)) {
        g_synth_err = "synth_generate: *n_train / *n_test do not match this spec (call with NULL arrays first)";
        rc = SBMF_ERR_INVALID;
        goto done;
    }
    *n_train = tot_tr;
    *n_test = tot_te;
    if (!fill) goto done;

    SY_CK(cudaMalloc(&d_P, (size_t)I * SY_RANK * 4)); SY_CK(cudaMalloc(&d_Q, (size_t)J * SY_RANK * 4));
    SY_CK(cudaMalloc(&d_ub, (size_t)I * 4)); SY_CK(cudaMalloc(&d_vb, (size_t)J * 4));
    sy_normals<<<1184, T>>>(d_P, (uint64_t)I * SY_RANK, SY_RANK, seed, SY_FACT_U, 0.3f);
    sy_normals<<<1184, T>>>(d_Q, (uint64_t)J * SY_RANK, SY_RANK, seed, SY_FACT_V, 0.3f);
    sy_normals<<<1184, T>>>(d_ub, I, 1, seed, SY_BIAS, 0.3f);
    sy_normals<<<1184, T>>>(d_vb, J, 1, seed, SY_BIAS + 1, 0.3f);
    SY_CK(cudaMalloc(&d_tru, (tot_tr + 1) * 4)); SY_CK(cudaMalloc(&d_tri, (tot_tr + 1) * 4)); SY_CK(cudaMalloc(&d_trr, (tot_tr + 1) * 4));
    SY_CK(cudaMalloc(&d_teu, (tot_te + 1) * 4)); SY_CK(cudaMalloc(&d_tei, (tot_te + 1) * 4)); SY_CK(cudaMalloc(&d_ter, (tot_te + 1) * 4));
    sy_fill<<<(I + 127) / 128, 128>>>(d_a, d_b, I, J, (float)c, seed, test_thr, d_otr, d_ote, d_P, d_Q, d_ub, d_vb, d_tru, d_tri, d_trr, d_teu, d_tei,
                                      d_ter);
    SY_CK(cudaGetLastError());
    SY_CK(cudaMemcpy(train_user, d_tru, tot_tr * 4, cudaMemcpyDeviceToHost));
    SY_CK(cudaMemcpy(train_item, d_tri, tot_tr * 4, cudaMemcpyDeviceToHost));
    SY_CK(cudaMemcpy(train_rating, d_trr, tot_tr * 4, cudaMemcpyDeviceToHost));
    SY_CK(cudaMemcpy(test_user, d_teu, tot_te * 4, cudaMemcpyDeviceToHost));
    SY_CK(cudaMemcpy(test_item, d_tei, tot_te * 4, cudaMemcpyDeviceToHost));
    SY_CK(cudaMemcpy(test_rating, d_ter, tot_te * 4, cudaMemcpyDeviceToHost));

done:
    cudaFree(d_keys); cudaFree(d_vals); cudaFree(d_keys2); cudaFree(d_order); cudaFree(d_ntr); cudaFree(d_nte);
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_P); cudaFree(d_Q); cudaFree(d_ub); cudaFree(d_vb);
    cudaFree(d_bs); cudaFree(d_bp); cudaFree(d_out); cudaFree(d_wtr); cudaFree(d_wte); cudaFree(d_otr); cudaFree(d_ote);
    cudaFree(d_tru); cudaFree(d_tri); cudaFree(d_teu); cudaFree(d_tei); cudaFree(d_trr); cudaFree(d_ter); cudaFree(d_tmp);
    return rc;
}
