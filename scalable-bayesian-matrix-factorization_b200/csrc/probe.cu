// probe.cu -- sbmf_cuda_probe: the two hardware rates the sweep kernels run against, measured on the device the job is
// running on, in the same process, so that bench.py's roofline denominators are not constants copied from another box:
//   * HBM streaming bandwidth (copy: read + write bytes) over buffers far larger than L2;
//   * the L2 -> SM return path for row gathers from an L2-resident table, for every ACCESS FORM the kernels use or could use.
//     The factor gathers of the row kernels are "one random 32-byte sector per lane, one LDG.E.256" (form G32_LANE).  The other
//     forms fetch the same bytes as wider contiguous pieces (64 B = two K8 blocks side by side, 128 B = a full line), per lane
//     or by 2 / 4 cooperating lanes; they bound what a different factor layout could reach.
// Row ids stream from HBM (4 bytes per gather, coalesced) exactly as the index stream of a rating slice does.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/sbmf_cuda.h"
#include "common.cuh"
#include "launch.h"

namespace sbmf {

// SPAN sectors per row.  COOP = 1: every lane fetches a whole row with SPAN LDG.E.256; COOP = SPAN: SPAN adjacent lanes fetch one
// row, one sector each (a quarter-warp pass of the 256-bit load then touches 8 / SPAN lines instead of 8).
template <int SPAN, int COOP, int UNR>
__global__ void __launch_bounds__(256) probe_gather_kernel(const float* __restrict__ table, const uint32_t* __restrict__ idx, uint64_t n_rows,
                                                           float* __restrict__ out)
{
    static_assert(COOP == 1 || COOP == SPAN, "per-lane or fully cooperative");
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nth = (uint64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int q = (COOP == 1) ? 0 : lane % COOP;
    const uint64_t g0 = (COOP == 1) ? tid : (tid >> 5) * (32 / COOP) + (uint64_t)(lane / COOP);
    const uint64_t gstep = nth / COOP;
    float acc = 0.f;
    for (uint64_t g = g0; g < n_rows; g += gstep * UNR) {
        uint32_t id[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) id[u] = (g + u * gstep < n_rows) ? idx[g + u * gstep] : 0u;
        constexpr int PER = (COOP == 1) ? SPAN : 1;
        f8 f[UNR][PER];
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int s = 0; s < PER; ++s) f[u][s] = ld256_nc(table + ((size_t)id[u] * SPAN + q + s) * 8);
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int s = 0; s < PER; ++s)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += f[u][s].v[k];
    }
    if (acc == 123.456f) out[tid] = acc;
}

__global__ void __launch_bounds__(256) probe_copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) probe_read_kernel(const float4* __restrict__ src, uint64_t n, float* __restrict__ out)
{
    float acc = 0.f;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        const float4 v = src[i];
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) out[threadIdx.x] = acc;
}

__global__ void probe_fill_idx_kernel(uint32_t* idx, uint64_t n, uint32_t rows, uint64_t seed)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 x = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0x9e37u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        idx[i] = x.x % rows;
    }
}

namespace {
thread_local std::string g_probe_err;

template <class F>
cudaError_t time_ms(F&& launch, int reps, float* ms)
{
    cudaEvent_t a, b;
    cudaError_t e;
    if ((e = cudaEventCreate(&a)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&b)) != cudaSuccess) return e;
    launch();   // warm-up (also pulls the table into L2)
    launch();
    // best of `reps` single launches: the figure is used as a ceiling
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a, 0);
        launch();
        cudaEventRecord(b, 0);
        if ((e = cudaEventSynchronize(b)) != cudaSuccess) break;
        float t = 0.f;
        cudaEventElapsedTime(&t, a, b);
        best = std::min(best, t);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *ms = best;
    return e;
}
}  // namespace

}  // namespace sbmf

using namespace sbmf;

extern "C" const char* sbmf_cuda_probe_last_error(void) { return g_probe_err.c_str(); }

extern "C" int sbmf_cuda_probe(int device, uint64_t table_bytes, uint64_t n_gathers, sbmf_probe_result* out)
{
    if (!out) return SBMF_ERR_INVALID;
    *out = sbmf_probe_result{};
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        g_probe_err = std::string("probe: ") + cudaGetErrorString(cudaGetLastError());
        return SBMF_ERR_CUDA;
    }
    if (table_bytes == 0) table_bytes = 15ull << 20;      // one K8 factor block of the Netflix-shaped user side
    if (n_gathers == 0) n_gathers = 64ull << 20;
    table_bytes = (table_bytes + 127) / 128 * 128;
    const int sms = prop.multiProcessorCount;
    out->sm_count = sms;
    out->table_bytes = table_bytes;
    out->n_gathers = n_gathers;
    const size_t copy_elems = (1ull << 30) / 16;          // 1 GiB read + 1 GiB written
    float *table = nullptr, *sink = nullptr;
    uint32_t* idx = nullptr;
    float4 *ca = nullptr, *cb = nullptr;
    auto cleanup = [&]() {
        cudaFree(table); cudaFree(sink); cudaFree(idx); cudaFree(ca); cudaFree(cb);
    };
#define PCK(call)                                                                       \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            g_probe_err = std::string("probe: " #call ": ") + cudaGetErrorString(e_);   \
            cleanup();                                                                  \
            return (e_ == cudaErrorMemoryAllocation) ? SBMF_ERR_NOMEM : SBMF_ERR_CUDA;  \
        }                                                                               \
    } while (0)
    const int grid = sms * 8;
    PCK(cudaMalloc((void**)&table, table_bytes));
    PCK(cudaMemset(table, 0, table_bytes));
    PCK(cudaMalloc((void**)&sink, (size_t)grid * 256 * 4));
    PCK(cudaMalloc((void**)&idx, n_gathers * 4));
    PCK(cudaMalloc((void**)&ca, copy_elems * 16));
    PCK(cudaMalloc((void**)&cb, copy_elems * 16));
    PCK(cudaMemset(ca, 0, copy_elems * 16));
    float ms = 0.f;
    // ---- HBM
    PCK(time_ms([&]() { SBMF_LAUNCH((probe_copy_kernel), sms * 16, 256, 0, 0, ca, cb, copy_elems); }, 5, &ms));
    out->hbm_copy_gbs = 2.0 * copy_elems * 16 / (ms * 1e-3) / 1e9;
    PCK(time_ms([&]() { SBMF_LAUNCH((probe_read_kernel), sms * 16, 256, 0, 0, ca, copy_elems, sink); }, 5, &ms));
    out->hbm_read_gbs = 1.0 * copy_elems * 16 / (ms * 1e-3) / 1e9;
    // ---- gathers: same number of SECTORS per launch in every form
    auto gather = [&](int form, int span, auto kernel) -> int {
        const uint32_t rows = (uint32_t)(table_bytes / (32 * (size_t)span));
        const uint64_t n_rows = n_gathers / (uint64_t)span;
        SBMF_LAUNCH((probe_fill_idx_kernel), sms * 8, 256, 0, 0, idx, n_rows, rows, 0x5bd1e995ull + (uint64_t)form);
        float t = 0.f;
        cudaError_t e = time_ms([&]() { SBMF_LAUNCH((kernel), grid, 256, 0, 0, table, idx, n_rows, sink); }, 5, &t);
        if (e != cudaSuccess) {
            g_probe_err = std::string("probe: gather form ") + std::to_string(form) + ": " + cudaGetErrorString(e);
            return -1;
        }
        out->gather_sectors_per_s[form] = (double)n_rows * span / (t * 1e-3);
        return 0;
    };
    int rc = 0;
    rc |= gather(SBMF_PROBE_G32_LANE, 1, probe_gather_kernel<1, 1, 4>);
    rc |= gather(SBMF_PROBE_G64_LANE, 2, probe_gather_kernel<2, 1, 2>);
    rc |= gather(SBMF_PROBE_G64_COOP2, 2, probe_gather_kernel<2, 2, 4>);
    rc |= gather(SBMF_PROBE_G128_LANE, 4, probe_gather_kernel<4, 1, 2>);
    rc |= gather(SBMF_PROBE_G128_COOP4, 4, probe_gather_kernel<4, 4, 4>);
    cleanup();
#undef PCK
    if (rc) return SBMF_ERR_CUDA;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    out->sm_clock_mhz_max = khz * 1e-3;
    return SBMF_OK;
}
