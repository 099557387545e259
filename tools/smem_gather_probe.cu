// tools/smem_gather_probe.cu -- developer microbenchmark for DESIGN.md 10, item 4: how fast are 32-byte row gathers when the
// table tile lives in SHARED memory instead of L2?  (tools/gather_probe.cu measured the L2 form: 0.89 sectors/clk/SM.)
// Every CTA stages a tile of ROWS x 32 B (96 KB by default, like one factor block of ~3,000 users), then streams a list of
// row ids (4 B each, coalesced) and reads each id's 32 bytes with two LDS.128.
//   layout 0: rows contiguous (row r at byte 32 r)            -- both halves of a row share a 32-byte bank group
//   layout 1: half h of row r at 16-byte slot (2r + h) ^ ((r >> 2) & 1)  -- spreads the lower halves over all eight bank groups
//   ids: random within the tile, or ascending with random gaps (the order of a CSC segment restricted to a user range)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/smem_gather_probe tools/smem_gather_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

template <int LAYOUT, int UNR>
__global__ void __launch_bounds__(256) smem_gather_kernel(const float4* __restrict__ table, int rows, const uint32_t* __restrict__ idx,
                                                          uint64_t n, float* __restrict__ out)
{
    extern __shared__ float4 tile[];   // [rows * 2] 16-byte slots
    for (int s = threadIdx.x; s < rows * 2; s += blockDim.x) {
        const int r = s >> 1, h = s & 1;
        const int slot = LAYOUT == 0 ? s : ((2 * r + h) ^ ((r >> 2) & 1));
        tile[slot] = table[s];
    }
    __syncthreads();
    const uint64_t per_cta = (n + gridDim.x - 1) / gridDim.x;
    const uint64_t beg = (uint64_t)blockIdx.x * per_cta, end = beg + per_cta < n ? beg + per_cta : n;
    float acc = 0.f;
    for (uint64_t g = beg + threadIdx.x; g < end; g += (uint64_t)blockDim.x * UNR) {
        uint32_t id[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) id[u] = (g + (uint64_t)u * blockDim.x < end) ? idx[g + (uint64_t)u * blockDim.x] : 0u;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int r = (int)id[u];
            const int s0 = LAYOUT == 0 ? 2 * r : ((2 * r) ^ ((r >> 2) & 1));
            const int s1 = LAYOUT == 0 ? 2 * r + 1 : ((2 * r + 1) ^ ((r >> 2) & 1));
            const float4 a = tile[s0], b = tile[s1];
            acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
        }
    }
    if (acc == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int LAYOUT, int UNR>
static void run(const char* name, const float4* table, int rows, const uint32_t* idx, uint64_t n, float* out, int sms, double ghz, int ctas_per_sm)
{
    const size_t smem = (size_t)rows * 32;
    cudaFuncSetAttribute(smem_gather_kernel<LAYOUT, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = sms * ctas_per_sm;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int w = 0; w < 2; ++w) smem_gather_kernel<LAYOUT, UNR><<<grid, 256, smem>>>(table, rows, idx, n, out);
    cudaEventRecord(a);
    const int reps = 5;
    for (int r = 0; r < reps; ++r) smem_gather_kernel<LAYOUT, UNR><<<grid, 256, smem>>>(table, rows, idx, n, out);
    cudaEventRecord(b);
    cudaError_t err = cudaEventSynchronize(b);
    if (err != cudaSuccess || (err = cudaGetLastError()) != cudaSuccess) {
        printf("CUDA error in %s: %s\n", name, cudaGetErrorString(err));
        exit(1);
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    ms /= reps;
    const double per_s = (double)n / (ms * 1e-3);
    printf("%-44s %8.3f ms  %7.1f G rows/s  %5.2f rows/clk/SM (L2 form: 0.89)\n", name, ms, per_s / 1e9, per_s / (sms * ghz * 1e9));
}

int main(int argc, char** argv)
{
    const int rows = argc > 1 ? atoi(argv[1]) : 3072;          // 96 KB tile
    const int ctas_per_sm = argc > 2 ? atoi(argv[2]) : 2;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const double ghz = p.clockRate * 1e-6;
    const uint64_t n = 256ull << 20;                           // gathers per launch
    float4* table;
    uint32_t* idx;
    float* out;
    cudaMalloc(&table, (size_t)rows * 32);
    cudaMemset(table, 0, (size_t)rows * 32);
    cudaMalloc(&idx, n * 4);
    cudaMalloc(&out, (size_t)sms * ctas_per_sm * 256 * 4);
    std::vector<uint32_t> h(n);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    printf("%s: %d SMs, %.3f GHz, tile %d rows = %d KB, %d CTAs/SM\n", p.name, sms, ghz, rows, rows * 32 / 1024, ctas_per_sm);
    for (int pattern = 0; pattern < 2; ++pattern) {
        if (pattern == 0) {
            for (uint64_t i = 0; i < n; ++i) h[i] = (uint32_t)(rnd() % (uint64_t)rows);
        } else {   // ascending runs: each run walks the tile once with random gaps (mean gap 8 rows), like one item's users in a range
            uint32_t r = 0;
            for (uint64_t i = 0; i < n; ++i) {
                r += 1 + (uint32_t)(rnd() % 15);
                if (r >= (uint32_t)rows) r = (uint32_t)(rnd() % 8);
                h[i] = r;
            }
        }
        cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);
        const char* pn = pattern == 0 ? "random ids" : "ascending runs";
        char name[96];
        snprintf(name, sizeof(name), "%s, contiguous rows, unroll 2", pn);
        run<0, 2>(name, table, rows, idx, n, out, sms, ghz, ctas_per_sm);
        snprintf(name, sizeof(name), "%s, contiguous rows, unroll 4", pn);
        run<0, 4>(name, table, rows, idx, n, out, sms, ghz, ctas_per_sm);
        snprintf(name, sizeof(name), "%s, swizzled halves, unroll 2", pn);
        run<1, 2>(name, table, rows, idx, n, out, sms, ghz, ctas_per_sm);
        snprintf(name, sizeof(name), "%s, swizzled halves, unroll 4", pn);
        run<1, 4>(name, table, rows, idx, n, out, sms, ghz, ctas_per_sm);
    }
    return 0;
}
