// tools/tma_gather_probe.cu -- developer microbenchmark (round 2): can the bulk-copy engine (cp.async.bulk, the descriptor-less
// TMA path) gather 32-byte factor rows faster than LDG does?  ncu of the resident-row kernels shows their limiter is the L1TEX
// data pipe: a scattered 32-byte sector costs one wavefront there (profiles/r2/ncu_rows_rk1.txt: 353 M wavefronts = 225 M gathered
// sectors + 115 M shuffle / shared-memory wavefronts at 72 % pipe utilisation).  A bulk copy lands in shared memory without
// passing through that pipe, and reading it back costs 8 wavefronts per 32 rows instead of 32.  The question is the engine's
// rate for 32-byte requests.
//   form A: every lane fetches its row with one LDG.E.256                (the kernels' form; L2-resident 15 MB table, random ids)
//   form B: every lane issues one 32-byte cp.async.bulk into a shared-memory ring (STAGES deep, UNR rows per lane and stage,
//           mbarrier complete_tx), then reads its rows back with two LDS.128 (48-byte row stride: conflict-free)
//   form C: like B without the read-back (the engine alone)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/tma_gather_probe tools/tma_gather_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

struct __align__(32) f8 { float v[8]; };
__device__ __forceinline__ f8 ld256_nc(const float* p)
{
    f8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}

template <int UNR>
__global__ void __launch_bounds__(128) ldg_gather_kernel(const float* __restrict__ table, const uint32_t* __restrict__ idx, uint64_t n, float* __restrict__ out)
{
    const uint64_t per_cta = (n + gridDim.x - 1) / gridDim.x;
    const uint64_t beg = (uint64_t)blockIdx.x * per_cta, end = beg + per_cta < n ? beg + per_cta : n;
    float acc = 0.f;
    for (uint64_t g = beg + threadIdx.x; g < end; g += 128ull * UNR) {
        uint32_t id[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) id[u] = (g + 128ull * u < end) ? idx[g + 128ull * u] : 0u;
        f8 f[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) f[u] = ld256_nc(table + (size_t)id[u] * 8);
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc += f[u].v[k];
    }
    if (acc == 123.456f) out[blockIdx.x * 128 + threadIdx.x] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s_32(void* dst, const void* src, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(smem_u32(bar))
                 : "memory");
}

constexpr int ROW_STRIDE = 48;   // bytes between rows in the ring: lanes r..r+7 of an LDS.128 phase fall on distinct bank groups

// one "batch" = 128 threads x UNR rows.  Stage s of the ring holds one batch.
template <int UNR, int STAGES, bool READBACK>
__global__ void __launch_bounds__(128) bulk_gather_kernel(const float* __restrict__ table, const uint32_t* __restrict__ idx, uint64_t n, float* __restrict__ out)
{
    extern __shared__ __align__(128) unsigned char ring[];   // [STAGES][UNR][128] rows of ROW_STRIDE bytes
    __shared__ __align__(8) uint64_t full[STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t per_cta = ((n + gridDim.x - 1) / gridDim.x + 128ull * UNR - 1) / (128ull * UNR) * (128ull * UNR);   // whole batches
    const uint64_t beg = (uint64_t)blockIdx.x * per_cta;
    const uint64_t end = beg + per_cta < n ? beg + per_cta : (n / (128ull * UNR)) * (128ull * UNR);
    const int nb = beg < end ? (int)((end - beg) / (128ull * UNR)) : 0;
    constexpr uint32_t STAGE_BYTES = UNR * 128 * ROW_STRIDE;
    auto issue = [&](int batch) {
        const int s = batch % STAGES;
        if (tid == 0) mbar_expect_tx(&full[s], UNR * 128 * 32);
        const uint64_t g = beg + (uint64_t)batch * 128 * UNR + tid;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const uint32_t id = idx[g + 128ull * u];
            bulk_g2s_32(ring + (size_t)s * STAGE_BYTES + (size_t)(u * 128 + tid) * ROW_STRIDE, table + (size_t)id * 8, &full[s]);
        }
    };
    for (int b = 0; b < STAGES - 1 && b < nb; ++b) issue(b);
    float acc = 0.f;
    for (int b = 0; b < nb; ++b) {
        if (b + STAGES - 1 < nb) issue(b + STAGES - 1);   // refills the stage consumed in iteration b - 1 (barrier below)
        const int s = b % STAGES;
        mbar_wait(&full[s], (uint32_t)((b / STAGES) & 1));
        if (READBACK) {
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const float4* row = reinterpret_cast<const float4*>(ring + (size_t)s * STAGE_BYTES + (size_t)(u * 128 + tid) * ROW_STRIDE);
                const float4 a = row[0], c = row[1];
                acc += a.x + a.y + a.z + a.w + c.x + c.y + c.z + c.w;
            }
        }
        __syncthreads();   // everybody is done with stage s before it is refilled
    }
    if (acc == 123.456f) out[blockIdx.x * 128 + threadIdx.x] = acc;
}

template <class F>
static float best_ms(F&& launch)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    launch();
    launch();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float t;
        cudaEventElapsedTime(&t, a, b);
        best = std::min(best, t);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(e));
        exit(1);
    }
    return best;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    const size_t table_bytes = 15ull << 20;
    const uint32_t rows = (uint32_t)(table_bytes / 32);
    const uint64_t n = 64ull << 20;
    float *table, *out;
    uint32_t* idx;
    cudaMalloc(&table, table_bytes);
    cudaMemset(table, 0, table_bytes);
    cudaMalloc(&out, (size_t)sms * 16 * 128 * 4);
    std::vector<uint32_t> h(n);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (uint64_t i = 0; i < n; ++i) {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        h[i] = (uint32_t)(s % rows);
    }
    cudaMalloc(&idx, n * 4);
    cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);
    printf("%s: %d SMs, %.3f GHz; %llu gathers of 32 B from a %zu MB table (L2-resident), random ids\n", prop.name, sms, ghz, (unsigned long long)n, table_bytes >> 20);
    auto report = [&](const char* name, float ms) {
        printf("%-72s %8.3f ms  %7.1f G rows/s  %5.2f rows/clk/SM\n", name, ms, n / (ms * 1e-3) / 1e9, n / (ms * 1e-3) / (sms * ghz * 1e9));
    };
    report("A  LDG.E.256 per lane, 4 in flight per lane, 16 CTAs/SM x 128", best_ms([&] { ldg_gather_kernel<4><<<sms * 16, 128>>>(table, idx, n, out); }));
    report("A  LDG.E.256 per lane, 8 in flight per lane", best_ms([&] { ldg_gather_kernel<8><<<sms * 16, 128>>>(table, idx, n, out); }));
#define BULK(UNR, STAGES, RB, CTAS, label)                                                                                            \
    do {                                                                                                                              \
        const size_t smem = (size_t)STAGES * UNR * 128 * ROW_STRIDE;                                                                   \
        cudaFuncSetAttribute(bulk_gather_kernel<UNR, STAGES, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
        report(label, best_ms([&] { bulk_gather_kernel<UNR, STAGES, RB><<<sms * CTAS, 128, smem>>>(table, idx, n, out); }));          \
    } while (0)
    BULK(2, 4, true, 4, "B  cp.async.bulk 32 B per lane + LDS read-back, 2 rows x 4 stages, 4 CTAs/SM");
    BULK(4, 4, true, 2, "B  cp.async.bulk 32 B per lane + LDS read-back, 4 rows x 4 stages, 2 CTAs/SM");
    BULK(4, 4, true, 4, "B  cp.async.bulk 32 B per lane + LDS read-back, 4 rows x 4 stages, 4 CTAs/SM");
    BULK(4, 8, true, 1, "B  cp.async.bulk 32 B per lane + LDS read-back, 4 rows x 8 stages, 1 CTA/SM");
    BULK(4, 4, false, 4, "C  cp.async.bulk 32 B per lane, no read-back, 4 rows x 4 stages, 4 CTAs/SM");
    BULK(8, 4, false, 2, "C  cp.async.bulk 32 B per lane, no read-back, 8 rows x 4 stages, 2 CTAs/SM");
    return 0;
}
