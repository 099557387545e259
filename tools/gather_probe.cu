// tools/gather_probe.cu -- developer microbenchmark: what bounds random sector gathers from an L2-resident table on B200?
// Each "rating" needs SPAN consecutive 32-byte sectors of one random table row; SPAN lanes cooperate on one rating
// (lane q of the group loads sector q with one LDG.E.256), so a warp-wide load touches 32/SPAN distinct 128-byte lines.
// Same bytes per warp load in every variant; only the number of lines (L1TEX tag-stage wavefronts) changes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe tools/gather_probe.cu && ./gather_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

struct f8 { float v[8]; };
__device__ __forceinline__ f8 ld256_nc(const float* p)
{
    f8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=&f"(r.v[0]), "=&f"(r.v[1]), "=&f"(r.v[2]), "=&f"(r.v[3]), "=&f"(r.v[4]), "=&f"(r.v[5]), "=&f"(r.v[6]), "=&f"(r.v[7])
                 : "l"(p));
    return r;
}

// table: rows x (SPAN*8) floats; idx: n random row ids (one per group of SPAN lanes per iteration)
template <int SPAN, int UNR, int MAP = 0>
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ table, const uint32_t* __restrict__ idx, uint64_t n_groups,
                                                     float* __restrict__ out)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nth = (uint64_t)gridDim.x * blockDim.x;
    // MAP 0: lanes q = lane % SPAN of one rating are adjacent; MAP 1: sector-major (lanes [q*32/SPAN, (q+1)*32/SPAN) load sector q)
    const int lane = threadIdx.x & 31;
    const int q = (MAP == 1) ? lane / (32 / SPAN) : lane % SPAN;
    const int rsub = (MAP == 1) ? lane % (32 / SPAN) : lane / SPAN;
    float acc = 0.f;
    const uint64_t warp_id = tid >> 5;
    const uint64_t g0 = warp_id * (32 / SPAN) + rsub, gstep = nth / SPAN;
    for (uint64_t g = g0; g < n_groups; g += gstep * UNR) {
        uint32_t id[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) id[u] = (g + u * gstep < n_groups) ? idx[g + u * gstep] : 0u;
        f8 f[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const float* p = table + ((size_t)id[u] * SPAN + q) * 8;
            if (MAP == 2) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
                f[u].v[0] = a.x; f[u].v[1] = a.y; f[u].v[2] = a.z; f[u].v[3] = a.w; f[u].v[4] = b.x; f[u].v[5] = b.y; f[u].v[6] = b.z; f[u].v[7] = b.w;
            } else {
                f[u] = ld256_nc(p);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc += f[u].v[k];
    }
    if (acc == 123.456f) out[tid] = acc;
}

// every lane loads BOTH sectors of a random 64-byte row with two LDG.E.256 (32 lines per instruction, 2 instructions per row)
template <int UNR>
__global__ void __launch_bounds__(256) gather2_perlane_kernel(const float* __restrict__ table, const uint32_t* __restrict__ idx, uint64_t n_groups,
                                                              float* __restrict__ out)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nth = (uint64_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (uint64_t g = tid; g < n_groups; g += nth * UNR) {
        uint32_t id[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) id[u] = (g + u * nth < n_groups) ? idx[g + u * nth] : 0u;
        f8 f[UNR], h[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            f[u] = ld256_nc(table + (size_t)id[u] * 16);
            h[u] = ld256_nc(table + (size_t)id[u] * 16 + 8);
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc += f[u].v[k] + h[u].v[k];
    }
    if (acc == 123.456f) out[tid] = acc;
}

template <int SPAN, int UNR, int MAP = 0>
static void run(const char* name, const float* table, const uint32_t* idx, uint64_t n_sectors, float* out, int sms, double ghz)
{
    const uint64_t n_groups = n_sectors / SPAN;
    const int grid = sms * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 2; ++w) gather_kernel<SPAN, UNR, MAP><<<grid, 256>>>(table, idx, n_groups, out);
    cudaEventRecord(a);
    const int reps = 5;
    for (int r = 0; r < reps; ++r) gather_kernel<SPAN, UNR, MAP><<<grid, 256>>>(table, idx, n_groups, out);
    cudaEventRecord(b);
    cudaError_t err = cudaEventSynchronize(b);
    if (err != cudaSuccess || (err = cudaGetLastError()) != cudaSuccess) {
        printf("CUDA error in %s: %s\n", name, cudaGetErrorString(err));
        exit(1);
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    ms /= reps;
    const double sect_per_s = (double)n_sectors / (ms * 1e-3);
    printf("%-28s %8.3f ms  %7.1f G sectors/s  %6.2f TB/s  %5.2f sectors/clk/SM (at %.3f GHz)\n", name, ms, sect_per_s / 1e9, sect_per_s * 32 / 1e12,
           sect_per_s / (sms * ghz * 1e9), ghz);
}

int main(int argc, char** argv)
{
    const size_t table_mb = argc > 1 ? atol(argv[1]) : 15;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const double ghz = p.clockRate * 1e-6;
    const size_t rows128 = table_mb * (1u << 20) / 128;   // table as 128-byte lines
    const uint64_t n_sectors = 128ull << 20;              // 4 GB of gathers per launch
    float* table;
    uint32_t* idx;
    float* out;
    cudaMalloc(&table, rows128 * 128);
    cudaMemset(table, 0, rows128 * 128);
    cudaMalloc(&idx, n_sectors * 4);
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    uint32_t* h = (uint32_t*)malloc(n_sectors * 4);
    uint64_t s = 88172645463325252ull;
    printf("%s: %d SMs, %.3f GHz, table %zu MB\n", p.name, sms, ghz, table_mb);
    const int only = argc > 2 ? atoi(argv[2]) : 0;
    for (int span = 1; span <= 4; span *= 2) {
        if (only && span != only) continue;
        const uint64_t rows = rows128 * 4 / span;   // rows of span sectors
        for (uint64_t i = 0; i < n_sectors / span; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            h[i] = (uint32_t)(s % rows);
        }
        cudaMemcpy(idx, h, n_sectors / span * 4, cudaMemcpyHostToDevice);
        if (span == 1) {
            run<1, 2>("1 sector/rating  unroll 2", table, idx, n_sectors, out, sms, ghz);
            run<1, 4>("1 sector/rating  unroll 4", table, idx, n_sectors, out, sms, ghz);
            run<1, 8>("1 sector/rating  unroll 8", table, idx, n_sectors, out, sms, ghz);
            run<1, 4, 2>("1 sect, LDG.128x2, unr 4", table, idx, n_sectors, out, sms, ghz);
        } else if (span == 2) {
            run<2, 2, 2>("2 sect, LDG.128x2, unr 2", table, idx, n_sectors, out, sms, ghz);
            run<2, 4, 2>("2 sect, LDG.128x2, unr 4", table, idx, n_sectors, out, sms, ghz);
            run<2, 4, 1>("2 sect, sector-major, unr 4", table, idx, n_sectors, out, sms, ghz);
            run<2, 2>("2 sectors/rating unroll 2", table, idx, n_sectors, out, sms, ghz);
            run<2, 4>("2 sectors/rating unroll 4", table, idx, n_sectors, out, sms, ghz);
            run<2, 8>("2 sectors/rating unroll 8", table, idx, n_sectors, out, sms, ghz);
            {
                cudaEvent_t a, b;
                cudaEventCreate(&a); cudaEventCreate(&b);
                gather2_perlane_kernel<2><<<sms * 8, 256>>>(table, idx, n_sectors / 2, out);
                cudaEventRecord(a);
                for (int r = 0; r < 5; ++r) gather2_perlane_kernel<2><<<sms * 8, 256>>>(table, idx, n_sectors / 2, out);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms = 0.f;
                cudaEventElapsedTime(&ms, a, b);
                ms /= 5;
                printf("%-28s %8.3f ms  %7.1f G sectors/s  %5.2f sectors/clk/SM\n", "64 B/lane as 2 loads, unr 2", ms, n_sectors / (ms * 1e-3) / 1e9,
                       n_sectors / (ms * 1e-3) / (sms * ghz * 1e9));
            }
        } else {
            run<4, 4, 2>("4 sect, LDG.128x2, unr 4", table, idx, n_sectors, out, sms, ghz);
            run<4, 4, 1>("4 sect, sector-major, unr 4", table, idx, n_sectors, out, sms, ghz);
            run<4, 2>("4 sectors/rating unroll 2", table, idx, n_sectors, out, sms, ghz);
            run<4, 4>("4 sectors/rating unroll 4", table, idx, n_sectors, out, sms, ghz);
            run<4, 8>("4 sectors/rating unroll 8", table, idx, n_sectors, out, sms, ghz);
        }
    }
    return 0;
}
