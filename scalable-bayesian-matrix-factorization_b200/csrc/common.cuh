// common.cuh -- device helpers shared by every kernel of libsbmf_cuda: counter-based Philox streams,
// the samplers that replace the reference's rand()-based src/util/random.h ("[R]"), 256-bit gathers,
// warp reductions.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sbmf {

// Philox stream ids (counter word 2).  The parity tests' CPU checker mirrors these values -- keep them stable.
enum Site : uint32_t {
    SITE_INIT_U = 0, SITE_INIT_V = 1, SITE_U = 2, SITE_V = 3, SITE_BI = 4, SITE_BJ = 5,
    SITE_MU_BI = 6, SITE_MU_BJ = 7, SITE_SIGMA_BI = 8, SITE_SIGMA_BJ = 9,
    SITE_SIGMA_U = 10, SITE_MU_U = 11, SITE_SIGMA_V = 12, SITE_MU_V = 13,
    SITE_ALPHA = 14, SITE_SIGMA_B0 = 15, SITE_MU_B0 = 16, SITE_B0 = 17
};

enum SampleMode : int { SAMPLE_REF = 0, SAMPLE_SQRT = 1, SAMPLE_ZERO = 2 };

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011).  counter = (row, c1, site, sweep), key = seed.
// A draw is a pure function of (seed, site, row, c1, sweep): independent of which kernel path, warp
// or GPU handles the row, which is what makes 1-GPU and G-GPU runs comparable.
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__host__ __device__ __forceinline__ uint4 philox_site(uint64_t seed, uint32_t site, uint32_t row, uint32_t c1, uint32_t sweep)
{
    return philox4x32_10(make_uint4(row, c1, site, sweep), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// standard normal, fp32 Box-Muller on words 0/1 -- used by the O(|Omega|K) factor / bias draws
__device__ __forceinline__ float normal_f32(uint4 x)
{
    const float u1 = ((float)(x.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((float)(x.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
}

// fp64 variant for the O(K + I + J) hyper-parameter draws
__device__ __forceinline__ double normal_f64(uint4 x)
{
    const double u1 = ((double)x.x + 0.5) * (1.0 / 4294967296.0);
    const double u2 = ((double)x.y + 0.5) * (1.0 / 4294967296.0);
    return sqrt(-2.0 * log(u1)) * cos(6.28318530717958647692 * u2);
}

// ran_gaussian(mean, stdev) of [R]:166-172, including its "stdev == 0 or NaN => mean" guard.
__device__ __forceinline__ double draw_gauss_f64(int mode, uint64_t seed, uint32_t site, uint32_t row, uint32_t c1,
                                                  uint32_t sweep, double mean, double var)
{
    if (mode == SAMPLE_ZERO) return mean;
    const double sd = (mode == SAMPLE_SQRT) ? sqrt(var) : var;   // [T] passes the variance as stdev
    if (sd == 0.0 || isnan(sd)) return mean;
    return mean + sd * normal_f64(philox_site(seed, site, row, c1, sweep));
}

// ran_gamma(shape, rate) of [R]:118-148.  shape >= 1 (true at every call site of [T] with its default priors): Marsaglia-Tsang;
// attempt a consumes counter (row, a, site, sweep): normal from words 0/1, uniform from word 2.  shape < 1 ([R]:120-125):
// Gamma(shape + 1) * u^(1/shape), u from word 3 of the counter (row, 0xffffffff, site, sweep).  The rejection loop is capped:
// Marsaglia-Tsang accepts > 95 % of the attempts, so 64 rejections in a row mean NaN / non-positive arguments (priors are
// validated at create, but a diverged chain can still produce them) -- the draw then degrades to the Gamma mean instead of
// spinning forever in a single-thread kernel.
constexpr uint32_t GAMMA_MAX_ATTEMPTS = 64;
__device__ __forceinline__ double draw_gamma_f64(int mode, uint64_t seed, uint32_t site, uint32_t row, uint32_t sweep,
                                                  double shape, double rate)
{
    if (mode == SAMPLE_ZERO) return shape / rate;
    double boost = 1.0, sh = shape;
    if (shape < 1.0) {
        const uint4 xb = philox_site(seed, site, row, 0xffffffffu, sweep);
        const double u = ((double)xb.w + 0.5) * (1.0 / 4294967296.0);   // in (0, 1): [R] redraws u == 0
        boost = pow(u, 1.0 / shape);
        sh = shape + 1.0;
    }
    const double d = sh - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    for (uint32_t a = 0; a < GAMMA_MAX_ATTEMPTS; ++a) {
        const uint4 x = philox_site(seed, site, row, a, sweep);
        const double z = normal_f64(x);
        double v = 1.0 + c * z;
        if (!(v > 0.0)) continue;
        v = v * v * v;
        const double u = ((double)x.z + 0.5) * (1.0 / 4294967296.0);
        if (u < 1.0 - 0.0331 * (z * z) * (z * z)) return d * v * boost / rate;
        if (log(u) < 0.5 * z * z + d * (1.0 - v + log(v))) return d * v * boost / rate;
    }
    return shape / rate;
}

// ---------------------------------------------------------------------------------------------------
// 256-bit global access (LDG.E.256 / STG.E.256 on sm_100): one K8 factor block of one row per instruction
struct __align__(32) f8 { float v[8]; };

#ifdef SBMF_SIMT_EMU   // test-only host build (launch.h): plain loads and stores
__device__ __forceinline__ f8 ld256_nc(const float* p) { return *reinterpret_cast<const f8*>(p); }
__device__ __forceinline__ f8 ld256(const float* p) { return *reinterpret_cast<const f8*>(p); }
__device__ __forceinline__ void st256(float* p, const f8& r) { *reinterpret_cast<f8*>(p) = r; }
#else
__device__ __forceinline__ f8 ld256_nc(const float* p)
{
    f8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}
// coherent variant: for data written earlier in the same kernel / by a concurrently running kernel
__device__ __forceinline__ f8 ld256(const float* p)
{
    f8 r;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st256(float* p, const f8& r)
{
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
                 : "memory");
}
#endif

// Streaming accesses (index / residual / rating streams that are touched once per pass): -DSBMF_STREAM_HINTS=1 marks them
// evict-first (ld.global.cs / st.global.cs) so that they do not displace the gathered factor block from L1 / L2.
#ifndef SBMF_STREAM_HINTS
#define SBMF_STREAM_HINTS 0
#endif
template <class T>
__device__ __forceinline__ T ld_stream(const T* p)
{
#if SBMF_STREAM_HINTS && !defined(SBMF_SIMT_EMU)
    return __ldcs(p);
#else
    return *p;
#endif
}
template <class T>
__device__ __forceinline__ void st_stream(T* p, T v)
{
#if SBMF_STREAM_HINTS && !defined(SBMF_SIMT_EMU)
    __stcs(p, v);
#else
    *p = v;
#endif
}

// Reduce 48 per-lane values over the warp with 48 shuffles instead of 48*5: each step halves the vector a lane
// still carries (the lane keeps one half and ships the other to its partner).  On return v[0..2] of every lane
// hold the warp totals of the original entries reduce_scatter_base(lane) + {0,1,2}; lanes l and l^1 hold the
// same three.  The summation tree is fixed, so results are reproducible run to run.
__device__ __forceinline__ int reduce_scatter_base(int lane)
{
    return ((lane >> 4) & 1) * 24 + ((lane >> 3) & 1) * 12 + ((lane >> 2) & 1) * 6 + ((lane >> 1) & 1) * 3;
}
#ifndef SBMF_FFMA2
#define SBMF_FFMA2 0
#endif
// two independent fp32 additions in one instruction (sm_100 add.rn.f32x2, SASS FADD2): same sums as two FADD (FFMA2=2 builds)
__device__ __forceinline__ void add2(float& x0, float& x1, float a0, float a1, float b0, float b1)
{
#if SBMF_FFMA2 >= 2
    asm("{ .reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd; }"
        : "=f"(x0), "=f"(x1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
#else
    x0 = a0 + b0;
    x1 = a1 + b1;
#endif
}
template <int OFF, int HALF>
__device__ __forceinline__ void reduce_scatter_step(float (&v)[48], int lane)
{
    const bool up = (lane & OFF) != 0;
#if SBMF_FFMA2 >= 2
#pragma unroll
    for (int i = 0; i + 1 < HALF; i += 2) {   // the packed build adds the kept and the received halves two at a time
        const float s0 = up ? v[i] : v[i + HALF], s1 = up ? v[i + 1] : v[i + 1 + HALF];
        const float k0 = up ? v[i + HALF] : v[i], k1 = up ? v[i + 1 + HALF] : v[i + 1];
        add2(v[i], v[i + 1], k0, k1, __shfl_xor_sync(0xffffffffu, s0, OFF), __shfl_xor_sync(0xffffffffu, s1, OFF));
    }
    if (HALF & 1) {
        constexpr int i = HALF - 1;
        const float send = up ? v[i] : v[i + HALF];
        const float keep = up ? v[i + HALF] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
#else
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        const float send = up ? v[i] : v[i + HALF];
        const float keep = up ? v[i + HALF] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
#endif
}
__device__ __forceinline__ void warp_reduce_scatter48(float (&v)[48], int lane)
{
    reduce_scatter_step<16, 24>(v, lane);
    reduce_scatter_step<8, 12>(v, lane);
    reduce_scatter_step<4, 6>(v, lane);
    reduce_scatter_step<2, 3>(v, lane);
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1);
}

__device__ __forceinline__ float warp_sum(float x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

}  // namespace sbmf
