// gram.cuh -- the per-rating accumulation of the Gram-blocked update (kernels.cu): for f = the 8 opposite-side factors of one
// rating and e its residual,   g[k] += f[k] * e  (8 sums)   and   G[k][l] += f[k] * f[l], k <= l  (36 sums).
//
// Two forms with IDENTICAL results (every sum receives the same fused multiply-adds in the same order):
//   scalar : 44 FFMA per rating into float acc[48] laid out by gi()                       (SBMF_FFMA2 = 0)
//   packed : 24 FFMA2 per rating (sm_100 fma.rn.f32x2: two independent IEEE fp32 FMAs per instruction, one operand may be a
//            broadcast scalar) into float2 a2[24].  A pair is two ADJACENT columns (2p, 2p+1) of one row k of G, so the
//            second operand is the register pair (f[2p], f[2p+1]) exactly as the 256-bit load delivers it; rows with odd k
//            start one column early and compute G[k][k-1] a second time (4 redundant products of 40).  (SBMF_FFMA2 = 1)
// The kernels are bound by issue slots, not by the FP32 pipe, which is what the packed form addresses.
// The index logic below is host-callable so that tools/ffma2_check.cu can compare the two forms bit for bit on the CPU.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"
#include "model.h"

#ifndef SBMF_FFMA2
#define SBMF_FFMA2 0
#endif

namespace sbmf {

// packed index of G[k][l], k <= l, in the scalar layout: g[0..7], then the rows of the upper triangle
__host__ __device__ constexpr int gi(int k, int l) { return 8 + k * 8 - (k * (k - 1)) / 2 + (l - k); }

// packed layout: a2[0..3] = (g[2q], g[2q+1]); then row k of G holds the column pairs p = k/2 .. 3 (columns 2p, 2p+1)
__host__ __device__ constexpr int pair_row_start(int k) { return 4 + (k == 0 ? 0 : k == 1 ? 4 : k == 2 ? 8 : k == 3 ? 11 : k == 4 ? 14 : k == 5 ? 16 : k == 6 ? 18 : 19); }
__host__ __device__ constexpr int pi(int k, int p) { return pair_row_start(k) + (p - k / 2); }
constexpr int NPAIR = 24;
static_assert(pi(7, 3) == NPAIR - 1 && pi(0, 0) == 4 && pi(1, 0) == 8 && pi(2, 1) == 12, "pair layout");

// d = s * b + c on both halves (s broadcast)
__host__ __device__ __forceinline__ float2 fma2_bcast(float s, float2 b, float2 c)
{
    float2 d;
#ifdef __CUDA_ARCH__
    asm("{ .reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %2};\n\t"
        "mov.b64 rb, {%3, %4};\n\t"
        "mov.b64 rc, {%5, %6};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y)
        : "f"(s), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
#else
    d.x = fmaf(s, b.x, c.x);
    d.y = fmaf(s, b.y, c.y);
#endif
    return d;
}

__host__ __device__ __forceinline__ void accumulate_scalar(float (&acc)[NACC], const f8& f, float e)
{
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = fmaf(f.v[k], e, acc[k]);
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int l = k; l < 8; ++l) acc[gi(k, l)] = fmaf(f.v[k], f.v[l], acc[gi(k, l)]);
}

__host__ __device__ __forceinline__ void accumulate_packed(float2 (&a2)[NPAIR], const f8& f, float e)
{
#pragma unroll
    for (int q = 0; q < 4; ++q) a2[q] = fma2_bcast(e, make_float2(f.v[2 * q], f.v[2 * q + 1]), a2[q]);
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int p = k / 2; p < 4; ++p) a2[pi(k, p)] = fma2_bcast(f.v[k], make_float2(f.v[2 * p], f.v[2 * p + 1]), a2[pi(k, p)]);
}

// packed sums -> the scalar layout every consumer (warp reduction, slice partials, solve) uses; padding entries are zero
__host__ __device__ __forceinline__ void unpack_pairs(const float2 (&a2)[NPAIR], float (&acc)[NACC])
{
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        acc[2 * q] = a2[q].x;
        acc[2 * q + 1] = a2[q].y;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int l = k; l < 8; ++l) acc[gi(k, l)] = (l & 1) ? a2[pi(k, l / 2)].y : a2[pi(k, l / 2)].x;
#pragma unroll
    for (int i = 44; i < NACC; ++i) acc[i] = 0.f;
}

// Residual update + prediction refresh of one rating: fd = <f, d>, fu = <f, u> (two 8-term chains, each summed in the order
// k = 0..7 like dot8 in kernels.cu).  Packed form: the pair (fd, fu) advances with ONE FFMA2 per k -- f[k] broadcast, the pair
// (d[k], u[k]) prepared once per row and block -- instead of two FFMA; same operations per chain, same results.
__host__ __device__ __forceinline__ float2 dot8_pair(const f8& f, const float2 (&du)[8])
{
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) s = fma2_bcast(f.v[k], du[k], s);
    return s;
}

// 64-bit register pair (lo, hi) as ONE operand: a pair prepared once per row and block stays in an aligned register pair instead of
// being re-assembled by two moves in front of every FFMA2 that reads it (what ptxas did to dot8_pair under register pressure)
#ifndef SBMF_SIMT_EMU
typedef unsigned long long pair64;
__device__ __forceinline__ pair64 pack_pair(float lo, float hi)
{
    pair64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack_pair(pair64 p)
{
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
// d = (s, s) * b + c
__device__ __forceinline__ pair64 fma2_bcast_p(float s, pair64 b, pair64 c)
{
    pair64 d;
    asm("{ .reg .b64 ra;\n\t"
        "mov.b64 ra, {%1, %1};\n\t"
        "fma.rn.f32x2 %0, ra, %2, %3; }"
        : "=l"(d)
        : "f"(s), "l"(b), "l"(c));
    return d;
}
#else   // test-only host build of the kernels (launch.h)
struct pair64 { float lo, hi; };
inline pair64 pack_pair(float lo, float hi) { return pair64{lo, hi}; }
inline float2 unpack_pair(pair64 p) { return float2{p.lo, p.hi}; }
inline pair64 fma2_bcast_p(float s, pair64 b, pair64 c) { return pair64{fmaf(s, b.lo, c.lo), fmaf(s, b.hi, c.hi)}; }
#endif
// (<f, x>, <f, y>) for the 8 prepared pairs xy[k] = (x[k], y[k]): the same two 8-term chains as dot8_pair, k = 0..7
__device__ __forceinline__ float2 dot8_pairs64(const f8& f, const pair64 (&xy)[8])
{
    pair64 s = pack_pair(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) s = fma2_bcast_p(f.v[k], xy[k], s);
    return unpack_pair(s);
}

// "Native" order of the 48 accumulators = the order the build's GramAcc holds them in registers (scalar: gi(); packed: the pair
// layout).  The shared-memory reduction of the row kernels moves them as 12 float4 quads without re-packing, and
// native_entry() tells where each one belongs: kind 0 = g[k], 1 = G[k][l] (k <= l), 2 = padding / redundant copy.
__host__ __device__ constexpr int native_entry(int n, int& k, int& l)
{
    k = 0;
    l = 0;
    if (n < 8) {
        k = n;
        return 0;
    }
#if SBMF_FFMA2
    const int P = n / 2, half = n & 1;
    for (int kk = 0; kk < 8; ++kk)
        for (int p = kk / 2; p < 4; ++p)
            if (pi(kk, p) == P) {
                k = kk;
                l = 2 * p + half;
                return l >= kk ? 1 : 2;
            }
    return 2;
#else
    for (int kk = 0; kk < 8; ++kk)
        for (int ll = kk; ll < 8; ++ll)
            if (gi(kk, ll) == n) {
                k = kk;
                l = ll;
                return 1;
            }
    return 2;
#endif
}

// The accumulator of one (row | slice, block) step in whichever form the build selects.
struct GramAcc {
#if SBMF_FFMA2
    float2 a2[NPAIR];
    __host__ __device__ __forceinline__ float4 quad(int j) const { return make_float4(a2[2 * j].x, a2[2 * j].y, a2[2 * j + 1].x, a2[2 * j + 1].y); }
    __host__ __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int i = 0; i < NPAIR; ++i) a2[i] = make_float2(0.f, 0.f);
    }
    __host__ __device__ __forceinline__ void add(const f8& f, float e) { accumulate_packed(a2, f, e); }
    __host__ __device__ __forceinline__ void finish(float (&acc)[NACC]) const { unpack_pairs(a2, acc); }
#else
    float a[NACC];
    __host__ __device__ __forceinline__ float4 quad(int j) const { return make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]); }
    __host__ __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int i = 0; i < NACC; ++i) a[i] = 0.f;
    }
    __host__ __device__ __forceinline__ void add(const f8& f, float e) { accumulate_scalar(a, f, e); }
    __host__ __device__ __forceinline__ void finish(float (&acc)[NACC]) const
    {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = a[i];
    }
#endif
};

}  // namespace sbmf
