// rows2.cuh -- second generation of the resident-row kernels; included by kernels.cu after PhaseArgs / solve_lanes.
// Same half-step as row_resident_kernel / row_group_kernel ([T]:514-558, 563-606), same Gram-blocked arithmetic per rating; what
// changed is everything AROUND the per-rating work (profiles/sass/SUMMARY_r2.txt, profiles/r2/ncu_rows_rk*.txt):
//
//   * rows owned by several warps: ONE block barrier per factor block instead of four.  Every warp publishes its 48 sums
//     (double-buffered by block parity), after the barrier every warp adds the W partials itself and runs the lane-parallel solve
//     redundantly, so there is no "warp 0 solves, the others wait twice" and no second hop for the deltas.
//   * (d, u_new) pairs of the residual update / prediction refresh prepared once per block as 64-bit operands (gram.cuh pair64):
//     8 FFMA2 + 2 FADD per rating instead of 8 FFMA2 + ~14 moves + 3 FADD.
//   * SMEM_RED = true (option row_kernels = 2): the 48 sums reduced through shared memory instead of the transposed shuffle tree.
//     A lane stores its accumulators as 12 float4 (conflict-free: row stride 52 floats), 24 lanes each sum one quad over 16 lanes'
//     rows (16 LDS.128 + 60 FADD), one shuffle stage joins the halves: ~100 instructions per row block instead of 212, ptxas
//     fuses accumulation and store (4 live accumulators instead of 48).  Measured SLOWER: the kernels are bound by the L1TEX data
//     pipe, where this reduction costs 96 wavefronts against 65 for the shuffles, and its buffers take L1 from the gathers.
//     SMEM_RED = false (option row_kernels = 3, the default): the shuffle tree of common.cuh inside the new structure.
#pragma once

namespace sbmf {

// solve-layout offsets of native accumulator n (gram.cuh native_entry): g[k] -> k; G[k][l] -> 8 + 12k + l and its mirror;
// padding / redundant copies -> the slack words behind the matrix
__constant__ uint8_t c_nat_a[NACC];
__constant__ uint8_t c_nat_b[NACC];

// exchange buffer of the shared-memory reduction, NR rounds: rows of 48 / NR values + 4 floats of padding (13 or 7 quads per row:
// odd, so the 8 lanes of an LDS.128 / STS.128 phase cover all banks)
template <int NR> struct XbGeom {
    static constexpr int QR = 12 / NR;              // quads per round
    static constexpr int STRIDE = 4 * QR + 4;       // floats per lane row
    static constexpr int PARTS = 24 / QR;           // lanes per quad = row ranges summed separately
    static constexpr int RP = 32 / PARTS;           // rows per part
    static constexpr int SH = (NR == 1) ? 4 : 2;    // row rotation per part that keeps the 8 lanes of an LDS.128 phase on distinct banks
    static constexpr int FLOATS = 32 * STRIDE;      // buffer per warp
};

__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 shfl4(float4 v, int src)
{
    return make_float4(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src), __shfl_sync(0xffffffffu, v.z, src),
                       __shfl_sync(0xffffffffu, v.w, src));
}

// Warp totals of the 48 native accumulators, NGRP = 1: the whole warp is one row.  On return lanes 0 .. QR-1 hold, in t[r],
// the totals of quad r * QR + lane.  Fixed summation order (reproducible).  xb = this warp's exchange buffer.
// Lane L < 24 sums quad q = L % QR over the RP rows of part p = L / QR, starting SH * p rows into its part and wrapping around:
// that rotation is what keeps the 8 lanes of one LDS.128 phase on distinct banks when they belong to different parts.  The row
// of step i is (base row + i), minus RP once the rotation has wrapped -- a second base pointer, so every access is base + constant.
template <int NR>
__device__ __forceinline__ void warp_sum48(const GramAcc& ga, float* xb, int lane, float4 (&t)[NR])
{
    using X = XbGeom<NR>;
    const int part = lane / X::QR, q = lane - part * X::QR;
    const float* baseA = xb + (X::RP * part + X::SH * part) * X::STRIDE + 4 * q;   // rows before the wrap
    const float* baseB = baseA - X::RP * X::STRIDE;                                 // rows after it (part 0 never wraps)
    float4* mine = reinterpret_cast<float4*>(xb + lane * X::STRIDE);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
#pragma unroll
        for (int j = 0; j < X::QR; ++j) mine[j] = ga.quad(r * X::QR + j);
        __syncwarp();
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
        if (lane < 24) {
#pragma unroll
            for (int i = 0; i < X::RP; i += 2) {
                // steps (i, i + 1) are on the same side of the wrap point RP - SH * part (SH is even)
                const float* src;
                if (NR == 1) src = (i < X::RP - X::SH) ? baseA : (part ? baseB : baseA);
                else src = (i + X::SH * part >= X::RP) ? baseB : baseA;
                s0 = add4(s0, *reinterpret_cast<const float4*>(src + i * X::STRIDE));
                s1 = add4(s1, *reinterpret_cast<const float4*>(src + (i + 1) * X::STRIDE));
            }
        }
        float4 s = add4(s0, s1);
        if (NR == 2) s = add4(s, shfl4(s, (lane + X::QR) & 31));        // parts (0,1) and (2,3)
        s = add4(s, shfl4(s, (lane + 12) & 31));                        // NR == 1: parts 0 + 1; NR == 2: (0,1) + (2,3)
        t[r] = s;
        __syncwarp();   // every lane has read the buffer before the next round / the next block overwrites it
    }
}

// NGRP = 2 (16 lanes per row) or 4 (8 lanes per row): lane L < 24 owns quad L % 12; NGRP == 2: of row group L / 12 (t[0]);
// NGRP == 4: of row groups 2 * (L / 12) (t[0]) and 2 * (L / 12) + 1 (t[1]).  Lanes 12..23 start 4 rows into their range (banks).
template <int NGRP>
__device__ __forceinline__ void group_sum48(const GramAcc& ga, float* xb, int lane, float4 (&t)[NGRP / 2])
{
    using X = XbGeom<1>;
    constexpr int NT = NGRP / 2;        // sums per lane
    constexpr int RG = 32 / NGRP;       // source lanes per row group
    const int half = lane / 12, q = lane - half * 12;
    const float* baseA = xb + (16 * half + 4 * half) * X::STRIDE + 4 * q;
    const float* baseB = baseA - RG * half * X::STRIDE;
    float4* mine = reinterpret_cast<float4*>(xb + lane * X::STRIDE);
#pragma unroll
    for (int j = 0; j < 12; ++j) mine[j] = ga.quad(j);
    __syncwarp();
#pragma unroll
    for (int u = 0; u < NT; ++u) {
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
        if (lane < 24) {
#pragma unroll
            for (int i = 0; i < RG; i += 2) {
                const float* src = ((i < RG - 4) ? baseA : baseB) + RG * u * X::STRIDE;
                s0 = add4(s0, *reinterpret_cast<const float4*>(src + i * X::STRIDE));
                s1 = add4(s1, *reinterpret_cast<const float4*>(src + (i + 1) * X::STRIDE));
            }
        }
        t[u] = add4(s0, s1);
    }
    __syncwarp();
}

// one quad of totals into the solve layout (g[8] + full symmetric G[8][12]); oa / ob = 4 x 8-bit offsets
__device__ __forceinline__ void scatter_quad(float* tot, float4 v, uint32_t oa, uint32_t ob)
{
    tot[oa & 0xffu] = v.x;
    tot[ob & 0xffu] = v.x;
    tot[(oa >> 8) & 0xffu] = v.y;
    tot[(ob >> 8) & 0xffu] = v.y;
    tot[(oa >> 16) & 0xffu] = v.z;
    tot[(ob >> 16) & 0xffu] = v.z;
    tot[oa >> 24] = v.w;
    tot[ob >> 24] = v.w;
}
// keep a loop-invariant value in its register: ptxas otherwise re-derives the packed offsets inside the block loop from their
// constant-memory bytes, and a constant load with a lane-dependent index replays once per distinct address (ncu: the address
// arithmetic behind those loads was 11 % of the stall samples of row_resident2_kernel<6,1>)
__device__ __forceinline__ void pin_register(uint32_t& v)
{
#ifndef SBMF_SIMT_EMU
    asm volatile("mov.b32 %0, %0;" : "+r"(v));
#endif
}
__device__ __forceinline__ void quad_offsets(int q, uint32_t& oa, uint32_t& ob)
{
    oa = 0;
    ob = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        oa |= (uint32_t)c_nat_a[4 * q + i] << (8 * i);
        ob |= (uint32_t)c_nat_b[4 * q + i] << (8 * i);
    }
    pin_register(oa);
    pin_register(ob);
}

// The lane-parallel solve of solve_lanes (kernels.cu) with the stores optional: rows owned by several warps solve in every
// warp, one of them stores.  Returns the 8 deltas in all lanes.
__device__ __forceinline__ void solve_lanes2(const float* sm, const PhaseArgs& a, size_t foff, float uo, float sig, float mu, bool live, int mode,
                                             float z, float alpha, int kq, bool store, float (&d)[8])
{
    const float4 g0 = *reinterpret_cast<const float4*>(sm + 8 + kq * G_STRIDE);
    const float4 g1 = *reinterpret_cast<const float4*>(sm + 8 + kq * G_STRIDE + 4);
    const float Grow[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float A = sm[8 + kq * (G_STRIDE + 1)];
    const float gk = sm[kq];
    float s = 1.0f / (sig + alpha * A);
    // ran_gaussian(mean, stdev) of random.h:166-172: stdev := 1/lambda or its sqrt; stdev == 0 or NaN -> mean
    float sd = (mode == SAMPLE_ZERO) ? 0.f : ((mode == SAMPLE_SQRT) ? sqrtf(s) : s);
    if (isnan(sd)) sd = 0.f;
    float smu = sig * mu;
    if (!live) {   // padding dimension: mean = 0, no noise, delta = 0
        s = 0.f;
        sd = 0.f;
        smu = 0.f;
    }
    float B = fmaf(A, uo, gk);
    float cand = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        const float mean = s * fmaf(alpha, B, smu);
        cand = uo - fmaf(sd, z, mean);
        const float dl = __shfl_sync(0xffffffffu, cand, l, 8);   // within the octet: octets of a warp may hold different rows
        d[l] = dl;
        if (kq > l) B = fmaf(dl, Grow[l], B);
    }
    // lane kq's B is final after step kq - 1, so the candidate of every later step -- the last one included -- is its own delta
    if (store) {
        const float un = uo - cand;
#pragma unroll 1
        for (int q = 0; q < a.nrep; ++q) a.Frep[q][foff + kq] = un;   // 8 lanes x 4 B = one sector per replica
    }
}

// e += <f, d>, pr += <f, u_new> for the RPL ratings a lane holds
template <int RPL, bool REFRESH>
__device__ __forceinline__ void apply_deltas(const f8 (&f)[RPL], const float (&d)[8], const float (&uo8)[8], float (&e)[RPL], float (&pr)[REFRESH ? RPL : 1])
{
    if (kPairedDots && REFRESH) {
        pair64 du[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) du[k] = pack_pair(d[k], uo8[k] - d[k]);
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const float2 s2 = dot8_pairs64(f[r], du);
            e[r] += s2.x;
            pr[r] += s2.y;
        }
    } else {
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            e[r] += dot8(f[r], d);
            if (REFRESH) {
                float fu = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) fu = fmaf(f[r].v[k], uo8[k] - d[k], fu);
                pr[r] += fu;
            }
        }
    }
}

// --------------------------------------------------------------------------------------------------------
// WARPS warps own one row (WARPS == 1: four independent rows per CTA), RPL (idx, e, f) per lane in registers.
// NR = rounds of the shared-memory reduction (2 halves the exchange buffer: the 8-warp shape would not fit 48 KB otherwise).
// SMEM_RED: the 48 sums go through the exchange buffer (warp_sum48); false: transposed shuffle tree of kernels.cu inside the new
// structure (one barrier per block, pair operands) -- option row_kernels = 3, to tell the two changes apart.
template <int RPL, int WARPS, bool REFRESH, int NR, bool SMEM_RED>
__global__ void __launch_bounds__(WARPS == 1 ? 128 : WARPS * 32, (RPL >= 7 ? 512 : RPL >= 4 ? 640 : 768) / (WARPS == 1 ? 128 : WARPS * 32))
row_resident2_kernel(PhaseArgs a, const uint32_t* __restrict__ rows, uint32_t nrows, int b_begin, int b_end, int do_bias)
{
    using X = XbGeom<NR>;
    static_assert(WARPS > 1 || NR == 1, "single-warp rows reduce in one round");
    constexpr int WPC = (WARPS == 1) ? 4 : WARPS;   // warps per CTA
    __shared__ __align__(16) float s_xb[SMEM_RED ? WPC : 1][SMEM_RED ? X::FLOATS : 4];
    __shared__ __align__(16) float s_tot[WPC][SOLVE_SMEM];
    __shared__ __align__(16) float s_part[(WARPS == 1) ? 1 : 2][(WARPS == 1) ? 1 : WARPS][NACC];
    __shared__ float s_bias[(WARPS == 1) ? 1 : WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t r_idx = (WARPS == 1) ? blockIdx.x * WPC + warp : blockIdx.x;
    if (r_idx >= nrows) return;   // WARPS == 1: whole warp leaves; no block-wide barrier is used in that shape
    const uint32_t row = rows[r_idx];
    const uint32_t rid = a.row_id ? a.row_id[row] : row;   // the caller's id of this row: the key of its draws (relabelled models)
    const int64_t beg = a.ptr[row];
    const int c = (int)(a.ptr[row + 1] - beg);
    const int t_in_row = (WARPS == 1) ? lane : threadIdx.x;
    constexpr int TPR = WARPS * 32;   // threads per row
    float* tot = s_tot[warp];
    float* xb = s_xb[SMEM_RED ? warp : 0];

    const int mode = a.mode;
    const uint32_t K = a.K;
    const float alpha = a.sc->alpha_f;
    const uint32_t sweep = a.sc->sweep;
    const uint32_t ns_other = a.ns_other, ns_self = a.ns_self;
    const uint32_t pad_row = ns_other - 1;
    const float* __restrict__ Fother = a.Fother;
    const int kq = lane & 7;
    // solve-layout offsets of quad `lane` (lanes 0..11 scatter): native accumulator order (SMEM_RED) or the scalar order the
    // shuffle tree delivers; single-warp rows on the shuffle path store their 3 sums per even lane like row_resident_kernel
    uint32_t off_a = 0, off_b = 0;
    if (SMEM_RED) quad_offsets(lane < 12 ? lane : 0, off_a, off_b);
    else if (WARPS > 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            off_a |= (uint32_t)c_pk_a[4 * (lane < 12 ? lane : 0) + i] << (8 * i);
            off_b |= (uint32_t)c_pk_b[4 * (lane < 12 ? lane : 0) + i] << (8 * i);
        }
        pin_register(off_a);
        pin_register(off_b);
    } else {
        const int base = reduce_scatter_base(lane);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            off_a |= (uint32_t)c_pk_a[base + i] << (8 * i);
            off_b |= (uint32_t)c_pk_b[base + i] << (8 * i);
        }
        pin_register(off_a);
        pin_register(off_b);
    }

    uint32_t id[RPL];
    float e[RPL];
    float pr[REFRESH ? RPL : 1];   // partial prediction sum_b <f_b, u_new_b>
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int p = r * TPR + t_in_row;
        const bool valid = p < c;
        id[r] = valid ? a.idx[beg + p] : pad_row;
        e[r] = valid ? load_e_first(a, beg + p) : 0.f;
        if (REFRESH) pr[r] = (valid && b_begin > 0) ? a.pacc[beg + p] : 0.f;
    }
    float bias_new = 0.f;
    if (REFRESH && !do_bias) bias_new = a.bias[row];
    // first block's gathers go out before the bias half-step
    f8 f[RPL];
    {
        const float* Fo = Fother + (size_t)b_begin * ns_other * 8;
#pragma unroll
        for (int r = 0; r < RPL; ++r) f[r] = ld256_nc(Fo + (size_t)id[r] * 8);
    }

    if (do_bias) {
        const float shift = a.apply_shift ? a.sc->shift_f : 0.f;
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < RPL; ++r)
            if (id[r] != pad_row) {
                e[r] += shift;
                t += e[r];
            }
        // every thread of the row reads the old bias BEFORE the reduction (shuffles / barriers): thread 0 overwrites it below
        const float bo = a.bias[row], sb = a.sigma_b[row], mb = a.mu_b[row];
        t = warp_sum(t);
        if (WARPS > 1) {
            if (lane == 0) s_bias[warp] = t;
            __syncthreads();
            t = 0.f;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) t += s_bias[w];
        }
        const float s = 1.0f / (sb + alpha * (float)c);
        const float mean = s * (sb * mb + alpha * (t + (float)c * bo));
        float z = 0.f;
        if (mode != SAMPLE_ZERO) z = normal_f32(philox_site(a.seed, a.site_b, rid, 0u, sweep));
        const float bn = draw_f32(mode, mean, s, z);
        const float d = bo - bn;
#pragma unroll
        for (int r = 0; r < RPL; ++r)
            if (id[r] != pad_row) e[r] += d;
        if (t_in_row == 0)
            for (int q = 0; q < a.nrep; ++q) a.brep[q][row] = bn;
        bias_new = bn;
    }

    float zq = 0.f;
    for (int b = b_begin; b < b_end; ++b) {
        // this lane's dimension of the block: old value and hyper-parameters (latency hidden behind the accumulation)
        const size_t foff = ((size_t)b * ns_self + row) * 8;
        const bool live = (uint32_t)(b * 8 + kq) < K;
        const float uo = live ? a.Fself[foff + kq] : 0.f;
        const float sig = a.sigma_kf[b * 8 + kq], mu = a.mu_kf[b * 8 + kq];
        // this row's noise, 4 blocks at a time: lane l draws dimension 32*(b/4) + l
        if (mode != SAMPLE_ZERO && (((b & 3) == 0) || b == b_begin))
            zq = normal_f32(philox_site(a.seed, a.site_f, rid, (uint32_t)((b & ~3) * 8 + lane), sweep));
        GramAcc ga;
        ga.clear();
#pragma unroll
        for (int r = 0; r < RPL; ++r) ga.add(f[r], e[r]);
        float* part = s_part[(WARPS == 1) ? 0 : (b & 1)][(WARPS == 1) ? 0 : warp];
        if (SMEM_RED) {
            float4 t4[NR];
            warp_sum48<NR>(ga, xb, lane, t4);
            if (WARPS == 1) {
                if (lane < 12) scatter_quad(tot, t4[0], off_a, off_b);
            } else if (lane < X::QR) {
#pragma unroll
                for (int r = 0; r < NR; ++r) *reinterpret_cast<float4*>(part + 4 * (r * X::QR + lane)) = t4[r];
            }
        } else {
            float acc[NACC];
            ga.finish(acc);
            warp_reduce_scatter48(acc, lane);
            if ((lane & 1) == 0) {
                if (WARPS == 1) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        tot[(off_a >> (8 * i)) & 0xff] = acc[i];
                        tot[(off_b >> (8 * i)) & 0xff] = acc[i];
                    }
                } else {
                    const int rb = reduce_scatter_base(lane);
                    part[rb] = acc[0];
                    part[rb + 1] = acc[1];
                    part[rb + 2] = acc[2];
                }
            }
        }
        if (WARPS > 1) {
            __syncthreads();   // the only block barrier of a factor block: s_part is double-buffered by block parity
            if (lane < 12) {
                float4 s = *reinterpret_cast<const float4*>(s_part[b & 1][0] + 4 * lane);
#pragma unroll
                for (int w = 1; w < WARPS; ++w) s = add4(s, *reinterpret_cast<const float4*>(s_part[b & 1][w] + 4 * lane));
                scatter_quad(tot, s, off_a, off_b);
            }
        }
        __syncwarp();
        const float z = __shfl_sync(0xffffffffu, zq, ((b & 3) << 3) + kq);
        float uo8[8];
        if (REFRESH) {
#pragma unroll
            for (int k = 0; k < 8; ++k) uo8[k] = __shfl_sync(0xffffffffu, uo, k);
        }
        float d[8];
        solve_lanes2(tot, a, foff, uo, sig, mu, live, mode, z, alpha, kq, lane < 8 && (WARPS == 1 || warp == 0), d);
        apply_deltas<RPL, REFRESH>(f, d, uo8, e, pr);
        if (b + 1 < b_end) {   // next block's gathers (a register double buffer for them costs more occupancy than it hides latency)
            const float* Fo = Fother + (size_t)(b + 1) * ns_other * 8;
#pragma unroll
            for (int r = 0; r < RPL; ++r) f[r] = ld256_nc(Fo + (size_t)id[r] * 8);
        }
        __syncwarp();   // tot (this warp's) is rewritten by the next block
    }

    if (REFRESH && b_end == a.KBtot) {   // the phase is complete for this row: fresh residual instead of the incremental one
        const float b0 = a.sc->b_0_f;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int p = r * TPR + t_in_row;
            if (p < c) store_e_final(a, beg + p, a.r[beg + p] - (b0 + bias_new + a.bias_other[id[r]] + pr[r]));
        }
        return;
    }
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int p = r * TPR + t_in_row;
        if (p < c) {
            if (b_end == a.KBtot) store_e_final(a, beg + p, e[r]);   // (not a REFRESH phase: its last launch returned above)
            else a.e[beg + p] = e[r];
            if (REFRESH) a.pacc[beg + p] = pr[r];
        }
    }
}

// --------------------------------------------------------------------------------------------------------
// Short rows: G = 8 or 16 lanes own a row, 32 / G rows per warp (see row_group_kernel); the reduction serves all rows of the
// warp with the same instructions.
template <int RPL, int G, bool REFRESH, bool SMEM_RED>
__global__ void __launch_bounds__(128, (RPL >= 6 ? 512 : RPL >= 4 ? 640 : 768) / 128)
row_group2_kernel(PhaseArgs a, const uint32_t* __restrict__ rows, uint32_t nrows, int b_begin, int b_end, int do_bias)
{
    using X = XbGeom<1>;
    constexpr int RPW = 32 / G;          // rows per warp
    constexpr int ZB = G / 8;            // blocks covered by one noise draw
    constexpr int NV = (G == 16) ? 3 : 6;   // shuffle path: reduced values per lane
    __shared__ __align__(16) float s_xb[SMEM_RED ? 4 : 1][SMEM_RED ? X::FLOATS : 4];
    __shared__ __align__(16) float s_tot[4 * RPW][SOLVE_SMEM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G, lg = lane % G, kq = lane & 7;
    const uint32_t r_idx = (blockIdx.x * 4 + warp) * RPW + grp;
    const bool have_row = r_idx < nrows;                 // idle groups run along (the shuffles are warp-wide) but touch nothing
    const uint32_t row = rows[have_row ? r_idx : nrows - 1];
    const uint32_t rid = a.row_id ? a.row_id[row] : row;   // the caller's id of this row: the key of its draws (relabelled models)
    const int64_t beg = a.ptr[row];
    const int c = have_row ? (int)(a.ptr[row + 1] - beg) : 0;
    float* tot = s_tot[warp * RPW + grp];
    float* xb = s_xb[SMEM_RED ? warp : 0];

    const int mode = a.mode;
    const uint32_t K = a.K;
    const float alpha = a.sc->alpha_f;
    const uint32_t sweep = a.sc->sweep;
    const uint32_t ns_other = a.ns_other, ns_self = a.ns_self;
    const uint32_t pad_row = ns_other - 1;
    const float* __restrict__ Fother = a.Fother;
    // lanes 0..23 scatter: quad lane % 12 of row group(s) lane / 12 (G == 16) or 2 * (lane / 12) + {0, 1} (G == 8)
    uint32_t off_a = 0, off_b = 0;
    uint64_t sh_a = 0, sh_b = 0;   // shuffle path: NV x 8-bit offsets of this lane's reduced sums
    if (SMEM_RED) quad_offsets(lane < 24 ? lane % 12 : 0, off_a, off_b);
    else {
        const int base0 = (G == 16) ? ((lg >> 3) & 1) * 24 + ((lg >> 2) & 1) * 12 + ((lg >> 1) & 1) * 6 + (lg & 1) * 3
                                    : ((lg >> 2) & 1) * 24 + ((lg >> 1) & 1) * 12 + (lg & 1) * 6;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            sh_a |= (uint64_t)c_pk_a[base0 + i] << (8 * i);
            sh_b |= (uint64_t)c_pk_b[base0 + i] << (8 * i);
        }
    }
    float* tot_s0 = s_tot[warp * RPW + ((G == 16) ? (lane < 24 ? lane / 12 : 0) : (lane < 24 ? 2 * (lane / 12) : 0))];

    uint32_t id[RPL];
    float e[RPL];
    float pr[REFRESH ? RPL : 1];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int p = r * G + lg;
        const bool valid = p < c;
        id[r] = valid ? a.idx[beg + p] : pad_row;
        e[r] = valid ? load_e_first(a, beg + p) : 0.f;
        if (REFRESH) pr[r] = (valid && b_begin > 0) ? a.pacc[beg + p] : 0.f;
    }
    float bias_new = 0.f;
    if (REFRESH && !do_bias) bias_new = a.bias[row];
    f8 f[RPL];
    {
        const float* Fo = Fother + (size_t)b_begin * ns_other * 8;
#pragma unroll
        for (int r = 0; r < RPL; ++r) f[r] = ld256_nc(Fo + (size_t)id[r] * 8);
    }

    if (do_bias) {
        const float shift = a.apply_shift ? a.sc->shift_f : 0.f;
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < RPL; ++r)
            if (id[r] != pad_row) {
                e[r] += shift;
                t += e[r];
            }
        // read before the (warp-wide) shuffles: lane 0 of the group overwrites the bias below (idle groups alias the last row)
        const float bo = have_row ? a.bias[row] : 0.f, sb = a.sigma_b[row], mb = a.mu_b[row];
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        const float s = 1.0f / (sb + alpha * (float)c);
        const float mean = s * (sb * mb + alpha * (t + (float)c * bo));
        float z = 0.f;
        if (mode != SAMPLE_ZERO) z = normal_f32(philox_site(a.seed, a.site_b, rid, 0u, sweep));
        const float bn = draw_f32(mode, mean, s, z);
        const float d = bo - bn;
#pragma unroll
        for (int r = 0; r < RPL; ++r)
            if (id[r] != pad_row) e[r] += d;
        if (lg == 0 && have_row)
            for (int q = 0; q < a.nrep; ++q) a.brep[q][row] = bn;
        bias_new = bn;
    }

    float zq = 0.f;
    for (int b = b_begin; b < b_end; ++b) {
        const size_t foff = ((size_t)b * ns_self + row) * 8;
        const bool live = have_row && (uint32_t)(b * 8 + kq) < K;   // idle groups alias the last row: they must not read what its owner writes
        const float uo = live ? a.Fself[foff + kq] : 0.f;
        const float sig = a.sigma_kf[b * 8 + kq], mu = a.mu_kf[b * 8 + kq];
        // this row's noise, ZB blocks at a time: lane lg of the group draws dimension 8 * ZB * (b / ZB) + lg
        if (mode != SAMPLE_ZERO && (((b % ZB) == 0) || b == b_begin))
            zq = normal_f32(philox_site(a.seed, a.site_f, rid, (uint32_t)((b - b % ZB) * 8 + lg), sweep));
        GramAcc ga;
        ga.clear();
#pragma unroll
        for (int r = 0; r < RPL; ++r) ga.add(f[r], e[r]);
        if (SMEM_RED) {
            float4 t4[RPW / 2];
            group_sum48<RPW>(ga, xb, lane, t4);
            if (lane < 24) {
                scatter_quad(tot_s0, t4[0], off_a, off_b);
                if (RPW == 4) scatter_quad(tot_s0 + SOLVE_SMEM, t4[RPW / 2 - 1], off_a, off_b);
            }
        } else {
            float acc[NACC];
            ga.finish(acc);
            group_reduce_scatter48<G>(acc, lg);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                tot[(uint32_t)(sh_a >> (8 * i)) & 0xffu] = acc[i];
                tot[(uint32_t)(sh_b >> (8 * i)) & 0xffu] = acc[i];
            }
        }
        __syncwarp();
        const float z = __shfl_sync(0xffffffffu, zq, ((b % ZB) << 3) + kq, G);
        float uo8[8];
        if (REFRESH) {
#pragma unroll
            for (int k = 0; k < 8; ++k) uo8[k] = __shfl_sync(0xffffffffu, uo, k, 8);
        }
        // every octet solves its group's row (for G == 16 both octets of the group compute the same values, the first stores)
        float d[8];
        solve_lanes2(tot, a, foff, uo, sig, mu, live, mode, z, alpha, kq, have_row && lg < 8, d);
        apply_deltas<RPL, REFRESH>(f, d, uo8, e, pr);
        if (b + 1 < b_end) {
            const float* Fo = Fother + (size_t)(b + 1) * ns_other * 8;
#pragma unroll
            for (int r = 0; r < RPL; ++r) f[r] = ld256_nc(Fo + (size_t)id[r] * 8);
        }
        __syncwarp();
    }

    if (REFRESH && b_end == a.KBtot) {
        const float b0 = a.sc->b_0_f;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int p = r * G + lg;
            if (p < c) store_e_final(a, beg + p, a.r[beg + p] - (b0 + bias_new + a.bias_other[id[r]] + pr[r]));
        }
        return;
    }
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int p = r * G + lg;
        if (p < c) {
            if (b_end == a.KBtot) store_e_final(a, beg + p, e[r]);   // (not a REFRESH phase: its last launch returned above)
            else a.e[beg + p] = e[r];
            if (REFRESH) a.pacc[beg + p] = pr[r];
        }
    }
}

}  // namespace sbmf
