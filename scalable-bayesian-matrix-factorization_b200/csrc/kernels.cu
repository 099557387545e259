// kernels.cu -- the device side of one SBMF Gibbs sweep (gibbs_sbpmf2.cpp "[T]":335-637), sm_100a.
//
// The factor half-steps ([T]:538-556 users, 586-605 items) are restructured, not translated.  [T] walks a
// row's ratings twice per latent dimension (reduce, then residual update).  Here 8 consecutive dimensions
// ("a block") are updated from ONE pass: with f_p = the 8 opposite-side factors of rating p (one 32-byte
// sector, one LDG.256) the kernel accumulates
//        g[k]    = sum_p f_p[k] * e_p                  (8 values)
//        G[k][l] = sum_p f_p[k] * f_p[l],  k <= l      (36 values)
// and then replays [T]'s sequential coordinate updates for k = 0..7 exactly, because the sums [T] would
// compute after the earlier dimensions of the block changed e are
//        A_k = G[k][k],   B_k = g[k] + sum_{l<k} G[l][k] * (u_l_old - u_l_new) + G[k][k] * u_k_old .
// The residual is then updated once, e_p += sum_k f_p[k] * (u_k_old - u_k_new).  Same chain of conditionals,
// same scan order, 1/8 of the residual/index traffic; rows of up to 2048 ratings keep (idx, e, f) in
// registers so the residual is read and written once per LAUNCH (possibly several blocks), longer rows stream
// through a sliced two-kernel pipeline that applies the previous block's delta while accumulating the next.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "gram.cuh"
#include "launch.h"
#include "model.h"

namespace sbmf {

// --------------------------------------------------------------------------------------------------------
struct PhaseArgs {
    const int64_t* ptr;
    const uint32_t* idx;
    float* e;
    float* Fself;
    // replicas of Fself / bias to write updated rows to: [0, nrep).  One GPU: just the local arrays.  G GPUs with peer access:
    // every rank's replica (peer-mapped over NVLink) -- the all-gather is fused into the update
    float* Frep[MAX_PEERS];
    float* brep[MAX_PEERS];
    int nrep;
    const float* Fother;
    // F2other[pb][row][16] = (block pb, block pb + 1) of every row of the other side side by side (pair_pack_kernel): the
    // streaming pipeline's pass b gathers the two blocks it needs (b - 1 to apply, b to accumulate) as ONE 64-byte row that two
    // adjacent lanes fetch together -- half the L1TEX wavefronts of two separate sector gathers (sbmf_cuda_probe: 1.50 vs 0.88
    // sectors per clock per SM)
    const float* F2other;
    uint32_t ns_self, ns_other;   // rows per factor block INCLUDING the all-zero pad row at index n (gather target of empty slots)
    float* bias;
    const float* mu_b;
    const float* sigma_b;
    const float* sigma_kf;
    const float* mu_kf;
    const Scalars* sc;
    uint32_t K;
    uint64_t seed;
    uint32_t site_f, site_b;
    const uint32_t* rows_of_hrow;   // heavy_rows of this side
    const uint32_t* row_id;         // [n] caller's id of each row (Side::id_at; nullptr = identity): Philox keys of the row's draws
    int mode;          // SampleMode
    int apply_shift;   // user phase: add sc->shift_f on the first touch of e ([T]:407-410)
    // residual refresh fused into this phase (REFRESH kernels): while the blocks are processed the prediction
    // p = sum_b <f_b, u_new_b> is accumulated from the factors already in registers, and the phase ends with
    // e = r - (b_0 + b_self + b_other + p): the rebuild of [T]:342-359 at the cost of one extra 4-byte gather per rating
    const float* r;           // rating per slot (this side's order)
    const float* bias_other;  // [n_other + 1] opposite-side biases (pad row -> 0 is never read)
    float* pacc;              // [slots] partial predictions carried between launches / streaming passes
    int KBtot;                // total number of blocks (to know the last one)
    // Residual hand-over between the phases, folded into the first touch: when e_map is set, the current residual of slot s
    // is e_src[e_map[s]] (it lives in the OTHER side's slot order; e_map = perm or its inverse) and e[] is write-only until
    // this phase has stored it.  Replaces the stand-alone permutation pass between the phases on one GPU.
    const float* e_src;
    const uint32_t* e_map;
    // Fused forward exchange (G GPUs, peer pushes; Model::xmap_fwd): the phase's FINAL residual of local slot s does not go to e[s]
    // but to xdst[x >> 28][x & 0x0fffffff], x = xmap[s] -- the receive buffer of the rank that owns the rating's row on the other
    // side, at the position its own first touch will read it from (e_src / e_map of its next phase).
    const uint32_t* xmap;
    float* xdst[MAX_PEERS];
};

// final store of a phase's residual (last block processed): local array, or straight into the owner rank's receive buffer
__device__ __forceinline__ void store_e_final(const PhaseArgs& a, int64_t slot, float v)
{
    if (a.xmap) {
        const uint32_t x = a.xmap[slot];
        a.xdst[x >> 28][x & 0x0fffffffu] = v;
    } else {
        a.e[slot] = v;
    }
}

__device__ __forceinline__ float load_e_first(const PhaseArgs& a, int64_t slot)
{
    return a.e_map ? __ldg(a.e_src + a.e_map[slot]) : a.e[slot];
}

// gi(k, l), the per-rating accumulation (GramAcc: 44 FFMA, or 24 FFMA2 with -DSBMF_FFMA2=1) and its layouts: gram.cuh
constexpr bool kPairedDots = SBMF_FFMA2 != 0;   // REFRESH kernels: (<f, d>, <f, u>) as one FFMA2 chain (dot8_pair)

__device__ __forceinline__ float dot8(const f8& a, const float (&d)[8])
{
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s = fmaf(a.v[k], d[k], s);
    return s;
}

// ran_gaussian(mean, stdev) of random.h:166-172 with [T]'s "variance passed as stdev" 
__device__ __forceinline__ float draw_f32(int mode, float mean, float var, float z)
{
    if (mode == SAMPLE_ZERO) return mean;
    const float sd = (mode == SAMPLE_SQRT) ? sqrtf(var) : var;
    if (sd == 0.f || isnan(sd)) return mean;
    return fmaf(sd, z, mean);
}

// The 8 sequential coordinate updates of one block ([T]:546-551 / 594-599) from the reduced sums.
// tot = g[8], G[36] as laid out by gi(); z[k] = this row's standard normal for dimension 8b+k.
__device__ __forceinline__ void solve_block(const float* tot, const f8& uo, const float (&z)[8], const float* sig, const float* mu,
                                            float alpha, int mode, int kvalid, f8& un, float (&d)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < kvalid) {
            const float A = tot[gi(k, k)];
            float B = fmaf(A, uo.v[k], tot[k]);
#pragma unroll
            for (int l = 0; l < k; ++l) B = fmaf(d[l], tot[gi(l, k)], B);
            const float s = 1.0f / (sig[k] + alpha * A);
            const float mean = s * (alpha * B + sig[k] * mu[k]);
            un.v[k] = draw_f32(mode, mean, s, z[k]);
            d[k] = uo.v[k] - un.v[k];
        } else {   // padding dimensions of the last block stay exactly zero
            un.v[k] = 0.f;
            d[k] = 0.f;
        }
    }
}

// --------------------------------------------------------------------------------------------------------
// Resident rows: WARPS warps own one row, each lane keeps RPL (idx, e, f) triples in registers.
// blocks [b_begin, b_end) are processed in one launch; do_bias runs the bias half-step ([T]:517-534 / 566-582) first.
//
// Per block: gather f (empty slots point at the all-zero pad row, so they need no masking), accumulate g/G, transposed
// warp reduction, sums to shared memory as g[8] + a FULL symmetric G[8][8], then the 8 sequential coordinate updates run
// lane-parallel: lane k owns dimension k (its row of G, its 1/lambda), and step l broadcasts lane l's delta to the lanes
// l' > l, which fold it into their B.  The row's noise is drawn 4 blocks at a time (lane -> one of 32 dimensions).
constexpr int G_STRIDE = 12;                      // row stride of the symmetric G in shared memory: LDS.128-aligned and bank-conflict-free
constexpr int SOLVE_SMEM = 8 + 8 * G_STRIDE + 8;  // g[8], G[8][12], 8 floats of slack for the 4 padding accumulators

// smem offset(s) of packed accumulator p: g[k] -> k; G[k][l] -> 8 + 12k + l and its mirror
__constant__ uint8_t c_pk_a[NACC];
__constant__ uint8_t c_pk_b[NACC];

struct SolveOut {
    float d[8];   // u_old - u_new for the 8 dimensions of the block (all lanes)
    float mine;   // d[lane & 7]
};

// uo / sig / mu are this lane's (dimension kq = lane & 7) old factor value and hyper-parameters, loaded by the caller
// early enough to hide their latency; lanes of padding dimensions pass uo = 0 and get delta 0.
__device__ __forceinline__ SolveOut solve_lanes(const float* sm, const PhaseArgs& a, size_t foff, float uo, float sig, float mu, bool live,
                                                int mode, float z, float alpha, int lane)
{
    const int kq = lane & 7;
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
    SolveOut o;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        const float mean = s * fmaf(alpha, B, smu);
        const float cand = uo - fmaf(sd, z, mean);
        const float dl = __shfl_sync(0xffffffffu, cand, l, 8);   // within the octet: octets of a warp may hold different rows
        o.d[l] = dl;
        if (kq > l) B = fmaf(dl, Grow[l], B);
    }
    float mine = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) mine = (kq == l) ? o.d[l] : mine;
    o.mine = mine;
    if (lane < 8) {   // callers pass the lane index WITHIN the row's lane group
        const float un = uo - mine;
        for (int q = 0; q < a.nrep; ++q) a.Frep[q][foff + kq] = un;   // 8 lanes x 4 B = one sector per replica
    }
    return o;
}

// register caps (min CTAs per SM): 8 ratings per lane -> 128 registers, 4 -> 102, fewer -> 80
template <int RPL, int WARPS, bool REFRESH>
__global__ void __launch_bounds__(WARPS == 1 ? 128 : WARPS * 32, (RPL >= 7 ? 512 : RPL >= 4 ? 640 : 768) / (WARPS == 1 ? 128 : WARPS * 32))
row_resident_kernel(PhaseArgs a, const uint32_t* __restrict__ rows, uint32_t nrows, int b_begin, int b_end, int do_bias)
{
    constexpr int WPC = (WARPS == 1) ? 4 : WARPS;   // warps per CTA
    __shared__ __align__(16) float s_tot[(WARPS == 1) ? WPC : 1][SOLVE_SMEM];
    __shared__ __align__(16) float s_part[(WARPS == 1) ? 1 : WARPS][NACC];
    __shared__ float s_d[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t r_idx = (WARPS == 1) ? blockIdx.x * WPC + warp : blockIdx.x;
    if (r_idx >= nrows) return;   // WARPS == 1: whole warp leaves; no block-wide barrier is used in that shape
    const uint32_t row = rows[r_idx];
    const uint32_t rid = a.row_id ? a.row_id[row] : row;   // the caller's id of this row: the key of its draws (relabelled models)
    const int64_t beg = a.ptr[row];
    const int c = (int)(a.ptr[row + 1] - beg);
    const int t_in_row = (WARPS == 1) ? lane : threadIdx.x;
    constexpr int TPR = WARPS * 32;   // threads per row
    float* tot = (WARPS == 1) ? s_tot[warp] : s_tot[0];

    const int mode = a.mode;
    const uint32_t K = a.K;
    const float alpha = a.sc->alpha_f;
    const uint32_t sweep = a.sc->sweep;
    const uint32_t ns_other = a.ns_other, ns_self = a.ns_self;
    const uint32_t pad_row = ns_other - 1;
    const float* __restrict__ Fother = a.Fother;
    const int kq = lane & 7;

    // where this lane's (or, WARPS > 1, this thread's) reduced sums go in shared memory
    const int base = (WARPS == 1) ? reduce_scatter_base(lane) : (int)threadIdx.x;
    uint32_t off_a = 0, off_b = 0;   // 3 x 8-bit offsets each
    if (WARPS == 1) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            off_a |= (uint32_t)c_pk_a[base + i] << (8 * i);
            off_b |= (uint32_t)c_pk_b[base + i] << (8 * i);
        }
    } else if (threadIdx.x < NACC) {
        off_a = c_pk_a[base];
        off_b = c_pk_b[base];
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
            if (lane == 0) s_part[warp][0] = t;
            __syncthreads();
            t = 0.f;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) t += s_part[w][0];
            __syncthreads();
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
        float acc[NACC];
        {
            GramAcc ga;
            ga.clear();
#pragma unroll
            for (int r = 0; r < RPL; ++r) ga.add(f[r], e[r]);
            ga.finish(acc);
        }
        warp_reduce_scatter48(acc, lane);
        if (WARPS == 1) {
            if ((lane & 1) == 0) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    tot[(off_a >> (8 * i)) & 0xff] = acc[i];
                    tot[(off_b >> (8 * i)) & 0xff] = acc[i];
                }
            }
            __syncwarp();
        } else {
            const int rb = reduce_scatter_base(lane);
            if ((lane & 1) == 0) {
                s_part[warp][rb] = acc[0];
                s_part[warp][rb + 1] = acc[1];
                s_part[warp][rb + 2] = acc[2];
            }
            __syncthreads();
            if (threadIdx.x < NACC) {
                float sum = 0.f;
#pragma unroll
                for (int w = 0; w < WARPS; ++w) sum += s_part[w][threadIdx.x];
                tot[off_a] = sum;
                tot[off_b] = sum;
            }
            __syncthreads();
        }
        const float z = __shfl_sync(0xffffffffu, zq, ((b & 3) << 3) + kq);
        float uo8[8];
        if (REFRESH) {
#pragma unroll
            for (int k = 0; k < 8; ++k) uo8[k] = __shfl_sync(0xffffffffu, uo, k);
        }
        SolveOut so;
        if (WARPS == 1) {
            so = solve_lanes(tot, a, foff, uo, sig, mu, live, mode, z, alpha, lane);
        } else {   // warp 0 solves and publishes the 8 deltas
            if (warp == 0) {
                so = solve_lanes(tot, a, foff, uo, sig, mu, live, mode, z, alpha, lane);
                if (lane < 8) s_d[lane] = so.mine;
            }
            __syncthreads();
#pragma unroll
            for (int l = 0; l < 8; ++l) so.d[l] = s_d[l];
        }
        if (kPairedDots && REFRESH) {   // (<f, d>, <f, u_old>) as one pair per rating: 8 FFMA2 instead of 16 FFMA (gram.cuh)
            float2 du[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) du[k] = make_float2(so.d[k], uo8[k]);
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const float2 s2 = dot8_pair(f[r], du);
                e[r] += s2.x;
                pr[r] += s2.y - s2.x;
            }
        } else {
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const float fd = dot8(f[r], so.d);
                e[r] += fd;
                if (REFRESH) {   // <f, u_new> = <f, u_old> - <f, d>; u_old broadcast from the 8 lanes that hold it
                    float fu = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) fu = fmaf(f[r].v[k], uo8[k], fu);
                    pr[r] += fu - fd;
                }
            }
        }
        if (b + 1 < b_end) {   // next block's gathers (a register double buffer for them costs more occupancy than it hides latency)
            const float* Fo = Fother + (size_t)(b + 1) * ns_other * 8;
#pragma unroll
            for (int r = 0; r < RPL; ++r) f[r] = ld256_nc(Fo + (size_t)id[r] * 8);
        }
        if (WARPS == 1) __syncwarp();   // tot is rewritten by the next block
        else __syncthreads();
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
// Short rows: G = 8 or 16 lanes own a row (RPL ratings per lane, rows of <= G * RPL ratings), so one warp works on 32 / G rows
// at once.  Same arithmetic as row_resident_kernel<RPL, 1>, but the fixed per-row-block costs are shared: the transposed
// reduction needs log2(G) halving steps and serves all rows of the warp with the same instructions, every octet runs the
// lane-parallel solve for its own row, and the noise is drawn G / 8 blocks at a time.
template <int G>
__device__ __forceinline__ int group_reduce_scatter48(float (&v)[NACC], int lg)   // returns the packed index of v[0]
{
    if (G == 16) {
        reduce_scatter_step<8, 24>(v, lg);
        reduce_scatter_step<4, 12>(v, lg);
        reduce_scatter_step<2, 6>(v, lg);
        reduce_scatter_step<1, 3>(v, lg);
        return ((lg >> 3) & 1) * 24 + ((lg >> 2) & 1) * 12 + ((lg >> 1) & 1) * 6 + (lg & 1) * 3;
    } else {
        reduce_scatter_step<4, 24>(v, lg);
        reduce_scatter_step<2, 12>(v, lg);
        reduce_scatter_step<1, 6>(v, lg);
        return ((lg >> 2) & 1) * 24 + ((lg >> 1) & 1) * 12 + (lg & 1) * 6;
    }
}

template <int RPL, int G, bool REFRESH>
__global__ void __launch_bounds__(128, (RPL >= 6 ? 512 : RPL >= 4 ? 640 : 768) / 128)
row_group_kernel(PhaseArgs a, const uint32_t* __restrict__ rows, uint32_t nrows, int b_begin, int b_end, int do_bias)
{
    constexpr int RPW = 32 / G;          // rows per warp
    constexpr int NV = (G == 16) ? 3 : 6;   // reduced values per lane
    constexpr int ZB = G / 8;            // blocks covered by one noise draw
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

    const int mode = a.mode;
    const uint32_t K = a.K;
    const float alpha = a.sc->alpha_f;
    const uint32_t sweep = a.sc->sweep;
    const uint32_t ns_other = a.ns_other, ns_self = a.ns_self;
    const uint32_t pad_row = ns_other - 1;
    const float* __restrict__ Fother = a.Fother;

    // where this lane's NV reduced sums go in shared memory (packed once: constant-memory look-ups with a lane-dependent index are slow)
    uint64_t off_a = 0, off_b = 0;
    {
        const int base0 = (G == 16) ? ((lg >> 3) & 1) * 24 + ((lg >> 2) & 1) * 12 + ((lg >> 1) & 1) * 6 + (lg & 1) * 3
                                    : ((lg >> 2) & 1) * 24 + ((lg >> 1) & 1) * 12 + (lg & 1) * 6;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            off_a |= (uint64_t)c_pk_a[base0 + i] << (8 * i);
            off_b |= (uint64_t)c_pk_b[base0 + i] << (8 * i);
        }
    }

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
        float acc[NACC];
        {
            GramAcc ga;
            ga.clear();
#pragma unroll
            for (int r = 0; r < RPL; ++r) ga.add(f[r], e[r]);
            ga.finish(acc);
        }
        group_reduce_scatter48<G>(acc, lg);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            tot[(uint32_t)(off_a >> (8 * i)) & 0xffu] = acc[i];
            tot[(uint32_t)(off_b >> (8 * i)) & 0xffu] = acc[i];
        }
        __syncwarp();
        const float z = __shfl_sync(0xffffffffu, zq, ((b % ZB) << 3) + kq, G);
        float uo8[8];
        if (REFRESH) {
#pragma unroll
            for (int k = 0; k < 8; ++k) uo8[k] = __shfl_sync(0xffffffffu, uo, k, 8);
        }
        // every octet solves its group's row (for G == 16 both octets of the group compute the same values); an idle group's
        // lanes report lane index 8 so that they never take the write branch
        const SolveOut so = solve_lanes(tot, a, foff, uo, sig, mu, live, mode, z, alpha, have_row ? lg : 8);
        if (kPairedDots && REFRESH) {
            float2 du[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) du[k] = make_float2(so.d[k], uo8[k]);
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const float2 s2 = dot8_pair(f[r], du);
                e[r] += s2.x;
                pr[r] += s2.y - s2.x;
            }
        } else {
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const float fd = dot8(f[r], so.d);
                e[r] += fd;
                if (REFRESH) {
                    float fu = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) fu = fmaf(f[r].v[k], uo8[k], fu);
                    pr[r] += fu - fd;
                }
            }
        }
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
#include "rows2.cuh"   // second-generation resident-row kernels (option row_kernels = 2)
namespace sbmf {

// --------------------------------------------------------------------------------------------------------
// Heavy rows: sliced streaming pipeline.  accumulate<PREV,CUR> applies the pending delta of the previous step
// to e (PREV: 0 = only the global shift, 1 = bias delta, 2 = delta of block pb, re-gathering f) and accumulates
// the partial sums of the current step (CUR: 0 = nothing, 1 = bias: sum e, 2 = block b: g, G) per slice;
// heavy_solve<CUR> combines a row's slices in slice order and performs the update(s).
// The update(s) of one streamed row, executed by the LAST slice CTA of the row to finish a pass (heavy_accumulate_kernel):
// warp w sums the partials of slices s0 + w, s0 + w + NW, ... (independent coalesced 192-byte reads, through L2: the lines were
// written by other SMs during this launch), the sums are combined in warp order -- a fixed tree whatever CTA happens to be last,
// so results are reproducible -- and warp 0 performs the update(s).  CUR == 1: bias half-step; CUR == 2: factor block b.
#ifndef SBMF_SOLVE_INLINE
#define SBMF_SOLVE_INLINE __forceinline__
#endif
template <int CUR, int NW>
__device__ SBMF_SOLVE_INLINE void heavy_row_solve(const PhaseArgs& a, uint32_t hrow, uint32_t s0, uint32_t s1, const float* __restrict__ hpart,
                                                float* __restrict__ hdelta, float* __restrict__ hbias_delta, int b, float (*s_part)[NACC], float* s_tot)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t row = a.rows_of_hrow[hrow];
    const uint32_t rid = a.row_id ? a.row_id[row] : row;   // the caller's id: key of the draws
    if (CUR == 1) {
        float t = 0.f;
        for (uint32_t s = s0 + threadIdx.x; s < s1; s += NW * 32) t += __ldcg(hpart + (size_t)s * NACC);
        t = warp_sum(t);
        if (lane == 0) s_part[warp][0] = t;
        __syncthreads();
        if (warp != 0) return;
        t = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) t += s_part[w][0];
        const float alpha = a.sc->alpha_f;
        const uint32_t sweep = a.sc->sweep;
        const float c = (float)(a.ptr[row + 1] - a.ptr[row]);
        const float bo = a.bias[row], sb = a.sigma_b[row], mb = a.mu_b[row];
        const float s = 1.0f / (sb + alpha * c);
        const float mean = s * (sb * mb + alpha * (t + c * bo));
        float z = 0.f;
        if (a.mode != SAMPLE_ZERO) z = normal_f32(philox_site(a.seed, a.site_b, rid, 0u, sweep));
        const float bn = draw_f32(a.mode, mean, s, z);
        __syncwarp();   // every lane has read the old bias (all lanes compute the same value, lane 0 stores it)
        if (lane == 0) {
            for (int q = 0; q < a.nrep; ++q) a.brep[q][row] = bn;
            hbias_delta[hrow] = bo - bn;
        }
    } else {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
        for (uint32_t s = s0 + warp; s < s1; s += NW) {
            a0 += __ldcg(hpart + (size_t)s * NACC + lane);
            if (lane < NACC - 32) a1 += __ldcg(hpart + (size_t)s * NACC + 32 + lane);
        }
        s_part[warp][lane] = a0;
        if (lane < NACC - 32) s_part[warp][32 + lane] = a1;
        __syncthreads();
        if (warp != 0) return;
        const float alpha = a.sc->alpha_f;
        const uint32_t sweep = a.sc->sweep;
        float zl = 0.f;
        if (a.mode != SAMPLE_ZERO && lane < 8) zl = normal_f32(philox_site(a.seed, a.site_f, rid, (uint32_t)(b * 8 + lane), sweep));
        a0 = 0.f;
        a1 = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            a0 += s_part[w][lane];
            if (lane < NACC - 32) a1 += s_part[w][32 + lane];
        }
        s_tot[lane] = a0;
        if (lane < NACC - 32) s_tot[32 + lane] = a1;
        __syncwarp();
        float z[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) z[k] = __shfl_sync(0xffffffffu, zl, k);
        const size_t foff = ((size_t)b * a.ns_self + row) * 8;
        const f8 uo = ld256(a.Fself + foff);
        f8 un;
        float d[8];
        solve_block(s_tot, uo, z, a.sigma_kf + b * 8, a.mu_kf + b * 8, alpha, a.mode, (int)a.K - b * 8, un, d);
        __syncwarp();   // every lane has read the old factors
        if (lane == 0) {
            for (int q = 0; q < a.nrep; ++q) st256(a.Frep[q] + foff, un);
#pragma unroll
            for (int k = 0; k < 8; ++k) hdelta[(size_t)hrow * 8 + k] = d[k];
        }
    }
}

// one ticket per slice CTA: fetch-and-add with release semantics at GPU scope (MEMBAR.ALL.GPU + ATOMG by the calling thread only)
__device__ __forceinline__ uint32_t ticket_release(uint32_t* counter)
{
#ifdef SBMF_SIMT_EMU
    __threadfence();
    return atomicAdd(counter, 1u);
#else
    uint32_t t;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(t) : "l"(counter) : "memory");
    return t;
#endif
}

// The row updates as a launch of their own, one CTA per streamed row (option fuse_solve = 0: A/B against the fused tail).
template <int CUR, int NW>
__global__ void __launch_bounds__(NW * 32)
heavy_solve_kernel(PhaseArgs a, const uint32_t* __restrict__ slice_ptr, const float* __restrict__ hpart, float* __restrict__ hdelta,
                   float* __restrict__ hbias_delta, int b, uint32_t hrow0)
{
    __shared__ __align__(16) float s_part[NW][NACC];
    __shared__ __align__(16) float s_tot[NACC];
    const uint32_t hrow = blockIdx.x + hrow0;   // hrow0: first streamed row of this chain
    heavy_row_solve<CUR, NW>(a, hrow, slice_ptr[hrow], slice_ptr[hrow + 1], hpart, hdelta, hbias_delta, b, s_part, s_tot);
}

// PAIR (PREV == 2 and CUR == 2 only): the factor rows come from the pair array F2other.  Lanes (2k, 2k+1) fetch the 64-byte row
// of the even lane's rating with one 256-bit load each (lower half = block pb to the even lane, upper half = block b to the odd
// lane), then the row of the odd lane's rating with the halves swapped.  Every lane thus holds the previous block of its OWN
// rating (delta apply, residual store) and the current block of its PARTNER's rating, whose updated residual arrives by one
// shuffle for the accumulation.  Which load delivered which is a matter of lane parity: 16 selects per rating, paid from the
// issue slots this gather-bound kernel leaves idle.
template <int PREV, int CUR, int UNR, int THREADS, bool REFRESH, bool PAIR = false, bool FUSE = true>
__global__ void __launch_bounds__(THREADS, REFRESH ? 8 : 10)   // 64-thread CTAs: 96 (124 with the prediction refresh) registers per thread
heavy_accumulate_kernel(PhaseArgs a, const Slice* __restrict__ slices, float* __restrict__ hdelta, float* __restrict__ hbias_delta,
                        float* __restrict__ hpart, const uint32_t* __restrict__ slice_ptr, uint32_t* __restrict__ hcount, int pb, int b)
{
    __shared__ __align__(16) float s_part[THREADS / 32][NACC];
    __shared__ __align__(16) float s_tot[NACC];
    __shared__ uint32_t s_last;
    const Slice sl = slices[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float dprev[8];
    float dscalar = 0.f;
    if (PREV == 0) dscalar = a.apply_shift ? a.sc->shift_f : 0.f;
    if (PREV == 1) dscalar = hbias_delta[sl.hrow];
    if (PREV == 2) {
#pragma unroll
        for (int k = 0; k < 8; ++k) dprev[k] = hdelta[(size_t)sl.hrow * 8 + k];
    }
    // REFRESH: the previous block's NEW factor values of this row (the solve kernel has stored them), for p += <f_prev, u_new_prev>
    float unprev[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float row_const = 0.f;
    if (REFRESH && PREV == 2) {
        const uint32_t row = a.rows_of_hrow[sl.hrow];
        const float* Fsp = a.Fself + ((size_t)pb * a.ns_self + row) * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) unprev[k] = Fsp[k];
        if (CUR == 0) row_const = a.sc->b_0_f + a.bias[row];
    }
    float2 dun[8];   // (d_prev[k], u_new_prev[k]) pairs of the packed form
#pragma unroll
    for (int k = 0; k < 8; ++k) dun[k] = make_float2((PREV == 2) ? dprev[k] : 0.f, unprev[k]);
    float* pp = a.pacc + sl.start;
    const float* rp = a.r + sl.start;
    const float* Fp = a.Fother + (size_t)pb * a.ns_other * 8;
    const float* Fc = a.Fother + (size_t)b * a.ns_other * 8;
    const uint32_t pad_row = a.ns_other - 1;
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;   // (CUR == 1 sums into acc[0]; CUR == 2 fills it from ga after the loop)
    GramAcc ga;
    ga.clear();
    const uint32_t* idx = a.idx + sl.start;
    float* ep = a.e + sl.start;
    if constexpr (PAIR) {
        static_assert(!PAIR || (PREV == 2 && CUR == 2), "the pair form serves the block-to-block step");
        const float* F2 = a.F2other + (size_t)pb * a.ns_other * 16;
        const int odd = lane & 1;
        for (uint32_t base0 = 0; base0 < sl.len; base0 += THREADS * UNR) {   // warp-uniform trip count: the shuffles need every lane
            uint32_t id[UNR];
            float e[UNR];
            float pr[REFRESH ? UNR : 1];
            bool ok[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const uint32_t i = base0 + u * THREADS + threadIdx.x;
                ok[u] = i < sl.len;
                id[u] = ok[u] ? idx[i] : pad_row;
                e[u] = ok[u] ? ep[i] : 0.f;
                if (REFRESH) pr[u] = (ok[u] && pb > 0) ? pp[i] : 0.f;
            }
            f8 r1[UNR], r2[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const uint32_t idp = __shfl_xor_sync(0xffffffffu, id[u], 1);
                const uint32_t id_even = odd ? idp : id[u], id_odd = odd ? id[u] : idp;
                r1[u] = ld256_nc(F2 + (size_t)id_even * 16 + odd * 8);         // even lane: own previous block; odd lane: partner's current block
                r2[u] = ld256_nc(F2 + (size_t)id_odd * 16 + (odd ^ 1) * 8);   // even lane: partner's current block; odd lane: own previous block
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const uint32_t i = base0 + u * THREADS + threadIdx.x;
                f8 fprev, fcur;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    fprev.v[k] = odd ? r2[u].v[k] : r1[u].v[k];
                    fcur.v[k] = odd ? r1[u].v[k] : r2[u].v[k];
                }
                if (kPairedDots && REFRESH) {
                    const float2 s2 = dot8_pair(fprev, dun);
                    e[u] += s2.x;
                    pr[u] += s2.y;
                } else {
                    e[u] += dot8(fprev, dprev);
                    if (REFRESH) pr[u] += dot8(fprev, unprev);
                }
                if (ok[u]) {
                    ep[i] = e[u];
                    if (REFRESH) pp[i] = pr[u];
                }
                // the partner's rating is the one whose current block this lane holds (an invalid slot carries e = 0 and the zero pad row)
                const float e_partner = __shfl_xor_sync(0xffffffffu, e[u], 1);
                ga.add(fcur, e_partner);
            }
        }
    } else
    // (the general form)
    // Batches of UNR ratings per thread, software-pipelined by one batch: the index / residual (/ partial prediction) words of
    // batch t + 1 are requested right after the gathers of batch t have been issued, so that an iteration's critical path is one
    // L2 round trip (the gathers) instead of a DRAM round trip (the index stream) FOLLOWED by the gathers -- with ~19 warps per SM
    // the serial chain, not the L1TEX rate, was what a pass ran at (ncu: issue active 18 %, long scoreboard 21 per issue).
    {
        uint32_t idn[UNR];
        float en[UNR];
        float prn[REFRESH ? UNR : 1];
        auto fetch = [&](uint32_t base) {
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const uint32_t i = base + u * THREADS;
                const bool ok = i < sl.len;
                idn[u] = ok ? ld_stream(idx + i) : pad_row;
                en[u] = ok ? ((PREV == 0) ? load_e_first(a, (int64_t)sl.start + i) : ld_stream(ep + i)) : 0.f;
                if (REFRESH && PREV == 2) prn[u] = (ok && pb > 0) ? ld_stream(pp + i) : 0.f;
            }
        };
        fetch(threadIdx.x);
        for (uint32_t base = threadIdx.x; base < sl.len; base += THREADS * UNR) {
            uint32_t id[UNR];
            float e[UNR];
            float pr[REFRESH ? UNR : 1];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                id[u] = idn[u];
                e[u] = en[u];
                if (REFRESH && PREV == 2) pr[u] = prn[u];
            }
            f8 fp[PREV == 2 ? UNR : 1], fc[CUR == 2 ? UNR : 1];
            if (PREV == 2) {
#pragma unroll
                for (int u = 0; u < UNR; ++u) fp[u] = ld256_nc(Fp + (size_t)id[u] * 8);
            }
            if (CUR == 2) {
#pragma unroll
                for (int u = 0; u < UNR; ++u) fc[u] = ld256_nc(Fc + (size_t)id[u] * 8);
            }
            fetch(base + THREADS * UNR);   // next batch (slots beyond the slice read nothing); other slots than the ones stored below
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const uint32_t i = base + u * THREADS;
                if (kPairedDots && REFRESH && PREV == 2) {   // (<f_prev, d_prev>, <f_prev, u_new_prev>) as one FFMA2 chain
                    const float2 s2 = dot8_pair(fp[u], dun);
                    e[u] += s2.x;
                    pr[u] += s2.y;
                } else {
                    if (PREV == 2) e[u] += dot8(fp[u], dprev);
                    else e[u] += dscalar;
                    if (REFRESH && PREV == 2) pr[u] += dot8(fp[u], unprev);
                }
                if (REFRESH && PREV == 2) {
                    if (CUR == 0) e[u] = (i < sl.len) ? ld_stream(rp + i) - (row_const + a.bias_other[id[u]] + pr[u]) : 0.f;   // phase done: fresh residual
                    else if (i < sl.len) st_stream(pp + i, pr[u]);
                }
                if (i < sl.len) {
                    if (CUR == 0) store_e_final(a, (int64_t)sl.start + i, e[u]);   // last pass of the phase
                    else st_stream(ep + i, e[u]);
                }
                if (CUR == 1 && i < sl.len) acc[0] += e[u];
                if (CUR == 2) ga.add(fc[u], e[u]);
            }
        }
    }
    if (CUR == 1) {
        const float t = warp_sum(acc[0]);
        if (lane == 0) s_part[warp][0] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < THREADS / 32; ++w) s += s_part[w][0];
            hpart[(size_t)blockIdx.x * NACC] = s;
        }
    }
    if (CUR == 2) {
        ga.finish(acc);
        warp_reduce_scatter48(acc, lane);
        const int base = reduce_scatter_base(lane);
        if ((lane & 1) == 0) {
            s_part[warp][base] = acc[0];
            s_part[warp][base + 1] = acc[1];
            s_part[warp][base + 2] = acc[2];
        }
        __syncthreads();
        if (threadIdx.x < NACC) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < THREADS / 32; ++w) s += s_part[w][threadIdx.x];
            hpart[(size_t)blockIdx.x * NACC + threadIdx.x] = s;
        }
    }
    if constexpr (CUR != 0 && FUSE) {   // FUSE == false: the updates run as a launch of their own (heavy_solve_kernel) and this tail does not exist
        // The row's update runs in the tail of this launch: every slice CTA publishes its partial, takes a ticket, and the CTA
        // that draws the last ticket of its row combines the partials (in slice order) and solves.  All CTAs of the row have
        // read the pending delta / bias delta of the previous step before they take a ticket, so the solve may overwrite them.
        // The ticket is a RELEASE atomic of thread 0 after the barrier (cumulative: it orders the partial stores of the whole CTA).
        // Not __threadfence() by every thread: on sm_100 that is MEMBAR.SC + CCTL.IVALL, i.e. every finishing CTA would also
        // invalidate its SM's L1 under the gathers of the CTAs still running (measured: 806 -> 902 us per pass).
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t n_row = slice_ptr[sl.hrow + 1] - slice_ptr[sl.hrow];
            const uint32_t ticket = ticket_release(&hcount[sl.hrow]);
            s_last = (ticket == n_row - 1) ? 1u : 0u;
            if (s_last) hcount[sl.hrow] = 0u;   // re-armed for the next pass (nobody else touches it before the next launch)
        }
        __syncthreads();
        if (s_last) {
            __threadfence();   // acquire side, once per row and pass; the partials are then read through L2 (__ldcg)
            heavy_row_solve<CUR, THREADS / 32>(a, sl.hrow, slice_ptr[sl.hrow], slice_ptr[sl.hrow + 1], hpart, hdelta, hbias_delta, b, s_part, s_tot);
        }
    }
}

// --------------------------------------------------------------------------------------------------------
// Residual rebuild [T]:342-359 in CSR slot order, with per-block partial statistics.
constexpr int RED_THREADS = 256;

__device__ __forceinline__ void block_reduce2_store(double s1, double s2, double* out)
{
    __shared__ double sh[2][RED_THREADS / 32];
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sh[0][warp] = s1;
        sh[1][warp] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int w = 0; w < RED_THREADS / 32; ++w) {
            t1 += sh[0][w];
            t2 += sh[1][w];
        }
        out[0] = t1;
        out[1] = t2;
    }
}

// I, J here are the factor strides (rows + 1 pad row)
__device__ __forceinline__ float dot_blocked(const float* __restrict__ Fu, const float* __restrict__ Fv, uint32_t I, uint32_t J, uint32_t KB,
                                             uint32_t u, uint32_t j)
{
    float dot = 0.f;
#pragma unroll 4
    for (uint32_t b = 0; b < KB; ++b) {
        const f8 fu = ld256_nc(Fu + ((size_t)b * I + u) * 8);
        const f8 fv = ld256_nc(Fv + ((size_t)b * J + j) * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) dot = fmaf(fu.v[k], fv.v[k], dot);
    }
    return dot;
}

__global__ void __launch_bounds__(RED_THREADS)
rebuild_kernel(const uint32_t* __restrict__ urow, const uint32_t* __restrict__ col, const float* __restrict__ r, float* __restrict__ e,
               const float* __restrict__ Fu, const float* __restrict__ Fv, const float* __restrict__ bi, const float* __restrict__ bj,
               const Scalars* __restrict__ sc, uint32_t I, uint32_t J, uint32_t KB, uint64_t N, double* __restrict__ part)
{
    const float b0 = sc->b_0_f;
    double s1 = 0.0, s2 = 0.0;
    for (uint64_t s = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x; s < N; s += (uint64_t)gridDim.x * RED_THREADS) {
        const uint32_t u = urow[s], j = col[s];
        const float dot = dot_blocked(Fu, Fv, I, J, KB, u, j);
        const float ev = r[s] - (b0 + bi[u] + bj[j] + dot);
        e[s] = ev;
        s1 += (double)ev;
        s2 += (double)ev * (double)ev;
    }
    block_reduce2_store(s1, s2, part + (size_t)blockIdx.x * 2);
}

__global__ void __launch_bounds__(RED_THREADS)
stats_kernel(const float* __restrict__ e, uint64_t N, double* __restrict__ part)
{
    double s1 = 0.0, s2 = 0.0;
    for (uint64_t s = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x; s < N; s += (uint64_t)gridDim.x * RED_THREADS) {
        const double ev = (double)e[s];
        s1 += ev;
        s2 += ev * ev;
    }
    block_reduce2_store(s1, s2, part + (size_t)blockIdx.x * 2);
}

// final reduce of the statistics + the scalar chain alpha, sigma_b_0, mu_b_0, b_0 of [T]:366-410
__global__ void __launch_bounds__(RED_THREADS)
reduce_pair_kernel(const double* __restrict__ part, uint32_t nparts, double* __restrict__ out)
{
    double s1 = 0.0, s2 = 0.0;
    for (uint32_t i = threadIdx.x; i < nparts; i += RED_THREADS) {
        s1 += part[(size_t)i * 2];
        s2 += part[(size_t)i * 2 + 1];
    }
    block_reduce2_store(s1, s2, out);
}

__global__ void __launch_bounds__(32)
global_hypers_kernel(Scalars* sc, const double* __restrict__ red2, uint64_t N, sbmf_priors pr, int mode, int hyper_mode, uint64_t seed)
{
    if (threadIdx.x != 0) return;
    const double S1 = red2[0], S2 = red2[1];
    const uint32_t sweep = sc->sweep;
    const double Nd = (double)N;
    sc->sum_e = S1;
    sc->sum_e2 = S2;
    if (hyper_mode != SBMF_HYPER_REF_T) {   // [S]:339-342; global mean and biases do not exist in [S]
        const double tau = draw_gamma_f64(mode, seed, SITE_ALPHA, 0, sweep, pr.ng_a_0 + 0.5 * Nd, pr.ng_b_0 + 0.5 * S2);
        sc->alpha = tau;
        sc->alpha_f = (float)tau;
        sc->shift_f = 0.f;
        return;
    }
    const double alpha = draw_gamma_f64(mode, seed, SITE_ALPHA, 0, sweep, pr.alpha_dash + Nd, pr.beta_dash + S2);
    double b0 = sc->b_0, mu_b0 = sc->mu_b_0;
    const double sigma_b0 = draw_gamma_f64(mode, seed, SITE_SIGMA_B0, 0, sweep, pr.alpha[0] + 1.0, pr.beta[0] + 0.5 * (b0 - mu_b0) * (b0 - mu_b0));
    {
        const double s = 1.0 / (pr.sigma[0] + sigma_b0);
        const double m = s * (pr.sigma[0] * pr.mu[0] + b0 * sigma_b0);
        mu_b0 = draw_gauss_f64(mode, seed, SITE_MU_B0, 0, 0, sweep, m, s);
    }
    {
        const double s = 1.0 / (sigma_b0 + alpha * Nd);
        const double m = s * (sigma_b0 * mu_b0 + alpha * (S1 + Nd * b0));
        const double old = b0;
        b0 = draw_gauss_f64(mode, seed, SITE_B0, 0, 0, sweep, m, s);
        sc->shift_f = (float)(old - b0);
    }
    sc->alpha = alpha;
    sc->sigma_b_0 = sigma_b0;
    sc->mu_b_0 = mu_b0;
    sc->b_0 = b0;
    sc->alpha_f = (float)alpha;
    sc->b_0_f = (float)b0;
}

// --------------------------------------------------------------------------------------------------------
// Per-dimension hyper-parameters [T]:415-467.  Stage 1: (sum, sum of squared deviations from the OLD mean) per
// chunk of rows for the 8 dimensions of a block; stage 2: ordered sum over chunks + the Gamma / Normal draws.
constexpr int HYP_CHUNK = 16384;

__global__ void __launch_bounds__(256)
dim_hyper_partial_kernel(const float* __restrict__ F, uint32_t n, uint32_t ns, const double* __restrict__ mu_k, double* __restrict__ part, uint32_t chunks)
{
    __shared__ double sh[8][16];
    const uint32_t b = blockIdx.y, chunk = blockIdx.x;
    double mu[8], S[8], SS[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        mu[k] = mu_k[b * 8 + k];
        S[k] = 0.0;
        SS[k] = 0.0;
    }
    const uint32_t r0 = chunk * HYP_CHUNK;
    const uint32_t r1 = min(n, r0 + HYP_CHUNK);
    for (uint32_t r = r0 + threadIdx.x; r < r1; r += 256) {
        const f8 f = ld256(F + ((size_t)b * ns + r) * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double u = (double)f.v[k];
            S[k] += u;
            SS[k] += (u - mu[k]) * (u - mu[k]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        S[k] = warp_sum(S[k]);
        SS[k] = warp_sum(SS[k]);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            sh[warp][k] = S[k];
            sh[warp][8 + k] = SS[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
        part[((size_t)b * chunks + chunk) * 16 + threadIdx.x] = t;
    }
}

__global__ void __launch_bounds__(32)
dim_hyper_final_kernel(const double* __restrict__ part, uint32_t chunks, uint32_t n, uint32_t K, double* sigma_k, double* mu_k, float* sigma_kf,
                       float* mu_kf, const Scalars* sc, double pa, double pb, double pmu, double psigma, int mode, uint64_t seed,
                       uint32_t site_sigma, uint32_t site_mu, int hyper_mode, sbmf_priors pr, double* post_var, const double* lead_var)
{
    const uint32_t b = blockIdx.x, k = threadIdx.x;
    if (k >= 8) return;
    const uint32_t kk = b * 8 + k;
    if (kk >= K) {
        sigma_k[kk] = 1.0;
        mu_k[kk] = 0.0;
        sigma_kf[kk] = 1.f;
        mu_kf[kk] = 0.f;
        return;
    }
    double S = 0.0, SS = 0.0;
    for (uint32_t c = 0; c < chunks; ++c) {
        S += part[((size_t)b * chunks + c) * 16 + k];
        SS += part[((size_t)b * chunks + c) * 16 + 8 + k];
    }
    const uint32_t sweep = sc->sweep;
    const double nd = (double)n;
    if (hyper_mode != SBMF_HYPER_REF_T) {   // Normal-Gamma step of [S]:383-413
        const double mo = mu_k[kk];
        const double a = pr.ng_alpha_0 + 0.5 * (nd + 1.0);
        const double b = pr.ng_beta_0 + pr.ng_nu_0 * (mo - pr.ng_mu_0) * (mo - pr.ng_mu_0) + 0.5 * SS;
        const double sg = draw_gamma_f64(mode, seed, site_sigma, kk, sweep, a, b);
        const double s = 1.0 / (pr.ng_nu_0 * sg + sg * nd);
        post_var[kk] = s;
        // [S]:412 scales the item-side mean by the USER side's posterior variance; lead_var != NULL reproduces that
        const double lead = lead_var ? lead_var[kk] : s;
        const double mn = lead * (pr.ng_nu_0 * pr.ng_mu_0 * sg + sg * S);
        const double mun = draw_gauss_f64(mode, seed, site_mu, kk, 0, sweep, mn, s);
        sigma_k[kk] = sg;
        mu_k[kk] = mun;
        sigma_kf[kk] = (float)sg;
        mu_kf[kk] = (float)mun;
        return;
    }
    const double sig = draw_gamma_f64(mode, seed, site_sigma, kk, sweep, pa + nd, pb + 0.5 * SS);
    const double s = 1.0 / (psigma + sig * nd);
    const double m = s * (psigma * pmu + sig * S);
    const double mu = draw_gauss_f64(mode, seed, site_mu, kk, 0, sweep, m, s);
    sigma_k[kk] = sig;
    mu_k[kk] = mu;
    sigma_kf[kk] = (float)sig;
    mu_kf[kk] = (float)mu;
}

// Per-row bias hyper-parameters [T]:469-511.
__global__ void __launch_bounds__(256)
bias_hyper_kernel(const float* __restrict__ bias, float* __restrict__ mu_b, float* __restrict__ sigma_b, uint32_t n, const Scalars* sc,
                  double pa, double pb, double pmu, double psigma, int mode, uint64_t seed, uint32_t site_sigma, uint32_t site_mu,
                  const uint32_t* __restrict__ row_id)
{
    const uint32_t r = blockIdx.x * 256 + threadIdx.x;
    if (r >= n) return;
    const uint32_t rid = row_id ? row_id[r] : r;   // draws are keyed by the caller's row id
    const uint32_t sweep = sc->sweep;
    const double b = (double)bias[r], mo = (double)mu_b[r];
    const double sig = draw_gamma_f64(mode, seed, site_sigma, rid, sweep, pa + 1.0, pb + 0.5 * (b - mo) * (b - mo));
    const double s = 1.0 / (psigma + sig);
    const double m = s * (psigma * pmu + b * sig);
    const double mu = draw_gauss_f64(mode, seed, site_mu, rid, 0, sweep, m, s);
    sigma_b[r] = (float)sig;
    mu_b[r] = (float)mu;
}

// --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
permute_gather_kernel(const float* __restrict__ src, const uint32_t* __restrict__ perm, float* __restrict__ dst, uint64_t N)
{
    for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < N; t += (uint64_t)gridDim.x * 256) dst[t] = src[perm[t]];
}
__global__ void __launch_bounds__(256)
permute_scatter_kernel(const float* __restrict__ src, const uint32_t* __restrict__ perm, float* __restrict__ dst, uint64_t N)
{
    for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < N; t += (uint64_t)gridDim.x * 256) dst[perm[t]] = src[t];
}

// --------------------------------------------------------------------------------------------------------
// Test prediction + RMSE of the running posterior mean [T]:610-636.
__global__ void __launch_bounds__(RED_THREADS)
eval_kernel(const uint32_t* __restrict__ tu, const uint32_t* __restrict__ ti, const float* __restrict__ tr, double* __restrict__ tsum,
            const float* __restrict__ Fu, const float* __restrict__ Fv, const float* __restrict__ bi, const float* __restrict__ bj,
            const Scalars* __restrict__ sc, uint32_t I, uint32_t J, uint32_t KB, uint64_t Nt, uint32_t burn_in, float lo, float hi,
            double* __restrict__ part)
{
    const float b0 = sc->b_0_f;
    const uint32_t sweep = sc->sweep;
    const bool collect = sweep >= burn_in;
    const double denom = collect ? (double)(sweep + 1 - burn_in) : 1.0;
    double s1 = 0.0, s2 = 0.0;
    for (uint64_t t = (uint64_t)blockIdx.x * RED_THREADS + threadIdx.x; t < Nt; t += (uint64_t)gridDim.x * RED_THREADS) {
        const uint32_t u = tu[t], j = ti[t];
        float p = b0 + bi[u] + bj[j] + dot_blocked(Fu, Fv, I, J, KB, u, j);
        p = fminf(hi, p);
        p = fmaxf(lo, p);
        const double target = (double)tr[t];
        double mean = (double)p;
        if (collect) {
            const double sm = tsum[t] + (double)p;
            tsum[t] = sm;
            mean = sm / denom;
        }
        s1 += (target - mean) * (target - mean);
        s2 += (target - (double)p) * (target - (double)p);
    }
    block_reduce2_store(s1, s2, part + (size_t)blockIdx.x * 2);
}

__global__ void __launch_bounds__(32)
eval_final_kernel(Scalars* sc, const double* __restrict__ out, uint64_t Nt, double* __restrict__ hist, uint32_t hist_cap)
{
    if (threadIdx.x != 0) return;
    const uint32_t sweep = sc->sweep;
    const double rm = Nt ? sqrt(out[0] / (double)Nt) : 0.0;
    const double rs = Nt ? sqrt(out[1] / (double)Nt) : 0.0;
    sc->rmse_mean = rm;
    sc->rmse_sweep = rs;
    if (sweep < hist_cap) {
        hist[(size_t)sweep * 2] = rm;
        hist[(size_t)sweep * 2 + 1] = rs;
    }
    sc->sweep = sweep + 1;
}

// --------------------------------------------------------------------------------------------------------
// Factor layout conversion / initialisation.
__global__ void __launch_bounds__(256)
init_factors_kernel(float* __restrict__ F, uint32_t n, uint32_t K, uint32_t KB, uint64_t seed, uint32_t site, float stdev,
                    const uint32_t* __restrict__ row_id)
{
    const uint32_t nreal = n - 1;   // n = stride = rows + 1 zero pad row
    const uint64_t total = (uint64_t)KB * n * 8;
    for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (uint64_t)gridDim.x * 256) {
        const uint32_t k8 = (uint32_t)(t & 7);
        const uint64_t br = t >> 3;
        const uint32_t row = (uint32_t)(br % n), b = (uint32_t)(br / n);
        const uint32_t k = b * 8 + k8;
        F[t] = (k < K && row < nreal) ? stdev * normal_f32(philox_site(seed, site, row_id ? row_id[row] : row, k, 0u)) : 0.f;
    }
}

__global__ void __launch_bounds__(256)
load_factors_kernel(float* __restrict__ F, const float* __restrict__ src, uint32_t n, uint32_t K, uint32_t KB, int dim_major,
                    const uint32_t* __restrict__ row_id)
{
    const uint32_t nreal = n - 1;   // n = stride = rows + 1 zero pad row
    const uint64_t total = (uint64_t)KB * n * 8;
    for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (uint64_t)gridDim.x * 256) {
        const uint32_t k8 = (uint32_t)(t & 7);
        const uint64_t br = t >> 3;
        const uint32_t row = (uint32_t)(br % n), b = (uint32_t)(br / n);
        const uint32_t k = b * 8 + k8;
        float v = 0.f;
        if (k < K && row < nreal) {
            const uint32_t sr = row_id ? row_id[row] : row;   // src is in the caller's row order
            v = dim_major ? src[(size_t)k * nreal + sr] : src[(size_t)sr * K + k];
        }
        F[t] = v;
    }
}

__global__ void __launch_bounds__(256)
export_factors_kernel(const float* __restrict__ F, float* __restrict__ dst, uint32_t n, uint32_t K, int dim_major,
                      const uint32_t* __restrict__ pos_of)
{
    const uint64_t total = (uint64_t)n * K;
    for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (uint64_t)gridDim.x * 256) {
        uint32_t row, k;
        if (dim_major) {
            k = (uint32_t)(t / n);
            row = (uint32_t)(t % n);
        } else {
            row = (uint32_t)(t / K);
            k = (uint32_t)(t % K);
        }
        dst[t] = F[((size_t)(k >> 3) * (n + 1) + (pos_of ? pos_of[row] : row)) * 8 + (k & 7)];
    }
}

// Pair array of one side (PhaseArgs::F2other): F2[pb][row][0..7] = F[pb][row], F2[pb][row][8..15] = F[pb + 1][row].  Rebuilt from
// the K8 blocks at the start of every phase whose streaming pipeline gathers this side (the rows were all rewritten by the
// previous phase): one streaming pass, ~0.1 ms for the Netflix-shaped user side, saved many times over by the pipeline.
__global__ void __launch_bounds__(256)
pair_pack_kernel(const float* __restrict__ F, float* __restrict__ F2, uint32_t ns, uint32_t npairs)
{
    const uint64_t total = (uint64_t)npairs * ns;
    for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (uint64_t)gridDim.x * 256) {
        const f8 lo = ld256(F + t * 8);                        // (pb, row) -> F[pb][row]
        const f8 hi = ld256(F + (t + ns) * 8);                 //           -> F[pb + 1][row]
        st256(F2 + t * 16, lo);
        st256(F2 + t * 16 + 8, hi);
    }
}

// ========================================================================================================
// launch wrappers
void init_constant_tables()
{
    uint8_t pa[NACC], pb[NACC];
    for (int p = 0; p < NACC; ++p) pa[p] = pb[p] = (uint8_t)(8 + 8 * G_STRIDE + (p & 3));   // padding accumulators -> slack
    for (int k = 0; k < 8; ++k) pa[k] = pb[k] = (uint8_t)k;
    for (int k = 0; k < 8; ++k)
        for (int l = k; l < 8; ++l) {
            pa[gi(k, l)] = (uint8_t)(8 + G_STRIDE * k + l);
            pb[gi(k, l)] = (uint8_t)(8 + G_STRIDE * l + k);
        }
    cudaMemcpyToSymbol(c_pk_a, pa, NACC);
    cudaMemcpyToSymbol(c_pk_b, pb, NACC);
    // the same two offsets for the accumulators in the order the build's GramAcc holds them (rows2.cuh)
    uint8_t na[NACC], nb[NACC];
    for (int n = 0; n < NACC; ++n) {
        int k = 0, l = 0;
        const int kind = native_entry(n, k, l);
        if (kind == 0) na[n] = nb[n] = (uint8_t)k;
        else if (kind == 1) {
            na[n] = (uint8_t)(8 + G_STRIDE * k + l);
            nb[n] = (uint8_t)(8 + G_STRIDE * l + k);
        } else na[n] = nb[n] = (uint8_t)(8 + 8 * G_STRIDE + (n & 3));
    }
    cudaMemcpyToSymbol(c_nat_a, na, NACC);
    cudaMemcpyToSymbol(c_nat_b, nb, NACC);
}

static inline uint32_t grid_for(uint64_t n, int threads, int cap)
{
    uint64_t g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > (uint64_t)cap) g = cap;
    return (uint32_t)g;
}

void launch_init_factors(Model& m, Side& s, uint32_t site, cudaStream_t st)
{
    const uint64_t total = (uint64_t)m.KB * (s.n + 1) * 8;
    SBMF_LAUNCH((init_factors_kernel), grid_for(total, 256, m.sm_count * 16), 256, 0, st, s.F, s.n + 1, m.K, m.KB, m.cfg.seed, site, (float)m.cfg.init_stdev, s.id_at);
    m.launches++;
}

void launch_load_factors(Model& m, Side& s, const float* d_src, bool dim_major, cudaStream_t st)
{
    const uint64_t total = (uint64_t)m.KB * (s.n + 1) * 8;
    SBMF_LAUNCH((load_factors_kernel), grid_for(total, 256, m.sm_count * 16), 256, 0, st, s.F, d_src, s.n + 1, m.K, m.KB, dim_major ? 1 : 0, s.id_at);
    m.launches++;
}

void launch_export_factors(Model& m, const Side& s, float* d_out, bool dim_major, cudaStream_t st)
{
    const uint64_t total = (uint64_t)s.n * m.K;
    SBMF_LAUNCH((export_factors_kernel), grid_for(total, 256, m.sm_count * 16), 256, 0, st, s.F, d_out, s.n, m.K, dim_major ? 1 : 0, s.pos_of);
    m.launches++;
}

void launch_rebuild(Model& m, cudaStream_t st)
{
    SBMF_LAUNCH((rebuild_kernel), m.red_blocks, RED_THREADS, 0, st, m.csr_urow, m.us.idx, m.csr_r, m.us.e, m.us.F, m.it.F, m.us.bias, m.it.bias, m.sc, m.I + 1,
                                                         m.J + 1, m.KB, m.n_csr, m.red_part);
    m.launches++;
}

void launch_stats(Model& m, cudaStream_t st, bool from_csc)
{
    SBMF_LAUNCH((stats_kernel), m.red_blocks, RED_THREADS, 0, st, from_csc ? m.it.e : m.us.e, from_csc ? m.n_csc : m.n_csr, m.red_part);
    m.launches++;
}

void launch_global_hypers(Model& m, cudaStream_t st)
{
    SBMF_LAUNCH((global_hypers_kernel), 1, 32, 0, st, m.sc, m.red2, m.N, m.cfg.priors, m.cfg.sample_mode, m.cfg.hyper_mode, m.cfg.seed);
    m.launches++;
}

static void dim_hypers_side(Model& m, Side& s, cudaStream_t st)
{
    const sbmf_priors& p = m.cfg.priors;
    SBMF_LAUNCH((dim_hyper_partial_kernel), dim3(s.hyp_chunks, m.KB), 256, 0, st, s.F, s.n, s.n + 1, s.mu_k, s.hyp_part, s.hyp_chunks);
    SBMF_LAUNCH((dim_hyper_final_kernel), m.KB, 32, 0, st, s.hyp_part, s.hyp_chunks, s.n, m.K, s.sigma_k, s.mu_k, s.sigma_kf, s.mu_kf, m.sc, p.alpha[s.prior],
                                                p.beta[s.prior], p.mu[s.prior], p.sigma[s.prior], m.cfg.sample_mode, m.cfg.seed, s.site_sigma_k,
                                                s.site_mu_k, m.cfg.hyper_mode, p, s.post_var,
                                                (m.cfg.hyper_mode == SBMF_HYPER_NG_S && &s == &m.it) ? m.us.post_var : nullptr);
    m.launches += 2;
}

void launch_dim_hypers(Model& m, cudaStream_t st)
{
    dim_hypers_side(m, m.us, st);
    dim_hypers_side(m, m.it, st);
}

void launch_bias_hypers(Model& m, cudaStream_t st)
{
    const sbmf_priors& p = m.cfg.priors;
    for (Side* s : {&m.us, &m.it}) {
        SBMF_LAUNCH((bias_hyper_kernel), (s->n + 255) / 256, 256, 0, st, s->bias, s->mu_b, s->sigma_b, s->n, m.sc, p.alpha[s->prior_b], p.beta[s->prior_b],
                                                              p.mu[s->prior_b], p.sigma[s->prior_b], m.cfg.sample_mode, m.cfg.seed,
                                                              s->site_sigma_b, s->site_mu_b, s->id_at);
        m.launches++;
    }
}

template <int BIN>
static void launch_bin(Model& m, const PhaseArgs& a, const Side& self, int b0, int b1, int do_bias, bool refresh, cudaStream_t st)
{
    constexpr int RPL = kBins[BIN].rpl, WARPS = kBins[BIN].warps;
    const uint32_t n = self.bin_count[BIN];
    if (!n) return;
    // option alt_bins: rows of 193..512 ratings by 2 warps x (6 | 8) ratings per lane instead of 4 warps x (3 | 4) -- same capacity,
    // half the per-block fixed work (reduction, solve) per rating
    if constexpr (WARPS == 4 && RPL <= 4) {
        if (m.opt.alt_bins) {
            const dim3 grid(n), block(64);
            if (refresh) SBMF_LAUNCH((row_resident2_kernel<2 * RPL, 2, true, 1, false>), grid, block, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
            else SBMF_LAUNCH((row_resident2_kernel<2 * RPL, 2, false, 1, false>), grid, block, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
            m.launches++;
            return;
        }
    }
    const int rk = (int)m.opt.row_kernels;   // 1: kernels.cu; 2: rows2.cuh with the shared-memory reduction; 3: rows2.cuh with the shuffle tree
    if constexpr (WARPS == 0) {   // short rows: G lanes per row (kBins[].cap = G * RPL)
        if (!m.opt.group_rows) {   // option group_rows = 0: one warp per short row instead (debugging / A-B)
            constexpr int R1 = kBins[BIN].cap / 32;
            const dim3 grid1((n + 3) / 4);
            if (rk == 2 && refresh) SBMF_LAUNCH((row_resident2_kernel<R1, 1, true, 1, true>), grid1, 128, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
            else if (rk == 2) SBMF_LAUNCH((row_resident2_kernel<R1, 1, false, 1, true>), grid1, 128, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
            else if (rk == 3 && refresh) SBMF_LAUNCH((row_resident2_kernel<R1, 1, true, 1, false>), grid1, 128, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
            else if (rk == 3) SBMF_LAUNCH((row_resident2_kernel<R1, 1, false, 1, false>), grid1, 128, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
            else if (refresh) SBMF_LAUNCH((row_resident_kernel<R1, 1, true>), grid1, 128, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
            else SBMF_LAUNCH((row_resident_kernel<R1, 1, false>), grid1, 128, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
            m.launches++;
            return;
        }
        constexpr int G = kBins[BIN].cap / (RPL > 0 ? RPL : 1) >= 16 ? 16 : 8;
        const uint32_t rows_per_cta = 4 * (32 / G);
        const dim3 grid((n + rows_per_cta - 1) / rows_per_cta);
        if (rk == 2 && refresh) SBMF_LAUNCH((row_group2_kernel<RPL, G, true, true>), grid, 128, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
        else if (rk == 2) SBMF_LAUNCH((row_group2_kernel<RPL, G, false, true>), grid, 128, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
        else if (rk == 3 && refresh) SBMF_LAUNCH((row_group2_kernel<RPL, G, true, false>), grid, 128, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
        else if (rk == 3) SBMF_LAUNCH((row_group2_kernel<RPL, G, false, false>), grid, 128, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
        else if (refresh) SBMF_LAUNCH((row_group_kernel<RPL, G, true>), grid, 128, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
        else SBMF_LAUNCH((row_group_kernel<RPL, G, false>), grid, 128, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
        m.launches++;
        return;
    }
    if constexpr (WARPS > 0) {
        constexpr int NR = (WARPS == 8) ? 2 : 1;   // rounds of the shared-memory reduction: the 8-warp exchange buffers must fit 48 KB
        const dim3 grid(WARPS == 1 ? (n + 3) / 4 : n), block(WARPS == 1 ? 128 : WARPS * 32);
        if (rk == 2 && refresh) SBMF_LAUNCH((row_resident2_kernel<RPL, WARPS, true, NR, true>), grid, block, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
        else if (rk == 2) SBMF_LAUNCH((row_resident2_kernel<RPL, WARPS, false, NR, true>), grid, block, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
        else if (rk == 3 && refresh) SBMF_LAUNCH((row_resident2_kernel<RPL, WARPS, true, 1, false>), grid, block, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
        else if (rk == 3) SBMF_LAUNCH((row_resident2_kernel<RPL, WARPS, false, 1, false>), grid, block, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
        else if (refresh) SBMF_LAUNCH((row_resident_kernel<RPL, WARPS, true>), grid, block, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
        else SBMF_LAUNCH((row_resident_kernel<RPL, WARPS, false>), grid, block, 0, st, a, self.bin_rows[BIN], n, b0, b1, do_bias);
        m.launches++;
    }
}

// One half-sweep: bias then all factor blocks of every row of `self` ([T]:514-558 users / 563-606 items).
void launch_phase(Model& m, Side& self, const Side& other, bool apply_shift, bool refresh, const float* e_src, const uint32_t* e_map,
                  bool push_final)
{
    PhaseArgs a;
    a.xmap = push_final ? m.xmap_fwd : nullptr;
    for (int q = 0; q < MAX_PEERS; ++q) a.xdst[q] = (push_final && q < m.world) ? m.precv[q] : nullptr;
    a.e_src = e_src;   // first touch of e (block 0 launches, first streaming pass) reads through the map, see PhaseArgs
    a.e_map = e_map;
    a.r = m.csr_r;              // only the user phase refreshes (refresh == false on the item side)
    a.bias_other = other.bias;
    a.pacc = m.pacc;
    a.KBtot = (int)m.KB;
    a.rows_of_hrow = self.heavy_rows;
    a.row_id = self.id_at;
    a.ptr = self.ptr;
    a.idx = self.idx;
    a.e = self.e;
    a.Fself = self.F;
    {
        const int side = (&self == &m.us) ? 0 : 1;
        a.nrep = m.peer_ok ? m.world : 1;
        for (int q = 0; q < MAX_PEERS; ++q) {
            a.Frep[q] = (m.peer_ok && q < m.world) ? m.pF[side][q] : self.F;
            a.brep[q] = (m.peer_ok && q < m.world) ? m.pbias[side][q] : self.bias;
        }
    }
    a.Fother = other.F;
    a.F2other = other.F2;
    a.ns_self = self.n + 1;
    a.ns_other = other.n + 1;
    a.bias = self.bias;
    a.mu_b = self.mu_b;
    a.sigma_b = self.sigma_b;
    a.sigma_kf = self.sigma_kf;
    a.mu_kf = self.mu_kf;
    a.sc = m.sc;
    a.K = m.K;
    a.seed = m.cfg.seed;
    a.site_f = self.site_f;
    a.site_b = self.site_b;
    a.mode = m.cfg.sample_mode;
    a.apply_shift = apply_shift ? 1 : 0;

    const int KB = (int)m.KB;
    const bool heavy = self.n_heavy > 0;
    const bool with_bias = m.cfg.hyper_mode == SBMF_HYPER_REF_T;   // [S] (the Normal-Gamma modes) has no bias half-step
    // row classes are independent (disjoint rows, slots and factor rows): the streaming pipeline runs on its own stream and the
    // resident bins are spread over three, so the tail of one launch overlaps the head of another
    cudaStream_t sr = m.s_main, sh = m.s_aux;
    cudaStream_t sb[3] = {m.s_main, m.s_res[0], m.s_res[1]};
    cudaEventRecord(m.ev_fork, sr);
    if (heavy) {
        cudaStreamWaitEvent(sh, m.ev_fork, 0);
        if (other.F2 && m.opt.pair_gather && KB > 1) {   // the other side's rows are final for this phase: lay its block pairs side by side
            const uint64_t total = (uint64_t)(KB - 1) * (other.n + 1);
            SBMF_LAUNCH((pair_pack_kernel), grid_for(total, 256, m.sm_count * 16), 256, 0, sh, other.F, other.F2, other.n + 1, (uint32_t)(KB - 1));
            m.launches++;
        }
    }
    cudaStreamWaitEvent(sb[1], m.ev_fork, 0);
    cudaStreamWaitEvent(sb[2], m.ev_fork, 0);
    // resident rows: several blocks per launch (e, idx and the row set-up are then touched once per launch, not once per
    // block), bounded by the bytes of gathered factor blocks one launch has in flight.  Measured: a 48 MB bound (blocks
    // strictly L2-resident) loses to 100-200 MB on both the Netflix-shaped matrix (item phase 10.94 -> 10.90 ms) and the
    // 10M x 1M one (user phase 378 -> 271 ms); beyond ~200 MB it is flat.
    const size_t block_bytes = (size_t)other.n * 32;
    const size_t l2_budget = (size_t)m.opt.l2_budget_mb << 20;   // option l2_budget_mb (tests shrink it to reach the multi-launch continuation)
    int nb = (int)(l2_budget / (block_bytes ? block_bytes : 1));
    if (m.opt.max_blocks_per_launch > 0 && nb > (int)m.opt.max_blocks_per_launch) nb = (int)m.opt.max_blocks_per_launch;
    if (nb < 1) nb = 1;
    if (nb > KB) nb = KB;
    for (int b0 = 0; b0 < KB; b0 += nb) {
        const int b1 = (b0 + nb < KB) ? b0 + nb : KB;
        const int do_bias = (b0 == 0 && with_bias) ? 1 : 0;
        if (b0 > 0) a.e_map = nullptr;   // later launches continue from e[] as stored by the first
        launch_bin<11>(m, a, self, b0, b1, do_bias, refresh, sb[0]);   // longest rows first
        launch_bin<10>(m, a, self, b0, b1, do_bias, refresh, sb[1]);
        launch_bin<9>(m, a, self, b0, b1, do_bias, refresh, sb[2]);
        launch_bin<8>(m, a, self, b0, b1, do_bias, refresh, sb[0]);
        launch_bin<7>(m, a, self, b0, b1, do_bias, refresh, sb[1]);
        launch_bin<6>(m, a, self, b0, b1, do_bias, refresh, sb[2]);
        launch_bin<5>(m, a, self, b0, b1, do_bias, refresh, sb[0]);
        launch_bin<4>(m, a, self, b0, b1, do_bias, refresh, sb[1]);
        launch_bin<3>(m, a, self, b0, b1, do_bias, refresh, sb[2]);
        launch_bin<2>(m, a, self, b0, b1, do_bias, refresh, sb[0]);
        launch_bin<1>(m, a, self, b0, b1, do_bias, refresh, sb[1]);
        launch_bin<0>(m, a, self, b0, b1, do_bias, refresh, sb[2]);
    }
    a.e_map = e_map;   // (only the PREV == 0 streaming pass looks at it)
    cudaEventRecord(m.ev_join_res[0], sb[1]);
    cudaEventRecord(m.ev_join_res[1], sb[2]);
    cudaStreamWaitEvent(sr, m.ev_join_res[0], 0);
    cudaStreamWaitEvent(sr, m.ev_join_res[1], 0);
    if (heavy) {
        const uint32_t ns = self.n_slices, nh = self.n_heavy;
        float* hbias = self.hdelta + (size_t)nh * 8;
    // 2 ratings per thread x 64 threads per slice CTA measured best on B200 (profiles/): small CTAs, ~20 warps per SM.
    // The row updates run as a launch of their own between two passes (heavy_solve_kernel: one CTA per streamed row, latency-bound),
    // or, option fuse_solve, in the tail of each pass by the last slice CTA of the row.
    // Option heavy_chains = 2: the streamed rows are cut into two halves (by slices, at a row boundary) that run as two independent
    // pass -> update -> pass chains on two streams, so the updates of one half execute under the passes of the other instead of
    // leaving the GPU to ~9,000 tiny CTAs 2 KB + 2 times per phase.  Per-pass timing (timing_detail) keeps the single chain: it
    // measures the kernel, not the overlap.
        const bool fuse = m.opt.fuse_solve != 0;
        uint32_t* hc = fuse ? self.hcount : nullptr;
        struct Chain {
            cudaStream_t st;
            uint32_t h0, nh, s0, ns;
        };
        Chain chains[2] = {{sh, 0, nh, 0, ns}, {m.s_aux2, 0, 0, 0, 0}};
        int nch = 1;
        const bool detail = m.timing_detail && !apply_shift;   // item phase only
        if (m.opt.heavy_chains == 2 && !fuse && !m.timing_detail && m.s_aux2 && self.h_split > 0 && self.s_split > 0 && self.s_split < ns) {
            chains[0] = Chain{sh, 0, self.h_split, 0, self.s_split};
            chains[1] = Chain{m.s_aux2, self.h_split, nh - self.h_split, self.s_split, ns - self.s_split};
            nch = 2;
            cudaStreamWaitEvent(m.s_aux2, m.ev_fork, 0);
        }
#define HEAVY_ACC_T(PREV, CUR, RF, PR, FU, PB, B)                                                                                            \
    SBMF_LAUNCH((heavy_accumulate_kernel<PREV, CUR, 2, 64, RF, PR, FU>), C.ns, 64, 0, C.st, a, self.slices + C.s0, self.hdelta, hbias,           \
                self.hpart + (size_t)C.s0 * NACC, self.heavy_slice_ptr, hc, PB, B)
#define HEAVY_ACC_P(PREV, CUR, PR, PB, B)                                     \
    do {                                                                      \
        if (refresh && fuse) HEAVY_ACC_T(PREV, CUR, true, PR, true, PB, B);   \
        else if (refresh) HEAVY_ACC_T(PREV, CUR, true, PR, false, PB, B);     \
        else if (fuse) HEAVY_ACC_T(PREV, CUR, false, PR, true, PB, B);        \
        else HEAVY_ACC_T(PREV, CUR, false, PR, false, PB, B);                 \
        m.launches++;                                                         \
    } while (0)
#define HEAVY_ACC(PREV, CUR, PB, B) HEAVY_ACC_P(PREV, CUR, false, PB, B)
#define HEAVY_SOLVE(CUR, B)                                                                                                                  \
    do {                                                                                                                                     \
        if (fuse) break;                                                                                                                     \
        if (C.ns >= 8u * C.nh)   /* >= 8 slices per streamed row on average */                                                               \
            SBMF_LAUNCH((heavy_solve_kernel<CUR, 8>), C.nh, 256, 0, C.st, a, self.heavy_slice_ptr, self.hpart, self.hdelta, hbias, B, C.h0);    \
        else SBMF_LAUNCH((heavy_solve_kernel<CUR, 2>), C.nh, 64, 0, C.st, a, self.heavy_slice_ptr, self.hpart, self.hdelta, hbias, B, C.h0);    \
        m.launches++;                                                                                                                        \
    } while (0)
#define EACH_CHAIN(STEP)                    \
    for (int c_ = 0; c_ < nch; ++c_) {      \
        const Chain& C = chains[c_];        \
        STEP;                               \
    }
        if (with_bias) {
            EACH_CHAIN(HEAVY_ACC(0, 1, 0, 0));
            EACH_CHAIN(HEAVY_SOLVE(1, 0));
            EACH_CHAIN(HEAVY_ACC(1, 2, 0, 0));
        } else {
            EACH_CHAIN(HEAVY_ACC(0, 2, 0, 0));
        }
        EACH_CHAIN(HEAVY_SOLVE(2, 0));
        if (detail && m.ev_top.size() < (size_t)2 * KB) {
            while (m.ev_top.size() < (size_t)2 * KB) {
                cudaEvent_t ev;
                cudaEventCreate(&ev);
                m.ev_top.push_back(ev);
            }
        }
        if (detail) m.ev_top_used = 0;
        const bool pair = other.F2 != nullptr && m.opt.pair_gather;
        for (int b = 1; b < KB; ++b) {
            if (detail) cudaEventRecord(m.ev_top[m.ev_top_used++], sh);   // (detail: one chain on sh)
            if (!pair) { EACH_CHAIN(HEAVY_ACC(2, 2, b - 1, b)); }
            else { EACH_CHAIN(HEAVY_ACC_P(2, 2, true, b - 1, b)); }
            if (detail) cudaEventRecord(m.ev_top[m.ev_top_used++], sh);
            EACH_CHAIN(HEAVY_SOLVE(2, b));
        }
        EACH_CHAIN(HEAVY_ACC(2, 0, KB - 1, 0));
#undef EACH_CHAIN
#undef HEAVY_ACC
#undef HEAVY_ACC_P
#undef HEAVY_ACC_T
#undef HEAVY_SOLVE
        cudaEventRecord(m.ev_join, sh);
        cudaStreamWaitEvent(sr, m.ev_join, 0);
        if (nch == 2) {
            cudaEventRecord(m.ev_join2, m.s_aux2);
            cudaStreamWaitEvent(sr, m.ev_join2, 0);
        }
    }
}

// Residual between the two slot orders.  One GPU: a gather through perm.  G GPUs: every value moves from the rank that owns
// its user to the rank that owns its item -- pack in the destination's CSC order, one grouped NCCL send/recv over NVLink,
// unpack through recv_pos (plan.cpp).
int launch_allgather_side(Model& m, Side& s, cudaStream_t st);
static int permute_peer(Model& m, bool csr_to_csc, cudaStream_t st);

__global__ void __launch_bounds__(256)
invert_perm_kernel(const uint32_t* __restrict__ perm, uint32_t* __restrict__ inv, uint64_t n)
{
    for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < n; t += (uint64_t)gridDim.x * 256) inv[perm[t]] = (uint32_t)t;
}

// the reverse direction (CSC order -> CSR order) as GATHERS through inverse maps, built on first use: random 4-byte
// reads are ~4x cheaper than random 4-byte writes
static bool ensure_inverse(Model& m, const uint32_t* perm, uint32_t** inv, uint64_t n, cudaStream_t st)
{
    if (*inv) return true;
    // one GPU: from the stream-ordered pool like the rest of the layout (storage.cu: palloc)
    if ((m.world == 1 ? cudaMallocAsync((void**)inv, (n ? n : 1) * 4, st) : cudaMalloc((void**)inv, (n ? n : 1) * 4)) != cudaSuccess) {
        *inv = nullptr;
        cudaGetLastError();
        return false;   // fall back to the scatter form
    }
    SBMF_LAUNCH((invert_perm_kernel), grid_for(n, 256, m.sm_count * 16), 256, 0, st, perm, *inv, n);
    return true;
}

bool ensure_perm_inverse(Model& m, cudaStream_t st) { return ensure_inverse(m, m.perm, &m.perm_inv, m.N, st); }

int launch_permute(Model& m, bool csr_to_csc, Side* gather_side, cudaStream_t st)
{
    if (m.world == 1) {
        const uint32_t g = grid_for(m.N, 256, m.sm_count * 16);
        if (csr_to_csc) SBMF_LAUNCH((permute_gather_kernel), g, 256, 0, st, m.us.e, m.perm, m.it.e, m.N);
        else if (ensure_inverse(m, m.perm, &m.perm_inv, m.N, st)) SBMF_LAUNCH((permute_gather_kernel), g, 256, 0, st, m.it.e, m.perm_inv, m.us.e, m.N);
        else SBMF_LAUNCH((permute_scatter_kernel), g, 256, 0, st, m.it.e, m.perm, m.us.e, m.N);
        m.launches++;
        return 0;
    }
    if (m.peer_ok) return permute_peer(m, csr_to_csc, st);   // updated rows were already written into every replica by the phase kernels
    // G GPUs: one grouped NCCL launch carries the residual all-to-all AND the all-gather of the rows this rank just updated
    // (both needed before the next phase; together they keep all NVLink links busy instead of running back to back)
    const uint32_t gs = grid_for(m.n_csr, 256, m.sm_count * 16), gr = grid_for(m.n_csc, 256, m.sm_count * 16);
    int rc = 0;
    if (csr_to_csc) {
        SBMF_LAUNCH((permute_gather_kernel), gs, 256, 0, st, m.us.e, m.send_idx, m.sendbuf, m.n_csr);
        if ((rc = comm_group_begin(m.err)) != 0) return rc;
        rc = comm_alltoallv_f32(m.comm, m.sendbuf, m.send_off.data(), m.send_cnt.data(), m.recvbuf, m.recv_off.data(), m.recv_cnt.data(), st, m.err);
        if (!rc && gather_side) rc = launch_allgather_side(m, *gather_side, st);
        if (rc) {   // a failed call: close the outer group (keeping the first error message) and stop issuing NCCL work
            std::string ignored;
            comm_group_end(ignored);
            return rc;
        }
        if ((rc = comm_group_end(m.err)) != 0) return rc;
        SBMF_LAUNCH((permute_gather_kernel), gr, 256, 0, st, m.recvbuf, m.recv_pos, m.it.e, m.n_csc);
    } else {
        if (ensure_inverse(m, m.recv_pos, &m.recv_pos_inv, m.n_csc, st)) SBMF_LAUNCH((permute_gather_kernel), gr, 256, 0, st, m.it.e, m.recv_pos_inv, m.recvbuf, m.n_csc);
        else SBMF_LAUNCH((permute_scatter_kernel), gr, 256, 0, st, m.it.e, m.recv_pos, m.recvbuf, m.n_csc);
        if ((rc = comm_group_begin(m.err)) != 0) return rc;
        rc = comm_alltoallv_f32(m.comm, m.recvbuf, m.recv_off.data(), m.recv_cnt.data(), m.sendbuf, m.send_off.data(), m.send_cnt.data(), st, m.err);
        if (!rc && gather_side) rc = launch_allgather_side(m, *gather_side, st);
        if (rc) {   // a failed call: close the outer group (keeping the first error message) and stop issuing NCCL work
            std::string ignored;
            comm_group_end(ignored);
            return rc;
        }
        if ((rc = comm_group_end(m.err)) != 0) return rc;
        if (ensure_inverse(m, m.send_idx, &m.send_idx_inv, m.n_csr, st)) SBMF_LAUNCH((permute_gather_kernel), gs, 256, 0, st, m.sendbuf, m.send_idx_inv, m.us.e, m.n_csr);
        else SBMF_LAUNCH((permute_scatter_kernel), gs, 256, 0, st, m.sendbuf, m.send_idx, m.us.e, m.n_csr);
    }
    m.launches += 2;
    return rc;
}

// Residual exchange as direct NVLink stores: position i of my send order belongs to destination rank q = segment of i, and goes
// to dst[q][dst_off[q] + (i - seg_off[q])] in q's receive buffer.  src_idx[i] = local slot to read.
struct PushArgs {
    float* dst[MAX_PEERS];
    uint64_t seg_off[MAX_PEERS + 1];
    uint64_t dst_off[MAX_PEERS];
    int world;
};
__global__ void __launch_bounds__(256)
push_kernel(const float* __restrict__ e, const uint32_t* __restrict__ src_idx, PushArgs pa, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        int q = 0;
#pragma unroll
        for (int k = 1; k < MAX_PEERS; ++k) q += (k < pa.world && i >= pa.seg_off[k]) ? 1 : 0;
        pa.dst[q][pa.dst_off[q] + (i - pa.seg_off[q])] = e[src_idx[i]];
    }
    __threadfence_system();
}

int launch_barrier(Model& m, cudaStream_t st)
{
    if (m.world == 1) return 0;
    return comm_allreduce_sum_f64(m.comm, m.bar, 1, st, m.err);
}

// forward (CSR -> CSC) / reverse exchange with peer pushes: pack+send in one kernel, barrier, unpack
static int permute_peer(Model& m, bool csr_to_csc, cudaStream_t st)
{
    PushArgs pa;
    pa.world = m.world;
    const uint32_t gs = grid_for(m.n_csr, 256, m.sm_count * 16), gr = grid_for(m.n_csc, 256, m.sm_count * 16);
    if (csr_to_csc) {
        for (int q = 0; q < m.world; ++q) {
            pa.dst[q] = m.precv[q];
            pa.seg_off[q] = m.send_off[q];
            pa.dst_off[q] = m.fwd_dst_off[q];
        }
        pa.seg_off[m.world] = m.n_csr;
        SBMF_LAUNCH((push_kernel), gs, 256, 0, st, m.us.e, m.send_idx, pa, m.n_csr);
        int rc = launch_barrier(m, st);   // also: every peer has finished its user phase, so all U rows / biases have landed here
        SBMF_LAUNCH((permute_gather_kernel), gr, 256, 0, st, m.recvbuf, m.recv_pos, m.it.e, m.n_csc);
        m.launches += 2;
        return rc;
    }
    if (!ensure_inverse(m, m.recv_pos, &m.recv_pos_inv, m.n_csc, st) || !ensure_inverse(m, m.send_idx, &m.send_idx_inv, m.n_csr, st)) {
        m.err = "out of device memory for the inverse exchange maps";
        return -1;
    }
    for (int q = 0; q < m.world; ++q) {
        pa.dst[q] = m.psend[q];
        pa.seg_off[q] = m.recv_off[q];
        pa.dst_off[q] = m.rev_dst_off[q];
    }
    pa.seg_off[m.world] = m.n_csc;
    // position i of my receive order holds local CSC slot recv_pos_inv[i]; it returns to the rank it came from
    SBMF_LAUNCH((push_kernel), gr, 256, 0, st, m.it.e, m.recv_pos_inv, pa, m.n_csc);
    int rc = launch_barrier(m, st);       // also: all V rows / item biases have landed
    SBMF_LAUNCH((permute_gather_kernel), gs, 256, 0, st, m.sendbuf, m.send_idx_inv, m.us.e, m.n_csr);
    m.launches += 2;
    return rc;
}

int launch_reduce_pair(Model& m, cudaStream_t st)
{
    SBMF_LAUNCH((reduce_pair_kernel), 1, RED_THREADS, 0, st, m.red_part, m.red_blocks, m.red2);
    m.launches++;
    if (m.world > 1) return comm_allreduce_sum_f64(m.comm, m.red2, 2, st, m.err);
    return 0;
}

int launch_allgather_side(Model& m, Side& s, cudaStream_t st)
{
    if (m.world == 1) return 0;
    if (m.peer_ok) return launch_barrier(m, st);   // rows were pushed by the phase kernels; only make sure every peer is done
    const std::vector<uint32_t>& bd = (&s == &m.us) ? m.ub : m.ib;
    std::vector<size_t> off(m.world), cnt(m.world), off8(m.world), cnt8(m.world);
    for (int q = 0; q < m.world; ++q) {
        off[q] = bd[q];
        cnt[q] = bd[q + 1] - bd[q];
        off8[q] = off[q] * 8;
        cnt8[q] = cnt[q] * 8;
    }
    int rc = comm_allgatherv_strided_f32(m.comm, s.F, (size_t)(s.n + 1) * 8, (int)m.KB, off8.data(), cnt8.data(), st, m.err);
    if (rc) return rc;
    return comm_allgatherv_f32(m.comm, s.bias, off.data(), cnt.data(), st, m.err);
}

void launch_eval(Model& m, cudaStream_t st)
{
    const uint64_t t0 = m.t_begin, nt = m.t_end - m.t_begin;   // this rank's slice of the test set
    SBMF_LAUNCH((eval_kernel), m.red_blocks, RED_THREADS, 0, st, m.t_user + t0, m.t_item + t0, m.t_r + t0, m.t_sum + t0, m.us.F, m.it.F, m.us.bias, m.it.bias,
                                                      m.sc, m.I + 1, m.J + 1, m.KB, nt, m.cfg.burn_in, (float)m.cfg.clamp_lo, (float)m.cfg.clamp_hi,
                                                      m.red_part);
    m.launches++;
}

__global__ void __launch_bounds__(256)
pred_mean_kernel(const double* __restrict__ tsum, float* __restrict__ out, uint64_t n, double denom)
{
    for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < n; t += (uint64_t)gridDim.x * 256) out[t] = (float)(tsum[t] / denom);
}

void launch_pred_mean(Model& m, float* d_out, double denom, cudaStream_t st)
{
    SBMF_LAUNCH((pred_mean_kernel), grid_for(m.Nt, 256, m.sm_count * 8), 256, 0, st, m.t_sum, d_out, m.Nt, denom);
    m.launches++;
}

void launch_eval_final(Model& m, cudaStream_t st)
{
    SBMF_LAUNCH((eval_final_kernel), 1, 32, 0, st, m.sc, m.red2, m.Nt, m.rmse_hist, m.hist_cap);
    m.launches++;
}

}  // namespace sbmf
