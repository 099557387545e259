// storage.cu -- rating storage on the device: COO in file order -> CSR (by user) + CSC (by item) + the
// CSC-slot -> CSR-slot permutation + residual arrays + row work lists.
//
// Replaces the jagged R / R_t build of gibbs_sbpmf2.cpp ("[T]"):156-221.  [T] appends every rating, in FILE
// order, to its user's row and to its item's row ([T]:209-214), so both layouts are stable w.r.t. file order;
// here that is two stable LSD radix sorts (cub::DeviceRadixSort) of the rating index by user and by item.
// [T]'s `.id` back-pointer (rating index of each slot) becomes csr_id / csc_id, and the pair of them the
// permutation perm[csc slot] = csr slot that moves the residual between the two orders.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cub/cub.cuh>
#include <vector>

#include "launch.h"
#include "model.h"

namespace sbmf {

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
    cudaError_t e = cudaMalloc((void**)p, (n ? n : 1) * sizeof(T));
    if (e == cudaErrorMemoryAllocation) {   // memory parked in the stream-ordered pool (see palloc) is given back before giving up
        cudaGetLastError();
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            cudaDeviceSynchronize();
            cudaMemPoolTrimTo(pool, 0);
            e = cudaMalloc((void**)p, (n ? n : 1) * sizeof(T));
        }
    }
    return e;
}

// The rating-sized arrays of a single-GPU model come from the device's stream-ordered pool (release threshold raised in
// sbmf_cuda_create): a second set_train in the same process then reuses the memory of the first instead of paying
// cudaMalloc / cudaFree of several GB again (measured 25-110 ms and varying for the layout alone).  Multi-GPU models keep
// cudaMalloc throughout: their buffers are exported through CUDA IPC, which pool memory does not support.
// Option mgpu_pool (default on): the pool also serves the multi-GPU layout -- only the six buffers setup_peer_access exports
// (factors, biases, exchange buffers; plain dmalloc below) have to be cudaMalloc memory.
template <class T>
static cudaError_t palloc(const Model& m, T** p, size_t n, cudaStream_t st)
{
    if (m.world > 1 && !m.opt.mgpu_pool) return dmalloc(p, n);
    return cudaMallocAsync((void**)p, (n ? n : 1) * sizeof(T), st);
}

__global__ void iota_kernel(uint32_t* v, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) v[i] = (uint32_t)i;
}

__global__ void max_id_kernel(const uint32_t* a, const uint32_t* b, uint64_t n, uint32_t* out)
{
    uint32_t ma = 0, mb = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        ma = max(ma, a[i]);
        mb = max(mb, b[i]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        ma = max(ma, __shfl_xor_sync(0xffffffffu, ma, o));
        mb = max(mb, __shfl_xor_sync(0xffffffffu, mb, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, ma);
        atomicMax(out + 1, mb);
    }
}

// ptr[r] = first slot whose (sorted) key is >= r
__global__ void row_ptr_kernel(const uint32_t* sorted_keys, uint64_t n, uint32_t nrows, int64_t* ptr)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nrows) return;
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < r) lo = mid + 1;
        else hi = mid;
    }
    ptr[r] = (int64_t)lo;
}

__global__ void gather_u32_kernel(const uint32_t* src, const uint32_t* id, uint32_t* dst, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[id[i]];
}
__global__ void gather_f32_kernel(const float* src, const uint32_t* id, float* dst, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[id[i]];
}
__global__ void invert_kernel(const uint32_t* id, uint32_t* inv, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) inv[id[i]] = (uint32_t)i;
}

// ---- row relabelling (option relabel) -----------------------------------------------------------------------------------------
// The gathers of the row kernels fetch one 32-byte factor sector per rating, addressed by the opposite side's row index; adjacent
// lanes hold consecutive ratings of a row.  With the caller's (arbitrary) ids two lanes almost never fall into the same 128-byte
// line.  Ordered by decreasing rating count, the popular rows are neighbours, every row lists them first, and the lanes that
// touch them share lines -- fewer L1TEX wavefronts for the same sectors (measured on the Netflix-shaped matrix: 24.5 -> 22.8 ms
// per sweep).  Everything inside the model is indexed by position; the boundary (api.cu) translates, draws stay keyed by id.
__global__ void degree_kernel(const uint32_t* __restrict__ ids, uint64_t n, uint32_t* __restrict__ deg)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) atomicAdd(&deg[ids[i]], 1u);
}
__global__ void rank_key_kernel(const uint32_t* __restrict__ deg, uint32_t n, uint32_t* __restrict__ key, uint32_t* __restrict__ id)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    key[r] = 0xffffffffu - deg[r];   // ascending key = descending degree; the sort is stable, ties stay in id order
    id[r] = r;
}
// rank r (0 = most ratings) -> position: dealt round-robin into `chunks` contiguous ranges (one per GPU), so every shard of a
// multi-GPU model sees the same mix of row lengths; one chunk = plain rank order
__global__ void assign_pos_kernel(const uint32_t* __restrict__ id_by_rank, uint32_t n, uint32_t chunks, uint32_t* __restrict__ id_at,
                                  uint32_t* __restrict__ pos_of)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t c = r % chunks, k = r / chunks;
    // chunk c holds ceil((n - c) / chunks) rows; chunks before it: c * (n / chunks) + min(c, n % chunks)
    const uint32_t off = c * (n / chunks) + min(c, n % chunks);
    const uint32_t pos = off + k;
    const uint32_t id = id_by_rank[r];
    id_at[pos] = id;
    pos_of[id] = pos;
}
__global__ void relabel_kernel(uint32_t* __restrict__ ids, uint64_t n, const uint32_t* __restrict__ pos_of)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) ids[i] = pos_of[ids[i]];
}
// the caller's COO back from the relabelled CSR arrays (reference-layout export): rating csr_id[s] was (id_u[urow[s]], id_v[idx[s]])
__global__ void restore_coo_kernel(const uint32_t* __restrict__ csr_id, const uint32_t* __restrict__ urow, const uint32_t* __restrict__ idx,
                                   const uint32_t* __restrict__ id_u, const uint32_t* __restrict__ id_v, uint64_t n, uint32_t* __restrict__ user,
                                   uint32_t* __restrict__ item)
{
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = csr_id[s];
        user[r] = id_u[urow[s]];
        item[r] = id_v[idx[s]];
    }
}

// fused forward exchange: position i of my send order (segment q = destination rank) is local CSR slot send_idx[i]; its value
// belongs at recvbuf(q)[dst_off[q] + i - seg_off[q]]
struct XSeg {
    uint64_t seg_off[MAX_PEERS + 1];
    uint64_t dst_off[MAX_PEERS];
    int world;
};
__global__ void __launch_bounds__(256) xmap_kernel(const uint32_t* __restrict__ send_idx, uint64_t n, XSeg sg, uint32_t* __restrict__ xmap)
{
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        int q = 0;
#pragma unroll
        for (int k = 1; k < MAX_PEERS; ++k) q += (k < sg.world && i >= sg.seg_off[k]) ? 1 : 0;
        xmap[send_idx[i]] = ((uint32_t)q << 28) | (uint32_t)(sg.dst_off[q] + (i - sg.seg_off[q]));
    }
}

static int bits_for(uint32_t n)
{
    int b = 1;
    while (b < 32 && (1ull << b) < (uint64_t)n) ++b;
    return b;
}

static void free_side(Side& s)
{
    cudaFree(s.ptr); cudaFree(s.idx); cudaFree(s.e); cudaFree(s.F); cudaFree(s.F2); cudaFree(s.id_at); cudaFree(s.pos_of);
    cudaFree(s.bias); cudaFree(s.mu_b); cudaFree(s.sigma_b);
    cudaFree(s.sigma_k); cudaFree(s.mu_k); cudaFree(s.post_var); cudaFree(s.sigma_kf); cudaFree(s.mu_kf); cudaFree(s.hyp_part);
    for (int b = 0; b < NBINS; ++b) cudaFree(s.bin_rows[b]);
    cudaFree(s.heavy_rows); cudaFree(s.heavy_slice_ptr); cudaFree(s.slices); cudaFree(s.hpart); cudaFree(s.hdelta); cudaFree(s.hcount);
    const uint32_t sf = s.site_f, sb = s.site_b, a = s.site_sigma_k, b_ = s.site_mu_k, c = s.site_sigma_b, d = s.site_mu_b;
    const int pr = s.prior, prb = s.prior_b;
    s = Side();
    s.site_f = sf; s.site_b = sb; s.site_sigma_k = a; s.site_mu_k = b_; s.site_sigma_b = c; s.site_mu_b = d;
    s.prior = pr; s.prior_b = prb;
}

// Map every peer's factor / bias replicas and exchange buffers into this process (CUDA IPC; NVLink peer access is enabled
// lazily by cudaIpcOpenMemHandle).  The 64-byte handles travel through one NCCL all-gather, so no extra host channel is needed.
// Any failure leaves peer_ok = false: the NCCL exchanges are used instead.
int setup_peer_access(Model& m)
{
    m.peer_ok = false;
    const int G = m.world;
    if (G > MAX_PEERS || !m.opt.peer) return SBMF_OK;
    void* mine[6] = {m.us.F, m.it.F, m.us.bias, m.it.bias, m.recvbuf, m.sendbuf};
    constexpr int HW = 6 * 64 / 4;   // floats per rank
    std::vector<float> host((size_t)G * HW, 0.f);
    float ok_flag = 1.f;
    for (int i = 0; i < 6; ++i) {
        cudaIpcMemHandle_t h;
        if (cudaIpcGetMemHandle(&h, mine[i]) != cudaSuccess) {
            cudaGetLastError();
            ok_flag = 0.f;
            memset(&h, 0, sizeof(h));
        }
        memcpy(&host[(size_t)m.rank * HW + i * 16], &h, 64);
    }
    float* d = nullptr;
    CK(cudaMalloc((void**)&d, ((size_t)G * HW + G) * 4));
    CK(cudaMemset(d, 0, ((size_t)G * HW + G) * 4));
    CK(cudaMemcpy(d + (size_t)m.rank * HW, &host[(size_t)m.rank * HW], HW * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d + (size_t)G * HW + m.rank, &ok_flag, 4, cudaMemcpyHostToDevice));
    std::vector<size_t> off(G), cnt(G, HW), off1(G), cnt1(G, 1);
    for (int q = 0; q < G; ++q) {
        off[q] = (size_t)q * HW;
        off1[q] = (size_t)G * HW + q;
    }
    if (comm_allgatherv_f32(m.comm, d, off.data(), cnt.data(), m.s_main, m.err) != 0 ||
        comm_allgatherv_f32(m.comm, d, off1.data(), cnt1.data(), m.s_main, m.err) != 0) {
        cudaFree(d);
        return SBMF_ERR_NCCL;
    }
    std::vector<float> flags(G);
    CK(cudaMemcpyAsync(host.data(), d, (size_t)G * HW * 4, cudaMemcpyDeviceToHost, m.s_main));
    CK(cudaMemcpyAsync(flags.data(), d + (size_t)G * HW, (size_t)G * 4, cudaMemcpyDeviceToHost, m.s_main));
    CK(cudaStreamSynchronize(m.s_main));
    cudaFree(d);
    bool all_ok = true;
    for (int q = 0; q < G; ++q) all_ok = all_ok && flags[q] == 1.f;
    // every rank must reach the same verdict: open, then agree through a second all-reduce-like exchange (sum of failures)
    double fails = 0.0;
    void* opened[MAX_PEERS][6] = {};
    if (all_ok) {
        for (int q = 0; q < G && fails == 0.0; ++q) {
            for (int i = 0; i < 6; ++i) {
                if (q == m.rank) {
                    opened[q][i] = mine[i];
                    continue;
                }
                cudaIpcMemHandle_t h;
                memcpy(&h, &host[(size_t)q * HW + i * 16], 64);
                void* p = nullptr;
                if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    fails = 1.0;
                    break;
                }
                opened[q][i] = p;
                m.ipc_opened.push_back(p);
            }
        }
    } else {
        fails = 1.0;
    }
    CK(cudaMemcpy(m.bar, &fails, 8, cudaMemcpyHostToDevice));
    if (comm_allreduce_sum_f64(m.comm, m.bar, 1, m.s_main, m.err) != 0) return SBMF_ERR_NCCL;
    CK(cudaMemcpyAsync(&fails, m.bar, 8, cudaMemcpyDeviceToHost, m.s_main));
    CK(cudaStreamSynchronize(m.s_main));
    CK(cudaMemset(m.bar, 0, 8));
    if (fails != 0.0) {
        close_peer_access(m);
        return SBMF_OK;   // NCCL fallback
    }
    for (int q = 0; q < G; ++q) {
        m.pF[0][q] = (float*)opened[q][0];
        m.pF[1][q] = (float*)opened[q][1];
        m.pbias[0][q] = (float*)opened[q][2];
        m.pbias[1][q] = (float*)opened[q][3];
        m.precv[q] = (float*)opened[q][4];
        m.psend[q] = (float*)opened[q][5];
    }
    m.peer_ok = true;
    return SBMF_OK;
}

void close_peer_access(Model& m)
{
    for (void* p : m.ipc_opened) cudaIpcCloseMemHandle(p);
    m.ipc_opened.clear();
    m.peer_ok = false;
}

void free_storage(Model& m)
{
    if (m.graph_exec) {   // the captured sweep points at the arrays freed below
        cudaGraphExecDestroy(m.graph_exec);
        m.graph_exec = nullptr;
    }
    m.graph_failed = false;
    cudaDeviceSynchronize();   // pool allocations are released with cudaFree below, which does not wait for work that still uses them
    if (m.world > 1 && m.comm.nccl && m.have_train) {
        // peers may still be writing into / reading from our buffers: meet them before anything is unmapped or freed
        cudaStreamSynchronize(m.s_main);
        cudaStreamSynchronize(m.s_aux);
        if (m.s_aux2) cudaStreamSynchronize(m.s_aux2);
        if (m.bar) {
            std::string err;
            comm_allreduce_sum_f64(m.comm, m.bar, 1, m.s_main, err);
            cudaStreamSynchronize(m.s_main);
        }
    }
    close_peer_access(m);
    cudaFree(m.bar);
    m.bar = nullptr;
    free_side(m.us);
    free_side(m.it);
    cudaFree(m.csr_urow); cudaFree(m.csr_r); cudaFree(m.csr_id); cudaFree(m.csc_id); cudaFree(m.perm); cudaFree(m.pacc);
    m.pacc = nullptr;
    m.csr_urow = nullptr; m.csr_r = nullptr; m.csr_id = nullptr; m.csc_id = nullptr; m.perm = nullptr;
    cudaFree(m.red_part); m.red_part = nullptr;
    cudaFree(m.red2); m.red2 = nullptr;
    cudaFree(m.send_idx); cudaFree(m.recv_pos); cudaFree(m.sendbuf); cudaFree(m.recvbuf);
    cudaFree(m.perm_inv); cudaFree(m.recv_pos_inv); cudaFree(m.send_idx_inv); cudaFree(m.xmap_fwd);
    m.perm_inv = m.recv_pos_inv = m.send_idx_inv = m.xmap_fwd = nullptr;
    m.send_idx = m.recv_pos = nullptr; m.sendbuf = m.recvbuf = nullptr;
    m.n_csr = m.n_csc = 0;
    m.have_train = false;
    m.have_factors = false;
    m.N = 0;
}

void free_test(Model& m)
{
    if (m.graph_exec) {
        cudaGraphExecDestroy(m.graph_exec);
        m.graph_exec = nullptr;
    }
    cudaDeviceSynchronize();   // see free_storage
    cudaFree(m.t_user); cudaFree(m.t_item); cudaFree(m.t_r); cudaFree(m.t_sum);
    m.t_user = m.t_item = nullptr; m.t_r = nullptr; m.t_sum = nullptr;
    m.have_test = false;
    m.Nt = 0;
}

// Row work lists: resident bins by row length, heavy rows cut into slices.  Built on the host from the row
// pointer (one-time; the order of rows inside a list does not influence any result).
static int build_worklists(Model& m, Side& s, uint32_t row0, uint32_t row1)
{
    std::vector<int64_t> ptr((size_t)s.n + 1);
    CK(cudaMemcpy(ptr.data(), s.ptr, ptr.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> bins[NBINS], heavy, hsp;
    std::vector<Slice> slices;
    s.nnz_resident = s.nnz_heavy = 0;
    hsp.push_back(0);
    // slice length: at most SLICE_LEN, but short enough that one streaming launch is >= ~8 waves of CTAs on this GPU
    // (a multi-GPU shard or a thin heavy tail would otherwise run 2-3 ragged waves per launch)
    const int64_t resident_max = (&s == &m.us) ? m.opt.resident_max_user : m.opt.resident_max_item;   // options (<= RESIDENT_MAX)
    uint64_t nnz_heavy_total = 0;
    for (uint32_t r = row0; r < row1; ++r) {
        const int64_t c = ptr[r + 1] - ptr[r];
        if (c > resident_max) nnz_heavy_total += (uint64_t)c;
    }
    int64_t slice_len = SLICE_LEN;
    {
        const int64_t want = (int64_t)(nnz_heavy_total / ((uint64_t)m.sm_count * 80u));
        slice_len = std::min<int64_t>(SLICE_LEN, std::max<int64_t>(1024, (want + 255) / 256 * 256));
        if (m.opt.slice_len > 0) slice_len = m.opt.slice_len;   // option
    }
    for (uint32_t r = row0; r < row1; ++r) {
        const int64_t c = ptr[r + 1] - ptr[r];
        if (c <= resident_max) {
            int b = 0;
            while (c > kBins[b].cap) ++b;
            bins[b].push_back(r);
            s.nnz_resident += (uint64_t)c;
        } else {
            const uint32_t h = (uint32_t)heavy.size();
            heavy.push_back(r);
            // equal-length slices (the last one is not a short tail)
            const int64_t ns = (c + slice_len - 1) / slice_len;
            const int64_t len = (c + ns - 1) / ns;
            for (int64_t o = 0; o < c; o += len) slices.push_back(Slice{ptr[r] + o, (uint32_t)std::min<int64_t>(len, c - o), h});
            hsp.push_back((uint32_t)slices.size());
            s.nnz_heavy += (uint64_t)c;
        }
    }
    for (int b = 0; b < NBINS; ++b) {
        s.bin_count[b] = (uint32_t)bins[b].size();
        CK(dmalloc(&s.bin_rows[b], bins[b].size()));
        if (!bins[b].empty()) CK(cudaMemcpy(s.bin_rows[b], bins[b].data(), bins[b].size() * 4, cudaMemcpyHostToDevice));
    }
    s.n_heavy = (uint32_t)heavy.size();
    s.n_slices = (uint32_t)slices.size();
    // the row boundary closest to half of the slices: the two chains of the streaming pipeline (option heavy_chains)
    s.h_split = s.s_split = 0;
    for (uint32_t h = 1; h < s.n_heavy; ++h)
        if (hsp[h] * 2 >= s.n_slices) {
            s.h_split = h;
            s.s_split = hsp[h];
            break;
        }
    CK(dmalloc(&s.heavy_rows, heavy.size()));
    CK(dmalloc(&s.heavy_slice_ptr, hsp.size()));
    CK(dmalloc(&s.slices, slices.size()));
    CK(dmalloc(&s.hpart, slices.size() * NACC));
    CK(dmalloc(&s.hdelta, heavy.size() * 9));   // [n_heavy][8] block deltas + [n_heavy] bias deltas
    CK(dmalloc(&s.hcount, heavy.size()));
    CK(cudaMemset(s.hcount, 0, (heavy.size() ? heavy.size() : 1) * 4));
    if (!heavy.empty()) {
        CK(cudaMemcpy(s.heavy_rows, heavy.data(), heavy.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(s.slices, slices.data(), slices.size() * sizeof(Slice), cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(s.heavy_slice_ptr, hsp.data(), hsp.size() * 4, cudaMemcpyHostToDevice));
    return SBMF_OK;
}

__global__ void rebase_kernel(int64_t* ptr, uint32_t n, int64_t base)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= n) ptr[r] -= base;
}

template <typename T>
static cudaError_t slice_inplace(const Model& m, T*& arr, uint64_t off, uint64_t cnt, cudaStream_t st)
{
    T* loc = nullptr;
    if (m.opt.mgpu_pool) {   // everything stream-ordered: no cudaMalloc / cudaFree of rating-sized arrays
        cudaError_t e = cudaMallocAsync((void**)&loc, (cnt ? cnt : 1) * sizeof(T), st);
        if (e != cudaSuccess) return e;
        if (cnt) e = cudaMemcpyAsync(loc, arr + off, cnt * sizeof(T), cudaMemcpyDeviceToDevice, st);
        cudaFreeAsync(arr, st);
        arr = loc;
        return e;
    }
    cudaError_t e = dmalloc(&loc, cnt);
    if (e != cudaSuccess) return e;
    if (cnt) e = cudaMemcpy(loc, arr + off, cnt * sizeof(T), cudaMemcpyDeviceToDevice);
    cudaFree(arr);
    arr = loc;
    return e;
}

// ---- exchange planning on the device: the same plan as plan.cpp (sbmf_cuda_plan_exchange + sbmf_cuda_plan_pair_counts)
// without moving the 4-byte-per-rating permutation to the host and walking it there three times.
//   pair counts : histogram of (owner of perm[t], owner of t) over all CSC slots t
//   send_idx    : stream compaction (stable, ascending t) of the perm values that fall into this rank's CSR shard -- ascending
//                 t is "grouped by destination rank, in the destination's CSC order", exactly plan.cpp's order
//   recv_pos    : stable sort of this rank's CSC slots by source rank gives the receive order (position -> slot); its inverse
//                 is recv_pos
constexpr int MAX_WORLD = 64;
struct ShardBounds {
    int64_t csr[MAX_WORLD + 1], csc[MAX_WORLD + 1];
    int world;
};

// last q with b[q] <= slot (plan.cpp: owner); empty shards are never chosen
__device__ __forceinline__ int owner_of(const int64_t* b, int world, int64_t slot)
{
    int lo = 0, hi = world - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (b[mid] <= slot) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
pair_count_kernel(const uint32_t* __restrict__ perm, uint64_t n, ShardBounds b, unsigned long long* __restrict__ counts)
{
#ifdef SBMF_SIMT_EMU
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(simt::block().dyn_smem.data());
#else
    extern __shared__ uint32_t s_hist[];   // [world * world]
#endif
    const int W = b.world;
    for (int i = threadIdx.x; i < W * W; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        const int src = owner_of(b.csr, W, (int64_t)perm[t]);
        const int dst = owner_of(b.csc, W, (int64_t)t);
        atomicAdd(&s_hist[src * W + dst], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < W * W; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&counts[i], (unsigned long long)s_hist[i]);
}

__global__ void source_rank_kernel(const uint32_t* __restrict__ perm_local, uint64_t n, ShardBounds b, uint32_t* __restrict__ key)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        key[i] = (uint32_t)owner_of(b.csr, b.world, (int64_t)perm_local[i]);
}

__global__ void subtract_kernel(uint32_t* v, uint64_t n, uint32_t base)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) v[i] -= base;
}

struct InCsrShard {
    uint32_t lo, hi;
    __host__ __device__ bool operator()(const uint32_t& s) const { return s >= lo && s < hi; }
};

// d_perm: [N] on the device.  d_send_idx [n_csr of rank], d_recv_pos [n_csc of rank]: device outputs.  h_pc: [G * G] host output,
// h_pc[src * G + dst].  Synchronises the stream.
static int plan_exchange_device(Model& m, cudaStream_t st, uint64_t N, const uint32_t* d_perm, int G, int r, const int64_t* cb,
                                const int64_t* tb, uint32_t* d_send_idx, uint32_t* d_recv_pos, int64_t* h_pc)
{
    if (G < 1 || G > MAX_WORLD || r < 0 || r >= G || (uint64_t)cb[G] != N || (uint64_t)tb[G] != N || N >= (1ull << 31)) {
        m.err = "exchange planning: bad bounds";
        return SBMF_ERR_INVALID;
    }
    ShardBounds b;
    memset(&b, 0, sizeof(b));
    b.world = G;
    for (int q = 0; q <= G; ++q) {
        b.csr[q] = cb[q];
        b.csc[q] = tb[q];
    }
    const uint64_t c0 = (uint64_t)cb[r], c1 = (uint64_t)cb[r + 1], t0 = (uint64_t)tb[r], t1 = (uint64_t)tb[r + 1];
    const uint64_t n_csr = c1 - c0, n_csc = t1 - t0;
    const int T = 256;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((N + T - 1) / T + 1, (uint64_t)m.sm_count * 16);
    const uint32_t grid_l = (uint32_t)std::min<uint64_t>((n_csc + T - 1) / T + 1, (uint64_t)m.sm_count * 16);
    unsigned long long* d_pc = nullptr;
    int* d_nsel = nullptr;
    uint32_t *d_key = nullptr, *d_key_out = nullptr, *d_iota = nullptr, *d_order = nullptr;
    void* d_tmp = nullptr;
    auto cleanup = [&]() {
        for (void* p : {(void*)d_pc, (void*)d_nsel, (void*)d_key, (void*)d_key_out, (void*)d_iota, (void*)d_order, d_tmp})
            if (p) cudaFreeAsync(p, st);
    };
#define CKP(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            m.err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            cleanup();                                                                             \
            return (e_ == cudaErrorMemoryAllocation) ? SBMF_ERR_NOMEM : SBMF_ERR_CUDA;             \
        }                                                                                          \
    } while (0)
    auto talloc = [&](void** p, size_t bytes) { return cudaMallocAsync(p, bytes ? bytes : 1, st); };
    CKP(talloc((void**)&d_pc, (size_t)G * G * 8));
    CKP(talloc((void**)&d_nsel, 4));
    CKP(talloc((void**)&d_key, n_csc * 4)); CKP(talloc((void**)&d_key_out, n_csc * 4));
    CKP(talloc((void**)&d_iota, n_csc * 4)); CKP(talloc((void**)&d_order, n_csc * 4));
    size_t sel_bytes = 0, sort_bytes = 0;
    const InCsrShard in_shard{(uint32_t)c0, (uint32_t)c1};
    CKP(cub::DeviceSelect::If(nullptr, sel_bytes, d_perm, d_send_idx, d_nsel, (int)N, in_shard, st));
    CKP(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, d_key, d_key_out, d_iota, d_order, (int)n_csc, 0, bits_for((uint32_t)G), st));
    CKP(talloc(&d_tmp, std::max(sel_bytes, sort_bytes)));
    // traffic matrix
    CKP(cudaMemsetAsync(d_pc, 0, (size_t)G * G * 8, st));
    if (N) SBMF_LAUNCH((pair_count_kernel), grid, T, (size_t)G * G * 4, st, d_perm, N, b, d_pc);
    // send order
    CKP(cub::DeviceSelect::If(d_tmp, sel_bytes, d_perm, d_send_idx, d_nsel, (int)N, in_shard, st));
    if (n_csr) SBMF_LAUNCH((subtract_kernel), grid, T, 0, st, d_send_idx, n_csr, (uint32_t)c0);
    // receive order
    if (n_csc) {
        SBMF_LAUNCH((source_rank_kernel), grid_l, T, 0, st, d_perm + t0, n_csc, b, d_key);
        SBMF_LAUNCH((iota_kernel), grid_l, T, 0, st, d_iota, n_csc);
        CKP(cub::DeviceRadixSort::SortPairs(d_tmp, sort_bytes, d_key, d_key_out, d_iota, d_order, (int)n_csc, 0, bits_for((uint32_t)G), st));
        SBMF_LAUNCH((invert_kernel), grid_l, T, 0, st, d_order, d_recv_pos, n_csc);
    }
    CKP(cudaGetLastError());
    std::vector<unsigned long long> pc((size_t)G * G);
    int nsel = 0;
    CKP(cudaMemcpyAsync(pc.data(), d_pc, (size_t)G * G * 8, cudaMemcpyDeviceToHost, st));
    CKP(cudaMemcpyAsync(&nsel, d_nsel, 4, cudaMemcpyDeviceToHost, st));
    CKP(cudaStreamSynchronize(st));
    cleanup();
#undef CKP
    if ((uint64_t)nsel != n_csr) {
        m.err = "exchange planning: perm is not a permutation";
        return SBMF_ERR_INVALID;
    }
    for (size_t i = 0; i < pc.size(); ++i) h_pc[i] = (int64_t)pc[i];
    return SBMF_OK;
}

// Test hook (single GPU is enough): the device planner on host arrays, for comparison with sbmf_cuda_plan_exchange.
extern "C" int sbmf_cuda_plan_exchange_device(uint64_t n, const uint32_t* perm, int world, int rank, const int64_t* csr_bounds,
                                              const int64_t* csc_bounds, uint32_t* send_idx, uint32_t* recv_pos, int64_t* pair_counts, int device)
{
    if (!perm || !csr_bounds || !csc_bounds || !send_idx || !recv_pos || !pair_counts || world < 1 || world > MAX_WORLD || rank < 0 || rank >= world)
        return SBMF_ERR_INVALID;
    Model m;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SBMF_ERR_CUDA;
    m.sm_count = prop.multiProcessorCount;
    const uint64_t n_csr = (uint64_t)(csr_bounds[rank + 1] - csr_bounds[rank]), n_csc = (uint64_t)(csc_bounds[rank + 1] - csc_bounds[rank]);
    uint32_t *d_perm = nullptr, *d_s = nullptr, *d_r = nullptr;
    cudaStream_t st = nullptr;
    int rc = SBMF_ERR_CUDA;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess && dmalloc(&d_perm, n) == cudaSuccess &&
        dmalloc(&d_s, n_csr) == cudaSuccess && dmalloc(&d_r, n_csc) == cudaSuccess &&
        cudaMemcpyAsync(d_perm, perm, n * 4, cudaMemcpyHostToDevice, st) == cudaSuccess) {
        rc = plan_exchange_device(m, st, n, d_perm, world, rank, csr_bounds, csc_bounds, d_s, d_r, pair_counts);
        if (rc == SBMF_OK && (cudaMemcpy(send_idx, d_s, n_csr * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
                              cudaMemcpy(recv_pos, d_r, n_csc * 4, cudaMemcpyDeviceToHost) != cudaSuccess))
            rc = SBMF_ERR_CUDA;
    }
    cudaGetLastError();
    cudaFree(d_perm); cudaFree(d_s); cudaFree(d_r);
    if (st) cudaStreamDestroy(st);
    return rc;
}

// multi-GPU: cut the global layout into this rank's CSR shard (its users) and CSC shard (its items) and plan the residual
// all-to-all (plan.cpp).  Every rank built the same global layout from the same COO, so all plans agree.
extern "C" int sbmf_cuda_plan_shards(const int64_t* ptr, uint32_t n_rows, int world, uint32_t* bounds);
extern "C" int sbmf_cuda_plan_pair_counts(uint64_t n, const uint32_t* perm, int world, const int64_t* csr_bounds, const int64_t* csc_bounds,
                                          int64_t* counts);
extern "C" int sbmf_cuda_plan_exchange(uint64_t n, const uint32_t* perm, int world, int rank, const int64_t* csr_bounds, const int64_t* csc_bounds,
                                       uint32_t* send_idx, int64_t* send_counts, uint32_t* recv_pos, int64_t* recv_counts);

static int shard_storage(Model& m)
{
    const int G = m.world, r = m.rank;
    std::vector<int64_t> up((size_t)m.I + 1), ip((size_t)m.J + 1);
    CK(cudaMemcpy(up.data(), m.us.ptr, up.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ip.data(), m.it.ptr, ip.size() * 8, cudaMemcpyDeviceToHost));
    m.ub.assign(G + 1, 0);
    m.ib.assign(G + 1, 0);
    sbmf_cuda_plan_shards(up.data(), m.I, G, m.ub.data());
    sbmf_cuda_plan_shards(ip.data(), m.J, G, m.ib.data());
    std::vector<int64_t> cb(G + 1), tb(G + 1);
    for (int q = 0; q <= G; ++q) {
        cb[q] = up[m.ub[q]];
        tb[q] = ip[m.ib[q]];
    }
    const uint64_t c0 = (uint64_t)cb[r], c1 = (uint64_t)cb[r + 1], t0 = (uint64_t)tb[r], t1 = (uint64_t)tb[r + 1];
    m.n_csr = c1 - c0;
    m.n_csc = t1 - t0;
    {
        // option device_plan (default on): plan on the device (plan_exchange_device above) instead of downloading perm and
        // walking it on the host (plan.cpp, which the CPU tests cover and the GPU tests compare with)
        const bool device_plan = m.opt.device_plan != 0;
        std::vector<int64_t> sc(G), rc(G), pc((size_t)G * G);
        CK(dmalloc(&m.send_idx, m.n_csr)); CK(dmalloc(&m.recv_pos, m.n_csc));
        CK(dmalloc(&m.sendbuf, m.n_csr)); CK(dmalloc(&m.recvbuf, m.n_csc));
        if (device_plan) {
            const int prc = plan_exchange_device(m, m.s_main, m.N, m.perm, G, r, cb.data(), tb.data(), m.send_idx, m.recv_pos, pc.data());
            if (prc != SBMF_OK) return prc;
            for (int q = 0; q < G; ++q) {
                sc[q] = pc[(size_t)r * G + q];   // what I send to q: my users' ratings of q's items
                rc[q] = pc[(size_t)q * G + r];   // what q sends to me
            }
        } else {
            std::vector<uint32_t> perm(m.N ? m.N : 1), sidx(m.n_csr ? m.n_csr : 1), rpos(m.n_csc ? m.n_csc : 1);
            CK(cudaMemcpy(perm.data(), m.perm, m.N * 4, cudaMemcpyDeviceToHost));
            if (sbmf_cuda_plan_exchange(m.N, perm.data(), G, r, cb.data(), tb.data(), sidx.data(), sc.data(), rpos.data(), rc.data()) != SBMF_OK) {
                m.err = "set_train: exchange planning failed (internal)";
                return SBMF_ERR_INVALID;
            }
            sbmf_cuda_plan_pair_counts(m.N, perm.data(), G, cb.data(), tb.data(), pc.data());
            CK(cudaMemcpy(m.send_idx, sidx.data(), m.n_csr * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(m.recv_pos, rpos.data(), m.n_csc * 4, cudaMemcpyHostToDevice));
        }
        {   // where my segments start in every peer's buffers (for the direct-push exchange)
            m.fwd_dst_off.assign(G, 0);
            m.rev_dst_off.assign(G, 0);
            for (int q = 0; q < G; ++q) {
                size_t f = 0, b = 0;
                for (int x = 0; x < r; ++x) {
                    f += (size_t)pc[(size_t)x * G + q];   // residuals rank x sends to q precede mine in q's recvbuf
                    b += (size_t)pc[(size_t)q * G + x];   // residuals q originally sent to x precede mine in q's sendbuf
                }
                m.fwd_dst_off[q] = f;
                m.rev_dst_off[q] = b;
            }
        }
        m.send_off.assign(G, 0); m.send_cnt.assign(G, 0); m.recv_off.assign(G, 0); m.recv_cnt.assign(G, 0);
        size_t so = 0, ro = 0;
        for (int q = 0; q < G; ++q) {
            m.send_off[q] = so; m.send_cnt[q] = (size_t)sc[q]; so += (size_t)sc[q];
            m.recv_off[q] = ro; m.recv_cnt[q] = (size_t)rc[q]; ro += (size_t)rc[q];
        }
    }
    if (m.opt.mgpu_pool) cudaFreeAsync(m.perm, m.s_main);
    else cudaFree(m.perm);
    m.perm = nullptr;
    cudaStream_t st = m.s_main;
    {   // option fuse_exchange: per-slot destination of the user phase's final residual (28 bits of position, 3 of rank)
        bool fits = G <= MAX_PEERS && m.opt.fuse_exchange && m.opt.peer;
        for (int q = 0; q < G && fits; ++q) fits = (uint64_t)(tb[q + 1] - tb[q]) < (1ull << 28);
        if (fits && m.n_csr) {
            XSeg sg;
            sg.world = G;
            for (int q = 0; q < MAX_PEERS; ++q) {
                sg.seg_off[q] = q < G ? m.send_off[q] : m.n_csr;
                sg.dst_off[q] = q < G ? m.fwd_dst_off[q] : 0;
            }
            sg.seg_off[MAX_PEERS] = m.n_csr;
            CK(palloc(m, &m.xmap_fwd, m.n_csr, st));
            SBMF_LAUNCH((xmap_kernel), (uint32_t)std::min<uint64_t>((m.n_csr + 255) / 256, (uint64_t)m.sm_count * 16), 256, 0, st, m.send_idx, m.n_csr, sg,
                        m.xmap_fwd);
        }
    }
    CK(slice_inplace(m, m.us.idx, c0, m.n_csr, st)); CK(slice_inplace(m, m.us.e, c0, m.n_csr, st)); CK(slice_inplace(m, m.csr_urow, c0, m.n_csr, st));
    CK(slice_inplace(m, m.csr_r, c0, m.n_csr, st)); CK(slice_inplace(m, m.csr_id, c0, m.n_csr, st));
    CK(slice_inplace(m, m.it.idx, t0, m.n_csc, st)); CK(slice_inplace(m, m.it.e, t0, m.n_csc, st)); CK(slice_inplace(m, m.csc_id, t0, m.n_csc, st));
    SBMF_LAUNCH((rebase_kernel), (m.I + 256) / 256, 256, 0, m.s_main, m.us.ptr, m.I, (int64_t)c0);
    SBMF_LAUNCH((rebase_kernel), (m.J + 256) / 256, 256, 0, m.s_main, m.it.ptr, m.J, (int64_t)t0);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(m.s_main));
    return SBMF_OK;
}

static int alloc_side_state(Model& m, Side& s, const Side& other)
{
    CK(dmalloc(&s.F, (size_t)m.KB * (s.n + 1) * 8));   // + the all-zero pad row
    // block pairs for the other side's streaming pipeline (kernels.cu: pair_pack_kernel); only where that pipeline has rows
    if (other.n_heavy > 0 && m.KB > 1 && m.opt.pair_gather) CK(dmalloc(&s.F2, (size_t)(m.KB - 1) * (s.n + 1) * 16));
    CK(dmalloc(&s.bias, s.n)); CK(dmalloc(&s.mu_b, s.n)); CK(dmalloc(&s.sigma_b, s.n));
    CK(dmalloc(&s.sigma_k, m.KP)); CK(dmalloc(&s.mu_k, m.KP)); CK(dmalloc(&s.post_var, m.KP)); CK(dmalloc(&s.sigma_kf, m.KP)); CK(dmalloc(&s.mu_kf, m.KP));
    s.hyp_chunks = (s.n + 16383) / 16384;
    if (s.hyp_chunks < 1) s.hyp_chunks = 1;
    CK(dmalloc(&s.hyp_part, (size_t)m.KB * s.hyp_chunks * 16));
    return SBMF_OK;
}

struct Trace {   // option trace = 1: wall-clock of the set_train stages on stderr (= 2: host time only, no device synchronisation)
    bool on, sync;
    explicit Trace(const Model& m) : on(m.opt.trace != 0), sync(m.opt.trace != 2) {}
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char* what)
    {
        if (!on) return;
        if (sync) cudaDeviceSynchronize();
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[sbmf trace] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

int build_storage(Model& m, uint64_t n, const uint32_t* user, const uint32_t* item, const float* rating, uint32_t num_users,
                  uint32_t num_items)
{
    Trace tr(m);
    free_storage(m);
    tr.lap("free previous");
    if (n >= (1ull << 31)) {
        m.err = "set_train: more than 2^31-1 ratings per GPU are not supported (shard across GPUs)";
        return SBMF_ERR_UNSUPPORTED;
    }
    m.N = n; m.I = num_users; m.J = num_items;
    m.us.n = num_users; m.it.n = num_items;
    cudaStream_t st = m.s_main;
    const int T = 256;
    const uint32_t G = (uint32_t)std::min<uint64_t>((n + T - 1) / T + 1, (uint64_t)m.sm_count * 16);

    uint32_t *d_user = nullptr, *d_item = nullptr, *d_iota = nullptr, *d_keys = nullptr, *d_inv = nullptr, *d_max = nullptr;
    float* d_rating = nullptr;
    void* d_tmp = nullptr;
    // temporaries come from the device's stream-ordered pool (release threshold raised in sbmf_cuda_create): a plain
    // cudaFree of a few GB can take hundreds of milliseconds
    auto talloc = [&](void** p, size_t bytes) { return cudaMallocAsync(p, bytes ? bytes : 1, st); };
    auto cleanup = [&]() {
        for (void* p : {(void*)d_user, (void*)d_item, (void*)d_iota, (void*)d_keys, (void*)d_inv, (void*)d_max, (void*)d_rating, d_tmp})
            if (p) cudaFreeAsync(p, st);
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

    CKC(talloc((void**)&d_user, n * 4)); CKC(talloc((void**)&d_item, n * 4)); CKC(talloc((void**)&d_rating, n * 4));
    CKC(talloc((void**)&d_iota, n * 4)); CKC(talloc((void**)&d_keys, n * 4)); CKC(talloc((void**)&d_inv, n * 4)); CKC(talloc((void**)&d_max, 8));
    CKC(cudaMemcpyAsync(d_user, user, n * 4, cudaMemcpyHostToDevice, st));
    CKC(cudaMemcpyAsync(d_item, item, n * 4, cudaMemcpyHostToDevice, st));
    CKC(cudaMemcpyAsync(d_rating, rating, n * 4, cudaMemcpyHostToDevice, st));
    tr.lap("alloc + H2D");
    CKC(cudaMemsetAsync(d_max, 0, 8, st));
    if (n) SBMF_LAUNCH((max_id_kernel), G, T, 0, st, d_user, d_item, n, d_max);
    uint32_t h_max[2] = {0, 0};
    CKC(cudaMemcpyAsync(h_max, d_max, 8, cudaMemcpyDeviceToHost, st));
    CKC(cudaStreamSynchronize(st));
    if (n && (h_max[0] >= num_users || h_max[1] >= num_items)) {
        m.err = "set_train: user/item id out of range (max user " + std::to_string(h_max[0]) + ", max item " + std::to_string(h_max[1]) + ")";
        cleanup();
        return SBMF_ERR_INVALID;
    }

    CKC(palloc(m, &m.us.ptr, (size_t)num_users + 1, st)); CKC(palloc(m, &m.it.ptr, (size_t)num_items + 1, st));
    CKC(palloc(m, &m.us.idx, n, st)); CKC(palloc(m, &m.it.idx, n, st)); CKC(palloc(m, &m.us.e, n, st)); CKC(palloc(m, &m.it.e, n, st));
    CKC(palloc(m, &m.csr_urow, n, st)); CKC(palloc(m, &m.csr_r, n, st)); CKC(palloc(m, &m.csr_id, n, st)); CKC(palloc(m, &m.csc_id, n, st));
    CKC(palloc(m, &m.perm, n, st));

    tr.lap("id check + alloc layout");
    size_t tmp_bytes = 0;
    const bool relabel = m.opt.relabel != 0;
    const uint64_t n_sort = relabel ? std::max<uint64_t>(n, std::max(num_users, num_items)) : n;   // the row ranking sorts too
    CKC(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_user, d_keys, d_iota, m.csr_id, (int)n_sort, 0, 32, st));
    CKC(talloc(&d_tmp, tmp_bytes));
    if (relabel) {
        // positions by decreasing rating count, per side; then the COO itself moves to position space
        for (Side* sd : {&m.us, &m.it}) {
            const uint32_t nr = sd->n;
            const uint32_t* ids = (sd == &m.us) ? d_user : d_item;
            CKC(dmalloc(&sd->id_at, nr)); CKC(dmalloc(&sd->pos_of, nr));
            uint32_t* d_rk = nullptr;   // degrees | keys | sorted keys | ranked ids, nr words each
            CKC(talloc((void**)&d_rk, (size_t)nr * 16));
            uint32_t *deg = d_rk, *key = d_rk + nr, *key_sorted = d_rk + 2 * (size_t)nr, *ranked = d_rk + 3 * (size_t)nr;
            cudaError_t e = cudaMemsetAsync(deg, 0, (size_t)nr * 4, st);
            if (e == cudaSuccess) {
                if (n) SBMF_LAUNCH((degree_kernel), G, T, 0, st, ids, n, deg);
                SBMF_LAUNCH((rank_key_kernel), (nr + T - 1) / T, T, 0, st, deg, nr, key, sd->pos_of);   // pos_of = identity, as the sort's values
                e = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, key, key_sorted, sd->pos_of, ranked, (int)nr, 0, 32, st);
            }
            if (e == cudaSuccess) {
                SBMF_LAUNCH((assign_pos_kernel), (nr + T - 1) / T, T, 0, st, ranked, nr, (uint32_t)m.world, sd->id_at, sd->pos_of);
                sd->h_id_at.resize(nr);
                e = cudaMemcpyAsync(sd->h_id_at.data(), sd->id_at, (size_t)nr * 4, cudaMemcpyDeviceToHost, st);
            }
            cudaFreeAsync(d_rk, st);
            CKC(e);
        }
        if (n) {
            SBMF_LAUNCH((relabel_kernel), G, T, 0, st, d_user, n, m.us.pos_of);
            SBMF_LAUNCH((relabel_kernel), G, T, 0, st, d_item, n, m.it.pos_of);
        }
        tr.lap("relabel");
    }
    SBMF_LAUNCH((iota_kernel), G, T, 0, st, d_iota, n);
    if (!relabel) {
        // CSR: stable sort of the rating index by user
        CKC(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_user, m.csr_urow, d_iota, m.csr_id, (int)n, 0, bits_for(num_users), st));
        // CSC: stable sort of the rating index by item
        CKC(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_item, d_keys, d_iota, m.csc_id, (int)n, 0, bits_for(num_items), st));
    } else {
        // rows sorted by the opposite side's position as well (that is what puts the popular rows into adjacent lanes): three
        // stable passes, LSD fashion -- by user; then by item = CSC order (item, user); then by user again = CSR order (user, item)
        CKC(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_user, m.csr_urow, d_iota, d_inv, (int)n, 0, bits_for(num_users), st));
        SBMF_LAUNCH((gather_u32_kernel), G, T, 0, st, d_item, d_inv, m.us.idx, n);
        CKC(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, m.us.idx, d_keys, d_inv, m.csc_id, (int)n, 0, bits_for(num_items), st));
        SBMF_LAUNCH((gather_u32_kernel), G, T, 0, st, d_user, m.csc_id, m.us.idx, n);
        CKC(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, m.us.idx, m.csr_urow, m.csc_id, m.csr_id, (int)n, 0, bits_for(num_users), st));
    }
    SBMF_LAUNCH((row_ptr_kernel), (num_users + 1 + T - 1) / T, T, 0, st, m.csr_urow, n, num_users, m.us.ptr);
    SBMF_LAUNCH((gather_u32_kernel), G, T, 0, st, d_item, m.csr_id, m.us.idx, n);
    SBMF_LAUNCH((gather_f32_kernel), G, T, 0, st, d_rating, m.csr_id, m.csr_r, n);
    SBMF_LAUNCH((invert_kernel), G, T, 0, st, m.csr_id, d_inv, n);
    SBMF_LAUNCH((row_ptr_kernel), (num_items + 1 + T - 1) / T, T, 0, st, d_keys, n, num_items, m.it.ptr);
    SBMF_LAUNCH((gather_u32_kernel), G, T, 0, st, d_user, m.csc_id, m.it.idx, n);
    SBMF_LAUNCH((gather_u32_kernel), G, T, 0, st, d_inv, m.csc_id, m.perm, n);
    CKC(cudaMemsetAsync(m.us.e, 0, (n ? n : 1) * 4, st));
    CKC(cudaMemsetAsync(m.it.e, 0, (n ? n : 1) * 4, st));
    CKC(cudaGetLastError());
    CKC(cudaStreamSynchronize(st));
    tr.lap("sorts + gathers");
    cleanup();
    tr.lap("free temporaries");
#undef CKC

    int rc;
    m.n_csr = m.n_csc = n;
    m.ub.assign(2, 0); m.ib.assign(2, 0);
    m.ub[1] = num_users; m.ib[1] = num_items;
    if (m.world > 1 && (rc = shard_storage(m)) != SBMF_OK) return rc;
    if ((rc = build_worklists(m, m.us, m.ub[m.rank], m.ub[m.rank + 1])) != SBMF_OK) return rc;
    if ((rc = build_worklists(m, m.it, m.ib[m.rank], m.ib[m.rank + 1])) != SBMF_OK) return rc;
    tr.lap("shard + work lists");
    if ((rc = alloc_side_state(m, m.us, m.it)) != SBMF_OK) return rc;
    if ((rc = alloc_side_state(m, m.it, m.us)) != SBMF_OK) return rc;
    m.red_blocks = (uint32_t)m.sm_count * 8;
    CK(dmalloc(&m.red_part, (size_t)m.red_blocks * 2));
    CK(dmalloc(&m.red2, 2));
    CK(palloc(m, &m.pacc, m.n_csr, m.s_main));
    CK(dmalloc(&m.bar, 1));
    CK(cudaMemset(m.bar, 0, 8));
    if (m.world > 1 && (rc = setup_peer_access(m)) != SBMF_OK) return rc;
    tr.lap("alloc state");
    m.have_train = true;
    m.e_in_csc = false;
    return SBMF_OK;
}

int build_test(Model& m, uint64_t nt, const uint32_t* user, const uint32_t* item, const float* rating)
{
    free_test(m);
    m.Nt = nt;
    CK(palloc(m, &m.t_user, nt, m.s_main)); CK(palloc(m, &m.t_item, nt, m.s_main)); CK(palloc(m, &m.t_r, nt, m.s_main));
    CK(palloc(m, &m.t_sum, nt, m.s_main));
    if (nt) {
        CK(cudaMemcpyAsync(m.t_user, user, nt * 4, cudaMemcpyHostToDevice, m.s_main));
        CK(cudaMemcpyAsync(m.t_item, item, nt * 4, cudaMemcpyHostToDevice, m.s_main));
        CK(cudaMemcpyAsync(m.t_r, rating, nt * 4, cudaMemcpyHostToDevice, m.s_main));
    }
    if (nt && m.us.pos_of) {   // relabelled model: the test pairs move to position space like the training triples
        const uint32_t g = (uint32_t)std::min<uint64_t>((nt + 255) / 256, (uint64_t)m.sm_count * 16);
        SBMF_LAUNCH((relabel_kernel), g, 256, 0, m.s_main, m.t_user, nt, m.us.pos_of);
        SBMF_LAUNCH((relabel_kernel), g, 256, 0, m.s_main, m.t_item, nt, m.it.pos_of);
    }
    CK(cudaMemsetAsync(m.t_sum, 0, (nt ? nt : 1) * 8, m.s_main));
    CK(cudaStreamSynchronize(m.s_main));
    m.t_begin = nt * (uint64_t)m.rank / (uint64_t)m.world;
    m.t_end = nt * (uint64_t)(m.rank + 1) / (uint64_t)m.world;
    m.have_test = true;
    return SBMF_OK;
}

// The layout of sbmf_cuda_get_layout for a relabelled model: the caller's COO is restored from the CSR arrays and [T]'s layout
// ([T]:156-221: both orders stable w.r.t. file order, caller's ids) is built from it by the same two stable sorts build_storage
// runs with relabel = 0, into temporaries that only serve this export.  One GPU only (like get_layout).
int export_reference_layout(Model& m, int64_t* row_ptr, uint32_t* col, uint64_t* csr_id, int64_t* col_ptr, uint32_t* row, uint64_t* csc_id, uint64_t* perm)
{
    const uint64_t n = m.N;
    cudaStream_t st = m.s_main;
    const int T = 256;
    const uint32_t G = (uint32_t)std::min<uint64_t>((n + T - 1) / T + 1, (uint64_t)m.sm_count * 16);
    uint32_t *d_user = nullptr, *d_item = nullptr, *d_iota = nullptr, *d_keys = nullptr, *d_id = nullptr, *d_inv = nullptr, *d_out = nullptr;
    int64_t* d_ptr = nullptr;
    void* d_tmp = nullptr;
    auto cleanup = [&]() {
        for (void* p : {(void*)d_user, (void*)d_item, (void*)d_iota, (void*)d_keys, (void*)d_id, (void*)d_inv, (void*)d_out, (void*)d_ptr, d_tmp})
            if (p) cudaFreeAsync(p, st);
        cudaStreamSynchronize(st);
    };
#define CKX(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            m.err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            cleanup();                                                                             \
            return (e_ == cudaErrorMemoryAllocation) ? SBMF_ERR_NOMEM : SBMF_ERR_CUDA;             \
        }                                                                                          \
    } while (0)
    auto talloc = [&](void** p, size_t bytes) { return cudaMallocAsync(p, bytes ? bytes : 1, st); };
    const size_t nrow_max = (size_t)std::max(m.I, m.J) + 1;
    CKX(talloc((void**)&d_user, n * 4)); CKX(talloc((void**)&d_item, n * 4)); CKX(talloc((void**)&d_iota, n * 4)); CKX(talloc((void**)&d_keys, n * 4));
    CKX(talloc((void**)&d_id, n * 4)); CKX(talloc((void**)&d_inv, n * 4)); CKX(talloc((void**)&d_out, n * 4)); CKX(talloc((void**)&d_ptr, nrow_max * 8));
    size_t tmp_bytes = 0;
    CKX(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_user, d_keys, d_iota, d_id, (int)n, 0, 32, st));
    CKX(talloc(&d_tmp, tmp_bytes));
    if (n) SBMF_LAUNCH((restore_coo_kernel), G, T, 0, st, m.csr_id, m.csr_urow, m.us.idx, m.us.id_at, m.it.id_at, n, d_user, d_item);
    SBMF_LAUNCH((iota_kernel), G, T, 0, st, d_iota, n);
    std::vector<uint32_t> tmp32(n ? n : 1);
    auto out_u64 = [&](uint64_t* dst, const uint32_t* src) -> cudaError_t {
        cudaError_t e = cudaMemcpyAsync(tmp32.data(), src, n * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess)
            for (uint64_t i = 0; i < n; ++i) dst[i] = tmp32[i];
        return e;
    };
    // CSR: stable sort of the rating index by user
    CKX(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_user, d_keys, d_iota, d_id, (int)n, 0, bits_for(m.I), st));
    SBMF_LAUNCH((row_ptr_kernel), (m.I + 1 + T - 1) / T, T, 0, st, d_keys, n, m.I, d_ptr);
    if (row_ptr) CKX(cudaMemcpyAsync(row_ptr, d_ptr, ((size_t)m.I + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (col) {
        SBMF_LAUNCH((gather_u32_kernel), G, T, 0, st, d_item, d_id, d_out, n);
        CKX(cudaMemcpyAsync(col, d_out, n * 4, cudaMemcpyDeviceToHost, st));
    }
    CKX(cudaStreamSynchronize(st));
    if (csr_id) CKX(out_u64(csr_id, d_id));
    SBMF_LAUNCH((invert_kernel), G, T, 0, st, d_id, d_inv, n);
    // CSC: stable sort of the rating index by item
    CKX(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_item, d_keys, d_iota, d_id, (int)n, 0, bits_for(m.J), st));
    SBMF_LAUNCH((row_ptr_kernel), (m.J + 1 + T - 1) / T, T, 0, st, d_keys, n, m.J, d_ptr);
    if (col_ptr) CKX(cudaMemcpyAsync(col_ptr, d_ptr, ((size_t)m.J + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (row) {
        SBMF_LAUNCH((gather_u32_kernel), G, T, 0, st, d_user, d_id, d_out, n);
        CKX(cudaMemcpyAsync(row, d_out, n * 4, cudaMemcpyDeviceToHost, st));
    }
    CKX(cudaStreamSynchronize(st));
    if (csc_id) CKX(out_u64(csc_id, d_id));
    if (perm) {
        SBMF_LAUNCH((gather_u32_kernel), G, T, 0, st, d_inv, d_id, d_out, n);
        CKX(out_u64(perm, d_out));
    }
    CKX(cudaGetLastError());
    cleanup();
#undef CKX
    return SBMF_OK;
}

}  // namespace sbmf
