// model.h -- host-side model of one SBMF replica and the launch wrappers of its kernels.
// Internal to libsbmf_cuda.so; the public boundary is include/sbmf_cuda.h.
//
// Storage (replaces the jagged R / R_t of gibbs_sbpmf2.cpp "[T]":156-221):
//   user side: CSR  ptr[I+1] (int64), idx[N] = item of each slot, e[N] = residual in CSR slot order
//   item side: CSC  ptr[J+1],         idx[N] = user of each slot, e[N] = residual in CSC slot order
//   perm[N] = CSR slot of each CSC slot (replaces [T]'s `.id` back-pointers)
// Factors are stored "K8-blocked": F[b][row][8] with b = k/8 -- the 8 values of one row in one block are one
// 32-byte sector, fetched by one LDG.256, which is the unit the row kernels gather.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/sbmf_cuda.h"
#include "comm.h"

namespace sbmf {

constexpr int KBLK = 8;             // latent dimensions per factor block (one 32-byte sector per row)
constexpr int NACC = 48;            // g[8] + upper-triangular G[36] = 44 accumulators, padded to 48
constexpr int NBINS = 12;            // resident-row bins, see kBin* below
constexpr int RESIDENT_MAX = 2048;  // longer rows go through the streaming ("heavy") pipeline
constexpr int SLICE_LEN = 4096;     // max ratings per heavy-row slice (one 64-thread CTA); shorter when a shard is small (storage.cu)
constexpr int MAX_PEERS = 8;         // replicas a phase kernel can write directly (one NVSwitch domain)

// resident bins: a row of c ratings is handled by WARPS warps holding RPL ratings per lane in registers
struct BinShape { int rpl, warps, cap; };
// warps == 0: short rows, cap / rpl lanes (8 or 16) per row, several rows per warp (row_group_kernel)
constexpr BinShape kBins[NBINS] = {{4, 0, 32},  {8, 0, 64},  {6, 0, 96},  {8, 0, 128},  {6, 1, 192},  {8, 1, 256},
                                   {3, 4, 384}, {4, 4, 512}, {6, 4, 768}, {8, 4, 1024}, {6, 8, 1536}, {8, 8, 2048}};

// Device-resident scalars of the sweep ([T]:315-318 plus the statistics of [T]:342-359).
struct Scalars {
    double b_0, alpha, mu_b_0, sigma_b_0;
    double sum_e, sum_e2;
    double rmse_mean, rmse_sweep;
    float alpha_f, b_0_f;
    float shift_f;       // b0_old - b0_new of [T]:407-410, folded into the user phase's first residual touch
    uint32_t sweep;      // 0-based sweep counter = Philox counter word 3; bumped by the eval kernel
};

struct Slice {
    int64_t start;       // first slot of the slice in this side's order
    uint32_t len;
    uint32_t hrow;       // index into Side::heavy_rows
};

// One side of the model: users with the CSR layout, or items with the CSC layout.
struct Side {
    uint32_t n = 0;                   // rows on this side (I or J)
    // Row relabelling (option relabel, storage.cu): every array of the model is indexed by a row's POSITION, positions ordered by
    // decreasing number of ratings.  id_at[pos] = the caller's row id (Philox keys, exports), pos_of[id] = position (imports).
    // nullptr = identity (relabel off).
    uint32_t* id_at = nullptr;        // [n]
    uint32_t* pos_of = nullptr;       // [n]
    std::vector<uint32_t> h_id_at;    // host copy of id_at (empty = identity)
    int64_t* ptr = nullptr;           // [n+1]
    uint32_t* idx = nullptr;          // [N] opposite-side id of each slot
    float* e = nullptr;               // [N] residual e_ij in this side's slot order
    float* F = nullptr;               // [KB][n][8]  K8-blocked factors (own rows; gather target of the other side)
    float* F2 = nullptr;              // [KB-1][n+1][16] block pairs (pb, pb+1) side by side, rebuilt per phase; only if the OTHER side streams rows
    float *bias = nullptr, *mu_b = nullptr, *sigma_b = nullptr;   // [n]
    double *sigma_k = nullptr, *mu_k = nullptr;                   // [KP] hyper-parameters (fp64 masters)
    double* post_var = nullptr;                                   // [KP] posterior variance of mu_k (Normal-Gamma modes, [S]:391/411)
    float *sigma_kf = nullptr, *mu_kf = nullptr;                  // [KP] fp32 mirrors read by the row kernels
    double* hyp_part = nullptr;       // [KB][hyp_chunks][16] partial (sum, sumsq) of the per-dimension hyper step
    uint32_t hyp_chunks = 0;
    // work lists
    uint32_t* bin_rows[NBINS] = {};
    uint32_t bin_count[NBINS] = {};
    uint32_t n_heavy = 0, n_slices = 0;
    uint32_t h_split = 0, s_split = 0;    // streamed rows [0, h_split) own slices [0, s_split): about half of them (two-chain pipeline)
    uint32_t* heavy_rows = nullptr;       // [n_heavy] row id
    uint32_t* heavy_slice_ptr = nullptr;  // [n_heavy+1] slice range of each heavy row
    Slice* slices = nullptr;              // [n_slices]
    float* hpart = nullptr;               // [n_slices][NACC]
    uint32_t* hcount = nullptr;           // [n_heavy] tickets of the slice CTAs of a row in the current pass (0 between launches)
    float* hdelta = nullptr;              // [n_heavy][8] pending factor deltas of the block just solved, then [n_heavy] bias deltas
    uint64_t nnz_resident = 0, nnz_heavy = 0;
    uint32_t site_f = 0, site_b = 0;      // Philox streams of the factor / bias draws
    uint32_t site_sigma_k = 0, site_mu_k = 0, site_sigma_b = 0, site_mu_b = 0;
    int prior = 0;                        // index into sbmf_priors of the factor hyper-prior (2 = users, 1 = items)
    int prior_b = 0;                      // ... of the bias hyper-prior (4 = users, 5 = items)
};

// Tuning / developer options of a handle (sbmf_cuda_set_option; include/sbmf_cuda.h lists them).  They replace the
// process-wide getenv knobs of round 1: every value lives in the handle, is set through the ABI and is covered by tests.
struct Options {
    int64_t l2_budget_mb = 192;   // resident bins: bytes of gathered factor blocks one launch keeps in flight (launch_phase)
    int64_t max_blocks_per_launch = 0;   // resident bins: cap on the factor blocks one launch processes (0 = only the L2 budget)
    int64_t resident_max = RESIDENT_MAX;   // rows longer than this stream through the sliced pipeline (<= RESIDENT_MAX); sets both sides
    int64_t resident_max_user = RESIDENT_MAX, resident_max_item = RESIDENT_MAX;   // ... per side
    int64_t slice_len = 0;        // heavy-row slice length; 0 = chosen from the shard size (build_worklists)
    int64_t relabel = 1;          // rows stored in order of decreasing rating count (adjacent lanes of a gather then share 128-byte lines
                                  // on the popular rows); 0 = the caller's ids.  Draws are keyed by the caller's ids either way
    int64_t group_rows = 1;       // short rows: several rows per warp (row_group_kernel); 0 = one warp per row
    int64_t row_kernels = 3;      // resident rows: 3 = rows2.cuh with the shuffle reduction (one barrier per block, 64-bit pair operands: 22.12 -> 21.87 ms
                                  // per sweep on the final build), 1 = kernels.cu (round 1), 2 = rows2.cuh with the shared-memory reduction (slower)
    int64_t alt_bins = 1;         // resident rows of 193..512 ratings: 2 warps x (6 | 8) per lane (rows2.cuh) instead of 4 warps x (3 | 4); measured 12.19 -> 11.77 ms user phase
    int64_t pair_gather = 0;      // streaming pipeline: gather (previous, current) block as one 64-byte row by lane pairs (0: two sector gathers)
    int64_t fuse_solve = 0;       // streaming pipeline: row updates in the tail of each pass by the row's last slice CTA (0: a launch of
                                  // their own; measured on B200: the release fence + ticket per slice CTA costs 160 us per pass, a launch 55)
    int64_t heavy_chains = 2;     // streaming pipeline: the streamed rows as two independent pass -> update -> pass chains on two streams, so
                                  // that the (latency-bound, one-CTA-per-row) updates of one half run under the passes of the other
    int64_t fold_user = 1;        // one GPU: CSC->CSR residual hand-over folded into the user phase's first touch
    int64_t fold_item = 1;        // one GPU: CSR->CSC hand-over folded into the item phase's first touch (22.46 -> 22.27 ms per sweep)
    int64_t graph = 1;            // replay the steady-state sweep from a CUDA graph when per-phase timing is off
    int64_t device_plan = 1;      // multi-GPU: exchange plan computed on the device (0 = host planner, plan.cpp)
    int64_t mgpu_pool = 1;        // multi-GPU: rating-sized arrays from the stream-ordered pool (0 = cudaMalloc)
    int64_t fuse_exchange = 0;    // multi-GPU with peer pushes: the user phase stores its final residual directly into the owner rank's receive
                                  // buffer (over NVLink, overlapped with the phase) and the item phase reads it through recv_pos: no push /
                                  // unpack kernels between the phases, only the barrier.  Measured: the exchange drops from 0.64 to 0.04 ms
                                  // at N = 2 and 0.24 to 0.05 at N = 8, but the peer stores slow the phase itself by more (N = 2: 12.34 vs
                                  // 12.38 ms per sweep; N = 8: 4.47 vs 3.91) -- off
    int64_t peer = 1;             // multi-GPU: peer-mapped replicas / direct NVLink pushes (0 = NCCL exchanges)
    int64_t trace = 0;            // 1: wall-clock of the set_train stages on stderr; 2: without device synchronisation
};

struct Model {
    sbmf_config cfg{};
    Options opt;
    int device = 0;
    int sm_count = 148;
    uint32_t K = 0, KB = 0, KP = 0;
    uint64_t N = 0, Nt = 0;
    uint32_t I = 0, J = 0;
    bool have_train = false, have_test = false, have_factors = false;
    uint32_t sweeps_done = 0;
    uint32_t sweeps_since_init = 0;   // sweeps this handle has launched since init_factors / set_state (lazily built maps exist after 2)

    Side us, it;                      // user side (CSR), item side (CSC)
    uint32_t* csr_urow = nullptr;     // [N] user of each CSR slot (COO row index for the flat rebuild)
    float* csr_r = nullptr;           // [N] rating per CSR slot
    float* pacc = nullptr;            // [N] partial predictions of the fused residual refresh (CSR slot order)
    uint32_t* csr_id = nullptr;       // [N] rating (file) index of each CSR slot
    uint32_t* csc_id = nullptr;       // [N] rating index of each CSC slot
    uint32_t* perm = nullptr;         // [N] CSR slot of each CSC slot
    uint32_t *perm_inv = nullptr, *recv_pos_inv = nullptr, *send_idx_inv = nullptr;   // inverse maps for the reverse direction, built on first use
    bool e_in_csc = false;            // where the freshest residual lives
    bool need_rebuild = false;        // set_state without a residual: the next sweep rebuilds it stand-alone, like sweep 0
    // multi-GPU: rank r owns users [ub[r], ub[r+1]) with their CSR slots and items [ib[r], ib[r+1]) with their
    // CSC slots; every slot array above is then the LOCAL shard, ptr[] is rebased to local slots, factors/biases are replicas
    Comm comm;
    int rank = 0, world = 1;
    uint64_t n_csr = 0, n_csc = 0;    // local slot counts (== N on one GPU)
    std::vector<uint32_t> ub, ib;     // [world+1] row bounds
    uint32_t* send_idx = nullptr;     // [n_csr] local CSR slot of each send-buffer position (grouped by destination rank)
    uint32_t* recv_pos = nullptr;     // [n_csc] receive-buffer position of each local CSC slot (grouped by source rank)
    float *sendbuf = nullptr, *recvbuf = nullptr;
    std::vector<size_t> send_off, send_cnt, recv_off, recv_cnt;
    // peer-mapped replicas (CUDA IPC over NVLink, world <= MAX_PEERS): every rank writes the rows it updates straight into all
    // replicas from inside the phase kernels, and pushes its residual segments into the peers' exchange buffers; NCCL then only
    // carries two 16-byte all-reduces per sweep plus tiny barriers.  Falls back to the NCCL exchanges if IPC is unavailable.
    bool peer_ok = false;
    float* pF[2][8] = {};             // [side: 0 users, 1 items][rank] factor replica of that rank (own entry = local pointer)
    float* pbias[2][8] = {};
    float* precv[8] = {};             // peers' recvbuf (forward exchange target)
    float* psend[8] = {};             // peers' sendbuf (reverse exchange target)
    std::vector<void*> ipc_opened;
    std::vector<size_t> fwd_dst_off, rev_dst_off;   // start of MY segment in rank q's recvbuf / sendbuf
    // fused forward exchange (option fuse_exchange): xmap_fwd[local CSR slot] = (destination rank << 28) | position in that rank's
    // recvbuf -- the user phase's final residual store goes straight there (kernels.cu store_e_final); nullptr = push_kernel path
    uint32_t* xmap_fwd = nullptr;
    double* bar = nullptr;            // 1 double: payload of the barrier all-reduce
    uint64_t t_begin = 0, t_end = 0;  // this rank's slice of the test set
    double* red2 = nullptr;           // [2] reduced (and all-reduced) pair of sums: (sum e, sum e^2) or the two squared-error sums
    // test set
    uint32_t *t_user = nullptr, *t_item = nullptr;
    float* t_r = nullptr;
    double* t_sum = nullptr;          // running prediction sum, [T]:145, 629
    Scalars* sc = nullptr;
    double* red_part = nullptr;       // [red_blocks][2] per-block partial sums of the reductions
    uint32_t red_blocks = 0;
    double* rmse_hist = nullptr;      // [hist_cap][2]
    uint32_t hist_cap = 0;

    cudaStream_t s_main = nullptr, s_aux = nullptr;
    cudaStream_t s_aux2 = nullptr;     // second chain of the streaming pipeline (option heavy_chains)
    cudaEvent_t ev_join2 = nullptr;
    cudaStream_t s_res[2] = {};        // extra streams for the resident bins
    cudaEvent_t ev_join_res[2] = {};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_t[8] = {};
    cudaEvent_t ev_call[2] = {};
    cudaEvent_t ev_c[2] = {};          // start of the post-phase all-gathers (multi-GPU)
    std::vector<cudaEvent_t> ev_top;   // 2 per launch of the dominant kernel in one phase
    uint32_t ev_top_used = 0;
    cudaGraphExec_t graph_exec = nullptr;   // captured steady-state sweep (api.cu: graph_sweep)
    bool gk[4] = {};                        // its SweepKey
    bool graph_failed = false;
    uint64_t graph_launches = 0;
    bool timing_enabled = true;
    bool timing_detail = false;
    sbmf_timing timing{};
    uint64_t launches = 0;
    std::string err;
};

// ---- storage.cu
int build_storage(Model& m, uint64_t n, const uint32_t* user, const uint32_t* item, const float* rating,
                  uint32_t num_users, uint32_t num_items);
int build_test(Model& m, uint64_t nt, const uint32_t* user, const uint32_t* item, const float* rating);
void free_storage(Model& m);
void free_test(Model& m);

// ---- kernels.cu: launch wrappers (all asynchronous on the given stream)
int export_reference_layout(Model& m, int64_t* row_ptr, uint32_t* col, uint64_t* csr_id, int64_t* col_ptr, uint32_t* row, uint64_t* csc_id, uint64_t* perm);
void init_constant_tables();   // once per device, after cudaSetDevice
void launch_init_factors(Model& m, Side& s, uint32_t site, cudaStream_t st);
// host layout -> K8-blocked.  dim_major: src is [K][n] (V of [T]:234-237), else [n][K] (U of [T]:229-232)
void launch_load_factors(Model& m, Side& s, const float* d_src, bool dim_major, cudaStream_t st);
void launch_export_factors(Model& m, const Side& s, float* d_out, bool dim_major, cudaStream_t st);
void launch_rebuild(Model& m, cudaStream_t st);          // [T]:342-359 into the CSR-order residual + partial stats
void launch_stats(Model& m, cudaStream_t st, bool from_csc = false);   // partial stats of the existing residual (either slot order: same multiset)
void launch_global_hypers(Model& m, cudaStream_t st);    // [T]:366-410 (final reduce of the stats + 4 scalar draws)
void launch_dim_hypers(Model& m, cudaStream_t st);       // [T]:415-467
void launch_bias_hypers(Model& m, cudaStream_t st);      // [T]:469-511
// [T]:514-558 / 563-606; e_map != nullptr: the residual is taken over from e_src[e_map[slot]] (the other side's slot order)
void launch_phase(Model& m, Side& self, const Side& other, bool apply_shift, bool refresh, const float* e_src = nullptr,
                  const uint32_t* e_map = nullptr, bool push_final = false);
bool ensure_perm_inverse(Model& m, cudaStream_t st);    // perm_inv (CSR slot -> CSC slot), built on first use; false: no memory
// one GPU: gather through perm.  G GPUs: all-to-all over NVLink, grouped with the all-gather of gather_side's updated rows (may be null)
int launch_permute(Model& m, bool csr_to_csc, Side* gather_side, cudaStream_t st);
int setup_peer_access(Model& m);                                  // storage.cu: exchange IPC handles, map the peers' buffers
void close_peer_access(Model& m);
int launch_barrier(Model& m, cudaStream_t st);                    // cross-GPU barrier on the stream (1-element all-reduce)
int launch_reduce_pair(Model& m, cudaStream_t st);                // red_part -> red2 (+ all-reduce over ranks)
int launch_allgather_side(Model& m, Side& s, cudaStream_t st);    // replicate the rows each rank just updated (factors + bias)
void launch_eval(Model& m, cudaStream_t st);             // [T]:610-636: prediction + partial squared errors of this rank's slice
void launch_eval_final(Model& m, cudaStream_t st);
void launch_pred_mean(Model& m, float* d_out, double denom, cudaStream_t st);   // posterior-mean prediction = running sum / collected sweeps       // RMSE from the reduced sums, history, sweep counter

}  // namespace sbmf

struct sbmf_handle {
    sbmf::Model m;
};
