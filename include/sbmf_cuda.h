/*
 * sbmf_cuda.h -- C ABI of the B200-native SBMF Gibbs sweep (libsbmf_cuda.so).
 *
 * The reference (rishabhmisra/Scalable-Bayesian-Matrix-Factorization) has no plugin / FFI boundary for
 * this path: gibbs_sbpmf2.cpp ("[T]") is one monolithic main().  This header therefore DEFINES the
 * boundary, one entry point per phase of [T]; each comment cites the reference lines it replaces.
 * INTEGRATION.md shows the host-side binding (C++ main, ctypes) a maintainer would add.
 *
 * Conventions: every function returns SBMF_OK (0) or a negative sbmf_status; nothing throws across the
 * ABI; sbmf_cuda_last_error() returns a human-readable message for the last failure on that handle
 * (or the last create failure when handle == NULL).  Host buffers are caller-owned and only touched
 * during the call.  Device memory is owned by the handle.  One handle = one model replica on one GPU
 * (one rank of a multi-GPU job); calls on a handle come from one host thread.  There is NO CPU
 * fallback: if no sm_100 device is usable, sbmf_cuda_create fails.
 */
#ifndef SBMF_CUDA_H_
#define SBMF_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBMF_CUDA_ABI_VERSION 2
#define SBMF_MAX_K 256

typedef struct sbmf_handle sbmf_handle;

typedef enum sbmf_status {
    SBMF_OK = 0,
    SBMF_ERR_INVALID = -1,      /* bad argument (null pointer, id out of range, K > SBMF_MAX_K, ...) */
    SBMF_ERR_CUDA = -2,         /* a CUDA runtime call or kernel failed */
    SBMF_ERR_NOMEM = -3,        /* device or host allocation failed */
    SBMF_ERR_STATE = -4,        /* call order violated (e.g. sweep before set_train / init_factors) */
    SBMF_ERR_NCCL = -5,         /* NCCL could not be loaded or a collective failed */
    SBMF_ERR_UNSUPPORTED = -6   /* valid request this build does not implement */
} sbmf_status;

/* How a draw x ~ N(mu*, 1/lambda*) is realised.  [T] passes the posterior VARIANCE 1/lambda* to
   ran_gaussian(mean, stdev) ([T]:393,406,439,465,487,509,530,551,578,599; random.h:166-172), so the
   reference-compatible mode draws x = mu* + (1/lambda*) z.  SURVEY.md 0.3. */
typedef enum sbmf_sample_mode {
    SBMF_SAMPLE_REF_VAR_AS_STDEV = 0,   /* x = mu* + (1/lambda*) z      -- what [T] does */
    SBMF_SAMPLE_SQRT = 1,               /* x = mu* + sqrt(1/lambda*) z  -- the textbook conditional */
    SBMF_SAMPLE_ZERO_NOISE = 2          /* x = mu*; Gamma(a,b) -> a/b   -- conditional-mean updates (parity mode) */
} sbmf_sample_mode;

typedef enum sbmf_hyper_mode {
    SBMF_HYPER_REF_T = 0,               /* hyper-parameter updates exactly as [T]:366-511 */
    /* the sibling program src/libfm/gibbs_sbpmf2.cpp ("[S]", the paper's Algorithm 1 hyper step): Normal-Gamma prior on the
       per-dimension (mu, sigma) ([S]:375-414), tau ~ Gamma(a0 + N/2, b0 + sum e^2 / 2) ([S]:339-342), and -- as in [S], where
       that code is commented out -- NO biases and NO global mean (they stay 0).  NG_S reproduces [S] including the slip at
       [S]:412 (the mean of mu_v is scaled by the USER posterior variance); NG uses the item one. */
    SBMF_HYPER_NG_S = 1,
    SBMF_HYPER_NG = 2
} sbmf_hyper_mode;

/* Prior constants of [T]:284-313, indexed like the reference's names: alpha_0 (sigma_b_0), alpha_1
   (sigma_v), alpha_2 (sigma_u), alpha_3 (unused), alpha_4 (sigma_b_i), alpha_5 (sigma_b_j); same for
   beta_*, mu_*, sigma_*; *_dash = alpha_0_dash / beta_0_dash of the noise precision.  All 1/1/0/1 in [T]. */
typedef struct sbmf_priors {
    double alpha[6], beta[6], mu[6], sigma[6];
    double alpha_dash, beta_dash;
    /* Normal-Gamma modes, [S]:260-269: a_0, b_0 (noise precision), alpha_0, beta_0, mu_0, nu_0 (factor hypers); all 1/1/1/1/0/1 */
    double ng_a_0, ng_b_0, ng_alpha_0, ng_beta_0, ng_mu_0, ng_nu_0;
} sbmf_priors;

typedef struct sbmf_config {
    uint32_t struct_size;        /* = sizeof(sbmf_config); set by sbmf_cuda_config_default */
    uint32_t K;                  /* latent dimension; [T]:224 `uint D = 20` */
    int32_t device;              /* CUDA device ordinal */
    int32_t sample_mode;         /* sbmf_sample_mode */
    int32_t hyper_mode;          /* sbmf_hyper_mode */
    uint32_t rebuild_every;      /* rebuild the residual from scratch every n-th sweep; 1 = every sweep = [T]:342-359 */
    uint32_t burn_in;            /* sweeps before test predictions are accumulated; 0 = [T]:323 */
    uint32_t residual_mode;      /* how the per-sweep residual rebuild of [T]:342-359 is done: 0 = fused into the user phase
                                    (the fresh residual r - prediction replaces the incrementally updated one at the end of the
                                    user phase; this sweep's sum e / sum e^2 come from the incremental residual of the previous
                                    sweep, equal up to fp32 rounding); 1 = stand-alone rebuild kernel at the start of every sweep */
    uint64_t seed;               /* Philox key.  ([T] never seeds rand(); libFM parses -seed and ignores it) */
    double init_stdev;           /* 0.1 = [T]:242, 248 */
    double clamp_lo, clamp_hi;   /* 0.5 / 5.0 = [T]:627-628 */
    sbmf_priors priors;
    /* multi-GPU: one handle per rank, users (CSR) and items (CSC) sharded by nnz-balanced contiguous ranges */
    int32_t rank, world_size;    /* 0 / 1 for a single GPU */
    uint8_t nccl_id[128];        /* ncclUniqueId from sbmf_cuda_nccl_unique_id on rank 0, broadcast by the host */
} sbmf_config;

/* Filled by sbmf_cuda_get_state; NULL members are skipped.  Layouts are the reference's own:
   U row-major per user ([T]:229-232), V dimension-major ([T]:234-237), E in rating (file) order. */
typedef struct sbmf_state {
    float* U;                    /* [num_users][K] */
    float* V;                    /* [K][num_items] */
    float *b_i, *b_j;            /* [num_users], [num_items] */
    float *mu_b_i, *sigma_b_i;   /* [num_users] */
    float *mu_b_j, *sigma_b_j;   /* [num_items] */
    double *sigma_u, *mu_u, *sigma_v, *mu_v;   /* [K] */
    float* E;                    /* [n_train], rating order; as left by the last sweep (multi-GPU: only this rank's shard is written) */
    double b_0, alpha, mu_b_0, sigma_b_0;      /* out */
    double sum_e, sum_e2;        /* out: the statistics of [T]:357-358 taken at the start of the last sweep */
    uint32_t sweeps_done;        /* out */
    uint32_t reserved0;
} sbmf_state;

typedef struct sbmf_timing {
    /* CUDA-event milliseconds accumulated since the last sbmf_cuda_reset_timing, per phase of [T]'s sweep */
    double ms_rebuild;           /* [T]:342-359: stand-alone rebuild, or (fused mode) residual CSC->CSR permute + sum e / sum e^2 pass */
    double ms_hypers;            /* [T]:366-511 */
    double ms_user_phase;        /* [T]:514-558 (+ the fused residual rebuild) */
    double ms_exchange;          /* residual CSR->CSC permutation (+ multi-GPU exchange) */
    double ms_item_phase;        /* [T]:563-606 */
    double ms_eval;              /* [T]:610-636 */
    double ms_total;
    uint64_t sweeps;
    uint64_t kernel_launches;    /* kernels of this library launched by those sweeps */
    uint64_t nnz_light_user, nnz_heavy_user, nnz_light_item, nnz_heavy_item;   /* ratings by kernel path: rows held in
                                    registers (<= 2048 ratings) vs rows streamed in slices; of this rank's shard */
    /* the dominant kernel, timed per launch with CUDA events on its own stream (only while detail timing is on):
       the item-phase streaming block step heavy_accumulate_kernel<2,2> (kernels.cu) */
    double ms_top_kernel;        /* sum of its launch durations */
    uint64_t top_kernel_launches;
    uint64_t top_kernel_ratings; /* ratings one launch streams (each for 8 latent dimensions) */
    double ms_allgather;         /* multi-GPU: grouped V all-gather + reverse residual all-to-all after the item phase (the U all-gather rides in ms_exchange) */
} sbmf_timing;

/* ---- life cycle ------------------------------------------------------------------------------------ */
int sbmf_cuda_abi_version(void);
/* Defaults = the constants hard-coded in [T]: K=20 ([T]:224), priors 1/1/0/1 ([T]:284-313), clamp [0.5,5]
   ([T]:627-628), init_stdev 0.1, burn_in 0, sample_mode REF_VAR_AS_STDEV, rebuild_every 1, seed 1. */
int sbmf_cuda_config_default(sbmf_config* cfg);
int sbmf_cuda_create(const sbmf_config* cfg, sbmf_handle** out);
int sbmf_cuda_destroy(sbmf_handle* h);
const char* sbmf_cuda_last_error(const sbmf_handle* h);
/* rank 0 of a multi-GPU job calls this and ships the 128 bytes to the other ranks (torch.distributed, MPI, ...) */
int sbmf_cuda_nccl_unique_id(uint8_t out[128]);

/* Tuning / developer options of a handle, by name (they replace the process-wide environment knobs of ABI version 1; the
   only environment variable the library reads is SBMF_OPTIONS="name=value,...", applied through this call inside
   sbmf_cuda_create).  None of them changes the chain: results are identical up to fp32 summation order for every setting.
     l2_budget_mb  (192)  resident rows: MB of gathered factor blocks one launch keeps in flight; small values split a phase
                          into several launches that continue from the stored residual / partial predictions
     max_blocks_per_launch (0)  additional cap on the factor blocks per launch of the resident rows (0: none)
     resident_max  (2048) rows with more ratings stream through the sliced pipeline          [before set_train]
     resident_max_user / resident_max_item  the same threshold per side (resident_max sets both)  [before set_train]
     slice_len     (0)    ratings per slice of a streamed row; 0 = chosen from the shard size  [before set_train]
     relabel       (1)    rows are stored in order of decreasing rating count and the slots of a row in that order of the
                          opposite side (sbmf_cuda_get_storage_layout): the popular rows become neighbours, so adjacent lanes of a
                          factor gather share 128-byte lines.  0: the caller's ids and file order.  Same chain either way: every
                          draw is keyed by the caller's row id; only fp32 summation order differs          [before set_train]
     group_rows    (1)    short rows share a warp (0: one warp per row)
     row_kernels   (3)    resident rows: 3 = csrc/rows2.cuh with the transposed shuffle reduction (one block barrier per factor block,
                          64-bit pair operands); 1 = csrc/kernels.cu (the round-1 kernels: 1 % slower on the final build); 2 = rows2.cuh
                          with the Gram sums reduced through shared memory (fewer instructions, more L1TEX wavefronts: slower)
     alt_bins      (1)    resident rows of 193..512 ratings are owned by 2 warps with 6 | 8 ratings per lane (one-barrier kernels of
                          csrc/rows2.cuh) instead of 4 warps with 3 | 4: half the per-block reduction / solve work per rating
     pair_gather   (0)    streamed rows gather (previous, current) factor block as one 64-byte row by lane pairs from a per-phase
                          pair array instead of two sector gathers (measured slower on sorted rating rows)   [before set_train]
     fuse_solve    (0)    streamed rows: the row updates run in the tail of each pass, by the last slice CTA of the row, instead of
                          a launch of their own (halves the launch count of a phase; measured slower on one GPU)
     heavy_chains  (2)    streamed rows as two independent pass -> update -> pass chains on two streams (the one-CTA-per-row updates of
                          one half run under the passes of the other); 1: a single chain
     fold_user / fold_item (1 / 1)  one GPU: residual hand-over between the slot orders folded into the phase's first touch
     graph         (1)    steady-state sweep replayed from a CUDA graph when per-phase timing is off
     device_plan   (1)    multi-GPU: exchange plan computed on the device (0: host planner)  [before set_train]
     mgpu_pool     (1)    multi-GPU: rating-sized arrays from the stream-ordered pool        [before set_train]
     fuse_exchange (0)    multi-GPU with peer pushes: the user phase writes its final residual straight into the receive buffer of
                          the rank owning the rating's item and the item phase reads it in place (0: push kernel + barrier +
                          unpack kernel between the phases; measured faster: peer stores from inside the phase kernels cost
                          more than the two kernels they replace)                            [before set_train]
     peer          (1)    multi-GPU: peer-mapped replicas + direct NVLink pushes (0: NCCL)   [before set_train]
     trace         (0)    stage times of set_train on stderr
   Unknown name or value out of range: SBMF_ERR_INVALID; a [before set_train] option after set_train: SBMF_ERR_STATE. */
int sbmf_cuda_set_option(sbmf_handle* h, const char* name, int64_t value);
int sbmf_cuda_get_option(sbmf_handle* h, const char* name, int64_t* value);

/* ---- multi-GPU (one handle per rank = per GPU) --------------------------------------------
   Every rank passes the SAME full COO to set_train / set_test; the library keeps only the rank's shards: a contiguous
   user range with its CSR slots and a contiguous item range with its CSC slots (both cut to balance ratings), full
   replicas of the factors.  Per sweep: all-reduce of (sum e, sum e^2), one residual all-to-all between the phases,
   all-gather of the rows each rank updated after each phase, all-reduce of the squared test errors.  Draws are keyed
   by global row id, so the chain does not depend on world_size.
   The two planning functions are pure host code (no GPU needed) and are what the CPU tests exercise:
   plan_shards: bounds[0..world], rank r owns rows [bounds[r], bounds[r+1]).
   plan_exchange: see csrc/plan.cpp; perm[csc slot] = csr slot, csr_bounds / csc_bounds = [world+1] global slot offsets. */
int sbmf_cuda_plan_shards(const int64_t* ptr, uint32_t n_rows, int world, uint32_t* bounds);
int sbmf_cuda_plan_exchange(uint64_t n, const uint32_t* perm, int world, int rank, const int64_t* csr_bounds,
                            const int64_t* csc_bounds, uint32_t* send_idx, int64_t* send_counts, uint32_t* recv_pos,
                            int64_t* recv_counts);
/* The same plan computed on the device (csrc/storage.cu: pair-count histogram, stable compaction, stable sort by source
   rank): what set_train uses (option device_plan, default on) instead of downloading perm.  This entry takes HOST arrays and runs on
   one GPU, so a single-GPU box can check it against sbmf_cuda_plan_exchange.  pair_counts[src * world + dst] = residuals
   whose user lives on rank src and whose item lives on rank dst (send_counts of rank r = row r, recv_counts = column r). */
int sbmf_cuda_plan_exchange_device(uint64_t n, const uint32_t* perm, int world, int rank, const int64_t* csr_bounds,
                                   const int64_t* csc_bounds, uint32_t* send_idx, uint32_t* recv_pos, int64_t* pair_counts,
                                   int device);

/* ---- rating storage: replaces the jagged R / R_t build of [T]:32-221 --------------------------------- */
/* COO in FILE ORDER (rating index n = position), 0-based ids, num_users = 1 + max user id over train U test
   ([T]:151-153).  Builds on device: CSR by user + CSC by item, both STABLE w.r.t. file order ([T]:209-214),
   the CSC-slot -> CSR-slot permutation that replaces [T]'s `.id` back-pointers, the residual arrays and the
   row work lists.  Host pointers may be pageable or pinned. */
int sbmf_cuda_set_train(sbmf_handle* h, uint64_t n, const uint32_t* user, const uint32_t* item, const float* rating,
                        uint32_t num_users, uint32_t num_items);
/* test pairs of [T]:98-149; resets the running prediction sums ([T]:145). */
int sbmf_cuda_set_test(sbmf_handle* h, uint64_t nt, const uint32_t* user, const uint32_t* item, const float* rating);
/* Integer layout for the bit-exact tests; any pointer may be NULL.  row_ptr[num_users+1], col[n] (item per
   CSR slot), csr_id[n] (rating index per CSR slot), col_ptr[num_items+1], row[n] (user per CSC slot),
   csc_id[n] (rating index per CSC slot), perm[n] (CSR slot of each CSC slot). */
int sbmf_cuda_get_layout(sbmf_handle* h, int64_t* row_ptr, uint32_t* col, uint64_t* csr_id,
                         int64_t* col_ptr, uint32_t* row, uint64_t* csc_id, uint64_t* perm);
/* The arrays the sweep kernels actually run on.  With option relabel = 0 that is get_layout.  With relabel = 1 (default) every
   row lives at a POSITION -- rows ordered by decreasing number of training ratings, ties by id (G GPUs: ranks dealt round-robin
   into G contiguous ranges) -- and the slots of a row are ordered by the opposite side's position, then file order:
   user_pos[u] / item_pos[j] = position of the caller's row id; row_ptr / col_ptr are indexed by position, col / row hold
   positions, csr_id / csc_id the caller's rating indices, perm the CSR slot of each CSC slot.  The same sampler runs on it
   (draws are keyed by the caller's ids); the order only decides which factor sectors adjacent lanes of a gather fetch. */
int sbmf_cuda_get_storage_layout(sbmf_handle* h, int64_t* row_ptr, uint32_t* col, uint64_t* csr_id,
                                 int64_t* col_ptr, uint32_t* row, uint64_t* csc_id, uint64_t* perm);
int sbmf_cuda_get_row_positions(sbmf_handle* h, uint32_t* user_pos, uint32_t* item_pos);

/* libFM's transposed binary design matrix (".xt", fmatrix.h:34-52) of the training data from the layout get_layout returns:
   feature f < num_users is the user's CSR row, feature item_offset + j the item's CSC row, entries {rating index, 1.0f} in
   rating order -- byte-identical to what the reference's src/libfm/tools/transpose.cpp (83-166) writes for the same .x
   (num_features = that file's num_cols = 1 + the largest feature id in it).  Pure host code (csrc/xt_writer.cpp). */
int sbmf_cuda_write_libfm_xt(const char* path, uint32_t num_features, uint32_t num_users, uint32_t num_items,
                             uint32_t item_offset, uint64_t n, const int64_t* row_ptr, const uint64_t* csr_id,
                             const int64_t* col_ptr, const uint64_t* csc_id);
const char* sbmf_cuda_write_libfm_xt_last_error(void);

/* ---- model state ----------------------------------------------------------------------------------- */
/* U0 [num_users][K] row-major, V0 [K][num_items]; NULL => 0.1*N(0,1) from the Philox INIT streams
   ([T]:239-250).  Zeroes biases/hypers ([T]:268-281, 315-318), the sweep counter and the prediction sums. */
int sbmf_cuda_init_factors(sbmf_handle* h, const float* U0, const float* V0);
int sbmf_cuda_get_state(sbmf_handle* h, sbmf_state* out);

/* ---- the hot path: n bodies of the loop [T]:335-637 ------------------------------------------------ */
/* Each sweep: residual rebuild + statistics ([T]:342-359), alpha / global-mean chain ([T]:366-410),
   per-dimension and bias hyper-parameters ([T]:415-511), user phase ([T]:514-558), item phase ([T]:563-606),
   test prediction + running-mean RMSE ([T]:610-636).  Asynchronous w.r.t. the host until a getter is called. */
int sbmf_cuda_sweep(sbmf_handle* h, uint32_t n_sweeps);
/* RMSE of the running posterior-mean prediction (the "rmse is" value of [T]:635) and of the last sweep's own
   prediction; synchronises and copies 2 doubles back. */
int sbmf_cuda_eval(sbmf_handle* h, double* rmse_running_mean, double* rmse_last_sweep);
/* RMSE trajectory: entries [first, first+count) of the per-sweep history since init_factors. */
int sbmf_cuda_get_rmse_history(sbmf_handle* h, uint32_t first, uint32_t count, double* rmse_running_mean,
                               double* rmse_sweep);
/* Posterior-mean (clamped) prediction per test pair = sum/(sweeps - burn_in), libFM's -out content. */
int sbmf_cuda_get_pred(sbmf_handle* h, float* pred);

/* ---- checkpoint / resume ---------------------------------------------------------------------------
   [T] has none (DVector/DMatrix::save/load exist in matrix.h:130-207, 268-328 but [T] never calls them).
   The complete sampler state at a sweep boundary is what sbmf_cuda_get_state returns plus the running prediction sums
   of [T]:145, 629; every draw is a pure function of (seed, site, row, k, sweep), so a chain restored with set_state +
   set_pred_sum continues exactly as the uninterrupted one would have. */
/* After set_train (+ set_test): replaces init_factors.  in->U and in->V are required; NULL bias / hyper arrays mean zeros
   (the state [T]:268-281 starts from).  in->sweeps_done becomes the sweep counter (and the Philox counter).  in->E
   (rating order, as written by get_state) restores the residual; NULL => the next sweep rebuilds it stand-alone like
   sweep 0 ([T]:342-359).  Multi-GPU: every rank passes the same factors / biases / hypers and its own E.  The RMSE
   history of the sweeps before the restore point reads 0.  Also zeroes the prediction sums: call set_pred_sum AFTER it. */
int sbmf_cuda_set_state(sbmf_handle* h, const sbmf_state* in);
/* running sums of the clamped test predictions over the collected sweeps, [n_test] doubles ([T]:629).  Multi-GPU: get
   is collective (every rank holds the sums of its slice of the test set), set takes the full array on every rank. */
int sbmf_cuda_get_pred_sum(sbmf_handle* h, double* sum);
int sbmf_cuda_set_pred_sum(sbmf_handle* h, const double* sum);

/* Checkpoint files (pure host code, csrc/checkpoint.cpp): a 128-byte little-endian header (magic "SBMFCKP2", the
   dimensions and chain flags below) followed by the arrays of sbmf_state in declaration order, then pred_sum; absent arrays
   have their bit in `present` cleared.  write: NULL members of st / NULL pred_sum are recorded as absent.  read_dims: header
   only.  read: `expect` holds the dimensions (num_users, num_items, K, n_train, n_test) st's non-NULL members and pred_sum
   were allocated for; a file with other dimensions is refused with SBMF_ERR_INVALID before anything is copied.  Members
   absent from the file are reported by clearing the pointer in st (and *pred_sum_present = 0).
   seed / sample_mode / burn_in / residual_mode / rebuild_every: the sbmf_config values of the chain that wrote the file -- a
   resumed chain only continues the interrupted one under the same values (the CLI's -load_state checks them). */
typedef struct sbmf_checkpoint_dims {
    uint32_t num_users, num_items, K;
    int32_t hyper_mode;
    uint64_t n_train, n_test;
    uint32_t sweeps_done;
    uint32_t present;            /* bit i = i-th pointer member of sbmf_state (U = bit 0 ... E = bit 12), bit 13 = pred_sum */
    uint64_t seed;
    int32_t sample_mode;
    uint32_t burn_in, residual_mode, rebuild_every;
} sbmf_checkpoint_dims;
int sbmf_cuda_checkpoint_write(const char* path, const sbmf_checkpoint_dims* dims, const sbmf_state* st, const double* pred_sum);
int sbmf_cuda_checkpoint_read_dims(const char* path, sbmf_checkpoint_dims* dims);
int sbmf_cuda_checkpoint_read(const char* path, const sbmf_checkpoint_dims* expect, sbmf_state* st, double* pred_sum,
                              int* pred_sum_present);
const char* sbmf_cuda_checkpoint_last_error(void);

/* ---- instrumentation -------------------------------------------------------------------------------- */
int sbmf_cuda_get_timing(sbmf_handle* h, sbmf_timing* out);
int sbmf_cuda_reset_timing(sbmf_handle* h);
int sbmf_cuda_synchronize(sbmf_handle* h);
/* Per-phase timing makes every sweep wait for its own CUDA events (one host/device round trip per sweep).
   enabled = 0 turns that off: sbmf_cuda_sweep(h, n) then only enqueues work.  Default: enabled. */
int sbmf_cuda_set_timing_enabled(sbmf_handle* h, int enabled);
/* enabled = 2 additionally brackets every launch of the dominant kernel with events (sbmf_timing.ms_top_kernel). */
/* Device time of the last sbmf_cuda_sweep call (all its sweeps), from two CUDA events on the library's main
   stream; waits for that call to finish. */
int sbmf_cuda_last_sweep_call_ms(sbmf_handle* h, double* ms);
/* Pinned host buffers for callers that want full-speed host<->device copies in set_train / set_test / get_*. */
int sbmf_cuda_host_alloc(void** ptr, size_t bytes);
int sbmf_cuda_host_free(void* ptr);

/* ---- hardware probes: the denominators of bench.py's roofline, measured on the device the job runs on, in the same run ----
   hbm_*: streaming kernels over 1 GiB buffers.  gather_sectors_per_s[form]: random row gathers from an L2-resident table of
   table_bytes (0 = 15 MB = one K8 factor block of the Netflix-shaped user side), row ids streamed from HBM, n_gathers sectors
   per launch (0 = 64 Mi); best of 5 launches each.  Forms: G32_LANE = one random 32-byte sector per lane with one LDG.E.256 --
   the access form of the factor gathers in csrc/kernels.cu; G64_* / G128_* = the same bytes as contiguous 64-byte rows / full
   128-byte lines, fetched per lane (2 / 4 loads) or by 2 / 4 cooperating adjacent lanes -- what a wider factor layout could reach. */
enum { SBMF_PROBE_G32_LANE = 0, SBMF_PROBE_G64_LANE = 1, SBMF_PROBE_G64_COOP2 = 2, SBMF_PROBE_G128_LANE = 3, SBMF_PROBE_G128_COOP4 = 4,
       SBMF_PROBE_FORMS = 5 };
typedef struct sbmf_probe_result {
    double hbm_copy_gbs;                            /* (bytes read + bytes written) / s */
    double hbm_read_gbs;
    double gather_sectors_per_s[8];                 /* [SBMF_PROBE_FORMS] 32-byte sectors delivered to the SMs per second */
    double sm_clock_mhz_max;                        /* cudaDevAttrClockRate */
    uint64_t table_bytes, n_gathers;
    int32_t sm_count, reserved0;
} sbmf_probe_result;
int sbmf_cuda_probe(int device, uint64_t table_bytes, uint64_t n_gathers, sbmf_probe_result* out);
const char* sbmf_cuda_probe_last_error(void);

/* ---- synthetic workloads (bench / tests): MovieLens/Netflix-shaped Zipf rating matrices, generated on
   the device.  Distinct (user,item) pairs with Zipf(s_user) x Zipf(s_item) marginals over
   randomly permuted ids, ratings from a planted rank-16 model rounded to 0.5 steps in [0.5,5], sorted by
   (user,item) like every shipped fixture, split train/test by Bernoulli(test_frac).  Two-call protocol:
   call with NULL arrays to learn n_train/n_test, then with buffers of that size. */
typedef struct sbmf_synth_spec {
    uint32_t num_users, num_items;
    uint64_t n_ratings;          /* train + test */
    double s_user, s_item;       /* Zipf exponents (0.8 / 1.0 in SURVEY.md 8d) */
    double test_frac;            /* 0.1 */
    uint64_t seed;
    int32_t device;
    int32_t reserved0;
} sbmf_synth_spec;
int sbmf_cuda_synth_generate(const sbmf_synth_spec* spec, uint64_t* n_train, uint64_t* n_test,
                             uint32_t* train_user, uint32_t* train_item, float* train_rating,
                             uint32_t* test_user, uint32_t* test_item, float* test_rating);
const char* sbmf_cuda_synth_last_error(void);

/* The same matrix family sampled sparsely on host threads (csrc/synth_host.cpp): O(ratings) work instead of one trial per
   (user, item) pair, for shapes like BASELINE.json's 10M x 1M / 1B ratings whose pair grid (1e13) cannot be visited.
   spec->device is ignored; threads <= 0: all host threads.  The six output arrays are allocated by the call
   (release each with sbmf_cuda_synth_host_free); the matrix does not depend on the number of threads. */
int sbmf_cuda_synth_host_generate(const sbmf_synth_spec* spec, int threads, uint64_t* n_train, uint64_t* n_test,
                             uint32_t** train_user, uint32_t** train_item, float** train_rating,
                             uint32_t** test_user, uint32_t** test_item, float** test_rating);
void sbmf_cuda_synth_host_free(void* ptr);
const char* sbmf_cuda_synth_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SBMF_CUDA_H_ */
