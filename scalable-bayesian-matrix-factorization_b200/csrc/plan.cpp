// plan.cpp -- host-side sharding logic of the multi-GPU sweep (no CUDA in this file; unit-tested on CPU).
//
// in the user phase rows are independent given V, in the item phase columns are independent given U,
// so rank r owns a contiguous user range (its CSR shard) and a contiguous item range (its CSC shard), both cut to
// balance the number of RATINGS, not rows.  Between the phases every residual value moves from the rank that owns
// its user to the rank that owns its item: one all-to-all per half-sweep, planned here once from the permutation
// perm[csc slot] = csr slot that the single-GPU path uses as a plain gather.
#include <stdint.h>

#include <vector>

#include "../../include/sbmf_cuda.h"

extern "C" {

// bounds[0..world]: rank r owns rows [bounds[r], bounds[r+1]).  Greedy cut at the row whose prefix count is closest to
// r * nnz / world; rows are never split; empty shards are legal (world > number of non-empty rows).
int sbmf_cuda_plan_shards(const int64_t* ptr, uint32_t n_rows, int world, uint32_t* bounds)
{
    if (!ptr || !bounds || world < 1) return SBMF_ERR_INVALID;
    const int64_t nnz = ptr[n_rows] - ptr[0];
    bounds[0] = 0;
    uint32_t row = 0;
    for (int r = 1; r < world; ++r) {
        const int64_t target = ptr[0] + (int64_t)(((__int128)nnz * r) / world);
        // first row whose start is >= target, then the closer of it and its predecessor
        uint32_t lo = row, hi = n_rows;
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (ptr[mid] < target) lo = mid + 1;
            else hi = mid;
        }
        if (lo > row && (target - ptr[lo - 1]) <= (ptr[lo] - target)) --lo;
        if (lo < row) lo = row;
        row = lo;
        bounds[r] = row;
    }
    bounds[world] = n_rows;
    return SBMF_OK;
}

// Exchange plan of rank `rank`.  csr_bounds / csc_bounds: [world+1] global SLOT offsets of the shards.
//   send_idx[i], i in [0, n_csr_local): LOCAL csr slot to place at position i of the send buffer; the buffer is grouped by
//       destination rank (send_counts[q] values for rank q, in ascending order of q's csc slots).
//   recv_pos[t], t in [0, n_csc_local): position in the receive buffer (grouped by source rank, recv_counts[q] values from
//       rank q in the order q sends them) of the residual of LOCAL csc slot t.
int sbmf_cuda_plan_exchange(uint64_t n, const uint32_t* perm, int world, int rank, const int64_t* csr_bounds, const int64_t* csc_bounds,
                            uint32_t* send_idx, int64_t* send_counts, uint32_t* recv_pos, int64_t* recv_counts)
{
    if (!perm || !csr_bounds || !csc_bounds || !send_counts || !recv_counts || world < 1 || rank < 0 || rank >= world) return SBMF_ERR_INVALID;
    if ((uint64_t)csr_bounds[world] != n || (uint64_t)csc_bounds[world] != n) return SBMF_ERR_INVALID;
    const int64_t my_csr0 = csr_bounds[rank], my_csr1 = csr_bounds[rank + 1];
    const int64_t my_csc0 = csc_bounds[rank], my_csc1 = csc_bounds[rank + 1];
    auto owner = [&](uint32_t csr_slot) {
        int lo = 0, hi = world - 1;   // last q with csr_bounds[q] <= slot
        while (lo < hi) {
            const int mid = (lo + hi + 1) / 2;
            if (csr_bounds[mid] <= (int64_t)csr_slot) lo = mid;
            else hi = mid - 1;
        }
        return lo;
    };
    // what I send: for each destination q, my csr slots that q's csc shard references, in q's csc order
    int64_t pos = 0;
    for (int q = 0; q < world; ++q) {
        int64_t cnt = 0;
        for (int64_t t = csc_bounds[q]; t < csc_bounds[q + 1]; ++t) {
            const int64_t s = perm[t];
            if (s >= my_csr0 && s < my_csr1) {
                if (send_idx) send_idx[pos + cnt] = (uint32_t)(s - my_csr0);
                ++cnt;
            }
        }
        send_counts[q] = cnt;
        pos += cnt;
    }
    if (pos != my_csr1 - my_csr0) return SBMF_ERR_INVALID;   // perm is not a permutation
    // what I receive: my csc slots grouped by owning rank of their csr slot, ascending csc slot inside a group
    std::vector<int64_t> base(world + 1, 0);
    for (int q = 0; q < world; ++q) recv_counts[q] = 0;
    for (int64_t t = my_csc0; t < my_csc1; ++t) recv_counts[owner(perm[t])]++;
    for (int q = 0; q < world; ++q) base[q + 1] = base[q] + recv_counts[q];
    if (recv_pos) {
        std::vector<int64_t> fill(base.begin(), base.end() - 1);
        for (int64_t t = my_csc0; t < my_csc1; ++t) recv_pos[t - my_csc0] = (uint32_t)(fill[owner(perm[t])]++);
    }
    return SBMF_OK;
}

// counts[src * world + dst] = number of residuals whose user lives on rank src (CSR owner) and whose item lives on rank dst
// (CSC owner): the full traffic matrix of the residual exchange, from which every rank derives where its segments start in
// its peers' receive buffers (direct NVLink pushes need no handshake).
int sbmf_cuda_plan_pair_counts(uint64_t n, const uint32_t* perm, int world, const int64_t* csr_bounds, const int64_t* csc_bounds, int64_t* counts)
{
    if (!perm || !csr_bounds || !csc_bounds || !counts || world < 1) return SBMF_ERR_INVALID;
    if ((uint64_t)csr_bounds[world] != n || (uint64_t)csc_bounds[world] != n) return SBMF_ERR_INVALID;
    for (int i = 0; i < world * world; ++i) counts[i] = 0;
    for (int dst = 0; dst < world; ++dst)
        for (int64_t t = csc_bounds[dst]; t < csc_bounds[dst + 1]; ++t) {
            const int64_t s = perm[t];
            int lo = 0, hi = world - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) / 2;
                if (csr_bounds[mid] <= s) lo = mid;
                else hi = mid - 1;
            }
            counts[lo * world + dst]++;
        }
    return SBMF_OK;
}

}  // extern "C"
