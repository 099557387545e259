// checkpoint.cpp -- checkpoint files of the sampler state (sbmf_cuda_checkpoint_* of include/sbmf_cuda.h).  Pure host code.
// The reference has no checkpointing on this path; the file simply carries what sbmf_cuda_get_state and
// sbmf_cuda_get_pred_sum return, so that sbmf_cuda_set_state / set_pred_sum can continue the chain in another process.
//
// Layout (little-endian): 128-byte header
//   char magic[8] = "SBMFCKP2"; u32 num_users, num_items, K; i32 hyper_mode; u64 n_train, n_test; u32 sweeps_done, present;
//   f64 b_0, alpha, mu_b_0, sigma_b_0, sum_e, sum_e2;
//   u64 seed; i32 sample_mode; u32 burn_in, residual_mode, rebuild_every; u8 reserved[8]
// (the chain a resumed handle continues is a function of the seed and the mode flags, so they travel with the state and the
// caller -- the CLI's -load_state -- can refuse a resume under different flags).  Then, for every bit set in `present`, in this order: U[I][K] f32, V[K][J] f32, b_i[I], b_j[J], mu_b_i[I], sigma_b_i[I],
// mu_b_j[J], sigma_b_j[J] f32, sigma_u[K], mu_u[K], sigma_v[K], mu_v[K] f64, E[n_train] f32, pred_sum[n_test] f64.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/sbmf_cuda.h"

namespace {

thread_local std::string g_err;
const char kMagic[8] = {'S', 'B', 'M', 'F', 'C', 'K', 'P', '2'};
constexpr int kArrays = 14;   // 13 pointer members of sbmf_state + pred_sum

struct Header {
    char magic[8];
    uint32_t num_users, num_items, K;
    int32_t hyper_mode;
    uint64_t n_train, n_test;
    uint32_t sweeps_done, present;
    double b_0, alpha, mu_b_0, sigma_b_0, sum_e, sum_e2;
    uint64_t seed;
    int32_t sample_mode;
    uint32_t burn_in, residual_mode, rebuild_every;
    uint8_t reserved[8];
};
static_assert(sizeof(Header) == 128, "checkpoint header is 128 bytes");

// byte size of array i for the dimensions in the header
uint64_t array_bytes(const Header& h, int i)
{
    const uint64_t I = h.num_users, J = h.num_items, K = h.K;
    switch (i) {
        case 0: return I * K * 4;
        case 1: return K * J * 4;
        case 2: case 4: case 5: return I * 4;
        case 3: case 6: case 7: return J * 4;
        case 8: case 9: case 10: case 11: return K * 8;
        case 12: return h.n_train * 4;
        default: return h.n_test * 8;
    }
}

const void* member(const sbmf_state* st, const double* pred_sum, int i)
{
    const void* p[kArrays] = {st->U, st->V, st->b_i, st->b_j, st->mu_b_i, st->sigma_b_i, st->mu_b_j, st->sigma_b_j,
                              st->sigma_u, st->mu_u, st->sigma_v, st->mu_v, st->E, pred_sum};
    return p[i];
}

void clear_member(sbmf_state* st, int i)
{
    switch (i) {
        case 0: st->U = nullptr; break;
        case 1: st->V = nullptr; break;
        case 2: st->b_i = nullptr; break;
        case 3: st->b_j = nullptr; break;
        case 4: st->mu_b_i = nullptr; break;
        case 5: st->sigma_b_i = nullptr; break;
        case 6: st->mu_b_j = nullptr; break;
        case 7: st->sigma_b_j = nullptr; break;
        case 8: st->sigma_u = nullptr; break;
        case 9: st->mu_u = nullptr; break;
        case 10: st->sigma_v = nullptr; break;
        case 11: st->mu_v = nullptr; break;
        case 12: st->E = nullptr; break;
        default: break;
    }
}

int read_header(FILE* f, const char* path, Header& h)
{
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, kMagic, 8) != 0) {
        g_err = std::string(path) + ": not an SBMF checkpoint (bad magic or truncated header)";
        return SBMF_ERR_INVALID;
    }
    if (h.K == 0 || h.K > SBMF_MAX_K || h.num_users == 0 || h.num_items == 0 || (h.present >> kArrays) != 0) {
        g_err = std::string(path) + ": corrupt checkpoint header";
        return SBMF_ERR_INVALID;
    }
    return SBMF_OK;
}

}  // namespace

extern "C" {

const char* sbmf_cuda_checkpoint_last_error(void) { return g_err.c_str(); }

int sbmf_cuda_checkpoint_write(const char* path, const sbmf_checkpoint_dims* dims, const sbmf_state* st, const double* pred_sum)
{
    if (!path || !dims || !st) {
        g_err = "checkpoint_write: null argument";
        return SBMF_ERR_INVALID;
    }
    if (dims->K == 0 || dims->K > SBMF_MAX_K || dims->num_users == 0 || dims->num_items == 0) {
        g_err = "checkpoint_write: bad dimensions";
        return SBMF_ERR_INVALID;
    }
    Header h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, kMagic, 8);
    h.num_users = dims->num_users;
    h.num_items = dims->num_items;
    h.K = dims->K;
    h.hyper_mode = dims->hyper_mode;
    h.n_train = dims->n_train;
    h.n_test = dims->n_test;
    h.sweeps_done = st->sweeps_done;
    h.seed = dims->seed;
    h.sample_mode = dims->sample_mode;
    h.burn_in = dims->burn_in;
    h.residual_mode = dims->residual_mode;
    h.rebuild_every = dims->rebuild_every;
    for (int i = 0; i < kArrays; ++i)
        if (member(st, pred_sum, i)) h.present |= 1u << i;
    h.b_0 = st->b_0;
    h.alpha = st->alpha;
    h.mu_b_0 = st->mu_b_0;
    h.sigma_b_0 = st->sigma_b_0;
    h.sum_e = st->sum_e;
    h.sum_e2 = st->sum_e2;
    // write next to the target and rename, so that an interrupted write never leaves a half checkpoint under the final name
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) {
        g_err = "checkpoint_write: unable to open " + tmp;
        return SBMF_ERR_INVALID;
    }
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    for (int i = 0; i < kArrays && ok; ++i) {
        const void* p = member(st, pred_sum, i);
        const uint64_t nb = array_bytes(h, i);
        if (p && nb) ok = fwrite(p, 1, nb, f) == nb;
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path) != 0) {
        remove(tmp.c_str());
        g_err = std::string("checkpoint_write: write to ") + path + " failed";
        return SBMF_ERR_INVALID;
    }
    return SBMF_OK;
}

int sbmf_cuda_checkpoint_read_dims(const char* path, sbmf_checkpoint_dims* dims)
{
    if (!path || !dims) {
        g_err = "checkpoint_read_dims: null argument";
        return SBMF_ERR_INVALID;
    }
    FILE* f = fopen(path, "rb");
    if (!f) {
        g_err = std::string("checkpoint_read_dims: unable to open ") + path;
        return SBMF_ERR_INVALID;
    }
    Header h;
    const int rc = read_header(f, path, h);
    fclose(f);
    if (rc != SBMF_OK) return rc;
    dims->num_users = h.num_users;
    dims->num_items = h.num_items;
    dims->K = h.K;
    dims->hyper_mode = h.hyper_mode;
    dims->n_train = h.n_train;
    dims->n_test = h.n_test;
    dims->sweeps_done = h.sweeps_done;
    dims->present = h.present;
    dims->seed = h.seed;
    dims->sample_mode = h.sample_mode;
    dims->burn_in = h.burn_in;
    dims->residual_mode = h.residual_mode;
    dims->rebuild_every = h.rebuild_every;
    return SBMF_OK;
}

int sbmf_cuda_checkpoint_read(const char* path, const sbmf_checkpoint_dims* expect, sbmf_state* st, double* pred_sum, int* pred_sum_present)
{
    if (!path || !st || !expect) {
        g_err = "checkpoint_read: null argument (expect = the dimensions the buffers in st were allocated for)";
        return SBMF_ERR_INVALID;
    }
    FILE* f = fopen(path, "rb");
    if (!f) {
        g_err = std::string("checkpoint_read: unable to open ") + path;
        return SBMF_ERR_INVALID;
    }
    Header h;
    int rc = read_header(f, path, h);
    if (rc != SBMF_OK) {
        fclose(f);
        return rc;
    }
    // the arrays are copied with the sizes of the FILE: refuse before touching the caller's buffers if they were sized otherwise
    if (h.num_users != expect->num_users || h.num_items != expect->num_items || h.K != expect->K || h.n_train != expect->n_train ||
        h.n_test != expect->n_test) {
        fclose(f);
        g_err = std::string(path) + ": checkpoint is for " + std::to_string(h.num_users) + " users x " + std::to_string(h.num_items) +
                " items, K=" + std::to_string(h.K) + ", " + std::to_string(h.n_train) + " train / " + std::to_string(h.n_test) +
                " test ratings; the caller's buffers are for " + std::to_string(expect->num_users) + " x " + std::to_string(expect->num_items) +
                ", K=" + std::to_string(expect->K) + ", " + std::to_string(expect->n_train) + " / " + std::to_string(expect->n_test);
        return SBMF_ERR_INVALID;
    }
    if (pred_sum_present) *pred_sum_present = 0;
    for (int i = 0; i < kArrays; ++i) {
        const uint64_t nb = array_bytes(h, i);
        void* dst = const_cast<void*>(member(st, pred_sum, i));
        if (!((h.present >> i) & 1u)) {
            clear_member(st, i);
            continue;
        }
        bool ok;
        if (dst) ok = nb == 0 || fread(dst, 1, nb, f) == nb;
        else ok = fseek(f, (long)nb, SEEK_CUR) == 0;
        if (!ok) {
            fclose(f);
            g_err = std::string(path) + ": truncated checkpoint";
            return SBMF_ERR_INVALID;
        }
        if (i == kArrays - 1 && dst && pred_sum_present) *pred_sum_present = 1;
    }
    fclose(f);
    st->b_0 = h.b_0;
    st->alpha = h.alpha;
    st->mu_b_0 = h.mu_b_0;
    st->sigma_b_0 = h.sigma_b_0;
    st->sum_e = h.sum_e;
    st->sum_e2 = h.sum_e2;
    st->sweeps_done = h.sweeps_done;
    return SBMF_OK;
}

}  // extern "C"
