// xt_writer.cpp -- libFM's transposed binary design matrix (".xt") from the device-built rating layout.  Pure host code.
//
// For matrix-factorisation data the design matrix X has one row per rating with two indicator features (user u -> feature u,
// item j -> feature item_offset + j), so X^T has one row per FEATURE listing the ratings that carry it, in rating (file)
// order -- which is exactly the CSR row of a user (rating index per slot = csr_id) and the CSC row of an item (csc_id): the
// storage build of storage.cu IS the transpose.  The reference produces the same file with src/libfm/tools/transpose.cpp
// (transpose.cpp:83-166: counts per column, then stable fill in input-row order); format of fmatrix.h:34-52:
//   file_header{u32 id = 2; u32 float_size = 4; u64 num_values; u32 num_rows = features; u32 num_cols = ratings}
//   per row: u32 size, then size x {u32 id = rating index; f32 value = 1}
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/sbmf_cuda.h"

namespace {
thread_local std::string g_err;
}

extern "C" {

const char* sbmf_cuda_write_libfm_xt_last_error(void) { return g_err.c_str(); }

int sbmf_cuda_write_libfm_xt(const char* path, uint32_t num_features, uint32_t num_users, uint32_t num_items, uint32_t item_offset,
                             uint64_t n, const int64_t* row_ptr, const uint64_t* csr_id, const int64_t* col_ptr, const uint64_t* csc_id)
{
    if (!path || !row_ptr || !col_ptr || (n && (!csr_id || !csc_id))) {
        g_err = "write_libfm_xt: null argument";
        return SBMF_ERR_INVALID;
    }
    if (item_offset < num_users || (uint64_t)row_ptr[num_users] != n || (uint64_t)col_ptr[num_items] != n || n >= (1ull << 32)) {
        g_err = "write_libfm_xt: item features must be numbered after the users, and the layout must hold n ratings";
        return SBMF_ERR_INVALID;
    }
    // every rated item must have a feature row: rows beyond num_features would be lost
    for (uint32_t j = num_items; j-- > 0;)
        if (col_ptr[j + 1] > col_ptr[j]) {
            if ((uint64_t)item_offset + j >= num_features) {
                g_err = "write_libfm_xt: num_features does not cover item " + std::to_string(j);
                return SBMF_ERR_INVALID;
            }
            break;
        }
    FILE* f = fopen(path, "wb");
    if (!f) {
        g_err = std::string("write_libfm_xt: unable to open ") + path;
        return SBMF_ERR_INVALID;
    }
    struct {
        uint32_t id, float_size;
        uint64_t num_values;
        uint32_t num_rows, num_cols;
    } fh = {2u, 4u, 2 * n, num_features, (uint32_t)n};
    static_assert(sizeof(fh) == 24, "file_header layout (fmatrix.h:46-52)");
    bool ok = fwrite(&fh, sizeof(fh), 1, f) == 1;
    struct Entry {
        uint32_t id;
        float value;
    };
    std::vector<Entry> row;
    auto put = [&](const int64_t* ptr, const uint64_t* id, uint32_t r) {
        const uint32_t size = (uint32_t)(ptr[r + 1] - ptr[r]);
        row.resize(size);
        for (uint32_t p = 0; p < size; ++p) row[p] = Entry{(uint32_t)id[ptr[r] + p], 1.0f};
        ok = ok && fwrite(&size, 4, 1, f) == 1 && (size == 0 || fwrite(row.data(), sizeof(Entry), size, f) == size);
    };
    const uint32_t zero = 0;
    for (uint32_t ft = 0; ft < num_features && ok; ++ft) {
        if (ft < num_users) put(row_ptr, csr_id, ft);
        else if (ft >= item_offset && ft - item_offset < num_items) put(col_ptr, csc_id, ft - item_offset);
        else ok = fwrite(&zero, 4, 1, f) == 1;   // a feature id nobody uses (gap between users and items)
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) {
        g_err = std::string("write_libfm_xt: write to ") + path + " failed";
        return SBMF_ERR_INVALID;
    }
    return SBMF_OK;
}

}  // extern "C"
