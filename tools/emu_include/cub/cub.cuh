// tools/emu_include/cub/cub.cuh -- TEST INFRASTRUCTURE: host stand-in for the one CUB entry point csrc/fm.cu uses (see cuda_runtime.h here)
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

#include <cuda_runtime.h>

namespace cub {
struct DeviceRadixSort {
    // stable sort of (key, value) pairs by the key bits [begin_bit, end_bit)
    template <class K, class V>
    static cudaError_t SortPairs(void* tmp, size_t& tmp_bytes, const K* kin, K* kout, const V* vin, V* vout, int n, int begin_bit, int end_bit, cudaStream_t)
    {
        if (!tmp) {
            tmp_bytes = 1;
            return cudaSuccess;
        }
        const K mask = (end_bit - begin_bit >= (int)(8 * sizeof(K))) ? ~(K)0 : (K)((((K)1 << (end_bit - begin_bit)) - 1) << begin_bit);
        std::vector<int> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return (kin[a] & mask) < (kin[b] & mask); });
        for (int i = 0; i < n; ++i) {
            kout[i] = kin[idx[i]];
            vout[i] = vin[idx[i]];
        }
        return cudaSuccess;
    }
};
struct DeviceSelect {
    // stable compaction of the items the predicate accepts
    template <class T, class Pred>
    static cudaError_t If(void* tmp, size_t& tmp_bytes, const T* in, T* out, int* num_selected, int n, Pred pred, cudaStream_t)
    {
        if (!tmp) {
            tmp_bytes = 1;
            return cudaSuccess;
        }
        int k = 0;
        for (int i = 0; i < n; ++i)
            if (pred(in[i])) out[k++] = in[i];
        *num_selected = k;
        return cudaSuccess;
    }
};
}  // namespace cub
