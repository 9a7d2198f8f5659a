// TEST INFRASTRUCTURE ONLY (see ../cuda_runtime.h): sequential host stand-ins for the four cub device primitives the
// streaming kernels' drivers call, with cub's signatures and its two-phase temp-storage protocol, so that the native
// sequencing code (workspace carving, argument order, run/scan logic) executes in the GPU-less container.
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

#include "../cuda_runtime.h"

namespace cub {

struct DeviceSelect {
    template <class In, class Flag, class Out, class Num>
    static cudaError_t Flagged(void *temp, size_t &bytes, In in, Flag flags, Out out, Num n_out, int n, cudaStream_t = 0) {
        if (!temp) { bytes = 1; return cudaSuccess; }
        int m = 0;
        for (int i = 0; i < n; ++i)
            if (flags[i]) out[m++] = in[i];
        *n_out = m;
        return cudaSuccess;
    }
};

struct DeviceRadixSort {
    template <class K, class V>
    static cudaError_t SortPairs(void *temp, size_t &bytes, const K *kin, K *kout, const V *vin, V *vout, long long n,
                                 int begin_bit = 0, int end_bit = sizeof(K) * 8, cudaStream_t = 0) {
        if (!temp) { bytes = 1; return cudaSuccess; }
        using U = unsigned long long;
        const U mask = (end_bit - begin_bit >= 64) ? ~0ull : (((1ull << (end_bit - begin_bit)) - 1) << begin_bit);
        std::vector<long long> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](long long a, long long b) { return ((U)kin[a] & mask) < ((U)kin[b] & mask); });
        for (long long i = 0; i < n; ++i) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
        return cudaSuccess;
    }
};

struct DeviceScan {
    template <class In, class Out>
    static cudaError_t InclusiveSum(void *temp, size_t &bytes, In in, Out out, int n, cudaStream_t = 0) {
        if (!temp) { bytes = 1; return cudaSuccess; }
        if (n <= 0) return cudaSuccess;
        auto acc = in[0];
        out[0] = acc;
        for (int i = 1; i < n; ++i) { acc = acc + in[i]; out[i] = acc; }
        return cudaSuccess;
    }

    template <class In, class Out, class Op>
    static cudaError_t InclusiveScan(void *temp, size_t &bytes, In in, Out out, Op op, int n, cudaStream_t = 0) {
        if (!temp) { bytes = 1; return cudaSuccess; }
        if (n <= 0) return cudaSuccess;
        auto acc = in[0];
        out[0] = acc;
        for (int i = 1; i < n; ++i) { acc = op(acc, in[i]); out[i] = acc; }
        return cudaSuccess;
    }
};

}  // namespace cub
