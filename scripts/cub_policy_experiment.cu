// Does a smaller onesweep tile make cub's radix sort faster at THIS problem size?  The two sorts of the binning stage
// are 1.0 M (depth bits, index) pairs on 32 bits and 2.3 M (tile id, Gaussian) pairs on 12 bits; cub's sm_100 policy uses
// 384 threads x 23 items = 8 832 pairs per CTA, i.e. 114 / 259 CTAs on 148 SMs -- less than one / two waves of one CTA
// per SM, each pass bound by one CTA's serial work.  This program times DispatchRadixSort with custom policy hubs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/cub_policy_experiment scripts/cub_policy_experiment.cu
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdint>
#include <vector>

template <int THREADS, int ITEMS>
struct Hub {
    using K = uint32_t; using V = uint32_t; using O = int;
    using Base = typename cub::detail::radix::policy_hub<K, V, O>::Policy1000;
    struct Policy : cub::ChainedPolicy<350, Policy, Policy> {
        static constexpr bool ONESWEEP = true;
        static constexpr int ONESWEEP_RADIX_BITS = 8;
        using HistogramPolicy = typename Base::HistogramPolicy;
        using ExclusiveSumPolicy = typename Base::ExclusiveSumPolicy;
        using OnesweepPolicy = cub::AgentRadixSortOnesweepPolicy<THREADS, ITEMS, uint32_t, 1, cub::RADIX_RANK_MATCH_EARLY_COUNTS_ANY,
                                                                 cub::BLOCK_SCAN_RAKING_MEMOIZE, cub::RADIX_SORT_STORE_DIRECT, 8>;
        using ScanPolicy = typename Base::ScanPolicy;
        using DownsweepPolicy = typename Base::DownsweepPolicy;
        using AltDownsweepPolicy = typename Base::AltDownsweepPolicy;
        using UpsweepPolicy = typename Base::UpsweepPolicy;
        using AltUpsweepPolicy = typename Base::AltUpsweepPolicy;
        using SingleTilePolicy = typename Base::SingleTilePolicy;
        using SegmentedPolicy = typename Base::SegmentedPolicy;
        using AltSegmentedPolicy = typename Base::AltSegmentedPolicy;
    };
    using MaxPolicy = Policy;
};

template <class HubT>
cudaError_t sort_with(void *tmp, size_t &bytes, const uint32_t *ki, uint32_t *ko, const uint32_t *vi, uint32_t *vo, int n, int b0,
                      int b1, cudaStream_t st) {
    cub::DoubleBuffer<uint32_t> k(const_cast<uint32_t *>(ki), ko), v(const_cast<uint32_t *>(vi), vo);
    return cub::DispatchRadixSort<false, uint32_t, uint32_t, int, HubT>::Dispatch(tmp, bytes, k, v, n, b0, b1, false, st);
}

static cudaError_t sort_default(void *tmp, size_t &bytes, const uint32_t *ki, uint32_t *ko, const uint32_t *vi, uint32_t *vo, int n,
                                int b0, int b1, cudaStream_t st) {
    return cub::DeviceRadixSort::SortPairs(tmp, bytes, ki, ko, vi, vo, n, b0, b1, st);
}

typedef cudaError_t (*SortFn)(void *, size_t &, const uint32_t *, uint32_t *, const uint32_t *, uint32_t *, int, int, int, cudaStream_t);

int main() {
    struct Case { const char *name; int n, b0, b1; } cases[2] = {{"depth sort 1.0M x 32 bits", 1002528, 0, 32}, {"tile sort 2.28M x 12 bits", 2284225, 0, 12}};
    struct Var { const char *name; SortFn fn; } vars[] = {
        {"cub default (384x23)", sort_default}, {"384x12", sort_with<Hub<384, 12>>}, {"256x12", sort_with<Hub<256, 12>>},
        {"256x8", sort_with<Hub<256, 8>>},     {"512x8", sort_with<Hub<512, 8>>},   {"256x16", sort_with<Hub<256, 16>>},
        {"128x12", sort_with<Hub<128, 12>>},   {"512x12", sort_with<Hub<512, 12>>}};
    printf("{");
    for (int c = 0; c < 2; ++c) {
        const int n = cases[c].n;
        std::vector<uint32_t> hk(n), hv(n);
        uint64_t s = 88172645463325252ull;
        for (int i = 0; i < n; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            float d = 2.0f + 3.0f * (float)((s >> 11) & 0xFFFFFF) / 16777216.0f;      // depths in [2, 5)
            uint32_t bits; memcpy(&bits, &d, 4);
            hk[i] = (cases[c].b1 == 32) ? bits : (uint32_t)((s >> 40) % 2500);
            hv[i] = i;
        }
        uint32_t *ki, *ko, *vi, *vo, *ref;
        cudaMalloc(&ki, 4 * n); cudaMalloc(&ko, 4 * n); cudaMalloc(&vi, 4 * n); cudaMalloc(&vo, 4 * n); cudaMalloc(&ref, 4 * n);
        cudaMemcpy(ki, hk.data(), 4 * n, cudaMemcpyHostToDevice); cudaMemcpy(vi, hv.data(), 4 * n, cudaMemcpyHostToDevice);
        void *tmp; cudaMalloc(&tmp, 64 << 20);
        std::vector<uint32_t> ref_v(n), got(n);
        printf("%s\"%s\": {", c ? ", " : "", cases[c].name);
        for (size_t v = 0; v < sizeof(vars) / sizeof(vars[0]); ++v) {
            size_t bytes = 64 << 20;
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            for (int it = 0; it < 3; ++it) { bytes = 64 << 20; vars[v].fn(tmp, bytes, ki, ko, vi, vo, n, cases[c].b0, cases[c].b1, 0); }
            cudaEventRecord(a);
            const int iters = 50;
            for (int it = 0; it < iters; ++it) { bytes = 64 << 20; vars[v].fn(tmp, bytes, ki, ko, vi, vo, n, cases[c].b0, cases[c].b1, 0); }
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            cudaError_t e = cudaGetLastError();
            cudaMemcpy(got.data(), vo, 4 * n, cudaMemcpyDeviceToHost);
            if (v == 0) ref_v = got;
            printf("%s\"%s\": {\"us\": %.2f, \"same_as_default\": %s, \"err\": \"%s\"}", v ? ", " : "", vars[v].name, 1e3 * ms / iters,
                   got == ref_v ? "true" : "false", e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
        printf("}");
        cudaFree(ki); cudaFree(ko); cudaFree(vi); cudaFree(vo); cudaFree(ref); cudaFree(tmp);
    }
    printf("}\n");
    return 0;
}
