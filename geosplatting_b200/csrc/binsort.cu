// Tile binning: prefix sum of tiles-per-Gaussian, stable radix sort of the (tile|depth) keys, per-tile
// offsets.  Integer/byte work, HBM-bound; the sort is cub::DeviceRadixSort restricted to the live key
// bits (32 depth bits + tile bits + camera bits), as gsplat 1.4.0 isect_tiles/isect_offset_encode do.
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>

#include "gsb_common.cuh"

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" __attribute__((visibility("default"))) int gsb_bin_workspace_bytes(int32_t N, int64_t M, size_t *bytes_host) {
    GSB_CHECK_ARG(N >= 0 && M >= 0 && bytes_host != nullptr);
    size_t scan_b = 0, sort_b = 0;
    GSB_CHECK_CUDA(cub::DeviceScan::InclusiveSum((void *)nullptr, scan_b, (const int32_t *)nullptr,
                                                 (int64_t *)nullptr, (int)N));
    GSB_CHECK_CUDA(cub::DeviceRadixSort::SortPairs((void *)nullptr, sort_b, (const int64_t *)nullptr,
                                                   (int64_t *)nullptr, (const int32_t *)nullptr,
                                                   (int32_t *)nullptr, M, 0, 64));
    *bytes_host = align256(scan_b > sort_b ? scan_b : sort_b) + 256;
    return GSB_OK;
}

struct I32ToI64 {
    __host__ __device__ __forceinline__ int64_t operator()(const int32_t &v) const { return (int64_t)v; }
};

extern "C" __attribute__((visibility("default"))) int gsb_isect_scan(int32_t N, const int32_t *tiles_per_gauss, int64_t *cum_tiles, void *workspace,
                              size_t workspace_bytes, void *stream) {
    GSB_CHECK_ARG(N >= 0);
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(tiles_per_gauss && cum_tiles && workspace);
    size_t need = 0;
    thrust::transform_iterator<I32ToI64, const int32_t *, int64_t> in(tiles_per_gauss, I32ToI64());
    GSB_CHECK_CUDA(cub::DeviceScan::InclusiveSum((void *)nullptr, need, in, cum_tiles, (int)N));
    if (need > workspace_bytes) {
        gsb_set_error("gsb_isect_scan: workspace too small (%zu < %zu)", workspace_bytes, need);
        return GSB_ENOMEM;
    }
    GSB_CHECK_CUDA(cub::DeviceScan::InclusiveSum(workspace, need, in, cum_tiles, (int)N, (cudaStream_t)stream));
    return GSB_OK;
}

__global__ void isect_total_kernel(int N, const int64_t *__restrict__ cum_tiles, volatile int64_t *total_out) {
    *total_out = cum_tiles[N - 1];
    __threadfence_system();
}

extern "C" __attribute__((visibility("default"))) int gsb_isect_total(int32_t N, const int64_t *cum_tiles, int64_t *total_out, void *stream) {
    GSB_CHECK_ARG(N >= 0 && total_out != nullptr);
    if (N == 0) {
        GSB_CHECK_CUDA(cudaMemsetAsync(total_out, 0, sizeof(int64_t), (cudaStream_t)stream));
        return GSB_OK;
    }
    GSB_CHECK_ARG(cum_tiles != nullptr);
    isect_total_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(N, cum_tiles, total_out);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

extern "C" __attribute__((visibility("default"))) int gsb_sort_pairs(int64_t M, int32_t key_bits, const int64_t *keys_in, const int32_t *vals_in,
                              int64_t *keys_out, int32_t *vals_out, void *workspace, size_t workspace_bytes,
                              void *stream) {
    GSB_CHECK_ARG(M >= 0 && key_bits > 0 && key_bits <= 64);
    if (M == 0) return GSB_OK;
    GSB_CHECK_ARG(keys_in && vals_in && keys_out && vals_out && workspace);
    size_t need = 0;
    GSB_CHECK_CUDA(cub::DeviceRadixSort::SortPairs((void *)nullptr, need, keys_in, keys_out, vals_in, vals_out, M,
                                                   0, key_bits));
    if (need > workspace_bytes) {
        gsb_set_error("gsb_sort_pairs: workspace too small (%zu < %zu)", workspace_bytes, need);
        return GSB_ENOMEM;
    }
    GSB_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(workspace, need, keys_in, keys_out, vals_in, vals_out, M, 0,
                                                   key_bits, (cudaStream_t)stream));
    return GSB_OK;
}

// offsets[id] = lower_bound over the sorted (camera,tile) ids.  One thread per sorted entry writes the
// offsets of every id in (prev_id, cur_id]; the thread of the last entry also fills the tail.
__global__ void __launch_bounds__(256) isect_offsets_kernel(int64_t M, const int64_t *__restrict__ keys,
                                                             int64_t total_ids, int n_tiles, int tile_n_bits,
                                                             int32_t *__restrict__ offsets) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    auto flat_id = [&](int64_t key) -> int64_t {
        int64_t hi = key >> 32;
        int64_t cam = hi >> tile_n_bits;
        int64_t tile = hi & (((int64_t)1 << tile_n_bits) - 1);
        return cam * n_tiles + tile;
    };
    int64_t cur = flat_id(keys[i]);
    int64_t prev = (i == 0) ? -1 : flat_id(keys[i - 1]);
    for (int64_t id = prev + 1; id <= cur; ++id) offsets[id] = (int32_t)i;
    if (i == M - 1)
        for (int64_t id = cur + 1; id < total_ids; ++id) offsets[id] = (int32_t)M;
}

extern "C" __attribute__((visibility("default"))) int gsb_isect_offsets(int64_t M, const int64_t *sorted_isect_ids, int32_t n_cameras, int32_t tile_w,
                                 int32_t tile_h, int32_t *offsets, void *stream) {
    GSB_CHECK_ARG(M >= 0 && n_cameras > 0 && tile_w > 0 && tile_h > 0 && offsets != nullptr);
    int n_tiles = tile_w * tile_h;
    int64_t total = (int64_t)n_cameras * n_tiles;
    if (M == 0) {
        GSB_CHECK_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int32_t) * total, (cudaStream_t)stream));
        return GSB_OK;
    }
    GSB_CHECK_ARG(sorted_isect_ids != nullptr);
    isect_offsets_kernel<<<gsb_div_up(M, 256), 256, 0, (cudaStream_t)stream>>>(
        M, sorted_isect_ids, total, n_tiles, gsb_tile_bits(n_tiles), offsets);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}
