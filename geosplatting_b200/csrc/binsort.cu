// Tile binning: prefix sum of tiles-per-Gaussian, stable radix sort of the (tile|depth) keys, per-tile
// offsets.  Integer/byte work, HBM-bound; the sort is cub::DeviceRadixSort restricted to the live key
// bits (32 depth bits + tile bits + camera bits), as gsplat 1.4.0 isect_tiles/isect_offset_encode do.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
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

// ---- two-stage binning ----------------------------------------------------------------------------------------
// The single sort above moves M 12-byte pairs through ceil((32 + tile bits) / 8) = 6 radix passes.  Sorting the N
// Gaussians by depth first (4 passes over N 8-byte pairs) and then the M (tile id, Gaussian) pairs by tile id alone
// (2 passes over M 8-byte pairs, stable) produces the identical order with about a third of the traffic, and never
// materialises the 64-bit keys.

struct OrderedTiles {
    const int32_t *tiles_per_gauss;
    const int32_t *order;
    __host__ __device__ __forceinline__ int64_t operator()(const int32_t &i) const {
        return (int64_t)tiles_per_gauss[order[i]];
    }
};

__global__ void __launch_bounds__(256) iota_kernel(int N, int32_t *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) out[i] = i;
}

static size_t depth_sort_temp_bytes(int32_t N) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs((void *)nullptr, b, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, (int)N, 0, 32);
    return b;
}

static size_t ordered_scan_temp_bytes(int32_t N) {
    size_t b = 0;
    thrust::transform_iterator<OrderedTiles, thrust::counting_iterator<int32_t>, int64_t> in(
        thrust::counting_iterator<int32_t>(0), OrderedTiles{nullptr, nullptr});
    cub::DeviceScan::InclusiveSum((void *)nullptr, b, in, (int64_t *)nullptr, (int)N);
    return b;
}

static size_t tile_sort_temp_bytes(int64_t M) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs((void *)nullptr, b, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, M, 0, 32);
    return b;
}

extern "C" __attribute__((visibility("default"))) int gsb_bin2_workspace_bytes(int32_t N, int64_t M, size_t *bytes_host) {
    GSB_CHECK_ARG(N >= 0 && M >= 0 && bytes_host != nullptr);
    // count step: sorted depth keys [N] u32 + iota [N] + cub temp;  sort step: tile keys [M] + sorted tile keys [M] +
    // unsorted Gaussian ids [M] + cub temp.  The two steps never overlap in time: take the larger.
    size_t a = 2 * align256(sizeof(uint32_t) * (size_t)N) +
               align256(depth_sort_temp_bytes(N) > ordered_scan_temp_bytes(N) ? depth_sort_temp_bytes(N)
                                                                                : ordered_scan_temp_bytes(N));
    size_t b = 3 * align256(sizeof(uint32_t) * (size_t)M) + align256(tile_sort_temp_bytes(M));
    *bytes_host = (a > b ? a : b) + 256;
    return GSB_OK;
}

// Depth order of the Gaussians (stable: ties keep ascending index), the prefix sum of tiles-per-Gaussian in that
// order, and M published to *total_out (device or pinned host memory, see gsb_isect_total).
// `total_out` may be null for the batch driver, whose gsb_bin2_publish reads M off cum_ordered itself (one launch less).
int gsb_bin2_order(int32_t N, const float *depths, const int32_t *tiles_per_gauss, int32_t *order, int64_t *cum_ordered,
                   int64_t *total_out, void *workspace, size_t workspace_bytes, void *stream) {
    GSB_CHECK_ARG(N >= 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        if (total_out) GSB_CHECK_CUDA(cudaMemsetAsync(total_out, 0, sizeof(int64_t), st));
        return GSB_OK;
    }
    GSB_CHECK_ARG(depths && tiles_per_gauss && order && cum_ordered && workspace);
    size_t need = 0;
    GSB_CHECK_ARG(gsb_bin2_workspace_bytes(N, 0, &need) == GSB_OK);
    if (need > workspace_bytes) {
        gsb_set_error("gsb_bin2_count: workspace too small (%zu < %zu)", workspace_bytes, need);
        return GSB_ENOMEM;
    }
    char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    uint32_t *keys_sorted = reinterpret_cast<uint32_t *>(p); p += align256(sizeof(uint32_t) * (size_t)N);
    int32_t *iota = reinterpret_cast<int32_t *>(p); p += align256(sizeof(uint32_t) * (size_t)N);
    void *temp = p;
    iota_kernel<<<gsb_div_up(N, 256), 256, 0, st>>>(N, iota);
    // depths of visible Gaussians are positive floats (>= near plane): their bit patterns order like the values;
    // culled Gaussians carry depth 0 and emit nothing.
    size_t tb = depth_sort_temp_bytes(N);
    GSB_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, reinterpret_cast<const uint32_t *>(depths), keys_sorted,
                                                   iota, order, (int)N, 0, 32, st));
    thrust::transform_iterator<OrderedTiles, thrust::counting_iterator<int32_t>, int64_t> in(
        thrust::counting_iterator<int32_t>(0), OrderedTiles{tiles_per_gauss, order});
    tb = ordered_scan_temp_bytes(N);
    GSB_CHECK_CUDA(cub::DeviceScan::InclusiveSum(temp, tb, in, cum_ordered, (int)N, st));
    if (total_out) isect_total_kernel<<<1, 1, 0, st>>>(N, cum_ordered, total_out);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

extern "C" __attribute__((visibility("default"))) int gsb_bin2_count(int32_t N, const float *depths, const int32_t *tiles_per_gauss, int32_t *order,
                              int64_t *cum_ordered, int64_t *total_out, void *workspace, size_t workspace_bytes,
                              void *stream) {
    GSB_CHECK_ARG(total_out != nullptr);
    return gsb_bin2_order(N, depths, tiles_per_gauss, order, cum_ordered, total_out, workspace, workspace_bytes, stream);
}

__global__ void __launch_bounds__(256) tile_offsets_kernel(int64_t M, const uint32_t *__restrict__ tile_keys, int n_tiles,
                                                            int32_t *__restrict__ offsets) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    int cur = (int)tile_keys[i];
    int prev = (i == 0) ? -1 : (int)tile_keys[i - 1];
    for (int id = prev + 1; id <= cur; ++id) offsets[id] = (int32_t)i;
    if (i == M - 1)
        for (int id = cur + 1; id < n_tiles; ++id) offsets[id] = (int32_t)M;
}

// Emission in depth order, stable sort by tile id, per-tile offsets.  -> flatten_ids[M], offsets[tile_w * tile_h]:
// bit-identical to gsb_isect_tiles + gsb_sort_pairs + gsb_isect_offsets.
extern "C" __attribute__((visibility("default"))) int gsb_bin2_sort(int32_t N, int64_t M, const float *means2d, const int32_t *radii, const int32_t *order,
                             const int64_t *cum_ordered, const gsb_camera *cam, int32_t *flatten_ids,
                             int32_t *offsets, void *workspace, size_t workspace_bytes, void *stream) {
    GSB_CHECK_ARG(N >= 0 && M >= 0 && cam != nullptr && offsets != nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    const int tile_w = (cam->width + GSB_TILE - 1) / GSB_TILE, tile_h = (cam->height + GSB_TILE - 1) / GSB_TILE;
    const int n_tiles = tile_w * tile_h;
    if (M == 0 || N == 0) {
        GSB_CHECK_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int32_t) * (size_t)n_tiles, st));
        return GSB_OK;
    }
    GSB_CHECK_ARG(means2d && radii && order && cum_ordered && flatten_ids && workspace);
    size_t need = 0;
    GSB_CHECK_ARG(gsb_bin2_workspace_bytes(0, M, &need) == GSB_OK);
    if (need > workspace_bytes) {
        gsb_set_error("gsb_bin2_sort: workspace too small (%zu < %zu)", workspace_bytes, need);
        return GSB_ENOMEM;
    }
    char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    uint32_t *keys = reinterpret_cast<uint32_t *>(p); p += align256(sizeof(uint32_t) * (size_t)M);
    uint32_t *keys_sorted = reinterpret_cast<uint32_t *>(p); p += align256(sizeof(uint32_t) * (size_t)M);
    int32_t *gids = reinterpret_cast<int32_t *>(p); p += align256(sizeof(uint32_t) * (size_t)M);
    void *temp = p;
    int rc = gsb_isect_tiles_ordered(N, means2d, radii, order, cum_ordered, cam, keys, gids, stream);
    if (rc != GSB_OK) return rc;
    size_t tb = tile_sort_temp_bytes(M);
    GSB_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, keys, keys_sorted, gids, flatten_ids, M, 0,
                                                   gsb_tile_bits(n_tiles), st));
    tile_offsets_kernel<<<gsb_div_up(M, 256), 256, 0, st>>>(M, keys_sorted, n_tiles, offsets);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

// ---- speculative capacity: nothing on the host waits for M --------------------------------------------------------
// The batch driver (view.cu: gsb_batch_*) sizes the M-dependent arrays from a capacity `cap` the caller chose (e.g. the
// count of this camera's previous visit + a margin) and keeps M on the device: m_eff = min(M, cap) is what every later
// kernel reads, the raw M goes to `total_out` (pinned host memory) for the caller to check and adapt the capacity.

int gsb_isect_tiles_ordered_cap(int32_t N, const float *means2d, const int32_t *radii, const int32_t *order,
                                const int64_t *cum_ordered, const gsb_camera *cam, int64_t cap, uint32_t *tile_keys,
                                int32_t *gauss_ids, void *stream);

__global__ void publish_total_kernel(int N, const int64_t *__restrict__ cum_ordered, int64_t cap, int64_t *m_eff,
                                     volatile int64_t *total_out) {
    const int64_t m = N > 0 ? cum_ordered[N - 1] : 0;
    *m_eff = m < cap ? m : cap;
    if (total_out) {
        *total_out = m;
        __threadfence_system();
    }
}

int gsb_bin2_publish(int32_t N, const int64_t *cum_ordered, int64_t cap, int64_t *m_eff, int64_t *total_out, void *stream) {
    GSB_CHECK_ARG(N >= 0 && cap >= 0 && m_eff != nullptr && (N == 0 || cum_ordered != nullptr));
    publish_total_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(N, cum_ordered, cap, m_eff, total_out);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

// Sorted keys past m_eff are the 0xFFFFFFFF fill: they compare above every tile id in the sorted bit range.
__global__ void __launch_bounds__(256) tile_offsets_cap_kernel(int64_t cap, const int64_t *__restrict__ m_eff,
                                                                const uint32_t *__restrict__ tile_keys, int n_tiles,
                                                                int32_t *__restrict__ offsets) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t M = *m_eff;
    if (i >= M || i >= cap) return;
    int cur = (int)tile_keys[i];
    int prev = (i == 0) ? -1 : (int)tile_keys[i - 1];
    for (int id = prev + 1; id <= cur; ++id) offsets[id] = (int32_t)i;
    if (i == M - 1)
        for (int id = cur + 1; id < n_tiles; ++id) offsets[id] = (int32_t)M;
}

// gsb_bin2_sort with a capacity instead of the count: emission (guarded), stable sort of `cap` pairs by tile id -- the
// unused tail is filled with all-ones keys and sorts behind every tile --, per-tile offsets from the device-side count.
int gsb_bin2_sort_cap(int32_t N, int64_t cap, const int64_t *m_eff, const float *means2d, const int32_t *radii,
                      const int32_t *order, const int64_t *cum_ordered, const gsb_camera *cam, int32_t *flatten_ids,
                      int32_t *offsets, void *workspace, size_t workspace_bytes, void *stream) {
    GSB_CHECK_ARG(N >= 0 && cap >= 0 && cam != nullptr && offsets != nullptr && m_eff != nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    const int tile_w = (cam->width + GSB_TILE - 1) / GSB_TILE, tile_h = (cam->height + GSB_TILE - 1) / GSB_TILE;
    const int n_tiles = tile_w * tile_h;
    GSB_CHECK_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int32_t) * (size_t)n_tiles, st));   // stays when M == 0
    if (cap == 0 || N == 0) return GSB_OK;
    GSB_CHECK_ARG(means2d && radii && order && cum_ordered && flatten_ids && workspace);
    size_t need = 0;
    GSB_CHECK_ARG(gsb_bin2_workspace_bytes(0, cap, &need) == GSB_OK);
    if (need > workspace_bytes) {
        gsb_set_error("gsb_bin2_sort_cap: workspace too small (%zu < %zu)", workspace_bytes, need);
        return GSB_ENOMEM;
    }
    char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    uint32_t *keys = reinterpret_cast<uint32_t *>(p); p += align256(sizeof(uint32_t) * (size_t)cap);
    uint32_t *keys_sorted = reinterpret_cast<uint32_t *>(p); p += align256(sizeof(uint32_t) * (size_t)cap);
    int32_t *gids = reinterpret_cast<int32_t *>(p); p += align256(sizeof(uint32_t) * (size_t)cap);
    void *temp = p;
    GSB_CHECK_CUDA(cudaMemsetAsync(keys, 0xFF, sizeof(uint32_t) * (size_t)cap, st));
    GSB_CHECK_CUDA(cudaMemsetAsync(gids, 0, sizeof(int32_t) * (size_t)cap, st));
    int rc = gsb_isect_tiles_ordered_cap(N, means2d, radii, order, cum_ordered, cam, cap, keys, gids, stream);
    if (rc != GSB_OK) return rc;
    size_t tb = tile_sort_temp_bytes(cap);
    GSB_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, keys, keys_sorted, gids, flatten_ids, cap, 0,
                                                   gsb_tile_bits(n_tiles), st));
    tile_offsets_cap_kernel<<<gsb_div_up(cap, 256), 256, 0, st>>>(cap, m_eff, keys_sorted, n_tiles, offsets);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}
