// Projection forward + tile-intersection emission.  THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false:
// its outputs (means2d, radii, depths) decide every tile/bin index and must be bit-identical to the
// IEEE-fp32 CPU oracle (oracle/raster_oracle.c).  Both kernels are streaming, HBM-bound passes.
//
// Replaces gsplat 1.4.0 fully_fused_projection_packed_fwd + isect_tiles (third-party), reached from
// rfstudio/model/gsplat.py:334-355.
#include "project_math.cuh"

__global__ void __launch_bounds__(256) project_fwd_kernel(int N, const float *__restrict__ means,
                                                           const float *__restrict__ quats,
                                                           const float *__restrict__ scales, CamK cam,
                                                           int32_t *__restrict__ radii, float2 *__restrict__ means2d,
                                                           float *__restrict__ depths, float *__restrict__ conics,
                                                           float *__restrict__ comps,
                                                           int32_t *__restrict__ tiles_per_gauss) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float m[3] = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
    float4 q4 = reinterpret_cast<const float4 *>(quats)[i];
    float q[4] = {q4.x, q4.y, q4.z, q4.w};
    float s[3] = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
    ProjOut o;
    bool ok = gsb_project_one(m, q, s, cam, o);
    if (!ok) {
        radii[i] = 0;
        means2d[i] = make_float2(0.f, 0.f);
        depths[i] = 0.f;
        conics[3 * i] = conics[3 * i + 1] = conics[3 * i + 2] = 0.f;
        comps[i] = 0.f;
        if (tiles_per_gauss) tiles_per_gauss[i] = 0;
        return;
    }
    int r = (int)o.radius;
    radii[i] = r;
    means2d[i] = make_float2(o.mean2d[0], o.mean2d[1]);
    depths[i] = o.pc[2];
    conics[3 * i] = o.conic[0];
    conics[3 * i + 1] = o.conic[1];
    conics[3 * i + 2] = o.conic[2];
    comps[i] = cam.antialiased ? o.comp : 1.0f;
    if (tiles_per_gauss) {
        int x0, x1, y0, y1;
        gsb_tile_range(o.mean2d[0], o.mean2d[1], r, cam.tile_w, cam.tile_h, x0, x1, y0, y1);
        tiles_per_gauss[i] = (x1 - x0) * (y1 - y0);
    }
}

__global__ void __launch_bounds__(256) isect_tiles_kernel(int N, const float2 *__restrict__ means2d,
                                                           const int32_t *__restrict__ radii,
                                                           const float *__restrict__ depths,
                                                           const int64_t *__restrict__ cum_tiles, CamK cam,
                                                           int tile_n_bits, int64_t *__restrict__ isect_ids,
                                                           int32_t *__restrict__ flatten_ids) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int r = radii[i];
    if (r <= 0) return;
    float2 m = means2d[i];
    int x0, x1, y0, y1;
    gsb_tile_range(m.x, m.y, r, cam.tile_w, cam.tile_h, x0, x1, y0, y1);
    int64_t pos = (i == 0) ? 0 : cum_tiles[i - 1];
    int64_t hi = ((int64_t)cam.camera_id << tile_n_bits);
    int64_t dbits = (int64_t)(uint32_t)__float_as_uint(depths[i]);
    for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
            int64_t tid = (int64_t)y * cam.tile_w + x;
            isect_ids[pos] = ((hi | tid) << 32) | dbits;
            flatten_ids[pos] = i;
            ++pos;
        }
}

// Two-stage binning, emission step: thread i handles the i-th Gaussian IN DEPTH ORDER and writes the ids of the tiles
// it touches (row-major) and its own index.  A stable sort of these pairs by tile id alone then yields exactly the
// order of a stable sort on (tile | depth) of Gaussian-major pairs: depth ties keep ascending Gaussian index in
// both (gsb_isect_tiles + gsb_sort_pairs is that single-sort path).
__global__ void __launch_bounds__(256) isect_tiles_ordered_kernel(int N, const float2 *__restrict__ means2d,
                                                                   const int32_t *__restrict__ radii,
                                                                   const int32_t *__restrict__ order,
                                                                   const int64_t *__restrict__ cum_ordered, CamK cam,
                                                                   int64_t cap, uint32_t *__restrict__ tile_keys,
                                                                   int32_t *__restrict__ gauss_ids) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int g = order[i];
    int r = radii[g];
    if (r <= 0) return;
    float2 m = means2d[g];
    int x0, x1, y0, y1;
    gsb_tile_range(m.x, m.y, r, cam.tile_w, cam.tile_h, x0, x1, y0, y1);
    int64_t pos = (i == 0) ? 0 : cum_ordered[i - 1];
    for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
            if (pos < cap) {      // speculative capacity (gsb_batch_*): the farthest intersections are dropped, never written out of bounds
                tile_keys[pos] = (uint32_t)(y * cam.tile_w + x);
                gauss_ids[pos] = g;
            }
            ++pos;
        }
}

// `cap`: capacity of tile_keys / gauss_ids (entries at positions >= cap are dropped).
int gsb_isect_tiles_ordered_cap(int32_t N, const float *means2d, const int32_t *radii, const int32_t *order,
                                const int64_t *cum_ordered, const gsb_camera *cam, int64_t cap, uint32_t *tile_keys,
                                int32_t *gauss_ids, void *stream) {
    GSB_CHECK_ARG(N >= 0 && cam != nullptr);
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(means2d && radii && order && cum_ordered && tile_keys && gauss_ids);
    CamK k = gsb_make_cam(cam);
    isect_tiles_ordered_kernel<<<gsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
        N, reinterpret_cast<const float2 *>(means2d), radii, order, cum_ordered, k, cap, tile_keys, gauss_ids);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

extern "C" __attribute__((visibility("default"))) int gsb_isect_tiles_ordered(int32_t N, const float *means2d, const int32_t *radii,
                                       const int32_t *order, const int64_t *cum_ordered, const gsb_camera *cam,
                                       uint32_t *tile_keys, int32_t *gauss_ids, void *stream) {
    return gsb_isect_tiles_ordered_cap(N, means2d, radii, order, cum_ordered, cam, INT64_MAX, tile_keys, gauss_ids, stream);
}

extern "C" __attribute__((visibility("default"))) int gsb_project_fwd(int32_t N, const float *means, const float *quats, const float *scales,
                               const gsb_camera *cam, int32_t *radii, float *means2d, float *depths,
                               float *conics, float *comps, int32_t *tiles_per_gauss, void *stream) {
    GSB_CHECK_ARG(N >= 0 && cam != nullptr);
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(means && quats && scales && radii && means2d && depths && conics && comps);
    GSB_CHECK_ARG(cam->width > 0 && cam->height > 0);
    CamK k = gsb_make_cam(cam);
    project_fwd_kernel<<<gsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
        N, means, quats, scales, k, radii, reinterpret_cast<float2 *>(means2d), depths, conics, comps,
        tiles_per_gauss);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

extern "C" __attribute__((visibility("default"))) int gsb_isect_tiles(int32_t N, const float *means2d, const int32_t *radii, const float *depths,
                               const int64_t *cum_tiles, const gsb_camera *cam, int64_t *isect_ids,
                               int32_t *flatten_ids, void *stream) {
    GSB_CHECK_ARG(N >= 0 && cam != nullptr);
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(means2d && radii && depths && cum_tiles && isect_ids && flatten_ids);
    CamK k = gsb_make_cam(cam);
    int bits = gsb_tile_bits(k.tile_w * k.tile_h);
    isect_tiles_kernel<<<gsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
        N, reinterpret_cast<const float2 *>(means2d), radii, depths, cum_tiles, k, bits, isect_ids, flatten_ids);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}
