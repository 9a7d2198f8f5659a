// Alpha compositing forward / backward: one 16x16 CTA per tile, the tile's depth-sorted list is staged
// through shared memory in batches of 256 records, every pixel walks it front-to-back (forward) or
// back-to-front from its own last contributor (backward).
//
// Replaces gsplat 1.4.0 rasterize_to_pixels_fwd/bwd (third-party), reached
// from rfstudio/model/gsplat.py:334-355.  Bound: FP32/MUFU issue, not HBM.
#include "gsb_common.cuh"

#define LOG2E 1.4426950408889634f

namespace {

constexpr int BLOCK = GSB_TILE * GSB_TILE;  // 256 threads, one per pixel

// Per-Gaussian record staged in shared memory.  The conic is pre-scaled so that
// exp(-sigma) == exp2(-(qa*dx*dx + qc*dy*dy + qb*dx*dy)).
struct GRec {
    float x, y, opac, qa;
};

template <int CH>
__global__ void __launch_bounds__(BLOCK)
composite_fwd_kernel(int W, int H, int tile_w, const float2 *__restrict__ means2d, const float *__restrict__ conics,
                     const float *__restrict__ colors, const float *__restrict__ opacities,
                     const float *__restrict__ background, const int32_t *__restrict__ offsets,
                     const int32_t *__restrict__ flatten_ids, int n_tiles, int M, float *__restrict__ render,
                     float *__restrict__ alphas, int32_t *__restrict__ last_ids) {
    __shared__ float4 s_g0[BLOCK];     // x, y, opac, qa
    __shared__ float2 s_g1[BLOCK];     // qb, qc
    __shared__ float s_rgb[BLOCK * CH];

    const int tile_id = blockIdx.y * tile_w + blockIdx.x;
    const int tr = threadIdx.y * GSB_TILE + threadIdx.x;
    const int i = blockIdx.y * GSB_TILE + threadIdx.y;
    const int j = blockIdx.x * GSB_TILE + threadIdx.x;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < H && j < W);
    bool done = !inside;

    const int range_start = offsets[tile_id];
    const int range_end = (tile_id == n_tiles - 1) ? M : offsets[tile_id + 1];
    const int num_batches = (range_end - range_start + BLOCK - 1) / BLOCK;

    float T = 1.0f;
    float acc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = 0.f;
    int cur_idx = 0;

    for (int b = 0; b < num_batches; ++b) {
        if (__syncthreads_count(done) >= BLOCK) break;
        const int batch_start = range_start + BLOCK * b;
        const int idx = batch_start + tr;
        if (idx < range_end) {
            const int g = flatten_ids[idx];
            const float2 xy = means2d[g];
            const float ca = conics[3 * g], cb = conics[3 * g + 1], cc = conics[3 * g + 2];
            s_g0[tr] = make_float4(xy.x, xy.y, opacities[g], 0.5f * LOG2E * ca);
            s_g1[tr] = make_float2(LOG2E * cb, 0.5f * LOG2E * cc);
#pragma unroll
            for (int k = 0; k < CH; ++k) s_rgb[tr * CH + k] = colors[(size_t)g * CH + k];
        }
        __syncthreads();
        const int batch_size = min(BLOCK, range_end - batch_start);
        for (int t = 0; t < batch_size && !done; ++t) {
            const float4 g0 = s_g0[t];
            const float2 g1 = s_g1[t];
            const float dx = g0.x - px, dy = g0.y - py;
            const float sigma = g0.w * dx * dx + g1.y * dy * dy + g1.x * dx * dy;
            const float alpha = fminf(GSB_ALPHA_CLAMP, g0.z * exp2f(-sigma));
            if (sigma < 0.f || alpha < GSB_ALPHA_MIN) continue;
            const float next_T = T * (1.0f - alpha);
            if (next_T <= GSB_T_STOP) {
                done = true;
                break;
            }
            const float vis = alpha * T;
#pragma unroll
            for (int k = 0; k < CH; ++k) acc[k] += s_rgb[t * CH + k] * vis;
            cur_idx = batch_start + t;
            T = next_T;
        }
    }
    if (inside) {
        const size_t pix = (size_t)i * W + j;
        alphas[pix] = 1.0f - T;
#pragma unroll
        for (int k = 0; k < CH; ++k)
            render[pix * CH + k] = background ? acc[k] + T * background[k] : acc[k];
        last_ids[pix] = cur_idx;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int CH>
__global__ void __launch_bounds__(BLOCK)
composite_bwd_kernel(int W, int H, int tile_w, const float2 *__restrict__ means2d, const float *__restrict__ conics,
                     const float *__restrict__ colors, const float *__restrict__ opacities,
                     const float *__restrict__ background, const int32_t *__restrict__ offsets,
                     const int32_t *__restrict__ flatten_ids, int n_tiles, int M,
                     const float *__restrict__ alphas, const int32_t *__restrict__ last_ids,
                     const float *__restrict__ v_render, const float *__restrict__ v_alphas,
                     float *__restrict__ v_means2d, float *__restrict__ v_conics, float *__restrict__ v_colors,
                     float *__restrict__ v_opacities) {
    __shared__ int32_t s_id[BLOCK];
    __shared__ float4 s_g0[BLOCK];  // x, y, opac, ca
    __shared__ float2 s_g1[BLOCK];  // cb, cc
    __shared__ float s_rgb[BLOCK * CH];
    __shared__ int s_max[BLOCK / 32];

    const int tile_id = blockIdx.y * tile_w + blockIdx.x;
    const int tr = threadIdx.y * GSB_TILE + threadIdx.x;
    const int lane = tr & 31, warp = tr >> 5;
    const int i = blockIdx.y * GSB_TILE + threadIdx.y;
    const int j = blockIdx.x * GSB_TILE + threadIdx.x;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < H && j < W);
    const size_t pix = inside ? (size_t)i * W + j : 0;

    const int range_start = offsets[tile_id];
    const int range_end = (tile_id == n_tiles - 1) ? M : offsets[tile_id + 1];
    const int num_batches = (range_end - range_start + BLOCK - 1) / BLOCK;

    const float T_final = inside ? 1.0f - alphas[pix] : 1.0f;
    float T = T_final;
    float buffer[CH];
    float v_out[CH];
    float bg_dot = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        buffer[k] = 0.f;
        v_out[k] = inside ? v_render[pix * CH + k] : 0.f;
        if (background) bg_dot += background[k] * v_out[k];
    }
    const float v_a_out = inside ? v_alphas[pix] : 0.f;
    const int bin_final = inside ? last_ids[pix] : 0;

    // the block starts from the deepest contributor of any of its pixels
    int wmax = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    int block_max = s_max[0];
#pragma unroll
    for (int w = 1; w < BLOCK / 32; ++w) block_max = max(block_max, s_max[w]);
    const int warp_bin_final = wmax;

    for (int b = 0; b < num_batches; ++b) {
        __syncthreads();
        // batches walk the list from its END: batch b covers [batch_end-BLOCK+1, batch_end]
        const int batch_end = range_end - 1 - BLOCK * b;
        const int batch_size = min(BLOCK, batch_end + 1 - range_start);
        if (batch_end - batch_size + 1 > block_max) continue;  // uniform across the block
        const int idx = batch_end - tr;
        if (idx >= range_start) {
            const int g = flatten_ids[idx];
            s_id[tr] = g;
            const float2 xy = means2d[g];
            s_g0[tr] = make_float4(xy.x, xy.y, opacities[g], conics[3 * g]);
            s_g1[tr] = make_float2(conics[3 * g + 1], conics[3 * g + 2]);
#pragma unroll
            for (int k = 0; k < CH; ++k) s_rgb[tr * CH + k] = colors[(size_t)g * CH + k];
        }
        __syncthreads();
        // entry t of the batch is list position batch_end - t (descending depth)
        const int t0 = max(0, batch_end - warp_bin_final);
        for (int t = t0; t < batch_size; ++t) {
            bool valid = inside && (batch_end - t <= bin_final);
            const float4 g0 = s_g0[t];
            const float2 g1 = s_g1[t];
            const float dx = g0.x - px, dy = g0.y - py;
            float alpha = 0.f, vis = 0.f;
            if (valid) {
                const float sigma = 0.5f * (g0.w * dx * dx + g1.y * dy * dy) + g1.x * dx * dy;
                vis = __expf(-sigma);
                alpha = fminf(GSB_ALPHA_CLAMP, g0.z * vis);
                if (sigma < 0.f || alpha < GSB_ALPHA_MIN) valid = false;
            }
            if (!__any_sync(0xffffffffu, valid)) continue;
            float v_rgb[CH];
            float v_ca = 0.f, v_cb = 0.f, v_cc = 0.f, v_x = 0.f, v_y = 0.f, v_op = 0.f;
#pragma unroll
            for (int k = 0; k < CH; ++k) v_rgb[k] = 0.f;
            if (valid) {
                const float ra = 1.0f / (1.0f - alpha);
                T *= ra;
                const float fac = alpha * T;
                float v_alpha = 0.f;
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const float c = s_rgb[t * CH + k];
                    v_rgb[k] = fac * v_out[k];
                    v_alpha += (c * T - buffer[k] * ra) * v_out[k];
                    buffer[k] += c * fac;
                }
                v_alpha += T_final * ra * v_a_out;
                if (background) v_alpha += -T_final * ra * bg_dot;
                if (g0.z * vis <= GSB_ALPHA_CLAMP) {
                    const float v_sigma = -g0.z * vis * v_alpha;
                    v_ca = 0.5f * v_sigma * dx * dx;
                    v_cb = v_sigma * dx * dy;
                    v_cc = 0.5f * v_sigma * dy * dy;
                    v_x = v_sigma * (g0.w * dx + g1.x * dy);
                    v_y = v_sigma * (g1.x * dx + g1.y * dy);
                    v_op = vis * v_alpha;
                }
            }
#pragma unroll
            for (int k = 0; k < CH; ++k) v_rgb[k] = warp_sum(v_rgb[k]);
            v_ca = warp_sum(v_ca); v_cb = warp_sum(v_cb); v_cc = warp_sum(v_cc);
            v_x = warp_sum(v_x); v_y = warp_sum(v_y); v_op = warp_sum(v_op);
            if (lane == 0) {
                const int g = s_id[t];
#pragma unroll
                for (int k = 0; k < CH; ++k) atomicAdd(v_colors + (size_t)g * CH + k, v_rgb[k]);
                atomicAdd(v_conics + 3 * g, v_ca);
                atomicAdd(v_conics + 3 * g + 1, v_cb);
                atomicAdd(v_conics + 3 * g + 2, v_cc);
                atomicAdd(v_means2d + 2 * g, v_x);
                atomicAdd(v_means2d + 2 * g + 1, v_y);
                atomicAdd(v_opacities + g, v_op);
            }
        }
    }
}

template <int CH>
int launch_fwd(int W, int H, const float *means2d, const float *conics, const float *colors,
               const float *opacities, const float *background, const int32_t *offsets,
               const int32_t *flatten_ids, int64_t M, float *render, float *alphas, int32_t *last_ids,
               cudaStream_t st) {
    int tw = (W + GSB_TILE - 1) / GSB_TILE, th = (H + GSB_TILE - 1) / GSB_TILE;
    dim3 grid(tw, th), block(GSB_TILE, GSB_TILE);
    composite_fwd_kernel<CH><<<grid, block, 0, st>>>(W, H, tw, reinterpret_cast<const float2 *>(means2d), conics,
                                                     colors, opacities, background, offsets, flatten_ids, tw * th,
                                                     (int)M, render, alphas, last_ids);
    return 0;
}

template <int CH>
int launch_bwd(int W, int H, const float *means2d, const float *conics, const float *colors,
               const float *opacities, const float *background, const int32_t *offsets,
               const int32_t *flatten_ids, int64_t M, const float *alphas, const int32_t *last_ids,
               const float *v_render, const float *v_alphas, float *v_means2d, float *v_conics, float *v_colors,
               float *v_opacities, cudaStream_t st) {
    int tw = (W + GSB_TILE - 1) / GSB_TILE, th = (H + GSB_TILE - 1) / GSB_TILE;
    dim3 grid(tw, th), block(GSB_TILE, GSB_TILE);
    composite_bwd_kernel<CH><<<grid, block, 0, st>>>(
        W, H, tw, reinterpret_cast<const float2 *>(means2d), conics, colors, opacities, background, offsets,
        flatten_ids, tw * th, (int)M, alphas, last_ids, v_render, v_alphas, v_means2d, v_conics, v_colors,
        v_opacities);
    return 0;
}

}  // namespace

#define GSB_DISPATCH_CH(CHV, CALL)             \
    switch (CHV) {                             \
        case 1: { constexpr int C_ = 1; CALL; break; }   \
        case 2: { constexpr int C_ = 2; CALL; break; }   \
        case 3: { constexpr int C_ = 3; CALL; break; }   \
        case 4: { constexpr int C_ = 4; CALL; break; }   \
        case 8: { constexpr int C_ = 8; CALL; break; }   \
        case 16: { constexpr int C_ = 16; CALL; break; } \
        default:                               \
            gsb_set_error("%s: unsupported channel count %d (supported: 1,2,3,4,8,16; pad to the next)", __func__, CHV); \
            return GSB_EINVAL;                 \
    }

extern "C" __attribute__((visibility("default"))) int gsb_composite_fwd(int32_t width, int32_t height, int32_t channels, const float *means2d,
                                 const float *conics, const float *colors, const float *opacities,
                                 const float *background, const int32_t *offsets, const int32_t *flatten_ids,
                                 int64_t M, float *render, float *alphas, int32_t *last_ids, void *stream) {
    GSB_CHECK_ARG(width > 0 && height > 0 && M >= 0 && M < 2147483647LL);
    GSB_CHECK_ARG(offsets && render && alphas && last_ids);
    GSB_CHECK_ARG(M == 0 || (means2d && conics && colors && opacities && flatten_ids));
    GSB_DISPATCH_CH(channels, (launch_fwd<C_>(width, height, means2d, conics, colors, opacities, background,
                                               offsets, flatten_ids, M, render, alphas, last_ids,
                                               (cudaStream_t)stream)));
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

extern "C" __attribute__((visibility("default"))) int gsb_composite_bwd(int32_t width, int32_t height, int32_t channels, const float *means2d,
                                 const float *conics, const float *colors, const float *opacities,
                                 const float *background, const int32_t *offsets, const int32_t *flatten_ids,
                                 int64_t M, const float *alphas, const int32_t *last_ids, const float *v_render,
                                 const float *v_alphas, float *v_means2d, float *v_conics, float *v_colors,
                                 float *v_opacities, void *stream) {
    GSB_CHECK_ARG(width > 0 && height > 0 && M >= 0 && M < 2147483647LL);
    GSB_CHECK_ARG(offsets && alphas && last_ids && v_render && v_alphas);
    if (M == 0) return GSB_OK;
    GSB_CHECK_ARG(means2d && conics && colors && opacities && flatten_ids);
    GSB_CHECK_ARG(v_means2d && v_conics && v_colors && v_opacities);
    GSB_DISPATCH_CH(channels, (launch_bwd<C_>(width, height, means2d, conics, colors, opacities, background,
                                               offsets, flatten_ids, M, alphas, last_ids, v_render, v_alphas,
                                               v_means2d, v_conics, v_colors, v_opacities,
                                               (cudaStream_t)stream)));
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}
