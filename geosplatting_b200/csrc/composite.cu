// Alpha compositing forward / backward: one 256-thread CTA per 16x16 tile.  The tile's depth-sorted list is
// staged through shared memory in batches of 256 records; each of the 8 warps owns an 8x4 pixel sub-rectangle
// and first CULLS the batch against it (conservative axis-aligned extent of the alpha >= 1/255 ellipse, one
// ballot per 32 records), then walks only its hits front-to-back (forward) or back-to-front from its pixels'
// last contributors (backward).  GeoSplatting's Gaussians are a few pixels wide, so a warp typically keeps
// ~1/3 of a tile's list; culling never changes results because a culled pair is exactly one the reference
// kernel would `continue` on (alpha < 1/255 at every pixel centre of the sub-rectangle).
// Backward: per-lane partial gradients are summed with a transposing butterfly (14 shuffles for 9 values instead
// of 45) and written with one atomic per value per (Gaussian, warp).
//
// Replaces gsplat 1.4.0 rasterize_to_pixels_fwd/bwd (third-party), reached from
// rfstudio/model/gsplat.py:334-355.  Bound: FP32 / MUFU issue, not HBM.
#include "gsb_common.cuh"

#define LOG2E 1.4426950408889634f

namespace {

constexpr int BLOCK = GSB_TILE * GSB_TILE;  // 256 threads
constexpr int WARPS = BLOCK / 32;
constexpr int SUB_W = 8, SUB_H = 4;         // pixel footprint of one warp

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Pixel owned by a thread: warp w -> sub-rectangle (w&1, w>>1), lane l -> (l&7, l>>3) inside it.
__device__ __forceinline__ void thread_pixel(int tr, int &lx, int &ly) {
    int w = tr >> 5, l = tr & 31;
    lx = (w & 1) * SUB_W + (l & 7);
    ly = (w >> 1) * SUB_H + (l >> 3);
}

// Half extents of {d : 0.5 d^T C d <= tau}, tau = ln(255 * opac): the only region where alpha >= 1/255.
// Returns a negative extent when the Gaussian can contribute nowhere.  Inflated so that rounding in the
// per-pixel evaluation can never contradict a cull.
__device__ __forceinline__ float2 alpha_extent(float ca, float cb, float cc, float opac) {
    float t = 255.0f * opac;
    if (!(t > 1.0f)) return make_float2(-1e30f, -1e30f);
    float tau2 = 2.0f * logf(t);
    float det = ca * cc - cb * cb;
    if (!(det > 0.f)) return make_float2(1e30f, 1e30f);  // degenerate conic: never cull
    float hx = sqrtf(tau2 * cc / det), hy = sqrtf(tau2 * ca / det);
    return make_float2(hx * 1.0005f + 0.02f, hy * 1.0005f + 0.02f);
}

template <int CH>
__global__ void __launch_bounds__(BLOCK)
composite_fwd_kernel(int W, int H, int tile_w, const float2 *__restrict__ means2d, const float *__restrict__ conics,
                     const float *__restrict__ colors, const float *__restrict__ opacities,
                     const float *__restrict__ background, const int32_t *__restrict__ offsets,
                     const int32_t *__restrict__ flatten_ids, int n_tiles, int M, float *__restrict__ render,
                     float *__restrict__ alphas, int32_t *__restrict__ last_ids) {
    __shared__ float4 s_k[BLOCK];     // x, y, hx, hy
    __shared__ float4 s_q[BLOCK];     // qa, qb, qc (log2e-scaled conic), opac
    __shared__ float s_rgb[BLOCK * CH];

    const int tile_id = blockIdx.y * tile_w + blockIdx.x;
    const int tr = threadIdx.x;
    const int lane = tr & 31, warp = tr >> 5;
    int lx, ly;
    thread_pixel(tr, lx, ly);
    const int i = blockIdx.y * GSB_TILE + ly;
    const int j = blockIdx.x * GSB_TILE + lx;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < H && j < W);
    bool done = !inside;
    // centre of the warp's sub-rectangle in pixel-centre coordinates
    const float rcx = (float)(blockIdx.x * GSB_TILE + (warp & 1) * SUB_W) + 0.5f * SUB_W;
    const float rcy = (float)(blockIdx.y * GSB_TILE + (warp >> 1) * SUB_H) + 0.5f * SUB_H;
    const float rhx = 0.5f * (SUB_W - 1), rhy = 0.5f * (SUB_H - 1);

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
            const float op = opacities[g];
            const float2 ext = alpha_extent(ca, cb, cc, op);
            s_k[tr] = make_float4(xy.x, xy.y, ext.x, ext.y);
            s_q[tr] = make_float4(0.5f * LOG2E * ca, LOG2E * cb, 0.5f * LOG2E * cc, op);
#pragma unroll
            for (int k = 0; k < CH; ++k) s_rgb[tr * CH + k] = colors[(size_t)g * CH + k];
        } else {
            s_k[tr] = make_float4(0.f, 0.f, -1e30f, -1e30f);
        }
        __syncthreads();
        if (__all_sync(0xffffffffu, done)) continue;  // this warp is finished; it still helps staging
#pragma unroll 1
        for (int c = 0; c < BLOCK / 32; ++c) {
            const float4 kc = s_k[c * 32 + lane];
            const bool hit = (fabsf(kc.x - rcx) <= kc.z + rhx) && (fabsf(kc.y - rcy) <= kc.w + rhy);
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                const int t = c * 32 + __ffs(m) - 1;
                m &= m - 1;
                if (done) continue;
                const float4 kk = s_k[t];
                const float4 q = s_q[t];
                const float dx = kk.x - px, dy = kk.y - py;
                const float sigma = q.x * dx * dx + q.z * dy * dy + q.y * dx * dy;
                const float alpha = fminf(GSB_ALPHA_CLAMP, q.w * ex2_approx(-sigma));
                if (sigma < 0.f || alpha < GSB_ALPHA_MIN) continue;
                const float next_T = T * (1.0f - alpha);
                if (next_T <= GSB_T_STOP) {
                    done = true;
                    continue;
                }
                const float vis = alpha * T;
#pragma unroll
                for (int k = 0; k < CH; ++k) acc[k] += s_rgb[t * CH + k] * vis;
                cur_idx = batch_start + t;
                T = next_T;
            }
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

// Sum 8 per-lane values over the warp with a transposing butterfly: after the call, lanes 4s..4s+3 all hold the
// warp total of value s (s = 0..7).  4+2+1+1+1 = 9 shuffles.
__device__ __forceinline__ float warp_reduce8(float v[8], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float send = b4 ? v[k] : v[k + 4];
        float keep = b4 ? v[k + 4] : v[k];
        a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float c[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        float send = b3 ? a[k] : a[k + 2];
        float keep = b3 ? a[k + 2] : a[k];
        c[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float send = b2 ? c[0] : c[1];
    float keep = b2 ? c[1] : c[0];
    float r = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;  // lane holds value index ((b4?4:0) + (b3?2:0) + (b2?1:0))
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
    __shared__ float4 s_k[BLOCK];   // x, y, hx, hy
    __shared__ float4 s_q[BLOCK];   // ca, cb, cc, opac
    __shared__ float s_rgb[BLOCK * CH];
    __shared__ int s_max[WARPS];

    const int tile_id = blockIdx.y * tile_w + blockIdx.x;
    const int tr = threadIdx.x;
    const int lane = tr & 31, warp = tr >> 5;
    int lx, ly;
    thread_pixel(tr, lx, ly);
    const int i = blockIdx.y * GSB_TILE + ly;
    const int j = blockIdx.x * GSB_TILE + lx;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < H && j < W);
    const size_t pix = inside ? (size_t)i * W + j : 0;
    const float rcx = (float)(blockIdx.x * GSB_TILE + (warp & 1) * SUB_W) + 0.5f * SUB_W;
    const float rcy = (float)(blockIdx.y * GSB_TILE + (warp >> 1) * SUB_H) + 0.5f * SUB_H;
    const float rhx = 0.5f * (SUB_W - 1), rhy = 0.5f * (SUB_H - 1);

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
    const int bin_final = inside ? last_ids[pix] : -1;

    int wmax = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    int block_max = s_max[0];
#pragma unroll
    for (int w = 1; w < WARPS; ++w) block_max = max(block_max, s_max[w]);
    const int warp_bin_final = wmax;

    for (int b = 0; b < num_batches; ++b) {
        // batches walk the list from its END: entry t of batch b is list position batch_end - t
        const int batch_end = range_end - 1 - BLOCK * b;
        const int batch_size = min(BLOCK, batch_end + 1 - range_start);
        if (batch_end - batch_size + 1 > block_max) continue;  // nothing in this batch contributed anywhere
        __syncthreads();
        const int idx = batch_end - tr;
        if (idx >= range_start) {
            const int g = flatten_ids[idx];
            s_id[tr] = g;
            const float2 xy = means2d[g];
            const float ca = conics[3 * g], cb = conics[3 * g + 1], cc = conics[3 * g + 2];
            const float op = opacities[g];
            const float2 ext = alpha_extent(ca, cb, cc, op);
            s_k[tr] = make_float4(xy.x, xy.y, ext.x, ext.y);
            s_q[tr] = make_float4(ca, cb, cc, op);
#pragma unroll
            for (int k = 0; k < CH; ++k) s_rgb[tr * CH + k] = colors[(size_t)g * CH + k];
        } else {
            s_k[tr] = make_float4(0.f, 0.f, -1e30f, -1e30f);
        }
        __syncthreads();
        if (batch_end - batch_size + 1 > warp_bin_final) continue;  // warp-uniform
#pragma unroll 1
        for (int c = 0; c < BLOCK / 32; ++c) {
            if (batch_end - (c * 32 + 31) > warp_bin_final) continue;  // whole chunk is behind this warp's pixels
            const float4 kc = s_k[c * 32 + lane];
            const bool hit = (fabsf(kc.x - rcx) <= kc.z + rhx) && (fabsf(kc.y - rcy) <= kc.w + rhy) &&
                             (batch_end - (c * 32 + lane) <= warp_bin_final);
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                const int t = c * 32 + __ffs(m) - 1;   // ascending t == descending list position
                m &= m - 1;
                const float4 kk = s_k[t];
                const float4 q = s_q[t];
                const float dx = kk.x - px, dy = kk.y - py;
                bool valid = (batch_end - t <= bin_final);
                float alpha = 0.f, vis = 0.f;
                const float sigma = 0.5f * (q.x * dx * dx + q.z * dy * dy) + q.y * dx * dy;
                vis = ex2_approx(-LOG2E * sigma);
                alpha = fminf(GSB_ALPHA_CLAMP, q.w * vis);
                if (sigma < 0.f || alpha < GSB_ALPHA_MIN) valid = false;
                if (!__any_sync(0xffffffffu, valid)) continue;
                float v[8];
                float v_op = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = 0.f;
                float v_rgb_extra[CH > 3 ? CH - 3 : 1];
#pragma unroll
                for (int k = 0; k < (CH > 3 ? CH - 3 : 1); ++k) v_rgb_extra[k] = 0.f;
                if (valid) {
                    const float ra = 1.0f / (1.0f - alpha);
                    T *= ra;
                    const float fac = alpha * T;
                    float v_alpha = 0.f;
#pragma unroll
                    for (int k = 0; k < CH; ++k) {
                        const float cch = s_rgb[t * CH + k];
                        const float g_ = fac * v_out[k];
                        if (k < 3) v[k] = g_; else v_rgb_extra[k - 3] = g_;
                        v_alpha += (cch * T - buffer[k] * ra) * v_out[k];
                        buffer[k] += cch * fac;
                    }
                    v_alpha += T_final * ra * v_a_out;
                    if (background) v_alpha += -T_final * ra * bg_dot;
                    if (q.w * vis <= GSB_ALPHA_CLAMP) {
                        const float v_sigma = -q.w * vis * v_alpha;
                        v[3] = 0.5f * v_sigma * dx * dx;
                        v[4] = v_sigma * dx * dy;
                        v[5] = 0.5f * v_sigma * dy * dy;
                        v[6] = v_sigma * (q.x * dx + q.y * dy);
                        v[7] = v_sigma * (q.y * dx + q.z * dy);
                        v_op = vis * v_alpha;
                    }
                }
                const float r = warp_reduce8(v, lane);
                v_op = warp_sum(v_op);
                const int g = s_id[t];
                if ((lane & 3) == 0) {
                    const int s = lane >> 2;   // value index held by this lane group
                    float *dst;
                    if (s < 3) dst = (s < min(CH, 3)) ? v_colors + (size_t)g * CH + s : nullptr;
                    else if (s < 6) dst = v_conics + 3 * (size_t)g + (s - 3);
                    else dst = v_means2d + 2 * (size_t)g + (s - 6);
                    if (dst) atomicAdd(dst, r);
                } else if (lane == 1) {
                    atomicAdd(v_opacities + g, v_op);
                }
                if (CH > 3) {
#pragma unroll
                    for (int k = 3; k < CH; ++k) {
                        float e = warp_sum(v_rgb_extra[k - 3]);
                        if (lane == 2) atomicAdd(v_colors + (size_t)g * CH + k, e);
                    }
                }
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
    dim3 grid(tw, th), block(BLOCK);
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
    dim3 grid(tw, th), block(BLOCK);
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
