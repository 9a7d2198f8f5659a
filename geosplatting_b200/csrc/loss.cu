// Per-view training loss that consumes the hot path's image, forward and backward fused with the random-background
// composite: rfstudio/trainer/geosplat_trainer.py:171-180 --
//     img1 = rgb + (1 - alpha) * bg,  img2 = gt_rgb * mask + (1 - mask) * bg,
//     loss = lambda * (1 - SSIM(img2, img1)) + (1 - lambda) * mean|img1 - img2| + c_mask * mean((mask - alpha)^2)
// with SSIML1Loss of rfstudio/loss/photometric_loss.py:72-112 and torchmetrics' SSIM (Gaussian 11x11, sigma 1.5,
// k1 0.01, k2 0.03, data_range 1, border of 5 cropped, variances clamped at 0) [third-party restatement, see
// oracle/loss.py].  The reference runs this as ~40 torch kernels and five 11x11 depthwise convolutions per view on
// [H,W] images; here it is two kernels (separable convolutions on shared-memory tiles) that read the rendered RGBA, the
// ground truth and the background once each and write the image cotangent the splat backward consumes.
// HBM-bound: forward 44 B in + 36 B of derivative maps out per pixel, backward 80 B in + 16 B out.
#include "gsb_common.cuh"

namespace {

constexpr int TILE = 16, HALO = 5, KS = 11, EXT = TILE + 2 * HALO;   // 26
constexpr float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;

struct Window {
    float g[KS];
};

Window make_window() {
    Window w;
    double s = 0.0, v[KS];
    for (int i = 0; i < KS; ++i) {
        double d = (double)i - 5.0;
        v[i] = exp(-(d / 1.5) * (d / 1.5) / 2.0);
        s += v[i];
    }
    for (int i = 0; i < KS; ++i) w.g[i] = (float)(v[i] / s);
    return w;
}

__device__ __forceinline__ void compose(const float4 *__restrict__ rgba, const float4 *__restrict__ gt,
                                        const float *__restrict__ bg, int H, int W, int y, int x, float img1[3],
                                        float img2[3]) {
    if (y < 0 || y >= H || x < 0 || x >= W) {
        img1[0] = img1[1] = img1[2] = img2[0] = img2[1] = img2[2] = 0.f;
        return;
    }
    const size_t p = (size_t)y * W + x;
    const float4 r = rgba[p], t = gt[p];
    const float b[3] = {bg[3 * p], bg[3 * p + 1], bg[3 * p + 2]};
    const float rc[3] = {r.x, r.y, r.z}, tc[3] = {t.x, t.y, t.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        img1[c] = rc[c] + (1.0f - r.w) * b[c];
        img2[c] = tc[c] * t.w + (1.0f - t.w) * b[c];
    }
}

__device__ __forceinline__ float block_sum(float v, float *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int tid = threadIdx.y * TILE + threadIdx.x;
    __syncthreads();
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (tid == 0)
        for (int w = 0; w < (TILE * TILE) / 32; ++w) t += s_red[w];
    return t;   // valid in thread 0
}

// sums[0] += sum of the SSIM map over the interior (3 channels), sums[1] += sum |img1 - img2|, sums[2] += sum (mask - alpha)^2;
// maps[c][k][H][W], k = 0..2: d m / d mu1, d m / d E[x x], d m / d E[x y]  (x = img1, the rendered image; 0 outside the interior)
__global__ void __launch_bounds__(TILE *TILE) loss_fwd_kernel(int H, int W, Window win, const float4 *__restrict__ rgba,
                                                              const float4 *__restrict__ gt, const float *__restrict__ bg,
                                                              float *__restrict__ sums, float *__restrict__ maps) {
    __shared__ float s_x[3][EXT][EXT + 1], s_y[3][EXT][EXT + 1];
    __shared__ float s_h[5][EXT][TILE + 1];
    __shared__ float s_red[8];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TILE + tx;
    const int x0 = blockIdx.x * TILE, y0 = blockIdx.y * TILE;
    for (int i = tid; i < EXT * EXT; i += TILE * TILE) {
        const int ly = i / EXT, lx = i % EXT;
        float a[3], b[3];
        compose(rgba, gt, bg, H, W, y0 + ly - HALO, x0 + lx - HALO, a, b);
#pragma unroll
        for (int c = 0; c < 3; ++c) { s_x[c][ly][lx] = a[c]; s_y[c][ly][lx] = b[c]; }
    }
    __syncthreads();
    const int py = y0 + ty, px = x0 + tx;
    const bool in_img = (py < H && px < W);
    const bool interior = in_img && py >= HALO && py < H - HALO && px >= HALO && px < W - HALO;
    float ssim_sum = 0.f, l1_sum = 0.f, mask_sum = 0.f;
    if (in_img) {
#pragma unroll
        for (int c = 0; c < 3; ++c) l1_sum += fabsf(s_x[c][ty + HALO][tx + HALO] - s_y[c][ty + HALO][tx + HALO]);
        const size_t p = (size_t)py * W + px;
        const float d = gt[p].w - rgba[p].w;
        mask_sum = d * d;
    }
    for (int c = 0; c < 3; ++c) {
        __syncthreads();
        // horizontal pass: EXT rows x TILE columns, five quantities
        for (int i = tid; i < EXT * TILE; i += TILE * TILE) {
            const int ly = i / TILE, lx = i % TILE;
            float h[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const float a = s_x[c][ly][lx + k], b = s_y[c][ly][lx + k], g = win.g[k];
                h[0] += g * a; h[1] += g * b; h[2] += g * a * a; h[3] += g * b * b; h[4] += g * a * b;
            }
#pragma unroll
            for (int q = 0; q < 5; ++q) s_h[q][ly][lx] = h[q];
        }
        __syncthreads();
        float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const float g = win.g[k];
#pragma unroll
            for (int q = 0; q < 5; ++q) v[q] += g * s_h[q][ty + k][tx];
        }
        float d_mu1 = 0.f, d_xx = 0.f, d_xy = 0.f;
        if (interior) {
            const float mu1 = v[0], mu2 = v[1];
            const float r11 = v[2] - mu1 * mu1, r22 = v[3] - mu2 * mu2;
            const float s11 = fmaxf(r11, 0.f), s22 = fmaxf(r22, 0.f), s12 = v[4] - mu1 * mu2;
            const float A1 = 2.f * mu1 * mu2 + C1, A2 = 2.f * s12 + C2;
            const float B1 = mu1 * mu1 + mu2 * mu2 + C1, B2 = s11 + s22 + C2;
            const float inv = 1.0f / (B1 * B2);
            const float m = A1 * A2 * inv;
            ssim_sum += m;
            const float dm_ds11 = (r11 > 0.f) ? -m / B2 : 0.f;
            const float dm_ds12 = 2.f * A1 * inv;
            const float dm_dmu1_fixed = 2.f * mu2 * A2 * inv - m * (2.f * mu1) / B1;
            d_mu1 = dm_dmu1_fixed + dm_ds11 * (-2.f * mu1) + dm_ds12 * (-mu2);
            d_xx = dm_ds11;
            d_xy = dm_ds12;
        }
        if (in_img) {
            const size_t plane = (size_t)H * W, p = (size_t)py * W + px;
            maps[(c * 3 + 0) * plane + p] = d_mu1;
            maps[(c * 3 + 1) * plane + p] = d_xx;
            maps[(c * 3 + 2) * plane + p] = d_xy;
        }
    }
    float t0 = block_sum(ssim_sum, s_red);
    if (tid == 0) atomicAdd(sums, t0);
    float t1 = block_sum(l1_sum, s_red);
    if (tid == 0) atomicAdd(sums + 1, t1);
    float t2 = block_sum(mask_sum, s_red);
    if (tid == 0) atomicAdd(sums + 2, t2);
}

__global__ void loss_finalize_kernel(int H, int W, float ssim_lambda, float mask_coeff, float *__restrict__ sums) {
    const float n_ssim = 3.0f * (float)(H - 2 * HALO) * (float)(W - 2 * HALO), n_pix = (float)H * (float)W;
    sums[3] = ssim_lambda * (1.0f - sums[0] / n_ssim) + (1.0f - ssim_lambda) * sums[1] / (3.0f * n_pix) +
              mask_coeff * sums[2] / n_pix;
}

// v_rgba = v_loss * d loss / d rgba
__global__ void __launch_bounds__(TILE *TILE) loss_bwd_kernel(int H, int W, Window win, const float4 *__restrict__ rgba,
                                                              const float4 *__restrict__ gt, const float *__restrict__ bg,
                                                              const float *__restrict__ maps, float ssim_lambda,
                                                              float mask_coeff, const float *__restrict__ v_loss,
                                                              float4 *__restrict__ v_rgba) {
    __shared__ float s_m[3][EXT][EXT + 1];
    __shared__ float s_h[3][EXT][TILE + 1];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TILE + tx;
    const int x0 = blockIdx.x * TILE, y0 = blockIdx.y * TILE;
    const int py = y0 + ty, px = x0 + tx;
    const bool in_img = (py < H && px < W);
    const size_t plane = (size_t)H * W;
    float img1[3] = {0.f, 0.f, 0.f}, img2[3] = {0.f, 0.f, 0.f};
    if (in_img) compose(rgba, gt, bg, H, W, py, px, img1, img2);
    const float n_ssim = 3.0f * (float)(H - 2 * HALO) * (float)(W - 2 * HALO);
    const float n_pix = (float)H * (float)W;
    float v1[3];
    for (int c = 0; c < 3; ++c) {
        __syncthreads();
        for (int i = tid; i < EXT * EXT; i += TILE * TILE) {
            const int ly = i / EXT, lx = i % EXT;
            const int gy = y0 + ly - HALO, gx = x0 + lx - HALO;
            const bool ok = (gy >= 0 && gy < H && gx >= 0 && gx < W);
            const size_t p = ok ? (size_t)gy * W + gx : 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) s_m[k][ly][lx] = ok ? maps[(c * 3 + k) * plane + p] : 0.f;
        }
        __syncthreads();
        for (int i = tid; i < EXT * TILE; i += TILE * TILE) {
            const int ly = i / TILE, lx = i % TILE;
            float h[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const float g = win.g[k];
#pragma unroll
                for (int q = 0; q < 3; ++q) h[q] += g * s_m[q][ly][lx + k];
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) s_h[q][ly][lx] = h[q];
        }
        __syncthreads();
        float v[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const float g = win.g[k];
#pragma unroll
            for (int q = 0; q < 3; ++q) v[q] += g * s_h[q][ty + k][tx];
        }
        // the window is symmetric: correlation with the derivative maps is the transpose of the forward convolution
        const float d_ssim = v[0] + 2.f * img1[c] * v[1] + img2[c] * v[2];
        const float diff = img1[c] - img2[c];
        const float sgn = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
        v1[c] = -ssim_lambda * d_ssim / n_ssim + (1.0f - ssim_lambda) * sgn / (3.0f * n_pix);
    }
    if (!in_img) return;
    const size_t p = (size_t)py * W + px;
    const float s = __ldg(v_loss);
    const float alpha = rgba[p].w, mask = gt[p].w;
    const float v_alpha = -(v1[0] * bg[3 * p] + v1[1] * bg[3 * p + 1] + v1[2] * bg[3 * p + 2]) +
                          mask_coeff * 2.f * (alpha - mask) / n_pix;
    v_rgba[p] = make_float4(s * v1[0], s * v1[1], s * v1[2], s * v_alpha);
}

}  // namespace

#define GSB_API extern "C" __attribute__((visibility("default")))

GSB_API int gsb_loss_fwd(int32_t H, int32_t W, const float *rgba, const float *gt_rgba, const float *bg,
                         float ssim_lambda, float mask_coeff, float *sums4, float *maps, void *stream) {
    GSB_CHECK_ARG(H > 2 * HALO && W > 2 * HALO && rgba && gt_rgba && bg && sums4 && maps);
    cudaStream_t st = (cudaStream_t)stream;
    float *sums3 = sums4;
    GSB_CHECK_CUDA(cudaMemsetAsync(sums4, 0, 4 * sizeof(float), st));
    dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE), block(TILE, TILE);
    loss_fwd_kernel<<<grid, block, 0, st>>>(H, W, make_window(), reinterpret_cast<const float4 *>(rgba),
                                            reinterpret_cast<const float4 *>(gt_rgba), bg, sums3, maps);
    loss_finalize_kernel<<<1, 1, 0, st>>>(H, W, ssim_lambda, mask_coeff, sums4);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_loss_bwd(int32_t H, int32_t W, const float *rgba, const float *gt_rgba, const float *bg,
                         const float *maps, float ssim_lambda, float mask_coeff, const float *v_loss, float *v_rgba,
                         void *stream) {
    GSB_CHECK_ARG(H > 2 * HALO && W > 2 * HALO && rgba && gt_rgba && bg && maps && v_loss && v_rgba);
    dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE), block(TILE, TILE);
    loss_bwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
        H, W, make_window(), reinterpret_cast<const float4 *>(rgba), reinterpret_cast<const float4 *>(gt_rgba), bg, maps,
        ssim_lambda, mask_coeff, v_loss, reinterpret_cast<float4 *>(v_rgba));
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}
