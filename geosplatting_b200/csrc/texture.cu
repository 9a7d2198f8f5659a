// Stand-alone texture sampling with the call shapes the reference hands to nvdiffrast.torch.texture:
//   2D  [H,W,C]    filter 'linear', boundary 'clamp'                 (rfstudio/model/geosplat.py:93-98)
//   cube [6,R,R,C] filter 'linear', boundary 'cube'                  (rfstudio/graphics/_mesh/_texture.py:220,:596)
//   cube + explicit mip stack + per-sample level, 'linear-mipmap-linear' (_texture.py:604-611)
// One thread per sample; gathers are L2-resident, gradients scatter with red.global.add.f32.
// The fused shade kernels (shade.cu) use the same device functions; these entry points exist so that the
// reference's own Python (TextureSplitSum.sample, _CubeMapMip.backward) can run on this library unchanged.
#include "texture_math.cuh"

namespace {

constexpr int MAX_LEVELS = 16;

struct LevelPtrs {
    float *p[MAX_LEVELS];
};

template <int C>
__device__ __forceinline__ void sample_level(const float *__restrict__ tex, const CubeTaps &t, float out[C],
                                             float d_fu[C], float d_fv[C]) {
    float a[4][C];
    float sum[C];
#pragma unroll
    for (int c = 0; c < C; ++c) sum[c] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            a[k][c] = (t.idx[k] >= 0) ? __ldg(tex + (size_t)t.idx[k] * C + c) : 0.f;
            sum[c] += a[k][c];
        }
    if (t.missing >= 0)
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k == t.missing)
#pragma unroll
                for (int c = 0; c < C; ++c) a[k][c] = sum[c] * 0.33333333f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float top = a[0][c] + (a[1][c] - a[0][c]) * t.fu;
        float bot = a[2][c] + (a[3][c] - a[2][c]) * t.fu;
        out[c] = top + (bot - top) * t.fv;
        d_fu[c] = (a[1][c] - a[0][c]) * (1.f - t.fv) + (a[3][c] - a[2][c]) * t.fv;
        d_fv[c] = bot - top;
    }
}

template <int C>
__device__ __forceinline__ void scatter_level(float *__restrict__ v_tex, const CubeTaps &t, const float v[C]) {
    float w[4];
    gsb_cube_weights(t, w);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (t.idx[k] < 0 || w[k] == 0.f) continue;
#pragma unroll
        for (int c = 0; c < C; ++c) atomicAdd(v_tex + (size_t)t.idx[k] * C + c, w[k] * v[c]);
    }
}

// MODE 0: forward.  MODE 1: backward.
template <int C, int MODE>
__global__ void __launch_bounds__(256) texture_cube_kernel(int N, int n_levels, LevelPtrs tex, int R0,
                                                            const float *__restrict__ dirs,
                                                            const float *__restrict__ level, float *__restrict__ out,
                                                            const float *__restrict__ v_out, LevelPtrs v_tex,
                                                            float *__restrict__ v_dirs, float *__restrict__ v_level) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
    float lv = (level && n_levels > 1) ? fminf(fmaxf(level[i], 0.f), (float)(n_levels - 1)) : 0.f;
    bool lv_clamped = (level && n_levels > 1) ? (lv != level[i]) : true;
    int l0 = min((int)floorf(lv), n_levels - 1);
    int l1 = min(l0 + 1, n_levels - 1);
    float lf = lv - (float)l0;
    CubeTaps t0 = gsb_cube_taps(x, y, z, R0 >> l0), t1 = t0;
    float c0[C], c1[C], d0u[C], d0v[C], d1u[C], d1v[C];
    sample_level<C>(tex.p[l0], t0, c0, d0u, d0v);
    if (l1 != l0) {
        t1 = gsb_cube_taps(x, y, z, R0 >> l1);
        sample_level<C>(tex.p[l1], t1, c1, d1u, d1v);
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) c1[c] = c0[c];
    }
    if (MODE == 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) out[(size_t)i * C + c] = c0[c] + (c1[c] - c0[c]) * lf;
        return;
    }
    float v[C], v0[C], v1[C];
    float vfu0 = 0.f, vfv0 = 0.f, vfu1 = 0.f, vfv1 = 0.f, vl = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        v[c] = v_out[(size_t)i * C + c];
        v0[c] = (l1 != l0) ? v[c] * (1.f - lf) : v[c];
        v1[c] = v[c] * lf;
        vfu0 += v0[c] * d0u[c]; vfv0 += v0[c] * d0v[c];
        vfu1 += v1[c] * d1u[c]; vfv1 += v1[c] * d1v[c];
        vl += v[c] * (c1[c] - c0[c]);
    }
    if (v_tex.p[l0]) scatter_level<C>(v_tex.p[l0], t0, v0);
    float vd[3], vd1[3] = {0.f, 0.f, 0.f};
    gsb_cube_dir_grad(t0.uv, x, y, z, R0 >> l0, vfu0, vfv0, vd);
    if (l1 != l0) {
        if (v_tex.p[l1]) scatter_level<C>(v_tex.p[l1], t1, v1);
        gsb_cube_dir_grad(t1.uv, x, y, z, R0 >> l1, vfu1, vfv1, vd1);
    }
    if (v_dirs) {
        v_dirs[3 * i] = vd[0] + vd1[0];
        v_dirs[3 * i + 1] = vd[1] + vd1[1];
        v_dirs[3 * i + 2] = vd[2] + vd1[2];
    }
    if (v_level) v_level[i] = (l1 != l0 && !lv_clamped) ? vl : 0.f;
}

template <int MODE>
__global__ void __launch_bounds__(256) texture2d_kernel(int N, int W, int H, const float2 *__restrict__ tex,
                                                         const float2 *__restrict__ uv, float2 *__restrict__ out,
                                                         const float2 *__restrict__ v_out, float2 *__restrict__ v_uv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float2 q = uv[i];
    Lut2D s = gsb_lut_sample(tex, W, H, q.x, q.y);
    if (MODE == 0) { out[i] = s.val; return; }
    float2 v = v_out[i];
    v_uv[i] = make_float2(v.x * s.d_u.x + v.y * s.d_u.y, v.x * s.d_v.x + v.y * s.d_v.y);
}

int fill_levels(LevelPtrs &lp, const float *const *ptrs_host, int n) {
    for (int k = 0; k < MAX_LEVELS; ++k) lp.p[k] = (ptrs_host && k < n) ? const_cast<float *>(ptrs_host[k]) : nullptr;
    return 0;
}

}  // namespace

#define GSB_DISPATCH_C(CV, CALL)                                         \
    switch (CV) {                                                        \
        case 1: { constexpr int C_ = 1; CALL; break; }                   \
        case 2: { constexpr int C_ = 2; CALL; break; }                   \
        case 3: { constexpr int C_ = 3; CALL; break; }                   \
        case 4: { constexpr int C_ = 4; CALL; break; }                   \
        default:                                                         \
            gsb_set_error("%s: unsupported channel count %d (1..4)", __func__, CV); \
            return GSB_EINVAL;                                           \
    }

extern "C" __attribute__((visibility("default"))) int gsb_texture_cube_fwd(
    int32_t N, int32_t channels, int32_t n_levels, const float *const *level_ptrs_host, int32_t R0,
    const float *dirs, const float *level, float *out, void *stream) {
    GSB_CHECK_ARG(N >= 0 && n_levels >= 1 && n_levels <= MAX_LEVELS && R0 >= 1 && level_ptrs_host != nullptr);
    GSB_CHECK_ARG((R0 >> (n_levels - 1)) >= 1);
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(dirs && out);
    LevelPtrs tex, none;
    fill_levels(tex, level_ptrs_host, n_levels);
    fill_levels(none, nullptr, 0);
    GSB_DISPATCH_C(channels, (texture_cube_kernel<C_, 0><<<gsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
                                 N, n_levels, tex, R0, dirs, level, out, nullptr, none, nullptr, nullptr)));
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

extern "C" __attribute__((visibility("default"))) int gsb_texture_cube_bwd(
    int32_t N, int32_t channels, int32_t n_levels, const float *const *level_ptrs_host, int32_t R0,
    const float *dirs, const float *level, const float *v_out, float *const *v_level_ptrs_host, float *v_dirs,
    float *v_level, void *stream) {
    GSB_CHECK_ARG(N >= 0 && n_levels >= 1 && n_levels <= MAX_LEVELS && R0 >= 1 && level_ptrs_host != nullptr);
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(dirs && v_out);
    LevelPtrs tex, vtex;
    fill_levels(tex, level_ptrs_host, n_levels);
    fill_levels(vtex, v_level_ptrs_host, v_level_ptrs_host ? n_levels : 0);
    GSB_DISPATCH_C(channels, (texture_cube_kernel<C_, 1><<<gsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
                                 N, n_levels, tex, R0, dirs, level, nullptr, v_out, vtex, v_dirs, v_level)));
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

extern "C" __attribute__((visibility("default"))) int gsb_texture2d_fwd(int32_t N, int32_t width, int32_t height,
                                                                        const float *tex, const float *uv, float *out,
                                                                        void *stream) {
    GSB_CHECK_ARG(N >= 0 && width > 1 && height > 1);
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(tex && uv && out);
    texture2d_kernel<0><<<gsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
        N, width, height, reinterpret_cast<const float2 *>(tex), reinterpret_cast<const float2 *>(uv),
        reinterpret_cast<float2 *>(out), nullptr, nullptr);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

extern "C" __attribute__((visibility("default"))) int gsb_texture2d_bwd(int32_t N, int32_t width, int32_t height,
                                                                        const float *tex, const float *uv,
                                                                        const float *v_out, float *v_uv, void *stream) {
    GSB_CHECK_ARG(N >= 0 && width > 1 && height > 1);
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(tex && uv && v_out && v_uv);
    texture2d_kernel<1><<<gsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
        N, width, height, reinterpret_cast<const float2 *>(tex), reinterpret_cast<const float2 *>(uv), nullptr,
        reinterpret_cast<const float2 *>(v_out), reinterpret_cast<float2 *>(v_uv));
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}
