// Fused per-Gaussian split-sum / Cook-Torrance shade, forward and backward: one thread per Gaussian,
// one pass over the per-Gaussian attributes (56 B in, 12 B out forward), texture taps served from L2
// (FG LUT 512 KB, env stack <= 34 MB), texel gradients scattered with red.global.add.v4.f32 -- those of the coarse
// levels into one of SHADE_REPLICAS private copies (summed afterwards), because half of all Gaussians hit the same
// few thousand texels there and same-address reductions serialise in L2 (measured: 0.13 of 0.24 ms).
// HBM-bound streaming kernel; replaces ~25 elementwise torch kernels + 3 nvdiffrast texture kernels
// per view of rfstudio/model/geosplat.py:83-121 and rfstudio/graphics/_mesh/_texture.py:571-613.
#include "texture_math.cuh"
#include "composite_rec.cuh"

namespace {

struct ShadeParams {
    float cam[3];
    float min_roughness, max_metallic;        // geosplat.py:85-86 (0.1, 1.0)
    float env_min_roughness, env_max_roughness;  // TextureSplitSum.min/max_roughness (0.08, 0.5)
    int mode;                                 // 0 pbr, 1 diffuse, 2 specular
    int lut_w, lut_h;
};

// _texture.py:584-594 (+ the clamp nvdiffrast applies to the level); returns d(level)/d(roughness).
__device__ __forceinline__ float mip_level(float r, const ShadeParams &p, int L, float &dlevel_dr) {
    float level;
    if (r < p.env_max_roughness) {
        float span = p.env_max_roughness - p.env_min_roughness;
        float t = (r - p.env_min_roughness) / span;
        bool in = (t > 0.f && t < 1.f);
        level = fminf(fmaxf(t, 0.f), 1.f) * (float)(L - 2);
        dlevel_dr = in ? (float)(L - 2) / span : 0.f;
    } else {
        float span = 1.0f - p.env_max_roughness;
        float t = (r - p.env_max_roughness) / span;
        bool in = (t > 0.f && t < 1.f);
        level = fminf(fmaxf(t, 0.f), 1.f) + (float)(L - 2);
        dlevel_dr = in ? 1.0f / span : 0.f;
    }
    float lc = fminf(fmaxf(level, 0.f), (float)(L - 1));
    if (lc != level) dlevel_dr = 0.f;
    return lc;
}

constexpr int SHADE_REPLICAS = 32;

// Where texel gradients go: the stack gradient itself, or -- for texels of the hot tail when replicas are in use --
// this CTA's private copy of the tail.
struct EnvGrad {
    float *stack;      // v_env_stack
    float *replica;    // this CTA's copy of the hot tail, or nullptr
    long long hot_begin;
    __device__ __forceinline__ float *level(long long texel_offset) const {
        if (replica && texel_offset >= hot_begin) return replica + 4 * (texel_offset - hot_begin);
        return stack + 4 * texel_offset;
    }
};

struct ShadeFwd {
    float3 color;
    // saved for the backward
    float r, met, len, s;         // roughness, metallic, |cam-m|, wo.n
    float wo[3], refl[3];
    float2 fg;
    float3 l_spec, l_diff, F0, diff;
    bool wo_fallback, ndv_clamped;
};

__device__ __forceinline__ float3 f3mul(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }

template <bool BWD>
__device__ __forceinline__ void shade_one(const float m[3], const float n[3], const float kd[3], const float ks[2],
                                          const ShadeParams &p, const float2 *__restrict__ lut, EnvStack env,
                                          ShadeFwd &o, const float vc[3], float v_m[3], float v_n[3], float v_kd[3],
                                          float v_ks[2], const EnvGrad &v_env) {
    o.r = ks[0] * (1.0f - p.min_roughness) + p.min_roughness;
    o.met = ks[1] * p.max_metallic;
    float omm = 1.0f - o.met;
    o.F0 = make_float3(omm * 0.04f + kd[0] * o.met, omm * 0.04f + kd[1] * o.met, omm * 0.04f + kd[2] * o.met);
    o.diff = make_float3(kd[0] * omm, kd[1] * omm, kd[2] * omm);
    float d[3] = {p.cam[0] - m[0], p.cam[1] - m[1], p.cam[2] - m[2]};
    o.len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    o.wo_fallback = o.len < 1e-6f;
    if (o.wo_fallback) { o.wo[0] = 0.f; o.wo[1] = 0.f; o.wo[2] = 1.f; }
    else { float il = 1.0f / fmaxf(o.len, 1e-6f); o.wo[0] = d[0] * il; o.wo[1] = d[1] * il; o.wo[2] = d[2] * il; }
    o.s = n[0] * o.wo[0] + n[1] * o.wo[1] + n[2] * o.wo[2];
    o.ndv_clamped = !(o.s >= 1e-6f);
    float ndv = fmaxf(o.s, 1e-6f);
    Lut2D fg = gsb_lut_sample(lut, p.lut_w, p.lut_h, ndv, o.r);
    o.fg = fg.val;
    o.refl[0] = 2.f * o.s * n[0] - o.wo[0];
    o.refl[1] = 2.f * o.s * n[1] - o.wo[1];
    o.refl[2] = 2.f * o.s * n[2] - o.wo[2];

    const bool need_spec = (p.mode != 1), need_diff = (p.mode == 1);
    float dlevel_dr = 0.f;
    float level = mip_level(o.r, p, env.L, dlevel_dr);
    int l0 = min((int)floorf(level), env.L - 1);
    int l1 = min(l0 + 1, env.L - 1);
    float lf = level - (float)l0;
    CubeTaps t0, t1, tb;
    float3 c0 = make_float3(0.f, 0.f, 0.f), c1 = c0, d0u = c0, d0v = c0, d1u = c0, d1v = c0, dbu = c0, dbv = c0;
    const float *lvl0 = env.data + 4 * env.level_offset(l0);
    const float *lvl1 = env.data + 4 * env.level_offset(l1);
    const float *base = env.data + 4 * env.base_offset();
    o.l_spec = c0;
    o.l_diff = c0;
    if (need_spec) {
        t0 = gsb_cube_taps(o.refl[0], o.refl[1], o.refl[2], env.R0 >> l0);
        c0 = gsb_cube_sample<4, BWD>(lvl0, t0, &d0u, &d0v);
        c1 = c0;
        if (l1 != l0) {
            t1 = gsb_cube_taps(o.refl[0], o.refl[1], o.refl[2], env.R0 >> l1);
            c1 = gsb_cube_sample<4, BWD>(lvl1, t1, &d1u, &d1v);
        }
        o.l_spec = make_float3(c0.x + (c1.x - c0.x) * lf, c0.y + (c1.y - c0.y) * lf, c0.z + (c1.z - c0.z) * lf);
    }
    if (need_diff) {
        tb = gsb_cube_taps(n[0], n[1], n[2], env.Rb);
        o.l_diff = gsb_cube_sample<4, BWD>(base, tb, &dbu, &dbv);
    }
    float3 reflc = make_float3(o.F0.x * o.fg.x + o.fg.y, o.F0.y * o.fg.x + o.fg.y, o.F0.z * o.fg.x + o.fg.y);
    if (p.mode == 0) {
        float3 sp = f3mul(o.l_spec, reflc);
        o.color = make_float3(o.diff.x + sp.x, o.diff.y + sp.y, o.diff.z + sp.z);
    } else if (p.mode == 1) {
        o.color = f3mul(o.l_diff, o.diff);
    } else {
        o.color = f3mul(o.l_spec, reflc);
    }
    if (!BWD) return;

    // ------------------------------------------------------------------------------------------ backward
    float3 v = make_float3(vc[0], vc[1], vc[2]);
    float3 v_diff = make_float3(0.f, 0.f, 0.f), v_lspec = v_diff, v_reflc = v_diff, v_ldiff = v_diff;
    if (p.mode == 0) { v_diff = v; v_lspec = f3mul(v, reflc); v_reflc = f3mul(v, o.l_spec); }
    else if (p.mode == 1) { v_ldiff = f3mul(v, o.diff); v_diff = f3mul(v, o.l_diff); }
    else { v_lspec = f3mul(v, reflc); v_reflc = f3mul(v, o.l_spec); }
    float3 v_F0 = make_float3(v_reflc.x * o.fg.x, v_reflc.y * o.fg.x, v_reflc.z * o.fg.x);
    float v_fgx = v_reflc.x * o.F0.x + v_reflc.y * o.F0.y + v_reflc.z * o.F0.z;
    float v_fgy = v_reflc.x + v_reflc.y + v_reflc.z;
    v_kd[0] = v_diff.x * omm + v_F0.x * o.met;
    v_kd[1] = v_diff.y * omm + v_F0.y * o.met;
    v_kd[2] = v_diff.z * omm + v_F0.z * o.met;
    float v_met = -(v_diff.x * kd[0] + v_diff.y * kd[1] + v_diff.z * kd[2]) +
                  (v_F0.x * (kd[0] - 0.04f) + v_F0.y * (kd[1] - 0.04f) + v_F0.z * (kd[2] - 0.04f));
    float v_ndv = v_fgx * fg.d_u.x + v_fgy * fg.d_u.y;
    float v_r = v_fgx * fg.d_v.x + v_fgy * fg.d_v.y;
    float v_refl[3] = {0.f, 0.f, 0.f};
    float v_nrm[3] = {0.f, 0.f, 0.f};
    if (need_spec) {
        float3 v0 = make_float3(v_lspec.x * (1.f - lf), v_lspec.y * (1.f - lf), v_lspec.z * (1.f - lf));
        float3 v1 = make_float3(v_lspec.x * lf, v_lspec.y * lf, v_lspec.z * lf);
        float vd[3];
        if (l1 != l0) {
            gsb_cube_scatter<4>(v_env.level(env.level_offset(l0)), t0, v0);
            gsb_cube_scatter<4>(v_env.level(env.level_offset(l1)), t1, v1);
            float vfu1 = v1.x * d1u.x + v1.y * d1u.y + v1.z * d1u.z;
            float vfv1 = v1.x * d1v.x + v1.y * d1v.y + v1.z * d1v.z;
            gsb_cube_dir_grad(t1.uv, o.refl[0], o.refl[1], o.refl[2], env.R0 >> l1, vfu1, vfv1, vd);
            v_refl[0] += vd[0]; v_refl[1] += vd[1]; v_refl[2] += vd[2];
            float v_level = v_lspec.x * (c1.x - c0.x) + v_lspec.y * (c1.y - c0.y) + v_lspec.z * (c1.z - c0.z);
            v_r += v_level * dlevel_dr;
        } else {
            gsb_cube_scatter<4>(v_env.level(env.level_offset(l0)), t0, v_lspec);
            v0 = v_lspec;
        }
        float vfu0 = v0.x * d0u.x + v0.y * d0u.y + v0.z * d0u.z;
        float vfv0 = v0.x * d0v.x + v0.y * d0v.y + v0.z * d0v.z;
        gsb_cube_dir_grad(t0.uv, o.refl[0], o.refl[1], o.refl[2], env.R0 >> l0, vfu0, vfv0, vd);
        v_refl[0] += vd[0]; v_refl[1] += vd[1]; v_refl[2] += vd[2];
    }
    if (need_diff) {
        gsb_cube_scatter<4>(v_env.level(env.base_offset()), tb, v_ldiff);
        float vfu = v_ldiff.x * dbu.x + v_ldiff.y * dbu.y + v_ldiff.z * dbu.z;
        float vfv = v_ldiff.x * dbv.x + v_ldiff.y * dbv.y + v_ldiff.z * dbv.z;
        gsb_cube_dir_grad(tb.uv, n[0], n[1], n[2], env.Rb, vfu, vfv, v_nrm);
    }
    // refl = 2 s n - wo ; s = wo . n
    float v_s = 2.f * (v_refl[0] * n[0] + v_refl[1] * n[1] + v_refl[2] * n[2]);
    if (!o.ndv_clamped) v_s += v_ndv;
    float v_wo[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        v_nrm[k] += 2.f * o.s * v_refl[k] + v_s * o.wo[k];
        v_wo[k] = -v_refl[k] + v_s * n[k];
    }
    if (o.wo_fallback) {
        v_m[0] = v_m[1] = v_m[2] = 0.f;
    } else {
        float dotw = o.wo[0] * v_wo[0] + o.wo[1] * v_wo[1] + o.wo[2] * v_wo[2];
        float il = 1.0f / fmaxf(o.len, 1e-6f);
#pragma unroll
        for (int k = 0; k < 3; ++k) v_m[k] = -(v_wo[k] - o.wo[k] * dotw) * il;
    }
    v_n[0] = v_nrm[0]; v_n[1] = v_nrm[1]; v_n[2] = v_nrm[2];
    v_ks[0] = v_r * (1.0f - p.min_roughness);
    v_ks[1] = v_met * p.max_metallic;
}

// What the compositing record needs besides the colour (composite_rec.cuh); `rec == nullptr`: colours only.
struct PackArgs {
    const float2 *means2d;
    const float *conics, *opacity_logits, *comps;
    GsbRec *rec;
};

#ifndef GSB_SHADE_FWD_MINB
#define GSB_SHADE_FWD_MINB 5   // 48 registers, 0.050 -> 0.047 ms
#endif
__global__ void __launch_bounds__(256, GSB_SHADE_FWD_MINB) shade_fwd_kernel(int N, const float *__restrict__ means,
                                                         const float *__restrict__ normals,
                                                         const float *__restrict__ kd, const float *__restrict__ ks,
                                                         ShadeParams p, const float2 *__restrict__ lut, EnvStack env,
                                                         float *__restrict__ colors, PackArgs pk) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float m[3] = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
    float n[3] = {normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]};
    float k3[3] = {kd[3 * i], kd[3 * i + 1], kd[3 * i + 2]};
    float2 s2 = reinterpret_cast<const float2 *>(ks)[i];
    float k2[2] = {s2.x, s2.y};
    ShadeFwd o;
    shade_one<false>(m, n, k3, k2, p, lut, env, o, nullptr, nullptr, nullptr, nullptr, nullptr, EnvGrad{nullptr, nullptr, 0});
    colors[3 * i] = o.color.x;
    colors[3 * i + 1] = o.color.y;
    colors[3 * i + 2] = o.color.z;
    if (pk.rec)
        pk.rec[i] = gsb_pack_record(pk.means2d[i], pk.conics[3 * i], pk.conics[3 * i + 1], pk.conics[3 * i + 2], o.color.x,
                                    o.color.y, o.color.z, pk.opacity_logits[i], 1, pk.comps ? pk.comps[i] : 1.0f);
}

#ifndef GSB_SHADE_BWD_MINB
#define GSB_SHADE_BWD_MINB 3   // 80 registers (172 B of spills): latency-bound kernel, +50 % resident warps wins 3 %
#endif
__global__ void __launch_bounds__(256, GSB_SHADE_BWD_MINB) shade_bwd_kernel(int N, const float *__restrict__ means,
                                                         const float *__restrict__ normals,
                                                         const float *__restrict__ kd, const float *__restrict__ ks,
                                                         ShadeParams p, const float2 *__restrict__ lut, EnvStack env,
                                                         const float *__restrict__ v_colors,
                                                         float *__restrict__ v_means, float *__restrict__ v_normals,
                                                         float *__restrict__ v_kd, float *__restrict__ v_ks,
                                                         float *__restrict__ v_env_stack, float *__restrict__ replicas,
                                                         long long hot_begin, long long hot_texels, int accumulate) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float vc0 = v_colors[3 * i], vc1 = v_colors[3 * i + 1], vc2 = v_colors[3 * i + 2];
    if (vc0 == 0.f && vc1 == 0.f && vc2 == 0.f) {
        // linear in the colour cotangent: a Gaussian no pixel's gradient reached (culled, occluded -- about half of a
        // closed mesh) adds nothing; in a batch (accumulate) not even its outputs are touched
        if (!accumulate) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { v_means[3 * i + k] = 0.f; v_normals[3 * i + k] = 0.f; v_kd[3 * i + k] = 0.f; }
            reinterpret_cast<float2 *>(v_ks)[i] = make_float2(0.f, 0.f);
        }
        return;
    }
    EnvGrad v_env{v_env_stack, replicas ? replicas + 4 * hot_texels * (blockIdx.x % SHADE_REPLICAS) : nullptr, hot_begin};
    float m[3] = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
    float n[3] = {normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]};
    float k3[3] = {kd[3 * i], kd[3 * i + 1], kd[3 * i + 2]};
    float2 s2 = reinterpret_cast<const float2 *>(ks)[i];
    float k2[2] = {s2.x, s2.y};
    float vc[3] = {vc0, vc1, vc2};
    float vm[3], vn[3], vkd[3], vks[2];
    ShadeFwd o;
    shade_one<true>(m, n, k3, k2, p, lut, env, o, vc, vm, vn, vkd, vks, v_env);
    if (accumulate) {   // several views of a batch (and the projection's v_means) collect in one buffer
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            vm[k] += v_means[3 * i + k];
            vn[k] += v_normals[3 * i + k];
            vkd[k] += v_kd[3 * i + k];
        }
        float2 k0 = reinterpret_cast<const float2 *>(v_ks)[i];
        vks[0] += k0.x; vks[1] += k0.y;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        v_means[3 * i + k] = vm[k];
        v_normals[3 * i + k] = vn[k];
        v_kd[3 * i + k] = vkd[k];
    }
    reinterpret_cast<float2 *>(v_ks)[i] = make_float2(vks[0], vks[1]);
}

// v_env_stack[hot tail] += sum of the replicas (runs after shade_bwd_kernel on the same stream; sole writer).
__global__ void __launch_bounds__(256) shade_replica_sum_kernel(long long hot_texels, const float4 *__restrict__ replicas,
                                                                 float4 *__restrict__ v_hot) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= hot_texels) return;
    float4 s = v_hot[t];
#pragma unroll 8
    for (int r = 0; r < SHADE_REPLICAS; ++r) {
        float4 v = replicas[r * hot_texels + t];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    v_hot[t] = s;
}

// ---- env-stack <-> reference layouts -----------------------------------------------------------------
// TextureSplitSum.mipmaps is a quad-tree pack [6,4,R,R] (rfstudio/graphics/_mesh/_texture.py:228-261):
// level 0 RGB in channels 0..2; level l>=1 R,G,B planes in three quadrants of channel 3, recursively.
__device__ __forceinline__ long long packed_index(int R0, int l, int face, int y, int x, int ch) {
    if (l == 0) return (((long long)face * 4 + ch) * R0 + y) * R0 + x;
    int o = 0, R = R0;
    for (int k = 1; k < l; ++k) { o += R / 2; R /= 2; }
    int h = R / 2;  // size of level l
    int oy = o + (ch == 2 ? h : 0), ox = o + (ch == 1 ? h : 0);
    return (((long long)face * 4 + 3) * R0 + (oy + y)) * R0 + (ox + x);
}

// dir 0: packed/base -> stack (forward);  dir 1: stack gradient -> packed/base gradients (overwrite).
__global__ void __launch_bounds__(256) envstack_convert_kernel(EnvStack env, float *__restrict__ stack,
                                                                float *__restrict__ packed, float *__restrict__ base,
                                                                int dir) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = env.total_texels();
    if (t >= total) return;
    long long boff = env.base_offset();
    float4 *st = reinterpret_cast<float4 *>(stack);
    if (t >= boff) {
        long long k = t - boff;
        if (dir == 0) st[t] = make_float4(base[3 * k], base[3 * k + 1], base[3 * k + 2], 0.f);
        else { float4 v = st[t]; base[3 * k] = v.x; base[3 * k + 1] = v.y; base[3 * k + 2] = v.z; }
        return;
    }
    int l = 0;
    long long rem = t;
    for (;; ++l) { long long r = env.R0 >> l; long long n = 6 * r * r; if (rem < n) break; rem -= n; }
    int R = env.R0 >> l;
    int face = (int)(rem / ((long long)R * R));
    int y = (int)((rem / R) % R), x = (int)(rem % R);
    long long i0 = packed_index(env.R0, l, face, y, x, 0), i1 = packed_index(env.R0, l, face, y, x, 1),
              i2 = packed_index(env.R0, l, face, y, x, 2);
    if (dir == 0) st[t] = make_float4(packed[i0], packed[i1], packed[i2], 0.f);
    else { float4 v = st[t]; packed[i0] = v.x; packed[i1] = v.y; packed[i2] = v.z; }
}

int fill_params(ShadeParams &p, const float *cam_pos_host, float min_roughness, float max_metallic,
                float env_min_roughness, float env_max_roughness, int mode, int lut_res) {
    p.cam[0] = cam_pos_host[0]; p.cam[1] = cam_pos_host[1]; p.cam[2] = cam_pos_host[2];
    p.min_roughness = min_roughness; p.max_metallic = max_metallic;
    p.env_min_roughness = env_min_roughness; p.env_max_roughness = env_max_roughness;
    p.mode = mode; p.lut_w = lut_res; p.lut_h = lut_res;
    return 0;
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int gsb_envstack_texels(int32_t R0, int32_t L, int32_t Rb,
                                                                          int64_t *texels_host) {
    GSB_CHECK_ARG(R0 > 0 && L >= 2 && Rb > 0 && (R0 >> (L - 1)) >= 1 && texels_host != nullptr);
    EnvStack e{nullptr, R0, L, Rb};
    *texels_host = e.total_texels();
    return GSB_OK;
}

extern "C" __attribute__((visibility("default"))) int gsb_envstack_pack(int32_t R0, int32_t L, int32_t Rb,
                                                                        const float *packed, const float *base,
                                                                        float *stack, void *stream) {
    GSB_CHECK_ARG(R0 > 0 && L >= 2 && Rb > 0 && packed && base && stack);
    EnvStack e{stack, R0, L, Rb};
    long long total = e.total_texels();
    envstack_convert_kernel<<<gsb_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
        e, stack, const_cast<float *>(packed), const_cast<float *>(base), 0);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

extern "C" __attribute__((visibility("default"))) int gsb_envstack_unpack_grad(int32_t R0, int32_t L, int32_t Rb,
                                                                               const float *v_stack, float *v_packed,
                                                                               float *v_base, void *stream) {
    GSB_CHECK_ARG(R0 > 0 && L >= 2 && Rb > 0 && v_stack && v_packed && v_base);
    EnvStack e{v_stack, R0, L, Rb};
    long long total = e.total_texels();
    envstack_convert_kernel<<<gsb_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
        e, const_cast<float *>(v_stack), v_packed, v_base, 1);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

// The shade forward, optionally also packing the compositing records (batch driver): `rec` = gsb_composite_records(ws).
int gsb_shade_fwd_impl(int32_t N, const float *means, const float *normals, const float *kd, const float *ks,
                       const float *cam_pos_host, const float *fg_lut, int32_t lut_res, const float *env_stack, int32_t R0,
                       int32_t L, int32_t Rb, float min_roughness, float max_metallic, float env_min_roughness,
                       float env_max_roughness, int32_t mode, float *colors, const float *means2d, const float *conics,
                       const float *opacity_logits, const float *comps, void *rec, void *stream) {
    GSB_CHECK_ARG(N >= 0 && cam_pos_host && mode >= 0 && mode <= 2 && lut_res > 1 && L >= 2 && R0 > 0 && Rb > 0);
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(means && normals && kd && ks && fg_lut && env_stack && colors);
    GSB_CHECK_ARG(rec == nullptr || (means2d && conics && opacity_logits));
    ShadeParams p;
    fill_params(p, cam_pos_host, min_roughness, max_metallic, env_min_roughness, env_max_roughness, mode, lut_res);
    EnvStack e{env_stack, R0, L, Rb};
    PackArgs pk{reinterpret_cast<const float2 *>(means2d), conics, opacity_logits, comps, reinterpret_cast<GsbRec *>(rec)};
    shade_fwd_kernel<<<gsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
        N, means, normals, kd, ks, p, reinterpret_cast<const float2 *>(fg_lut), e, colors, pk);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

extern "C" __attribute__((visibility("default"))) int gsb_shade_fwd(
    int32_t N, const float *means, const float *normals, const float *kd, const float *ks, const float *cam_pos_host,
    const float *fg_lut, int32_t lut_res, const float *env_stack, int32_t R0, int32_t L, int32_t Rb,
    float min_roughness, float max_metallic, float env_min_roughness, float env_max_roughness, int32_t mode,
    float *colors, void *stream) {
    return gsb_shade_fwd_impl(N, means, normals, kd, ks, cam_pos_host, fg_lut, lut_res, env_stack, R0, L, Rb, min_roughness,
                              max_metallic, env_min_roughness, env_max_roughness, mode, colors, nullptr, nullptr, nullptr,
                              nullptr, nullptr, stream);
}

extern "C" __attribute__((visibility("default"))) int gsb_shade_bwd(
    int32_t N, const float *means, const float *normals, const float *kd, const float *ks, const float *cam_pos_host,
    const float *fg_lut, int32_t lut_res, const float *env_stack, int32_t R0, int32_t L, int32_t Rb,
    float min_roughness, float max_metallic, float env_min_roughness, float env_max_roughness, int32_t mode,
    const float *v_colors, float *v_means, float *v_normals, float *v_kd, float *v_ks, float *v_env_stack,
    void *workspace, size_t workspace_bytes, int32_t accumulate, void *stream) {
    GSB_CHECK_ARG(N >= 0 && cam_pos_host && mode >= 0 && mode <= 2 && lut_res > 1 && L >= 2 && R0 > 0 && Rb > 0);
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(means && normals && kd && ks && fg_lut && env_stack && v_colors);
    GSB_CHECK_ARG(v_means && v_normals && v_kd && v_ks && v_env_stack);
    EnvStack es{env_stack, R0, L, Rb};
    const long long hot_begin = es.hot_begin(), hot_texels = es.total_texels() - hot_begin;
    const size_t need = sizeof(float4) * (size_t)hot_texels * SHADE_REPLICAS;
    float *replicas = nullptr;
    if (workspace != nullptr) {
        if (workspace_bytes < need) {
            gsb_set_error("gsb_shade_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
            return GSB_ENOMEM;
        }
        GSB_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0);
        replicas = reinterpret_cast<float *>(workspace);
        GSB_CHECK_CUDA(cudaMemsetAsync(replicas, 0, need, (cudaStream_t)stream));
    }
    ShadeParams p;
    fill_params(p, cam_pos_host, min_roughness, max_metallic, env_min_roughness, env_max_roughness, mode, lut_res);
    EnvStack e{env_stack, R0, L, Rb};
    shade_bwd_kernel<<<gsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
        N, means, normals, kd, ks, p, reinterpret_cast<const float2 *>(fg_lut), e, v_colors, v_means, v_normals,
        v_kd, v_ks, v_env_stack, replicas, hot_begin, hot_texels, accumulate);
    GSB_CHECK_LAUNCH();
    if (replicas) {
        shade_replica_sum_kernel<<<gsb_div_up(hot_texels, 256), 256, 0, (cudaStream_t)stream>>>(
            hot_texels, reinterpret_cast<const float4 *>(replicas), reinterpret_cast<float4 *>(v_env_stack) + hot_begin);
        GSB_CHECK_LAUNCH();
    }
    return GSB_OK;
}

extern "C" __attribute__((visibility("default"))) int gsb_shade_workspace_bytes(int32_t R0, int32_t L, int32_t Rb,
                                                                                size_t *bytes_host) {
    GSB_CHECK_ARG(R0 > 0 && L >= 2 && Rb > 0 && bytes_host != nullptr);
    EnvStack es{nullptr, R0, L, Rb};
    *bytes_host = sizeof(float4) * (size_t)(es.total_texels() - es.hot_begin()) * SHADE_REPLICAS;
    return GSB_OK;
}
