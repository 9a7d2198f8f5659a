// Device-side texture sampling used by the fused shade kernels and by the nvdiffrast-compatible texture
// entry points: 2D bilinear with clamp, cube-map bilinear with seamless cross-face wrap and 3-texel
// corners, trilinear over an explicit mip stack.  All filtering is fp32 in registers (the hardware
// texture unit's 1.8 fixed-point weights would break the 1e-4 parity budget on HDR maps).
//
// Replaces nvdiffrast.torch.texture (third-party, unpinned) as called at
// rfstudio/model/geosplat.py:93-98 and rfstudio/graphics/_mesh/_texture.py:220,:596,:604.
#pragma once
#include "gsb_common.cuh"

// ---- cube-map indexing ------------------------------------------------------------------------------
// Face order +x,-x,+y,-y,+z,-z; (u,v) in [0,1]; inverse of rfstudio/graphics/_mesh/_texture.py:178-197.
struct CubeUV {
    int face;
    float u, v;
    // d(u,v)/d(dir): u = 0.5 + su * a / (2|c|), v = 0.5 + sv * b / (2|c|)
    int ia, ib, ic;   // component indices of a (u-slot), b (v-slot), c (major axis)
    float su, sv;     // +-1
};

__device__ __forceinline__ CubeUV gsb_cube_uv(float x, float y, float z) {
    CubeUV r;
    float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
    float c, a, b;
    if (az > fmaxf(ax, ay)) { r.face = 4; c = z; a = x; b = y; r.ia = 0; r.ib = 1; r.ic = 2; }
    else if (ay > ax)       { r.face = 2; c = y; a = x; b = z; r.ia = 0; r.ib = 2; r.ic = 1; }
    else                    { r.face = 0; c = x; a = z; b = y; r.ia = 2; r.ib = 1; r.ic = 0; }
    if (c < 0.f) r.face += 1;
    float m = 0.5f / fabsf(c);
    r.su = (r.face == 0 || r.face == 5) ? -1.f : 1.f;
    r.sv = (r.face == 2) ? 1.f : -1.f;
    r.u = fminf(fmaxf(a * (r.su * m) + 0.5f, 0.f), 1.f);
    r.v = fminf(fmaxf(b * (r.sv * m) + 0.5f, 0.f), 1.f);
    return r;
}

// Point on (the plane of) face f at face coordinates (gx, gy).
__device__ __forceinline__ void gsb_face_point(int f, float gx, float gy, float &px, float &py, float &pz) {
    switch (f) {
        case 0: px = 1.f;  py = -gy; pz = -gx; break;
        case 1: px = -1.f; py = -gy; pz = gx;  break;
        case 2: px = gx;   py = 1.f; pz = gy;  break;
        case 3: px = gx;   py = -1.f; pz = -gy; break;
        case 4: px = gx;   py = -gy; pz = 1.f; break;
        default: px = -gx; py = -gy; pz = -1.f; break;
    }
}

// Texel (face, iu, iv) with iu / iv possibly -1 or R -> linear texel index inside the level
// ((face*R + iv)*R + iu), or -1 when the tap leaves over two edges at once (cube corner).
__device__ __forceinline__ int gsb_cube_texel(int face, int iu, int iv, int R) {
    bool ou = (iu < 0) | (iu >= R), ov = (iv < 0) | (iv >= R);
    if (!(ou | ov)) return (face * R + iv) * R + iu;
    if (ou & ov) return -1;
    // fold the one-texel overshoot over the cube edge onto the adjacent face
    float inv = 1.0f / (float)R;
    float gx = (2.f * (float)iu + 1.f) * inv - 1.f;
    float gy = (2.f * (float)iv + 1.f) * inv - 1.f;
    float e = inv;  // overshoot is exactly one texel = 2/R * 0.5 in face units -> |g|-1 = 1/R
    float cgx = fminf(fmaxf(gx, -1.f), 1.f), cgy = fminf(fmaxf(gy, -1.f), 1.f);
    float px, py, pz, mx, my, mz;
    gsb_face_point(face, cgx, cgy, px, py, pz);
    gsb_face_point(face, 0.f, 0.f, mx, my, mz);
    px -= mx * e; py -= my * e; pz -= mz * e;
    CubeUV q = gsb_cube_uv(px, py, pz);
    int ju = min(max((int)floorf(q.u * (float)R), 0), R - 1);
    int jv = min(max((int)floorf(q.v * (float)R), 0), R - 1);
    return (q.face * R + jv) * R + ju;
}

struct CubeTaps {
    int idx[4];      // a00 a10 a01 a11 (first index = u); -1 = missing corner texel
    float fu, fv;    // fractional position
    CubeUV uv;
    int missing;     // which tap is the missing corner (-1 none)
};

__device__ __forceinline__ CubeTaps gsb_cube_taps(float x, float y, float z, int R) {
    CubeTaps t;
    t.uv = gsb_cube_uv(x, y, z);
    float u = t.uv.u * (float)R - 0.5f, v = t.uv.v * (float)R - 0.5f;
    float fl_u = floorf(u), fl_v = floorf(v);
    int iu0 = (int)fl_u, iv0 = (int)fl_v;
    t.fu = u - fl_u;
    t.fv = v - fl_v;
    t.idx[0] = gsb_cube_texel(t.uv.face, iu0, iv0, R);
    t.idx[1] = gsb_cube_texel(t.uv.face, iu0 + 1, iv0, R);
    t.idx[2] = gsb_cube_texel(t.uv.face, iu0, iv0 + 1, R);
    t.idx[3] = gsb_cube_texel(t.uv.face, iu0 + 1, iv0 + 1, R);
    t.missing = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (t.idx[k] < 0) t.missing = k;
    return t;
}

// Effective per-texel weights (the missing corner texel is the mean of the other three, factor
// 0.33333333 as in nvdiffrast's fetchQuad).
__device__ __forceinline__ void gsb_cube_weights(const CubeTaps &t, float w[4]) {
    w[0] = (1.f - t.fu) * (1.f - t.fv);
    w[1] = t.fu * (1.f - t.fv);
    w[2] = (1.f - t.fu) * t.fv;
    w[3] = t.fu * t.fv;
    if (t.missing >= 0) {
        // (selects, not w[t.missing]: a register array indexed by a run-time value lives in local memory)
        float wm = (t.missing == 0 ? w[0] : t.missing == 1 ? w[1] : t.missing == 2 ? w[2] : w[3]) * 0.33333333f;
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = (k == t.missing) ? 0.f : w[k] + wm;
    }
}

// Texel fetch with STRIDE floats per texel (3 = packed RGB, 4 = RGBA-padded stack), C channels used.
template <int STRIDE>
__device__ __forceinline__ float3 gsb_fetch3(const float *__restrict__ tex, int idx) {
    if (STRIDE == 4) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(tex) + idx);
        return make_float3(v.x, v.y, v.z);
    }
    const float *p = tex + (size_t)idx * STRIDE;
    return make_float3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}

// Bilinear cube sample; also returns d(out)/d(fu), d(out)/d(fv) when WITH_GRAD.
template <int STRIDE, bool WITH_GRAD>
__device__ __forceinline__ float3 gsb_cube_sample(const float *__restrict__ tex, const CubeTaps &t, float3 *d_fu,
                                                  float3 *d_fv) {
    // four named texels, not an array: "if (k == t.missing) a[k] = ..." is turned into a run-time indexed store by
    // the compiler, which puts the whole array into local memory (48 bytes of stack traffic per sample)
    const float3 zero = make_float3(0.f, 0.f, 0.f);
    float3 a0 = t.idx[0] >= 0 ? gsb_fetch3<STRIDE>(tex, t.idx[0]) : zero;
    float3 a1 = t.idx[1] >= 0 ? gsb_fetch3<STRIDE>(tex, t.idx[1]) : zero;
    float3 a2 = t.idx[2] >= 0 ? gsb_fetch3<STRIDE>(tex, t.idx[2]) : zero;
    float3 a3 = t.idx[3] >= 0 ? gsb_fetch3<STRIDE>(tex, t.idx[3]) : zero;
    if (t.missing >= 0) {
        // same order of additions as a loop over k = 0..3 that skips the missing tap
        float3 sum = zero;
        if (t.idx[0] >= 0) { sum.x += a0.x; sum.y += a0.y; sum.z += a0.z; }
        if (t.idx[1] >= 0) { sum.x += a1.x; sum.y += a1.y; sum.z += a1.z; }
        if (t.idx[2] >= 0) { sum.x += a2.x; sum.y += a2.y; sum.z += a2.z; }
        if (t.idx[3] >= 0) { sum.x += a3.x; sum.y += a3.y; sum.z += a3.z; }
        const float3 third = make_float3(sum.x * 0.33333333f, sum.y * 0.33333333f, sum.z * 0.33333333f);
        const int ms = t.missing;
        a0 = ms == 0 ? third : a0;
        a1 = ms == 1 ? third : a1;
        a2 = ms == 2 ? third : a2;
        a3 = ms == 3 ? third : a3;
    }
    float fu = t.fu, fv = t.fv;
    float3 top = make_float3(a0.x + (a1.x - a0.x) * fu, a0.y + (a1.y - a0.y) * fu, a0.z + (a1.z - a0.z) * fu);
    float3 bot = make_float3(a2.x + (a3.x - a2.x) * fu, a2.y + (a3.y - a2.y) * fu, a2.z + (a3.z - a2.z) * fu);
    if (WITH_GRAD) {
        *d_fu = make_float3((a1.x - a0.x) * (1.f - fv) + (a3.x - a2.x) * fv,
                            (a1.y - a0.y) * (1.f - fv) + (a3.y - a2.y) * fv,
                            (a1.z - a0.z) * (1.f - fv) + (a3.z - a2.z) * fv);
        *d_fv = make_float3(bot.x - top.x, bot.y - top.y, bot.z - top.z);
    }
    return make_float3(top.x + (bot.x - top.x) * fv, top.y + (bot.y - top.y) * fv, top.z + (bot.z - top.z) * fv);
}

// Scatter v (cotangent of the sampled colour, already scaled) into the texel gradients.  For the float4 env stack
// one vector reduction (red.global.add.v4.f32, sm_90+) per tap replaces three scalar ones.
template <int STRIDE>
__device__ __forceinline__ void gsb_cube_scatter(float *__restrict__ v_tex, const CubeTaps &t, float3 v) {
    float w[4];
    gsb_cube_weights(t, w);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (t.idx[k] < 0 || w[k] == 0.f) continue;
        if (STRIDE == 4) {
            atomicAdd(reinterpret_cast<float4 *>(v_tex) + t.idx[k], make_float4(w[k] * v.x, w[k] * v.y, w[k] * v.z, 0.f));
        } else {
            float *p = v_tex + (size_t)t.idx[k] * STRIDE;
            atomicAdd(p, w[k] * v.x);
            atomicAdd(p + 1, w[k] * v.y);
            atomicAdd(p + 2, w[k] * v.z);
        }
    }
}

// Chain (v_fu, v_fv) (cotangents of the fractional texel coordinates) back to the direction.
__device__ __forceinline__ void gsb_cube_dir_grad(const CubeUV &q, float x, float y, float z, int R, float v_fu,
                                                  float v_fv, float v_dir[3]) {
    // (ia, ib, ic) is one of (2,1,0), (0,2,1), (0,1,2): selects instead of run-time indexed arrays (local memory)
    float c = q.ic == 0 ? x : (q.ic == 1 ? y : z);
    float a = q.ia == 0 ? x : z;
    float b = q.ib == 1 ? y : z;
    float ac = fabsf(c);
    float m = 0.5f / ac;
    float Rf = (float)R;
    // clamp(u,0,1): zero gradient when saturated (ties only)
    float gu = (q.u > 0.f && q.u < 1.f) ? v_fu * Rf : 0.f;
    float gv = (q.v > 0.f && q.v < 1.f) ? v_fv * Rf : 0.f;
    // d(1/|c|)/dc = -sign(c)/c^2
    float dm = -0.5f * ((c < 0.f) ? -1.f : 1.f) / (c * c);
    const float ga = gu * q.su * m, gb = gv * q.sv * m, gc = (gu * q.su * a + gv * q.sv * b) * dm;
    // the three indices are distinct: every component receives exactly one of the three terms
    v_dir[0] = q.ia == 0 ? ga : gc;                            // ia is 0 or 2; when it is 2, ic is 0
    v_dir[1] = q.ib == 1 ? gb : gc;                            // ib is 1 or 2; when it is 2, ic is 1
    v_dir[2] = q.ic == 2 ? gc : (q.ia == 2 ? ga : gb);
}

// ---- 2D bilinear, clamp-to-edge (FG LUT: [H,W,2], u -> column, v -> row) -------------------------------
struct Lut2D {
    float2 val;
    float2 d_u, d_v;  // d(val)/du, d(val)/dv in uv units
};

__device__ __forceinline__ Lut2D gsb_lut_sample(const float2 *__restrict__ lut, int W, int H, float u, float v) {
    float U = fminf(fmaxf(u * (float)W - 0.5f, 0.f), (float)W - 1.f);
    float V = fminf(fmaxf(v * (float)H - 0.5f, 0.f), (float)H - 1.f);
    bool cu = (U == 0.f) || (U == (float)W - 1.f);
    bool cv = (V == 0.f) || (V == (float)H - 1.f);
    int iu0 = (int)floorf(U), iv0 = (int)floorf(V);
    int iu1 = iu0 + (cu ? 0 : 1), iv1 = iv0 + (cv ? 0 : 1);
    float fu = U - (float)iu0, fv = V - (float)iv0;
    float2 t00 = __ldg(lut + iv0 * W + iu0), t10 = __ldg(lut + iv0 * W + iu1);
    float2 t01 = __ldg(lut + iv1 * W + iu0), t11 = __ldg(lut + iv1 * W + iu1);
    float2 top = make_float2(t00.x + (t10.x - t00.x) * fu, t00.y + (t10.y - t00.y) * fu);
    float2 bot = make_float2(t01.x + (t11.x - t01.x) * fu, t01.y + (t11.y - t01.y) * fu);
    Lut2D r;
    r.val = make_float2(top.x + (bot.x - top.x) * fv, top.y + (bot.y - top.y) * fv);
    float gu = cu ? 0.f : (float)W, gv = cv ? 0.f : (float)H;
    r.d_u = make_float2(((t10.x - t00.x) * (1.f - fv) + (t11.x - t01.x) * fv) * gu,
                        ((t10.y - t00.y) * (1.f - fv) + (t11.y - t01.y) * fv) * gu);
    r.d_v = make_float2((bot.x - top.x) * gv, (bot.y - top.y) * gv);
    return r;
}

// ---- env-map stack: spec levels 0..L-1 (R0 >> l) followed by the diffuse base (Rb), RGBA texels ---------
struct EnvStack {
    const float *data;  // float4 texels
    int R0, L, Rb;
    __host__ __device__ __forceinline__ long long level_offset(int l) const {  // in texels
        // sum over k < l of 6 (R0 >> k)^2.  A power-of-two R0 whose levels stay >= 1 texel has the closed form
        // 8 (R0^2 - (R0 >> l)^2): the shade kernels ask for three offsets per Gaussian, and the loop was 64-bit
        // arithmetic per level each time.
        if ((R0 & (R0 - 1)) == 0 && (R0 >> l) >= 1) {
            const long long r0 = R0, rl = R0 >> l;
            return 8 * (r0 * r0 - rl * rl);
        }
        long long o = 0;
        for (int k = 0; k < l; ++k) { long long r = R0 >> k; o += 6 * r * r; }
        return o;
    }
    __host__ __device__ __forceinline__ long long base_offset() const { return level_offset(L); }
    __host__ __device__ __forceinline__ long long total_texels() const { return level_offset(L) + 6LL * Rb * Rb; }
    // First texel of the "hot" tail: the coarse levels (resolution <= 32, at most 6144 texels each) and the base.
    // Every Gaussian with a high roughness lands its texel gradients there, hundreds to thousands per texel.
    __host__ __device__ __forceinline__ long long hot_begin() const {
        int l = 0;
        while (l < L && (R0 >> l) > 32) ++l;
        return level_offset(l);
    }
};
