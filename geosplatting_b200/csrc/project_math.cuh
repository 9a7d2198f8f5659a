// EWA projection of one Gaussian: forward pieces shared by the forward kernel (compiled with
// -fmad=false so that means2d / radii / depths -- and therefore every tile/bin index -- are
// bit-identical to the CPU oracle's IEEE arithmetic) and by the backward kernel (which recomputes them).
//
// Replaces gsplat 1.4.0 fully_fused_projection (third-party; semantics per SURVEY.md Appendix C.2),
// reached from rfstudio/model/gsplat.py:334-355.
#pragma once
#include "gsb_common.cuh"

struct ProjOut {
    float R[9];   // rotation of the normalised quaternion
    float M[9];   // R * diag(scale)
    float Sc[9];  // camera-space covariance
    float pc[3];  // camera-space mean
    float J[6];   // 2x3 perspective Jacobian (with the fov clamp)
    float tx, ty, rz;
    bool x_in, y_in;
    float comp;
    float conic[3];
    float mean2d[2];
    float radius;
};

__device__ __forceinline__ void gsb_quat_to_rot(const float *q, float *R, float &inv_norm) {
    float w = q[0], x = q[1], y = q[2], z = q[3];
    float n2 = ((x * x + y * y) + z * z) + w * w;
    inv_norm = 1.0f / sqrtf(n2);
    w *= inv_norm; x *= inv_norm; y *= inv_norm; z *= inv_norm;
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float xy = x * y, xz = x * z, yz = y * z;
    float wx = w * x, wy = w * y, wz = w * z;
    R[0] = 1.0f - 2.0f * (y2 + z2); R[1] = 2.0f * (xy - wz);        R[2] = 2.0f * (xz + wy);
    R[3] = 2.0f * (xy + wz);        R[4] = 1.0f - 2.0f * (x2 + z2); R[5] = 2.0f * (yz - wx);
    R[6] = 2.0f * (xz - wy);        R[7] = 2.0f * (yz + wx);        R[8] = 1.0f - 2.0f * (x2 + y2);
}

// Returns false when the Gaussian is culled.  Operation order is part of the parity contract.
// DECIDE = false (the backward): the forward's verdict (radii > 0) is final -- the recomputation runs in a translation
// unit compiled WITH fma contraction, so a borderline near-plane / radius-clip / screen-edge predicate could disagree
// with the forward's and silently drop a gradient; only the division guard (det > 0) is kept.
template <bool DECIDE = true>
__device__ __forceinline__ bool gsb_project_one(const float *mean, const float *quat, const float *scale,
                                                const CamK &cam, ProjOut &o) {
    const float *Rcw = cam.r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        o.pc[i] = ((Rcw[i * 3 + 0] * mean[0] + Rcw[i * 3 + 1] * mean[1]) + Rcw[i * 3 + 2] * mean[2]) + cam.t[i];
    float x = o.pc[0], y = o.pc[1], z = o.pc[2];
    if (DECIDE && (z < cam.near_plane || z > cam.far_plane)) return false;

    float inv_norm;
    gsb_quat_to_rot(quat, o.R, inv_norm);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o.M[i * 3 + j] = o.R[i * 3 + j] * scale[j];
    float S[9], tmp[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            S[i * 3 + j] = (o.M[i * 3 + 0] * o.M[j * 3 + 0] + o.M[i * 3 + 1] * o.M[j * 3 + 1]) + o.M[i * 3 + 2] * o.M[j * 3 + 2];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            tmp[i * 3 + j] = (Rcw[i * 3 + 0] * S[0 * 3 + j] + Rcw[i * 3 + 1] * S[1 * 3 + j]) + Rcw[i * 3 + 2] * S[2 * 3 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            o.Sc[i * 3 + j] = (tmp[i * 3 + 0] * Rcw[j * 3 + 0] + tmp[i * 3 + 1] * Rcw[j * 3 + 1]) + tmp[i * 3 + 2] * Rcw[j * 3 + 2];

    const float fx = cam.fx, fy = cam.fy, cx = cam.cx, cy = cam.cy;
    const float Wf = (float)cam.W, Hf = (float)cam.H;
    float tan_fovx = 0.5f * Wf / fx, tan_fovy = 0.5f * Hf / fy;
    float lim_x_pos = (Wf - cx) / fx + GSB_FOV_MARGIN * tan_fovx;
    float lim_x_neg = cx / fx + GSB_FOV_MARGIN * tan_fovx;
    float lim_y_pos = (Hf - cy) / fy + GSB_FOV_MARGIN * tan_fovy;
    float lim_y_neg = cy / fy + GSB_FOV_MARGIN * tan_fovy;
    float rz = 1.0f / z;
    float rz2 = rz * rz;
    float xr = x * rz, yr = y * rz;
    o.x_in = (xr <= lim_x_pos && xr >= -lim_x_neg);
    o.y_in = (yr <= lim_y_pos && yr >= -lim_y_neg);
    float tx = z * fminf(lim_x_pos, fmaxf(-lim_x_neg, xr));
    float ty = z * fminf(lim_y_pos, fmaxf(-lim_y_neg, yr));
    o.tx = tx; o.ty = ty; o.rz = rz;
    float *J = o.J;
    J[0] = fx * rz; J[1] = 0.0f; J[2] = -(fx * tx) * rz2;
    J[3] = 0.0f; J[4] = fy * rz; J[5] = -(fy * ty) * rz2;
    const float *Sc = o.Sc;
    float T00 = J[0] * Sc[0] + J[2] * Sc[6], T01 = J[0] * Sc[1] + J[2] * Sc[7], T02 = J[0] * Sc[2] + J[2] * Sc[8];
    float T11 = J[4] * Sc[4] + J[5] * Sc[7], T12 = J[4] * Sc[5] + J[5] * Sc[8];
    float c00 = T00 * J[0] + T02 * J[2];
    float c01 = T01 * J[4] + T02 * J[5];
    float c11 = T11 * J[4] + T12 * J[5];
    o.mean2d[0] = (fx * x) * rz + cx;
    o.mean2d[1] = (fy * y) * rz + cy;

    float det0 = c00 * c11 - c01 * c01;
    c00 += cam.eps2d;
    c11 += cam.eps2d;
    float det = c00 * c11 - c01 * c01;
    o.comp = sqrtf(fmaxf(0.0f, det0 / det));
    if (!(det > 0.0f)) return false;
    o.conic[0] = c11 / det;
    o.conic[1] = -c01 / det;
    o.conic[2] = c00 / det;
    float b = 0.5f * (c00 + c11);
    float v1 = b + sqrtf(fmaxf(GSB_RADIUS_DET_FLOOR, b * b - det));
    float rad = ceilf(3.0f * sqrtf(v1));
    if (DECIDE && rad <= cam.radius_clip) return false;
    if (DECIDE && (o.mean2d[0] + rad <= 0.0f || o.mean2d[0] - rad >= Wf || o.mean2d[1] + rad <= 0.0f ||
        o.mean2d[1] - rad >= Hf))
        return false;
    o.radius = rad;
    return true;
}

// Tile rectangle [x0,x1) x [y0,y1) touched by a projected Gaussian (gsplat isect_tiles).
__device__ __forceinline__ void gsb_tile_range(float mx, float my, int radius, int tile_w, int tile_h, int &x0,
                                               int &x1, int &y0, int &y1) {
    const float ts = (float)GSB_TILE;
    float tr = (float)radius / ts;
    float txf = mx / ts, tyf = my / ts;
    x0 = (int)fminf(fmaxf(0.0f, floorf(txf - tr)), (float)tile_w);
    x1 = (int)fminf(fmaxf(0.0f, ceilf(txf + tr)), (float)tile_w);
    y0 = (int)fminf(fmaxf(0.0f, floorf(tyf - tr)), (float)tile_h);
    y1 = (int)fminf(fmaxf(0.0f, ceilf(tyf + tr)), (float)tile_h);
}
