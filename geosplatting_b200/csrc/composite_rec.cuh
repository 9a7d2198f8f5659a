// The packed per-Gaussian record the compositing kernels gather (48 bytes), and how it is made.  Shared by composite.cu
// (pack_records_kernel) and shade.cu (the batch driver's shade forward writes the record itself: the colours never
// make a round trip through HBM and one launch per view disappears).
#pragma once
#include "gsb_common.cuh"

#define GSB_LOG2E 1.4426950408889634f

struct GsbRec {
    float4 k;  // x, y, hx, hy
    float4 q;  // -log2e * (a/2, b, c/2), log2(opacity)
    float4 c;  // r, g, b, opacity
};

// Half extents of {d : 0.5 d^T C d <= tau}, tau = ln(255 * opac): the only region where alpha >= 1/255.
// Negative when the Gaussian can contribute nowhere.  Inflated so that rounding in the per-pixel evaluation can
// never contradict a cull.
__device__ __forceinline__ float2 gsb_alpha_extent(float ca, float cb, float cc, float opac) {
    float t = 255.0f * opac;
    if (!(t > 1.0f)) return make_float2(-1e30f, -1e30f);
    float tau2 = 2.0f * logf(t);
    float det = ca * cc - cb * cb;
    if (!(det > 0.f)) return make_float2(1e30f, 1e30f);  // degenerate conic: never cull
    float hx = sqrtf(tau2 * cc / det), hy = sqrtf(tau2 * ca / det);
    return make_float2(hx * 1.0005f + 0.02f, hy * 1.0005f + 0.02f);
}

// `op` is the raw opacity or its logit; `comp` the antialiasing compensation (1 when there is none)
__device__ __forceinline__ GsbRec gsb_pack_record(float2 xy, float ca, float cb, float cc, float r, float g, float b,
                                                  float op, int opacity_is_logit, float comp) {
    if (opacity_is_logit) op = 1.0f / (1.0f + expf(-op));   // torch.sigmoid (rfstudio/model/gsplat.py:338)
    op *= comp;                                              // gsplat: opacities * compensations
    const float2 ext = gsb_alpha_extent(ca, cb, cc, op);
    GsbRec rec;
    rec.k = make_float4(xy.x, xy.y, ext.x, ext.y);
    // op <= 0 (or NaN): log2 -> -inf / NaN, every comparison in the kernels fails, the record contributes nowhere
    rec.q = make_float4(-0.5f * GSB_LOG2E * ca, -GSB_LOG2E * cb, -0.5f * GSB_LOG2E * cc, log2f(op));
    rec.c = make_float4(r, g, b, op);
    return rec;
}
