// Projection backward: per-Gaussian VJP of the EWA projection (streaming, HBM-bound; recomputes the
// forward intermediates instead of saving them: 40 B of inputs vs ~200 B of saved state per Gaussian).
//
// Replaces gsplat 1.4.0 fully_fused_projection_packed_bwd (third-party), reached
// through autograd from rfstudio/model/gsplat.py:334-355.
#include "project_math.cuh"

#ifndef GSB_PROJECT_BWD_MINB
#define GSB_PROJECT_BWD_MINB 5   // 48 registers: the kernel waits for memory (long scoreboard), 0.056 -> 0.046 ms with 40 resident warps
#endif
__global__ void __launch_bounds__(256, GSB_PROJECT_BWD_MINB) project_bwd_kernel(
    int N, const float *__restrict__ means, const float *__restrict__ quats, const float *__restrict__ scales,
    CamK cam, const int32_t *__restrict__ radii, const float2 *__restrict__ v_means2d,
    const float *__restrict__ v_depths, const float *__restrict__ v_conics, const float *__restrict__ v_comps,
    float *__restrict__ v_means, float *__restrict__ v_quats, float *__restrict__ v_scales,
    const float *__restrict__ opacity_logits, const float *__restrict__ v_opacities_eff,
    float *__restrict__ v_opacity_logits, int accumulate) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float vm3[3] = {0.f, 0.f, 0.f}, vq[4] = {0.f, 0.f, 0.f, 0.f}, vs[3] = {0.f, 0.f, 0.f};
    float v_logit = 0.f;
    ProjOut o;
    float m[3] = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
    float4 q4 = reinterpret_cast<const float4 *>(quats)[i];
    float q[4] = {q4.x, q4.y, q4.z, q4.w};
    float s[3] = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
    {
        // The VJP is linear in the cotangents.  A Gaussian that was culled, or that no pixel's gradient reached (behind
        // an opaque surface, off screen: about half of a closed mesh), has nothing to add: leave before the arithmetic --
        // in a batch (accumulate) also before its outputs are read and rewritten.  (Its parameters were requested above
        // together with the cotangents: one round trip to memory, not two.)  (x != 0 is true for NaN: those still propagate.)
        bool touched = false;
        if (radii[i] > 0) {
            const float2 m2 = v_means2d[i];
            touched = m2.x != 0.f || m2.y != 0.f || v_conics[3 * i] != 0.f || v_conics[3 * i + 1] != 0.f ||
                      v_conics[3 * i + 2] != 0.f || (v_depths && v_depths[i] != 0.f) || (v_comps && v_comps[i] != 0.f) ||
                      (v_opacities_eff && v_opacities_eff[i] != 0.f);
        }
        if (!touched) {
            if (!accumulate) {
                v_means[3 * i] = 0.f; v_means[3 * i + 1] = 0.f; v_means[3 * i + 2] = 0.f;
                reinterpret_cast<float4 *>(v_quats)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                v_scales[3 * i] = 0.f; v_scales[3 * i + 1] = 0.f; v_scales[3 * i + 2] = 0.f;
                if (v_opacity_logits) v_opacity_logits[i] = 0.f;
            }
            return;
        }
    }
    if (gsb_project_one<false>(m, q, s, cam, o)) {   // radii > 0 (checked above): the forward's cull verdict is final
        const float fx = cam.fx, fy = cam.fy;
        float a = o.conic[0], b = o.conic[1], c = o.conic[2];
        float va = v_conics[3 * i], vb = 0.5f * v_conics[3 * i + 1], vc = v_conics[3 * i + 2];
        // G = -Cinv^T v_Cinv Cinv^T
        float t00 = a * va + b * vb, t01 = a * vb + b * vc;
        float t10 = b * va + c * vb, t11 = b * vb + c * vc;
        float g00 = -(t00 * a + t01 * b), g01 = -(t00 * b + t01 * c);
        float g10 = -(t10 * a + t11 * b), g11 = -(t10 * b + t11 * c);
        // fused opacity activation (opacity_eff = sigmoid(logit) * comp): chain rule back to the logit and to comp
        float v_comp_opac = 0.f;
        if (opacity_logits) {
            const float sig = 1.0f / (1.0f + expf(-opacity_logits[i]));
            const float ve = v_opacities_eff[i];
            v_comp_opac = ve * sig;
            v_logit = ve * (cam.antialiased ? o.comp : 1.0f) * sig * (1.0f - sig);
        }
        if (cam.antialiased) {
            float comp = o.comp;
            float det_conic = a * c - b * b;
            float v_sq = ((v_comps ? v_comps[i] : 0.f) + v_comp_opac) * 0.5f / (comp + GSB_COMP_EPS);
            float om = 1.0f - comp * comp;
            g00 += v_sq * (om * a - cam.eps2d * det_conic);
            g01 += v_sq * (om * b);
            g10 += v_sq * (om * b);
            g11 += v_sq * (om * c - cam.eps2d * det_conic);
        }
        const float *J = o.J, *S = o.Sc;
        float GJ[6], GtJ[6];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            GJ[j] = g00 * J[j] + g01 * J[3 + j];
            GJ[3 + j] = g10 * J[j] + g11 * J[3 + j];
            GtJ[j] = g00 * J[j] + g10 * J[3 + j];
            GtJ[3 + j] = g01 * J[j] + g11 * J[3 + j];
        }
        float vSc[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int j = 0; j < 3; ++j) vSc[r * 3 + j] = J[r] * GJ[j] + J[3 + r] * GJ[3 + j];
        // only the third column of v_J and its two focal entries are consumed
        float vJ00 = 0.f, vJ02 = 0.f, vJ11 = 0.f, vJ12 = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            vJ00 += GJ[k] * S[0 * 3 + k] + GtJ[k] * S[k * 3 + 0];
            vJ02 += GJ[k] * S[2 * 3 + k] + GtJ[k] * S[k * 3 + 2];
            vJ11 += GJ[3 + k] * S[1 * 3 + k] + GtJ[3 + k] * S[k * 3 + 1];
            vJ12 += GJ[3 + k] * S[2 * 3 + k] + GtJ[3 + k] * S[k * 3 + 2];
        }
        float x = o.pc[0], y = o.pc[1];
        float rz = o.rz, rz2 = rz * rz, rz3 = rz2 * rz;
        float2 vm2 = v_means2d[i];
        float vpc[3];
        vpc[0] = fx * rz * vm2.x;
        vpc[1] = fy * rz * vm2.y;
        vpc[2] = -(fx * x * vm2.x + fy * y * vm2.y) * rz2;
        if (o.x_in) vpc[0] += -fx * rz2 * vJ02;
        else vpc[2] += -fx * rz3 * vJ02 * o.tx;
        if (o.y_in) vpc[1] += -fy * rz2 * vJ12;
        else vpc[2] += -fy * rz3 * vJ12 * o.ty;
        vpc[2] += -fx * rz2 * vJ00 - fy * rz2 * vJ11 + 2.0f * fx * o.tx * rz3 * vJ02 + 2.0f * fy * o.ty * rz3 * vJ12;
        if (v_depths) vpc[2] += v_depths[i];
        const float *Rcw = cam.r;
#pragma unroll
        for (int j = 0; j < 3; ++j) vm3[j] = Rcw[0 * 3 + j] * vpc[0] + Rcw[1 * 3 + j] * vpc[1] + Rcw[2 * 3 + j] * vpc[2];
        // v_S = Rcw^T vSc Rcw ; v_M = (v_S + v_S^T) M
        float tmp[9], vS[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                tmp[r * 3 + j] = Rcw[0 * 3 + r] * vSc[0 * 3 + j] + Rcw[1 * 3 + r] * vSc[1 * 3 + j] + Rcw[2 * 3 + r] * vSc[2 * 3 + j];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                vS[r * 3 + j] = tmp[r * 3 + 0] * Rcw[0 * 3 + j] + tmp[r * 3 + 1] * Rcw[1 * 3 + j] + tmp[r * 3 + 2] * Rcw[2 * 3 + j];
        float vM[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) acc += (vS[r * 3 + k] + vS[k * 3 + r]) * o.M[k * 3 + j];
                vM[r * 3 + j] = acc;
            }
        const float *R = o.R;
        float vR[9];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            vs[j] = R[0 * 3 + j] * vM[0 * 3 + j] + R[1 * 3 + j] * vM[1 * 3 + j] + R[2 * 3 + j] * vM[2 * 3 + j];
#pragma unroll
            for (int r = 0; r < 3; ++r) vR[r * 3 + j] = vM[r * 3 + j] * s[j];
        }
        float n2 = ((q[1] * q[1] + q[2] * q[2]) + q[3] * q[3]) + q[0] * q[0];
        float inv = 1.0f / sqrtf(n2);
        float w = q[0] * inv, qx = q[1] * inv, qy = q[2] * inv, qz = q[3] * inv;
#define VR(r, c) vR[(r) * 3 + (c)]
        float vqn0 = 2.0f * (qx * (VR(2, 1) - VR(1, 2)) + qy * (VR(0, 2) - VR(2, 0)) + qz * (VR(1, 0) - VR(0, 1)));
        float vqn1 = 2.0f * (-2.0f * qx * (VR(1, 1) + VR(2, 2)) + qy * (VR(1, 0) + VR(0, 1)) + qz * (VR(2, 0) + VR(0, 2)) + w * (VR(2, 1) - VR(1, 2)));
        float vqn2 = 2.0f * (qx * (VR(1, 0) + VR(0, 1)) - 2.0f * qy * (VR(0, 0) + VR(2, 2)) + qz * (VR(2, 1) + VR(1, 2)) + w * (VR(0, 2) - VR(2, 0)));
        float vqn3 = 2.0f * (qx * (VR(2, 0) + VR(0, 2)) + qy * (VR(2, 1) + VR(1, 2)) - 2.0f * qz * (VR(0, 0) + VR(1, 1)) + w * (VR(1, 0) - VR(0, 1)));
#undef VR
        float d = vqn0 * w + vqn1 * qx + vqn2 * qy + vqn3 * qz;
        vq[0] = (vqn0 - d * w) * inv;
        vq[1] = (vqn1 - d * qx) * inv;
        vq[2] = (vqn2 - d * qy) * inv;
        vq[3] = (vqn3 - d * qz) * inv;
    }
    if (accumulate) {   // several views of a batch collect in one buffer (same stream: plain read-modify-write)
#pragma unroll
        for (int k = 0; k < 3; ++k) { vm3[k] += v_means[3 * i + k]; vs[k] += v_scales[3 * i + k]; }
        float4 q0 = reinterpret_cast<const float4 *>(v_quats)[i];
        vq[0] += q0.x; vq[1] += q0.y; vq[2] += q0.z; vq[3] += q0.w;
        if (v_opacity_logits) v_logit += v_opacity_logits[i];
    }
    v_means[3 * i] = vm3[0]; v_means[3 * i + 1] = vm3[1]; v_means[3 * i + 2] = vm3[2];
    reinterpret_cast<float4 *>(v_quats)[i] = make_float4(vq[0], vq[1], vq[2], vq[3]);
    v_scales[3 * i] = vs[0]; v_scales[3 * i + 1] = vs[1]; v_scales[3 * i + 2] = vs[2];
    if (v_opacity_logits) v_opacity_logits[i] = v_logit;
}

extern "C" __attribute__((visibility("default"))) int gsb_project_bwd(int32_t N, const float *means, const float *quats, const float *scales,
                               const gsb_camera *cam, const int32_t *radii, const float *v_means2d,
                               const float *v_depths, const float *v_conics, const float *v_comps,
                               float *v_means, float *v_quats, float *v_scales, const float *opacity_logits,
                               const float *v_opacities_eff, float *v_opacity_logits, int32_t accumulate,
                               void *stream) {
    GSB_CHECK_ARG(N >= 0 && cam != nullptr);
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(means && quats && scales && radii && v_means2d && v_conics && v_means && v_quats && v_scales);
    GSB_CHECK_ARG((opacity_logits == nullptr) == (v_opacities_eff == nullptr) &&
                  (opacity_logits == nullptr) == (v_opacity_logits == nullptr));
    GSB_CHECK_ARG(!cam->antialiased || v_comps != nullptr || opacity_logits != nullptr);
    CamK k = gsb_make_cam(cam);
    project_bwd_kernel<<<gsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
        N, means, quats, scales, k, radii, reinterpret_cast<const float2 *>(v_means2d), v_depths, v_conics,
        v_comps, v_means, v_quats, v_scales, opacity_logits, v_opacities_eff, v_opacity_logits, accumulate);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}
