/*
 * raster_oracle.c -- CPU restatement of the Gaussian-splat rasterizer the reference calls.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in geosplatting_b200/ may import, link or execute this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * What it restates: gsplat 1.4.0 `rasterization(packed=True, tile_size=16, near_plane=0.01,
 * far_plane=1e10, render_mode='RGB', rasterize_mode in {'classic','antialiased'})`, i.e. exactly the
 * call the reference makes at rfstudio/model/gsplat.py:334-355 (and geosplat.py:276).  gsplat is a
 * third-party dependency pinned `gsplat~=1.4.0` at pyproject.toml:24 and is NOT vendored in
 * /root/reference, nor installable here, and the reference holds no golden vectors for it
 *  => PARITY UNPINNED: this file follows the published algorithm
 *  and is anchored on the reference's call sites only.
 *
 * Arithmetic contract (mirrored by the CUDA path so that tile/bin indices are bit-exact):
 *   - IEEE fp32 everywhere, no FMA contraction (build with -ffp-contract=off), correctly rounded
 *     division and sqrt, operation order exactly as written below;
 *   - the composite uses expf() (the CUDA path uses ex2.approx, tolerance 1e-4 L-inf per pixel).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* Named constants (provenance: gsplat 1.4.0, recalled -- SURVEY.md Appendix C). */
#define ORC_ALPHA_CLAMP 0.999f         /* max alpha                        (C.4) */
#define ORC_ALPHA_MIN (1.0f / 255.0f)  /* skip threshold                   (C.4) */
#define ORC_T_STOP 1e-4f               /* transmittance stop               (C.4) */
#define ORC_RADIUS_DET_FLOOR 0.01f     /* max(0.01, b^2-det) in the radius (C.2) */
#define ORC_FOV_MARGIN 0.3f            /* 0.3*tan_fov clamp margin         (C.2) */
#define ORC_COMP_EPS 1e-6f             /* v_sq = 0.5 v_comp/(comp+1e-6)    (C.6) */

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------- */
/* C.2 projection (one camera).  Outputs are UNPACKED (length N); radii[i]==0 marks a culled     */
/* Gaussian.  The python wrapper does the `packed=True` compaction (ascending Gaussian id).      */
/* ------------------------------------------------------------------------------------------- */

static void quat_to_rot(const float *q, float R[9]) {
    float w = q[0], x = q[1], y = q[2], z = q[3];
    float n2 = ((x * x + y * y) + z * z) + w * w;
    float inv = 1.0f / sqrtf(n2);
    w *= inv; x *= inv; y *= inv; z *= inv;
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float xy = x * y, xz = x * z, yz = y * z;
    float wx = w * x, wy = w * y, wz = w * z;
    R[0] = 1.0f - 2.0f * (y2 + z2); R[1] = 2.0f * (xy - wz);        R[2] = 2.0f * (xz + wy);
    R[3] = 2.0f * (xy + wz);        R[4] = 1.0f - 2.0f * (x2 + z2); R[5] = 2.0f * (yz - wx);
    R[6] = 2.0f * (xz - wy);        R[7] = 2.0f * (yz + wx);        R[8] = 1.0f - 2.0f * (x2 + y2);
}

/* C = A * B, 3x3 row-major, each dot accumulated left to right. */
static void mm3(const float *A, const float *B, float *C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = (A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j]) + A[i * 3 + 2] * B[2 * 3 + j];
}

/* C = A * B^T */
static void mm3_bt(const float *A, const float *B, float *C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = (A[i * 3 + 0] * B[j * 3 + 0] + A[i * 3 + 1] * B[j * 3 + 1]) + A[i * 3 + 2] * B[j * 3 + 2];
}

typedef struct {
    float fx, fy, cx, cy;
    int W, H;
    float near_plane, far_plane, eps2d, radius_clip;
    int antialiased;
} orc_cam_t;

/* Shared forward pieces; returns 0 if culled.  Fills intermediates needed by the backward. */
typedef struct {
    float R[9], M[9], S[9];      /* rotation, R*diag(s), world covariance */
    float pc[3], Sc[9];          /* camera-space mean and covariance */
    float J[6];                  /* 2x3 */
    float tx, ty, rz;
    int x_in, y_in;              /* x/z (y/z) inside the clamp limits */
    float c00, c01, c11;         /* blurred 2D covariance */
    float det0, det;
    float comp;
    float conic[3];
    float mean2d[2];
    int radius;
} orc_proj_t;

static int project_one(const float *mean, const float *quat, const float *scale, const float *vm,
                       const orc_cam_t *cam, orc_proj_t *o) {
    float Rcw[9] = {vm[0], vm[1], vm[2], vm[4], vm[5], vm[6], vm[8], vm[9], vm[10]};
    float t[3] = {vm[3], vm[7], vm[11]};
    for (int i = 0; i < 3; ++i)
        o->pc[i] = ((Rcw[i * 3 + 0] * mean[0] + Rcw[i * 3 + 1] * mean[1]) + Rcw[i * 3 + 2] * mean[2]) + t[i];
    float x = o->pc[0], y = o->pc[1], z = o->pc[2];
    if (z < cam->near_plane || z > cam->far_plane) return 0;

    quat_to_rot(quat, o->R);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) o->M[i * 3 + j] = o->R[i * 3 + j] * scale[j];
    mm3_bt(o->M, o->M, o->S);
    float tmp[9];
    mm3(Rcw, o->S, tmp);
    mm3_bt(tmp, Rcw, o->Sc);

    float fx = cam->fx, fy = cam->fy, cx = cam->cx, cy = cam->cy;
    float Wf = (float)cam->W, Hf = (float)cam->H;
    float tan_fovx = 0.5f * Wf / fx, tan_fovy = 0.5f * Hf / fy;
    float lim_x_pos = (Wf - cx) / fx + ORC_FOV_MARGIN * tan_fovx;
    float lim_x_neg = cx / fx + ORC_FOV_MARGIN * tan_fovx;
    float lim_y_pos = (Hf - cy) / fy + ORC_FOV_MARGIN * tan_fovy;
    float lim_y_neg = cy / fy + ORC_FOV_MARGIN * tan_fovy;
    float rz = 1.0f / z;
    float rz2 = rz * rz;
    float xr = x * rz, yr = y * rz;
    o->x_in = (xr <= lim_x_pos && xr >= -lim_x_neg);
    o->y_in = (yr <= lim_y_pos && yr >= -lim_y_neg);
    float tx = z * fminf(lim_x_pos, fmaxf(-lim_x_neg, xr));
    float ty = z * fminf(lim_y_pos, fmaxf(-lim_y_neg, yr));
    o->tx = tx; o->ty = ty; o->rz = rz;
    float *J = o->J;
    J[0] = fx * rz; J[1] = 0.0f; J[2] = -(fx * tx) * rz2;
    J[3] = 0.0f; J[4] = fy * rz; J[5] = -(fy * ty) * rz2;
    const float *S = o->Sc;
    /* T = J * Sc (2x3), cov2d = T * J^T */
    float T00 = J[0] * S[0] + J[2] * S[6], T01 = J[0] * S[1] + J[2] * S[7], T02 = J[0] * S[2] + J[2] * S[8];
    float T10 = J[4] * S[3] + J[5] * S[6], T11 = J[4] * S[4] + J[5] * S[7], T12 = J[4] * S[5] + J[5] * S[8];
    (void)T10;
    float c00 = T00 * J[0] + T02 * J[2];
    float c01 = T01 * J[4] + T02 * J[5];
    float c11 = T11 * J[4] + T12 * J[5];
    o->mean2d[0] = (fx * x) * rz + cx;
    o->mean2d[1] = (fy * y) * rz + cy;

    float det0 = c00 * c11 - c01 * c01;
    c00 += cam->eps2d;
    c11 += cam->eps2d;
    float det = c00 * c11 - c01 * c01;
    o->det0 = det0; o->det = det;
    o->c00 = c00; o->c01 = c01; o->c11 = c11;
    o->comp = sqrtf(fmaxf(0.0f, det0 / det));
    if (!(det > 0.0f)) return 0;
    o->conic[0] = c11 / det;
    o->conic[1] = -c01 / det;
    o->conic[2] = c00 / det;
    float b = 0.5f * (c00 + c11);
    float v1 = b + sqrtf(fmaxf(ORC_RADIUS_DET_FLOOR, b * b - det));
    float rad = ceilf(3.0f * sqrtf(v1));
    if (rad <= cam->radius_clip) return 0;
    if (o->mean2d[0] + rad <= 0.0f || o->mean2d[0] - rad >= Wf || o->mean2d[1] + rad <= 0.0f ||
        o->mean2d[1] - rad >= Hf)
        return 0;
    o->radius = (int)rad;
    return 1;
}

void orc_project_fwd(int N, const float *means, const float *quats, const float *scales,
                     const float *viewmat, float fx, float fy, float cx, float cy, int W, int H,
                     float near_plane, float far_plane, float eps2d, float radius_clip, int antialiased,
                     int32_t *radii, float *means2d, float *depths, float *conics, float *comps) {
    orc_cam_t cam = {fx, fy, cx, cy, W, H, near_plane, far_plane, eps2d, radius_clip, antialiased};
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        orc_proj_t o;
        int ok = project_one(means + 3 * i, quats + 4 * i, scales + 3 * i, viewmat, &cam, &o);
        if (!ok) {
            radii[i] = 0;
            means2d[2 * i] = means2d[2 * i + 1] = 0.0f;
            depths[i] = 0.0f;
            conics[3 * i] = conics[3 * i + 1] = conics[3 * i + 2] = 0.0f;
            comps[i] = 0.0f;
            continue;
        }
        radii[i] = o.radius;
        means2d[2 * i] = o.mean2d[0];
        means2d[2 * i + 1] = o.mean2d[1];
        depths[i] = o.pc[2];
        conics[3 * i] = o.conic[0];
        conics[3 * i + 1] = o.conic[1];
        conics[3 * i + 2] = o.conic[2];
        comps[i] = antialiased ? o.comp : 1.0f;
    }
}

/* C.6 projection backward.  Inputs are per-Gaussian (unpacked) cotangents; culled rows get zeros. */
void orc_project_bwd(int N, const float *means, const float *quats, const float *scales,
                     const float *viewmat, float fx, float fy, float cx, float cy, int W, int H,
                     float near_plane, float far_plane, float eps2d, float radius_clip, int antialiased,
                     const int32_t *radii, const float *v_means2d, const float *v_depths,
                     const float *v_conics, const float *v_comps, float *v_means, float *v_quats,
                     float *v_scales) {
    orc_cam_t cam = {fx, fy, cx, cy, W, H, near_plane, far_plane, eps2d, radius_clip, antialiased};
    const float *vm = viewmat;
    float Rcw[9] = {vm[0], vm[1], vm[2], vm[4], vm[5], vm[6], vm[8], vm[9], vm[10]};
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        float *vm3 = v_means + 3 * i, *vq = v_quats + 4 * i, *vs = v_scales + 3 * i;
        vm3[0] = vm3[1] = vm3[2] = 0.0f;
        vq[0] = vq[1] = vq[2] = vq[3] = 0.0f;
        vs[0] = vs[1] = vs[2] = 0.0f;
        if (radii[i] <= 0) continue;
        orc_proj_t o;
        if (!project_one(means + 3 * i, quats + 4 * i, scales + 3 * i, viewmat, &cam, &o)) continue;
        float a = o.conic[0], b = o.conic[1], c = o.conic[2];
        float va = v_conics[3 * i], vb = 0.5f * v_conics[3 * i + 1], vc = v_conics[3 * i + 2];
        /* v_cov2d = -Cinv^T * v_Cinv * Cinv^T, Cinv = [[a,b],[b,c]], v_Cinv = [[va,vb],[vb,vc]] */
        float t00 = a * va + b * vb, t01 = a * vb + b * vc;
        float t10 = b * va + c * vb, t11 = b * vb + c * vc;
        float g00 = -(t00 * a + t01 * b), g01 = -(t00 * b + t01 * c);
        float g10 = -(t10 * a + t11 * b), g11 = -(t10 * b + t11 * c);
        if (antialiased) {
            float comp = o.comp;
            float v_comp = v_comps[i];
            float det_conic = a * c - b * b;
            float v_sq = v_comp * 0.5f / (comp + ORC_COMP_EPS);
            float om = 1.0f - comp * comp;
            g00 += v_sq * (om * a - eps2d * det_conic);
            g01 += v_sq * (om * b);
            g10 += v_sq * (om * b);
            g11 += v_sq * (om * c - eps2d * det_conic);
        }
        const float *J = o.J, *S = o.Sc;
        /* v_Sc = J^T G J  (3x3) */
        float GJ[6]; /* 2x3 = G * J */
        for (int j = 0; j < 3; ++j) {
            GJ[j] = g00 * J[j] + g01 * J[3 + j];
            GJ[3 + j] = g10 * J[j] + g11 * J[3 + j];
        }
        float vSc[9];
        for (int r = 0; r < 3; ++r)
            for (int j = 0; j < 3; ++j) vSc[r * 3 + j] = J[r] * GJ[j] + J[3 + r] * GJ[3 + j];
        /* v_J = G J Sc^T + G^T J Sc */
        float GtJ[6];
        for (int j = 0; j < 3; ++j) {
            GtJ[j] = g00 * J[j] + g10 * J[3 + j];
            GtJ[3 + j] = g01 * J[j] + g11 * J[3 + j];
        }
        float vJ[6];
        for (int r = 0; r < 2; ++r)
            for (int j = 0; j < 3; ++j) {
                float s1 = 0.0f, s2 = 0.0f;
                for (int k = 0; k < 3; ++k) {
                    s1 += GJ[r * 3 + k] * S[j * 3 + k];  /* (GJ) Sc^T */
                    s2 += GtJ[r * 3 + k] * S[k * 3 + j]; /* (G^T J) Sc */
                }
                vJ[r * 3 + j] = s1 + s2;
            }
        float x = o.pc[0], y = o.pc[1];
        float rz = o.rz, rz2 = rz * rz, rz3 = rz2 * rz;
        float vmx = v_means2d[2 * i], vmy = v_means2d[2 * i + 1];
        float vpc[3];
        vpc[0] = fx * rz * vmx;
        vpc[1] = fy * rz * vmy;
        vpc[2] = -(fx * x * vmx + fy * y * vmy) * rz2;
        if (o.x_in) vpc[0] += -fx * rz2 * vJ[2];
        else vpc[2] += -fx * rz3 * vJ[2] * o.tx;
        if (o.y_in) vpc[1] += -fy * rz2 * vJ[5];
        else vpc[2] += -fy * rz3 * vJ[5] * o.ty;
        vpc[2] += -fx * rz2 * vJ[0] - fy * rz2 * vJ[4] + 2.0f * fx * o.tx * rz3 * vJ[2] +
                  2.0f * fy * o.ty * rz3 * vJ[5];
        vpc[2] += v_depths ? v_depths[i] : 0.0f;
        /* world <- camera */
        for (int j = 0; j < 3; ++j) vm3[j] = Rcw[0 * 3 + j] * vpc[0] + Rcw[1 * 3 + j] * vpc[1] + Rcw[2 * 3 + j] * vpc[2];
        /* v_S = Rcw^T vSc Rcw */
        float tmp[9], vS[9];
        for (int r = 0; r < 3; ++r)
            for (int j = 0; j < 3; ++j)
                tmp[r * 3 + j] = Rcw[0 * 3 + r] * vSc[0 * 3 + j] + Rcw[1 * 3 + r] * vSc[1 * 3 + j] + Rcw[2 * 3 + r] * vSc[2 * 3 + j];
        mm3(tmp, Rcw, vS);
        /* v_M = (vS + vS^T) M */
        float sym[9], vM[9];
        for (int r = 0; r < 3; ++r)
            for (int j = 0; j < 3; ++j) sym[r * 3 + j] = vS[r * 3 + j] + vS[j * 3 + r];
        mm3(sym, o.M, vM);
        const float *R = o.R;
        const float *s = scales + 3 * i;
        float vR[9];
        for (int j = 0; j < 3; ++j) {
            vs[j] = R[0 * 3 + j] * vM[0 * 3 + j] + R[1 * 3 + j] * vM[1 * 3 + j] + R[2 * 3 + j] * vM[2 * 3 + j];
            for (int r = 0; r < 3; ++r) vR[r * 3 + j] = vM[r * 3 + j] * s[j];
        }
        /* quaternion VJP through the normalisation */
        const float *q = quats + 4 * i;
        float n2 = ((q[1] * q[1] + q[2] * q[2]) + q[3] * q[3]) + q[0] * q[0];
        float inv = 1.0f / sqrtf(n2);
        float w = q[0] * inv, qx = q[1] * inv, qy = q[2] * inv, qz = q[3] * inv;
        /* vR[r][c] is d/dR[r][c] (row-major) */
#define VR(r, c) vR[(r) * 3 + (c)]
        float vqn[4];
        vqn[0] = 2.0f * (qx * (VR(2, 1) - VR(1, 2)) + qy * (VR(0, 2) - VR(2, 0)) + qz * (VR(1, 0) - VR(0, 1)));
        vqn[1] = 2.0f * (-2.0f * qx * (VR(1, 1) + VR(2, 2)) + qy * (VR(1, 0) + VR(0, 1)) + qz * (VR(2, 0) + VR(0, 2)) +
                         w * (VR(2, 1) - VR(1, 2)));
        vqn[2] = 2.0f * (qx * (VR(1, 0) + VR(0, 1)) - 2.0f * qy * (VR(0, 0) + VR(2, 2)) + qz * (VR(2, 1) + VR(1, 2)) +
                         w * (VR(0, 2) - VR(2, 0)));
        vqn[3] = 2.0f * (qx * (VR(2, 0) + VR(0, 2)) + qy * (VR(2, 1) + VR(1, 2)) - 2.0f * qz * (VR(0, 0) + VR(1, 1)) +
                         w * (VR(1, 0) - VR(0, 1)));
#undef VR
        float d = vqn[0] * w + vqn[1] * qx + vqn[2] * qy + vqn[3] * qz;
        vq[0] = (vqn[0] - d * w) * inv;
        vq[1] = (vqn[1] - d * qx) * inv;
        vq[2] = (vqn[2] - d * qy) * inv;
        vq[3] = (vqn[3] - d * qz) * inv;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* C.3 binning.  Arrays here are PACKED (length nnz).                                          */
/* ------------------------------------------------------------------------------------------- */

static void tile_range(const float *m2, int radius, int tile, int tw, int th, int *x0, int *x1, int *y0,
                       int *y1) {
    float ts = (float)tile;
    float tr = (float)radius / ts;
    float txf = m2[0] / ts, tyf = m2[1] / ts;
    float fx0 = floorf(txf - tr), fx1 = ceilf(txf + tr);
    float fy0 = floorf(tyf - tr), fy1 = ceilf(tyf + tr);
    *x0 = (int)fminf(fmaxf(0.0f, fx0), (float)tw);
    *x1 = (int)fminf(fmaxf(0.0f, fx1), (float)tw);
    *y0 = (int)fminf(fmaxf(0.0f, fy0), (float)th);
    *y1 = (int)fminf(fmaxf(0.0f, fy1), (float)th);
}

/* tiles_per_gauss[nnz]; returns total M. */
int64_t orc_isect_count(int nnz, const float *means2d, const int32_t *radii, int tile, int tw, int th,
                        int32_t *tiles_per_gauss) {
    int64_t M = 0;
    for (int i = 0; i < nnz; ++i) {
        int x0, x1, y0, y1;
        tile_range(means2d + 2 * i, radii[i], tile, tw, th, &x0, &x1, &y0, &y1);
        int n = (radii[i] > 0) ? (x1 - x0) * (y1 - y0) : 0;
        tiles_per_gauss[i] = n;
        M += n;
    }
    return M;
}

/* Unsorted (key,val) pairs in generation order: Gaussian-major, tile row-major (y outer). */
void orc_isect_tiles(int nnz, const float *means2d, const int32_t *radii, const float *depths, int tile,
                     int tw, int th, int camera_id, int64_t *keys, int32_t *vals) {
    int n_tiles = tw * th;
    int tile_n_bits = 0;
    while ((1 << tile_n_bits) <= n_tiles) ++tile_n_bits; /* floor(log2(n_tiles)) + 1 */
    int64_t k = 0;
    for (int i = 0; i < nnz; ++i) {
        if (radii[i] <= 0) continue;
        int x0, x1, y0, y1;
        tile_range(means2d + 2 * i, radii[i], tile, tw, th, &x0, &x1, &y0, &y1);
        uint32_t dbits;
        memcpy(&dbits, depths + i, 4);
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) {
                int64_t tid = (int64_t)y * tw + x;
                keys[k] = ((int64_t)camera_id << (32 + tile_n_bits)) | (tid << 32) | (int64_t)dbits;
                vals[k] = i;
                ++k;
            }
    }
}

/* Stable LSD radix sort on the low `bits` bits (8 bits per pass), restating cub::DeviceRadixSort. */
void orc_radix_sort_pairs(int64_t M, int bits, int64_t *keys, int32_t *vals) {
    if (M <= 1) return;
    int64_t *k2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)M);
    int32_t *v2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)M);
    int64_t *ka = keys, *kb = k2;
    int32_t *va = vals, *vb = v2;
    for (int shift = 0; shift < bits; shift += 8) {
        int64_t hist[257];
        memset(hist, 0, sizeof(hist));
        for (int64_t i = 0; i < M; ++i) hist[(((uint64_t)ka[i]) >> shift & 0xFF) + 1]++;
        for (int b = 0; b < 256; ++b) hist[b + 1] += hist[b];
        for (int64_t i = 0; i < M; ++i) {
            int64_t p = hist[((uint64_t)ka[i]) >> shift & 0xFF]++;
            kb[p] = ka[i];
            vb[p] = va[i];
        }
        int64_t *tk = ka; ka = kb; kb = tk;
        int32_t *tv = va; va = vb; vb = tv;
    }
    if (ka != keys) {
        memcpy(keys, ka, sizeof(int64_t) * (size_t)M);
        memcpy(vals, va, sizeof(int32_t) * (size_t)M);
    }
    free(k2);
    free(v2);
}

/* offsets[c*n_tiles + t] = first sorted position whose (camera,tile) id >= (c,t). */
void orc_isect_offsets(int64_t M, const int64_t *sorted_keys, int C, int tw, int th, int32_t *offsets) {
    int n_tiles = tw * th;
    int tile_n_bits = 0;
    while ((1 << tile_n_bits) <= n_tiles) ++tile_n_bits;
    int64_t total = (int64_t)C * n_tiles;
    int64_t pos = 0;
    for (int64_t id = 0; id < total; ++id) {
        int64_t cam = id / n_tiles, t = id % n_tiles;
        int64_t want = (cam << tile_n_bits) | t;
        while (pos < M && (sorted_keys[pos] >> 32) < want) ++pos;
        offsets[id] = (int32_t)pos;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* C.4 / C.5 compositing (one camera, CH colour channels, packed per-Gaussian inputs).          */
/* ------------------------------------------------------------------------------------------- */

void orc_composite_fwd(int W, int H, int tile, int tw, int th, int CH, const float *means2d,
                       const float *conics, const float *colors, const float *opacities,
                       const float *background, const int32_t *offsets, const int32_t *flatten_ids, int64_t M,
                       float *render, float *alphas, int32_t *last_ids) {
    int n_tiles = tw * th;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < n_tiles; ++t) {
        int ty = t / tw, txi = t % tw;
        int start = offsets[t];
        int end = (t == n_tiles - 1) ? (int)M : offsets[t + 1];
        for (int py = ty * tile; py < (ty + 1) * tile && py < H; ++py)
            for (int px = txi * tile; px < (txi + 1) * tile && px < W; ++px) {
                float fxp = (float)px + 0.5f, fyp = (float)py + 0.5f;
                float T = 1.0f;
                float acc[64];
                for (int k = 0; k < CH; ++k) acc[k] = 0.0f;
                int last = 0;
                for (int idx = start; idx < end; ++idx) {
                    int g = flatten_ids[idx];
                    float dx = means2d[2 * g] - fxp, dy = means2d[2 * g + 1] - fyp;
                    float ca = conics[3 * g], cb = conics[3 * g + 1], cc = conics[3 * g + 2];
                    float sigma = 0.5f * (ca * dx * dx + cc * dy * dy) + cb * dx * dy;
                    float alpha = fminf(ORC_ALPHA_CLAMP, opacities[g] * expf(-sigma));
                    if (sigma < 0.0f || alpha < ORC_ALPHA_MIN) continue;
                    float nT = T * (1.0f - alpha);
                    if (nT <= ORC_T_STOP) break;
                    float vis = alpha * T;
                    for (int k = 0; k < CH; ++k) acc[k] += colors[(size_t)g * CH + k] * vis;
                    last = idx;
                    T = nT;
                }
                size_t pix = (size_t)py * W + px;
                for (int k = 0; k < CH; ++k)
                    render[pix * CH + k] = background ? acc[k] + T * background[k] : acc[k];
                alphas[pix] = 1.0f - T;
                last_ids[pix] = last;
            }
    }
}

static inline void atomic_addf(float *p, float v) {
#pragma omp atomic
    *p += v;
}

void orc_composite_bwd(int W, int H, int tile, int tw, int th, int CH, const float *means2d,
                       const float *conics, const float *colors, const float *opacities,
                       const float *background, const int32_t *offsets, const int32_t *flatten_ids, int64_t M,
                       const float *alphas, const int32_t *last_ids, const float *v_render,
                       const float *v_alphas, int nnz, float *v_means2d, float *v_conics, float *v_colors,
                       float *v_opacities) {
    int n_tiles = tw * th;
    memset(v_means2d, 0, sizeof(float) * 2 * (size_t)nnz);
    memset(v_conics, 0, sizeof(float) * 3 * (size_t)nnz);
    memset(v_colors, 0, sizeof(float) * (size_t)CH * (size_t)nnz);
    memset(v_opacities, 0, sizeof(float) * (size_t)nnz);
    (void)M;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < n_tiles; ++t) {
        int ty = t / tw, txi = t % tw;
        int start = offsets[t];
        for (int py = ty * tile; py < (ty + 1) * tile && py < H; ++py)
            for (int px = txi * tile; px < (txi + 1) * tile && px < W; ++px) {
                size_t pix = (size_t)py * W + px;
                float fxp = (float)px + 0.5f, fyp = (float)py + 0.5f;
                float T_final = 1.0f - alphas[pix];
                float T = T_final;
                float buffer[64];
                for (int k = 0; k < CH; ++k) buffer[k] = 0.0f;
                const float *vr = v_render + pix * CH;
                float va_out = v_alphas[pix];
                int bin_final = last_ids[pix];
                for (int idx = bin_final; idx >= start; --idx) {
                    int g = flatten_ids[idx];
                    float dx = means2d[2 * g] - fxp, dy = means2d[2 * g + 1] - fyp;
                    float ca = conics[3 * g], cb = conics[3 * g + 1], cc = conics[3 * g + 2];
                    float sigma = 0.5f * (ca * dx * dx + cc * dy * dy) + cb * dx * dy;
                    float opac = opacities[g];
                    float vis = expf(-sigma);
                    float alpha = fminf(ORC_ALPHA_CLAMP, opac * vis);
                    if (sigma < 0.0f || alpha < ORC_ALPHA_MIN) continue;
                    float ra = 1.0f / (1.0f - alpha);
                    T *= ra;
                    float fac = alpha * T;
                    float v_alpha = 0.0f;
                    for (int k = 0; k < CH; ++k) {
                        float c = colors[(size_t)g * CH + k];
                        atomic_addf(&v_colors[(size_t)g * CH + k], fac * vr[k]);
                        v_alpha += (c * T - buffer[k] * ra) * vr[k];
                    }
                    v_alpha += T_final * ra * va_out;
                    if (background) {
                        float accum = 0.0f;
                        for (int k = 0; k < CH; ++k) accum += background[k] * vr[k];
                        v_alpha += -T_final * ra * accum;
                    }
                    if (opac * vis <= ORC_ALPHA_CLAMP) {
                        float v_sigma = -opac * vis * v_alpha;
                        atomic_addf(&v_conics[3 * g + 0], 0.5f * v_sigma * dx * dx);
                        atomic_addf(&v_conics[3 * g + 1], v_sigma * dx * dy);
                        atomic_addf(&v_conics[3 * g + 2], 0.5f * v_sigma * dy * dy);
                        atomic_addf(&v_means2d[2 * g + 0], v_sigma * (ca * dx + cb * dy));
                        atomic_addf(&v_means2d[2 * g + 1], v_sigma * (cb * dx + cc * dy));
                        atomic_addf(&v_opacities[g], vis * v_alpha);
                    }
                    for (int k = 0; k < CH; ++k) buffer[k] += colors[(size_t)g * CH + k] * fac;
                }
            }
    }
}

/* Marks pixels whose result hinges on a discrete decision that is within `eps` (relative) of flipping:
 * alpha vs 1/255, next-T vs 1e-4, sigma vs 0.  Two correct fp32 implementations that round exp()
 * differently may legitimately disagree on exactly these pixels; parity tests compare the rest at
 * 1e-4 and bound the number of fragile pixels. */
void orc_composite_fragile(int W, int H, int tile, int tw, int th, const float *means2d, const float *conics,
                           const float *opacities, const int32_t *offsets, const int32_t *flatten_ids,
                           int64_t M, float eps, uint8_t *fragile) {
    int n_tiles = tw * th;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < n_tiles; ++t) {
        int ty = t / tw, txi = t % tw;
        int start = offsets[t];
        int end = (t == n_tiles - 1) ? (int)M : offsets[t + 1];
        for (int py = ty * tile; py < (ty + 1) * tile && py < H; ++py)
            for (int px = txi * tile; px < (txi + 1) * tile && px < W; ++px) {
                float fxp = (float)px + 0.5f, fyp = (float)py + 0.5f;
                float T = 1.0f;
                uint8_t flag = 0;
                for (int idx = start; idx < end; ++idx) {
                    int g = flatten_ids[idx];
                    float dx = means2d[2 * g] - fxp, dy = means2d[2 * g + 1] - fyp;
                    float ca = conics[3 * g], cb = conics[3 * g + 1], cc = conics[3 * g + 2];
                    float sigma = 0.5f * (ca * dx * dx + cc * dy * dy) + cb * dx * dy;
                    float alpha = fminf(ORC_ALPHA_CLAMP, opacities[g] * expf(-sigma));
                    if (fabsf(sigma) < 1e-6f) flag = 1;
                    if (fabsf(alpha - ORC_ALPHA_MIN) <= eps * ORC_ALPHA_MIN) flag = 1;
                    if (sigma < 0.0f || alpha < ORC_ALPHA_MIN) continue;
                    float nT = T * (1.0f - alpha);
                    if (fabsf(nT - ORC_T_STOP) <= eps * ORC_T_STOP) flag = 1;
                    if (nT <= ORC_T_STOP) break;
                    T = nT;
                }
                fragile[(size_t)py * W + px] = flag;
            }
    }
}
