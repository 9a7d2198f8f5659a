/*
 * prefilter_oracle.c -- CPU restatement of the reference's split-sum cube-map prefilter plugin
 * `rfstudio_render_utils` (the only hot-path native code the reference owns).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Follows rfstudio/graphics/_mesh/_splitsum/c_src/cubemap.cu:
 *     pixel_area            :17-30     cube_to_dir        :32-46      dir_to_side (unused by the kernels)
 *     DiffuseCubemapFwd/Bwd :110-169   SpecularBounds     :181-244
 *     SpecularCubemapFwd/Bwd:246-350   ndfGGX             :174-179
 * and the host glue rfstudio/graphics/_mesh/_splitsum/_wrap.py:120-157 (cut-off angle from a 1e6-sample
 * cumulative GGX, output = rgb / wsum).  Literal constants are kept: pi = 3.141592f in the diffuse weight,
 * M_PI (double) inside ndfGGX, clamp 0.999f, tile size 16 in the bounds search, and pixel_area()'s
 * |x - N/2| indexing exactly as written (it is asymmetric; do not "fix" it).
 * Pinned against the real plugin (oracle/_ref, built from the reference sources) on the GPU box:
 * tests/test_prefilter_gpu.py::test_oracle_matches_reference_plugin.
 *
 * Layouts: cubemap [6,R,R,3], bounds [6,R,R,24] (float-encoded ints, as the plugin stores them),
 * specular output [6,R,R,4] = (sum w*rgb, sum w).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static float pixel_area(int x, int y, int N) {
    if (N > 1) {
        int H = N / 2;
        x = x - H; if (x < 0) x = -x;
        y = y - H; if (y < 0) y = -y;
        float dx = atanf((float)(x + 1) / (float)H) - atanf((float)x / (float)H);
        float dy = atanf((float)(y + 1) / (float)H) - atanf((float)y / (float)H);
        return dx * dy;
    }
    return 1.0f;
}

static void safe_normalize(float v[3]) {
    float l = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (l > 0.0f) { v[0] /= l; v[1] /= l; v[2] /= l; }
    else { v[0] = v[1] = v[2] = 0.0f; }
}

static void cube_to_dir(int x, int y, int side, int N, float out[3]) {
    float fx = 2.0f * (((float)x + 0.5f) / (float)N) - 1.0f;
    float fy = 2.0f * (((float)y + 0.5f) / (float)N) - 1.0f;
    switch (side) {
        case 0: out[0] = 1; out[1] = -fy; out[2] = -fx; break;
        case 1: out[0] = -1; out[1] = -fy; out[2] = fx; break;
        case 2: out[0] = fx; out[1] = 1; out[2] = fy; break;
        case 3: out[0] = fx; out[1] = -1; out[2] = -fy; break;
        case 4: out[0] = fx; out[1] = -fy; out[2] = 1; break;
        default: out[0] = -fx; out[1] = -fy; out[2] = -1; break;
    }
    safe_normalize(out);
}

static float dot3(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static float diffuse_weight(const float Nrm[3], int x, int y, int s, int R) {
    float L[3];
    cube_to_dir(x, y, s, R, L);
    float costheta = fminf(fmaxf(dot3(Nrm, L), 0.0f), 0.999f);
    return costheta * pixel_area(x, y, R) / 3.141592f;
}

void orc_diffuse_cubemap_fwd(int R, const float *cubemap, float *out) {
#pragma omp parallel for schedule(static)
    for (int o = 0; o < 6 * R * R; ++o) {
        int pz = o / (R * R), py = (o / R) % R, px = o % R;
        float Nrm[3];
        cube_to_dir(px, py, pz, R, Nrm);
        float col[3] = {0, 0, 0};
        for (int s = 0; s < 6; ++s)
            for (int y = 0; y < R; ++y)
                for (int x = 0; x < R; ++x) {
                    float w = diffuse_weight(Nrm, x, y, s, R);
                    const float *c = cubemap + (((size_t)s * R + y) * R + x) * 3;
                    col[0] += c[0] * w; col[1] += c[1] * w; col[2] += c[2] * w;
                }
        out[3 * o] = col[0]; out[3 * o + 1] = col[1]; out[3 * o + 2] = col[2];
    }
}

void orc_diffuse_cubemap_bwd(int R, const float *grad_out, float *grad_in) {
    memset(grad_in, 0, sizeof(float) * 18 * (size_t)R * R);
    /* gather form of the reference's atomic scatter: deterministic, same sums */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < 6 * R * R; ++i) {
        int s = i / (R * R), y = (i / R) % R, x = i % R;
        float acc[3] = {0, 0, 0};
        for (int o = 0; o < 6 * R * R; ++o) {
            int pz = o / (R * R), py = (o / R) % R, px = o % R;
            float Nrm[3];
            cube_to_dir(px, py, pz, R, Nrm);
            float w = diffuse_weight(Nrm, x, y, s, R);
            acc[0] += grad_out[3 * o] * w; acc[1] += grad_out[3 * o + 1] * w; acc[2] += grad_out[3 * o + 2] * w;
        }
        grad_in[3 * i] = acc[0]; grad_in[3 * i + 1] = acc[1]; grad_in[3 * i + 2] = acc[2];
    }
}

void orc_specular_bounds(int R, float costheta_cutoff, float *bounds) {
    const int TILE_SIZE = 16;
#pragma omp parallel for schedule(dynamic, 64)
    for (int o = 0; o < 6 * R * R; ++o) {
        int pz = o / (R * R), py = (o / R) % R, px = o % R;
        float V[3];
        cube_to_dir(px, py, pz, R, V);
        for (int s = 0; s < 6; ++s) {
            int min_x = R - 1, max_x = 0, min_y = R - 1, max_y = 0;
            for (int tx = 0; tx < (R + TILE_SIZE - 1) / TILE_SIZE; tx++)
                for (int ty = 0; ty < (R + TILE_SIZE - 1) / TILE_SIZE; ty++) {
                    int tsx = tx * TILE_SIZE, tsy = ty * TILE_SIZE;
                    int tex = (tx + 1) * TILE_SIZE < R ? (tx + 1) * TILE_SIZE : R;
                    int tey = (ty + 1) * TILE_SIZE < R ? (ty + 1) * TILE_SIZE : R;
                    float L0[3], L1[3], L2[3], L3[3];
                    cube_to_dir(tsx, tsy, s, R, L0); cube_to_dir(tex, tsy, s, R, L1);
                    cube_to_dir(tsx, tey, s, R, L2); cube_to_dir(tex, tey, s, R, L3);
                    float mn[3], mx[3];
                    for (int k = 0; k < 3; ++k) {
                        mn[k] = fminf(fminf(L0[k], L1[k]), fminf(L2[k], L3[k]));
                        mx[k] = fmaxf(fmaxf(L0[k], L1[k]), fmaxf(L2[k], L3[k]));
                    }
                    float maxdp = fmaxf(mn[0] * V[0], mx[0] * V[0]) + fmaxf(mn[1] * V[1], mx[1] * V[1]) +
                                  fmaxf(mn[2] * V[2], mx[2] * V[2]);
                    if (maxdp >= costheta_cutoff) {
                        for (int y = tsy; y < tey; ++y)
                            for (int x = tsx; x < tex; ++x) {
                                float L[3];
                                cube_to_dir(x, y, s, R, L);
                                if (dot3(L, V) >= costheta_cutoff) {
                                    if (x < min_x) min_x = x;
                                    if (x > max_x) max_x = x;
                                    if (y < min_y) min_y = y;
                                    if (y > max_y) max_y = y;
                                }
                            }
                    }
                }
            float *b = bounds + (size_t)o * 24 + s * 4;
            b[0] = (float)min_x; b[1] = (float)max_x; b[2] = (float)min_y; b[3] = (float)max_y;
        }
    }
}

static float ndf_ggx(float alphaSqr, float cosTheta) {
    float c = fminf(fmaxf(cosTheta, 0.0f), 1.0f);
    float d = (c * alphaSqr - c) * c + 1.0f;
    return (float)((double)alphaSqr / ((double)(d * d) * M_PI));
}

static float specular_weight(const float V[3], int x, int y, int s, int R, float alphaSqr, float cutoff, int *hit) {
    float L[3];
    cube_to_dir(x, y, s, R, L);
    float d = dot3(L, V);
    if (!(d >= cutoff)) { *hit = 0; return 0.0f; }
    *hit = 1;
    float Hv[3] = {L[0] + V[0], L[1] + V[1], L[2] + V[2]};
    safe_normalize(Hv);
    float wiDotN = fmaxf(d, 0.0f);
    float VdotH = fmaxf(dot3(V, Hv), 0.0f);
    return wiDotN * ndf_ggx(alphaSqr, VdotH) * pixel_area(x, y, R) / 4.0f;
}

void orc_specular_cubemap_fwd(int R, const float *cubemap, const float *bounds, float roughness,
                              float costheta_cutoff, float *out) {
    float alpha = roughness * roughness;
    float alphaSqr = alpha * alpha;
#pragma omp parallel for schedule(dynamic, 64)
    for (int o = 0; o < 6 * R * R; ++o) {
        int pz = o / (R * R), py = (o / R) % R, px = o % R;
        float V[3];
        cube_to_dir(px, py, pz, R, V);
        float wsum = 0.0f, col[3] = {0, 0, 0};
        for (int s = 0; s < 6; ++s) {
            const float *b = bounds + (size_t)o * 24 + s * 4;
            int xmin = (int)b[0], xmax = (int)b[1], ymin = (int)b[2], ymax = (int)b[3];
            if (xmin <= xmax)
                for (int y = ymin; y <= ymax; ++y)
                    for (int x = xmin; x <= xmax; ++x) {
                        int hit;
                        float w = specular_weight(V, x, y, s, R, alphaSqr, costheta_cutoff, &hit);
                        if (!hit) continue;
                        const float *c = cubemap + (((size_t)s * R + y) * R + x) * 3;
                        col[0] += c[0] * w; col[1] += c[1] * w; col[2] += c[2] * w;
                        wsum += w;
                    }
        }
        out[4 * o] = col[0]; out[4 * o + 1] = col[1]; out[4 * o + 2] = col[2]; out[4 * o + 3] = wsum;
    }
}

/* grad_out is [6,R,R,4]; like the plugin, only channels 0..2 are read (cubemap.cu:311). */
void orc_specular_cubemap_bwd(int R, const float *bounds, const float *grad_out, float roughness,
                              float costheta_cutoff, float *grad_in) {
    float alpha = roughness * roughness;
    float alphaSqr = alpha * alpha;
    memset(grad_in, 0, sizeof(float) * 18 * (size_t)R * R);
    for (int o = 0; o < 6 * R * R; ++o) {  /* serial scatter: deterministic reference order */
        int pz = o / (R * R), py = (o / R) % R, px = o % R;
        float V[3];
        cube_to_dir(px, py, pz, R, V);
        const float *g = grad_out + 4 * (size_t)o;
        for (int s = 0; s < 6; ++s) {
            const float *b = bounds + (size_t)o * 24 + s * 4;
            int xmin = (int)b[0], xmax = (int)b[1], ymin = (int)b[2], ymax = (int)b[3];
            if (xmin <= xmax)
                for (int y = ymin; y <= ymax; ++y)
                    for (int x = xmin; x <= xmax; ++x) {
                        int hit;
                        float w = specular_weight(V, x, y, s, R, alphaSqr, costheta_cutoff, &hit);
                        if (!hit) continue;
                        float *gi = grad_in + (((size_t)s * R + y) * R + x) * 3;
                        gi[0] += g[0] * w; gi[1] += g[1] * w; gi[2] += g[2] * w;
                    }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Double-precision evaluation of the SAME specular forward (same texel set: the cone test is taken
 * from the fp32 path so that membership is identical), for a strided subset of output texels.
 * The reference's GGX weight alpha^2/(pi*(1-c^2(1-alpha^2))^2) cancels catastrophically in fp32 near
 * c = 1 (relative error of one weight ~ 2*6e-8/alpha^2: 3e-3 at roughness 0.08), so two correct fp32
 * builds legitimately differ at that level; this function is the yardstick both are measured against.
 * out[(o/stride)*5 ..] = (sum w*rgb, sum w, fragile) for o = 0, stride, 2*stride, ...; fragile = 1 when a tap
 * sits within 2e-6 of the cone cut-off (its membership may differ between two fp32 builds).
 * ------------------------------------------------------------------------------------------------ */
static void cube_to_dir_d(int x, int y, int side, int N, double out[3]) {
    double fx = 2.0 * (((double)x + 0.5) / (double)N) - 1.0;
    double fy = 2.0 * (((double)y + 0.5) / (double)N) - 1.0;
    switch (side) {
        case 0: out[0] = 1; out[1] = -fy; out[2] = -fx; break;
        case 1: out[0] = -1; out[1] = -fy; out[2] = fx; break;
        case 2: out[0] = fx; out[1] = 1; out[2] = fy; break;
        case 3: out[0] = fx; out[1] = -1; out[2] = -fy; break;
        case 4: out[0] = fx; out[1] = -fy; out[2] = 1; break;
        default: out[0] = -fx; out[1] = -fy; out[2] = -1; break;
    }
    double l = sqrt(out[0] * out[0] + out[1] * out[1] + out[2] * out[2]);
    out[0] /= l; out[1] /= l; out[2] /= l;
}

static double pixel_area_d(int x, int y, int N) {
    if (N <= 1) return 1.0;
    int H = N / 2;
    x = x - H; if (x < 0) x = -x;
    y = y - H; if (y < 0) y = -y;
    return (atan((double)(x + 1) / H) - atan((double)x / H)) * (atan((double)(y + 1) / H) - atan((double)y / H));
}

void orc_specular_cubemap_fwd_f64(int R, const float *cubemap, const float *bounds, float roughness,
                                  float costheta_cutoff, int stride, double *out) {
    double alpha = (double)roughness * (double)roughness;
    double alphaSqr = alpha * alpha;
    int n_out = (6 * R * R + stride - 1) / stride;
#pragma omp parallel for schedule(dynamic, 16)
    for (int j = 0; j < n_out; ++j) {
        int o = j * stride;
        int pz = o / (R * R), py = (o / R) % R, px = o % R;
        float Vf[3];
        double V[3];
        cube_to_dir(px, py, pz, R, Vf);
        cube_to_dir_d(px, py, pz, R, V);
        double wsum = 0.0, col[3] = {0, 0, 0}, fragile = 0.0;
        for (int s = 0; s < 6; ++s) {
            const float *b = bounds + (size_t)o * 24 + s * 4;
            int xmin = (int)b[0], xmax = (int)b[1], ymin = (int)b[2], ymax = (int)b[3];
            if (xmin <= xmax)
                for (int y = ymin; y <= ymax; ++y)
                    for (int x = xmin; x <= xmax; ++x) {
                        float Lf[3];
                        cube_to_dir(x, y, s, R, Lf);
                        float dm = dot3(Lf, Vf);
                        if (fabsf(dm - costheta_cutoff) < 2e-6f) fragile = 1.0; /* membership may flip */
                        if (!(dm >= costheta_cutoff)) continue; /* fp32 membership */
                        double L[3];
                        cube_to_dir_d(x, y, s, R, L);
                        double d = L[0] * V[0] + L[1] * V[1] + L[2] * V[2];
                        double Hh[3] = {L[0] + V[0], L[1] + V[1], L[2] + V[2]};
                        double hl = sqrt(Hh[0] * Hh[0] + Hh[1] * Hh[1] + Hh[2] * Hh[2]);
                        double c = (V[0] * Hh[0] + V[1] * Hh[1] + V[2] * Hh[2]) / hl;
                        c = c < 0 ? 0 : (c > 1 ? 1 : c);
                        double dd = (c * alphaSqr - c) * c + 1.0;
                        double w = (d > 0 ? d : 0) * (alphaSqr / (dd * dd * M_PI)) * pixel_area_d(x, y, R) / 4.0;
                        const float *cc = cubemap + (((size_t)s * R + y) * R + x) * 3;
                        col[0] += cc[0] * w; col[1] += cc[1] * w; col[2] += cc[2] * w;
                        wsum += w;
                    }
        }
        out[5 * j] = col[0]; out[5 * j + 1] = col[1]; out[5 * j + 2] = col[2]; out[5 * j + 3] = wsum;
        out[5 * j + 4] = fragile;
    }
}
