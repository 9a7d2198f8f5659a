// Split-sum cube-map prefilter: diffuse irradiance, GGX specular prefilter inside a cached cone AABB,
// the AABB search itself, and the 2x2 mip chain with the reference's (non-transpose) backward.
//
// Replaces the reference's own plugin `rfstudio_render_utils`
// (rfstudio/graphics/_mesh/_splitsum/c_src/cubemap.cu:110-350, torch_bindings.cpp:112-271) and
// `_CubeMapMip` (rfstudio/graphics/_mesh/_texture.py:199-226).
//
// Differences in HOW (results agree to fp32 summation order):
//  * both backward passes are written in GATHER form -- the weight k(V,L) = max(L.V,0) * D_ggx(V.H) / 4 and the
//    cone test L.V >= cos(theta_c) are symmetric in (V,L), so the cone AABB table of texel L bounds exactly the
//    output texels V that L contributes to; no atomics (the reference issues 3 atomicAdd per tap);
//  * the texel solid angle is separable (pixel_area(x,y) = a(x) * a(y)); a 1-D table lives in shared memory instead
//    of four atanf per tap;
//  * the specular forward can write rgb/wsum straight into this library's env stack (float4 texels, .w = wsum).
#include "texture_math.cuh"

namespace {

// Unit direction of texel (x, y) on `side`.  The cone membership test L.V >= cos(theta_c) sits on a fp32
// knife edge at high resolution (neighbouring boundary texels differ by ~1e-6 in L.V), so the arithmetic here
// keeps the reference's expression forms (cubemap.cu:32-46: divide by N, sqrtf, three IEEE divisions) to make
// the same texels pass the test.
__device__ __forceinline__ float3 texel_dir(int x, int y, int side, float N) {
    float fx = 2.0f * (((float)x + 0.5f) / N) - 1.0f;
    float fy = 2.0f * (((float)y + 0.5f) / N) - 1.0f;
    float px, py, pz;
    gsb_face_point(side, fx, fy, px, py, pz);
    float l = sqrtf(px * px + py * py + pz * pz);
    return make_float3(px / l, py / l, pz / l);
}

// cubemap.cu:17-30 verbatim semantics (including the |x - N/2| indexing), one axis.
__device__ __forceinline__ float axis_area(int x, int N) {
    if (N <= 1) return 1.0f;
    int H = N / 2;
    int a = abs(x - H);
    return atanf((float)(a + 1) / (float)H) - atanf((float)a / (float)H);
}

__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// ---- diffuse --------------------------------------------------------------------------------------------
// res[t] = post(t) * sum_u pre(u) * src[u] * clamp(D_t . D_u, 0, 0.999)
//   forward : pre = area(u)/3.141592, post = 1        backward: pre = 1, post = area(t)/3.141592
template <bool BWD>
__global__ void __launch_bounds__(256) diffuse_kernel(int R, const float *__restrict__ src, int src_stride,
                                                       float *__restrict__ dst, int dst_stride) {
    extern __shared__ float smem[];
    float *s_area = smem;              // [R]
    float4 *s_dir = reinterpret_cast<float4 *>(smem + ((R + 3) & ~3));  // chunk of source texels: dir.xyz, pre
    float4 *s_val = s_dir + 256;                                        // rgb
    const int total = 6 * R * R;
    const float invN = (float)R;  // texel_dir takes N
    for (int i = threadIdx.x; i < R; i += blockDim.x) s_area[i] = axis_area(i, R);
    __syncthreads();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = t < total;
    int ts = live ? t / (R * R) : 0, ty = live ? (t / R) % R : 0, tx = live ? t % R : 0;
    float3 Dt = texel_dir(tx, ty, ts, invN);
    float3 acc = make_float3(0.f, 0.f, 0.f);
    for (int base = 0; base < total; base += 256) {
        int u = base + threadIdx.x;
        __syncthreads();
        if (u < total) {
            int us = u / (R * R), uy = (u / R) % R, ux = u % R;
            float3 Du = texel_dir(ux, uy, us, invN);
            float pre = BWD ? 1.0f : s_area[ux] * s_area[uy] / 3.141592f;
            s_dir[threadIdx.x] = make_float4(Du.x, Du.y, Du.z, pre);
            const float *p = src + (size_t)u * src_stride;
            s_val[threadIdx.x] = make_float4(p[0], p[1], p[2], 0.f);
        }
        __syncthreads();
        int n = min(256, total - base);
        for (int k = 0; k < n; ++k) {
            float4 d = s_dir[k];
            float4 v = s_val[k];
            float c = fminf(fmaxf(Dt.x * d.x + Dt.y * d.y + Dt.z * d.z, 0.0f), 0.999f);
            float w = c * d.w;
            acc.x += v.x * w; acc.y += v.y * w; acc.z += v.z * w;
        }
    }
    if (!live) return;
    float post = BWD ? s_area[tx] * s_area[ty] / 3.141592f : 1.0f;
    float *o = dst + (size_t)t * dst_stride;
    o[0] = acc.x * post; o[1] = acc.y * post; o[2] = acc.z * post;
}

// ---- specular bounds: on every face, the texel AABB of the cone { L : L . V >= cos(theta_c) } ---------------------------
// Output contract of the plugin's `specular_bounds` (torch_bindings.cpp:168-191): bounds[6,R,R,24] = per source face
// (xmin, xmax, ymin, ymax) as floats, (R-1, 0, R-1, 0) when the cone misses the face.  The plugin finds them by testing
// every texel of every face against every output texel (O(R^4), 137 ms per level at 512^2 on a B200, ~0.8 s before the
// first training step).  Here the box is found from the geometry, one thread per (output texel, face):
//   * on face s a direction is q(fx, fy) = A + fx U + fy W (unnormalised), so L . V >= c reads
//         k(fx, fy) = a + u fx + w fy >= c sqrt(1 + fx^2 + fy^2),   a = A.V, u = U.V, w = W.V:
//     a face whose axis is further than theta_c + acos(1/sqrt 3) from V cannot be touched (most are not);
//   * for one texel row (fy fixed) that is a quadratic inequality in fx: its solution set clipped to [-1, 1] is the
//     row's candidate column interval, solved in double precision for a cone widened by 2e-6;
//   * the interval's ends are then moved to the first / last texel that passes the EXACT fp32 test the gather kernels
//     evaluate (texel_dir + dot3 >= cutoff), so the box is the bounding box of exactly the texels those kernels accept.
// O(R) quadratics + a handful of exact tests per touched face instead of O(R^2) exact tests per face.
__device__ __forceinline__ bool in_cone(int x, int y, int s, float N, float3 V, float cutoff) {
    return dot3(texel_dir(x, y, s, N), V) >= cutoff;
}

__global__ void __launch_bounds__(128) specular_bounds_kernel(int R, float cutoff, float *__restrict__ bounds) {
    int id = blockIdx.x * blockDim.x + threadIdx.x;  // (texel, face)
    if (id >= 36 * R * R) return;
    const int s = id % 6, o = id / 6;
    const int pz = o / (R * R), py = (o / R) % R, px = o % R;
    const float N = (float)R;
    const float3 V = texel_dir(px, py, pz, N);
    int min_x = R - 1, max_x = 0, min_y = R - 1, max_y = 0;     // the empty box

    // frame of face s: q(fx, fy) = A + fx U + fy W
    float ax, ay, az, bx, by, bz, cx, cy, cz;
    gsb_face_point(s, 0.f, 0.f, ax, ay, az);
    gsb_face_point(s, 1.f, 0.f, bx, by, bz);
    gsb_face_point(s, 0.f, 1.f, cx, cy, cz);
    const double a = (double)ax * V.x + (double)ay * V.y + (double)az * V.z;
    const double u = (double)(bx - ax) * V.x + (double)(by - ay) * V.y + (double)(bz - az) * V.z;
    const double w = (double)(cx - ax) * V.x + (double)(cy - ay) * V.y + (double)(cz - az) * V.z;
    const double c = fmax((double)cutoff - 2e-6, -1.0);        // slightly wider cone: candidates, not the verdict
    // quick reject: every direction of a face is within acos(1/sqrt 3) of its axis A
    const double theta = acos(fmin(fmax(c, -1.0), 1.0));
    const bool reachable = theta + 0.9553166181245093 >= 3.141592653589793 ||
                           a >= cos(theta + 0.9553166181245093) - 1e-9;
    if (reachable) {
        const double alpha = u * u - c * c;
        for (int y = 0; y < R; ++y) {
            const double fy = 2.0 * ((y + 0.5) / (double)R) - 1.0;
            const double k0 = a + w * fy, g = 1.0 + fy * fy;
            // { fx : (k0 + u fx)^2 >= c^2 (g + fx^2), k0 + u fx >= 0 } for c >= 0; for c < 0 the second condition drops
            // and the complement of the mirrored cone is wanted -- the plugin's levels never get there (cutoff > 0)
            double lo = -1.0, hi = 1.0;
            if (c > 0.0) {
                const double beta = u * k0, gamma = k0 * k0 - c * c * g;
                const double disc = beta * beta - alpha * gamma;
                if (fabs(alpha) < 1e-14) {                       // linear: 2 beta fx + gamma >= 0, with k0 + u fx >= 0
                    if (fabs(beta) < 1e-300) { if (gamma < 0.0 || k0 < 0.0) continue; }
                    else if (beta > 0.0) lo = fmax(lo, -gamma / (2.0 * beta));
                    else hi = fmin(hi, -gamma / (2.0 * beta));
                } else {
                    if (disc < 0.0) continue;                     // (alpha > 0 always has real roots)
                    const double sq = sqrt(disc);
                    double r1 = (-beta - sq) / alpha, r2 = (-beta + sq) / alpha;
                    if (r1 > r2) { const double t = r1; r1 = r2; r2 = t; }
                    if (alpha < 0.0) {                            // between the roots, on the side where k0 + u fx >= 0
                        if (k0 + u * 0.5 * (r1 + r2) < 0.0) continue;
                        lo = fmax(lo, r1); hi = fmin(hi, r2);
                    } else if (u > 0.0) lo = fmax(lo, r2);        // outside the roots: the ray on which k0 + u fx >= 0
                    else hi = fmin(hi, r1);
                }
                if (lo > hi) continue;
            }
            // candidate columns (one texel of slack each side), then the exact fp32 test decides the ends
            int xl = max(0, (int)floor((lo + 1.0) * 0.5 * R - 0.5) - 1);
            int xr = min(R - 1, (int)ceil((hi + 1.0) * 0.5 * R - 0.5) + 1);
            while (xl <= xr && !in_cone(xl, y, s, N, V, cutoff)) ++xl;
            if (xl > xr) continue;
            while (!in_cone(xr, y, s, N, V, cutoff)) --xr;
            min_x = min(min_x, xl); max_x = max(max_x, xr);
            min_y = min(min_y, y); max_y = max(max_y, y);
        }
    }
    float4 *bptr = reinterpret_cast<float4 *>(bounds) + (size_t)o * 6 + s;
    *bptr = make_float4((float)min_x, (float)max_x, (float)min_y, (float)max_y);
}

__device__ __forceinline__ float ndf_ggx(float alphaSqr, float cosTheta) {
    float c = fminf(fmaxf(cosTheta, 0.0f), 1.0f);
    float d = (c * alphaSqr - c) * c + 1.0f;
    return alphaSqr / (d * d * 3.14159265358979323846f);
}

// ---- specular (cubemap.cu:246-350) -----------------------------------------------------------------------
// forward : out[t] = (sum_u area(u) k(t,u) src[u], sum_u area(u) k(t,u))            [optionally rgb /= wsum]
// backward: gin[t] = area(t) * sum_u k(t,u) * g[u] (/ wsum[u] when the forward normalised)
template <bool BWD>
__global__ void __launch_bounds__(128) specular_kernel(int R, const float *__restrict__ src, int src_stride,
                                                        const float *__restrict__ wsum_src,
                                                        const float4 *__restrict__ bounds, float alphaSqr,
                                                        float cutoff, int normalize, float *__restrict__ dst) {
    extern __shared__ float s_area[];
    for (int i = threadIdx.x; i < R; i += blockDim.x) s_area[i] = axis_area(i, R);
    __syncthreads();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 6 * R * R) return;
    int ts = t / (R * R), ty = (t / R) % R, tx = t % R;
    const float invN = (float)R;  // texel_dir takes N
    float3 V = texel_dir(tx, ty, ts, invN);
    float3 acc = make_float3(0.f, 0.f, 0.f);
    float wsum = 0.f;
    for (int s = 0; s < 6; ++s) {
        float4 b = __ldg(bounds + (size_t)t * 6 + s);
        int xmin = (int)b.x, xmax = (int)b.y, ymin = (int)b.z, ymax = (int)b.w;
        if (xmin > xmax) continue;
        for (int y = ymin; y <= ymax; ++y) {
            float ay = s_area[y];
            for (int x = xmin; x <= xmax; ++x) {
                float3 L = texel_dir(x, y, s, invN);
                float d = dot3(L, V);
                if (!(d >= cutoff)) continue;
                // GGX D(V.H) with H = normalize(L+V).  The reference evaluates 1 - c^2 (1 - a^2) from c = V.H,
                // which cancels catastrophically near c = 1; here sin^2 = |V x L|^2 / |L+V|^2 is formed
                // directly, so the weight is accurate to ~1e-6 instead of ~1e-3.
                float3 Hh = make_float3(L.x + V.x, L.y + V.y, L.z + V.z);
                float3 cr = make_float3(V.y * L.z - V.z * L.y, V.z * L.x - V.x * L.z, V.x * L.y - V.y * L.x);
                float s2 = fminf(dot3(cr, cr) / dot3(Hh, Hh), 1.0f);
                float dd = s2 + (1.0f - s2) * alphaSqr;
                float k = fmaxf(d, 0.f) * (alphaSqr / (dd * dd * 3.14159265358979323846f));
                size_t u = ((size_t)s * R + y) * R + x;
                const float *p = src + u * src_stride;
                float w;
                if (BWD) {
                    w = k / 4.0f;
                    if (normalize) w /= __ldg(wsum_src + u * 4 + 3);
                } else {
                    w = k * (s_area[x] * ay) / 4.0f;
                }
                acc.x += __ldg(p) * w; acc.y += __ldg(p + 1) * w; acc.z += __ldg(p + 2) * w;
                wsum += w;
            }
        }
    }
    if (BWD) {
        float a = s_area[tx] * s_area[ty];
        float *o = dst + (size_t)t * 3;
        o[0] = acc.x * a; o[1] = acc.y * a; o[2] = acc.z * a;
    } else {
        float4 r = normalize ? make_float4(acc.x / wsum, acc.y / wsum, acc.z / wsum, wsum)
                             : make_float4(acc.x, acc.y, acc.z, wsum);
        reinterpret_cast<float4 *>(dst)[t] = r;
    }
}

// ---- specular, warp-cooperative fast path (R % 8 == 0) ----------------------------------------------------
// Same sums as specular_kernel, reorganised for issue efficiency (that kernel spends ~150 instructions per tap
// on IEEE divisions and square roots):
//   * per-texel unit directions + solid angle are tabulated once per call (dirs[u] = L.xyz, area) with the
//     reference-form arithmetic of texel_dir, so the cone membership is unchanged;
//   * the source is pre-multiplied per texel (forward: rgb * area with .w = area, so the same loop also yields
//     wsum; backward: grad / wsum), so forward and backward share ONE gather loop
//         acc[t] += k(t,u) * pre[u],   k = max(L.V,0) * D_ggx / 4
//   * a warp owns an 8x4 patch of output texels and walks the UNION of its lanes' cone AABBs, so every tap is a
//     broadcast 2 x LDG.128 and there is no loop divergence; |L+V|^2 = 2 + 2 L.V for unit vectors.
__global__ void __launch_bounds__(256) dir_table_kernel(int R, float4 *__restrict__ dirs) {
    __shared__ float s_area_tab[1024];
    for (int i = threadIdx.x; i < R; i += blockDim.x) s_area_tab[i] = axis_area(i, R);
    __syncthreads();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 6 * R * R) return;
    int s = t / (R * R), y = (t / R) % R, x = t % R;
    float3 d = texel_dir(x, y, s, (float)R);
    dirs[t] = make_float4(d.x, d.y, d.z, s_area_tab[x] * s_area_tab[y]);
}

// mode 0: forward  pre = (rgb * area, area)      mode 1: backward pre = (g, 0)      mode 2: backward pre = (g / wsum, 0)
__global__ void __launch_bounds__(256) prep_source_kernel(int R, const float *__restrict__ src, int src_stride,
                                                           const float *__restrict__ wsum_src,
                                                           const float4 *__restrict__ dirs, int mode,
                                                           float4 *__restrict__ pre) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 6 * R * R) return;
    const float *p = src + (size_t)t * src_stride;
    float3 v = make_float3(p[0], p[1], p[2]);
    if (mode == 0) {
        float a = dirs[t].w;
        pre[t] = make_float4(v.x * a, v.y * a, v.z * a, a);
    } else {
        float iw = (mode == 2) ? 1.0f / wsum_src[(size_t)t * 4 + 3] : 1.0f;
        pre[t] = make_float4(v.x * iw, v.y * iw, v.z * iw, 0.f);
    }
}

// The traversal shared by the on-the-fly gather, and by the two passes that BUILD a cached plan (below): a warp owns an
// 8x4 patch of output texels and visits, face by face and row by row, the union of its lanes' cone spans; `vis.row(s, y,
// xmin, xmax)` is called warp-uniformly for every non-empty union span.
struct PatchGeom {
    int t;            // this lane's output texel
    float3 V;         // its direction
    float area;
};

template <class Visitor>
__device__ __forceinline__ void walk_cone_rows(int R, const float4 *__restrict__ bounds, float cutoff, const PatchGeom &g,
                                               int s_begin, int s_end, Visitor &vis) {
    const float3 V = g.V;
    // conservative copy of the cone test for the per-row spans below (the exact test on d decides membership)
    const float cs = cutoff * (1.0f - 1e-5f) - 1e-6f;
    const float halfR = 0.5f * (float)R, twoOverR = 2.0f / (float)R;
    for (int s = s_begin; s < s_end; ++s) {
        const float4 b = __ldg(bounds + (size_t)g.t * 6 + s);
        const int lxmin = (int)b.x, lxmax = (int)b.y, lymin = (int)b.z, lymax = (int)b.w;   // this lane's cone AABB
        int ymin = lymin, ymax = lymax;
        if (lxmin > lxmax) { ymin = 1 << 30; ymax = -1; }   // this lane's cone misses face s
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
            ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        }
        vis.face(s);
        if (ymax < ymin) continue;   // warp-uniform
        // face frame: p(gx, gy) = n + gx u + gy w, so p.V = a gx + (gy bw + cn)
        float nx, ny, nz, ux, uy, uz, wx, wy, wz;
        gsb_face_point(s, 0.f, 0.f, nx, ny, nz);
        gsb_face_point(s, 1.f, 0.f, ux, uy, uz);
        gsb_face_point(s, 0.f, 1.f, wx, wy, wz);
        const float a = (ux - nx) * V.x + (uy - ny) * V.y + (uz - nz) * V.z;
        const float bw = (wx - nx) * V.x + (wy - ny) * V.y + (wz - nz) * V.z;
        const float cn = nx * V.x + ny * V.y + nz * V.z;
        const float A = a * a - cs * cs;
        for (int y = ymin; y <= ymax; ++y) {
            // Row span of THIS lane's cone: the texels with (a gx + b0)^2 >= cs^2 (gx^2 + gy^2 + 1), a gx + b0 >= 0 -- a
            // chord of a conic, far tighter than the AABB (the AABB of a disc wastes 1 - pi/4 of its taps).  The
            // span only has to be a superset (one texel of margin); where the quadratic does not describe a bounded
            // interval (A >= 0: the face is steep against the cone axis) the AABB span is kept.
            int xlo = 1 << 30, xhi = -1;
            if (y >= lymin && y <= lymax && lxmin <= lxmax) {
                xlo = lxmin; xhi = lxmax;
                if (cs > 0.f && A < -1e-6f) {
                    const float gy = ((float)y + 0.5f) * twoOverR - 1.0f;
                    const float b0 = gy * bw + cn;
                    const float Bq = 2.0f * a * b0, Cq = b0 * b0 - cs * cs * (gy * gy + 1.0f);
                    const float disc = Bq * Bq - 4.0f * A * Cq;
                    if (disc < 0.f) {
                        xlo = 1 << 30; xhi = -1;
                    } else {
                        const float sq = sqrtf(disc), inv2A = 0.5f / A;
                        const float r1 = (-Bq - sq) * inv2A, r2 = (-Bq + sq) * inv2A;
                        const float glo = fminf(r1, r2), ghi = fmaxf(r1, r2);
                        if (a * (0.5f * (glo + ghi)) + b0 < 0.f) {
                            xlo = 1 << 30; xhi = -1;                       // the mirror branch of the squared test
                        } else {
                            xlo = max(xlo, (int)floorf((glo + 1.0f) * halfR - 0.5f) - 1);
                            xhi = min(xhi, (int)ceilf((ghi + 1.0f) * halfR - 0.5f) + 1);
                        }
                    }
                }
            }
            int xmin = xlo, xmax = xhi;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
                xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
            }
            if (xmax < xmin) continue;   // warp-uniform
            vis.row(s, y, xmin, xmax);
        }
    }
}

__device__ __forceinline__ bool patch_geom(int R, const float4 *__restrict__ dirs, int patch, int lane, PatchGeom &g) {
    const int ppf = (R / 8) * (R / 4);                                 // 8x4 patches, (R/8) x (R/4) per face
    if (patch >= 6 * ppf) return false;
    const int ts = patch / ppf, pr = patch % ppf;
    const int tx = (pr % (R / 8)) * 8 + (lane & 7), ty = (pr / (R / 8)) * 4 + (lane >> 3);
    g.t = (ts * R + ty) * R + tx;
    const float4 Vd = __ldg(dirs + g.t);
    g.V = make_float3(Vd.x, Vd.y, Vd.z);
    g.area = Vd.w;
    return true;
}

// GGX weight of one tap for this lane: k = max(L.V, 0) D_ggx(V.H) / 4 inside the cone, 0 outside.  H bisects the unit
// vectors L and V, so 1 - (V.H)^2 = sin^2(theta/2) = |V - L|^2 / 4: exact differences of nearby unit vectors, no
// cancellation, no division.
__device__ __forceinline__ float tap_weight(const float4 Ld, const float3 V, float alphaSqr, float cutoff, float kscale,
                                            float e_scale) {
    const float3 L = make_float3(Ld.x, Ld.y, Ld.z);
    const float d = dot3(L, V);
    const float ex = V.x - L.x, ey = V.y - L.y, ez = V.z - L.z;
    const float e = ex * ex + ey * ey + ez * ez;
    const float dd = fmaf(e, e_scale, alphaSqr);     // s2 + (1 - s2) * alphaSqr with s2 = e / 4
    return (d >= cutoff) ? d * __fdividef(kscale, dd * dd) : 0.f;
}

struct GatherVisitor {      // on-the-fly: weight evaluated at every tap
    int R;
    const float4 *dirs, *pre;
    float3 V;
    float alphaSqr, cutoff, kscale, e_scale;
    float4 acc;
    __device__ __forceinline__ void face(int) {}
    __device__ __forceinline__ void row(int s, int y, int xmin, int xmax) {
        const float4 *drow = dirs + ((size_t)s * R + y) * R;
        const float4 *prow = pre + ((size_t)s * R + y) * R;
        // branch-free body, 4 taps in flight: the loads are warp-broadcast L1 hits and the four weight
        // evaluations are independent, which is what a latency-bound lone warp needs
#pragma unroll 4
        for (int x = xmin; x <= xmax; ++x) {
            const float4 P = __ldg(prow + x);
            const float k = tap_weight(__ldg(drow + x), V, alphaSqr, cutoff, kscale, e_scale);
            acc.x += P.x * k; acc.y += P.y * k; acc.z += P.z * k; acc.w += P.w * k;
        }
    }
};

// SPLIT == false: one launch dimension, a warp sums all six faces and writes the result.
// SPLIT == true : gridDim.y = 6, a warp sums ONE face and adds into a zeroed float4 accumulator (small levels have
//                 too few 8x4 patches to fill 148 SMs); specular_finalize_kernel then writes the result.
template <bool BWD, bool SPLIT>
__global__ void __launch_bounds__(128) specular_gather_kernel(int R, const float4 *__restrict__ dirs,
                                                               const float4 *__restrict__ pre,
                                                               const float4 *__restrict__ bounds, float alphaSqr,
                                                               float cutoff, int normalize, float *__restrict__ dst,
                                                               float4 *__restrict__ accum) {
    const int lane = threadIdx.x & 31;
    PatchGeom g;
    if (!patch_geom(R, dirs, blockIdx.x * 4 + (threadIdx.x >> 5), lane, g)) return;
    GatherVisitor vis{R, dirs, pre, g.V, alphaSqr, cutoff, alphaSqr * (0.25f / 3.14159265358979323846f),
                      0.25f * (1.0f - alphaSqr), make_float4(0.f, 0.f, 0.f, 0.f)};
    walk_cone_rows(R, bounds, cutoff, g, SPLIT ? (int)blockIdx.y : 0, SPLIT ? (int)blockIdx.y + 1 : 6, vis);
    const float4 acc = vis.acc;
    const int t = g.t;
    if (SPLIT) {
        atomicAdd(accum + t, acc);
        return;
    }
    if (BWD) {
        float *o = dst + (size_t)t * 3;
        o[0] = acc.x * g.area; o[1] = acc.y * g.area; o[2] = acc.z * g.area;
    } else {
        float4 r = normalize ? make_float4(acc.x / acc.w, acc.y / acc.w, acc.z / acc.w, acc.w) : acc;
        reinterpret_cast<float4 *>(dst)[t] = r;
    }
}

// ---- cached plan: the GGX weights are a function of the geometry only ------------------------------------------------
// The cube map changes every training step (it is a trainable parameter, geosplat.py:741-748); roughness, resolution and
// cone do not.  A level's "plan" stores, once, per 8x4 patch the row segments of the union of its cones and per
// (tap, lane) the weight k -- 32 consecutive floats per tap, so the per-step pass streams the weights with perfectly
// coalesced 128-byte loads and does 4 FMAs per tap instead of evaluating a GGX lobe: ~5.5 GB of HBM for the six levels of
// a 512^2 cube map (of 180), read once forward and once backward (the kernel k(V, L) is symmetric, so the backward
// gathers with the same weights).  Built by the same traversal and the same weight expression as the on-the-fly
// kernels, in the same tap order: the results are bit-identical to theirs.
struct CountVisitor {
    int segs, taps;
    int2 *out;           // per (patch, face): {segments, taps}
    int lane;
    __device__ __forceinline__ void face(int s) {
        if (s > 0 && lane == 0) out[s - 1] = make_int2(segs, taps);
        if (s > 0) { segs = 0; taps = 0; }
    }
    __device__ __forceinline__ void row(int, int, int xmin, int xmax) { ++segs; taps += (xmax - xmin + 4) & ~3; }   // padded to whole float4 groups
};

__global__ void __launch_bounds__(128) specular_plan_count_kernel(int R, const float4 *__restrict__ dirs,
                                                                   const float4 *__restrict__ bounds, float cutoff,
                                                                   int2 *__restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int patch = blockIdx.x * 4 + (threadIdx.x >> 5);
    PatchGeom g;
    if (!patch_geom(R, dirs, patch, lane, g)) return;
    CountVisitor vis{0, 0, counts + (size_t)patch * 6, lane};
    walk_cone_rows(R, bounds, cutoff, g, 0, 6, vis);
    if (lane == 0) counts[(size_t)patch * 6 + 5] = make_int2(vis.segs, vis.taps);
}

struct FillVisitor {
    int R;
    const float4 *dirs;
    float3 V;
    float alphaSqr, cutoff, kscale, e_scale;
    const int *seg_start, *tap_start;    // exclusive prefix sums over (patch, face)
    int4 *segs;
    float *weights;
    int lane, seg, tap;
    size_t pf;                            // patch * 6
    __device__ __forceinline__ void face(int s) { seg = seg_start[pf + s]; tap = tap_start[pf + s]; }
    __device__ __forceinline__ void row(int s, int y, int xmin, int xmax) {
        const int row0 = (s * R + y) * R;
        if (lane == 0) segs[seg] = make_int4(row0 + xmin, xmax - xmin + 1, tap, 0);
        ++seg;
        // layout: groups of 4 consecutive taps, [group][lane][4] -- one 16-byte load per lane and group when streaming;
        // every segment starts a new group and its tail is padded with zero weights
        const float4 *drow = dirs + row0;
        const int len = xmax - xmin + 1, padded = (len + 3) & ~3;
        float4 *w4 = reinterpret_cast<float4 *>(weights) + (size_t)(tap >> 2) * 32 + lane;
        for (int j = 0; j < padded; j += 4, w4 += 32) {
            float k[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                k[i] = (j + i < len) ? tap_weight(__ldg(drow + xmin + j + i), V, alphaSqr, cutoff, kscale, e_scale) : 0.f;
            *w4 = make_float4(k[0], k[1], k[2], k[3]);
        }
        tap += padded;
    }
};

__global__ void __launch_bounds__(128) specular_plan_fill_kernel(int R, const float4 *__restrict__ dirs,
                                                                  const float4 *__restrict__ bounds, float alphaSqr,
                                                                  float cutoff, const int *__restrict__ seg_start,
                                                                  const int *__restrict__ tap_start,
                                                                  int4 *__restrict__ segs, float *__restrict__ weights) {
    const int lane = threadIdx.x & 31;
    const int patch = blockIdx.x * 4 + (threadIdx.x >> 5);
    PatchGeom g;
    if (!patch_geom(R, dirs, patch, lane, g)) return;
    FillVisitor vis{R, dirs, g.V, alphaSqr, cutoff, alphaSqr * (0.25f / 3.14159265358979323846f),
                    0.25f * (1.0f - alphaSqr), seg_start, tap_start, segs, weights, lane, 0, 0, (size_t)patch * 6};
    walk_cone_rows(R, bounds, cutoff, g, 0, 6, vis);
}

#ifndef GSB_PLAN_DEPTH
#define GSB_PLAN_DEPTH 4
#endif
#ifndef GSB_PLAN_MINB
#define GSB_PLAN_MINB 1
#endif
#ifndef GSB_PLAN_PARTS_TARGET
#define GSB_PLAN_PARTS_TARGET 49152   // warps per level the split aims for (measured, as_envstack fwd+bwd: 6 k 5.36 ms, 12 k 5.01, 24 k 4.85, 48 k 4.60, 96 k 4.65)
#endif
int plan_parts(int R) {
    const int patches = 6 * R * R / 32;
    int parts = 1;
    while (parts < 16 && patches * parts < GSB_PLAN_PARTS_TARGET) parts *= 2;
    return parts;
}

// acc[t] += sum over the patch's taps of W[tap][lane] * pre[source texel]: same epilogues as the gather kernel
template <bool BWD, bool SPLIT>
__global__ void __launch_bounds__(128, GSB_PLAN_MINB) specular_apply_kernel(int R, const float4 *__restrict__ dirs,
                                                              const float4 *__restrict__ pre,
                                                              const int *__restrict__ seg_start,
                                                              const int4 *__restrict__ segs,
                                                              const float *__restrict__ weights, int normalize,
                                                              float *__restrict__ dst, float4 *__restrict__ accum) {
    const int lane = threadIdx.x & 31;
    const int patch = blockIdx.x * 4 + (threadIdx.x >> 5);
    PatchGeom g;
    if (!patch_geom(R, dirs, patch, lane, g)) return;
    const size_t pf = (size_t)patch * 6;
    int s0 = seg_start[pf], s1 = seg_start[pf + 6];
    if (SPLIT) {
        // an EVEN share of the patch's row segments (a split by source face leaves most of a cone in one part: the
        // coarse levels have few patches with long tap lists, and a level's duration was one warp's serial stream)
        const long long n = s1 - s0;
        const int a = s0 + (int)(n * blockIdx.y / gridDim.y), b = s0 + (int)(n * (blockIdx.y + 1) / gridDim.y);
        s0 = a;
        s1 = b;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int last = 6 * R * R - 1;
    if (s1 > s0) {
        // The weights of consecutive segments are consecutive in memory ([group of 4 taps][lane] float4), so the stream
        // is read through a register ring DEPTH groups deep that does not care where a segment ends: DEPTH 16-byte
        // loads per lane stay in flight (a row segment is only ~6 groups long at 512^2, and a per-segment loop drained
        // the pipeline at every segment: 4.2 TB/s).  The segment headers, one ahead, only say which source texels a
        // group multiplies.
        constexpr int DEPTH = GSB_PLAN_DEPTH;
        int4 sg = __ldg(segs + s0);
        int4 sg_next = (s0 + 1 < s1) ? __ldg(segs + s0 + 1) : sg;
        const int4 sg_last = __ldg(segs + s1 - 1);
        const int g0 = sg.z >> 2, G = ((sg_last.z >> 2) + ((sg_last.y + 3) >> 2)) - g0;   // groups of my share
        const float4 *w4 = reinterpret_cast<const float4 *>(weights) + (size_t)g0 * 32 + lane;
        int seg = s0, left = (sg.y + 3) >> 2, x = sg.x;        // current segment, groups left in it, next source texel
        float4 ring[DEPTH];
#pragma unroll
        for (int j = 0; j < DEPTH; ++j) ring[j] = (j < G) ? __ldcs(w4 + (size_t)j * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int base = 0; base < G; base += DEPTH) {
#pragma unroll
            for (int j = 0; j < DEPTH; ++j) {
                const int gi = base + j;
                const float4 k = ring[j];
                if (gi + DEPTH < G) ring[j] = __ldcs(w4 + (size_t)(gi + DEPTH) * 32);   // streamed once per pass
                if (gi < G) {
                    while (left == 0) {                         // next segment (its header was fetched one ahead)
                        ++seg;
                        sg = sg_next;
                        if (seg + 1 < s1) sg_next = __ldg(segs + seg + 1);
                        left = (sg.y + 3) >> 2;
                        x = sg.x;
                    }
                    // the padded tail of a segment (zero weights) stays inside the buffer
                    const float4 P0 = __ldg(pre + x), P1 = __ldg(pre + min(x + 1, last));
                    const float4 P2 = __ldg(pre + min(x + 2, last)), P3 = __ldg(pre + min(x + 3, last));
                    acc.x += P0.x * k.x; acc.y += P0.y * k.x; acc.z += P0.z * k.x; acc.w += P0.w * k.x;
                    acc.x += P1.x * k.y; acc.y += P1.y * k.y; acc.z += P1.z * k.y; acc.w += P1.w * k.y;
                    acc.x += P2.x * k.z; acc.y += P2.y * k.z; acc.z += P2.z * k.z; acc.w += P2.w * k.z;
                    acc.x += P3.x * k.w; acc.y += P3.y * k.w; acc.z += P3.z * k.w; acc.w += P3.w * k.w;
                    x += 4;
                    --left;
                }
            }
        }
    }
    const int t = g.t;
    if (SPLIT) {
        atomicAdd(accum + t, acc);
        return;
    }
    if (BWD) {
        float *o = dst + (size_t)t * 3;
        o[0] = acc.x * g.area; o[1] = acc.y * g.area; o[2] = acc.z * g.area;
    } else {
        float4 r = normalize ? make_float4(acc.x / acc.w, acc.y / acc.w, acc.z / acc.w, acc.w) : acc;
        reinterpret_cast<float4 *>(dst)[t] = r;
    }
}

template <bool BWD>
__global__ void __launch_bounds__(256) specular_finalize_kernel(int R, const float4 *__restrict__ accum,
                                                                 const float4 *__restrict__ dirs, int normalize,
                                                                 float *__restrict__ dst) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 6 * R * R) return;
    const float4 acc = accum[t];
    if (BWD) {
        const float a = dirs[t].w;
        float *o = dst + (size_t)t * 3;
        o[0] = acc.x * a; o[1] = acc.y * a; o[2] = acc.z * a;
    } else {
        float4 r = normalize ? make_float4(acc.x / acc.w, acc.y / acc.w, acc.z / acc.w, acc.w) : acc;
        reinterpret_cast<float4 *>(dst)[t] = r;
    }
}

// ---- mip chain (_texture.py:199-226) -------------------------------------------------------------------
__global__ void __launch_bounds__(256) mip_fwd_kernel(int Ro, const float *__restrict__ in, int in_stride,
                                                       float *__restrict__ out, int out_stride) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 6 * Ro * Ro) return;
    int s = t / (Ro * Ro), y = (t / Ro) % Ro, x = t % Ro;
    int Ri = Ro * 2;
    const float *p00 = in + (((size_t)s * Ri + 2 * y) * Ri + 2 * x) * in_stride;
    const float *p10 = p00 + in_stride, *p01 = p00 + (size_t)Ri * in_stride, *p11 = p01 + in_stride;
    float *o = out + (size_t)t * out_stride;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = (p00[c] + p10[c] + p01[c] + p11[c]) * 0.25f;
}

// d(in)[fine texel] = bilinear cube sample of 0.25 * d(out) along the fine texel's direction.
__global__ void __launch_bounds__(256) mip_bwd_kernel(int Ro, const float *__restrict__ dout, float *__restrict__ din) {
    int Ri = Ro * 2;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 6 * Ri * Ri) return;
    int s = t / (Ri * Ri), y = (t / Ri) % Ri, x = t % Ri;
    // _texture.py:213-219: linspace(-1 + 1/res, 1 - 1/res, res) == texel centres
    float3 d = texel_dir(x, y, s, (float)Ri);
    CubeTaps taps = gsb_cube_taps(d.x, d.y, d.z, Ro);
    float3 v = gsb_cube_sample<3, false>(dout, taps, nullptr, nullptr);
    float *o = din + (size_t)t * 3;
    o[0] = 0.25f * v.x; o[1] = 0.25f * v.y; o[2] = 0.25f * v.z;
}

}  // namespace

#define GSB_API extern "C" __attribute__((visibility("default")))

GSB_API int gsb_diffuse_cubemap_fwd(int32_t R, const float *cubemap, float *out, int32_t out_stride, void *stream) {
    GSB_CHECK_ARG(R >= 1 && R <= 64 && cubemap && out && (out_stride == 3 || out_stride == 4));
    int total = 6 * R * R;
    size_t smem = sizeof(float) * ((R + 3) & ~3) + 2 * 256 * sizeof(float4);
    diffuse_kernel<false><<<gsb_div_up(total, 256), 256, smem, (cudaStream_t)stream>>>(R, cubemap, 3, out, out_stride);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_diffuse_cubemap_bwd(int32_t R, const float *grad_out, int32_t grad_stride, float *grad_in,
                                    void *stream) {
    GSB_CHECK_ARG(R >= 1 && R <= 64 && grad_out && grad_in && (grad_stride == 3 || grad_stride == 4));
    int total = 6 * R * R;
    size_t smem = sizeof(float) * ((R + 3) & ~3) + 2 * 256 * sizeof(float4);
    diffuse_kernel<true><<<gsb_div_up(total, 256), 256, smem, (cudaStream_t)stream>>>(R, grad_out, grad_stride,
                                                                                      grad_in, 3);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_specular_bounds(int32_t R, float costheta_cutoff, float *bounds, void *stream) {
    GSB_CHECK_ARG(R >= 1 && R <= 4096 && bounds);
    long long n = 36LL * R * R;
    specular_bounds_kernel<<<gsb_div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(R, costheta_cutoff, bounds);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_specular_workspace_bytes(int32_t R, size_t *bytes_host) {
    GSB_CHECK_ARG(R >= 1 && R <= 4096 && bytes_host != nullptr);
    *bytes_host = 3 * sizeof(float4) * 6 * (size_t)R * R + 512;
    return GSB_OK;
}

static float4 *ws_align(void *ws) {
    return reinterpret_cast<float4 *>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
}

GSB_API int gsb_specular_cubemap_fwd(int32_t R, const float *cubemap, const float *bounds, float roughness,
                                     float costheta_cutoff, int32_t normalize, float *out, void *workspace,
                                     void *stream) {
    GSB_CHECK_ARG(R >= 1 && R <= 4096 && cubemap && bounds && out);
    float alpha = roughness * roughness;
    cudaStream_t st = (cudaStream_t)stream;
    if (workspace != nullptr && R % 8 == 0 && R <= 1024) {
        float4 *dirs = ws_align(workspace), *pre = dirs + 6 * (size_t)R * R;
        int total = 6 * R * R;
        dir_table_kernel<<<gsb_div_up(total, 256), 256, 0, st>>>(R, dirs);
        prep_source_kernel<<<gsb_div_up(total, 256), 256, 0, st>>>(R, cubemap, 3, nullptr, dirs, 0, pre);
        if (R <= 64) {
            float4 *accum = pre + 6 * (size_t)R * R;
            GSB_CHECK_CUDA(cudaMemsetAsync(accum, 0, sizeof(float4) * total, st));
            specular_gather_kernel<false, true><<<dim3(gsb_div_up(total / 32, 4), 6), 128, 0, st>>>(
                R, dirs, pre, reinterpret_cast<const float4 *>(bounds), alpha * alpha, costheta_cutoff, normalize, out,
                accum);
            specular_finalize_kernel<false><<<gsb_div_up(total, 256), 256, 0, st>>>(R, accum, dirs, normalize, out);
        } else {
            specular_gather_kernel<false, false><<<gsb_div_up(total / 32, 4), 128, 0, st>>>(
                R, dirs, pre, reinterpret_cast<const float4 *>(bounds), alpha * alpha, costheta_cutoff, normalize, out,
                nullptr);
        }
    } else {
        specular_kernel<false><<<gsb_div_up(6 * R * R, 128), 128, sizeof(float) * R, st>>>(
            R, cubemap, 3, nullptr, reinterpret_cast<const float4 *>(bounds), alpha * alpha, costheta_cutoff, normalize,
            out);
    }
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_specular_cubemap_bwd(int32_t R, const float *bounds, const float *grad_out, const float *fwd_out,
                                     float roughness, float costheta_cutoff, float *grad_in, void *workspace,
                                     void *stream) {
    GSB_CHECK_ARG(R >= 1 && R <= 4096 && bounds && grad_out && grad_in);
    float alpha = roughness * roughness;
    cudaStream_t st = (cudaStream_t)stream;
    if (workspace != nullptr && R % 8 == 0 && R <= 1024) {
        float4 *dirs = ws_align(workspace), *pre = dirs + 6 * (size_t)R * R;
        int total = 6 * R * R;
        dir_table_kernel<<<gsb_div_up(total, 256), 256, 0, st>>>(R, dirs);
        prep_source_kernel<<<gsb_div_up(total, 256), 256, 0, st>>>(R, grad_out, 4, fwd_out, dirs, fwd_out ? 2 : 1, pre);
        if (R <= 64) {
            float4 *accum = pre + 6 * (size_t)R * R;
            GSB_CHECK_CUDA(cudaMemsetAsync(accum, 0, sizeof(float4) * total, st));
            specular_gather_kernel<true, true><<<dim3(gsb_div_up(total / 32, 4), 6), 128, 0, st>>>(
                R, dirs, pre, reinterpret_cast<const float4 *>(bounds), alpha * alpha, costheta_cutoff, 0, grad_in, accum);
            specular_finalize_kernel<true><<<gsb_div_up(total, 256), 256, 0, st>>>(R, accum, dirs, 0, grad_in);
        } else {
            specular_gather_kernel<true, false><<<gsb_div_up(total / 32, 4), 128, 0, st>>>(
                R, dirs, pre, reinterpret_cast<const float4 *>(bounds), alpha * alpha, costheta_cutoff, 0, grad_in,
                nullptr);
        }
    } else {
        specular_kernel<true><<<gsb_div_up(6 * R * R, 128), 128, sizeof(float) * R, st>>>(
            R, grad_out, 4, fwd_out, reinterpret_cast<const float4 *>(bounds), alpha * alpha, costheta_cutoff,
            fwd_out != nullptr, grad_in);
    }
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

// ---- cached plan entry points -------------------------------------------------------------------------------------
// Build: gsb_specular_plan_count -> counts[(6 R^2 / 32) * 6][2] = {segments, taps} per (8x4 patch, face); the caller
// forms the exclusive prefix sums seg_start / tap_start (one more entry than counts) and allocates segs[n_segs][4] int32
// and weights[n_taps * 32] floats; gsb_specular_plan_fill writes them.  Per step: gsb_specular_plan_fwd / _bwd.
static bool plan_ok(int32_t R) { return R % 8 == 0 && R >= 8 && R <= 1024; }

GSB_API int gsb_specular_plan_count(int32_t R, const float *bounds, float costheta_cutoff, int32_t *counts,
                                    void *workspace, void *stream) {
    GSB_CHECK_ARG(plan_ok(R) && bounds && counts && workspace);
    cudaStream_t st = (cudaStream_t)stream;
    float4 *dirs = ws_align(workspace);
    const int total = 6 * R * R;
    dir_table_kernel<<<gsb_div_up(total, 256), 256, 0, st>>>(R, dirs);
    specular_plan_count_kernel<<<gsb_div_up(total / 32, 4), 128, 0, st>>>(
        R, dirs, reinterpret_cast<const float4 *>(bounds), costheta_cutoff, reinterpret_cast<int2 *>(counts));
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_specular_plan_fill(int32_t R, const float *bounds, float roughness, float costheta_cutoff,
                                   const int32_t *seg_start, const int32_t *tap_start, int32_t *segs, float *weights,
                                   void *workspace, void *stream) {
    GSB_CHECK_ARG(plan_ok(R) && bounds && seg_start && tap_start && segs && weights && workspace);
    cudaStream_t st = (cudaStream_t)stream;
    float4 *dirs = ws_align(workspace);
    const int total = 6 * R * R;
    const float alpha = roughness * roughness;
    dir_table_kernel<<<gsb_div_up(total, 256), 256, 0, st>>>(R, dirs);
    specular_plan_fill_kernel<<<gsb_div_up(total / 32, 4), 128, 0, st>>>(
        R, dirs, reinterpret_cast<const float4 *>(bounds), alpha * alpha, costheta_cutoff, seg_start, tap_start,
        reinterpret_cast<int4 *>(segs), weights);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

template <bool BWD>
static int plan_apply(int32_t R, const float *src, int src_stride, const float *fwd_out, int mode, const int32_t *seg_start,
                      const int32_t *segs, const float *weights, int normalize, float *dst, void *workspace,
                      cudaStream_t st) {
    float4 *dirs = ws_align(workspace), *pre = dirs + 6 * (size_t)R * R;
    const int total = 6 * R * R;
    dir_table_kernel<<<gsb_div_up(total, 256), 256, 0, st>>>(R, dirs);
    prep_source_kernel<<<gsb_div_up(total, 256), 256, 0, st>>>(R, src, src_stride, fwd_out, dirs, mode, pre);
    // parts per patch: enough warps to keep every SM streaming (a 64^2 level has 768 patches of ~2 500 taps)
    const int parts = plan_parts(R);
    if (parts > 1) {
        float4 *accum = pre + 6 * (size_t)R * R;
        GSB_CHECK_CUDA(cudaMemsetAsync(accum, 0, sizeof(float4) * total, st));
        specular_apply_kernel<BWD, true><<<dim3(gsb_div_up(total / 32, 4), parts), 128, 0, st>>>(
            R, dirs, pre, seg_start, reinterpret_cast<const int4 *>(segs), weights, normalize, dst, accum);
        specular_finalize_kernel<BWD><<<gsb_div_up(total, 256), 256, 0, st>>>(R, accum, dirs, normalize, dst);
    } else {
        specular_apply_kernel<BWD, false><<<gsb_div_up(total / 32, 4), 128, 0, st>>>(
            R, dirs, pre, seg_start, reinterpret_cast<const int4 *>(segs), weights, normalize, dst, nullptr);
    }
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_specular_plan_fwd(int32_t R, const float *cubemap, const int32_t *seg_start, const int32_t *segs,
                                  const float *weights, int32_t normalize, float *out, void *workspace, void *stream) {
    GSB_CHECK_ARG(plan_ok(R) && cubemap && seg_start && segs && weights && out && workspace);
    return plan_apply<false>(R, cubemap, 3, nullptr, 0, seg_start, segs, weights, normalize, out, workspace,
                             (cudaStream_t)stream);
}

GSB_API int gsb_specular_plan_bwd(int32_t R, const int32_t *seg_start, const int32_t *segs, const float *weights,
                                  const float *grad_out, const float *fwd_out, float *grad_in, void *workspace,
                                  void *stream) {
    GSB_CHECK_ARG(plan_ok(R) && seg_start && segs && weights && grad_out && grad_in && workspace);
    return plan_apply<true>(R, grad_out, 4, fwd_out, fwd_out ? 2 : 1, seg_start, segs, weights, 0, grad_in, workspace,
                            (cudaStream_t)stream);
}

GSB_API int gsb_cubemap_mip_fwd(int32_t R_out, const float *in, int32_t in_stride, float *out, int32_t out_stride,
                                void *stream) {
    GSB_CHECK_ARG(R_out >= 1 && in && out && (in_stride == 3 || in_stride == 4) && (out_stride == 3 || out_stride == 4));
    mip_fwd_kernel<<<gsb_div_up(6 * R_out * R_out, 256), 256, 0, (cudaStream_t)stream>>>(R_out, in, in_stride, out,
                                                                                         out_stride);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_cubemap_mip_bwd(int32_t R_out, const float *grad_out, float *grad_in, void *stream) {
    GSB_CHECK_ARG(R_out >= 1 && grad_out && grad_in);
    mip_bwd_kernel<<<gsb_div_up(24 * R_out * R_out, 256), 256, 0, (cudaStream_t)stream>>>(R_out, grad_out, grad_in);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}
