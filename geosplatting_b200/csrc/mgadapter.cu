// MGAdaptor mesh -> Gaussian sampling, area-weighted vertex normals, and the naive tone map.
//
// Replaces (same arithmetic, one kernel each instead of ~100 tiny elementwise/gather launches):
//   MGAdapter.make / bary2gs            rfstudio/model/geosplat.py:390-472
//   rot2quat / safe_normalize           rfstudio/graphics/math.py:246-278, :119-128
//   TriangleMesh.compute_vertex_normals_(fix=True)   rfstudio/graphics/_mesh/_triangle_mesh.py:588-614
//   _tone_mapping_naive                 rfstudio/model/geosplat.py:474-476
//
// The MGAdaptor backward is evaluated in FORWARD mode on dual numbers (9 partials: the face's three
// positions for means/scales/quats, its three vertex normals for the interpolated normals): the VJP is
// accumulated on the fly as every output is produced, so there is no hand-derived reverse chain through
// rot2quat's branch selection to get wrong, and the cost (~10x the forward flops per face) is irrelevant
// next to the HBM traffic of a streaming per-face kernel.
#include "gsb_common.cuh"

namespace {

// Sum of `v` over the 256 threads of the block, added to *out by one atomic.
__device__ __forceinline__ void block_add(float v, float *out) {
#ifndef GSB_HOST_EMULATION   // tests/emu runs the threads one after another: every thread adds its own term
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __shared__ float s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x != 0) return;
    v = 0.f;
    for (int w = 0; w < 8; ++w) v += s[w];
#endif
    atomicAdd(out, v);
}


template <int N>
struct Dual {
    float v;
    float d[N];
};

template <int N> __device__ __forceinline__ Dual<N> mk(float v) {
    Dual<N> r; r.v = v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = 0.f;
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> var(float v, int k) { Dual<N> r = mk<N>(v); r.d[k] = 1.f; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator+(Dual<N> a, Dual<N> b) {
    Dual<N> r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator-(Dual<N> a, Dual<N> b) {
    Dual<N> r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator*(Dual<N> a, Dual<N> b) {
    Dual<N> r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator*(Dual<N> a, float s) {
    Dual<N> r; r.v = a.v * s;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s;
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator/(Dual<N> a, Dual<N> b) {
    Dual<N> r; r.v = a.v / b.v;
    float ib = 1.0f / b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> dsqrt(Dual<N> a) {
    Dual<N> r; r.v = sqrtf(a.v);
    float h = 0.5f / r.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * h;
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> dlog(Dual<N> a) {
    Dual<N> r; r.v = logf(a.v);
    float h = 1.0f / a.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * h;
    return r;
}
// torch clamp(min=m): gradient passes where a >= m
template <int N> __device__ __forceinline__ Dual<N> dclamp_min(Dual<N> a, float m) { return (a.v >= m) ? a : mk<N>(m); }

// plain float "dual" so the same code yields the forward
struct F0 { float v; };
__device__ __forceinline__ F0 operator+(F0 a, F0 b) { return {a.v + b.v}; }
__device__ __forceinline__ F0 operator-(F0 a, F0 b) { return {a.v - b.v}; }
__device__ __forceinline__ F0 operator*(F0 a, F0 b) { return {a.v * b.v}; }
__device__ __forceinline__ F0 operator*(F0 a, float s) { return {a.v * s}; }
__device__ __forceinline__ F0 operator/(F0 a, F0 b) { return {a.v / b.v}; }
__device__ __forceinline__ F0 dsqrt(F0 a) { return {sqrtf(a.v)}; }
__device__ __forceinline__ F0 dlog(F0 a) { return {logf(a.v)}; }
__device__ __forceinline__ F0 dclamp_min(F0 a, float m) { return {fmaxf(a.v, m)}; }
template <class T> __device__ __forceinline__ T cst(float v);
template <> __device__ __forceinline__ F0 cst<F0>(float v) { return {v}; }
template <> __device__ __forceinline__ Dual<9> cst<Dual<9>>(float v) { return mk<9>(v); }

template <class T> struct V3 { T x, y, z; };
template <class T> __device__ __forceinline__ V3<T> operator+(V3<T> a, V3<T> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> __device__ __forceinline__ V3<T> operator-(V3<T> a, V3<T> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> __device__ __forceinline__ V3<T> operator*(V3<T> a, float s) { return {a.x * s, a.y * s, a.z * s}; }
template <class T> __device__ __forceinline__ V3<T> cross(V3<T> a, V3<T> b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class T> __device__ __forceinline__ T norm(V3<T> a) { return dsqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
template <class T> __device__ __forceinline__ V3<T> vdiv(V3<T> a, T s) { return {a.x / s, a.y / s, a.z / s}; }

// math.py:119-128
template <class T> __device__ __forceinline__ V3<T> safe_normalize(V3<T> a) {
    T len = norm(a);
    if (len.v < 1e-6f) return {cst<T>(0.f), cst<T>(0.f), cst<T>(1.f)};
    return vdiv(a, dclamp_min(len, 1e-6f));
}

struct MGConst {
    float c1[2], c2[2];     // (1 - 2 u), u per ring        geosplat.py:447,:451-453
    float a_coeff[2];       // area coefficient per ring    :448
    float s_max[2], s_min[2];  // g_scale_ratio * s, g_scale_ratio / s    :405-406
    float logit_opacity;    // logit(0.99)                  :422
};

__host__ MGConst make_consts() {
    MGConst k;
    double u[2] = {1.0 / 9.0 + (-1.0 / 24.0), 2.0 / 9.0 + 0.0};
    double a[2] = {1.0 / 4.0 * (1.0 / 3.0), 1.0 / 12.0 * 3.0};
    double s[2] = {0.5, 1.3};
    for (int r = 0; r < 2; ++r) {
        k.c1[r] = (float)(1.0 - 2.0 * u[r]);
        k.c2[r] = (float)u[r];
        k.a_coeff[r] = (float)a[r];
        k.s_max[r] = (float)(1.6 * s[r]);
        k.s_min[r] = (float)(1.6 / s[r]);
    }
    // torch: empty.fill_(0.99).logit() in fp32
    float p = 0.99f;
    k.logit_opacity = logf(p / (1.0f - p));
    return k;
}

// rot2quat (math.py:246-278) for R = [c0 | c1 | c2] (columns); returns wxyz.
template <class T> __device__ __forceinline__ void rot2quat(V3<T> c0, V3<T> c1, V3<T> c2, T q[4]) {
    T m00 = c0.x, m10 = c0.y, m20 = c0.z, m01 = c1.x, m11 = c1.y, m21 = c1.z, m02 = c2.x, m12 = c2.y, m22 = c2.z;
    T one = cst<T>(1.0f);
    T qq[4] = {one + m00 + m11 + m22, one + m00 - m11 - m22, one - m00 + m11 - m22, one - m00 - m11 + m22};
    T qa[4];
    int best = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        qa[i] = (qq[i].v > 0.f) ? dsqrt(qq[i]) : cst<T>(0.f);
    }
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (qa[i].v > qa[best].v) best = i;
    T cand[4];
    T sq;
    if (best == 0) { sq = qa[0] * qa[0]; cand[0] = sq; cand[1] = m21 - m12; cand[2] = m02 - m20; cand[3] = m10 - m01; }
    else if (best == 1) { sq = qa[1] * qa[1]; cand[0] = m21 - m12; cand[1] = sq; cand[2] = m10 + m01; cand[3] = m02 + m20; }
    else if (best == 2) { sq = qa[2] * qa[2]; cand[0] = m02 - m20; cand[1] = m10 + m01; cand[2] = sq; cand[3] = m12 + m21; }
    else { sq = qa[3] * qa[3]; cand[0] = m10 - m01; cand[1] = m20 + m02; cand[2] = m21 + m12; cand[3] = sq; }
    // 2 * max(q_abs, 0.1): torch `max` with a tensor passes the gradient to the larger argument
    T den = (qa[best].v > 0.1f) ? qa[best] * 2.0f : cst<T>(0.2f);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = cand[i] / den;
}

// One Gaussian of bary2gs (geosplat.py:390-424).  Emits mean[3], log-scales[2], quat[4] through `emit(slot, value)`.
template <class T, class Emit>
__device__ __forceinline__ void bary2gs(V3<T> pa, V3<T> pb, T area, V3<T> n, float s_max, float s_min, Emit emit) {
    V3<T> mean = (pa + pb) * 0.5f;
    V3<T> mr = pb - mean;
    T max_s = dclamp_min(norm(mr), 1e-10f);
    T min_s = (area * 0.25f) / max_s;     // area / 4 / max_scales
    V3<T> max_rot = vdiv(mr, max_s);
    V3<T> min_rot = cross(n, max_rot);
    T q[4];
    rot2quat(max_rot, min_rot, n, q);
    emit(0, mean.x); emit(1, mean.y); emit(2, mean.z);
    emit(3, dlog(max_s * s_max)); emit(4, dlog(min_s * s_min));
    emit(5, q[0]); emit(6, q[1]); emit(7, q[2]); emit(8, q[3]);
}

__device__ __forceinline__ float3 ld3(const float *p, long long i) { return make_float3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }

__global__ void __launch_bounds__(128) mgadapter_fwd_kernel(int F, const float *__restrict__ verts,
                                                             const float *__restrict__ vnormals,
                                                             const long long *__restrict__ faces, MGConst k,
                                                             float *__restrict__ means, float *__restrict__ scales,
                                                             float *__restrict__ quats, float *__restrict__ normals,
                                                             float *__restrict__ opacities, float *__restrict__ offsets) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    long long i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    float3 a = ld3(verts, i0), b = ld3(verts, i1), c = ld3(verts, i2);
    V3<F0> p0 = {{a.x}, {a.y}, {a.z}}, p1 = {{b.x}, {b.y}, {b.z}}, p2 = {{c.x}, {c.y}, {c.z}};
    V3<F0> nrm = cross(p1 - p0, p2 - p0);
    F0 area = dclamp_min(norm(nrm), 1e-10f) * 0.5f;
    V3<F0> n = safe_normalize(nrm);
    float sa = sqrtf(area.v);
    float3 na = make_float3(0, 0, 1), nb = na, nc = na;
    if (vnormals) { na = ld3(vnormals, i0); nb = ld3(vnormals, i1); nc = ld3(vnormals, i2); }
    V3<F0> vn0 = {{na.x}, {na.y}, {na.z}}, vn1 = {{nb.x}, {nb.y}, {nb.z}}, vn2 = {{nc.x}, {nc.y}, {nc.z}};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        V3<F0> u[3] = {p0 * k.c1[r] + (p1 + p2) * k.c2[r], p1 * k.c1[r] + (p2 + p0) * k.c2[r],
                       p2 * k.c1[r] + (p0 + p1) * k.c2[r]};
        V3<F0> nn[3] = {vn0 * k.c1[r] + (vn1 + vn2) * k.c2[r], vn1 * k.c1[r] + (vn2 + vn0) * k.c2[r],
                        vn2 * k.c1[r] + (vn0 + vn1) * k.c2[r]};
        F0 ar = area * k.a_coeff[r];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            long long g = (long long)(r * 3 + e) * F + f;
            float out[9];
            bary2gs(u[e], u[(e + 1) % 3], ar, n, k.s_max[r], k.s_min[r], [&](int slot, F0 v) { out[slot] = v.v; });
            means[3 * g] = out[0]; means[3 * g + 1] = out[1]; means[3 * g + 2] = out[2];
            scales[3 * g] = out[3]; scales[3 * g + 1] = out[4]; scales[3 * g + 2] = -10.0f;
            reinterpret_cast<float4 *>(quats)[g] = make_float4(out[5], out[6], out[7], out[8]);
            V3<F0> col = vnormals ? safe_normalize((nn[e] + nn[(e + 1) % 3]) * 0.5f) : n;
            normals[3 * g] = col.x.v; normals[3 * g + 1] = col.y.v; normals[3 * g + 2] = col.z.v;
            opacities[g] = k.logit_opacity;
            offsets[3 * g] = n.x.v * sa; offsets[3 * g + 1] = n.y.v * sa; offsets[3 * g + 2] = n.z.v * sa;
        }
    }
}

__global__ void __launch_bounds__(128) mgadapter_bwd_kernel(int F, const float *__restrict__ verts,
                                                             const float *__restrict__ vnormals,
                                                             const long long *__restrict__ faces, MGConst k,
                                                             const float *__restrict__ v_means,
                                                             const float *__restrict__ v_scales,
                                                             const float *__restrict__ v_quats,
                                                             const float *__restrict__ v_normals,
                                                             float *__restrict__ v_verts, float *__restrict__ v_vnormals) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    typedef Dual<9> D;
    long long idx[3] = {faces[3 * f], faces[3 * f + 1], faces[3 * f + 2]};
    float acc[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i] = 0.f;
    {   // ---- geometry: partials w.r.t. (p0, p1, p2)
        float3 a = ld3(verts, idx[0]), b = ld3(verts, idx[1]), c = ld3(verts, idx[2]);
        V3<D> p0 = {var<9>(a.x, 0), var<9>(a.y, 1), var<9>(a.z, 2)};
        V3<D> p1 = {var<9>(b.x, 3), var<9>(b.y, 4), var<9>(b.z, 5)};
        V3<D> p2 = {var<9>(c.x, 6), var<9>(c.y, 7), var<9>(c.z, 8)};
        V3<D> nrm = cross(p1 - p0, p2 - p0);
        D area = dclamp_min(norm(nrm), 1e-10f) * 0.5f;
        V3<D> n = safe_normalize(nrm);
        bool flat = (vnormals == nullptr);
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
            V3<D> u[3] = {p0 * k.c1[r] + (p1 + p2) * k.c2[r], p1 * k.c1[r] + (p2 + p0) * k.c2[r],
                          p2 * k.c1[r] + (p0 + p1) * k.c2[r]};
            D ar = area * k.a_coeff[r];
#pragma unroll 1
            for (int e = 0; e < 3; ++e) {
                long long g = (long long)(r * 3 + e) * F + f;
                float w[9] = {v_means[3 * g], v_means[3 * g + 1], v_means[3 * g + 2], v_scales[3 * g], v_scales[3 * g + 1],
                              v_quats[4 * g], v_quats[4 * g + 1], v_quats[4 * g + 2], v_quats[4 * g + 3]};
                bary2gs(u[e], u[(e + 1) % 3], ar, n, k.s_max[r], k.s_min[r], [&](int slot, D v) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) acc[i] += w[slot] * v.d[i];
                });
                if (flat) {  // colours = face normal
                    float wn[3] = {v_normals[3 * g], v_normals[3 * g + 1], v_normals[3 * g + 2]};
#pragma unroll
                    for (int i = 0; i < 9; ++i) acc[i] += wn[0] * n.x.d[i] + wn[1] * n.y.d[i] + wn[2] * n.z.d[i];
                }
            }
        }
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int c2 = 0; c2 < 3; ++c2) atomicAdd(v_verts + 3 * idx[v] + c2, acc[3 * v + c2]);
    }
    if (vnormals == nullptr || v_vnormals == nullptr) return;
    {   // ---- interpolated normals: partials w.r.t. (vn0, vn1, vn2)
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] = 0.f;
        float3 a = ld3(vnormals, idx[0]), b = ld3(vnormals, idx[1]), c = ld3(vnormals, idx[2]);
        V3<D> vn0 = {var<9>(a.x, 0), var<9>(a.y, 1), var<9>(a.z, 2)};
        V3<D> vn1 = {var<9>(b.x, 3), var<9>(b.y, 4), var<9>(b.z, 5)};
        V3<D> vn2 = {var<9>(c.x, 6), var<9>(c.y, 7), var<9>(c.z, 8)};
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
            V3<D> nn[3] = {vn0 * k.c1[r] + (vn1 + vn2) * k.c2[r], vn1 * k.c1[r] + (vn2 + vn0) * k.c2[r],
                           vn2 * k.c1[r] + (vn0 + vn1) * k.c2[r]};
#pragma unroll 1
            for (int e = 0; e < 3; ++e) {
                long long g = (long long)(r * 3 + e) * F + f;
                V3<D> col = safe_normalize((nn[e] + nn[(e + 1) % 3]) * 0.5f);
                float wn[3] = {v_normals[3 * g], v_normals[3 * g + 1], v_normals[3 * g + 2]};
#pragma unroll
                for (int i = 0; i < 9; ++i) acc[i] += wn[0] * col.x.d[i] + wn[1] * col.y.d[i] + wn[2] * col.z.d[i];
            }
        }
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int c2 = 0; c2 < 3; ++c2) atomicAdd(v_vnormals + 3 * idx[v] + c2, acc[3 * v + c2]);
    }
}

// ---- vertex normals (_triangle_mesh.py:588-614) ----------------------------------------------------------
__global__ void __launch_bounds__(256) vnormal_scatter_kernel(int F, const float *__restrict__ verts,
                                                               const long long *__restrict__ faces,
                                                               float *__restrict__ raw) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    long long i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    float3 a = ld3(verts, i0), b = ld3(verts, i1), c = ld3(verts, i2);
    float3 e1 = make_float3(b.x - a.x, b.y - a.y, b.z - a.z), e2 = make_float3(c.x - a.x, c.y - a.y, c.z - a.z);
    float3 w = make_float3(e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x);
    long long id[3] = {i0, i1, i2};
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        atomicAdd(raw + 3 * id[v], w.x); atomicAdd(raw + 3 * id[v] + 1, w.y); atomicAdd(raw + 3 * id[v] + 2, w.z);
    }
}

// fwd: normals = raw/len (or (0,0,1));  bwd (g != nullptr): g_raw = (g - n (n.g)) / len, in place into out
__global__ void __launch_bounds__(256) vnormal_normalize_kernel(int V, const float *__restrict__ raw,
                                                                 const float *__restrict__ g, float *__restrict__ out) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    float3 r = ld3(raw, v);
    float len = sqrtf(r.x * r.x + r.y * r.y + r.z * r.z);
    bool ok = len > 1e-10f;
    float il = 1.0f / fmaxf(len, 1e-10f);
    float3 n = ok ? make_float3(r.x * il, r.y * il, r.z * il) : make_float3(0.f, 0.f, 1.f);
    if (g == nullptr) { out[3 * v] = n.x; out[3 * v + 1] = n.y; out[3 * v + 2] = n.z; return; }
    float3 gv = ld3(g, v);
    float d = n.x * gv.x + n.y * gv.y + n.z * gv.z;
    float3 o = ok ? make_float3((gv.x - n.x * d) * il, (gv.y - n.y * d) * il, (gv.z - n.z * d) * il)
                  : make_float3(0.f, 0.f, 0.f);
    out[3 * v] = o.x; out[3 * v + 1] = o.y; out[3 * v + 2] = o.z;
}

__global__ void __launch_bounds__(256) vnormal_bwd_faces_kernel(int F, const float *__restrict__ verts,
                                                                 const long long *__restrict__ faces,
                                                                 const float *__restrict__ g_raw,
                                                                 float *__restrict__ v_verts) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    long long i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    float3 a = ld3(verts, i0), b = ld3(verts, i1), c = ld3(verts, i2);
    float3 e1 = make_float3(b.x - a.x, b.y - a.y, b.z - a.z), e2 = make_float3(c.x - a.x, c.y - a.y, c.z - a.z);
    float3 g0 = ld3(g_raw, i0), g1 = ld3(g_raw, i1), g2 = ld3(g_raw, i2);
    float3 vw = make_float3(g0.x + g1.x + g2.x, g0.y + g1.y + g2.y, g0.z + g1.z + g2.z);
    float3 ve1 = make_float3(e2.y * vw.z - e2.z * vw.y, e2.z * vw.x - e2.x * vw.z, e2.x * vw.y - e2.y * vw.x);
    float3 ve2 = make_float3(vw.y * e1.z - vw.z * e1.y, vw.z * e1.x - vw.x * e1.z, vw.x * e1.y - vw.y * e1.x);
    atomicAdd(v_verts + 3 * i1, ve1.x); atomicAdd(v_verts + 3 * i1 + 1, ve1.y); atomicAdd(v_verts + 3 * i1 + 2, ve1.z);
    atomicAdd(v_verts + 3 * i2, ve2.x); atomicAdd(v_verts + 3 * i2 + 1, ve2.y); atomicAdd(v_verts + 3 * i2 + 2, ve2.z);
    atomicAdd(v_verts + 3 * i0, -(ve1.x + ve2.x)); atomicAdd(v_verts + 3 * i0 + 1, -(ve1.y + ve2.y));
    atomicAdd(v_verts + 3 * i0 + 2, -(ve1.z + ve2.z));
}

// ---- tone map (geosplat.py:474-476; torch.nn.Softplus(beta=100, threshold=20)) ------------------------------
__device__ __forceinline__ float softplus100(float z, float &dz) {
    float bz = 100.0f * z;
    if (bz > 20.0f) { dz = 1.0f; return z; }
    float e = expf(bz);
    dz = e / (1.0f + e);
    return log1pf(e) / 100.0f;
}

__global__ void __launch_bounds__(256) tonemap_fwd_kernel(long long P, const float4 *__restrict__ rgba,
                                                           const float *__restrict__ exposure, float4 *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float e = __ldg(exposure);
    float4 p = rgba[i];
    float d;
    out[i] = make_float4(1.0f - softplus100(1.0f - p.x * e, d), 1.0f - softplus100(1.0f - p.y * e, d),
                         1.0f - softplus100(1.0f - p.z * e, d), p.w);
}

__global__ void __launch_bounds__(256) tonemap_bwd_kernel(long long P, const float4 *__restrict__ rgba,
                                                           const float *__restrict__ exposure,
                                                           const float4 *__restrict__ v_out, float4 *__restrict__ v_rgba,
                                                           float *__restrict__ v_exposure) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float ve = 0.f;
    if (i < P) {
        float e = __ldg(exposure);
        float4 p = rgba[i], v = v_out[i];
        float dx, dy, dz;
        softplus100(1.0f - p.x * e, dx); softplus100(1.0f - p.y * e, dy); softplus100(1.0f - p.z * e, dz);
        // y = 1 - sp(1 - x e): dy/dx = e sp', dy/de = x sp'
        v_rgba[i] = make_float4(v.x * dx * e, v.y * dy * e, v.z * dz * e, v.w);
        ve = v.x * dx * p.x + v.y * dy * p.y + v.z * dz * p.z;
    }
    block_add(ve, v_exposure);
}

// Planar variants: read the rasterizer's own outputs (render[P,3], alphas[P]) and write the RGBA image, so no
// concatenated copy of the image is ever made (the reference builds it with torch.cat, gsplat.py:358).
__global__ void __launch_bounds__(256) tonemap_planar_fwd_kernel(long long P, const float *__restrict__ render,
                                                                  const float *__restrict__ alphas,
                                                                  const float *__restrict__ exposure, int naive,
                                                                  float4 *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float e = __ldg(exposure);
    float r = render[3 * i] * e, g = render[3 * i + 1] * e, b = render[3 * i + 2] * e, d;
    if (naive) {
        r = 1.0f - softplus100(1.0f - r, d);
        g = 1.0f - softplus100(1.0f - g, d);
        b = 1.0f - softplus100(1.0f - b, d);
    }
    out[i] = make_float4(r, g, b, alphas[i]);
}

__global__ void __launch_bounds__(256) tonemap_planar_bwd_kernel(long long P, const float *__restrict__ render,
                                                                  const float *__restrict__ exposure, int naive,
                                                                  const float4 *__restrict__ v_out,
                                                                  float *__restrict__ v_render, float *__restrict__ v_alphas,
                                                                  float *__restrict__ v_exposure) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float ve = 0.f;
    if (i < P) {
        float e = __ldg(exposure);
        float4 v = v_out[i];
        float x = render[3 * i], y = render[3 * i + 1], z = render[3 * i + 2];
        float dx = 1.f, dy = 1.f, dz = 1.f;
        if (naive) { softplus100(1.0f - x * e, dx); softplus100(1.0f - y * e, dy); softplus100(1.0f - z * e, dz); }
        v_render[3 * i] = v.x * dx * e;
        v_render[3 * i + 1] = v.y * dy * e;
        v_render[3 * i + 2] = v.z * dz * e;
        v_alphas[i] = v.w;
        ve = v.x * dx * x + v.y * dy * y + v.z * dz * z;
    }
    block_add(ve, v_exposure);
}

}  // namespace

#define GSB_API extern "C" __attribute__((visibility("default")))

GSB_API int gsb_vertex_normals_fwd(int32_t V, int32_t F, const float *vertices, const int64_t *faces, float *raw,
                                   float *normals, void *stream) {
    GSB_CHECK_ARG(V >= 0 && F >= 0);
    if (V == 0) return GSB_OK;
    GSB_CHECK_ARG(vertices && raw && normals && (F == 0 || faces));
    cudaStream_t st = (cudaStream_t)stream;
    GSB_CHECK_CUDA(cudaMemsetAsync(raw, 0, sizeof(float) * 3 * (size_t)V, st));
    if (F > 0)
        vnormal_scatter_kernel<<<gsb_div_up(F, 256), 256, 0, st>>>(F, vertices, (const long long *)faces, raw);
    vnormal_normalize_kernel<<<gsb_div_up(V, 256), 256, 0, st>>>(V, raw, nullptr, normals);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_vertex_normals_bwd(int32_t V, int32_t F, const float *vertices, const int64_t *faces, const float *raw,
                                   const float *v_normals, float *scratch, float *v_vertices, void *stream) {
    GSB_CHECK_ARG(V >= 0 && F >= 0);
    if (V == 0 || F == 0) return GSB_OK;
    GSB_CHECK_ARG(vertices && faces && raw && v_normals && scratch && v_vertices);
    cudaStream_t st = (cudaStream_t)stream;
    vnormal_normalize_kernel<<<gsb_div_up(V, 256), 256, 0, st>>>(V, raw, v_normals, scratch);
    vnormal_bwd_faces_kernel<<<gsb_div_up(F, 256), 256, 0, st>>>(F, vertices, (const long long *)faces, scratch,
                                                                 v_vertices);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_mgadapter_fwd(int32_t F, const float *vertices, const float *vertex_normals, const int64_t *faces,
                              float *means, float *scales, float *quats, float *normals, float *opacities,
                              float *offsets, void *stream) {
    GSB_CHECK_ARG(F >= 0);
    if (F == 0) return GSB_OK;
    GSB_CHECK_ARG(vertices && faces && means && scales && quats && normals && opacities && offsets);
    mgadapter_fwd_kernel<<<gsb_div_up(F, 128), 128, 0, (cudaStream_t)stream>>>(
        F, vertices, vertex_normals, (const long long *)faces, make_consts(), means, scales, quats, normals, opacities,
        offsets);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_mgadapter_bwd(int32_t F, const float *vertices, const float *vertex_normals, const int64_t *faces,
                              const float *v_means, const float *v_scales, const float *v_quats,
                              const float *v_normals, float *v_vertices, float *v_vertex_normals, void *stream) {
    GSB_CHECK_ARG(F >= 0);
    if (F == 0) return GSB_OK;
    GSB_CHECK_ARG(vertices && faces && v_means && v_scales && v_quats && v_normals && v_vertices);
    mgadapter_bwd_kernel<<<gsb_div_up(F, 128), 128, 0, (cudaStream_t)stream>>>(
        F, vertices, vertex_normals, (const long long *)faces, make_consts(), v_means, v_scales, v_quats, v_normals,
        v_vertices, v_vertex_normals);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_tonemap_fwd(int64_t P, const float *rgba, const float *exposure, float *out, void *stream) {
    GSB_CHECK_ARG(P >= 0);
    if (P == 0) return GSB_OK;
    GSB_CHECK_ARG(rgba && exposure && out);
    tonemap_fwd_kernel<<<gsb_div_up(P, 256), 256, 0, (cudaStream_t)stream>>>(
        P, reinterpret_cast<const float4 *>(rgba), exposure, reinterpret_cast<float4 *>(out));
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_tonemap_bwd(int64_t P, const float *rgba, const float *exposure, const float *v_out, float *v_rgba,
                            float *v_exposure, void *stream) {
    GSB_CHECK_ARG(P >= 0);
    if (P == 0) return GSB_OK;
    GSB_CHECK_ARG(rgba && exposure && v_out && v_rgba && v_exposure);
    tonemap_bwd_kernel<<<gsb_div_up(P, 256), 256, 0, (cudaStream_t)stream>>>(
        P, reinterpret_cast<const float4 *>(rgba), exposure, reinterpret_cast<const float4 *>(v_out),
        reinterpret_cast<float4 *>(v_rgba), v_exposure);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_tonemap_planar_fwd(int64_t P, const float *render, const float *alphas, const float *exposure,
                                   int32_t naive, float *out, void *stream) {
    GSB_CHECK_ARG(P >= 0);
    if (P == 0) return GSB_OK;
    GSB_CHECK_ARG(render && alphas && exposure && out);
    tonemap_planar_fwd_kernel<<<gsb_div_up(P, 256), 256, 0, (cudaStream_t)stream>>>(
        P, render, alphas, exposure, naive, reinterpret_cast<float4 *>(out));
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_tonemap_planar_bwd(int64_t P, const float *render, const float *exposure, int32_t naive,
                                   const float *v_out, float *v_render, float *v_alphas, float *v_exposure,
                                   void *stream) {
    GSB_CHECK_ARG(P >= 0);
    if (P == 0) return GSB_OK;
    GSB_CHECK_ARG(render && exposure && v_out && v_render && v_alphas && v_exposure);
    tonemap_planar_bwd_kernel<<<gsb_div_up(P, 256), 256, 0, (cudaStream_t)stream>>>(
        P, render, exposure, naive, reinterpret_cast<const float4 *>(v_out), v_render, v_alphas, v_exposure);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}
