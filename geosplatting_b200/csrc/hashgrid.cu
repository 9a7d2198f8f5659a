// Multi-resolution hash-grid encoding, forward and backward: the fields behind kd / ks / z of the hot path
// (SURVEY.md section 8f rank 1).  Semantics are those of the reference's own `torch` backend of HashEncoding
// (rfstudio/model/components/encoding.py:124-138 scalings/offsets, :164-180 hash_fn, :182-229 pytorch_fwd): every level
// is hashed (primes 1, 2654435761, 805459861, modulo 2^log2), corners are ceil/floor of x01 * scaling, trilinear blend
// in the reference's operation order.  THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false so that the features are
// bit-identical to that code (oracle/encoding.py); tinycudann -- the reference's default backend, unpinned and absent
// here -- uses a different level layout and fp16 storage.
//
// One thread per (point, level): the 16 threads of a point share its coordinates through L1, the feature pair is one
// float2, the 8 corner gathers hit the L2-resident table (2^18 entries x 16 levels x 8 B = 33.5 MB).  Algorithmic
// bytes: 12 B in + 8 B/level out per point (+ 64 B/level of L2 gathers).  The backward scatters with one
// red.global.add.v2.f32 per corner and writes d/dx once per point after a 16-lane segmented reduction.
#include <stdlib.h>

#include "gsb_common.cuh"

namespace {

constexpr int MAX_LEVELS = 32;

struct Scalings {
    float s[MAX_LEVELS];
};

__device__ __forceinline__ uint32_t hash3(int x, int y, int z, uint32_t mask) {
    // low bits of the reference's int64 products / xor / modulo 2^k: 32-bit wrap-around arithmetic keeps them
    return ((uint32_t)x ^ ((uint32_t)y * 2654435761u) ^ ((uint32_t)z * 805459861u)) & mask;
}

struct Cell {
    uint32_t idx[8];   // f0..f7 in the reference's naming (c = ceil corner, f = floor corner)
    float ox, oy, oz;
    bool paired;       // floor x even and ceil x = floor x + 1: every (ceil-x, floor-x) corner pair is one aligned 16-byte
                       // pair of table entries (the hashes differ only in bit 0)
};

// The four x-neighbour pairs (ceil-x corner, floor-x corner) in the reference's numbering.
__device__ constexpr int PAIR_C[4] = {0, 1, 4, 5};
__device__ constexpr int PAIR_F[4] = {3, 2, 7, 6};

// Gathers of divergent 8-byte entries are bound by the L1TEX sector-lookup rate (ncu: l1tex 90 %): fetch an aligned pair
// with ONE 16-byte load where the hash allows it (half of all cells).
__device__ __forceinline__ void gather8(const float2 *__restrict__ table, const Cell &c, float2 f[8]) {
    if (c.paired) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t ic = c.idx[PAIR_C[k]];
            const float4 v = __ldg(reinterpret_cast<const float4 *>(table) + (ic >> 1));
            const float2 e0 = make_float2(v.x, v.y), e1 = make_float2(v.z, v.w);
            f[PAIR_C[k]] = (ic & 1u) ? e1 : e0;
            f[PAIR_F[k]] = (ic & 1u) ? e0 : e1;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = __ldg(table + c.idx[k]);
    }
}

__device__ __forceinline__ Cell locate(const float *__restrict__ x, int n, float scaling, uint32_t level_base,
                                       uint32_t mask) {
    Cell c;
    const float sx = (x[3 * n] * 0.5f + 0.5f) * scaling;
    const float sy = (x[3 * n + 1] * 0.5f + 0.5f) * scaling;
    const float sz = (x[3 * n + 2] * 0.5f + 0.5f) * scaling;
    const float fx = floorf(sx), fy = floorf(sy), fz = floorf(sz);
    const int ifx = (int)fx, ify = (int)fy, ifz = (int)fz;
    const int icx = (int)ceilf(sx), icy = (int)ceilf(sy), icz = (int)ceilf(sz);
    c.ox = sx - fx; c.oy = sy - fy; c.oz = sz - fz;
    c.paired = ((ifx & 1) == 0) && (icx == ifx + 1);
    c.idx[0] = level_base + hash3(icx, icy, icz, mask);
    c.idx[1] = level_base + hash3(icx, ify, icz, mask);
    c.idx[2] = level_base + hash3(ifx, ify, icz, mask);
    c.idx[3] = level_base + hash3(ifx, icy, icz, mask);
    c.idx[4] = level_base + hash3(icx, icy, ifz, mask);
    c.idx[5] = level_base + hash3(icx, ify, ifz, mask);
    c.idx[6] = level_base + hash3(ifx, ify, ifz, mask);
    c.idx[7] = level_base + hash3(ifx, icy, ifz, mask);
    return c;
}

__global__ void __launch_bounds__(256) hashgrid_fwd_kernel(long long total, int L, int log2_T, Scalings sc,
                                                            const float *__restrict__ x,
                                                            const float2 *__restrict__ table,
                                                            float2 *__restrict__ feats) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int n = (int)(t / L), l = (int)(t % L);
    const Cell c = locate(x, n, sc.s[l], (uint32_t)l << log2_T, (1u << log2_T) - 1u);
    float2 f[8];
    gather8(table, c, f);
    const float ox = c.ox, oy = c.oy, oz = c.oz, rx = 1.0f - ox, ry = 1.0f - oy, rz = 1.0f - oz;
    float2 out;
    {
        const float f03 = f[0].x * ox + f[3].x * rx, f12 = f[1].x * ox + f[2].x * rx;
        const float f56 = f[5].x * ox + f[6].x * rx, f47 = f[4].x * ox + f[7].x * rx;
        out.x = (f03 * oy + f12 * ry) * oz + (f47 * oy + f56 * ry) * rz;
    }
    {
        const float f03 = f[0].y * ox + f[3].y * rx, f12 = f[1].y * ox + f[2].y * rx;
        const float f56 = f[5].y * ox + f[6].y * rx, f47 = f[4].y * ox + f[7].y * rx;
        out.y = (f03 * oy + f12 * ry) * oz + (f47 * oy + f56 * ry) * rz;
    }
    feats[t] = out;
}

// v_table += (corner weight) * v_feats * table_grad_scale;  v_x[n] = sum over levels of d feats / d x . v_feats.
// L must divide 32 or be a multiple of it for the segmented reduction; otherwise v_x goes through atomics.
__global__ void __launch_bounds__(256) hashgrid_bwd_kernel(long long total, int L, int log2_T, Scalings sc,
                                                            const float *__restrict__ x,
                                                            const float2 *__restrict__ table,
                                                            const float2 *__restrict__ v_feats, float table_grad_scale,
                                                            float2 *__restrict__ v_table, float *__restrict__ v_x,
                                                            int segmented) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < total;
    const int n = live ? (int)(t / L) : 0, l = live ? (int)(t % L) : 0;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (live) {
        const float scaling = sc.s[l];
        const Cell c = locate(x, n, scaling, (uint32_t)l << log2_T, (1u << log2_T) - 1u);
        const float2 v = v_feats[t];
        const float ox = c.ox, oy = c.oy, oz = c.oz, rx = 1.0f - ox, ry = 1.0f - oy, rz = 1.0f - oz;
        // value = ((f0 ox + f3 rx) oy + (f1 ox + f2 rx) ry) oz + ((f4 ox + f7 rx) oy + (f5 ox + f6 rx) ry) rz
        const float w[8] = {ox * oy * oz, ox * ry * oz, rx * ry * oz, rx * oy * oz,
                            ox * oy * rz, ox * ry * rz, rx * ry * rz, rx * oy * rz};
        if (v_table) {
            if (c.paired) {   // one red.global.add.v4.f32 per aligned pair of entries
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t ic = c.idx[PAIR_C[k]];
                    const float sc_ = w[PAIR_C[k]] * table_grad_scale, sf_ = w[PAIR_F[k]] * table_grad_scale;
                    const float s0 = (ic & 1u) ? sf_ : sc_, s1 = (ic & 1u) ? sc_ : sf_;
                    atomicAdd(reinterpret_cast<float4 *>(v_table) + (ic >> 1),
                              make_float4(s0 * v.x, s0 * v.y, s1 * v.x, s1 * v.y));
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float s = w[k] * table_grad_scale;
                    atomicAdd(v_table + c.idx[k], make_float2(s * v.x, s * v.y));
                }
            }
        }
        if (v_x) {
            float2 f[8];
            gather8(table, c, f);
            float d[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float a[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] = ch ? f[k].y : f[k].x;
                const float vv = ch ? v.y : v.x;
                const float f03 = a[0] * ox + a[3] * rx, f12 = a[1] * ox + a[2] * rx;
                const float f56 = a[5] * ox + a[6] * rx, f47 = a[4] * ox + a[7] * rx;
                const float d03 = a[0] - a[3], d12 = a[1] - a[2], d56 = a[5] - a[6], d47 = a[4] - a[7];
                d[0] += vv * ((d03 * oy + d12 * ry) * oz + (d47 * oy + d56 * ry) * rz);
                d[1] += vv * ((f03 - f12) * oz + (f47 - f56) * rz);
                d[2] += vv * ((f03 * oy + f12 * ry) - (f47 * oy + f56 * ry));
            }
            const float j = 0.5f * scaling;   // d offset / d x  (floor and ceil carry no gradient)
            gx = d[0] * j; gy = d[1] * j; gz = d[2] * j;
        }
    }
    if (!v_x) return;
    if (segmented) {   // L consecutive lanes hold one point: xor-butterfly inside the segment
        for (int o = 1; o < L && o < 32; o <<= 1) {
            gx += __shfl_xor_sync(0xffffffffu, gx, o);
            gy += __shfl_xor_sync(0xffffffffu, gy, o);
            gz += __shfl_xor_sync(0xffffffffu, gz, o);
        }
        if (live && (l % 32) == 0) {
            if (L <= 32) { v_x[3 * n] = gx; v_x[3 * n + 1] = gy; v_x[3 * n + 2] = gz; }
            else { atomicAdd(v_x + 3 * n, gx); atomicAdd(v_x + 3 * n + 1, gy); atomicAdd(v_x + 3 * n + 2, gz); }
        }
    } else if (live) {
        atomicAdd(v_x + 3 * n, gx); atomicAdd(v_x + 3 * n + 1, gy); atomicAdd(v_x + 3 * n + 2, gz);
    }
}

// ---- level-major variants (L <= 32): a CTA = 32 consecutive points x L levels, WARP w = level w of those 32 points ----
// The points arrive in mesh order (MGAdaptor emits Gaussians face by face), so the 32 points of a warp are neighbours in
// space: at one level their cells coincide or touch, and the corner gathers of a warp fall into a few sectors instead of
// 8 x 32 scattered ones (with a (point, level)-per-thread layout the lanes of a warp address 16 different level tables).
// The [point][level] feature rows are transposed through shared memory so that global loads / stores stay coalesced.
constexpr int TILE_P = 32;

__global__ void __launch_bounds__(1024) hashgrid_fwd_lm_kernel(int64_t N, int L, int log2_T, Scalings sc,
                                                                const float *__restrict__ x,
                                                                const float2 *__restrict__ table,
                                                                float2 *__restrict__ feats) {
    __shared__ float2 s_out[TILE_P * (MAX_LEVELS + 1)];
    const int l = threadIdx.x >> 5, p = threadIdx.x & 31;
    const int64_t n0 = (int64_t)blockIdx.x * TILE_P, n = n0 + p;
    if (n < N) {
        const Cell c = locate(x, (int)n, sc.s[l], (uint32_t)l << log2_T, (1u << log2_T) - 1u);
        float2 f[8];
        gather8(table, c, f);
        const float ox = c.ox, oy = c.oy, oz = c.oz, rx = 1.0f - ox, ry = 1.0f - oy, rz = 1.0f - oz;
        float2 out;
        {
            const float f03 = f[0].x * ox + f[3].x * rx, f12 = f[1].x * ox + f[2].x * rx;
            const float f56 = f[5].x * ox + f[6].x * rx, f47 = f[4].x * ox + f[7].x * rx;
            out.x = (f03 * oy + f12 * ry) * oz + (f47 * oy + f56 * ry) * rz;
        }
        {
            const float f03 = f[0].y * ox + f[3].y * rx, f12 = f[1].y * ox + f[2].y * rx;
            const float f56 = f[5].y * ox + f[6].y * rx, f47 = f[4].y * ox + f[7].y * rx;
            out.y = (f03 * oy + f12 * ry) * oz + (f47 * oy + f56 * ry) * rz;
        }
        s_out[p * (L + 1) + l] = out;
    }
    __syncthreads();
    const int e = threadIdx.x, pp = e / L, ll = e % L;         // consecutive threads -> consecutive floats of the tile
    if (n0 + pp < N) feats[(n0 + pp) * L + ll] = s_out[pp * (L + 1) + ll];
}

// Table gradients of one warp = one level of 32 consecutive points.  The points arrive in mesh order, so at a coarse level
// (cells wider than the warp's patch of surface) all 32 lanes scatter into the same one to three cells: 32 same-address
// reductions per corner, which serialise in L2 -- the kernel is bound by that traffic (141 M red.v2 per 10^6 points).
// When the lanes' floor corners fall into at most AGG_MAX_GROUPS distinct entries the warp sums each group with a
// butterfly and its first lane issues ONE reduction per distinct entry and corner; finer levels keep one reduction per
// lane (paired into 16-byte reductions where the hash allows).  Dead lanes of a ragged last tile ride along with weight 0.
// Measured per 10^6 mesh-ordered points, encode forward + backward: off 1.10 ms, <= 3 groups 0.97, <= 6 groups 1.20,
// <= 12 groups 1.56 (a round is ~12 dependent shuffles); summing all 8 corners of a CELL in one round (16 butterflies in
// flight) needs more than the 64 registers a 1024-thread CTA leaves and was slower everywhere.
#ifndef GSB_HASHGRID_AGG_MAX_GROUPS
#define GSB_HASHGRID_AGG_MAX_GROUPS 3
#endif
constexpr int AGG_MAX_GROUPS = GSB_HASHGRID_AGG_MAX_GROUPS;

__device__ __forceinline__ void warp_grouped_add(float2 *__restrict__ v_table, uint32_t idx, float2 val, int lane) {
    unsigned remaining = 0xffffffffu;
    while (remaining) {                                   // warp-uniform: one round per distinct entry
        const int leader = __ffs((int)remaining) - 1;
        const uint32_t key = __shfl_sync(0xffffffffu, idx, leader);
        const bool in = (idx == key);
        float sx = in ? val.x : 0.f, sy = in ? val.y : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        if (lane == leader) atomicAdd(v_table + key, make_float2(sx, sy));
        remaining &= ~__ballot_sync(0xffffffffu, in);
    }
}

__global__ void __launch_bounds__(1024) hashgrid_bwd_lm_kernel(int64_t N, int L, int log2_T, Scalings sc,
                                                                const float *__restrict__ x,
                                                                const float2 *__restrict__ table,
                                                                const float2 *__restrict__ v_feats,
                                                                float table_grad_scale, float2 *__restrict__ v_table,
                                                                float *__restrict__ v_x) {
    __shared__ float2 s_v[TILE_P * (MAX_LEVELS + 1)];
    __shared__ float s_g[3][MAX_LEVELS][TILE_P];
    const int l = threadIdx.x >> 5, p = threadIdx.x & 31;
    const int64_t n0 = (int64_t)blockIdx.x * TILE_P, n = n0 + p;
    {
        const int e = threadIdx.x, pp = e / L, ll = e % L;
        if (n0 + pp < N) s_v[pp * (L + 1) + ll] = v_feats[(n0 + pp) * L + ll];
    }
    __syncthreads();
    float gx = 0.f, gy = 0.f, gz = 0.f;
    const bool live = n < N;
    {
        const float scaling = sc.s[l];
        // a dead lane (ragged last tile) takes the last point's cell with a zero cotangent: it joins the warp collectives
        const Cell c = locate(x, (int)(live ? n : N - 1), scaling, (uint32_t)l << log2_T, (1u << log2_T) - 1u);
        const float2 v = live ? s_v[p * (L + 1) + l] : make_float2(0.f, 0.f);
        const float ox = c.ox, oy = c.oy, oz = c.oz, rx = 1.0f - ox, ry = 1.0f - oy, rz = 1.0f - oz;
        const float w[8] = {ox * oy * oz, ox * ry * oz, rx * ry * oz, rx * oy * oz,
                            ox * oy * rz, ox * ry * rz, rx * ry * rz, rx * oy * rz};
        if (v_table) {
            // distinct floor-corner entries among the 32 lanes
            int groups = 0;
            {
                unsigned remaining = 0xffffffffu;
                while (remaining && groups <= AGG_MAX_GROUPS) {
                    const uint32_t key = __shfl_sync(0xffffffffu, c.idx[6], __ffs((int)remaining) - 1);
                    remaining &= ~__ballot_sync(0xffffffffu, c.idx[6] == key);
                    ++groups;
                }
            }
            if (groups <= AGG_MAX_GROUPS) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float s_ = w[k] * table_grad_scale;
                    warp_grouped_add(v_table, c.idx[k], make_float2(s_ * v.x, s_ * v.y), p);
                }
            } else if (!live) {
                // nothing to scatter
            } else if (c.paired) {   // one red.global.add.v4.f32 per aligned pair of entries
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t ic = c.idx[PAIR_C[k]];
                    const float sc_ = w[PAIR_C[k]] * table_grad_scale, sf_ = w[PAIR_F[k]] * table_grad_scale;
                    const float s0 = (ic & 1u) ? sf_ : sc_, s1 = (ic & 1u) ? sc_ : sf_;
                    atomicAdd(reinterpret_cast<float4 *>(v_table) + (ic >> 1),
                              make_float4(s0 * v.x, s0 * v.y, s1 * v.x, s1 * v.y));
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float s_ = w[k] * table_grad_scale;
                    atomicAdd(v_table + c.idx[k], make_float2(s_ * v.x, s_ * v.y));
                }
            }
        }
        if (v_x && live) {
            float2 f[8];
            gather8(table, c, f);
            float d[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float a[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] = ch ? f[k].y : f[k].x;
                const float vv = ch ? v.y : v.x;
                const float f03 = a[0] * ox + a[3] * rx, f12 = a[1] * ox + a[2] * rx;
                const float f56 = a[5] * ox + a[6] * rx, f47 = a[4] * ox + a[7] * rx;
                const float d03 = a[0] - a[3], d12 = a[1] - a[2], d56 = a[5] - a[6], d47 = a[4] - a[7];
                d[0] += vv * ((d03 * oy + d12 * ry) * oz + (d47 * oy + d56 * ry) * rz);
                d[1] += vv * ((f03 - f12) * oz + (f47 - f56) * rz);
                d[2] += vv * ((f03 * oy + f12 * ry) - (f47 * oy + f56 * ry));
            }
            const float j = 0.5f * scaling;   // d offset / d x  (floor and ceil carry no gradient)
            gx = d[0] * j; gy = d[1] * j; gz = d[2] * j;
        }
    }
    if (!v_x) return;
    s_g[0][l][p] = gx; s_g[1][l][p] = gy; s_g[2][l][p] = gz;
    __syncthreads();
    if (l == 0 && n < N) {                     // level 0's warp sums the levels of its 32 points, in level order
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (int k = 0; k < L; ++k) { sx += s_g[0][k][p]; sy += s_g[1][k][p]; sz += s_g[2][k][p]; }
        v_x[3 * n] = sx; v_x[3 * n + 1] = sy; v_x[3 * n + 2] = sz;
    }
}

int check(int64_t N, int32_t L, int32_t F, int32_t log2_T, const float *scalings_host) {
    GSB_CHECK_ARG(N >= 0 && L >= 1 && L <= MAX_LEVELS && log2_T >= 1 && log2_T <= 24 && scalings_host != nullptr);
    if (F != 2) {
        gsb_set_error("gsb_hashgrid: features_per_level must be 2 (GeoSplatting's fields, geosplat.py:485-518), got %d", F);
        return GSB_EINVAL;
    }
    GSB_CHECK_ARG((int64_t)L << log2_T < (int64_t)1 << 31);
    return GSB_OK;
}

}  // namespace

#define GSB_API extern "C" __attribute__((visibility("default")))

GSB_API int gsb_hashgrid_fwd(int64_t N, const float *x, const float *table, int32_t L, int32_t F, int32_t log2_T,
                             const float *scalings_host, float *feats, void *stream) {
    int rc = check(N, L, F, log2_T, scalings_host);
    if (rc != GSB_OK) return rc;
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(x && table && feats);
    Scalings sc;
    for (int l = 0; l < L; ++l) sc.s[l] = scalings_host[l];
    const long long total = (long long)N * L;
#ifndef GSB_HOST_EMULATION
    if (L <= MAX_LEVELS && !getenv("GSB_HASHGRID_POINT_MAJOR")) {       // level-major warps (see hashgrid_fwd_lm_kernel)
        hashgrid_fwd_lm_kernel<<<gsb_div_up(N, TILE_P), TILE_P * L, 0, (cudaStream_t)stream>>>(
            N, L, log2_T, sc, x, reinterpret_cast<const float2 *>(table), reinterpret_cast<float2 *>(feats));
        GSB_CHECK_LAUNCH();
        return GSB_OK;
    }
#endif
    hashgrid_fwd_kernel<<<gsb_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
        total, L, log2_T, sc, x, reinterpret_cast<const float2 *>(table), reinterpret_cast<float2 *>(feats));
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_hashgrid_bwd(int64_t N, const float *x, const float *table, int32_t L, int32_t F, int32_t log2_T,
                             const float *scalings_host, const float *v_feats, float table_grad_scale, float *v_table,
                             float *v_x, void *stream) {
    int rc = check(N, L, F, log2_T, scalings_host);
    if (rc != GSB_OK) return rc;
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(x && table && v_feats && (v_table || v_x));
    Scalings sc;
    for (int l = 0; l < L; ++l) sc.s[l] = scalings_host[l];
    const long long total = (long long)N * L;
#ifndef GSB_HOST_EMULATION
    if (L <= MAX_LEVELS && !getenv("GSB_HASHGRID_POINT_MAJOR")) {
        hashgrid_bwd_lm_kernel<<<gsb_div_up(N, TILE_P), TILE_P * L, 0, (cudaStream_t)stream>>>(
            N, L, log2_T, sc, x, reinterpret_cast<const float2 *>(table), reinterpret_cast<const float2 *>(v_feats),
            table_grad_scale, reinterpret_cast<float2 *>(v_table), v_x);
        GSB_CHECK_LAUNCH();
        return GSB_OK;
    }
#endif
#ifdef GSB_HOST_EMULATION   // tests/emu runs the threads one after another: no warp to reduce over, d/dx through atomics
    const int segmented = 0;
#else
    const int segmented = (L <= 32) ? ((32 % L) == 0) : ((L % 32) == 0);
#endif
    if (v_x && (!segmented || L > 32))
        GSB_CHECK_CUDA(cudaMemsetAsync(v_x, 0, sizeof(float) * 3 * (size_t)N, (cudaStream_t)stream));
    hashgrid_bwd_kernel<<<gsb_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
        total, L, log2_T, sc, x, reinterpret_cast<const float2 *>(table), reinterpret_cast<const float2 *>(v_feats),
        table_grad_scale, reinterpret_cast<float2 *>(v_table), v_x, segmented);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}
