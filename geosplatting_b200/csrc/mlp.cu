// The small bias-free MLPs behind the kd / ks / z fields, fused: all layers, the ReLUs,
// the output activation and -- in the backward -- the recomputation, the chain to the input and the three weight
// gradients in ONE kernel each way, fp32 throughout.
//
// Replaces, for the shapes GeoSplatting's fields use (rfstudio/model/geosplat.py:485-518: 32 -> 32 [-> 32] -> {3, 2, 1},
// bias=False, ReLU between layers, sigmoid / none after the last): rfstudio/nn/mlp.py:125-145 (nn.Linear + F.relu per
// layer), i.e. per call 2-3 cuBLAS GEMMs forward, 4-6 backward (the [32 x N] . [N x 32] weight gradients run at a few
// percent of peak: K = 10^6, M = N = 32) and ~8 elementwise passes over [N, 32] activations.
//
// Shape of the kernels.  A WARP owns a tile of TP = 32 points; every [TP x 32] activation of the tile lives in shared
// memory TRANSPOSED ([feature][point], row stride TP + 4 floats) and every layer is a register-tiled TP x 32 x 32 product: a
// lane computes 4 points x 8 features (32 accumulators) and per contraction step loads 4 + 8 operands with three LDS.128
// -- 10.7 FMAs per 16 bytes loaded.  (Round 2's first version kept a point's activations in one thread's registers and
// broadcast the weights: one LDS.128 per FOUR FMAs, and the kernels sat on the shared-memory return path at 28 % of the
// FP32 peak.)  The lane -> (point group, feature group) map and the feature ownership (feature = group + 4 j) are chosen
// so that all operand loads and tile stores are bank-conflict-free; the weights are kept in shared memory in the two
// permuted orders those loads want.  The sums run over the contraction index in ascending order with one fmaf each,
// exactly as a per-point dot product would: values are the same to the bit as the per-thread formulation.
// One persistent CTA per SM, its shared memory filled with the tiles of its warps (12 backward / 20 forward warps).
// Measured per 10^6 points, 2 hidden layers (kd): forward 0.18 -> 0.15 ms, backward 0.63 -> 0.39 ms (31 TFLOP/s fp32);
// 8 points per lane (64-point tiles, 16 FMAs per 16 bytes) leaves room for only 6-7 warps per SM and is slower (0.46 ms):
// at 1.5 warps per scheduler the products are latency-bound however good their operand ratio
// (profiles/r02_prof_mlp_tiled.summary.csv: issue 51 %, L1 data pipe 51 %).
//   forward : x -> h1 = relu(W0 x) [-> h2 = relu(W1 h1)] -> y = act(Wout h)
//   backward: recompute h1, h2; dz = v_y act'; dWout += dz h^T; dh2 = relu'(h2) Wout^T dz; dW1 += dh2 h1^T (a 32 x 32 x TP
//             product, 4 x 8 accumulators per lane, persistent over the warp's tiles); dh1 = relu'(h1) W1^T dh2;
//             dW0 += dh1 x^T; dx = W0^T dh1.  One round of float atomics per warp at the end (~3 K values).
// Optional input rounding: the reference hands the MLP feats * s + feats.detach() * (1 - s) (encoding.py:239-240): the
// same values up to an fp32 rounding, which decides the side of a ReLU kink; `ref_round_scale` != 0 reproduces it.
#include "gsb_common.cuh"

namespace {

constexpr int D = 32;            // input and hidden width
#ifndef GSB_MLP_PPL
#define GSB_MLP_PPL 4
#endif
constexpr int PPL = GSB_MLP_PPL; // points per lane in the layer products (8: 16 FMAs per 16 bytes loaded; 4: 10.7, half the tile)
static_assert(PPL == 4 || PPL == 8, "GSB_MLP_PPL: 4 or 8");
constexpr int TP = 8 * PPL;      // points per warp tile
constexpr int LDP = TP + 4;      // row stride of a [D][TP] tile: 16-byte rows, consecutive rows 4 banks apart
constexpr int TILE = D * LDP;    // floats per tile
#ifndef GSB_MLP_FWD_WARPS
#define GSB_MLP_FWD_WARPS 20
#endif
#ifndef GSB_MLP_BWD_WARPS
#define GSB_MLP_BWD_WARPS 12
#endif
#ifndef GSB_MLP_UNROLL
#define GSB_MLP_UNROLL 4
#endif
constexpr int FWD_WARPS = GSB_MLP_FWD_WARPS, BWD_WARPS = GSB_MLP_BWD_WARPS, UNROLL = GSB_MLP_UNROLL;

struct Weights {
    const float *w0, *w1, *wout;   // [32,32], [32,32] or null (one hidden layer), [dout,32]
};

// Shared-memory weights of one CTA.  A lane owns the features {g + 4 j : j < 8} of a layer's output (g = its feature
// group), so the copies are permuted to make those 8 weights contiguous:
//   f[k][8 g + j] = W[g + 4 j][k]   for  out[p][o] = sum_k in[p][k] W[o][k]    (forward-type product, contraction over k)
//   b[o][8 g + j] = W[o][g + 4 j]   for  din[p][k] = sum_o dout[p][o] W[o][k]  (backward-type product, contraction over o)
struct SharedWeights {
    float w0f[D * D], w1f[D * D], w0b[D * D], w1b[D * D];
    float wout[4 * D];    // [o][k], rows >= dout are zero
    float woutT[D * 4];   // [k][o]
};

__device__ __forceinline__ void load_weights(const Weights &w, int n_hidden, int dout, bool backward, SharedWeights &s,
                                             int tid, int nthreads) {
    for (int i = tid; i < D * D; i += nthreads) {
        const int o = i / D, k = i % D;
        const float a = w.w0[i];
        s.w0f[k * D + (o & 3) * 8 + (o >> 2)] = a;
        if (backward) s.w0b[o * D + (k & 3) * 8 + (k >> 2)] = a;
        if (n_hidden > 1) {
            const float c = w.w1[i];
            s.w1f[k * D + (o & 3) * 8 + (o >> 2)] = c;
            if (backward) s.w1b[o * D + (k & 3) * 8 + (k >> 2)] = c;
        }
    }
    for (int i = tid; i < 4 * D; i += nthreads) {
        const int o = i / D, k = i % D;
        const float a = (o < dout) ? w.wout[i] : 0.f;
        s.wout[i] = a;
        s.woutT[k * 4 + o] = a;
    }
}

// lane -> point group (8 points p0 .. p0+7) and feature group g (features g + 4 j).  The low three lane bits hold two
// bits of the point group and one of the feature group: the eight lanes of a quarter-warp then write 8 x 16 bytes that
// cover all 32 banks (four point groups 32 bytes apart in two consecutive rows, which sit 4 banks apart).
struct LaneMap {
    int p0, g;
};
__device__ __forceinline__ LaneMap lane_map(int lane) {
    LaneMap m;
    if (PPL == 8) {
        m.p0 = 8 * ((lane & 3) + 4 * (lane >> 4));
        m.g = ((lane >> 2) & 1) + 2 * ((lane >> 3) & 1);
    } else {   // 4 points per lane: the eight lanes of a quarter-warp are the eight point groups of one row
        m.p0 = 4 * (lane & 7);
        m.g = lane >> 3;
    }
    return m;
}

__device__ __forceinline__ void zero(float (&acc)[PPL][8]) {
#pragma unroll
    for (int i = 0; i < PPL; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

// acc[i][j] += sum_c inT[c][p0 + i] * wp[c][8 g + j], c ascending (one fmaf per term): both product types, with
// wp = the f- or the b-copy of the layer's weights
__device__ __forceinline__ void tile_product(const float *__restrict__ inT, const float *__restrict__ wp, const LaneMap m,
                                             float (&acc)[PPL][8]) {
#pragma unroll UNROLL
    for (int c = 0; c < D; ++c) {
        float a[PPL];
#pragma unroll
        for (int i4 = 0; i4 < PPL / 4; ++i4) {
            const float4 v = *reinterpret_cast<const float4 *>(inT + c * LDP + m.p0 + 4 * i4);
            a[4 * i4] = v.x; a[4 * i4 + 1] = v.y; a[4 * i4 + 2] = v.z; a[4 * i4 + 3] = v.w;
        }
        const float4 b0 = *reinterpret_cast<const float4 *>(wp + c * D + 8 * m.g);
        const float4 b1 = *reinterpret_cast<const float4 *>(wp + c * D + 8 * m.g + 4);
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < PPL; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(b[j], a[i], acc[i][j]);
    }
}

// outT[g + 4 j][p0 + i] = relu(acc[i][j])
__device__ __forceinline__ void store_relu(float *__restrict__ outT, const LaneMap m, const float (&acc)[PPL][8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float *row = outT + (m.g + 4 * j) * LDP + m.p0;
#pragma unroll
        for (int i4 = 0; i4 < PPL / 4; ++i4)
            *reinterpret_cast<float4 *>(row + 4 * i4) =
                make_float4(fmaxf(acc[4 * i4][j], 0.f), fmaxf(acc[4 * i4 + 1][j], 0.f), fmaxf(acc[4 * i4 + 2][j], 0.f),
                            fmaxf(acc[4 * i4 + 3][j], 0.f));
    }
}

// hT[g + 4 j][p0 + i] = hT[..] > 0 ? acc[i][j] : 0   (the tile of activations becomes the tile of their gradients)
__device__ __forceinline__ void store_masked(float *__restrict__ hT, const LaneMap m, const float (&acc)[PPL][8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float *row = hT + (m.g + 4 * j) * LDP + m.p0;
#pragma unroll
        for (int i4 = 0; i4 < PPL / 4; ++i4) {
            const float4 h = *reinterpret_cast<const float4 *>(row + 4 * i4);
            *reinterpret_cast<float4 *>(row + 4 * i4) =
                make_float4(h.x > 0.f ? acc[4 * i4][j] : 0.f, h.y > 0.f ? acc[4 * i4 + 1][j] : 0.f,
                            h.z > 0.f ? acc[4 * i4 + 2][j] : 0.f, h.w > 0.f ? acc[4 * i4 + 3][j] : 0.f);
        }
    }
}

// xT[k][p] = x[n0 + p][k] (with the reference's rounding), zeros for points past the end
__device__ __forceinline__ void load_tile(const float *__restrict__ x, int64_t n0, int64_t N, float ref_round_scale,
                                          float *__restrict__ xT, int lane) {
    const float s = ref_round_scale, r = 1.0f - ref_round_scale;
#pragma unroll 4
    for (int it = 0; it < TP * D / 4 / 32; ++it) {
        const int idx = it * 32 + lane, p = idx >> 3, k4 = idx & 7;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + p < N) q = *reinterpret_cast<const float4 *>(x + (n0 + p) * D + 4 * k4);
        if (ref_round_scale != 0.f) {   // two roundings, no fma
            q.x = __fadd_rn(__fmul_rn(q.x, s), __fmul_rn(q.x, r));
            q.y = __fadd_rn(__fmul_rn(q.y, s), __fmul_rn(q.y, r));
            q.z = __fadd_rn(__fmul_rn(q.z, s), __fmul_rn(q.z, r));
            q.w = __fadd_rn(__fmul_rn(q.w, s), __fmul_rn(q.w, r));
        }
        float *col = xT + (4 * k4) * LDP + p;
        col[0] = q.x; col[LDP] = q.y; col[2 * LDP] = q.z; col[3 * LDP] = q.w;
    }
}

__device__ __forceinline__ float activate(float z, int act) { return act == 1 ? 1.0f / (1.0f + expf(-z)) : z; }

// z[o] = sum_k Wout[o][k] hT[k][p], k ascending: the output layer of point p (rows >= dout of Wout are zero)
__device__ __forceinline__ void output_layer(const float *__restrict__ hT, const float *__restrict__ woutT, int p,
                                             float (&z)[4]) {
    z[0] = z[1] = z[2] = z[3] = 0.f;
#pragma unroll 8
    for (int k = 0; k < D; ++k) {
        const float h = hT[k * LDP + p];
        const float4 w = *reinterpret_cast<const float4 *>(woutT + 4 * k);
        z[0] = fmaf(w.x, h, z[0]);
        z[1] = fmaf(w.y, h, z[1]);
        z[2] = fmaf(w.z, h, z[2]);
        z[3] = fmaf(w.w, h, z[3]);
    }
}

__global__ void __launch_bounds__(32 * FWD_WARPS) mlp_fwd_kernel(int64_t N, const float *__restrict__ x, Weights w,
                                                                  int n_hidden, int dout, int act, float ref_round_scale,
                                                                  float *__restrict__ y) {
    extern __shared__ float4 dyn_smem[];
    SharedWeights &sw = *reinterpret_cast<SharedWeights *>(dyn_smem);
    float *const tiles = reinterpret_cast<float *>(dyn_smem) + sizeof(SharedWeights) / sizeof(float);
    load_weights(w, n_hidden, dout, false, sw, threadIdx.x, 32 * FWD_WARPS);
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float *const t_a = tiles + (size_t)wib * 2 * TILE, *const t_b = t_a + TILE;
    const LaneMap m = lane_map(lane);
    const int64_t n_tiles = (N + TP - 1) / TP;
    for (int64_t tile = (int64_t)blockIdx.x * FWD_WARPS + wib; tile < n_tiles; tile += (int64_t)gridDim.x * FWD_WARPS) {
        const int64_t n0 = tile * TP;
        __syncwarp();
        load_tile(x, n0, N, ref_round_scale, t_a, lane);
        __syncwarp();
        float acc[PPL][8];
        zero(acc);
        tile_product(t_a, sw.w0f, m, acc);
        store_relu(t_b, m, acc);
        __syncwarp();
        const float *h = t_b;
        if (n_hidden > 1) {
            zero(acc);
            tile_product(t_b, sw.w1f, m, acc);
            store_relu(t_a, m, acc);       // x is dead
            __syncwarp();
            h = t_a;
        }
#pragma unroll
        for (int half = 0; half < TP / 32; ++half) {
            const int p = lane + 32 * half;
            float z[4];
            output_layer(h, sw.woutT, p, z);
            if (n0 + p < N) {
#pragma unroll
                for (int o = 0; o < 4; ++o)       // static indices: z[] stays in registers
                    if (o < dout) y[(n0 + p) * dout + o] = activate(z[o], act);
            }
        }
    }
}

// acc[a][b] += sum_p dT[i0 + a][p] * inT[jg + 4 b][p]: the weight-gradient product of one tile
__device__ __forceinline__ void weight_grad(const float *__restrict__ dT, const float *__restrict__ inT, int i0, int jg,
                                            float (&acc)[4][8]) {
#pragma unroll 2
    for (int q = 0; q < TP / 4; ++q) {
        float4 A[4], B[8];
#pragma unroll
        for (int a = 0; a < 4; ++a) A[a] = *reinterpret_cast<const float4 *>(dT + (i0 + a) * LDP + 4 * q);
#pragma unroll
        for (int b = 0; b < 8; ++b) B[b] = *reinterpret_cast<const float4 *>(inT + (jg + 4 * b) * LDP + 4 * q);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                float s = acc[a][b];
                s = fmaf(A[a].x, B[b].x, s);
                s = fmaf(A[a].y, B[b].y, s);
                s = fmaf(A[a].z, B[b].z, s);
                s = fmaf(A[a].w, B[b].w, s);
                acc[a][b] = s;
            }
    }
}

__global__ void __launch_bounds__(32 * BWD_WARPS) mlp_bwd_kernel(int64_t N, const float *__restrict__ x, Weights w,
                                                                  int n_hidden, int dout, int act,
                                                                  float ref_round_scale, const float *__restrict__ v_y,
                                                                  float *__restrict__ v_x, float *__restrict__ v_w0,
                                                                  float *__restrict__ v_w1, float *__restrict__ v_wout) {
    extern __shared__ float4 dyn_smem[];
    SharedWeights &sw = *reinterpret_cast<SharedWeights *>(dyn_smem);
    float *const tiles = reinterpret_cast<float *>(dyn_smem) + sizeof(SharedWeights) / sizeof(float);
    load_weights(w, n_hidden, dout, true, sw, threadIdx.x, 32 * BWD_WARPS);
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // per warp: x^T | h1^T (later dh1^T) | h-last^T (later dh-last^T) | dz[64][4]
    float *const t_x = tiles + (size_t)wib * (3 * TILE + TP * 4), *const t_h1 = t_x + TILE, *const t_c = t_h1 + TILE;
    float *const t_dz = t_c + TILE;
    const LaneMap m = lane_map(lane);
    // weight-gradient products: lane -> rows i0 .. i0+3 of dW (contiguous) x columns jg + 4 b (strided: the four column
    // groups of a quarter-warp read four consecutive tile rows, 4 banks apart)
    const int i0 = 4 * (lane >> 2), jg = lane & 3;
    float acc0[4][8], acc1[4][8], accout[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc0[a][b] = acc1[a][b] = 0.f;

    const int64_t n_tiles = (N + TP - 1) / TP;
    for (int64_t tile = (int64_t)blockIdx.x * BWD_WARPS + wib; tile < n_tiles; tile += (int64_t)gridDim.x * BWD_WARPS) {
        const int64_t n0 = tile * TP;
        __syncwarp();                          // the previous tile's readers are done with the shared tiles
        load_tile(x, n0, N, ref_round_scale, t_x, lane);
        __syncwarp();
        // ---- recompute the forward
        float acc[PPL][8];
        zero(acc);
        tile_product(t_x, sw.w0f, m, acc);
        if (n_hidden > 1) {
            store_relu(t_h1, m, acc);
            __syncwarp();
            zero(acc);
            tile_product(t_h1, sw.w1f, m, acc);
        }
        store_relu(t_c, m, acc);               // the activations of the last hidden layer
        __syncwarp();
        // ---- output layer per point: dz = v_y * act'(z)
#pragma unroll
        for (int half = 0; half < TP / 32; ++half) {
            const int p = lane + 32 * half;
            const bool live = n0 + p < N;
            float z[4], dz[4] = {0.f, 0.f, 0.f, 0.f};
            output_layer(t_c, sw.woutT, p, z);
#pragma unroll
            for (int o = 0; o < 4; ++o) {         // static indices: z[] / dz[] stay in registers
                if (o < dout) {
                    const float g = live ? v_y[(n0 + p) * dout + o] : 0.f;
                    if (act == 1) {
                        const float yv = 1.0f / (1.0f + expf(-z[o]));
                        dz[o] = g * yv * (1.0f - yv);
                    } else {
                        dz[o] = g;
                    }
                }
            }
            *reinterpret_cast<float4 *>(t_dz + 4 * p) = make_float4(dz[0], dz[1], dz[2], dz[3]);
        }
        __syncwarp();
        // ---- dWout[o][lane] += sum_p dz[p][o] * hlast[p][lane]
#pragma unroll 4
        for (int q = 0; q < TP / 4; ++q) {
            const float4 h4 = *reinterpret_cast<const float4 *>(t_c + lane * LDP + 4 * q);
            const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 z = *reinterpret_cast<const float4 *>(t_dz + 4 * (4 * q + c));
                accout[0] = fmaf(z.x, hv[c], accout[0]);
                accout[1] = fmaf(z.y, hv[c], accout[1]);
                accout[2] = fmaf(z.z, hv[c], accout[2]);
                accout[3] = fmaf(z.w, hv[c], accout[3]);
            }
        }
        __syncwarp();
        // ---- dh_last[p][k] = relu'(h[p][k]) * sum_o Wout[o][k] dz[p][o], in place
#pragma unroll
        for (int half = 0; half < TP / 32; ++half) {
            const int p = lane + 32 * half;
            const float4 z = *reinterpret_cast<const float4 *>(t_dz + 4 * p);
#pragma unroll 8
            for (int k = 0; k < D; ++k) {
                const float4 wk = *reinterpret_cast<const float4 *>(sw.woutT + 4 * k);
                float a = 0.f;
                a = fmaf(wk.x, z.x, a);
                a = fmaf(wk.y, z.y, a);
                a = fmaf(wk.z, z.z, a);
                a = fmaf(wk.w, z.w, a);                     // rows >= dout are zero
                float *cell = t_c + k * LDP + p;
                *cell = *cell > 0.f ? a : 0.f;
            }
        }
        __syncwarp();
        const float *d_first = t_c;            // gradient w.r.t. the first hidden layer's activations
        if (n_hidden > 1) {
            weight_grad(t_c, t_h1, i0, jg, acc1);                  // dW1[i][j] += sum_p dh2[p][i] * h1[p][j]
            zero(acc);
            tile_product(t_c, sw.w1b, m, acc);                     // W1^T dh2
            __syncwarp();                                          // every lane has read h1 for dW1
            store_masked(t_h1, m, acc);                            // h1 becomes dh1
            __syncwarp();
            d_first = t_h1;
        }
        weight_grad(d_first, t_x, i0, jg, acc0);                   // dW0[i][j] += sum_p dh1[p][i] * x[p][j]
        if (v_x != nullptr) {
            zero(acc);
            tile_product(d_first, sw.w0b, m, acc);                 // dx = W0^T dh1: point p0 + i, feature g + 4 j
#pragma unroll
            for (int i = 0; i < PPL; ++i) {
                if (n0 + m.p0 + i < N) {
                    float *row = v_x + (n0 + m.p0 + i) * D + m.g;
#pragma unroll
                    for (int j = 0; j < 8; ++j) row[4 * j] = acc[i][j];
                }
            }
        }
    }
    // ---- one round of atomics per warp
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) atomicAdd(v_w0 + (i0 + a) * D + jg + 4 * b, acc0[a][b]);
    if (n_hidden > 1) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) atomicAdd(v_w1 + (i0 + a) * D + jg + 4 * b, acc1[a][b]);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o)
        if (o < dout) atomicAdd(v_wout + o * D + lane, accout[o]);
}

int check(int64_t N, int32_t n_hidden, int32_t dout, int32_t act) {
    GSB_CHECK_ARG(N >= 0 && (n_hidden == 1 || n_hidden == 2) && dout >= 1 && dout <= 4 && (act == 0 || act == 1));
    return GSB_OK;
}

constexpr size_t FWD_SMEM = sizeof(SharedWeights) + sizeof(float) * FWD_WARPS * 2 * TILE;
constexpr size_t BWD_SMEM = sizeof(SharedWeights) + sizeof(float) * BWD_WARPS * (3 * TILE + TP * 4);
static_assert(FWD_SMEM <= 227 * 1024 && BWD_SMEM <= 227 * 1024, "mlp.cu: shared memory per CTA");

// persistent: one CTA per SM (the tiles of its warps fill the shared memory), fewer for small inputs
int grid_for(int64_t N, int warps) {
    static const int sms = gsb_sm_count();
    const int64_t want = ((N + TP - 1) / TP + warps - 1) / warps;
    return (int)(want < 1 ? 1 : (want < sms ? want : sms));
}

}  // namespace

#define GSB_API extern "C" __attribute__((visibility("default")))

GSB_API int gsb_mlp_fwd(int64_t N, const float *x, const float *w0, const float *w1, const float *wout,
                        int32_t n_hidden, int32_t dout, int32_t activation, float ref_round_scale, float *y,
                        void *stream) {
    int rc = check(N, n_hidden, dout, activation);
    if (rc != GSB_OK) return rc;
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(x && w0 && wout && y && (n_hidden == 1 || w1));
    Weights w{w0, w1, wout};
    GSB_CHECK_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
    mlp_fwd_kernel<<<grid_for(N, FWD_WARPS), 32 * FWD_WARPS, FWD_SMEM, (cudaStream_t)stream>>>(
        N, x, w, n_hidden, dout, activation, ref_round_scale, y);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_mlp_bwd(int64_t N, const float *x, const float *w0, const float *w1, const float *wout,
                        int32_t n_hidden, int32_t dout, int32_t activation, float ref_round_scale, const float *v_y,
                        float *v_x, float *v_w0, float *v_w1, float *v_wout, void *stream) {
    int rc = check(N, n_hidden, dout, activation);
    if (rc != GSB_OK) return rc;
    GSB_CHECK_ARG(v_w0 && v_wout && (n_hidden == 1 || v_w1));
    cudaStream_t st = (cudaStream_t)stream;
    GSB_CHECK_CUDA(cudaMemsetAsync(v_w0, 0, sizeof(float) * D * D, st));
    if (n_hidden > 1) GSB_CHECK_CUDA(cudaMemsetAsync(v_w1, 0, sizeof(float) * D * D, st));
    GSB_CHECK_CUDA(cudaMemsetAsync(v_wout, 0, sizeof(float) * dout * D, st));
    if (N == 0) return GSB_OK;
    GSB_CHECK_ARG(x && w0 && wout && v_y && (n_hidden == 1 || w1));
    Weights w{w0, w1, wout};
    GSB_CHECK_CUDA(cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
    mlp_bwd_kernel<<<grid_for(N, BWD_WARPS), 32 * BWD_WARPS, BWD_SMEM, st>>>(N, x, w, n_hidden, dout, activation,
                                                                            ref_round_scale, v_y, v_x, v_w0, v_w1, v_wout);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}
