// The small bias-free MLPs behind the kd / ks / z fields, fused: all layers, the ReLUs,
// the output activation and -- in the backward -- the recomputation, the chain to the input and the three weight
// gradients in ONE kernel each way, fp32 throughout.
//
// Replaces, for the shapes GeoSplatting's fields use (rfstudio/model/geosplat.py:485-518: 32 -> 32 [-> 32] -> {3, 2, 1},
// bias=False, ReLU between layers, sigmoid / none after the last): rfstudio/nn/mlp.py:125-145 (nn.Linear + F.relu per
// layer), i.e. per call 2-3 cuBLAS GEMMs forward, 4-6 backward (the [32 x N] . [N x 32] weight gradients run at a few
// percent of peak: K = 10^6, M = N = 32) and ~8 elementwise passes over [N, 32] activations.  Here the activations of
// a point never leave its thread's registers; the weights (<= 9 KB) are broadcast from shared memory.
//
// Forward : one thread per point, x[32] in registers, h = relu(W x) layer by layer.
// Backward: one thread per point recomputes h1, h2, then dz -> dh2 -> dh1 -> dx along the chain; the weight gradients
//           are sums over points of outer products: the 32 points of a warp are transposed through shared memory
//           ([32][36] tiles, conflict-free), lane L accumulates ROW L of every dW over the tile in registers across a
//           persistent loop, and adds it to global memory once at the end (float atomics, ~3 K per warp).
// Optional input rounding: the reference hands the MLP feats * s + feats.detach() * (1 - s) (encoding.py:239-240): the
// same values up to an fp32 rounding, which decides the side of a ReLU kink; `ref_round_scale` != 0 reproduces it.
#include "gsb_common.cuh"

namespace {

constexpr int D = 32;          // input and hidden width
constexpr int LD = 36;         // row stride of the transposition tiles (float4-aligned, conflict-free)
constexpr int FWD_THREADS = 128;
constexpr int BWD_WARPS = 2;

struct Weights {
    const float *w0, *w1, *wout;   // [32,32], [32,32] or null (one hidden layer), [dout,32]
};

__device__ __forceinline__ void load_weights(const Weights &w, int n_hidden, int dout, float *s_w0, float *s_w1,
                                             float *s_wout, int tid, int nthreads) {
    for (int i = tid; i < D * D; i += nthreads) {
        s_w0[i] = w.w0[i];
        if (n_hidden > 1) s_w1[i] = w.w1[i];
    }
    for (int i = tid; i < 4 * D; i += nthreads) s_wout[i] = (i < dout * D) ? w.wout[i] : 0.f;
}

// out[j] = sum_k W[j][k] v[k]   (W row-major in shared memory, broadcast float4 reads)
__device__ __forceinline__ void matvec(const float *__restrict__ s_w, const float (&v)[D], float (&out)[D]) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < D / 4; ++k4) {
            const float4 w = *reinterpret_cast<const float4 *>(s_w + j * D + 4 * k4);
            acc = fmaf(w.x, v[4 * k4], acc);
            acc = fmaf(w.y, v[4 * k4 + 1], acc);
            acc = fmaf(w.z, v[4 * k4 + 2], acc);
            acc = fmaf(w.w, v[4 * k4 + 3], acc);
        }
        out[j] = acc;
    }
}

// out[k] = sum_j W[j][k] v[j]   (the transpose product, same broadcast rows)
__device__ __forceinline__ void matvec_t(const float *__restrict__ s_w, const float (&v)[D], float (&out)[D]) {
#pragma unroll
    for (int k = 0; k < D; ++k) out[k] = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
#pragma unroll
        for (int k4 = 0; k4 < D / 4; ++k4) {
            const float4 w = *reinterpret_cast<const float4 *>(s_w + j * D + 4 * k4);
            out[4 * k4] = fmaf(w.x, v[j], out[4 * k4]);
            out[4 * k4 + 1] = fmaf(w.y, v[j], out[4 * k4 + 1]);
            out[4 * k4 + 2] = fmaf(w.z, v[j], out[4 * k4 + 2]);
            out[4 * k4 + 3] = fmaf(w.w, v[j], out[4 * k4 + 3]);
        }
    }
}

__device__ __forceinline__ void load_point(const float *__restrict__ x, int64_t n, float ref_round_scale, float (&v)[D]) {
    const float4 *row = reinterpret_cast<const float4 *>(x + n * D);
#pragma unroll
    for (int k4 = 0; k4 < D / 4; ++k4) {
        const float4 q = row[k4];
        v[4 * k4] = q.x; v[4 * k4 + 1] = q.y; v[4 * k4 + 2] = q.z; v[4 * k4 + 3] = q.w;
    }
    if (ref_round_scale != 0.f) {
        const float s = ref_round_scale, r = 1.0f - ref_round_scale;
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = __fadd_rn(__fmul_rn(v[k], s), __fmul_rn(v[k], r));   // two roundings, no fma
    }
}

__device__ __forceinline__ void relu(float (&v)[D]) {
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = fmaxf(v[k], 0.f);
}

__device__ __forceinline__ float activate(float z, int act) { return act == 1 ? 1.0f / (1.0f + expf(-z)) : z; }

__global__ void __launch_bounds__(FWD_THREADS, 3) mlp_fwd_kernel(int64_t N, const float *__restrict__ x, Weights w,
                                                               int n_hidden, int dout, int act, float ref_round_scale,
                                                               float *__restrict__ y) {
    __shared__ __align__(16) float s_w0[D * D], s_w1[D * D], s_wout[4 * D];
    load_weights(w, n_hidden, dout, s_w0, s_w1, s_wout, threadIdx.x, FWD_THREADS);
    __syncthreads();
    for (int64_t n = (int64_t)blockIdx.x * FWD_THREADS + threadIdx.x; n < N; n += (int64_t)gridDim.x * FWD_THREADS) {
        float v[D], h[D];
        load_point(x, n, ref_round_scale, v);
        matvec(s_w0, v, h);
        relu(h);
        if (n_hidden > 1) {
            matvec(s_w1, h, v);
            relu(v);
#pragma unroll
            for (int k = 0; k < D; ++k) h[k] = v[k];
        }
        for (int o = 0; o < dout; ++o) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) acc = fmaf(s_wout[o * D + k], h[k], acc);
            y[n * dout + o] = activate(acc, act);
        }
    }
}

// dW[L][k] += sum_p dOut[p][L] * in[p][k]   for this lane's row L, over the 32 points of the warp's tile
__device__ __forceinline__ void outer_rows(const float *__restrict__ t_dout, const float *__restrict__ t_in, int lane,
                                           float (&acc)[D]) {
#pragma unroll 4
    for (int p = 0; p < 32; ++p) {
        const float d = t_dout[p * LD + lane];
#pragma unroll
        for (int k4 = 0; k4 < D / 4; ++k4) {
            const float4 v = *reinterpret_cast<const float4 *>(t_in + p * LD + 4 * k4);
            acc[4 * k4] = fmaf(d, v.x, acc[4 * k4]);
            acc[4 * k4 + 1] = fmaf(d, v.y, acc[4 * k4 + 1]);
            acc[4 * k4 + 2] = fmaf(d, v.z, acc[4 * k4 + 2]);
            acc[4 * k4 + 3] = fmaf(d, v.w, acc[4 * k4 + 3]);
        }
    }
}

__device__ __forceinline__ void store_row(float *__restrict__ tile, int lane, const float (&v)[D]) {
#pragma unroll
    for (int k4 = 0; k4 < D / 4; ++k4)
        *reinterpret_cast<float4 *>(tile + lane * LD + 4 * k4) = make_float4(v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]);
}

__global__ void __launch_bounds__(32 * BWD_WARPS) mlp_bwd_kernel(int64_t N, const float *__restrict__ x, Weights w,
                                                                  int n_hidden, int dout, int act,
                                                                  float ref_round_scale, const float *__restrict__ v_y,
                                                                  float *__restrict__ v_x, float *__restrict__ v_w0,
                                                                  float *__restrict__ v_w1, float *__restrict__ v_wout) {
    __shared__ __align__(16) float s_w0[D * D], s_w1[D * D], s_wout[4 * D];
    __shared__ __align__(16) float s_tiles[BWD_WARPS][3][32 * LD];   // per warp: x | h1 | h-last, then dh
    __shared__ float s_dz[BWD_WARPS][32][4];
    load_weights(w, n_hidden, dout, s_w0, s_w1, s_wout, threadIdx.x, 32 * BWD_WARPS);
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float *const t_x = s_tiles[wib][0], *const t_h1 = s_tiles[wib][1], *const t_c = s_tiles[wib][2];
    float acc0[D], acc1[D], accout[4];
#pragma unroll
    for (int k = 0; k < D; ++k) acc0[k] = acc1[k] = 0.f;
    accout[0] = accout[1] = accout[2] = accout[3] = 0.f;

    const int64_t n_tiles = (N + 31) / 32;
    for (int64_t tile = (int64_t)blockIdx.x * BWD_WARPS + wib; tile < n_tiles; tile += (int64_t)gridDim.x * BWD_WARPS) {
        const int64_t n = tile * 32 + lane;
        const bool live = n < N;
        float h2[D];                          // the activations of the last hidden layer
        // ---- recompute the forward of my point (what is needed again later is parked in the shared tiles)
        {
            float v[D];
            if (live) load_point(x, n, ref_round_scale, v);
            else {
#pragma unroll
                for (int k = 0; k < D; ++k) v[k] = 0.f;
            }
            __syncwarp();                     // the previous tile's readers are done with the shared tiles
            store_row(t_x, lane, v);
            if (n_hidden > 1) {
                float h1[D];
                matvec(s_w0, v, h1);
                relu(h1);
                store_row(t_h1, lane, h1);
                matvec(s_w1, h1, h2);
            } else {
                matvec(s_w0, v, h2);
            }
            relu(h2);
        }
        store_row(t_c, lane, h2);
        float dz[4] = {0.f, 0.f, 0.f, 0.f};
        for (int o = 0; o < dout; ++o) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) acc = fmaf(s_wout[o * D + k], h2[k], acc);
            const float g = live ? v_y[n * dout + o] : 0.f;
            if (act == 1) {
                const float yv = 1.0f / (1.0f + expf(-acc));
                dz[o] = g * yv * (1.0f - yv);
            } else {
                dz[o] = g;
            }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) s_dz[wib][lane][o] = dz[o];
        __syncwarp();
        // ---- output layer: dWout[o][lane] += sum_p dz[p][o] * hlast[p][lane];  dh_last = relu'(h) * Wout^T dz
#pragma unroll 4
        for (int p = 0; p < 32; ++p) {
            const float hv = t_c[p * LD + lane];
            const float4 z = *reinterpret_cast<const float4 *>(&s_dz[wib][p][0]);
            accout[0] = fmaf(z.x, hv, accout[0]);
            accout[1] = fmaf(z.y, hv, accout[1]);
            accout[2] = fmaf(z.z, hv, accout[2]);
            accout[3] = fmaf(z.w, hv, accout[3]);
        }
        float dh[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float a = 0.f;
#pragma unroll
            for (int o = 0; o < 4; ++o) a = fmaf(s_wout[o * D + k], dz[o], a);     // rows >= dout are zero
            dh[k] = h2[k] > 0.f ? a : 0.f;
        }
        __syncwarp();
        store_row(t_c, lane, dh);             // h-last is dead: the tile now holds dh-last of every point
        __syncwarp();
        if (n_hidden > 1) {
            outer_rows(t_c, t_h1, lane, acc1);                    // dW1[lane][:] += sum_p dh2[p][lane] * h1[p][:]
            float dh1[D];
            matvec_t(s_w1, dh, dh1);
#pragma unroll
            for (int k = 0; k < D; ++k) dh[k] = t_h1[lane * LD + k] > 0.f ? dh1[k] : 0.f;     // my own row of h1
            __syncwarp();
            store_row(t_c, lane, dh);
            __syncwarp();
        }
        outer_rows(t_c, t_x, lane, acc0);                         // dW0[lane][:] += sum_p dh1[p][lane] * x[p][:]
        if (v_x != nullptr) {
            float dx[D];
            matvec_t(s_w0, dh, dx);
            if (live) {
                float4 *row = reinterpret_cast<float4 *>(v_x + n * D);
#pragma unroll
                for (int k4 = 0; k4 < D / 4; ++k4) row[k4] = make_float4(dx[4 * k4], dx[4 * k4 + 1], dx[4 * k4 + 2], dx[4 * k4 + 3]);
            }
        }
    }
    // ---- one round of atomics per warp
#pragma unroll
    for (int k = 0; k < D; ++k) atomicAdd(v_w0 + lane * D + k, acc0[k]);
    if (n_hidden > 1) {
#pragma unroll
        for (int k = 0; k < D; ++k) atomicAdd(v_w1 + lane * D + k, acc1[k]);
    }
    for (int o = 0; o < dout; ++o) atomicAdd(v_wout + o * D + lane, accout[o]);
}

int check(int64_t N, int32_t n_hidden, int32_t dout, int32_t act) {
    GSB_CHECK_ARG(N >= 0 && (n_hidden == 1 || n_hidden == 2) && dout >= 1 && dout <= 4 && (act == 0 || act == 1));
    return GSB_OK;
}

int grid_for(int64_t units, int per_cta) {
    const int64_t want = (units + per_cta - 1) / per_cta;
    const int64_t cap = 148 * 8;                 // persistent: a few CTAs per SM
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
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
    mlp_fwd_kernel<<<grid_for(N, FWD_THREADS), FWD_THREADS, 0, (cudaStream_t)stream>>>(N, x, w, n_hidden, dout, activation,
                                                                                        ref_round_scale, y);
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
    mlp_bwd_kernel<<<grid_for((N + 31) / 32, BWD_WARPS), 32 * BWD_WARPS, 0, st>>>(N, x, w, n_hidden, dout, activation,
                                                                                   ref_round_scale, v_y, v_x, v_w0, v_w1,
                                                                                   v_wout);
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}
