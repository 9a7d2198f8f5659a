// Staging experiment for the composite kernels (north_star: "TMA/shared-memory tile staging").
//
// What the composite kernels stage per step: 32 records of 48 bytes, GATHERED through a sorted index list, into a
// warp-private shared-memory slab.  Two ways to do that, same access pattern, same consumer:
//   A  lanes : every lane loads one index and its record (3 x LDG.128) into registers one chunk ahead and stores it to the
//              slab (what csrc/composite.cu does);
//   B  TMA   : eight lanes each issue one cp.async.bulk.tensor.2d ... tile::gather4 (4 rows of a [N, 12] fp32 tensor per
//              instruction), completion on one mbarrier per slab, double-buffered.
// Prints one JSON line: ms per pass, staged GB/s, checksums (must agree), and whether the TMA path ran at all.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tma_gather_experiment tma_gather_experiment.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

struct Rec { float4 k, q, c; };
constexpr int WARPS = 4;
constexpr int PER_WARP = 1024;   // list entries walked by one warp (a long sub-list)

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ float consume(const Rec *slab, int lane, float acc) {
    // stand-in for the alpha evaluation: every lane reads every row (broadcast), light arithmetic
#pragma unroll 8
    for (int t = 0; t < 32; ++t) {
        const float4 k = slab[t].k, q = slab[t].q;
        acc = fmaf(k.x - (float)lane, q.x, fmaf(k.y, q.y, acc));
    }
    return acc;
}

__device__ __forceinline__ float consume_groups(const unsigned char (*groups)[256], int lane, float acc) {
#pragma unroll 8
    for (int t = 0; t < 32; ++t) {
        const Rec *r = reinterpret_cast<const Rec *>(&groups[t >> 2][0]) + (t & 3);
        const float4 k = r->k, q = r->q;
        acc = fmaf(k.x - (float)lane, q.x, fmaf(k.y, q.y, acc));
    }
    return acc;
}

__global__ void __launch_bounds__(32 * WARPS) stage_lanes(const Rec *__restrict__ rec, const int *__restrict__ list,
                                                           int n_warps, float *__restrict__ out) {
    __shared__ Rec slab[WARPS][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, w = blockIdx.x * WARPS + wib;
    if (w >= n_warps) return;
    const int *my = list + (size_t)w * PER_WARP;
    float acc = 0.f;
    int e = my[lane], e_next = my[32 + lane];
    Rec r = rec[e];
    for (int base = 0; base < PER_WARP; base += 32) {
        __syncwarp();
        slab[wib][lane] = r;
        __syncwarp();
        e = e_next;
        if (base + 32 < PER_WARP) r = rec[e];
        if (base + 64 < PER_WARP) e_next = my[base + 64 + lane];
        acc = consume(slab[wib], lane, acc);
    }
    out[(size_t)w * 32 + lane] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32 * WARPS) stage_tma(const __grid_constant__ CUtensorMap tmap,
                                                         const int *__restrict__ list, int n_warps,
                                                         float *__restrict__ out, int *__restrict__ err) {
    // a gather4 lands 4 rows x 48 B = 192 B and needs a 128-byte aligned destination: every group of 4 rows gets 256 B
    __shared__ __align__(128) unsigned char slab_raw[WARPS][2][8][256];
    __shared__ __align__(8) uint64_t bar[WARPS][2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, w = blockIdx.x * WARPS + wib;
    if (lane == 0) {
        for (int b = 0; b < 2; ++b)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[wib][b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (w >= n_warps) return;
    const int *my = list + (size_t)w * PER_WARP;
    float acc = 0.f;
    auto issue = [&](int base, int b) {
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[wib][b])),
                         "r"(32 * (int)sizeof(Rec)) : "memory");
        __syncwarp();
        if (lane < 8) {
            const int4 rows = *reinterpret_cast<const int4 *>(my + base + 4 * lane);
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
                " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(&slab_raw[wib][b][lane][0])),
                "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&bar[wib][b])), "r"(0), "r"(rows.x), "r"(rows.y),
                "r"(rows.z), "r"(rows.w) : "memory");
        }
    };
    issue(0, 0);
    int phase[2] = {0, 0};
    for (int base = 0, b = 0; base < PER_WARP; base += 32, b ^= 1) {
        if (base + 32 < PER_WARP) issue(base + 32, b ^ 1);
        uint32_t done = 0;
        for (int spin = 0; spin < (1 << 22) && !done; ++spin)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&bar[wib][b])), "r"(phase[b]) : "memory");
        if (!done) { if (lane == 0) atomicAdd(err, 1); return; }      // never hang the box
        phase[b] ^= 1;
        acc = consume_groups(slab_raw[wib][b], lane, acc);
        __syncwarp();                                                  // the slab is free for the next gather
    }
    out[(size_t)w * 32 + lane] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int N = 1002528, n_warps = 4096;            // 4 M list entries, as one view of the bench scene
    const size_t E = (size_t)n_warps * PER_WARP;
    std::vector<Rec> h_rec(N);
    std::vector<int> h_list(E);
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    for (int i = 0; i < N; ++i) {
        float *f = reinterpret_cast<float *>(&h_rec[i]);
        for (int k = 0; k < 12; ++k) f[k] = (float)(rnd() % 1000) * 1e-3f;
    }
    for (size_t i = 0; i < E; ++i) h_list[i] = (int)(((i / 512) * 1237u + rnd() % 6000u) % N);   // a tile's Gaussians: a window
    Rec *d_rec; int *d_list, *d_err; float *d_a, *d_b;
    CK(cudaMalloc(&d_rec, sizeof(Rec) * N)); CK(cudaMalloc(&d_list, sizeof(int) * E)); CK(cudaMalloc(&d_err, 4));
    CK(cudaMalloc(&d_a, 4 * n_warps * 32)); CK(cudaMalloc(&d_b, 4 * n_warps * 32));
    CK(cudaMemcpy(d_rec, h_rec.data(), sizeof(Rec) * N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_list, h_list.data(), sizeof(int) * E, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_err, 0, 4)); CK(cudaMemset(d_b, 0, 4 * n_warps * 32));

    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time_it = [&](auto launch) -> float {
        for (int i = 0; i < 3; ++i) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        return ms / 20;
    };
    const int grid = (n_warps + WARPS - 1) / WARPS;
    const float ms_a = time_it([&] { stage_lanes<<<grid, 32 * WARPS>>>(d_rec, d_list, n_warps, d_a); });
    CK(cudaDeviceSynchronize());

    // tensor map: [N rows, 12 floats]; box rows for gather4 tried as 1 then 4
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &q));
    float ms_b = -1.f; int tma_err = -1, box_rows_used = 0; const char *why = "";
    for (int box_rows : {1, 4}) {
        CUtensorMap tmap;
        cuuint64_t dims[2] = {12, (cuuint64_t)N}, strides[1] = {sizeof(Rec)};
        cuuint32_t box[2] = {12, (cuuint32_t)box_rows}, estr[2] = {1, 1};
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_rec, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { why = "cuTensorMapEncodeTiled failed"; continue; }
        cudaMemset(d_err, 0, 4); cudaMemset(d_b, 0, 4 * n_warps * 32);
        stage_tma<<<grid, 32 * WARPS>>>(tmap, d_list, n_warps, d_b, d_err);
        cudaError_t ce = cudaDeviceSynchronize();
        if (ce != cudaSuccess) { why = cudaGetErrorString(ce); break; }        // sticky: stop here
        cudaMemcpy(&tma_err, d_err, 4, cudaMemcpyDeviceToHost);
        std::vector<float> a(n_warps * 32), b(n_warps * 32);
        cudaMemcpy(a.data(), d_a, 4 * a.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(b.data(), d_b, 4 * b.size(), cudaMemcpyDeviceToHost);
        size_t bad = 0; for (size_t i = 0; i < a.size(); ++i) bad += a[i] != b[i];
        if (tma_err == 0 && bad == 0) {
            box_rows_used = box_rows;
            ms_b = time_it([&] { stage_tma<<<grid, 32 * WARPS>>>(tmap, d_list, n_warps, d_b, d_err); });
            cudaDeviceSynchronize();
            break;
        }
        why = tma_err ? "mbarrier never completed" : "results differ from the lane-staged pass";
    }
    const double gb = (double)E * sizeof(Rec) / 1e9;
    printf("{\"entries\": %zu, \"bytes_staged\": %.0f, \"lanes_ms\": %.4f, \"lanes_gbs\": %.1f, \"tma_ms\": %.4f, \"tma_gbs\": %.1f, "
           "\"tma_box_rows\": %d, \"tma_ran\": %s, \"note\": \"%s\"}\n",
           E, gb * 1e9, ms_a, gb / (ms_a * 1e-3), ms_b, ms_b > 0 ? gb / (ms_b * 1e-3) : 0.0, box_rows_used,
           ms_b > 0 ? "true" : "false", why);
    return 0;
}
