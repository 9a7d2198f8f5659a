// Alpha compositing forward / backward.
//
// Work decomposition (B200): the unit of work is one WARP owning an 8x4 pixel sub-rectangle of a 16x16 tile
// (8 units per tile, ~20k units for 800x800).  Three kernels:
//   pack_records    : per Gaussian, one 48-byte record {x, y, hx, hy | conic a,b,c, opacity | r, g, b, -} where
//                     (hx, hy) is the axis-aligned half extent of the region in which alpha >= 1/255;
//   build_sublists  : per (tile, sub-rectangle), the tile's depth-sorted list filtered to the records whose
//                     extent overlaps the sub-rectangle (order preserved, ballot + popc compaction);
//   composite fwd/bwd: each warp walks ITS sub-list only, 32 records at a time staged in a private shared-memory
//                     slab with register prefetch of the next 32 -- no block-level barrier anywhere, per-warp
//                     early termination, and the hardware block scheduler balances ~20k small units.
// GeoSplatting's Gaussians are a few pixels wide, so a sub-list holds ~1/3 of its tile's list.  Filtering never
// changes results: a filtered pair is exactly one the reference kernel would `continue` on (alpha < 1/255 at
// every pixel centre of the sub-rectangle), and `last_ids` still indexes the 16x16 tile list.
// Backward: same units and sub-lists; pixel-parallel recurrence and Gaussian-parallel gradient accumulation are
// separated by a transpose through shared memory (see composite_bwd_kernel), one atomic per value per (Gaussian, warp).
//
// Replaces gsplat 1.4.0 rasterize_to_pixels_fwd/bwd (third-party), reached from
// rfstudio/model/gsplat.py:334-355.  Bound: FP32 / MUFU issue, not HBM.
#include "gsb_common.cuh"

#define LOG2E 1.4426950408889634f

namespace {

constexpr int SUB_W = 8, SUB_H = 4;                       // pixel footprint of one warp
constexpr int SUBS = (GSB_TILE / SUB_W) * (GSB_TILE / SUB_H);  // 8 sub-rectangles per tile
#ifndef GSB_WPB
#define GSB_WPB 2
#endif
#ifndef GSB_FG
#define GSB_FG 8
#endif
#ifndef GSB_BG
#define GSB_BG 8
#endif
#ifndef GSB_WPB_B
#define GSB_WPB_B 2
#endif
constexpr int WPB = GSB_WPB;                              // warps (units) per CTA in the composite kernels
constexpr int FG = GSB_FG;                                // entries evaluated together in the forward

struct Rec {
    float4 k;  // x, y, hx, hy
    float4 q;  // conic a, b, c, opacity
    float4 c;  // r, g, b, unused
};

__device__ __forceinline__ float ex2_approx(float x) {
#ifdef GSB_NO_INLINE_PTX   // host compilation of this file by tests/emu
    return exp2f(x);
#else
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}

// Half extents of {d : 0.5 d^T C d <= tau}, tau = ln(255 * opac): the only region where alpha >= 1/255.
// Negative when the Gaussian can contribute nowhere.  Inflated so that rounding in the per-pixel evaluation can
// never contradict a cull.
__device__ __forceinline__ float2 alpha_extent(float ca, float cb, float cc, float opac) {
    float t = 255.0f * opac;
    if (!(t > 1.0f)) return make_float2(-1e30f, -1e30f);
    float tau2 = 2.0f * logf(t);
    float det = ca * cc - cb * cb;
    if (!(det > 0.f)) return make_float2(1e30f, 1e30f);  // degenerate conic: never cull
    float hx = sqrtf(tau2 * cc / det), hy = sqrtf(tau2 * ca / det);
    return make_float2(hx * 1.0005f + 0.02f, hy * 1.0005f + 0.02f);
}

template <int CH>
__global__ void __launch_bounds__(256) pack_records_kernel(int N, const float2 *__restrict__ means2d,
                                                            const float *__restrict__ conics,
                                                            const float *__restrict__ colors,
                                                            const float *__restrict__ opacities, int opacity_is_logit,
                                                            const float *__restrict__ comps, Rec *__restrict__ rec) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    float2 xy = means2d[g];
    float ca = conics[3 * g], cb = conics[3 * g + 1], cc = conics[3 * g + 2];
    float op = opacities[g];
    if (opacity_is_logit) op = 1.0f / (1.0f + expf(-op));   // torch.sigmoid (rfstudio/model/gsplat.py:338)
    if (comps) op *= comps[g];                               // antialiasing compensation (gsplat: opacities * compensations)
    float2 ext = alpha_extent(ca, cb, cc, op);
    Rec r;
    r.k = make_float4(xy.x, xy.y, ext.x, ext.y);
    r.q = make_float4(ca, cb, cc, op);
    float c3[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < (CH < 3 ? CH : 3); ++k) c3[k] = colors[(size_t)g * CH + k];
    r.c = make_float4(c3[0], c3[1], c3[2], 0.f);
    rec[g] = r;
}

// One CTA (8 warps) per tile walks the tile's depth-sorted list ONCE, 256 entries per step, and appends every entry
// to the sub-lists of the sub-rectangles its extent overlaps, order preserved (per sub-rectangle: ballot + popc
// inside a warp, an 8 x 8 table of warp counts across the CTA).  Loads are software-pipelined two steps deep
// (list entry -> record is a dependent gather).  Sub-list w of tile t lives at entries[SUBS * start_t + w * len_t ...]
// (worst-case capacity, no global prefix sum needed).
constexpr int BUILD_THREADS = 256;

__global__ void __launch_bounds__(BUILD_THREADS) build_sublists_kernel(int tile_w, int n_tiles, int M,
                                                                        const int32_t *__restrict__ offsets,
                                                                        const int32_t *__restrict__ flatten_ids,
                                                                        const Rec *__restrict__ rec,
                                                                        int2 *__restrict__ entries,
                                                                        int32_t *__restrict__ counts) {
    static_assert(SUBS == 8 && BUILD_THREADS == 256, "build_sublists: 8 warps x 8 sub-rectangles");
    __shared__ int s_cnt[2][8][SUBS];   // [parity][warp][sub-rectangle] hits of this step
    __shared__ int s_pre[2][8][SUBS];   // exclusive prefix over warps
    __shared__ int s_base[2][SUBS];     // sub-list length before this step
    const int tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int start = offsets[tile];
    const int end = (tile == n_tiles - 1) ? M : offsets[tile + 1];
    const int len = end - start;
    if (len <= 0) {
        if (tid < SUBS) counts[tile * SUBS + tid] = 0;
        return;
    }
    const int tx = tile % tile_w, ty = tile / tile_w;
    const float x0 = (float)(tx * GSB_TILE) + 0.5f * SUB_W, y0 = (float)(ty * GSB_TILE) + 0.5f * SUB_H;
    const float rhx = 0.5f * (SUB_W - 1), rhy = 0.5f * (SUB_H - 1);
    int2 *const out = entries + (size_t)SUBS * start;
    if (tid < SUBS) s_base[0][tid] = 0;

    const float4 miss = make_float4(0.f, 0.f, -1e30f, -1e30f);
    // pipeline: gid of step i+2 and record of step i+1 are in flight while step i is processed
    int gid0 = (start + tid < end) ? flatten_ids[start + tid] : -1;
    int gid1 = (start + BUILD_THREADS + tid < end) ? flatten_ids[start + BUILD_THREADS + tid] : -1;
    float4 k0 = (gid0 >= 0) ? __ldg(&rec[gid0].k) : miss;
    int par = 0;
    for (int base = start; base < end; base += BUILD_THREADS, par ^= 1) {
        const int p2 = base + 2 * BUILD_THREADS + tid;
        const int gid2 = (p2 < end) ? flatten_ids[p2] : -1;
        const float4 k1 = (gid1 >= 0) ? __ldg(&rec[gid1].k) : miss;

        unsigned hits = 0u;          // bit w: this entry overlaps sub-rectangle w
        int my_pre[SUBS];            // hits of lower lanes of my warp
#pragma unroll
        for (int w = 0; w < SUBS; ++w) {
            const float rcx = x0 + (float)((w & 1) * SUB_W), rcy = y0 + (float)((w >> 1) * SUB_H);
            const bool hit = (fabsf(k0.x - rcx) <= k0.z + rhx) && (fabsf(k0.y - rcy) <= k0.w + rhy);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            my_pre[w] = __popc(m & ((1u << lane) - 1u));
            if (hit) hits |= 1u << w;
            if (lane == w) s_cnt[par][warp][w] = __popc(m);
        }
        __syncthreads();
        if (tid < 8 * SUBS) {
            const int wp = tid >> 3, w = tid & 7;
            int pre = 0, tot = 0;
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const int c = s_cnt[par][v][w];
                pre += (v < wp) ? c : 0;
                tot += c;
            }
            s_pre[par][wp][w] = pre;
            if (wp == 0) s_base[par ^ 1][w] = s_base[par][w] + tot;
        }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < SUBS; ++w) {
            if (hits & (1u << w))
                out[(size_t)w * len + s_base[par][w] + s_pre[par][warp][w] + my_pre[w]] = make_int2(base + tid, gid0);
        }
        gid0 = gid1; gid1 = gid2; k0 = k1;
    }
    // `par` now names the buffer the last step wrote its totals to
    __syncthreads();
    if (tid < SUBS) counts[tile * SUBS + tid] = s_base[par][tid];
}

// Longest-processing-time-first order of the TILES (their 8 units stay adjacent in launch order so that they
// share the tile's records in L1/L2): single-CTA counting sort on the mean per-unit work (4096 buckets of width 4,
// heaviest first; order inside a bucket is irrelevant).  `n_units` here is the number of tiles.
__device__ __forceinline__ int tile_work(const int32_t *__restrict__ counts, int tile) {
    int s = 0;
#pragma unroll
    for (int k = 0; k < SUBS; ++k) s += counts[tile * SUBS + k];
    return s >> 3;  // mean sub-list length of the tile's 8 units
}

__global__ void __launch_bounds__(1024) lpt_order_kernel(int n_units, const int32_t *__restrict__ counts,
                                                          int32_t *__restrict__ order) {
    constexpr int NB = 4096;
    __shared__ int hist[NB];
    __shared__ int warp_tot[32];
    const int tid = threadIdx.x;
    for (int i = tid; i < NB; i += 1024) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n_units; i += 1024) atomicAdd(&hist[NB - 1 - min(tile_work(counts, i) >> 2, NB - 1)], 1);
    __syncthreads();
    // exclusive scan of hist: each thread owns 4 consecutive buckets
    int v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = hist[tid * 4 + k]; sum += v[k]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        int w = warp_tot[tid], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (tid >= o) wi += t;
        }
        warp_tot[tid] = wi - w;
    }
    __syncthreads();
    int base = warp_tot[tid >> 5] + incl - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k) { hist[tid * 4 + k] = base; base += v[k]; }
    __syncthreads();
    for (int i = tid; i < n_units; i += 1024) {
        int pos = atomicAdd(&hist[NB - 1 - min(tile_work(counts, i) >> 2, NB - 1)], 1);
        order[pos] = i;
    }
}

struct Unit {
    int tile, w, i, j;
    bool inside;
    float px, py;
};

__device__ __forceinline__ Unit make_unit(int unit, int lane, int tile_w, int W, int H) {
    Unit u;
    u.tile = unit / SUBS;
    u.w = unit % SUBS;
    const int tx = u.tile % tile_w, ty = u.tile / tile_w;
    u.j = tx * GSB_TILE + (u.w & 1) * SUB_W + (lane & 7);
    u.i = ty * GSB_TILE + (u.w >> 1) * SUB_H + (lane >> 3);
    u.inside = (u.i < H && u.j < W);
    u.px = (float)u.j + 0.5f;
    u.py = (float)u.i + 0.5f;
    return u;
}

template <int CH>
__global__ void __launch_bounds__(32 * WPB)
composite_fwd_kernel(int W, int H, int tile_w, int n_units, const Rec *__restrict__ rec,
                     const float *__restrict__ colors, const float *__restrict__ background,
                     const int32_t *__restrict__ offsets, int n_tiles, int M, const int2 *__restrict__ entries,
                     const int32_t *__restrict__ counts, const int32_t *__restrict__ order,
                     int32_t *__restrict__ work, float *__restrict__ render, float *__restrict__ alphas,
                     int32_t *__restrict__ last_ids) {
    __shared__ Rec s_rec[WPB][32];
    __shared__ int2 s_ent[WPB][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int slot = blockIdx.x * WPB + wib;
    if (slot >= n_units) return;
    const int unit = order[slot / SUBS] * SUBS + (slot % SUBS);   // heaviest tiles are scheduled first (LPT)
    const Unit u = make_unit(unit, lane, tile_w, W, H);
    bool done = !u.inside;

    const int start = offsets[u.tile];
    const int end = (u.tile == n_tiles - 1) ? M : offsets[u.tile + 1];
    const int2 *list = entries + (size_t)SUBS * start + (size_t)u.w * (end - start);
    const int n = counts[unit];

    float T = 1.0f;
    float acc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = 0.f;
    int last_k = -1;   // sub-list index of the pixel's last contributor

    // prefetch chunk 0
    int2 e = make_int2(0, 0);
    Rec r;
    int processed = n;
    if (lane < n) { e = list[lane]; r = rec[e.y]; }
    for (int base = 0; base < n; base += 32) {
        __syncwarp();
        s_ent[wib][lane] = e;
        s_rec[wib][lane] = r;
        __syncwarp();
        const int nb = base + 32;
        if (nb + lane < n) { e = list[nb + lane]; r = rec[e.y]; }   // next chunk in flight during the loop
        const int cnt = min(32, n - base);
        // Groups of FG entries.  The FG alpha evaluations are independent (ILP for a warp that runs alone: the longest
        // sub-list of a view is this kernel's critical path); the transmittance chain is one multiply per entry.
        // alpha == 0 stands for "does not contribute" and makes every update the identity, so the common case has no
        // branch.  Only a group in which some pixel reaches the T <= 1e-4 stop replays the exact per-entry logic.
        bool all_done = false;
#pragma unroll 1
        for (int t0 = 0; t0 < cnt; t0 += FG) {
            float a[FG];
#pragma unroll
            for (int jj = 0; jj < FG; ++jj) {
                const int t = min(t0 + jj, 31);
                const float4 kk = s_rec[wib][t].k;
                const float4 q = s_rec[wib][t].q;
                const float dx = kk.x - u.px, dy = kk.y - u.py;
                const float sigma = 0.5f * (q.x * dx * dx + q.z * dy * dy) + q.y * dx * dy;
                const float alpha = fminf(GSB_ALPHA_CLAMP, q.w * ex2_approx(-LOG2E * sigma));
                a[jj] = (done || sigma < 0.f || alpha < GSB_ALPHA_MIN || t0 + jj >= cnt) ? 0.f : alpha;
            }
            float Tend = T;
#pragma unroll
            for (int jj = 0; jj < FG; ++jj) Tend *= 1.0f - a[jj];
            // T only decreases, so a pixel trips the stop inside this group iff the group's end value is below it
            if (!__any_sync(0xffffffffu, Tend <= GSB_T_STOP)) {
#pragma unroll
                for (int jj = 0; jj < FG; ++jj) {
                    const float alpha = a[jj];
                    const int t = min(t0 + jj, 31);
                    const float vis = alpha * T;
                    T *= 1.0f - alpha;
                    if (alpha > 0.f) {   // predicated: a non-contributing Gaussian's colour is never read into the sum
                        const float4 c = s_rec[wib][t].c;
                        acc[0] += c.x * vis;
                        if (CH > 1) acc[1] += c.y * vis;
                        if (CH > 2) acc[2] += c.z * vis;
                        if (CH > 3) {
                            const int g = s_ent[wib][t].y;
#pragma unroll
                            for (int k = 3; k < CH; ++k) acc[k] += __ldg(colors + (size_t)g * CH + k) * vis;
                        }
                        last_k = base + t;
                    }
                }
                continue;
            }
#pragma unroll
            for (int jj = 0; jj < FG; ++jj) {
                const float alpha = a[jj];
                if (done || alpha == 0.f) continue;
                const int t = t0 + jj;
                const float next_T = T * (1.0f - alpha);
                if (next_T <= GSB_T_STOP) {   // this entry is excluded
                    done = true;
                    continue;
                }
                const float vis = alpha * T;
                const float4 c = s_rec[wib][t].c;
                acc[0] += c.x * vis;
                if (CH > 1) acc[1] += c.y * vis;
                if (CH > 2) acc[2] += c.z * vis;
                if (CH > 3) {
                    const int g = s_ent[wib][t].y;
#pragma unroll
                    for (int k = 3; k < CH; ++k) acc[k] += __ldg(colors + (size_t)g * CH + k) * vis;
                }
                last_k = base + t;
                T = next_T;
            }
            if (__all_sync(0xffffffffu, done)) { all_done = true; break; }
        }
        if (all_done) { processed = min(n, base + 32); break; }
    }
    int cur_idx = 0;
    if (last_k >= 0) cur_idx = list[last_k].x;
    if (lane == 0) work[unit] = processed;   // entries actually walked: the backward's work estimate
    if (u.inside) {
        const size_t pix = (size_t)u.i * W + u.j;
        alphas[pix] = 1.0f - T;
#pragma unroll
        for (int k = 0; k < CH; ++k)
            render[pix * CH + k] = background ? acc[k] + T * background[k] : acc[k];
        last_ids[pix] = cur_idx;
    }
}

__device__ __forceinline__ float rcp_approx(float x) {
#ifdef GSB_NO_INLINE_PTX
    return 1.0f / x;
#else
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}

// Backward.  The per-Gaussian gradient is a sum over pixels, the transmittance recurrence runs over Gaussians: the
// kernel does each along the axis where it is register-local and TRANSPOSES through shared memory in between, so no
// warp shuffle and no cross-lane reduction is left.  Per chunk of 32 sub-list entries (walked back to front):
//   phase A (lane = pixel)   : evaluate the 32 entries, run the T / suffix-colour recurrence, and store per (entry,
//                              pixel) the three scalars the gradient needs -- vis = exp(-sigma) (0 if the pair does not
//                              contribute), T before the Gaussian, E = (suffix . v_out - T_final (v_alpha - bg . v_out))
//                              / (1 - alpha) -- into 32 x 33 slabs (row = entry, conflict-free both ways);
//   phase B (lane = Gaussian): read its row, accumulate the 9 gradient values over the 32 pixels in registers, one
//                              atomic per value per (Gaussian, warp) -- only for Gaussians that touched a pixel.
// v_alpha = T (c . v_out) - E is the reference's expression ((c T - buffer / (1 - alpha)) . v_out + T_final / (1 - alpha)
// (v_alpha_out - bg . v_out)) with the per-pixel constants folded into E.
constexpr int BSTRIDE = 33;
constexpr int BG = GSB_BG;   // entries evaluated together in phase A
constexpr int WPB_B = GSB_WPB_B;   // warps per CTA in the backward (15 KB of shared memory per warp)

template <int CH>
__global__ void __launch_bounds__(32 * WPB_B)
composite_bwd_kernel(int W, int H, int tile_w, int n_units, const Rec *__restrict__ rec,
                     const float *__restrict__ colors, const float *__restrict__ background,
                     const int32_t *__restrict__ offsets, int n_tiles, int M, const int2 *__restrict__ entries,
                     const int32_t *__restrict__ counts, const int32_t *__restrict__ order,
                     const float *__restrict__ alphas,
                     const int32_t *__restrict__ last_ids, const float *__restrict__ v_render,
                     const float *__restrict__ v_alphas, float *__restrict__ v_means2d, float *__restrict__ v_conics,
                     float *__restrict__ v_colors, float *__restrict__ v_opacities) {
    constexpr int C3 = CH < 3 ? CH : 3;                 // channels carried in the packed record
    constexpr int NV4 = (CH + 3) / 4;                   // float4s of v_out per pixel
    __shared__ float s_vis[WPB_B][32 * BSTRIDE];
    __shared__ float s_T[WPB_B][32 * BSTRIDE];
    __shared__ float s_E[WPB_B][32 * BSTRIDE];
    __shared__ Rec s_rec[WPB_B][32];
    __shared__ int2 s_ent[WPB_B][32];
    __shared__ float4 s_vo[WPB_B][32][NV4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int slot = blockIdx.x * WPB_B + wib;
    if (slot >= n_units) return;
    const int unit = order[slot / SUBS] * SUBS + (slot % SUBS);
    const Unit u = make_unit(unit, lane, tile_w, W, H);
    const size_t pix = u.inside ? (size_t)u.i * W + u.j : 0;

    const int start = offsets[u.tile];
    const int end = (u.tile == n_tiles - 1) ? M : offsets[u.tile + 1];
    const int2 *list = entries + (size_t)SUBS * start + (size_t)u.w * (end - start);
    int n = counts[unit];

    const float T_final = u.inside ? 1.0f - alphas[pix] : 1.0f;
    float v_out[CH];
    float bg_dot = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        v_out[k] = u.inside ? v_render[pix * CH + k] : 0.f;
        if (background) bg_dot += background[k] * v_out[k];
    }
    const float c0 = T_final * ((u.inside ? v_alphas[pix] : 0.f) - bg_dot);
    const int bin_final = u.inside ? last_ids[pix] : -1;
    {
        float vo4[NV4 * 4];
#pragma unroll
        for (int k = 0; k < NV4 * 4; ++k) vo4[k] = (k < CH) ? v_out[k] : 0.f;
#pragma unroll
        for (int k = 0; k < NV4; ++k)
            s_vo[wib][lane][k] = make_float4(vo4[4 * k], vo4[4 * k + 1], vo4[4 * k + 2], vo4[4 * k + 3]);
    }
    int wmax = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));

    // Entries are sorted by tile-list position: drop the tail that lies behind every pixel's last contributor
    // (binary search for the first position > wmax).
    {
        int lo = 0, hi = n;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (list[mid].x <= wmax) lo = mid + 1; else hi = mid;
        }
        n = lo;
    }
    if (n == 0) return;

    float *const my_vis = s_vis[wib], *const my_T = s_T[wib], *const my_E = s_E[wib];
    const float bx = (float)(u.j - (lane & 7)) + 0.5f, by = (float)(u.i - (lane >> 3)) + 0.5f;  // pixel (0,0) of the unit
    float T = T_final;
    float B = 0.f;    // suffix colour behind the current Gaussian, dotted with v_out

    // walk back to front: chunk c covers sub-list indices [top - 32, top), slab row t holds index top - 1 - t
    int2 e = make_int2(0, 0);
    Rec r;
    r.k = r.q = r.c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n - 1 - lane >= 0) { e = list[n - 1 - lane]; r = rec[e.y]; }
    for (int top = n; top > 0; top -= 32) {
        __syncwarp();
        s_ent[wib][lane] = e;
        s_rec[wib][lane] = r;
        __syncwarp();
        const int nt = top - 32;
        if (nt - 1 - lane >= 0) { e = list[nt - 1 - lane]; r = rec[e.y]; }
        const int cnt = min(32, top);

        // ---- phase A: lane = pixel -------------------------------------------------------------------------
        unsigned touched = 0u;   // bit t: entry t contributes to at least one pixel of the unit
#pragma unroll 1
        for (int t0 = 0; t0 < cnt; t0 += BG) {
            // independent per entry: alpha, 1 / (1 - alpha), colour . v_out  (ILP for a warp that runs alone)
            float visg[BG], alg[BG], rag[BG], wg[BG];
#pragma unroll
            for (int jj = 0; jj < BG; ++jj) {
                const int t = min(t0 + jj, 31);
                const float4 kk = s_rec[wib][t].k;
                const float4 q = s_rec[wib][t].q;
                const float4 c = s_rec[wib][t].c;
                const int2 en = s_ent[wib][t];
                const float dx = kk.x - u.px, dy = kk.y - u.py;
                const float sigma = 0.5f * (q.x * dx * dx + q.z * dy * dy) + q.y * dx * dy;
                const float vis = ex2_approx(-LOG2E * sigma);
                const float alpha = fminf(GSB_ALPHA_CLAMP, q.w * vis);
                const bool valid = (t0 + jj < cnt) && (en.x <= bin_final) && !(sigma < 0.f || alpha < GSB_ALPHA_MIN);
                visg[jj] = valid ? vis : 0.f;
                alg[jj] = valid ? alpha : 0.f;
                rag[jj] = rcp_approx(1.0f - alg[jj]);
                float w = c.x * v_out[0];
                if (C3 > 1) w += c.y * v_out[1];
                if (C3 > 2) w += c.z * v_out[2];
                if (CH > 3) {
#pragma unroll
                    for (int k = 3; k < CH; ++k) w += __ldg(colors + (size_t)en.y * CH + k) * v_out[k];
                }
                wg[jj] = w;
                if (__any_sync(0xffffffffu, valid)) touched |= 1u << t;
            }
            // the recurrence: alpha == 0 (pair does not contribute) makes every update the identity, so no branch
#pragma unroll
            for (int jj = 0; jj < BG; ++jj) {
                const int t = min(t0 + jj, 31);
                T *= rag[jj];
                my_vis[t * BSTRIDE + lane] = visg[jj];
                my_T[t * BSTRIDE + lane] = T;
                my_E[t * BSTRIDE + lane] = rag[jj] * (B - c0);
                B += wg[jj] * (alg[jj] * T);
            }
        }
        __syncwarp();
        if (touched == 0u) continue;

        // ---- phase B: lane = Gaussian (slab row `lane`) ------------------------------------------------------
        {
            const float4 kk = s_rec[wib][lane].k;
            const float4 q = s_rec[wib][lane].q;
            const float4 c = s_rec[wib][lane].c;
            const int g = s_ent[wib][lane].y;
            float col[CH];
            col[0] = c.x;
            if (CH > 1) col[1] = c.y;
            if (CH > 2) col[2] = c.z;
            const bool mine = (touched >> lane) & 1u;
            if (CH > 3) {
#pragma unroll
                for (int k = 3; k < CH; ++k) col[k] = mine ? __ldg(colors + (size_t)g * CH + k) : 0.f;
            }
            float g_col[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) g_col[k] = 0.f;
            float sxx = 0.f, sxy = 0.f, syy = 0.f, sx = 0.f, sy = 0.f, g_op = 0.f;
            const float *row_vis = my_vis + lane * BSTRIDE, *row_T = my_T + lane * BSTRIDE,
                        *row_E = my_E + lane * BSTRIDE;
#pragma unroll 8
            for (int p = 0; p < 32; ++p) {
                const float vis = row_vis[p], Tp = row_T[p], Ep = row_E[p];
                float vo[NV4 * 4];
#pragma unroll
                for (int k = 0; k < NV4; ++k) {
                    const float4 v4 = s_vo[wib][p][k];
                    vo[4 * k] = v4.x; vo[4 * k + 1] = v4.y; vo[4 * k + 2] = v4.z; vo[4 * k + 3] = v4.w;
                }
                const float ov = q.w * vis;
                const float fac = fminf(GSB_ALPHA_CLAMP, ov) * Tp;
                float w = 0.f;
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    w += col[k] * vo[k];
                    g_col[k] += fac * vo[k];
                }
                float v_alpha = Tp * w - Ep;
                v_alpha = (ov <= GSB_ALPHA_CLAMP) ? v_alpha : 0.f;   // clamped alpha passes no gradient to sigma / opacity
                g_op += vis * v_alpha;
                const float v_sigma = -ov * v_alpha;
                const float dx = kk.x - (bx + (float)(p & 7)), dy = kk.y - (by + (float)(p >> 3));   // exact pixel centre
                const float t1 = v_sigma * dx, t2 = v_sigma * dy;
                sxx += t1 * dx;
                sxy += t1 * dy;
                syy += t2 * dy;
                sx += t1;
                sy += t2;
            }
            if (mine) {
#pragma unroll
                for (int k = 0; k < CH; ++k) atomicAdd(v_colors + (size_t)g * CH + k, g_col[k]);
                atomicAdd(v_conics + 3 * (size_t)g, 0.5f * sxx);
                atomicAdd(v_conics + 3 * (size_t)g + 1, sxy);
                atomicAdd(v_conics + 3 * (size_t)g + 2, 0.5f * syy);
                atomicAdd(v_means2d + 2 * (size_t)g, q.x * sx + q.y * sy);
                atomicAdd(v_means2d + 2 * (size_t)g + 1, q.y * sx + q.z * sy);
                atomicAdd(v_opacities + g, g_op);
            }
        }
    }
}

struct Workspace {
    Rec *rec;
    int32_t *counts, *work, *order;   // sub-list lengths, entries the forward walked (the backward's LPT key), tile order
    int2 *entries;
};


size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t workspace_bytes(int64_t N, int64_t M, int n_tiles) {
    return align256(sizeof(Rec) * (size_t)N) + 3 * align256(sizeof(int32_t) * (size_t)n_tiles * SUBS) +
           align256(sizeof(int2) * (size_t)SUBS * (size_t)M) +
           256;
}

Workspace carve(void *ws, int64_t N, int64_t M, int n_tiles) {
    char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    Workspace w;
    w.rec = reinterpret_cast<Rec *>(p);
    p += align256(sizeof(Rec) * (size_t)N);
    const size_t ub = align256(sizeof(int32_t) * (size_t)n_tiles * SUBS);
    w.counts = reinterpret_cast<int32_t *>(p); p += ub;
    w.work = reinterpret_cast<int32_t *>(p); p += ub;
    w.order = reinterpret_cast<int32_t *>(p); p += ub;
    w.entries = reinterpret_cast<int2 *>(p);
    (void)M;
    return w;
}

template <int CH>
int launch_fwd(int W, int H, int64_t N, const float *means2d, const float *conics, const float *colors,
               const float *opacities, int opacity_is_logit, const float *comps, const float *background,
               const int32_t *offsets, const int32_t *flatten_ids, int64_t M, float *render, float *alphas,
               int32_t *last_ids, void *ws, cudaStream_t st) {
    int tw = (W + GSB_TILE - 1) / GSB_TILE, th = (H + GSB_TILE - 1) / GSB_TILE;
    int n_tiles = tw * th, n_units = n_tiles * SUBS;
    Workspace w = carve(ws, N, M, n_tiles);
    if (N > 0)
        pack_records_kernel<CH><<<gsb_div_up(N, 256), 256, 0, st>>>((int)N, reinterpret_cast<const float2 *>(means2d),
                                                                    conics, colors, opacities, opacity_is_logit,
                                                                    comps, w.rec);
    build_sublists_kernel<<<n_tiles, BUILD_THREADS, 0, st>>>(tw, n_tiles, (int)M, offsets, flatten_ids, w.rec,
                                                             w.entries, w.counts);
    lpt_order_kernel<<<1, 1024, 0, st>>>(n_tiles, w.counts, w.order);
    composite_fwd_kernel<CH><<<gsb_div_up(n_units, WPB), 32 * WPB, 0, st>>>(
        W, H, tw, n_units, w.rec, colors, background, offsets, n_tiles, (int)M, w.entries, w.counts, w.order,
        w.work, render, alphas, last_ids);
    return 0;
}

template <int CH>
int launch_bwd(int W, int H, int64_t N, const float *colors, const float *background, const int32_t *offsets,
               int64_t M, const float *alphas, const int32_t *last_ids, const float *v_render, const float *v_alphas,
               float *v_means2d, float *v_conics, float *v_colors, float *v_opacities, void *ws, cudaStream_t st) {
    int tw = (W + GSB_TILE - 1) / GSB_TILE, th = (H + GSB_TILE - 1) / GSB_TILE;
    int n_tiles = tw * th, n_units = n_tiles * SUBS;
    Workspace w = carve(ws, N, M, n_tiles);
    lpt_order_kernel<<<1, 1024, 0, st>>>(n_tiles, w.work, w.order);   // order by the forward's measured work
    composite_bwd_kernel<CH><<<gsb_div_up(n_units, WPB_B), 32 * WPB_B, 0, st>>>(
        W, H, tw, n_units, w.rec, colors, background, offsets, n_tiles, (int)M, w.entries, w.counts, w.order, alphas,
        last_ids, v_render, v_alphas, v_means2d, v_conics, v_colors, v_opacities);
    return 0;
}

}  // namespace

#define GSB_DISPATCH_CH(CHV, CALL)             \
    switch (CHV) {                             \
        case 1: { constexpr int C_ = 1; CALL; break; }   \
        case 2: { constexpr int C_ = 2; CALL; break; }   \
        case 3: { constexpr int C_ = 3; CALL; break; }   \
        case 4: { constexpr int C_ = 4; CALL; break; }   \
        case 8: { constexpr int C_ = 8; CALL; break; }   \
        case 16: { constexpr int C_ = 16; CALL; break; } \
        default:                               \
            gsb_set_error("%s: unsupported channel count %d (supported: 1,2,3,4,8,16; pad to the next)", __func__, CHV); \
            return GSB_EINVAL;                 \
    }

#define GSB_API extern "C" __attribute__((visibility("default")))

GSB_API int gsb_composite_workspace_bytes(int64_t N, int64_t M, int32_t width, int32_t height, size_t *bytes_host) {
    GSB_CHECK_ARG(N >= 0 && M >= 0 && width > 0 && height > 0 && bytes_host != nullptr);
    int tw = (width + GSB_TILE - 1) / GSB_TILE, th = (height + GSB_TILE - 1) / GSB_TILE;
    *bytes_host = workspace_bytes(N, M, tw * th);
    return GSB_OK;
}

GSB_API int gsb_composite_fwd(int32_t width, int32_t height, int32_t channels, int64_t N, const float *means2d,
                              const float *conics, const float *colors, const float *opacities,
                              int32_t opacity_is_logit, const float *comps, const float *background,
                              const int32_t *offsets, const int32_t *flatten_ids, int64_t M, float *render,
                              float *alphas, int32_t *last_ids, void *workspace, size_t workspace_bytes_,
                              void *stream) {
    GSB_CHECK_ARG(width > 0 && height > 0 && N >= 0 && M >= 0 && M < 268435455LL);
    GSB_CHECK_ARG(offsets && render && alphas && last_ids && workspace);
    GSB_CHECK_ARG(M == 0 || (means2d && conics && colors && opacities && flatten_ids));
    int tw = (width + GSB_TILE - 1) / GSB_TILE, th = (height + GSB_TILE - 1) / GSB_TILE;
    if (workspace_bytes(N, M, tw * th) > workspace_bytes_) {
        gsb_set_error("gsb_composite_fwd: workspace too small (%zu < %zu)", workspace_bytes_, workspace_bytes(N, M, tw * th));
        return GSB_ENOMEM;
    }
    int rc = 0;
    GSB_DISPATCH_CH(channels, (rc = launch_fwd<C_>(width, height, N, means2d, conics, colors, opacities,
                                                    opacity_is_logit, comps, background, offsets, flatten_ids, M,
                                                    render, alphas, last_ids, workspace, (cudaStream_t)stream)));
    if (rc != 0) {
        gsb_set_error("gsb_composite_fwd: internal sort scratch too small");
        return GSB_ENOMEM;
    }
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_composite_bwd(int32_t width, int32_t height, int32_t channels, int64_t N, const float *colors,
                              const float *background, const int32_t *offsets, int64_t M, const float *alphas,
                              const int32_t *last_ids, const float *v_render, const float *v_alphas,
                              float *v_means2d, float *v_conics, float *v_colors, float *v_opacities,
                              const void *workspace, void *stream) {
    GSB_CHECK_ARG(width > 0 && height > 0 && N >= 0 && M >= 0 && M < 268435455LL);
    GSB_CHECK_ARG(offsets && alphas && last_ids && v_render && v_alphas && workspace);
    if (M == 0) return GSB_OK;
    GSB_CHECK_ARG(colors && v_means2d && v_conics && v_colors && v_opacities);
    GSB_DISPATCH_CH(channels, (launch_bwd<C_>(width, height, N, colors, background, offsets, M, alphas, last_ids,
                                               v_render, v_alphas, v_means2d, v_conics, v_colors, v_opacities,
                                               const_cast<void *>(workspace), (cudaStream_t)stream)));
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}
