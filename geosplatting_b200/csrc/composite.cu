// Alpha compositing forward / backward.
//
// Work decomposition (B200): the unit of work is one WARP owning a 4x4 pixel sub-rectangle of a 16x16 tile (16 units per
// tile, ~40k units for 800x800), and the warp evaluates TWO list entries per step: lanes 0-15 hold the 16 pixels for
// the even entry of a pair, lanes 16-31 the same 16 pixels for the odd entry.  GeoSplatting's Gaussians are ~5 px wide:
// with 8x4 units (round 1) 87 % of the (entry, pixel) evaluations failed the alpha test; 4x4 units evaluate 1.5x fewer
// pairs, and pairing the entries keeps all 32 lanes busy.  Three kernels:
//   pack_records    : per Gaussian, one 48-byte record {x, y, hx, hy | qa, qb, qc, log2(opacity) | r, g, b, opacity}
//                     where (hx, hy) is the axis-aligned half extent of the region in which alpha >= 1/255 and
//                     (qa, qb, qc) = -log2(e) * (a/2, b, c/2) of the conic, so that alpha = 2^(qa dx^2 + qb dx dy +
//                     qc dy^2 + log2(opacity)) costs 5 FMA + one ex2 per pixel;
//   build_sublists  : per (tile, sub-rectangle), the tile's depth-sorted list filtered to the records whose
//                     extent overlaps the sub-rectangle (order preserved, ballot + popc compaction);
//   composite fwd/bwd: each warp walks ITS sub-list only, 32 records at a time staged in a private shared-memory
//                     slab with register prefetch of the next 32 -- no block-level barrier anywhere, per-warp
//                     early termination, and the hardware block scheduler balances ~40k small units.
// Filtering never changes results: a filtered pair is exactly one the reference kernel would `continue` on (alpha <
// 1/255 at every pixel centre of the sub-rectangle), and `last_ids` still indexes the 16x16 tile list.
// Backward: same units and sub-lists; pixel-parallel recurrence and Gaussian-parallel gradient accumulation are
// separated by a transpose through shared memory (see composite_bwd_kernel), one atomic per value per (Gaussian, warp).
//
// Replaces gsplat 1.4.0 rasterize_to_pixels_fwd/bwd (third-party), reached from
// rfstudio/model/gsplat.py:334-355.  Bound: FP32 / MUFU issue, not HBM.
#include <stdlib.h>

#include "gsb_common.cuh"
#include "composite_rec.cuh"

#define LOG2E 1.4426950408889634f
#define LOG2_ALPHA_MIN (-7.994353436858858f)   // log2(1/255)

namespace {

constexpr int SUB_W = 4, SUB_H = 4;                            // pixel footprint of one warp
constexpr int SUBS = (GSB_TILE / SUB_W) * (GSB_TILE / SUB_H);  // 16 sub-rectangles per tile
constexpr int PIX = SUB_W * SUB_H;                             // 16 pixels, held twice per warp
#ifndef GSB_WPB
#define GSB_WPB 4
#endif
#ifndef GSB_FG
#define GSB_FG 8
#endif
#ifndef GSB_BG
#define GSB_BG 16
#endif
#ifndef GSB_WPB_B
#define GSB_WPB_B 4
#endif
#ifndef GSB_MINB      // minimum resident CTAs per SM asked of ptxas (register cap) for the forward / backward kernels
#define GSB_MINB 1
#endif
#ifndef GSB_MINB_B
#define GSB_MINB_B 1
#endif
constexpr int WPB = GSB_WPB;                              // warps (pairs of units) per CTA in the composite kernels
constexpr int FSZ = GSB_FG;                               // entries evaluated together in the forward
constexpr int BSZ = GSB_BG;                               // entries evaluated together in the backward's phase A
static_assert(GSB_FG >= 1 && 16 % GSB_FG == 0, "GSB_FG: divisor of 16");
static_assert(GSB_BG >= 1 && 16 % GSB_BG == 0, "GSB_BG: divisor of 16");

using Rec = GsbRec;   // composite_rec.cuh

__device__ __forceinline__ float ex2_approx(float x) {
#ifdef GSB_NO_INLINE_PTX   // host compilation of this file by tests/emu
    return exp2f(x);
#else
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
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

// log2 of alpha before the 0.999 clamp at offset (dx, dy) from the centre; 5 FMA-class instructions
__device__ __forceinline__ float log2_alpha(const float4 q, float dx, float dy) {
    return fmaf(q.z * dy, dy, fmaf(fmaf(q.y, dy, q.x * dx), dx, q.w));
}

template <int CH>
__global__ void __launch_bounds__(256) pack_records_kernel(int N, const float2 *__restrict__ means2d,
                                                            const float *__restrict__ conics,
                                                            const float *__restrict__ colors,
                                                            const float *__restrict__ opacities, int opacity_is_logit,
                                                            const float *__restrict__ comps, Rec *__restrict__ rec) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    float c3[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < (CH < 3 ? CH : 3); ++k) c3[k] = colors[(size_t)g * CH + k];
    const Rec r = gsb_pack_record(means2d[g], conics[3 * g], conics[3 * g + 1], conics[3 * g + 2], c3[0], c3[1], c3[2],
                                  opacities[g], opacity_is_logit, comps ? comps[g] : 1.0f);
    rec[g] = r;
}

// One CTA per tile walks the tile's depth-sorted list ONCE, 512 entries per step, and appends every entry to the
// sub-lists of the sub-rectangles its extent overlaps, order preserved (per sub-rectangle: ballot + popc inside a
// warp, a 16 x 16 table of warp counts across the CTA).  Loads are software-pipelined two steps deep (list entry ->
// record is a dependent gather).  Sub-list w of tile t lives at entries[SUBS * start_t + w * len_t ...] (worst-case
// capacity, no global prefix sum needed).  32 registers per thread (the per-lane prefix counts are packed 5 bits each)
// keep four CTAs -- four tiles -- resident per SM: a tile's first step exposes the whole gather chain, and the tiles of
// the bench scene are 2.6 steps long on average.  Measured (composite_fwd entry point, alone): 1024 threads x 1 CTA
// 0.249 ms, 1024 x 2 0.243, 512 x 4 0.231, 256 x 8 0.242 (the 11.6 k-entry silhouette tiles become 46-step tails).
#ifndef GSB_BUILD_THREADS
#define GSB_BUILD_THREADS 512
#endif
#ifndef GSB_BUILD_MINB
#define GSB_BUILD_MINB 4    // 32 registers: four 512-thread CTAs (four tiles) per SM hide each other's gather latency
#endif
constexpr int BUILD_THREADS = GSB_BUILD_THREADS;
constexpr int BUILD_WARPS = BUILD_THREADS / 32;

__global__ void __launch_bounds__(BUILD_THREADS, GSB_BUILD_MINB) build_sublists_kernel(int tile_w, int n_tiles, int M_host,
                                                                        const int64_t *__restrict__ m_dev,
                                                                        const int32_t *__restrict__ offsets,
                                                                        const int32_t *__restrict__ flatten_ids,
                                                                        const Rec *__restrict__ rec,
                                                                        int2 *__restrict__ entries,
                                                                        int32_t *__restrict__ counts,
                                                                        int32_t *__restrict__ tile_len,
                                                                        int32_t *__restrict__ ctrl) {
    static_assert(SUBS == 16 && BUILD_THREADS % 32 == 0 && BUILD_WARPS * SUBS <= BUILD_THREADS, "build_sublists shape");
    __shared__ int s_cnt[2][BUILD_WARPS][SUBS];   // [parity][warp][sub-rectangle] hits of this step
    __shared__ int s_pre[2][BUILD_WARPS][SUBS];   // exclusive prefix over warps
    __shared__ int s_base[2][SUBS];     // sub-list length before this step
    const int tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = m_dev ? (int)*m_dev : M_host;
    const int start = offsets[tile];
    const int end = (tile == n_tiles - 1) ? M : offsets[tile + 1];
    const int len = end - start;
    if (tile == 0 && tid < 4) ctrl[tid] = 0;      // job count, checkpoint count, job cursor (the forward fills them)
    if (len <= 0) {
        if (tid < SUBS) counts[tile * SUBS + tid] = 0;
        if (tid == 0) tile_len[tile] = 0;
        return;
    }
    const int tx = tile % tile_w, ty = tile / tile_w;
    // centre of sub-rectangle column / row 0; columns and rows are SUB_W / SUB_H apart
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

        // 4 column tests x 4 row tests -> 16-bit overlap mask (bit w = row * 4 + column)
        unsigned colm = 0u, rowm = 0u;
#pragma unroll
        for (int c = 0; c < GSB_TILE / SUB_W; ++c)
            if (fabsf(k0.x - (x0 + (float)(c * SUB_W))) <= k0.z + rhx) colm |= 1u << c;
#pragma unroll
        for (int r = 0; r < GSB_TILE / SUB_H; ++r)
            if (fabsf(k0.y - (y0 + (float)(r * SUB_H))) <= k0.w + rhy) rowm |= 0xFu << (4 * r);
        const unsigned hits = (colm * 0x1111u) & rowm;
        unsigned my_pre[3] = {0u, 0u, 0u};   // hits of lower lanes of my warp, 5 bits per sub-rectangle (6 + 6 + 4)
#pragma unroll
        for (int w = 0; w < SUBS; ++w) {
            const unsigned m = __ballot_sync(0xffffffffu, (hits >> w) & 1u);
            my_pre[w / 6] |= (unsigned)__popc(m & ((1u << lane) - 1u)) << (5 * (w % 6));
            if (lane == w) s_cnt[par][warp][w] = __popc(m);
        }
        __syncthreads();
        if (tid < BUILD_WARPS * SUBS) {
            const int wp = tid >> 4, w = tid & (SUBS - 1);
            int pre = 0, tot = 0;
#pragma unroll
            for (int v = 0; v < BUILD_WARPS; ++v) {
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
                out[(size_t)w * len + s_base[par][w] + s_pre[par][warp][w] + (int)((my_pre[w / 6] >> (5 * (w % 6))) & 31u)] =
                    make_int2(base + tid, gid0);
        }
        gid0 = gid1; gid1 = gid2; k0 = k1;
    }
    // `par` now names the buffer the last step wrote its totals to
    __syncthreads();
    if (tid < SUBS) counts[tile * SUBS + tid] = s_base[par][tid];
    if (tid == 0) {
        // the forward's scheduling key: the tile's LONGEST sub-list (a warp is busy as long as its longer list; a
        // silhouette tile with one 2 500-entry unit and fourteen empty ones must start early, which its mean hides)
        int key = 0;
#pragma unroll
        for (int k = 0; k < SUBS; ++k) key = max(key, s_base[par][k]);
        tile_len[tile] = key;
    }
}

// Longest-processing-time-first order of the TILES (their 16 units stay adjacent in launch order so that they
// share the tile's records in L1/L2): single-CTA counting sort on the key build_sublists left per tile (4096 buckets
// of width 4, heaviest first; order inside a bucket is irrelevant).  `n_units` here is the number of tiles.
__device__ __forceinline__ int tile_work(const int32_t *__restrict__ per_tile, int tile) { return per_tile[tile]; }

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

// A warp owns TWO horizontally adjacent 4x4 units of a tile (an 8x4 pixel region) and walks BOTH sub-lists at once:
// lanes 0-15 are the pixels of unit 2k with its list, lanes 16-31 the pixels of unit 2k+1 with its own list.  The two
// half-warps never exchange data inside the loops (no shuffles, identical code on different lists); the warp runs as
// long as the longer list, and neighbouring units' lists differ by a few percent.
constexpr int UPW = 2;                 // units per warp
constexpr int CHUNK = 32 / UPW;        // list entries staged per unit and step (shared-memory rows 2 * i + half)

// Segments.  The backward walks a sub-list back to front, and the longest sub-list of a view (~2 500 entries at a
// silhouette, 156 steps on one warp) used to be the kernel's critical path.  The forward therefore CHECKPOINTS its
// per-pixel state {T, colour} every SEG entries of a sub-list longer than SEG, which makes every SEG-entry segment an
// independent work item of the backward: T before the segment's last entry comes from the checkpoint behind it, the
// colour accumulated behind it is (final - checkpoint).  Checkpoints exist for <= 3 channels (one float4 per pixel);
// wider splats are walked whole.
#ifndef GSB_SEG
#define GSB_SEG 256
#endif
constexpr int SEG = GSB_SEG;
static_assert(SEG >= CHUNK && SEG % CHUNK == 0, "GSB_SEG: multiple of 16");
template <int CH> constexpr bool segmented() { return CH <= 3; }
// control words in the workspace
enum { CTRL_JOBS = 0, CTRL_CKPTS = 1, CTRL_CURSOR = 2 };

struct Unit {
    int tile, w, i, j;
    bool inside;
    float px, py;
};

// lane -> (half h = lane >> 4 selects the unit, pixel p = lane & 15)
__device__ __forceinline__ Unit make_unit(int pair, int lane, int tile, int tile_w, int W, int H) {
    Unit u;
    u.tile = tile;
    u.w = pair * UPW + (lane >> 4);
    const int p = lane & (PIX - 1);
    const int tx = u.tile % tile_w, ty = u.tile / tile_w;
    u.j = tx * GSB_TILE + (u.w & 3) * SUB_W + (p & 3);
    u.i = ty * GSB_TILE + (u.w >> 2) * SUB_H + (p >> 2);
    u.inside = (u.i < H && u.j < W);
    u.px = (float)u.j + 0.5f;
    u.py = (float)u.i + 0.5f;
    return u;
}

__device__ __forceinline__ Rec null_record() {
    Rec r;
    r.k = make_float4(0.f, 0.f, -1e30f, -1e30f);
    r.q = make_float4(0.f, 0.f, 0.f, -1e30f);     // log2(alpha) = -1e30: fails the alpha test at every pixel
    r.c = make_float4(0.f, 0.f, 0.f, 0.f);
    return r;
}

constexpr float FAR_AWAY = 1e15f;   // pixel coordinate of a lane that must not contribute any more: log2(alpha) = -inf

template <int CH>
__global__ void __launch_bounds__(32 * WPB, GSB_MINB)
composite_fwd_kernel(int W, int H, int tile_w, int n_pairs, const Rec *__restrict__ rec,
                     const float *__restrict__ colors, const float *__restrict__ background,
                     const int32_t *__restrict__ offsets, int n_tiles, int M_host, const int64_t *__restrict__ m_dev,
                     const int2 *__restrict__ entries, const int32_t *__restrict__ counts,
                     const int32_t *__restrict__ order, int32_t *__restrict__ walk, int32_t *__restrict__ ckpt_base,
                     int32_t *ctrl, int2 *__restrict__ jobs, float4 *__restrict__ ckpt, float *__restrict__ render,
                     float *__restrict__ alphas, int32_t *__restrict__ last_ids) {
    constexpr int PAIRS = SUBS / UPW;
    constexpr bool SEGD = segmented<CH>();
    __shared__ Rec s_rec[WPB][32];     // k.z / k.w carry the entry's tile-list position / Gaussian id (bit patterns)
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int slot = blockIdx.x * WPB + wib;
    if (slot >= n_pairs) return;
    const int tile = order[slot / PAIRS];          // heaviest tiles are scheduled first (LPT)
    const int half = lane >> 4;
    const Unit u = make_unit(slot % PAIRS, lane, tile, tile_w, W, H);
    const int unit = tile * SUBS + u.w;
    bool done = !u.inside;
    float px = done ? FAR_AWAY : u.px;             // a finished pixel moves out of every Gaussian's reach
    const float py = u.py;

    const int M = m_dev ? (int)*m_dev : M_host;
    const int start = offsets[tile];
    const int end = (tile == n_tiles - 1) ? M : offsets[tile + 1];
    const int2 *list = entries + (size_t)SUBS * start + (size_t)u.w * (end - start);
    const int n = counts[unit];                                           // this half-warp's list
    const int n_max = max(n, __shfl_xor_sync(0xffffffffu, n, 16));        // the warp walks the longer one

    float T = 1.0f;
    float acc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = 0.f;
    int last_k = -1;   // sub-list index of the pixel's last contributor

    // prefetch chunk 0: lane (half, i) fetches entry i of its unit's list into shared-memory row 2 * i + half
    // staging: lane l fills shared-memory row l = entry l / 2 of unit l % 2 of the step (48-byte rows at a 48-byte lane
    // stride: conflict-free 16-byte stores), so a lane stages for the other half-warp's unit as often as for its own
    const int li = lane >> 1;
    const bool own = ((lane & 1) == half);
    const int2 *const list_other = (const int2 *)__shfl_xor_sync(0xffffffffu, (unsigned long long)list, 16);
    const int n_other = __shfl_xor_sync(0xffffffffu, n, 16);
    const int2 *const slist = own ? list : list_other;
    const int sn = own ? n : n_other;
    const Rec *const rows = &s_rec[wib][half];     // rows of my unit: rows[2 * s]
    // Software pipeline of the dependent gather (list entry -> record): while chunk c is evaluated, the records of
    // chunk c + 1 and the list entries of chunk c + 2 are in flight, so neither latency is on the warp's critical path.
    int2 e = make_int2(0, 0), e_next = make_int2(0, 0);
    Rec r = null_record();
    int cb = -1;       // first checkpoint block of my unit (allocated when the walk passes entry SEG)
    if (li < sn) { e = slist[li]; r = rec[e.y]; }
    if (CHUNK + li < sn) e_next = slist[CHUNK + li];
    for (int base = 0; base < n_max; base += CHUNK) {
        if (SEGD && base > 0 && base % SEG == 0) {
            // state BEFORE entry `base` -> checkpoint base / SEG of my unit (slot 0 takes the final state, below)
            if (base == SEG) {
                int b = -1;
                if ((lane & (PIX - 1)) == 0 && n > SEG) b = atomicAdd(ctrl + CTRL_CKPTS, (n + SEG - 1) / SEG);
                cb = __shfl_sync(0xffffffffu, b, lane & 16);
            }
            if (cb >= 0 && base < n)
                ckpt[((size_t)cb + base / SEG) * PIX + (lane & (PIX - 1))] =
                    make_float4(T, acc[0], CH > 1 ? acc[CH > 1 ? 1 : 0] : 0.f, CH > 2 ? acc[CH > 2 ? 2 : 0] : 0.f);
        }
        __syncwarp();
        r.k.z = __int_as_float(e.x);
        r.k.w = __int_as_float(e.y);
        s_rec[wib][lane] = r;
        __syncwarp();
        const int nb = base + CHUNK;
        e = e_next;
        r = null_record();                                            // rows past the end of a list never contribute
        if (nb + li < sn) r = rec[e.y];                                // records of the next chunk
        if (nb + CHUNK + li < sn) e_next = slist[nb + CHUNK + li];    // list entries of the chunk after it
        // Groups of FS entries.  The alpha evaluations are independent (ILP for a warp that runs alone: the longest
        // sub-list of a view is this kernel's critical path); alpha == 0 stands for "does not contribute" and makes
        // every update the identity, so the common case has no branch.  The state is saved before a group; only a group
        // in which some pixel reaches the T <= 1e-4 stop is rolled back and replayed with the exact per-entry logic
        // (T only decreases, so the group's end value decides).
        bool all_done = false;
#pragma unroll 1
        for (int t0 = 0; t0 < CHUNK; t0 += FSZ) {
            float a[FSZ];
#pragma unroll
            for (int s = 0; s < FSZ; ++s) {
                const float4 kk = rows[2 * (t0 + s)].k;
                const float4 q = rows[2 * (t0 + s)].q;
                const float l2a = log2_alpha(q, kk.x - px, kk.y - py);
                const float alpha = fminf(GSB_ALPHA_CLAMP, ex2_approx(l2a));
                // sigma < 0 <=> log2(alpha) > log2(opacity); alpha < 1/255 <=> log2(alpha) < log2(1/255)
                a[s] = (l2a <= q.w && l2a >= LOG2_ALPHA_MIN) ? alpha : 0.f;
            }
            const float T_saved = T;
            const int last_saved = last_k;
            float acc_saved[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) acc_saved[k] = acc[k];
#pragma unroll
            for (int s = 0; s < FSZ; ++s) {
                const float vis = a[s] * T;
                T = T * (1.0f - a[s]);
                const float4 c = rows[2 * (t0 + s)].c;      // finite by contract: alpha == 0 makes the products vanish
                acc[0] = fmaf(c.x, vis, acc[0]);
                if (CH > 1) acc[1] = fmaf(c.y, vis, acc[1]);
                if (CH > 2) acc[2] = fmaf(c.z, vis, acc[2]);
                if (CH > 3) {
                    if (a[s] > 0.f) {
                        const int g = __float_as_int(rows[2 * (t0 + s)].k.w);
#pragma unroll
                        for (int k = 3; k < CH; ++k) acc[k] += __ldg(colors + (size_t)g * CH + k) * vis;
                    }
                }
                last_k = a[s] > 0.f ? base + t0 + s : last_k;
            }
            if (!__any_sync(0xffffffffu, T <= GSB_T_STOP)) continue;
            // ---- roll back and replay this group entry by entry
            T = T_saved;
            last_k = last_saved;
#pragma unroll
            for (int k = 0; k < CH; ++k) acc[k] = acc_saved[k];
#pragma unroll
            for (int s = 0; s < FSZ; ++s) {
                const float alpha = a[s];
                if (done || alpha == 0.f) continue;
                const float next_T = T * (1.0f - alpha);
                if (next_T <= GSB_T_STOP) {   // this entry is excluded
                    done = true;
                    px = FAR_AWAY;
                    continue;
                }
                const float vis = alpha * T;
                const float4 c = rows[2 * (t0 + s)].c;
                acc[0] += c.x * vis;
                if (CH > 1) acc[1] += c.y * vis;
                if (CH > 2) acc[2] += c.z * vis;
                if (CH > 3) {
                    const int g = __float_as_int(rows[2 * (t0 + s)].k.w);
#pragma unroll
                    for (int k = 3; k < CH; ++k) acc[k] += __ldg(colors + (size_t)g * CH + k) * vis;
                }
                last_k = base + t0 + s;
                T = next_T;
            }
            if (__all_sync(0xffffffffu, done)) { all_done = true; break; }
        }
        if (all_done) break;
    }
    {
        // what the backward has to walk: up to the unit's last contributor; one job per SEG entries of the longer list
        int nb = last_k + 1;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) nb = max(nb, __shfl_xor_sync(0xffffffffu, nb, o));
        const int nb_max = max(nb, __shfl_xor_sync(0xffffffffu, nb, 16));
        if ((lane & (PIX - 1)) == 0) {
            walk[unit] = nb;
            ckpt_base[unit] = cb;
        }
        if (SEGD && cb >= 0)
            ckpt[(size_t)cb * PIX + (lane & (PIX - 1))] =
                make_float4(T, acc[0], CH > 1 ? acc[CH > 1 ? 1 : 0] : 0.f, CH > 2 ? acc[CH > 2 ? 2 : 0] : 0.f);
        if (lane == 0 && nb_max > 0) {
            const int nseg = SEGD ? (nb_max + SEG - 1) / SEG : 1;
            const int jb = atomicAdd(ctrl + CTRL_JOBS, nseg);
            for (int sgm = 0; sgm < nseg; ++sgm) jobs[jb + sgm] = make_int2(tile * PAIRS + slot % PAIRS, sgm);
        }
    }
    if (u.inside) {
        int cur_idx = 0;
        if (last_k >= 0) cur_idx = list[last_k].x;
        const size_t pix = (size_t)u.i * W + u.j;
        alphas[pix] = 1.0f - T;
#pragma unroll
        for (int k = 0; k < CH; ++k)
            render[pix * CH + k] = background ? acc[k] + T * background[k] : acc[k];
        last_ids[pix] = cur_idx;
    }
}

// Backward.  The per-Gaussian gradient is a sum over pixels, the transmittance recurrence runs over Gaussians: the
// kernel does each along the axis where it is register-local and TRANSPOSES through shared memory in between, so no
// cross-lane reduction is left.  Per step, 16 entries of each of the warp's two sub-lists (walked back to front):
//   phase A (lane = (unit, pixel)): evaluate the unit's 16 entries, run the T / suffix-colour recurrence, and store
//                              per (entry, pixel) the two scalars the sums need -- alpha T (the weight of v_out in the
//                              colour gradient) and v_sigma -- as one float2 into a 32 x 16 slab whose columns are
//                              rotated by the row (conflict-free for both access patterns, no padding);
//   phase B (lane = Gaussian): read its row, accumulate the colour gradient and the six moments of v_sigma over the 16
//                              pixels in registers, one atomic per value per (Gaussian, unit) -- only for Gaussians
//                              that touched a pixel.
// The kernel is bound by the L1 / shared-memory data pipe (87 % busy with a float4 {A, T, E, -} slab; ncu, round 2):
// what crosses the transpose is kept to 8 bytes per pair.
constexpr int WPB_B = GSB_WPB_B;   // warps per CTA in the backward (6 KB of shared memory per warp)

// slab[row][column] of float2; the column of pixel p in row r is (p + r) mod 16: the 8-byte accesses of a half-warp then
// cover all 32 banks both when the lanes are pixels of one row (phase A) and when they are rows at one pixel (phase B)
__device__ __forceinline__ int slab_at(int row, int p) { return row * PIX + ((p + row) & (PIX - 1)); }

// The backward's staged record, permuted so that what phase A needs per entry -- centre, folded conic, log2(opacity),
// colour, list position: 10 floats -- is two 16-byte loads and one 8-byte load (a 16-byte shared-memory load costs four
// wavefronts of the L1 data pipe however few distinct addresses it has, and that pipe is this kernel's bound).
struct SRec {
    float4 a;   // x, y, q.x, q.y
    float4 b;   // q.z, q.w = log2(opacity), r, g
    float4 c;   // b, tile-list position, Gaussian id (bit patterns), opacity
};
__device__ __forceinline__ SRec stage_record(const Rec &r, int2 e) {
    SRec s;
    s.a = make_float4(r.k.x, r.k.y, r.q.x, r.q.y);
    s.b = make_float4(r.q.z, r.q.w, r.c.x, r.c.y);
    s.c = make_float4(r.c.z, __int_as_float(e.x), __int_as_float(e.y), r.c.w);
    return s;
}

template <int CH>
__global__ void __launch_bounds__(32 * WPB_B, GSB_MINB_B)
composite_bwd_kernel(int W, int H, int tile_w, const Rec *__restrict__ rec,
                     const float *__restrict__ colors, const float *__restrict__ background,
                     const int32_t *__restrict__ offsets, int n_tiles, int M_host, const int64_t *__restrict__ m_dev,
                     const int2 *__restrict__ entries, const int32_t *__restrict__ walk,
                     const int32_t *__restrict__ ckpt_base, int32_t *ctrl, const int2 *__restrict__ jobs,
                     const float4 *__restrict__ ckpt, const float *__restrict__ alphas,
                     const int32_t *__restrict__ last_ids, const float *__restrict__ v_render,
                     const float *__restrict__ v_alphas, float *__restrict__ v_means2d, float *__restrict__ v_conics,
                     float *__restrict__ v_colors, float *__restrict__ v_opacities) {
    constexpr int PAIRS = SUBS / UPW;
    constexpr int C3 = CH < 3 ? CH : 3;                 // channels carried in the packed record
    constexpr int NV4 = (CH + 3) / 4;                   // float4s of v_out per pixel
    constexpr bool SEGD = segmented<CH>();
    __shared__ float2 s_slab[WPB_B][32 * PIX];          // {alpha T, v_sigma} per (row, pixel)
    __shared__ SRec s_rec[WPB_B][32];
    // v_out of the two units' pixels, side by side per pixel (distinct banks).  Up to three channels: 8 + 4 bytes in two
    // arrays (three wavefronts per load pair; as one float4 the compiler fuses them back into a four-wavefront LDS.128)
    constexpr bool NARROW = CH <= 3;
    __shared__ float4 s_vo[NARROW ? 1 : WPB_B][NARROW ? 1 : PIX][UPW][NV4];
    __shared__ float2 s_vo01[NARROW ? WPB_B : 1][NARROW ? PIX : 1][UPW];
    __shared__ float s_vo2[NARROW ? WPB_B : 1][NARROW ? PIX : 1][UPW];
    static_assert(sizeof(SRec) == 48, "staged record");
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int half = lane >> 4, p = lane & (PIX - 1);
    const int M = m_dev ? (int)*m_dev : M_host;
    const int n_jobs = ctrl[CTRL_JOBS];                 // written by the forward

    // Persistent warps draw jobs (pair of units, segment) from the queue the forward filled; the next ticket is in flight
    // while a job is processed.
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(ctrl + CTRL_CURSOR, 1);
    for (;;) {
    const int job = __shfl_sync(0xffffffffu, ticket, 0);
    if (job >= n_jobs) break;
    if (lane == 0) ticket = atomicAdd(ctrl + CTRL_CURSOR, 1);
    const int2 jb = jobs[job];
    const int tile = jb.x / PAIRS;
    const Unit u = make_unit(jb.x % PAIRS, lane, tile, tile_w, W, H);
    const int unit = tile * SUBS + u.w;
    const size_t pix = u.inside ? (size_t)u.i * W + u.j : 0;

    const int start = offsets[tile];
    const int end = (tile == n_tiles - 1) ? M : offsets[tile + 1];
    // my unit's part of the job: sub-list entries [lo, hi) of the walk_n the forward found worth walking
    const int walk_n = walk[unit];
    const int lo = jb.y * SEG;
    const int hi = SEGD ? min(walk_n, lo + SEG) : walk_n;
    const int n = max(hi - lo, 0);
    const int2 *list = entries + (size_t)SUBS * start + (size_t)u.w * (end - start) + lo;

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
    __syncwarp();   // the previous job's phase B has read s_vo
    {
        float vo4[NV4 * 4];
#pragma unroll
        for (int k = 0; k < NV4 * 4; ++k) vo4[k] = (k < CH) ? v_out[k] : 0.f;
        if (NARROW) {
            s_vo01[wib][p][half] = make_float2(vo4[0], vo4[1]);
            s_vo2[wib][p][half] = vo4[2];
        } else {
#pragma unroll
            for (int k = 0; k < NV4; ++k)
                s_vo[wib][p][half][k] = make_float4(vo4[4 * k], vo4[4 * k + 1], vo4[4 * k + 2], vo4[4 * k + 3]);
        }
    }
    const int n_max = max(n, __shfl_xor_sync(0xffffffffu, n, 16));

    float2 *const slab = s_slab[wib];
    const float bx = (float)(u.j - (p & 3)) + 0.5f, by = (float)(u.i - (p >> 2)) + 0.5f;  // pixel (0,0) of my unit
    const float bx_other = __shfl_xor_sync(0xffffffffu, bx, 16);
    float T = T_final;
    float B = 0.f;    // suffix colour behind the current Gaussian, dotted with v_out
    if (SEGD && hi < walk_n) {
        // not my unit's last segment: T before entry `hi` and the colour accumulated from `hi` on, from the forward's
        // checkpoints (slot 0 = final state, slot j = state before entry j * SEG)
        const size_t cbk = (size_t)ckpt_base[unit];
        const float4 ck = ckpt[(cbk + jb.y + 1) * PIX + p], fin = ckpt[cbk * PIX + p];
        // The reference derives every T of the backward from T_final = 1 - alpha_out (rfstudio's gsplat: rasterize_to_
        // pixels_bwd), i.e. from a value that carries the rounding of 1 - (1 - T): up to 6e-8 / T relative, 1e-4 for a
        // nearly opaque pixel, in EVERY T of that pixel.  The checkpoint holds the forward's own T; scaling by
        // (1 - alpha_out) / T_final reproduces the reference's chain.
        const float ratio = T_final / fin.x;
        T = ck.x * ratio;
        B = (fin.y - ck.y) * v_out[0];
        if (C3 > 1) B += (fin.z - ck.z) * v_out[C3 > 1 ? 1 : 0];
        if (C3 > 2) B += (fin.w - ck.w) * v_out[C3 > 2 ? 2 : 0];
        B *= ratio;
    }
    // walk back to front: per unit, a step covers sub-list indices [top - 16, top); index top - 1 - i goes to
    // shared-memory row 2 * i + unit
    // lane l stages shared-memory row l = entry l / 2 (from the top) of unit l % 2, as in the forward
    const int li = lane >> 1;
    const bool own = ((lane & 1) == half);
    const int2 *const list_other = (const int2 *)__shfl_xor_sync(0xffffffffu, (unsigned long long)list, 16);
    const int n_other = __shfl_xor_sync(0xffffffffu, n, 16);
    const int2 *const slist = own ? list : list_other;
    const int sn = own ? n : n_other;
    const SRec *const rows = &s_rec[wib][half];
    // same software pipeline as the forward: records one step ahead, list entries two steps ahead
    const int2 none = make_int2(0x7fffffff, 0);   // position beyond every last_id: a null row is never valid
    int2 e = none, e_next = none;
    Rec r = null_record();
    if (sn - 1 - li >= 0) { e = slist[sn - 1 - li]; r = rec[e.y]; }
    if (sn - CHUNK - 1 - li >= 0) e_next = slist[sn - CHUNK - 1 - li];
    for (int top = sn, walked = 0; walked < n_max; top -= CHUNK, walked += CHUNK) {
        __syncwarp();
        s_rec[wib][lane] = stage_record(r, e);
        __syncwarp();
        const int nt = top - CHUNK;
        e = e_next;
        e_next = none;
        r = null_record();
        if (nt - 1 - li >= 0) r = rec[e.y];
        if (nt - CHUNK - 1 - li >= 0) e_next = slist[nt - CHUNK - 1 - li];

        // ---- phase A: lane = (unit, pixel) --------------------------------------------------------------------
        unsigned mine_mask = 0u;   // bit t: shared-memory row t (one of my unit's) contributes to my pixel
#pragma unroll 1
        for (int t0 = 0; t0 < CHUNK; t0 += BSZ) {
            // independent per entry: alpha, 1 / (1 - alpha), colour . v_out  (ILP for a warp that runs alone)
            float Ag[BSZ], al[BSZ], rag[BSZ], wg[BSZ];
#pragma unroll
            for (int s = 0; s < BSZ; ++s) {
                const int t = 2 * (t0 + s);
                const float4 ra = rows[t].a, rb = rows[t].b;
                const float2 rc = *reinterpret_cast<const float2 *>(&rows[t].c);      // blue, list position
                const float4 q = make_float4(ra.z, ra.w, rb.x, rb.y);
                const float l2a = log2_alpha(q, ra.x - u.px, ra.y - u.py);
                const float araw = ex2_approx(l2a);
                const bool valid = (__float_as_int(rc.y) <= bin_final) && (l2a <= q.w) && (l2a >= LOG2_ALPHA_MIN);
                Ag[s] = valid ? araw : 0.f;
                al[s] = fminf(GSB_ALPHA_CLAMP, Ag[s]);
                rag[s] = rcp_approx(1.0f - al[s]);
                float w = rb.z * v_out[0];
                if (C3 > 1) w += rb.w * v_out[1];
                if (C3 > 2) w += rc.x * v_out[2];
                if (CH > 3) {
                    const int gid = __float_as_int(rows[t].c.z);
#pragma unroll
                    for (int k = 3; k < CH; ++k) w += __ldg(colors + (size_t)gid * CH + k) * v_out[k];
                }
                wg[s] = w;
                mine_mask |= valid ? (1u << (t + half)) : 0u;
            }
            // the recurrence: alpha == 0 (pair does not contribute) makes every update the identity, so no branch.
            // v_alpha = T (c . v_out) - E with E = (B - c0) / (1 - alpha) is the reference's expression ((c T - buffer /
            // (1 - alpha)) . v_out + T_final / (1 - alpha) (v_alpha_out - bg . v_out)) with the per-pixel constants folded
            // into c0; a clamped alpha passes no gradient to sigma / opacity; d alpha / d sigma = -alpha.
#pragma unroll
            for (int s = 0; s < BSZ; ++s) {
                T *= rag[s];                                  // T before this Gaussian
                const float fac = al[s] * T;
                const float v_alpha = T * wg[s] - rag[s] * (B - c0);
                const float v_sigma = (Ag[s] <= GSB_ALPHA_CLAMP) ? -Ag[s] * v_alpha : 0.f;
                slab[slab_at(2 * (t0 + s) + half, p)] = make_float2(fac, v_sigma);
                B += wg[s] * fac;
            }
        }
        const unsigned touched = __reduce_or_sync(0xffffffffu, mine_mask);   // also orders phase A's stores before phase B
        __syncwarp();
        if (touched == 0u) continue;

        // ---- phase B: lane = Gaussian (shared-memory row `lane`: entry lane / 2 of unit lane % 2) -------------
        {
            const float4 ra = s_rec[wib][lane].a, rb = s_rec[wib][lane].b, rc = s_rec[wib][lane].c;
            const float2 kk = make_float2(ra.x, ra.y);                 // centre
            const float3 q = make_float3(ra.z, ra.w, rb.x);            // folded conic
            const int g = __float_as_int(rc.z);
            const float opac = rc.w;
            const int hb = lane & 1;                                   // the unit this row belongs to
            const float ox = (hb == half) ? bx : bx_other;             // pixel (0,0) of that unit (same row of pixels)
            const bool mine = (touched >> lane) & 1u;
            float g_col[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) g_col[k] = 0.f;
            float s0 = 0.f, sxx = 0.f, sxy = 0.f, syy = 0.f, sx = 0.f, sy = 0.f;   // moments of v_sigma
            const float2 *row = slab + lane * PIX;
#pragma unroll 8
            for (int pp = 0; pp < PIX; ++pp) {
                const float2 fv = row[(pp + lane) & (PIX - 1)];
                const float fac = fv.x, v_sigma = fv.y;
                if (NARROW) {
                    const float2 v01 = s_vo01[wib][pp][hb];
                    g_col[0] += fac * v01.x;
                    if (CH > 1) g_col[CH > 1 ? 1 : 0] += fac * v01.y;
                    if (CH > 2) g_col[CH > 2 ? 2 : 0] += fac * s_vo2[wib][pp][hb];
                } else {
#pragma unroll
                    for (int k = 0; k < NV4; ++k) {
                        const float4 v4 = s_vo[wib][pp][hb][k];
                        g_col[4 * k] += fac * v4.x;
                        if (4 * k + 1 < CH) g_col[4 * k + 1 < CH ? 4 * k + 1 : 0] += fac * v4.y;
                        if (4 * k + 2 < CH) g_col[4 * k + 2 < CH ? 4 * k + 2 : 0] += fac * v4.z;
                        if (4 * k + 3 < CH) g_col[4 * k + 3 < CH ? 4 * k + 3 : 0] += fac * v4.w;
                    }
                }
                const float dx = kk.x - (ox + (float)(pp & 3)), dy = kk.y - (by + (float)(pp >> 2));   // exact pixel centre
                const float t1 = v_sigma * dx, t2 = v_sigma * dy;
                s0 += v_sigma;
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
                // conic back from the folded coefficients: a = qa / (-log2e / 2), b = qb / (-log2e), c = qc / (-log2e / 2)
                const float ca = q.x * (-2.0f / LOG2E), cb = q.y * (-1.0f / LOG2E), cc = q.z * (-2.0f / LOG2E);
                atomicAdd(v_means2d + 2 * (size_t)g, ca * sx + cb * sy);
                atomicAdd(v_means2d + 2 * (size_t)g + 1, cb * sx + cc * sy);
                // d alpha / d opacity = exp(-sigma) = A / opacity, and sum A v_alpha = -s0
                atomicAdd(v_opacities + g, -s0 / opac);
            }
        }
    }
    }   // job loop
}

struct Workspace {
    Rec *rec;
    int32_t *counts, *walk, *ckpt_base;   // per unit: sub-list length, entries the backward walks, first checkpoint block
    int32_t *tile_len, *order;            // per tile: longest sub-list (the forward's LPT key), tile order
    int32_t *ctrl;                        // CTRL_JOBS, CTRL_CKPTS, CTRL_CURSOR
    int2 *jobs;                           // backward work items {tile * 8 + pair, segment}
    float4 *ckpt;                         // checkpoint blocks of 16 float4 {T, r, g, b}
    int2 *entries;
};

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// sum over units of ceil(n / SEG) <= sum n / SEG + #(units with n > SEG) <= 2 * SUBS * M / SEG
size_t ckpt_blocks(int64_t M) { return 2 * (size_t)SUBS * (size_t)M / SEG + 1; }
// sum over pairs of ceil(n_max / SEG) <= #pairs + SUBS * M / SEG
size_t job_capacity(int64_t M, int n_tiles) { return (size_t)n_tiles * (SUBS / UPW) + (size_t)SUBS * (size_t)M / SEG + 1; }

size_t workspace_bytes(int64_t N, int64_t M, int n_tiles) {
    return align256(sizeof(Rec) * (size_t)N) + 3 * align256(sizeof(int32_t) * (size_t)n_tiles * SUBS) +
           2 * align256(sizeof(int32_t) * (size_t)n_tiles) + 256 + align256(sizeof(int2) * job_capacity(M, n_tiles)) +
           align256(sizeof(float4) * PIX * ckpt_blocks(M)) + align256(sizeof(int2) * (size_t)SUBS * (size_t)M) + 256;
}

Workspace carve(void *ws, int64_t N, int64_t M, int n_tiles) {
    char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    Workspace w;
    w.rec = reinterpret_cast<Rec *>(p);
    p += align256(sizeof(Rec) * (size_t)N);
    const size_t ub = align256(sizeof(int32_t) * (size_t)n_tiles * SUBS), tb = align256(sizeof(int32_t) * (size_t)n_tiles);
    w.counts = reinterpret_cast<int32_t *>(p); p += ub;
    w.walk = reinterpret_cast<int32_t *>(p); p += ub;
    w.ckpt_base = reinterpret_cast<int32_t *>(p); p += ub;
    w.tile_len = reinterpret_cast<int32_t *>(p); p += tb;
    w.order = reinterpret_cast<int32_t *>(p); p += tb;
    w.ctrl = reinterpret_cast<int32_t *>(p); p += 256;
    w.jobs = reinterpret_cast<int2 *>(p); p += align256(sizeof(int2) * job_capacity(M, n_tiles));
    w.ckpt = reinterpret_cast<float4 *>(p); p += align256(sizeof(float4) * PIX * ckpt_blocks(M));
    w.entries = reinterpret_cast<int2 *>(p);
    return w;
}

template <int CH>
int launch_fwd(int W, int H, int64_t N, const float *means2d, const float *conics, const float *colors,
               const float *opacities, int opacity_is_logit, const float *comps, const float *background,
               const int32_t *offsets, const int32_t *flatten_ids, int64_t M, const int64_t *m_dev, float *render,
               float *alphas, int32_t *last_ids, void *ws, bool prepacked, cudaStream_t st) {
    int tw = (W + GSB_TILE - 1) / GSB_TILE, th = (H + GSB_TILE - 1) / GSB_TILE;
    int n_tiles = tw * th, n_units = n_tiles * SUBS;
    Workspace w = carve(ws, N, M, n_tiles);
    if (N > 0 && !prepacked)   // the batch driver's shade forward has written the records already
        pack_records_kernel<CH><<<gsb_div_up(N, 256), 256, 0, st>>>((int)N, reinterpret_cast<const float2 *>(means2d),
                                                                    conics, colors, opacities, opacity_is_logit,
                                                                    comps, w.rec);
    build_sublists_kernel<<<n_tiles, BUILD_THREADS, 0, st>>>(tw, n_tiles, (int)M, m_dev, offsets, flatten_ids, w.rec,
                                                             w.entries, w.counts, w.tile_len, w.ctrl);
    lpt_order_kernel<<<1, 1024, 0, st>>>(n_tiles, w.tile_len, w.order);
    // experiment knob: unused dynamic shared memory per CTA caps the forward's residency and leaves registers to the
    // kernels of other views (GSB_FWD_PAD_KB, read once)
    static const int pad_kb = getenv("GSB_FWD_PAD_KB") ? atoi(getenv("GSB_FWD_PAD_KB")) : 0;
    if (pad_kb > 0) cudaFuncSetAttribute(composite_fwd_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad_kb * 1024);
    composite_fwd_kernel<CH><<<gsb_div_up(n_units / UPW, WPB), 32 * WPB, (size_t)pad_kb * 1024, st>>>(
        W, H, tw, n_units / UPW, w.rec, colors, background, offsets, n_tiles, (int)M, m_dev, w.entries, w.counts, w.order,
        w.walk, w.ckpt_base, w.ctrl, w.jobs, w.ckpt, render, alphas, last_ids);
    return 0;
}

// Resident CTAs per SM the persistent backward asks for (40 KB of shared memory each); surplus CTAs find the queue
// empty and leave.
#ifndef GSB_BWD_CTAS_PER_SM
#define GSB_BWD_CTAS_PER_SM 5
#endif

template <int CH>
int launch_bwd(int W, int H, int64_t N, const float *colors, const float *background, const int32_t *offsets,
               int64_t M, const int64_t *m_dev, const float *alphas, const int32_t *last_ids, const float *v_render,
               const float *v_alphas, float *v_means2d, float *v_conics, float *v_colors, float *v_opacities, void *ws,
               cudaStream_t st) {
    int tw = (W + GSB_TILE - 1) / GSB_TILE, th = (H + GSB_TILE - 1) / GSB_TILE;
    int n_tiles = tw * th;
    Workspace w = carve(ws, N, M, n_tiles);
    static const int sms = gsb_sm_count();
    const int64_t want = (int64_t)gsb_div_up((int64_t)job_capacity(M, n_tiles), WPB_B);
    static const int per_sm = getenv("GSB_BWD_CTAS") ? atoi(getenv("GSB_BWD_CTAS")) : GSB_BWD_CTAS_PER_SM;   // experiment knob
    const int grid = (int)(want < (int64_t)sms * per_sm ? want : (int64_t)sms * per_sm);
    if (cudaMemsetAsync(w.ctrl + CTRL_CURSOR, 0, sizeof(int32_t), st) != cudaSuccess) return 1;   // a forward may be walked twice
    composite_bwd_kernel<CH><<<grid, 32 * WPB_B, 0, st>>>(
        W, H, tw, w.rec, colors, background, offsets, n_tiles, (int)M, m_dev, w.entries, w.walk, w.ckpt_base, w.ctrl,
        w.jobs, w.ckpt, alphas, last_ids, v_render, v_alphas, v_means2d, v_conics, v_colors, v_opacities);
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

// Where the records live inside a compositing workspace (its first block): for a producer that packs them itself.
void *gsb_composite_records(void *workspace) {
    return reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
}

// `M` is the CAPACITY of the tile lists when `m_dev` (the device-resident intersection count) is given, else the count.
// `prepacked`: the records of all N Gaussians are already at gsb_composite_records(workspace) (colours included), and
// means2d / conics / opacities / comps are not read.
int gsb_composite_fwd_impl(int32_t width, int32_t height, int32_t channels, int64_t N, const float *means2d,
                           const float *conics, const float *colors, const float *opacities, int32_t opacity_is_logit,
                           const float *comps, const float *background, const int32_t *offsets,
                           const int32_t *flatten_ids, int64_t M, const int64_t *m_dev, float *render, float *alphas,
                           int32_t *last_ids, void *workspace, size_t workspace_bytes_, int32_t prepacked, void *stream) {
    GSB_CHECK_ARG(width > 0 && height > 0 && N >= 0 && M >= 0 && M < 134217727LL);
    GSB_CHECK_ARG(offsets && render && alphas && last_ids && workspace);
    GSB_CHECK_ARG(M == 0 || (means2d && conics && colors && opacities && flatten_ids));
    int tw = (width + GSB_TILE - 1) / GSB_TILE, th = (height + GSB_TILE - 1) / GSB_TILE;
    if (workspace_bytes(N, M, tw * th) > workspace_bytes_) {
        gsb_set_error("gsb_composite_fwd: workspace too small (%zu < %zu)", workspace_bytes_, workspace_bytes(N, M, tw * th));
        return GSB_ENOMEM;
    }
    int rc = 0;
    GSB_DISPATCH_CH(channels, (rc = launch_fwd<C_>(width, height, N, means2d, conics, colors, opacities,
                                                    opacity_is_logit, comps, background, offsets, flatten_ids, M, m_dev,
                                                    render, alphas, last_ids, workspace, prepacked != 0,
                                                    (cudaStream_t)stream)));
    if (rc != 0) {
        gsb_set_error("gsb_composite_fwd: internal sort scratch too small");
        return GSB_ENOMEM;
    }
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

int gsb_composite_bwd_impl(int32_t width, int32_t height, int32_t channels, int64_t N, const float *colors,
                           const float *background, const int32_t *offsets, int64_t M, const int64_t *m_dev,
                           const float *alphas, const int32_t *last_ids, const float *v_render, const float *v_alphas,
                           float *v_means2d, float *v_conics, float *v_colors, float *v_opacities,
                           const void *workspace, void *stream) {
    GSB_CHECK_ARG(width > 0 && height > 0 && N >= 0 && M >= 0 && M < 134217727LL);
    GSB_CHECK_ARG(offsets && alphas && last_ids && v_render && v_alphas && workspace);
    if (M == 0) return GSB_OK;
    GSB_CHECK_ARG(colors && v_means2d && v_conics && v_colors && v_opacities);
    GSB_DISPATCH_CH(channels, (launch_bwd<C_>(width, height, N, colors, background, offsets, M, m_dev, alphas, last_ids,
                                               v_render, v_alphas, v_means2d, v_conics, v_colors, v_opacities,
                                               const_cast<void *>(workspace), (cudaStream_t)stream)));
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}

GSB_API int gsb_composite_fwd(int32_t width, int32_t height, int32_t channels, int64_t N, const float *means2d,
                              const float *conics, const float *colors, const float *opacities,
                              int32_t opacity_is_logit, const float *comps, const float *background,
                              const int32_t *offsets, const int32_t *flatten_ids, int64_t M, float *render,
                              float *alphas, int32_t *last_ids, void *workspace, size_t workspace_bytes_,
                              void *stream) {
    return gsb_composite_fwd_impl(width, height, channels, N, means2d, conics, colors, opacities, opacity_is_logit, comps,
                                  background, offsets, flatten_ids, M, nullptr, render, alphas, last_ids, workspace,
                                  workspace_bytes_, 0, stream);
}

GSB_API int gsb_composite_bwd(int32_t width, int32_t height, int32_t channels, int64_t N, const float *colors,
                              const float *background, const int32_t *offsets, int64_t M, const float *alphas,
                              const int32_t *last_ids, const float *v_render, const float *v_alphas,
                              float *v_means2d, float *v_conics, float *v_colors, float *v_opacities,
                              const void *workspace, void *stream) {
    return gsb_composite_bwd_impl(width, height, channels, N, colors, background, offsets, M, nullptr, alphas, last_ids,
                                  v_render, v_alphas, v_means2d, v_conics, v_colors, v_opacities, workspace, stream);
}
