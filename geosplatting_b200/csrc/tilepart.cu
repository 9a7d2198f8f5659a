// Second stage of the two-stage binning for the batch driver: a STABLE partition of the emitted (tile, Gaussian) pairs
// by tile in one pass over the pairs, instead of a radix sort.
//
// What it replaces: gsplat sorts (tile | depth) keys of all intersections (isect_tiles + radix sort, SURVEY.md
// Appendix C); this library sorts the Gaussians by depth first and emits the pairs in that order (binsort.cu), after
// which a stable sort by the 12-bit tile id finishes the job -- two cub onesweep passes over a CAPACITY of pairs plus a
// histogram, two fills and an offsets kernel (0.115 ms per view, the third largest item of a view after the two
// compositing kernels).  But a stable sort on a key with n_tiles distinct values whose pairs are already in the wanted
// relative order is a stable PARTITION, and the per-tile offsets it needs are the output `isect_offsets` anyway:
//   1. tile_count   : per chunk of 16 384 pairs, a histogram over the tiles (shared-memory atomics) -> H[chunk][tile];
//   2. tile_prefix  : per tile, the exclusive prefix of H over the chunks and the tile's total; tile_offsets: the exclusive
//                     scan of the totals over the tiles = isect_offsets (one small CTA);
//   3. tile_scatter : per chunk, every pair's rank among the pairs of its tile inside the chunk -- warp w owns the w-th
//                     slice of 1 024 pairs and walks it 32 pairs at a time: the lanes that hold the same tile find each
//                     other with 12 ballots, the first of them bumps the warp's private counter of that tile, a prefix
//                     over the 16 warps' counters orders the slices -- and gid goes to
//                     offsets[tile] + H[chunk][tile] + prefix[warp][tile] + rank.
// Only the m_eff pairs that exist are touched (the count stays on the device): no sentinel fill, no capacity-sized
// passes.  The scatter kernel's shape follows the image: 16 / 8 / 4 warps per CTA for up to 4096 / 10 240 / 16 384
// tiles (W warps x n_tiles 16-bit counters + n_tiles bases in shared memory, at most 205 KB); larger images keep the
// radix sort.
#include "gsb_common.cuh"

int gsb_isect_tiles_ordered_cap(int32_t N, const float *means2d, const int32_t *radii, const int32_t *order,
                                const int64_t *cum_ordered, const gsb_camera *cam, int64_t cap, uint32_t *tile_keys,
                                int32_t *gauss_ids, void *stream);

namespace {

constexpr int TP_SLICE = 1024;                  // consecutive pairs owned by one warp of the scatter kernel
constexpr int TP_ROUNDS = TP_SLICE / 32;
constexpr int TP_THREADS = 512;                 // histogram / offsets kernels
constexpr int TP_MAX_TILES = 16384;
// Shape of the scatter kernel by image size: W warps per CTA, a chunk = W slices, (4 + 2 W) bytes of shared memory per tile.
//   n_tiles <= 4096 : 16 warps, 16 384-pair chunks, 12 key bits (147 KB at 4096 tiles)
//   n_tiles <= 10240:  8 warps,  8 192-pair chunks, 14 key bits (205 KB at 10 240 tiles; 1600 x 1600 has 10 000)
//   n_tiles <= 16384:  4 warps,  4 096-pair chunks, 14 key bits (197 KB)
struct TpShape {
    int warps, bits;
    int chunk() const { return warps * TP_SLICE; }
};
TpShape tp_shape(int n_tiles) {
    if (n_tiles <= 4096) return {16, 12};
    if (n_tiles <= 10240) return {8, 14};
    return {4, 14};
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__global__ void __launch_bounds__(TP_THREADS) tile_count_kernel(const int64_t *__restrict__ m_eff,
                                                                 const uint32_t *__restrict__ keys, int n_tiles, int chunk,
                                                                 uint32_t *__restrict__ H) {
    extern __shared__ uint32_t tp_smem[];
    const int64_t M = *m_eff, c0 = (int64_t)blockIdx.x * chunk;
    if (c0 >= M) return;
    for (int t = threadIdx.x; t < n_tiles; t += TP_THREADS) tp_smem[t] = 0u;
    __syncthreads();
    // eight independent loads in flight per thread, then their eight shared-memory atomics
    for (int i0 = threadIdx.x; i0 < chunk; i0 += 8 * TP_THREADS) {
        uint32_t k[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t idx = c0 + i0 + j * TP_THREADS;
            k[j] = idx < M ? keys[idx] : 0xffffffffu;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (k[j] != 0xffffffffu) atomicAdd(&tp_smem[k[j]], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += TP_THREADS) H[(size_t)blockIdx.x * n_tiles + t] = tp_smem[t];
}

// Per tile: the exclusive prefix of the tile's counts over the chunks, and its total.  Thread (x, y) of a 64 x 8 CTA
// takes tile x of the CTA's 64 and the y-th eighth of the chunks: partial totals first, then the prefixes on top of the
// eighths before it (two coalesced sweeps over H, 8 x the threads of one thread per tile).  Separate input and output
// arrays keep the loads of the unrolled loops independent of the stores.
constexpr int TPX = 64, TPY = 8;
__global__ void __launch_bounds__(TPX * TPY) tile_prefix_kernel(const int64_t *__restrict__ m_eff, int n_tiles, int chunk,
                                                                 const uint32_t *__restrict__ H, uint32_t *__restrict__ Hx,
                                                                 uint32_t *__restrict__ totals) {
    __shared__ uint32_t part[TPY][TPX];
    const int x = threadIdx.x % TPX, y = threadIdx.x / TPX;
    const int t = blockIdx.x * TPX + x;
    const int64_t M = *m_eff;
    const int n_chunks = (int)((M + chunk - 1) / chunk);
    const int per = (n_chunks + TPY - 1) / TPY, c_lo = min(y * per, n_chunks), c_hi = min(c_lo + per, n_chunks);
    uint32_t sum = 0u;
    if (t < n_tiles) {
#pragma unroll 8
        for (int c = c_lo; c < c_hi; ++c) sum += H[(size_t)c * n_tiles + t];
    }
    part[y][x] = sum;
    __syncthreads();
    uint32_t run = 0u;
    for (int k = 0; k < y; ++k) run += part[k][x];
    if (t < n_tiles) {
#pragma unroll 8
        for (int c = c_lo; c < c_hi; ++c) {
            const uint32_t v = H[(size_t)c * n_tiles + t];
            Hx[(size_t)c * n_tiles + t] = run;
            run += v;
        }
        if (y == TPY - 1) totals[t] = run;
    }
}

// isect_offsets = exclusive scan of the tiles' totals: one CTA, 4096 tiles at a time (thread tid scans tiles
// 8 tid .. 8 tid + 7 of the group from a coalesced copy in shared memory), a carry between the groups.
__global__ void __launch_bounds__(TP_THREADS) tile_offsets_kernel(int n_tiles, const uint32_t *__restrict__ totals,
                                                                   int32_t *__restrict__ offsets) {
    constexpr int GROUP = 4096, PER = GROUP / TP_THREADS, NW = TP_THREADS / 32;
    __shared__ uint32_t stage[GROUP];
    __shared__ uint32_t warp_tot[NW];
    __shared__ uint32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0u;
    for (int g0 = 0; g0 < n_tiles; g0 += GROUP) {
        const int n = min(GROUP, n_tiles - g0);
        __syncthreads();                      // carry written, stage free
        for (int t = tid; t < n; t += TP_THREADS) stage[t] = totals[g0 + t];
        __syncthreads();
        uint32_t v[PER], sum = 0u;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int t = tid * PER + k;
            v[k] = t < n ? stage[t] : 0u;
            sum += v[k];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        uint32_t before = carry + incl - sum;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int t = tid * PER + k;
            if (t < n) stage[t] = before;
            before += v[k];
        }
        __syncthreads();                      // every thread has read the carry
        if (tid == TP_THREADS - 1) carry = before;
        for (int t = tid; t < n; t += TP_THREADS) offsets[g0 + t] = (int32_t)stage[t];
    }
}

template <int WARPS, int BITS>
__global__ void __launch_bounds__(32 * WARPS) tile_scatter_kernel(const int64_t *__restrict__ m_eff,
                                                                   const uint32_t *__restrict__ keys,
                                                                   const int32_t *__restrict__ gids, int n_tiles,
                                                                   const uint32_t *__restrict__ Hx,
                                                                   const int32_t *__restrict__ offsets,
                                                                   int32_t *__restrict__ flatten_ids) {
    extern __shared__ uint32_t tp_smem[];
    uint32_t *const base32 = tp_smem;                                                // [n_tiles]
    uint16_t *const cnt = reinterpret_cast<uint16_t *>(tp_smem + n_tiles);          // [WARPS][n_tiles]
    constexpr int THREADS = 32 * WARPS;
    const int64_t M = *m_eff, c0 = (int64_t)blockIdx.x * (WARPS * TP_SLICE);
    if (c0 >= M) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < WARPS * n_tiles / 2; i += THREADS)                // 32-bit stores (WARPS is even)
        reinterpret_cast<uint32_t *>(cnt)[i] = 0u;
    for (int t = tid; t < n_tiles; t += THREADS)
        base32[t] = (uint32_t)offsets[t] + Hx[(size_t)blockIdx.x * n_tiles + t];
    __syncthreads();
    uint16_t *const mine = cnt + (size_t)warp * n_tiles;
    const int64_t s0 = c0 + (int64_t)warp * TP_SLICE;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t rank[TP_ROUNDS];
    uint32_t kbuf[16];                         // the keys of 16 rounds at a time: loads in flight instead of one per round
#pragma unroll
    for (int r = 0; r < TP_ROUNDS; ++r) {
        if ((r & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int64_t ix = s0 + (r + j) * 32 + lane;
                kbuf[j] = ix < M ? keys[ix] : 0u;
            }
        }
        const int64_t idx = s0 + r * 32 + lane;
        const bool valid = idx < M;
        const uint32_t key = kbuf[r & 15];
        // the lanes that hold my tile (match.any written with ballots: the host build of this file has no match)
        unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int b = 0; b < BITS; ++b) {
            const unsigned bal = __ballot_sync(0xffffffffu, (key >> b) & 1u);
            peers &= ((key >> b) & 1u) ? bal : ~bal;
        }
        const int leader = valid ? __ffs((int)peers) - 1 : lane;
        uint32_t old = 0u;
        if (valid && lane == leader) {
            old = mine[key];
            mine[key] = (uint16_t)(old + __popc(peers));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + __popc(peers & lt);
        __syncwarp();                          // the next round reads the counters this one wrote
    }
    __syncthreads();
    // exclusive prefix of the 16 warps' counts per tile (fits 16 bits: a chunk has 16 384 pairs)
    for (int t = tid; t < n_tiles; t += THREADS) {
        uint32_t run = 0u;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t v = cnt[(size_t)w * n_tiles + t];
            cnt[(size_t)w * n_tiles + t] = (uint16_t)run;
            run += v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TP_ROUNDS; ++r) {
        const int64_t idx = s0 + r * 32 + lane;
        if (idx < M) {
            const uint32_t key = keys[idx];
            flatten_ids[base32[key] + mine[key] + rank[r]] = gids[idx];
        }
    }
}

}  // namespace

// (C linkage so that the host-build tests can call them; hidden in the product library, which exports the header only)
// Bytes of scratch gsb_tile_partition_cap needs for `cap` pairs: keys | gids | H | Hx | totals.
extern "C" size_t gsb_tile_partition_bytes(int64_t cap, int n_tiles) {
    const int chunk = tp_shape(n_tiles).chunk();
    const size_t n_chunks = (size_t)((cap + chunk - 1) / chunk);
    return 2 * align256(sizeof(uint32_t) * (size_t)cap) + 2 * align256(sizeof(uint32_t) * n_chunks * (size_t)n_tiles) +
           align256(sizeof(uint32_t) * (size_t)n_tiles) + 256;
}

// 1 when the batch driver should take the partition path for an image of n_tiles tiles.  The kernels work up to 16 384
// tiles, but the chunk-by-tile histogram matrix grows with both: at 1600 x 1600 with 5 M Gaussians (10 000 tiles, 12 M
// pairs in 1 500 chunks: 2 x 60 MB of H) the partition LOSES to the radix sort (281 against 299 views/s); at 800 x 800 it
// wins (2 500 tiles: +3.5 % at 1 M Gaussians, +3 % at 2 M).
extern "C" int gsb_tile_partition_supported(int n_tiles) { return n_tiles >= 1 && n_tiles <= 4096; }

namespace {
template <int WARPS, int BITS>
int launch_scatter(int n_chunks, const int64_t *m_eff, const uint32_t *keys, const int32_t *gids, int n_tiles,
                   const uint32_t *Hx, const int32_t *offsets, int32_t *flatten_ids, cudaStream_t st) {
    const size_t smem = sizeof(uint32_t) * (size_t)n_tiles + sizeof(uint16_t) * (size_t)WARPS * n_tiles;
    GSB_CHECK_CUDA(cudaFuncSetAttribute(tile_scatter_kernel<WARPS, BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tile_scatter_kernel<WARPS, BITS><<<n_chunks, 32 * WARPS, smem, st>>>(m_eff, keys, gids, n_tiles, Hx, offsets, flatten_ids);
    return GSB_OK;
}
}  // namespace

// Same contract as gsb_bin2_sort_cap (binsort.cu): emission in depth order (guarded by the capacity), then
// flatten_ids = the pairs' Gaussian ids grouped by tile in emission order, offsets = first pair of every tile.
extern "C" int gsb_tile_partition_cap(int32_t N, int64_t cap, const int64_t *m_eff, const float *means2d, const int32_t *radii,
                           const int32_t *order, const int64_t *cum_ordered, const gsb_camera *cam, int32_t *flatten_ids,
                           int32_t *offsets, void *workspace, size_t workspace_bytes, void *stream) {
    GSB_CHECK_ARG(N >= 0 && cap >= 0 && cam != nullptr && offsets != nullptr && m_eff != nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    const int tile_w = (cam->width + GSB_TILE - 1) / GSB_TILE, tile_h = (cam->height + GSB_TILE - 1) / GSB_TILE;
    const int n_tiles = tile_w * tile_h;
    GSB_CHECK_ARG(n_tiles >= 1 && n_tiles <= TP_MAX_TILES);
    GSB_CHECK_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int32_t) * (size_t)n_tiles, st));   // stays when N == 0
    if (cap == 0 || N == 0) return GSB_OK;
    GSB_CHECK_ARG(means2d && radii && order && cum_ordered && flatten_ids && workspace);
    if (gsb_tile_partition_bytes(cap, n_tiles) > workspace_bytes) {
        gsb_set_error("gsb_tile_partition_cap: workspace too small (%zu < %zu)", workspace_bytes,
                      gsb_tile_partition_bytes(cap, n_tiles));
        return GSB_ENOMEM;
    }
    const TpShape shape = tp_shape(n_tiles);
    const int chunk = shape.chunk(), n_chunks = (int)((cap + chunk - 1) / chunk);
    char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    uint32_t *keys = reinterpret_cast<uint32_t *>(p); p += align256(sizeof(uint32_t) * (size_t)cap);
    int32_t *gids = reinterpret_cast<int32_t *>(p); p += align256(sizeof(uint32_t) * (size_t)cap);
    const size_t hb = align256(sizeof(uint32_t) * (size_t)n_chunks * (size_t)n_tiles);
    uint32_t *H = reinterpret_cast<uint32_t *>(p); p += hb;
    uint32_t *Hx = reinterpret_cast<uint32_t *>(p); p += hb;
    uint32_t *totals = reinterpret_cast<uint32_t *>(p);
    int rc = gsb_isect_tiles_ordered_cap(N, means2d, radii, order, cum_ordered, cam, cap, keys, gids, stream);
    if (rc != GSB_OK) return rc;
    const size_t smem_count = sizeof(uint32_t) * (size_t)n_tiles;
    if (smem_count > 48 * 1024)
        GSB_CHECK_CUDA(cudaFuncSetAttribute(tile_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count));
    tile_count_kernel<<<n_chunks, TP_THREADS, smem_count, st>>>(m_eff, keys, n_tiles, chunk, H);
    tile_prefix_kernel<<<gsb_div_up(n_tiles, TPX), TPX * TPY, 0, st>>>(m_eff, n_tiles, chunk, H, Hx, totals);
    tile_offsets_kernel<<<1, TP_THREADS, 0, st>>>(n_tiles, totals, offsets);
    if (shape.warps == 16) rc = launch_scatter<16, 12>(n_chunks, m_eff, keys, gids, n_tiles, Hx, offsets, flatten_ids, st);
    else if (shape.warps == 8) rc = launch_scatter<8, 14>(n_chunks, m_eff, keys, gids, n_tiles, Hx, offsets, flatten_ids, st);
    else rc = launch_scatter<4, 14>(n_chunks, m_eff, keys, gids, n_tiles, Hx, offsets, flatten_ids, st);
    if (rc != GSB_OK) return rc;
    GSB_CHECK_LAUNCH();
    return GSB_OK;
}
