// Native per-view driver: the three phases of one training view (prepare / finish / backward) as ONE C-ABI call each,
// composed from the stage entry points of this library on caller-provided arenas.
//
// The per-view host work of the reference is a Python loop over ~40 torch / plugin calls (rfstudio/model/geosplat.py:
// 53-132 + rfstudio/model/gsplat.py:284-358).  At ~1 ms of device time per view that orchestration is what a B200 ends
// up waiting for (measured: ~0.9 ms of host time per view even with one autograd node per batch, bench.py `batches`),
// so the sequencing lives here and the Python host makes three calls and a handful of allocations per view.
//
// Arena layouts (every block 256-byte aligned; sizes from gsb_view_bytes):
//   keep1 : radii[N] i32 | colors[N,3] | m_eff i64 (batch driver)                    prepare -> backward
//   tmp1  : means2d[N,2] | depths[N] | conics[N,3] | comps[N] | tiles_per_gauss[N] | order[N] | cum[N] i64 | scratch
//                                                                                   prepare -> finish
//   keep2 : offsets[T] i32 | render[P,3] | alphas[P] | last_ids[P] i32 | composite workspace    finish -> backward
//   tmp2  : flatten_ids[M] i32 | sort scratch                                       finish only
//   tmp3  : v_render[P,3] | v_alphas[P] | v_means2d[N,2] v_conics[N,3] v_colors[N,3] v_opacities[N] | shade replicas
//                                                                                   backward only
#include <stdlib.h>

#include "gsb_common.cuh"

// internal entry points with a device-resident intersection count (composite.cu, binsort.cu)
int gsb_composite_fwd_impl(int32_t width, int32_t height, int32_t channels, int64_t N, const float *means2d,
                           const float *conics, const float *colors, const float *opacities, int32_t opacity_is_logit,
                           const float *comps, const float *background, const int32_t *offsets,
                           const int32_t *flatten_ids, int64_t M, const int64_t *m_dev, float *render, float *alphas,
                           int32_t *last_ids, void *workspace, size_t workspace_bytes_, int32_t prepacked, void *stream);
void *gsb_composite_records(void *workspace);
extern "C" int gsb_tile_partition_supported(int n_tiles);
extern "C" int gsb_tile_partition_cap(int32_t N, int64_t cap, const int64_t *m_eff, const float *means2d, const int32_t *radii,
                           const int32_t *order, const int64_t *cum_ordered, const gsb_camera *cam, int32_t *flatten_ids,
                           int32_t *offsets, void *workspace, size_t workspace_bytes, void *stream);
int gsb_shade_fwd_impl(int32_t N, const float *means, const float *normals, const float *kd, const float *ks,
                       const float *cam_pos_host, const float *fg_lut, int32_t lut_res, const float *env_stack, int32_t R0,
                       int32_t L, int32_t Rb, float min_roughness, float max_metallic, float env_min_roughness,
                       float env_max_roughness, int32_t mode, float *colors, const float *means2d, const float *conics,
                       const float *opacity_logits, const float *comps, void *rec, void *stream);
int gsb_composite_bwd_impl(int32_t width, int32_t height, int32_t channels, int64_t N, const float *colors,
                           const float *background, const int32_t *offsets, int64_t M, const int64_t *m_dev,
                           const float *alphas, const int32_t *last_ids, const float *v_render, const float *v_alphas,
                           float *v_means2d, float *v_conics, float *v_colors, float *v_opacities,
                           const void *workspace, void *stream);
int gsb_bin2_order(int32_t N, const float *depths, const int32_t *tiles_per_gauss, int32_t *order, int64_t *cum_ordered,
                   int64_t *total_out, void *workspace, size_t workspace_bytes, void *stream);
int gsb_bin2_publish(int32_t N, const int64_t *cum_ordered, int64_t cap, int64_t *m_eff, int64_t *total_out, void *stream);
int gsb_bin2_sort_cap(int32_t N, int64_t cap, const int64_t *m_eff, const float *means2d, const int32_t *radii,
                      const int32_t *order, const int64_t *cum_ordered, const gsb_camera *cam, int32_t *flatten_ids,
                      int32_t *offsets, void *workspace, size_t workspace_bytes, void *stream);

namespace {

size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

struct Carver {
    char *p;
    explicit Carver(void *base) : p(reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(base) + 255) & ~(uintptr_t)255)) {}
    template <class T>
    T *take(size_t count) {
        T *r = reinterpret_cast<T *>(p);
        p += al(sizeof(T) * count);
        return r;
    }
};

struct Sizes {
    size_t keep1, tmp1, keep2, tmp2, tmp3, bin_n, bin_m, comp_ws, shade_ws;
};

int sizes(const gsb_view_config *c, int64_t M, Sizes &s) {
    const size_t N = (size_t)c->N, P = (size_t)c->width * c->height;
    const size_t T = (size_t)((c->width + GSB_TILE - 1) / GSB_TILE) * ((c->height + GSB_TILE - 1) / GSB_TILE);
    int rc;
    if ((rc = gsb_bin2_workspace_bytes(c->N, 0, &s.bin_n)) != GSB_OK) return rc;
    if ((rc = gsb_bin2_workspace_bytes(0, M, &s.bin_m)) != GSB_OK) return rc;
    if ((rc = gsb_composite_workspace_bytes(c->N, M, c->width, c->height, &s.comp_ws)) != GSB_OK) return rc;
    if ((rc = gsb_shade_workspace_bytes(c->R0, c->L, c->Rb, &s.shade_ws)) != GSB_OK) return rc;
    s.keep1 = al(4 * N) + al(12 * N) + al(16) + 256;
    s.tmp1 = al(8 * N) + al(4 * N) + al(12 * N) + al(4 * N) + al(4 * N) + al(4 * N) + al(8 * N) + al(s.bin_n) + 256;
    s.keep2 = al(4 * T) + al(12 * P) + al(4 * P) + al(4 * P) + al(s.comp_ws) + 256;
    s.tmp2 = al(4 * (size_t)M) + al(s.bin_m) + 256;
    s.tmp3 = al(12 * P) + al(4 * P) + al(36 * N) + al(s.shade_ws) + 256;
    return GSB_OK;
}

struct Keep1 { int32_t *radii; float *colors; int64_t *m_eff; };
struct Tmp1 { float *means2d, *depths, *conics, *comps; int32_t *tpg, *order; int64_t *cum; void *scratch; };
struct Keep2 { int32_t *offsets; float *render, *alphas; int32_t *last_ids; void *comp_ws; };

Keep1 carve_keep1(void *a, size_t N) {
    Carver c(a);
    Keep1 k;
    k.radii = c.take<int32_t>(N);
    k.colors = c.take<float>(3 * N);
    k.m_eff = c.take<int64_t>(2);      // min(M, capacity) on the device (batch driver); unused by the exact-size calls
    return k;
}

Tmp1 carve_tmp1(void *a, size_t N, size_t scratch_bytes) {
    Carver c(a);
    Tmp1 t;
    t.means2d = c.take<float>(2 * N);
    t.depths = c.take<float>(N);
    t.conics = c.take<float>(3 * N);
    t.comps = c.take<float>(N);
    t.tpg = c.take<int32_t>(N);
    t.order = c.take<int32_t>(N);
    t.cum = c.take<int64_t>(N);
    t.scratch = c.take<char>(scratch_bytes);
    return t;
}

Keep2 carve_keep2(void *a, size_t T, size_t P, size_t comp_ws) {
    Carver c(a);
    Keep2 k;
    k.offsets = c.take<int32_t>(T);
    k.render = c.take<float>(3 * P);
    k.alphas = c.take<float>(P);
    k.last_ids = c.take<int32_t>(P);
    k.comp_ws = c.take<char>(comp_ws);
    return k;
}

#define VIEW_CHECK_CFG(c)                                                                              \
    GSB_CHECK_ARG((c) != nullptr && (c)->N >= 0 && (c)->width > 0 && (c)->height > 0 && (c)->mode >= 0 && \
                  (c)->mode <= 2 && (c)->lut_res > 1 && (c)->L >= 2 && (c)->R0 > 0 && (c)->Rb > 0)

#define VIEW_TRY(expr)            \
    do {                          \
        int _rc = (expr);         \
        if (_rc != GSB_OK) return _rc; \
    } while (0)

}  // namespace

#define GSB_API extern "C" __attribute__((visibility("default")))

GSB_API int gsb_view_bytes(const gsb_view_config *cfg, int64_t M, size_t *bytes5_host) {
    VIEW_CHECK_CFG(cfg);
    GSB_CHECK_ARG(M >= 0 && bytes5_host != nullptr);
    Sizes s;
    VIEW_TRY(sizes(cfg, M, s));
    bytes5_host[0] = s.keep1; bytes5_host[1] = s.tmp1; bytes5_host[2] = s.keep2; bytes5_host[3] = s.tmp2;
    bytes5_host[4] = s.tmp3;
    return GSB_OK;
}

GSB_API int gsb_view_prepare(const gsb_view_config *cfg, const gsb_camera *cam, const float *cam_pos_host,
                             const float *means, const float *quats, const float *scales, const float *normals,
                             const float *kd, const float *ks, const float *fg_lut, const float *env_stack,
                             void *keep1, void *tmp1, int64_t *total_out, void *stream) {
    VIEW_CHECK_CFG(cfg);
    GSB_CHECK_ARG(cam && cam_pos_host && keep1 && tmp1 && total_out);
    GSB_CHECK_ARG(cam->width == cfg->width && cam->height == cfg->height);
    Sizes s;
    VIEW_TRY(sizes(cfg, 0, s));
    const size_t N = (size_t)cfg->N;
    Keep1 k = carve_keep1(keep1, N);
    Tmp1 t = carve_tmp1(tmp1, N, s.bin_n);
    VIEW_TRY(gsb_project_fwd(cfg->N, means, quats, scales, cam, k.radii, t.means2d, t.depths, t.conics, t.comps, t.tpg,
                             stream));
    VIEW_TRY(gsb_bin2_count(cfg->N, t.depths, t.tpg, t.order, t.cum, total_out, t.scratch, s.bin_n, stream));
    VIEW_TRY(gsb_shade_fwd(cfg->N, means, normals, kd, ks, cam_pos_host, fg_lut, cfg->lut_res, env_stack, cfg->R0, cfg->L,
                           cfg->Rb, cfg->min_roughness, cfg->max_metallic, cfg->env_min_roughness,
                           cfg->env_max_roughness, cfg->mode, k.colors, stream));
    return GSB_OK;
}

GSB_API int gsb_view_finish(const gsb_view_config *cfg, const gsb_camera *cam, int64_t M, const float *opacity_logits,
                            const float *exposure, void *keep1, void *tmp1, void *keep2, void *tmp2, float *out,
                            void *stream) {
    VIEW_CHECK_CFG(cfg);
    GSB_CHECK_ARG(cam && M >= 0 && exposure && keep1 && tmp1 && keep2 && tmp2 && out);
    Sizes s;
    VIEW_TRY(sizes(cfg, M, s));
    const size_t N = (size_t)cfg->N, P = (size_t)cfg->width * cfg->height;
    const size_t T = (size_t)((cfg->width + GSB_TILE - 1) / GSB_TILE) * ((cfg->height + GSB_TILE - 1) / GSB_TILE);
    Keep1 k1 = carve_keep1(keep1, N);
    Tmp1 t1 = carve_tmp1(tmp1, N, s.bin_n);
    Keep2 k2 = carve_keep2(keep2, T, P, s.comp_ws);
    Carver c2(tmp2);
    int32_t *flatten_ids = c2.take<int32_t>((size_t)M);
    void *sort_scratch = c2.take<char>(s.bin_m);
    VIEW_TRY(gsb_bin2_sort(cfg->N, M, t1.means2d, k1.radii, t1.order, t1.cum, cam, flatten_ids, k2.offsets, sort_scratch,
                           s.bin_m, stream));
    VIEW_TRY(gsb_composite_fwd(cfg->width, cfg->height, 3, cfg->N, t1.means2d, t1.conics, k1.colors, opacity_logits, 1,
                               cam->antialiased ? t1.comps : nullptr, nullptr, k2.offsets, flatten_ids, M, k2.render,
                               k2.alphas, k2.last_ids, k2.comp_ws, s.comp_ws, stream));
    VIEW_TRY(gsb_tonemap_planar_fwd((int64_t)P, k2.render, k2.alphas, exposure, cfg->naive_tonemap, out, stream));
    return GSB_OK;
}

GSB_API int gsb_view_backward(const gsb_view_config *cfg, const gsb_camera *cam, const float *cam_pos_host, int64_t M,
                              const float *means, const float *quats, const float *scales,
                              const float *opacity_logits, const float *normals, const float *kd, const float *ks,
                              const float *fg_lut, const float *env_stack, const float *exposure, const void *keep1,
                              const void *keep2, void *tmp3, const float *v_out, float *v_means, float *v_quats,
                              float *v_scales, float *v_opacity_logits, float *v_normals, float *v_kd, float *v_ks,
                              float *v_env_stack, float *v_exposure, void *probe_start, void *probe_stop,
                              void *stream) {
    VIEW_CHECK_CFG(cfg);
    GSB_CHECK_ARG(cam && cam_pos_host && M >= 0 && keep1 && keep2 && tmp3 && v_out && v_exposure);
    Sizes s;
    VIEW_TRY(sizes(cfg, M, s));
    const size_t N = (size_t)cfg->N, P = (size_t)cfg->width * cfg->height;
    const size_t T = (size_t)((cfg->width + GSB_TILE - 1) / GSB_TILE) * ((cfg->height + GSB_TILE - 1) / GSB_TILE);
    Keep1 k1 = carve_keep1(const_cast<void *>(keep1), N);
    Keep2 k2 = carve_keep2(const_cast<void *>(keep2), T, P, s.comp_ws);
    Carver c3(tmp3);
    float *v_render = c3.take<float>(3 * P);
    float *v_alphas = c3.take<float>(P);
    float *acc = c3.take<float>(9 * N);          // the four atomic accumulators of this view, one zero-fill
    void *shade_ws = c3.take<char>(s.shade_ws);
    float *v_means2d = acc, *v_conics = acc + 2 * N, *v_colors = acc + 5 * N, *v_opac = acc + 8 * N;
    VIEW_TRY(gsb_tonemap_planar_bwd((int64_t)P, k2.render, exposure, cfg->naive_tonemap, v_out, v_render, v_alphas,
                                    v_exposure, stream));
    GSB_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(float) * 9 * N, (cudaStream_t)stream));
    // optional probe: two caller-owned cudaEvent_t recorded around the dominant stage (bench.py's live roofline timing)
    if (probe_start) GSB_CHECK_CUDA(cudaEventRecord((cudaEvent_t)probe_start, (cudaStream_t)stream));
    VIEW_TRY(gsb_composite_bwd(cfg->width, cfg->height, 3, cfg->N, k1.colors, nullptr, k2.offsets, M, k2.alphas,
                               k2.last_ids, v_render, v_alphas, v_means2d, v_conics, v_colors, v_opac, k2.comp_ws, stream));
    if (probe_stop) GSB_CHECK_CUDA(cudaEventRecord((cudaEvent_t)probe_stop, (cudaStream_t)stream));
    // both ADD into the caller's gradient buffers (views of a batch on one stream share them)
    VIEW_TRY(gsb_project_bwd(cfg->N, means, quats, scales, cam, k1.radii, v_means2d, nullptr, v_conics, nullptr, v_means,
                             v_quats, v_scales, opacity_logits, v_opac, v_opacity_logits, 1, stream));
    VIEW_TRY(gsb_shade_bwd(cfg->N, means, normals, kd, ks, cam_pos_host, fg_lut, cfg->lut_res, env_stack, cfg->R0, cfg->L,
                           cfg->Rb, cfg->min_roughness, cfg->max_metallic, cfg->env_min_roughness,
                           cfg->env_max_roughness, cfg->mode, v_colors, v_means, v_normals, v_kd, v_ks, v_env_stack,
                           s.shade_ws ? shade_ws : nullptr, s.shade_ws, 1, stream));
    return GSB_OK;
}

/* =====================================================================================================================
 * Batch driver: ALL views of a training batch in one call each way (the per-view loop of GeoSplatter.render_report,
 * rfstudio/model/geosplat.py:869-879, and its backward).
 *
 * Nothing on the host waits for a view's intersection count M: the M-sized arrays are carved for a capacity `m_cap` the
 * caller chooses (a previous count plus a margin), the count stays on the device (keep1.m_eff = min(M, m_cap), read by
 * the binning and compositing kernels) and the raw M of every view is published to `totals_out` (pinned host memory)
 * so that the caller can detect an overflow (M > m_cap: the farthest intersections of that view were dropped) and grow
 * the capacity.  The views are spread round-robin over the caller's streams; `main_stream` is forked into them and
 * joined again with events, so the two calls are ordinary stream-ordered work (and CUDA-graph capturable).
 *
 * Arenas (gsb_batch_bytes): keep    = n_views x (keep1 | keep2(m_cap))            forward -> backward
 *                           scratch = n_streams x max(tmp1 | tmp2(m_cap), tmp3)    inside either call
 * Gradient buffers (one per stream, zero-filled by the caller, summed into buffer 0 by gsb_batch_backward):
 *   [ env 4T | quats 4N | ks 2N | means 3N | scales 3N | logits N | normals 3N | kd 3N | exposure n_views | 1 spare ]
 * (v_scales is w.r.t. the LINEAR scales; `grad_scale` multiplies the summed result, e.g. 1 / (views x ranks)).
 * ================================================================================================================== */
namespace {

struct BatchSizes { Sizes v; size_t keep_view, scratch_fwd, scratch_bwd, scratch_stream; };

int batch_sizes(const gsb_view_config *c, int64_t m_cap, BatchSizes &b) {
    int rc = sizes(c, m_cap, b.v);
    if (rc != GSB_OK) return rc;
    b.keep_view = al(b.v.keep1) + al(b.v.keep2);
    b.scratch_fwd = al(b.v.tmp1) + al(b.v.tmp2);
    b.scratch_bwd = al(b.v.tmp3);
    b.scratch_stream = b.scratch_fwd > b.scratch_bwd ? b.scratch_fwd : b.scratch_bwd;
    return GSB_OK;
}

constexpr int MAX_STREAMS = 8;
struct GradBufs { float *p[MAX_STREAMS]; };

__global__ void __launch_bounds__(256) grad_sum_kernel(int n_bufs, GradBufs b, int64_t n4, float scale) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 a = reinterpret_cast<float4 *>(b.p[0])[i];
#pragma unroll
    for (int k = 1; k < MAX_STREAMS; ++k)
        if (k < n_bufs) {
            const float4 v = reinterpret_cast<const float4 *>(b.p[k])[i];
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
    a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
    reinterpret_cast<float4 *>(b.p[0])[i] = a;
}

// fork: every side stream waits for what `main` has queued so far; join: `main` waits for every side stream
int fork_join(cudaStream_t main, void *const *streams, int n, bool fork) {
    cudaEvent_t ev;
    if (fork) {
        GSB_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        GSB_CHECK_CUDA(cudaEventRecord(ev, main));
        for (int s = 0; s < n; ++s)
            if ((cudaStream_t)streams[s] != main) GSB_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)streams[s], ev, 0));
        GSB_CHECK_CUDA(cudaEventDestroy(ev));
        return GSB_OK;
    }
    for (int s = 0; s < n; ++s) {
        if ((cudaStream_t)streams[s] == main) continue;
        GSB_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        GSB_CHECK_CUDA(cudaEventRecord(ev, (cudaStream_t)streams[s]));
        GSB_CHECK_CUDA(cudaStreamWaitEvent(main, ev, 0));
        GSB_CHECK_CUDA(cudaEventDestroy(ev));
    }
    return GSB_OK;
}

}  // namespace

GSB_API int gsb_batch_bytes(const gsb_view_config *cfg, int32_t n_views, int32_t n_streams, int64_t m_cap,
                            size_t *bytes2_host) {
    VIEW_CHECK_CFG(cfg);
    GSB_CHECK_ARG(n_views >= 0 && n_streams >= 1 && n_streams <= MAX_STREAMS && m_cap >= 0 && bytes2_host != nullptr);
    BatchSizes b;
    VIEW_TRY(batch_sizes(cfg, m_cap, b));
    bytes2_host[0] = b.keep_view * (size_t)n_views + 256;
    bytes2_host[1] = b.scratch_stream * (size_t)n_streams + 256;
    return GSB_OK;
}

GSB_API int gsb_batch_grad_floats(const gsb_view_config *cfg, int32_t n_views, int64_t env_texels, int64_t *floats_host) {
    VIEW_CHECK_CFG(cfg);
    GSB_CHECK_ARG(n_views >= 0 && env_texels >= 0 && floats_host != nullptr);
    const int64_t n = 4 * env_texels + 19 * (int64_t)cfg->N + n_views + 1;   // + 1: room for the caller's sum over views
    *floats_host = (n + 3) & ~(int64_t)3;      // whole float4s
    return GSB_OK;
}

GSB_API int gsb_batch_forward(const gsb_view_config *cfg, int32_t n_views, const gsb_camera *cams,
                              const float *cam_pos_host, const float *means, const float *quats, const float *scales,
                              const float *opacity_logits, const float *normals, const float *kd, const float *ks,
                              const float *fg_lut, const float *env_stack, const float *exposures,
                              int32_t exposure_stride, void *keep, void *scratch, int64_t m_cap, int64_t *totals_out,
                              float *out, void *const *streams, int32_t n_streams, void *main_stream) {
    VIEW_CHECK_CFG(cfg);
    GSB_CHECK_ARG(n_views >= 0 && n_streams >= 1 && n_streams <= MAX_STREAMS && m_cap >= 0);
    if (n_views == 0) return GSB_OK;
    GSB_CHECK_ARG(cams && cam_pos_host && exposures && keep && scratch && out && streams);
    BatchSizes b;
    VIEW_TRY(batch_sizes(cfg, m_cap, b));
    const size_t N = (size_t)cfg->N, P = (size_t)cfg->width * cfg->height;
    const size_t T = (size_t)((cfg->width + GSB_TILE - 1) / GSB_TILE) * ((cfg->height + GSB_TILE - 1) / GSB_TILE);
    char *keep_base = Carver(keep).p, *scr_base = Carver(scratch).p;
    VIEW_TRY(fork_join((cudaStream_t)main_stream, streams, n_streams, true));
    for (int v = 0; v < n_views; ++v) {
        const gsb_camera *cam = cams + v;
        GSB_CHECK_ARG(cam->width == cfg->width && cam->height == cfg->height);
        void *st = streams[v % n_streams];
        char *kv = keep_base + b.keep_view * (size_t)v;
        char *sv = scr_base + b.scratch_stream * (size_t)(v % n_streams);
        void *keep1 = kv, *keep2 = kv + al(b.v.keep1), *tmp1 = sv, *tmp2 = sv + al(b.v.tmp1);
        Keep1 k1 = carve_keep1(keep1, N);
        Tmp1 t1 = carve_tmp1(tmp1, N, b.v.bin_n);
        Keep2 k2 = carve_keep2(keep2, T, P, b.v.comp_ws);
        Carver c2(tmp2);
        int32_t *flatten_ids = c2.take<int32_t>((size_t)m_cap);
        void *sort_scratch = c2.take<char>(b.v.bin_m);
        // ---- prepare: projection, depth order + intersection count (published, not awaited), shade
        VIEW_TRY(gsb_project_fwd(cfg->N, means, quats, scales, cam, k1.radii, t1.means2d, t1.depths, t1.conics,
                                 t1.comps, t1.tpg, st));
        VIEW_TRY(gsb_bin2_order(cfg->N, t1.depths, t1.tpg, t1.order, t1.cum, nullptr, t1.scratch, b.v.bin_n, st));
        VIEW_TRY(gsb_bin2_publish(cfg->N, t1.cum, m_cap, k1.m_eff, totals_out ? totals_out + v : nullptr, st));
        // the shade writes the compositing records (colour, folded conic, extents, opacity) straight into the
        // compositing workspace: no pack pass, no colour round trip
        VIEW_TRY(gsb_shade_fwd_impl(cfg->N, means, normals, kd, ks, cam_pos_host + 3 * v, fg_lut, cfg->lut_res, env_stack,
                                    cfg->R0, cfg->L, cfg->Rb, cfg->min_roughness, cfg->max_metallic,
                                    cfg->env_min_roughness, cfg->env_max_roughness, cfg->mode, k1.colors, t1.means2d,
                                    t1.conics, opacity_logits, cam->antialiased ? t1.comps : nullptr,
                                    gsb_composite_records(k2.comp_ws), st));
        // ---- finish: binning on the capacity, compositing, tone map
        // stable partition by tile in one pass where the image has at most 4096 tiles (tilepart.cu), else the radix sort
        if (gsb_tile_partition_supported((int)T) && !getenv("GSB_BIN2_RADIX"))
            VIEW_TRY(gsb_tile_partition_cap(cfg->N, m_cap, k1.m_eff, t1.means2d, k1.radii, t1.order, t1.cum, cam,
                                            flatten_ids, k2.offsets, sort_scratch, b.v.bin_m, st));
        else
            VIEW_TRY(gsb_bin2_sort_cap(cfg->N, m_cap, k1.m_eff, t1.means2d, k1.radii, t1.order, t1.cum, cam, flatten_ids,
                                       k2.offsets, sort_scratch, b.v.bin_m, st));
        VIEW_TRY(gsb_composite_fwd_impl(cfg->width, cfg->height, 3, cfg->N, t1.means2d, t1.conics, k1.colors,
                                        opacity_logits, 1, cam->antialiased ? t1.comps : nullptr, nullptr, k2.offsets,
                                        flatten_ids, m_cap, k1.m_eff, k2.render, k2.alphas, k2.last_ids, k2.comp_ws,
                                        b.v.comp_ws, 1, st));
        VIEW_TRY(gsb_tonemap_planar_fwd((int64_t)P, k2.render, k2.alphas, exposures + (size_t)exposure_stride * v,
                                        cfg->naive_tonemap, out + 4 * P * (size_t)v, st));
    }
    VIEW_TRY(fork_join((cudaStream_t)main_stream, streams, n_streams, false));
    return GSB_OK;
}

GSB_API int gsb_batch_backward(const gsb_view_config *cfg, int32_t n_views, const gsb_camera *cams,
                               const float *cam_pos_host, const float *means, const float *quats, const float *scales,
                               const float *opacity_logits, const float *normals, const float *kd, const float *ks,
                               const float *fg_lut, const float *env_stack, const float *exposures,
                               int32_t exposure_stride, const void *keep, void *scratch, int64_t m_cap,
                               const float *const *v_outs_host, int64_t env_texels,
                               float *const *grad_bufs, float grad_scale, void *const *streams, int32_t n_streams,
                               void *probe_events, void *main_stream) {
    VIEW_CHECK_CFG(cfg);
    GSB_CHECK_ARG(n_views >= 0 && n_streams >= 1 && n_streams <= MAX_STREAMS && m_cap >= 0 && env_texels >= 0);
    if (n_views == 0) return GSB_OK;
    GSB_CHECK_ARG(cams && cam_pos_host && exposures && keep && scratch && v_outs_host && grad_bufs && streams);
    BatchSizes b;
    VIEW_TRY(batch_sizes(cfg, m_cap, b));
    const size_t N = (size_t)cfg->N, P = (size_t)cfg->width * cfg->height;
    const size_t T = (size_t)((cfg->width + GSB_TILE - 1) / GSB_TILE) * ((cfg->height + GSB_TILE - 1) / GSB_TILE);
    int64_t n_floats = 0;
    VIEW_TRY(gsb_batch_grad_floats(cfg, n_views, env_texels, &n_floats));
    char *keep_base = Carver(const_cast<void *>(keep)).p, *scr_base = Carver(scratch).p;
    cudaEvent_t *probes = reinterpret_cast<cudaEvent_t *>(probe_events);   // optional: 2 per view, around the compositing backward
    VIEW_TRY(fork_join((cudaStream_t)main_stream, streams, n_streams, true));
    const int used = n_views < n_streams ? n_views : n_streams;
    for (int v = 0; v < n_views; ++v) {
        if (v_outs_host[v] == nullptr) continue;                   // no cotangent for this view
        const gsb_camera *cam = cams + v;
        cudaStream_t st = (cudaStream_t)streams[v % n_streams];
        char *kv = keep_base + b.keep_view * (size_t)v;
        Keep1 k1 = carve_keep1(kv, N);
        Keep2 k2 = carve_keep2(kv + al(b.v.keep1), T, P, b.v.comp_ws);
        Carver c3(scr_base + b.scratch_stream * (size_t)(v % n_streams));
        float *v_render = c3.take<float>(3 * P);
        float *v_alphas = c3.take<float>(P);
        float *acc = c3.take<float>(9 * N);
        void *shade_ws = c3.take<char>(b.v.shade_ws);
        float *v_means2d = acc, *v_conics = acc + 2 * N, *v_colors = acc + 5 * N, *v_opac = acc + 8 * N;
        float *g = grad_bufs[v % n_streams];
        float *g_env = g, *g_quats = g_env + 4 * (size_t)env_texels, *g_ks = g_quats + 4 * N, *g_means = g_ks + 2 * N;
        float *g_scales = g_means + 3 * N, *g_logits = g_scales + 3 * N, *g_normals = g_logits + N;
        float *g_kd = g_normals + 3 * N, *g_exp = g_kd + 3 * N;
        const float *exposure = exposures + (size_t)exposure_stride * v;
        VIEW_TRY(gsb_tonemap_planar_bwd((int64_t)P, k2.render, exposure, cfg->naive_tonemap, v_outs_host[v],
                                        v_render, v_alphas, g_exp + v, st));
        GSB_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(float) * 9 * N, st));
        if (probes) GSB_CHECK_CUDA(cudaEventRecord(probes[2 * v], st));
        VIEW_TRY(gsb_composite_bwd_impl(cfg->width, cfg->height, 3, cfg->N, k1.colors, nullptr, k2.offsets, m_cap,
                                        k1.m_eff, k2.alphas, k2.last_ids, v_render, v_alphas, v_means2d, v_conics,
                                        v_colors, v_opac, k2.comp_ws, st));
        if (probes) GSB_CHECK_CUDA(cudaEventRecord(probes[2 * v + 1], st));
        VIEW_TRY(gsb_project_bwd(cfg->N, means, quats, scales, cam, k1.radii, v_means2d, nullptr, v_conics, nullptr,
                                 g_means, g_quats, g_scales, opacity_logits, v_opac, g_logits, 1, st));
        VIEW_TRY(gsb_shade_bwd(cfg->N, means, normals, kd, ks, cam_pos_host + 3 * v, fg_lut, cfg->lut_res, env_stack,
                               cfg->R0, cfg->L, cfg->Rb, cfg->min_roughness, cfg->max_metallic, cfg->env_min_roughness,
                               cfg->env_max_roughness, cfg->mode, v_colors, g_means, g_normals, g_kd, g_ks, g_env,
                               b.v.shade_ws ? shade_ws : nullptr, b.v.shade_ws, 1, st));
    }
    VIEW_TRY(fork_join((cudaStream_t)main_stream, streams, n_streams, false));
    if (used > 1 || grad_scale != 1.0f) {
        const int64_t n4 = n_floats / 4;
        GradBufs gb;
        for (int k = 0; k < MAX_STREAMS; ++k) gb.p[k] = k < used ? grad_bufs[k] : nullptr;
        grad_sum_kernel<<<gsb_div_up(n4, 256), 256, 0, (cudaStream_t)main_stream>>>(used, gb, n4, grad_scale);
        GSB_CHECK_LAUNCH();
    }
    return GSB_OK;
}
