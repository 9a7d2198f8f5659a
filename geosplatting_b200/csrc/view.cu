// Native per-view driver: the three phases of one training view (prepare / finish / backward) as ONE C-ABI call each,
// composed from the stage entry points of this library on caller-provided arenas.
//
// The per-view host work of the reference is a Python loop over ~40 torch / plugin calls (rfstudio/model/geosplat.py:
// 53-132 + rfstudio/model/gsplat.py:284-358).  At ~1 ms of device time per view that orchestration is what a B200 ends
// up waiting for (measured: ~0.9 ms of host time per view even with one autograd node per batch, bench.py `batches`),
// so the sequencing lives here and the Python host makes three calls and a handful of allocations per view.
//
// Arena layouts (every block 256-byte aligned; sizes from gsb_view_bytes):
//   keep1 : radii[N] i32 | colors[N,3]                                              prepare -> backward
//   tmp1  : means2d[N,2] | depths[N] | conics[N,3] | comps[N] | tiles_per_gauss[N] | order[N] | cum[N] i64 | scratch
//                                                                                   prepare -> finish
//   keep2 : offsets[T] i32 | render[P,3] | alphas[P] | last_ids[P] i32 | composite workspace    finish -> backward
//   tmp2  : flatten_ids[M] i32 | sort scratch                                       finish only
//   tmp3  : v_render[P,3] | v_alphas[P] | v_means2d[N,2] v_conics[N,3] v_colors[N,3] v_opacities[N] | shade replicas
//                                                                                   backward only
#include "gsb_common.cuh"

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
    s.keep1 = al(4 * N) + al(12 * N) + 256;
    s.tmp1 = al(8 * N) + al(4 * N) + al(12 * N) + al(4 * N) + al(4 * N) + al(4 * N) + al(8 * N) + al(s.bin_n) + 256;
    s.keep2 = al(4 * T) + al(12 * P) + al(4 * P) + al(4 * P) + al(s.comp_ws) + 256;
    s.tmp2 = al(4 * (size_t)M) + al(s.bin_m) + 256;
    s.tmp3 = al(12 * P) + al(4 * P) + al(36 * N) + al(s.shade_ws) + 256;
    return GSB_OK;
}

struct Keep1 { int32_t *radii; float *colors; };
struct Tmp1 { float *means2d, *depths, *conics, *comps; int32_t *tpg, *order; int64_t *cum; void *scratch; };
struct Keep2 { int32_t *offsets; float *render, *alphas; int32_t *last_ids; void *comp_ws; };

Keep1 carve_keep1(void *a, size_t N) {
    Carver c(a);
    Keep1 k;
    k.radii = c.take<int32_t>(N);
    k.colors = c.take<float>(3 * N);
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
