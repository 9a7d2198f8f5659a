/*
 * geosplat_b200.h -- C ABI of libgeosplat_b200.so: the B200 (sm_100a) splat + PBR-shade hot path of
 * GeoSplatting, written from scratch.
 *
 * Conventions (every entry point):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in `_host`
 *     or the parameter is a `const gsb_camera*` (host struct, passed by value to the kernels);
 *   - the caller allocates every buffer (PyTorch `tensor.data_ptr()` in the Python host code) and passes
 *     the CUDA stream to launch on (`cudaStream_t` as `void*`; NULL = legacy default stream);
 *   - no hidden streams, host threads, allocations or synchronisation, except where stated;
 *   - returns 0 on success, a negative `GSB_E*` code on failure; `gsb_last_error()` returns the
 *     message for the calling thread.  The Python host raises RuntimeError on non-zero.
 *   - fp32 data, int32 ids, int64 sort keys, row-major, densely packed.
 *
 * Each group cites the reference interface it replaces (paths relative to the GeoSplatting tree).
 */
#ifndef GEOSPLAT_B200_H
#define GEOSPLAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSB_OK 0
#define GSB_EINVAL (-1)   /* bad argument */
#define GSB_ECUDA (-2)    /* CUDA runtime error (launch / cub) */
#define GSB_ENOMEM (-3)   /* caller-provided workspace too small */

#define GSB_TILE 16       /* rfstudio/model/gsplat.py:30 block_width = 16 */

const char *gsb_last_error(void);
int gsb_version(void);

/* One pinhole camera, as rfstudio/model/gsplat.py:340-348 passes it to gsplat.rasterization:
 * viewmat = Cameras.view_matrix (rfstudio/graphics/_cameras.py:299-314, world->camera, OpenCV axes),
 * fx..cy = Cameras.intrinsic_matrix (:289-297). */
typedef struct gsb_camera {
    float viewmat[16]; /* row-major 4x4 */
    float fx, fy, cx, cy;
    int32_t width, height;
    float near_plane, far_plane; /* 0.01, 1e10 (gsplat.py:346-347) */
    float eps2d;                 /* 0.3 (gsplat default) */
    float radius_clip;           /* 0.0 */
    int32_t antialiased;         /* rasterize_mode == 'antialiased' (geosplat.py:796) */
    int32_t camera_id;           /* goes into the high bits of the sort key */
} gsb_camera;

/* ---------------------------------------------------------------------------------------------
 * Rasterizer: replaces gsplat.rasterization (third-party gsplat~=1.4.0, pyproject.toml:24), called at
 * rfstudio/model/gsplat.py:334-355 and rfstudio/model/geosplat.py:276.  Stage split follows gsplat's
 * fully_fused_projection -> isect_tiles -> radix sort -> isect_offset_encode -> rasterize_to_pixels.
 * ------------------------------------------------------------------------------------------- */

/* EWA projection of N Gaussians (unpacked outputs; radii[i]==0 marks a culled Gaussian).
 * means[N,3] quats[N,4](wxyz, unnormalised) scales[N,3](linear) -> radii[N] means2d[N,2] depths[N]
 * conics[N,3] comps[N] (1.0 in classic mode) tiles_per_gauss[N]. */
int gsb_project_fwd(int32_t N, const float *means, const float *quats, const float *scales,
                    const gsb_camera *cam, int32_t *radii, float *means2d, float *depths, float *conics,
                    float *comps, int32_t *tiles_per_gauss, void *stream);

/* VJP of gsb_project_fwd.  v_depths may be NULL; v_comps is ignored in classic mode (may be NULL).
 * accumulate == 0: writes v_means[N,3] v_quats[N,4] v_scales[N,3] (and v_opacity_logits); culled rows get zeros.
 * accumulate != 0: adds to them (plain read-modify-write: the views of a batch that share a buffer must be on one
 * stream).
 * Fused opacity activation (all three NULL to disable): when gsb_composite_fwd was given opacity logits and comps,
 * pass the logits and gsb_composite_bwd's v_opacities here; the kernel adds d(opacity)/d(comp) to v_comps and writes
 * v_opacity_logits[N] (the backward of torch.sigmoid(opacities) * compensations, rfstudio/model/gsplat.py:338 +
 * gsplat rendering.py). */
int gsb_project_bwd(int32_t N, const float *means, const float *quats, const float *scales,
                    const gsb_camera *cam, const int32_t *radii, const float *v_means2d,
                    const float *v_depths, const float *v_conics, const float *v_comps, float *v_means,
                    float *v_quats, float *v_scales, const float *opacity_logits, const float *v_opacities_eff,
                    float *v_opacity_logits, int32_t accumulate, void *stream);

/* Bytes of scratch the scan / sort calls below need for N Gaussians and up to M intersections. */
int gsb_bin_workspace_bytes(int32_t N, int64_t M, size_t *bytes_host);

/* Inclusive prefix sum of tiles_per_gauss -> cum_tiles[N] (int64).  M = cum_tiles[N-1]. */
int gsb_isect_scan(int32_t N, const int32_t *tiles_per_gauss, int64_t *cum_tiles, void *workspace,
                   size_t workspace_bytes, void *stream);

/* Publishes M = cum_tiles[N-1] to *total_out with a plain store from the device.  total_out may be device memory
 * or PINNED HOST memory (cudaHostAlloc / torch pin_memory; device-addressable under unified addressing): the host
 * then learns M by waiting on an event of `stream` instead of queueing a copy behind whatever the copy engines are
 * doing (gsplat's isect_tiles reads M with a blocking device->host copy). */
int gsb_isect_total(int32_t N, const int64_t *cum_tiles, int64_t *total_out, void *stream);

/* Emits the M (key,val) pairs: key = camera_id << (32+tile_n_bits) | tile << 32 | bits(depth),
 * val = Gaussian index; Gaussian-major, tiles row-major. */
int gsb_isect_tiles(int32_t N, const float *means2d, const int32_t *radii, const float *depths,
                    const int64_t *cum_tiles, const gsb_camera *cam, int64_t *isect_ids,
                    int32_t *flatten_ids, void *stream);

/* Stable radix sort of the low `key_bits` bits (cub::DeviceRadixSort::SortPairs). */
int gsb_sort_pairs(int64_t M, int32_t key_bits, const int64_t *keys_in, const int32_t *vals_in,
                   int64_t *keys_out, int32_t *vals_out, void *workspace, size_t workspace_bytes, void *stream);

/* offsets[n_cameras * n_tiles]: first sorted position of every (camera, tile). */
int gsb_isect_offsets(int64_t M, const int64_t *sorted_isect_ids, int32_t n_cameras, int32_t tile_w,
                      int32_t tile_h, int32_t *offsets, void *stream);

/* ---- Two-stage binning: the same flatten_ids / offsets as the three calls above, without the 64-bit keys ----------
 * gsb_bin2_count : stable depth order of the N Gaussians (order[N]), prefix sum of tiles-per-Gaussian in that order
 *                  (cum_ordered[N]), and M stored to *total_out (device or pinned host memory, see gsb_isect_total);
 * gsb_bin2_sort  : (tile id, Gaussian) pairs emitted in depth order, stable radix sort on the tile id alone,
 *                  per-tile offsets -> flatten_ids[M], offsets[tile_w*tile_h].
 * A stable sort by tile of depth-ordered pairs equals a stable sort by (tile | depth) of Gaussian-major pairs, ties
 * included; 4 radix passes over N pairs + 2 over M instead of 6 over M.  One camera per call. */
int gsb_bin2_workspace_bytes(int32_t N, int64_t M, size_t *bytes_host);
int gsb_bin2_count(int32_t N, const float *depths, const int32_t *tiles_per_gauss, int32_t *order,
                   int64_t *cum_ordered, int64_t *total_out, void *workspace, size_t workspace_bytes, void *stream);
int gsb_bin2_sort(int32_t N, int64_t M, const float *means2d, const int32_t *radii, const int32_t *order,
                  const int64_t *cum_ordered, const gsb_camera *cam, int32_t *flatten_ids, int32_t *offsets,
                  void *workspace, size_t workspace_bytes, void *stream);
/* The emission step of gsb_bin2_sort on its own (tile_keys[M] u32, gauss_ids[M]). */
int gsb_isect_tiles_ordered(int32_t N, const float *means2d, const int32_t *radii, const int32_t *order,
                            const int64_t *cum_ordered, const gsb_camera *cam, uint32_t *tile_keys,
                            int32_t *gauss_ids, void *stream);

/* Bytes of scratch gsb_composite_fwd needs (per-Gaussian records + per-(tile, 8x4 sub-rectangle) lists). */
int gsb_composite_workspace_bytes(int64_t N, int64_t M, int32_t width, int32_t height, size_t *bytes_host);

/* Front-to-back alpha compositing of one camera (gsplat rasterize_to_pixels_fwd).  Per-Gaussian inputs [N,.]
 * are indexed by flatten_ids.  channels in {1,2,3,4,8,16}.  background[channels] may be NULL.
 * The opacity a Gaussian composites with is (opacity_is_logit ? sigmoid(opacities[g]) : opacities[g]) *
 * (comps ? comps[g] : 1): pass activated, compensated opacities with (0, NULL) -- gsplat's contract -- or the raw
 * logits and gsb_project_fwd's comps to fuse rfstudio/model/gsplat.py:338 and gsplat's `opacities * compensations`.
 * -> render[H,W,channels] alphas[H,W] last_ids[H,W] (position of the last contributor in the tile list).
 * `workspace` is filled here and must be handed UNCHANGED to gsb_composite_bwd of the same view. */
int gsb_composite_fwd(int32_t width, int32_t height, int32_t channels, int64_t N, const float *means2d,
                      const float *conics, const float *colors, const float *opacities, int32_t opacity_is_logit,
                      const float *comps, const float *background, const int32_t *offsets,
                      const int32_t *flatten_ids, int64_t M, float *render, float *alphas, int32_t *last_ids,
                      void *workspace, size_t workspace_bytes, void *stream);

/* VJP of gsb_composite_fwd (gsplat rasterize_to_pixels_bwd).  ACCUMULATES into v_means2d[N,2] v_conics[N,3]
 * v_colors[N,channels] v_opacities[N]: the caller zero-fills them.  v_opacities is the gradient w.r.t. the opacity
 * the Gaussian composited with (after any fused activation; gsb_project_bwd finishes that chain). */
int gsb_composite_bwd(int32_t width, int32_t height, int32_t channels, int64_t N, const float *colors,
                      const float *background, const int32_t *offsets, int64_t M, const float *alphas,
                      const int32_t *last_ids, const float *v_render, const float *v_alphas, float *v_means2d,
                      float *v_conics, float *v_colors, float *v_opacities, const void *workspace, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Split-sum shade: replaces the per-Gaussian shade block of RenderableAttrs.splat
 * (rfstudio/model/geosplat.py:83-121), TextureSplitSum.sample (rfstudio/graphics/_mesh/_texture.py:
 * 571-613, including the per-view _split_mipmaps re-materialisation :247-261) and the three
 * nvdiffrast.torch.texture calls behind them (geosplat.py:93, _texture.py:596,:604).
 *
 * Env-map stack (native layout of this library): RGBA float4 texels; specular levels l = 0..L-1 of
 * resolution R0>>l as [6,R,R] each, concatenated, followed by the diffuse base [6,Rb,Rb].
 * ------------------------------------------------------------------------------------------- */

/* Number of float4 texels of a stack (host-side helper, no GPU work). */
int gsb_envstack_texels(int32_t R0, int32_t L, int32_t Rb, int64_t *texels_host);

/* TextureSplitSum.mipmaps [6,4,R0,R0] (quad-tree pack, _texture.py:228-244) + base [6,Rb,Rb,3] -> stack. */
int gsb_envstack_pack(int32_t R0, int32_t L, int32_t Rb, const float *packed, const float *base, float *stack,
                      void *stream);

/* Stack gradient -> gradients in the reference layouts (overwrites the texels it owns; the caller
 * zero-fills v_packed because the quad-tree leaves one unused tail). */
int gsb_envstack_unpack_grad(int32_t R0, int32_t L, int32_t Rb, const float *v_stack, float *v_packed,
                             float *v_base, void *stream);

/* colors[N,3] from means[N,3] normals[N,3] kd[N,3] ks[N,2]; cam_pos_host[3] is a HOST pointer;
 * fg_lut [lut_res,lut_res,2] (rfstudio/graphics/shaders.py:22-26); mode 0 'pbr', 1 'diffuse', 2 'specular'
 * (geosplat.py:111-121). */
int gsb_shade_fwd(int32_t N, const float *means, const float *normals, const float *kd, const float *ks,
                  const float *cam_pos_host, const float *fg_lut, int32_t lut_res, const float *env_stack,
                  int32_t R0, int32_t L, int32_t Rb, float min_roughness, float max_metallic,
                  float env_min_roughness, float env_max_roughness, int32_t mode, float *colors, void *stream);

/* Bytes of scratch gsb_shade_bwd can use: private copies of the coarse env levels' gradients (their few thousand
 * texels take the gradients of every rough Gaussian; same-address reductions serialise in L2). */
int gsb_shade_workspace_bytes(int32_t R0, int32_t L, int32_t Rb, size_t *bytes_host);

/* VJP of gsb_shade_fwd.  Writes (accumulate == 0) or adds to (accumulate != 0, same-stream read-modify-write)
 * v_means/v_normals/v_kd/v_ks; always ACCUMULATES texel gradients into
 * v_env_stack (same layout as env_stack; lets one buffer collect all views of a step).  `workspace` (16-byte
 * aligned, gsb_shade_workspace_bytes) may be NULL: every texel gradient then goes straight to v_env_stack. */
int gsb_shade_bwd(int32_t N, const float *means, const float *normals, const float *kd, const float *ks,
                  const float *cam_pos_host, const float *fg_lut, int32_t lut_res, const float *env_stack,
                  int32_t R0, int32_t L, int32_t Rb, float min_roughness, float max_metallic,
                  float env_min_roughness, float env_max_roughness, int32_t mode, const float *v_colors,
                  float *v_means, float *v_normals, float *v_kd, float *v_ks, float *v_env_stack, void *workspace,
                  size_t workspace_bytes, int32_t accumulate, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Stand-alone texture sampling: replaces nvdiffrast.torch.texture (third-party, unpinned, README.md:36)
 * for the three call shapes on the path, so that the reference's own TextureSplitSum.sample
 * (_texture.py:596-611) and _CubeMapMip.backward (_texture.py:220-225) run on this library.
 * `level_ptrs_host` is a HOST array of n_levels DEVICE pointers to [6,R0>>l,R0>>l,channels] levels.
 * ------------------------------------------------------------------------------------------- */

/* filter_mode 'linear' (n_levels == 1, level == NULL) or 'linear-mipmap-linear' with explicit mips and a
 * per-sample mip_level_bias `level[N]`; boundary_mode 'cube'.  dirs[N,3] -> out[N,channels]. */
int gsb_texture_cube_fwd(int32_t N, int32_t channels, int32_t n_levels, const float *const *level_ptrs_host,
                         int32_t R0, const float *dirs, const float *level, float *out, void *stream);

/* VJP: ACCUMULATES texel gradients into v_level_ptrs_host[l] (entries or the array may be NULL), writes
 * v_dirs[N,3] and v_level[N] (either may be NULL). */
int gsb_texture_cube_bwd(int32_t N, int32_t channels, int32_t n_levels, const float *const *level_ptrs_host,
                         int32_t R0, const float *dirs, const float *level, const float *v_out,
                         float *const *v_level_ptrs_host, float *v_dirs, float *v_level, void *stream);

/* 2D, filter 'linear', boundary 'clamp', 2 channels (the FG LUT): tex[H,W,2], uv[N,2] -> out[N,2]. */
int gsb_texture2d_fwd(int32_t N, int32_t width, int32_t height, const float *tex, const float *uv, float *out,
                      void *stream);
/* VJP w.r.t. uv only (the LUT is a constant asset in the reference). */
int gsb_texture2d_bwd(int32_t N, int32_t width, int32_t height, const float *tex, const float *uv,
                      const float *v_out, float *v_uv, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Split-sum prefilter: replaces the reference's own plugin `rfstudio_render_utils`
 * (rfstudio/graphics/_mesh/_splitsum/c_src/torch_bindings.cpp:112-271 -> cubemap.cu:110-350) and the mip
 * chain `_CubeMapMip` (rfstudio/graphics/_mesh/_texture.py:199-226).  cubemap is [6,R,R,3]; `*_stride`
 * arguments (3 or 4 floats per texel) let a kernel read/write this library's float4 env stack directly.
 * ------------------------------------------------------------------------------------------- */

/* torch_bindings.cpp:112 diffuse_cubemap_fwd: cosine-weighted irradiance of every texel direction. */
int gsb_diffuse_cubemap_fwd(int32_t R, const float *cubemap, float *out, int32_t out_stride, void *stream);
/* torch_bindings.cpp:140 diffuse_cubemap_bwd (gather form, writes grad_in[6,R,R,3]). */
int gsb_diffuse_cubemap_bwd(int32_t R, const float *grad_out, int32_t grad_stride, float *grad_in, void *stream);
/* torch_bindings.cpp:168 specular_bounds: bounds[6,R,R,24] = per face (xmin,xmax,ymin,ymax) as floats. */
int gsb_specular_bounds(int32_t R, float costheta_cutoff, float *bounds, void *stream);
/* torch_bindings.cpp:193 specular_cubemap_fwd: out[6,R,R,4] = (sum w*rgb, sum w); normalize != 0 stores
 * (rgb/wsum, wsum) instead, which is what _wrap.py:157 computes next. */
int gsb_specular_cubemap_fwd(int32_t R, const float *cubemap, const float *bounds, float roughness,
                             float costheta_cutoff, int32_t normalize, float *out, void *workspace, void *stream);
/* Scratch for the two calls around this comment (direction table + pre-multiplied source); a NULL workspace
 * selects the slower table-free kernel. */
int gsb_specular_workspace_bytes(int32_t R, size_t *bytes_host);
/* torch_bindings.cpp:226 specular_cubemap_bwd: grad_out[6,R,R,4] (channels 0..2 read, like the plugin) ->
 * grad_in[6,R,R,3].  fwd_out != NULL: grad_out is the cotangent of the NORMALISED rgb and fwd_out[...,3]
 * holds wsum (the forward's own output). */
int gsb_specular_cubemap_bwd(int32_t R, const float *bounds, const float *grad_out, const float *fwd_out,
                             float roughness, float costheta_cutoff, float *grad_in, void *workspace, void *stream);
/* _CubeMapMip.forward: 2x2 box filter per face, in [6,2R,2R,.] -> out [6,R,R,.]. */
int gsb_cubemap_mip_fwd(int32_t R_out, const float *in, int32_t in_stride, float *out, int32_t out_stride,
                        void *stream);
/* _CubeMapMip.backward: grad_in[6,2R,2R,3] = bilinear cube resample of 0.25*grad_out[6,R,R,3]. */
int gsb_cubemap_mip_bwd(int32_t R_out, const float *grad_out, float *grad_in, void *stream);

/* ---------------------------------------------------------------------------------------------
 * MGAdaptor mesh -> Gaussian sampling, vertex normals, tone map.
 * ------------------------------------------------------------------------------------------- */

/* TriangleMesh.compute_vertex_normals_(fix=True) (rfstudio/graphics/_mesh/_triangle_mesh.py:588-614):
 * vertices[V,3], faces[F,3] int64 -> raw[V,3] (area-weighted sums, kept for the backward), normals[V,3]. */
int gsb_vertex_normals_fwd(int32_t V, int32_t F, const float *vertices, const int64_t *faces, float *raw,
                           float *normals, void *stream);
/* VJP: ACCUMULATES into v_vertices[V,3]; scratch[V,3] is caller-provided workspace. */
int gsb_vertex_normals_bwd(int32_t V, int32_t F, const float *vertices, const int64_t *faces, const float *raw,
                           const float *v_normals, float *scratch, float *v_vertices, void *stream);

/* MGAdapter().make(mesh, normal_interpolation = vertex_normals != NULL) (rfstudio/model/geosplat.py:426-472):
 * -> N = 6F Gaussians ordered [ring0:(e01,e12,e20), ring1:(...)] in blocks of F:
 * means[N,3] scales[N,3](log) quats[N,4](wxyz) normals[N,3] (Splats.colors) opacities[N](logit 0.99)
 * offsets[N,3] (= n * sqrt(area), detached). */
int gsb_mgadapter_fwd(int32_t F, const float *vertices, const float *vertex_normals, const int64_t *faces,
                      float *means, float *scales, float *quats, float *normals, float *opacities, float *offsets,
                      void *stream);
/* VJP: ACCUMULATES into v_vertices[V,3] and v_vertex_normals[V,3] (may be NULL). */
int gsb_mgadapter_bwd(int32_t F, const float *vertices, const float *vertex_normals, const int64_t *faces,
                      const float *v_means, const float *v_scales, const float *v_quats, const float *v_normals,
                      float *v_vertices, float *v_vertex_normals, void *stream);

/* _tone_mapping_naive (rfstudio/model/geosplat.py:474-476): rgba[P,4], exposure[1] (device) -> out[P,4]. */
int gsb_tonemap_fwd(int64_t P, const float *rgba, const float *exposure, float *out, void *stream);
/* VJP: writes v_rgba[P,4], ACCUMULATES into v_exposure[1]. */
int gsb_tonemap_bwd(int64_t P, const float *rgba, const float *exposure, const float *v_out, float *v_rgba,
                    float *v_exposure, void *stream);


/* The same two, fed by the rasterizer's own outputs render[P,3] / alphas[P] instead of a concatenated image
 * (torch.cat at rfstudio/model/gsplat.py:358).  naive != 0: _tone_mapping_naive; naive == 0: tone_type 'none'
 * (rgb * exposure, geosplat.py:124).  out / v_out are [P,4]; v_exposure[1] is ACCUMULATED. */
int gsb_tonemap_planar_fwd(int64_t P, const float *render, const float *alphas, const float *exposure,
                           int32_t naive, float *out, void *stream);
int gsb_tonemap_planar_bwd(int64_t P, const float *render, const float *exposure, int32_t naive,
                           const float *v_out, float *v_render, float *v_alphas, float *v_exposure, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Hash-grid encoding of the fields that produce kd / ks / z: replaces
 * HashEncoding.pytorch_fwd / tcnn.Encoding("HashGrid") behind rfstudio/model/components/encoding.py:231-241, used at
 * rfstudio/model/geosplat.py:579-598,:644-669.  Semantics = the reference's own `torch` backend (encoding.py:124-229):
 * x[N,3] in [-1,1]; table[L * 2^log2_T, F] (level-major, F must be 2); scalings_host[L] = floor(min_res * growth^l)
 * (HOST array); feats[N, L*F] level-major.  Features are bit-identical to that code (the file is built -fmad=false).
 * The MLP behind the encoding (rfstudio/nn/mlp.py) is three bias-free GEMMs and stays on the BLAS library.
 * ------------------------------------------------------------------------------------------- */
int gsb_hashgrid_fwd(int64_t N, const float *x, const float *table, int32_t L, int32_t F, int32_t log2_T,
                     const float *scalings_host, float *feats, void *stream);
/* VJP: ACCUMULATES table_grad_scale * d feats/d table . v_feats into v_table (same shape as table; may be NULL) and
 * WRITES v_x[N,3] (may be NULL).  table_grad_scale carries HashEncoding.grad_scaling (encoding.py:232-240: the
 * gradient reaching the table is multiplied by it, the one reaching x is not). */
int gsb_hashgrid_bwd(int64_t N, const float *x, const float *table, int32_t L, int32_t F, int32_t log2_T,
                     const float *scalings_host, const float *v_feats, float table_grad_scale, float *v_table,
                     float *v_x, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Per-view training loss on the hot path's image: replaces the per-view body of
 * rfstudio/trainer/geosplat_trainer.py:171-180 (random-background composite, SSIML1Loss of
 * rfstudio/loss/photometric_loss.py:72-112 over torchmetrics' SSIM, mask MSE).  rgba[H,W,4] rendered (tone-mapped
 * linear rgb + alpha), gt_rgba[H,W,4] ground truth (linear rgb + mask), bg[H,W,3] random background.
 *   gsb_loss_fwd: sums4 (device, overwritten) = {sum of the SSIM map over the interior and 3 channels,
 *                 sum |img1 - img2|, sum (mask - alpha)^2, loss}; maps[9,H,W] derivative maps for the backward.
 *                 loss = l (1 - sums4[0] / (3 (H-10)(W-10))) + (1 - l) sums4[1] / (3 H W) + c sums4[2] / (H W).
 *   gsb_loss_bwd: v_rgba[H,W,4] = *v_loss (device scalar) * d loss / d rgba.
 * ------------------------------------------------------------------------------------------- */
int gsb_loss_fwd(int32_t H, int32_t W, const float *rgba, const float *gt_rgba, const float *bg, float ssim_lambda,
                 float mask_coeff, float *sums4, float *maps, void *stream);
int gsb_loss_bwd(int32_t H, int32_t W, const float *rgba, const float *gt_rgba, const float *bg, const float *maps,
                 float ssim_lambda, float mask_coeff, const float *v_loss, float *v_rgba, void *stream);

/* ---------------------------------------------------------------------------------------------
 * FlexiCubes dual marching cubes: replaces FlexiCubes._get_case_id,
 * _identify_surf_edges, dual_marching_cubes, _compute_reg_loss, _triangulate, compute_entropy
 * (rfstudio/graphics/_mesh/_flexicubes.py:460-802) as GeoSplatter.get_geometry drives them (geosplat.py:751-769).
 * Per-cube / per-group / per-quad arithmetic and the ordering bookkeeping (one stable radix sort of the edge keys,
 * three prefix sums) are all here; geosplatting_b200/flexicubes.py allocates and states every buffer's meaning.  cubes[F,8] int32 grid-vertex ids (16-byte aligned); tables: check[256,5], num_vd[256], dmc[256,4,7], cube_edges[12,2]
 * (int32, device).  N surface cubes, E surface edges, Q dual vertices, K (group, edge) entries.
 * ------------------------------------------------------------------------------------------- */
/* Native sequencing of the integer part (what _get_case_id, _identify_surf_edges, the num_vd loop and the no-grad part
 * of _triangulate compute, _flexicubes.py:460-538, :640-690, :758-771).  Two calls, one device->host read each:
 *   gsb_fc_surface : classify all F cubes, compact the surface cube ids (ascending) into surf_ids[F]; *n_surf_host = N.
 *   gsb_fc_topology: for N surface cubes -- resolved cases, dual-vertex counts, numbering (vd_base, k_base), surface
 *                    edges (edge_of[N,12], surf_edges[<=12N,2]), quads (quad_entry[<=3N,4]: positions cube*12+edge into
 *                    vd_of, winding order, reference quad order); counts_host[4] = {E, n_quads, Q, K}.
 * Both synchronise `stream` before returning.  workspace: gsb_fc_workspace_bytes(F, N) bytes (N = 0 for gsb_fc_surface).
 * After gsb_fc_dual_fwd, gsb_fc_quad_gather turns quad_entry into quad_vd for gsb_fc_quad_fwd/bwd. */
int gsb_fc_workspace_bytes(int32_t F, int32_t N, size_t *bytes_host);
int gsb_fc_surface(int32_t F, const float *sdf, const int32_t *cubes, int32_t *cases, int32_t *surf_flag,
                   int32_t *surf_ids, void *workspace, size_t workspace_bytes, int32_t *n_surf_host, void *stream);
int gsb_fc_topology(int32_t F, int32_t N, int64_t V, int32_t R0, int32_t R1, int32_t R2, const float *sdf,
                    const int32_t *cubes, const int32_t *surf_ids, const int32_t *cases, const int32_t *surf_flag,
                    const int32_t *check_table, const int32_t *num_vd_table, const int32_t *dmc_table,
                    const int32_t *cube_edges, int32_t *case_ids, int32_t *num_vd, int32_t *vd_base, int32_t *k_base,
                    int32_t *edge_of, int32_t *surf_edges, int32_t *quad_entry, void *workspace, size_t workspace_bytes,
                    int32_t *counts_host, void *stream);
int gsb_fc_quad_gather(int32_t n_quads, const int32_t *quad_entry, const int32_t *vd_of, int32_t *quad_vd, void *stream);
/* The same stages one kernel per call (the pieces gsb_fc_surface / gsb_fc_topology sequence). */
int gsb_fc_classify(int32_t F, const float *sdf, const int32_t *cubes, int32_t *cases, int32_t *surf_flag, void *stream);
int gsb_fc_resolve(int32_t N, int32_t R0, int32_t R1, int32_t R2, const int32_t *surf_ids, const int32_t *cases,
                   const int32_t *surf_flag, const int32_t *check_table, const int32_t *num_vd_table,
                   const int32_t *dmc_table, int32_t *case_ids, int32_t *num_vd, int32_t *n_entries, void *stream);
int gsb_fc_edge_keys(int32_t N, int64_t V, const int32_t *surf_ids, const int32_t *cubes, const int32_t *cube_edges,
                     int64_t *keys, void *stream);
/* vd[Q,3] dual vertices, vd_gamma[Q] activated gamma of the owning cube, vd_of[N,12] dual vertex of every cube edge,
 * l_dev[K] (_compute_reg_loss). */
int gsb_fc_dual_fwd(int32_t N, const int32_t *surf_ids, const int32_t *case_ids, const int32_t *num_vd,
                    const int32_t *vd_base, const int32_t *k_base, const int32_t *dmc_table, const int32_t *cube_edges,
                    const int32_t *edge_of, const int32_t *surf_edges, const float *vertices, const float *sdf,
                    const float *alpha, const float *beta, const float *gamma, float *vd, float *vd_gamma,
                    int32_t *vd_of, float *l_dev, void *stream);
/* VJP: ACCUMULATES into v_vertices[V,3] v_sdf[V] v_alpha[F,8] v_beta[F,12] v_gamma[F] (raw parameters). */
int gsb_fc_dual_bwd(int32_t N, const int32_t *surf_ids, const int32_t *case_ids, const int32_t *num_vd,
                    const int32_t *vd_base, const int32_t *k_base, const int32_t *dmc_table, const int32_t *cube_edges,
                    const int32_t *edge_of, const int32_t *surf_edges, const float *vertices, const float *sdf,
                    const float *alpha, const float *beta, const float *gamma, const float *v_vd,
                    const float *v_vd_gamma, const float *v_l_dev, float *v_vertices, float *v_sdf, float *v_alpha,
                    float *v_beta, float *v_gamma, void *stream);
/* quad_vd[n_quads,4] in winding order -> centres[n_quads,3], faces[4 n_quads,3] int64 (centre of quad q = vertex Q + q). */
int gsb_fc_quad_fwd(int32_t n_quads, int32_t Q, const int32_t *quad_vd, const float *vd, const float *vd_gamma,
                    float *centres, int64_t *faces, void *stream);
/* VJP: ACCUMULATES into v_vd[Q,3] and v_vd_gamma[Q]. */
int gsb_fc_quad_bwd(int32_t n_quads, int32_t Q, const int32_t *quad_vd, const float *vd, const float *vd_gamma,
                    const float *v_centres, float *v_vd, float *v_vd_gamma, void *stream);
/* compute_entropy over grid_edges[U,2] int32 (every grid edge once): sums3 = {sum BCE(a|b), sum BCE(b|a), #sign-changing
 * edges}; entropy = (sums3[0] + sums3[1]) / sums3[2].  partials: scratch of 3 * GSB_FC_ENTROPY_REPLICAS floats.
 * bwd ACCUMULATES *v_loss * d entropy / d sdf into v_sdf[V]. */
#define GSB_FC_ENTROPY_REPLICAS 64
int gsb_fc_entropy_fwd(int64_t U, const int32_t *grid_edges, const float *sdf, float *sums3, float *partials,
                       void *stream);
int gsb_fc_entropy_bwd(int64_t U, const int32_t *grid_edges, const float *sdf, const float *sums3, const float *v_loss,
                       float *v_sdf, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Native per-view driver: one training view of RenderableAttrs.splat (rfstudio/model/geosplat.py:53-132, culling
 * off, tone_type 'naive' / 'none') + GSplatter.render_rgba (rfstudio/model/gsplat.py:284-358) as three calls that
 * sequence the stage entry points above on caller-provided arenas.  The host keeps three calls and a handful of
 * allocations per view instead of ~40 framework calls; the kernels and the results are the same.
 *   prepare  : projection, depth order + intersection count (M -> *total_out, device or pinned host), shade;
 *   finish   : binning by tile, compositing, tone map -> out[H,W,4]        (needs M on the host);
 *   backward : tone map / compositing / projection / shade VJPs; ADDS the per-Gaussian gradients, the env-stack
 *              texel gradients and v_exposure[1] into the caller's buffers (the views of a batch that share them
 *              must be on one stream; the caller zero-fills once per batch).  v_scales is w.r.t. the LINEAR scales.
 * Arenas: gsb_view_bytes -> {keep1, tmp1, keep2, tmp2, tmp3}; keep1 lives prepare..backward, tmp1 prepare..finish,
 * keep2 finish..backward, tmp2 inside finish, tmp3 inside backward (layouts: csrc/view.cu).
 * ------------------------------------------------------------------------------------------- */
typedef struct gsb_view_config {
    int32_t N, width, height;
    int32_t lut_res, R0, L, Rb;                 /* FG LUT resolution; env stack shape */
    float min_roughness, max_metallic;          /* geosplat.py:85-86 */
    float env_min_roughness, env_max_roughness; /* TextureSplitSum.min/max_roughness */
    int32_t mode;                               /* 0 pbr, 1 diffuse, 2 specular */
    int32_t naive_tonemap;                      /* 1: _tone_mapping_naive, 0: rgb * exposure */
} gsb_view_config;

int gsb_view_bytes(const gsb_view_config *cfg, int64_t M, size_t *bytes5_host);
int gsb_view_prepare(const gsb_view_config *cfg, const gsb_camera *cam, const float *cam_pos_host, const float *means,
                     const float *quats, const float *scales, const float *normals, const float *kd, const float *ks,
                     const float *fg_lut, const float *env_stack, void *keep1, void *tmp1, int64_t *total_out,
                     void *stream);
int gsb_view_finish(const gsb_view_config *cfg, const gsb_camera *cam, int64_t M, const float *opacity_logits,
                    const float *exposure, void *keep1, void *tmp1, void *keep2, void *tmp2, float *out, void *stream);
int gsb_view_backward(const gsb_view_config *cfg, const gsb_camera *cam, const float *cam_pos_host, int64_t M,
                      const float *means, const float *quats, const float *scales, const float *opacity_logits,
                      const float *normals, const float *kd, const float *ks, const float *fg_lut,
                      const float *env_stack, const float *exposure, const void *keep1, const void *keep2, void *tmp3,
                      const float *v_out, float *v_means, float *v_quats, float *v_scales, float *v_opacity_logits,
                      float *v_normals, float *v_kd, float *v_ks, float *v_env_stack, float *v_exposure,
                      void *probe_start, void *probe_stop, void *stream);
/* probe_start / probe_stop: optional caller-owned cudaEvent_t (NULL to skip) recorded on `stream` right before and
 * after the compositing backward, the dominant stage, so that a benchmark can time it inside a running batch. */

/* -------------------------------------------------------------------------------------------
 * Cached specular prefilter ("plan"): the GGX weights of specular_cubemap_fwd / _bwd depend on the resolution, the
 * roughness and the cone only -- not on the cube map, which is what changes every training step (it is a parameter:
 * rfstudio/model/geosplat.py:741-748, :780-785).  A plan stores them once (32 floats per tap of every 8x4 patch of
 * output texels, ~5.5 GB for the six levels of a 512^2 map) and the per-step passes stream them.
 *   count : counts[(6 R^2 / 32) * 6][2] int32 = {row segments, taps} per (patch, face);
 *   fill  : seg_start / tap_start = exclusive prefix sums of the counts (one more entry), segs[n_segs][4] int32,
 *           weights[n_taps * 32] float;
 *   fwd / bwd : same arguments and results as gsb_specular_cubemap_fwd / _bwd with the plan in place of `bounds`.
 * R must be a multiple of 8, <= 1024.  workspace: gsb_specular_workspace_bytes(R).
 * ------------------------------------------------------------------------------------------- */
int gsb_specular_plan_count(int32_t R, const float *bounds, float costheta_cutoff, int32_t *counts, void *workspace,
                            void *stream);
int gsb_specular_plan_fill(int32_t R, const float *bounds, float roughness, float costheta_cutoff,
                           const int32_t *seg_start, const int32_t *tap_start, int32_t *segs, float *weights,
                           void *workspace, void *stream);
int gsb_specular_plan_fwd(int32_t R, const float *cubemap, const int32_t *seg_start, const int32_t *segs,
                          const float *weights, int32_t normalize, float *out, void *workspace, void *stream);
int gsb_specular_plan_bwd(int32_t R, const int32_t *seg_start, const int32_t *segs, const float *weights,
                          const float *grad_out, const float *fwd_out, float *grad_in, void *workspace, void *stream);

/* -------------------------------------------------------------------------------------------
 * The fields' MLPs, fused: replaces rfstudio/nn/mlp.py:125-145 (nn.Linear + F.relu per layer,
 * activation after the last) for the shapes of rfstudio/model/geosplat.py:485-518: x[N,32] -> 32 [-> 32] -> dout <= 4,
 * no bias.  Weights are nn.Linear's [out,in] row-major.  n_hidden: 1 or 2 (w1 NULL for 1).  activation: 0 none,
 * 1 sigmoid.  ref_round_scale s != 0: the input is taken as x*s + x*(1-s) in fp32 (two roundings), the value
 * rfstudio/model/components/encoding.py:239-240 hands the MLP.  gsb_mlp_bwd recomputes the forward and writes v_x[N,32]
 * (NULL to skip), v_w0[32,32], v_w1[32,32], v_wout[dout,32] (zero-filled by the call).
 * ------------------------------------------------------------------------------------------- */
int gsb_mlp_fwd(int64_t N, const float *x, const float *w0, const float *w1, const float *wout, int32_t n_hidden,
                int32_t dout, int32_t activation, float ref_round_scale, float *y, void *stream);
int gsb_mlp_bwd(int64_t N, const float *x, const float *w0, const float *w1, const float *wout, int32_t n_hidden,
                int32_t dout, int32_t activation, float ref_round_scale, const float *v_y, float *v_x, float *v_w0,
                float *v_w1, float *v_wout, void *stream);

/* -------------------------------------------------------------------------------------------
 * Batch driver: every view of a training batch in ONE call each way -- the per-view loop of
 * GeoSplatter.render_report (rfstudio/model/geosplat.py:869-879: `for i in range(batch_size): attrs.splat(...)`) and
 * its backward.  No host wait on any view's intersection count: the M-sized arrays are carved for the capacity
 * `m_cap`, the count stays on the device, the raw counts are published to `totals_out[n_views]` (pinned host memory;
 * M > m_cap means the farthest intersections of that view were dropped -- grow m_cap and redo the batch).
 *   keep / scratch: gsb_batch_bytes -> {keep, scratch}; keep lives forward..backward, scratch inside either call.
 *   cams[n_views], cam_pos_host[n_views][3]: host arrays.  exposures: device, element v at exposures[v * stride].
 *   out: [n_views][H][W][4]; v_outs_host[n_views]: host array of device pointers to the [H][W][4] cotangents (NULL =
 *   this view has none).  streams[n_streams <= 8]: the views are spread round-robin over them;
 *   main_stream is forked into them and joined again (results are ready on main_stream).
 *   grad_bufs[n_streams]: zero-filled buffers of gsb_batch_grad_floats floats each, layout
 *     [ env 4T | quats 4N | ks 2N | means 3N | scales(linear) 3N | logits N | normals 3N | kd 3N | exposure n_views |
 *       1 spare float (the caller's sum over views) ];
 *     the sum over streams times grad_scale ends up in grad_bufs[0].
 *   probe_events: optional
 *   cudaEvent_t[2 * n_views] recorded around each view's compositing backward.
 * ------------------------------------------------------------------------------------------- */
int gsb_batch_bytes(const gsb_view_config *cfg, int32_t n_views, int32_t n_streams, int64_t m_cap, size_t *bytes2_host);
int gsb_batch_grad_floats(const gsb_view_config *cfg, int32_t n_views, int64_t env_texels, int64_t *floats_host);
int gsb_batch_forward(const gsb_view_config *cfg, int32_t n_views, const gsb_camera *cams, const float *cam_pos_host,
                      const float *means, const float *quats, const float *scales, const float *opacity_logits,
                      const float *normals, const float *kd, const float *ks, const float *fg_lut,
                      const float *env_stack, const float *exposures, int32_t exposure_stride, void *keep,
                      void *scratch, int64_t m_cap, int64_t *totals_out, float *out, void *const *streams,
                      int32_t n_streams, void *main_stream);
int gsb_batch_backward(const gsb_view_config *cfg, int32_t n_views, const gsb_camera *cams, const float *cam_pos_host,
                       const float *means, const float *quats, const float *scales, const float *opacity_logits,
                       const float *normals, const float *kd, const float *ks, const float *fg_lut,
                       const float *env_stack, const float *exposures, int32_t exposure_stride, const void *keep,
                       void *scratch, int64_t m_cap, const float *const *v_outs_host,
                       int64_t env_texels, float *const *grad_bufs, float grad_scale, void *const *streams,
                       int32_t n_streams, void *probe_events, void *main_stream);

#ifdef __cplusplus
}
#endif
#endif /* GEOSPLAT_B200_H */
