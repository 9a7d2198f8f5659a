"""One autograd node per view for the whole hot path of RenderableAttrs.splat (rfstudio/model/geosplat.py:53-132
with culling=False): activations -> projection -> intersection count -> split-sum shade -> binning/sort ->
compositing -> tone map, and the reverse chain in one backward.

Same C-ABI kernels as the stage-by-stage operators (rasterization.py, shade.py, mgadapter.py) and the same results;
what it removes is host work and glue kernels: one autograd node instead of ~15, no torch.cat / stack / slice copies
of the image (gsb_tonemap_planar_*), sigmoid(opacity) * compensation folded into the record packing and its chain rule
into gsb_project_bwd, one zero-fill for the four per-Gaussian gradient accumulators, cached camera structs.  At
~1.3 ms of device time per view the host would otherwise be the bottleneck (scripts/host_overhead.py).
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import torch
from torch import Tensor

from ._lib import call, f32c, ptr, stream_ptr
from .rasterization import BinCount, bin_finish, make_camera
from .scenes import PinholeCamera
from .shade import MODES, EnvStack, shade_workspace


class ViewMeta(NamedTuple):
    R0: int
    L: int
    Rb: int
    env_min_roughness: float
    env_max_roughness: float
    min_roughness: float
    max_metallic: float
    mode: int
    antialiased: bool
    naive_tonemap: int


def _camera_struct(camera: PinholeCamera, antialiased: bool):
    """(gsb_camera, cam_pos float[3]) for a camera, cached on the camera object (cameras are immutable here, as
    the reference's `Cameras` tensors are)."""
    cache = camera.__dict__.setdefault("_gsb_cache", {})
    hit = cache.get(antialiased)
    if hit is None:
        cam = make_camera(camera.view_matrix, camera.intrinsic_matrix, camera.width, camera.height, near_plane=0.01,
                          far_plane=1e10, antialiased=antialiased)
        pos = (C.c_float * 3)(*[float(x) for x in camera.c2w[:, 3]])
        hit = cache[antialiased] = (cam, pos)
    return hit


class _SplatView(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, log_scales, quats, logits, kd, ks, normals, env_data, exposure, camera, lut, meta):
        means_c, quats_c, kd_c, ks_c, normals_c = f32c(means), f32c(quats), f32c(kd), f32c(ks), f32c(normals)
        env_c, exp_c = f32c(env_data), f32c(exposure).reshape(1)
        dev = means_c.device
        N = means_c.shape[0]
        logit_c = f32c(logits).reshape(N)
        scales = f32c(log_scales).exp()                       # rfstudio/model/gsplat.py:337
        cam, cam_pos = _camera_struct(camera, meta.antialiased)
        W, H = cam.width, cam.height
        st = stream_ptr(dev)

        fbuf = torch.empty(7 * N, dtype=torch.float32, device=dev)
        means2d, depths = fbuf[:2 * N].view(N, 2), fbuf[2 * N:3 * N]
        conics, comps = fbuf[3 * N:6 * N].view(N, 3), fbuf[6 * N:]
        ibuf = torch.empty(2 * N, dtype=torch.int32, device=dev)
        radii, tpg = ibuf[:N], ibuf[N:]
        call("gsb_project_fwd", dev, C.c_int32(N), ptr(means_c), ptr(quats_c), ptr(scales), C.byref(cam), ptr(radii),
             ptr(means2d), ptr(depths), ptr(conics), ptr(comps), ptr(tpg), st)
        count = BinCount(tpg, depths)                          # M is on its way to the host ...

        colors = torch.empty(N, 3, dtype=torch.float32, device=dev)
        call("gsb_shade_fwd", dev, C.c_int32(N), ptr(means_c), ptr(normals_c), ptr(kd_c), ptr(ks_c), cam_pos, ptr(lut),
             C.c_int32(lut.shape[0]), ptr(env_c), C.c_int32(meta.R0), C.c_int32(meta.L), C.c_int32(meta.Rb),
             C.c_float(meta.min_roughness), C.c_float(meta.max_metallic), C.c_float(meta.env_min_roughness),
             C.c_float(meta.env_max_roughness), C.c_int32(meta.mode), ptr(colors), st)

        flatten_ids, offsets = bin_finish(count, means2d, radii, cam)       # ... and is awaited only here
        M = flatten_ids.shape[0]
        render = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
        alphas = torch.empty(H, W, dtype=torch.float32, device=dev)
        last_ids = torch.empty(H, W, dtype=torch.int32, device=dev)
        nbytes = C.c_size_t(0)
        call("gsb_composite_workspace_bytes", dev, C.c_int64(N), C.c_int64(M), C.c_int32(W), C.c_int32(H),
             C.byref(nbytes))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)    # kept alive for the backward
        call("gsb_composite_fwd", dev, C.c_int32(W), C.c_int32(H), C.c_int32(3), C.c_int64(N), ptr(means2d),
             ptr(conics), ptr(colors), ptr(logit_c), C.c_int32(1), ptr(comps) if meta.antialiased else None, None,
             ptr(offsets), ptr(flatten_ids), C.c_int64(M), ptr(render), ptr(alphas), ptr(last_ids), ptr(ws),
             C.c_size_t(ws.numel()), st)
        out = torch.empty(H, W, 4, dtype=torch.float32, device=dev)
        call("gsb_tonemap_planar_fwd", dev, C.c_int64(H * W), ptr(render), ptr(alphas), ptr(exp_c),
             C.c_int32(meta.naive_tonemap), ptr(out), st)

        ctx.save_for_backward(means_c, quats_c, scales, logit_c, kd_c, ks_c, normals_c, env_c, exp_c, lut, radii,
                              colors, offsets, render, alphas, last_ids, ws)
        ctx.misc = (cam, cam_pos, meta, M, tuple(logits.shape), tuple(exposure.shape))
        return out

    @staticmethod
    def backward(ctx, v_out):
        (means, quats, scales, logit, kd, ks, normals, env, exp_c, lut, radii, colors, offsets, render, alphas,
         last_ids, ws) = ctx.saved_tensors
        cam, cam_pos, meta, M, logits_shape, exposure_shape = ctx.misc
        dev = means.device
        N = means.shape[0]
        W, H = cam.width, cam.height
        st = stream_ptr(dev)
        v_out = f32c(v_out)

        v_render = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
        v_alphas = torch.empty(H, W, dtype=torch.float32, device=dev)
        v_exp = torch.zeros(1, dtype=torch.float32, device=dev)
        call("gsb_tonemap_planar_bwd", dev, C.c_int64(H * W), ptr(render), ptr(exp_c), C.c_int32(meta.naive_tonemap),
             ptr(v_out), ptr(v_render), ptr(v_alphas), ptr(v_exp), st)

        acc = torch.zeros(9 * N, dtype=torch.float32, device=dev)       # the four atomic accumulators, one fill
        v_means2d, v_conics = acc[:2 * N], acc[2 * N:5 * N]
        v_colors, v_opac = acc[5 * N:8 * N], acc[8 * N:]
        call("gsb_composite_bwd", dev, C.c_int32(W), C.c_int32(H), C.c_int32(3), C.c_int64(N), ptr(colors), None,
             ptr(offsets), C.c_int64(M), ptr(alphas), ptr(last_ids), ptr(v_render), ptr(v_alphas), ptr(v_means2d),
             ptr(v_conics), ptr(v_colors), ptr(v_opac), ptr(ws), st)

        v_means = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_quats = torch.empty(N, 4, dtype=torch.float32, device=dev)
        v_scales = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_logits = torch.empty(N, dtype=torch.float32, device=dev)
        call("gsb_project_bwd", dev, C.c_int32(N), ptr(means), ptr(quats), ptr(scales), C.byref(cam), ptr(radii),
             ptr(v_means2d), None, ptr(v_conics), None, ptr(v_means), ptr(v_quats), ptr(v_scales), ptr(logit),
             ptr(v_opac), ptr(v_logits), st)
        v_scales.mul_(scales)                                            # d exp(s) / d s

        v_means_s = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_normals = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_kd = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_ks = torch.empty(N, 2, dtype=torch.float32, device=dev)
        v_env = torch.zeros_like(env)
        sws = shade_workspace(dev, meta.R0, meta.L, meta.Rb)
        call("gsb_shade_bwd", dev, C.c_int32(N), ptr(means), ptr(normals), ptr(kd), ptr(ks), cam_pos, ptr(lut),
             C.c_int32(lut.shape[0]), ptr(env), C.c_int32(meta.R0), C.c_int32(meta.L), C.c_int32(meta.Rb),
             C.c_float(meta.min_roughness), C.c_float(meta.max_metallic), C.c_float(meta.env_min_roughness),
             C.c_float(meta.env_max_roughness), C.c_int32(meta.mode), ptr(v_colors), ptr(v_means_s), ptr(v_normals),
             ptr(v_kd), ptr(v_ks), ptr(v_env), ptr(sws), C.c_size_t(sws.numel()), st)
        v_means.add_(v_means_s)                                          # via the projection and via the view vector
        return (v_means, v_scales, v_quats, v_logits.view(logits_shape), v_kd, v_ks, v_normals, v_env,
                v_exp.reshape(exposure_shape), None, None, None)


def splat_view(means: Tensor, log_scales: Tensor, quats: Tensor, opacity_logits: Tensor, kd: Tensor, ks: Tensor,
               normals: Tensor, camera: PinholeCamera, *, exposure: Tensor, envmap: EnvStack, fg_lut: Tensor,
               min_roughness: float, max_metallic: float, mode: str = "pbr", tone_type: str = "naive",
               rasterize_mode: str = "antialiased") -> Tensor:
    """[H,W,4] tone-mapped RGBA of one view; differentiable w.r.t. every tensor argument and `envmap.data`."""
    if mode not in MODES:
        raise ValueError(mode)
    if tone_type not in ("naive", "none"):
        raise ValueError(tone_type)
    if rasterize_mode not in ("antialiased", "classic"):
        raise ValueError(f"Unknown rasterize_mode: {rasterize_mode}")
    if not means.is_cuda:
        raise RuntimeError("geosplatting_b200.splat_view needs CUDA tensors; there is no CPU path")
    lut = f32c(fg_lut).reshape(fg_lut.shape[-3], fg_lut.shape[-2], 2)
    meta = ViewMeta(envmap.R0, envmap.L, envmap.Rb, envmap.min_roughness, envmap.max_roughness, float(min_roughness),
                    float(max_metallic), MODES[mode], rasterize_mode == "antialiased", int(tone_type == "naive"))
    return _SplatView.apply(means, log_scales, quats, opacity_logits, kd, ks, normals, envmap.data, exposure, camera,
                            lut, meta)
