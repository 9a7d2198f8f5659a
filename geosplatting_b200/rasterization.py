"""Drop-in for ``gsplat.rasterization`` (gsplat~=1.4.0) as the reference calls it.

Reference call sites: rfstudio/model/gsplat.py:151-172 (ED), :240-261 (RGB), :334-355 (RGB, the
GeoSplatter hot path) and rfstudio/model/geosplat.py:276-295 (D=14 G-buffer).  Same names, argument
meaning and error behaviour; the work is done by libgeosplat_b200.so through the C ABI in
include/geosplat_b200.h.  There is no CPU path: CPU tensors raise.

Stage split (each an autograd Function over one C-ABI call pair):
    _Project   : gsb_project_fwd / gsb_project_bwd          (EWA projection, packed=True semantics)
    binning    : gsb_bin2_count + gsb_bin2_sort (depth order of the Gaussians, then a stable sort by tile; no grad)
    _Composite : gsb_composite_fwd / gsb_composite_bwd      (alpha compositing)
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import GsbCamera, call, f32c, ptr, stream_ptr
from ._lib import require_cuda as _require_cuda

TILE = 16
_SUPPORTED_CH = (1, 2, 3, 4, 8, 16)

_workspaces: Dict[Tuple[int, int], Tensor] = {}


def _workspace(device: torch.device, nbytes: int) -> Tensor:
    """Scan / sort scratch, one buffer per (device, stream): views on different streams run concurrently."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def make_camera(viewmat, K, width: int, height: int, *, near_plane=0.01, far_plane=1e10, eps2d=0.3,
                radius_clip=0.0, antialiased=False, camera_id=0) -> GsbCamera:
    """Host-side camera struct from a [4,4] world->camera matrix and a [3,3] intrinsic matrix.

    Matches rfstudio/graphics/_cameras.py:289-314 (`intrinsic_matrix`, `view_matrix`).  CUDA tensors
    are copied to the host (one small D2H copy), CPU tensors / nested lists are read directly.
    """
    vm = torch.as_tensor(viewmat).detach().to("cpu", torch.float32).reshape(16).tolist()
    k = torch.as_tensor(K).detach().to("cpu", torch.float32).reshape(9).tolist()
    cam = GsbCamera()
    cam.viewmat = (C.c_float * 16)(*vm)
    cam.fx, cam.fy, cam.cx, cam.cy = k[0], k[4], k[2], k[5]
    cam.width, cam.height = int(width), int(height)
    cam.near_plane, cam.far_plane = float(near_plane), float(far_plane)
    cam.eps2d, cam.radius_clip = float(eps2d), float(radius_clip)
    cam.antialiased = int(bool(antialiased))
    cam.camera_id = int(camera_id)
    return cam


class _Project(torch.autograd.Function):
    """means[N,3], quats[N,4], scales[N,3] -> means2d[N,2], depths[N], conics[N,3], comps[N],
    radii[N] i32 (0 = culled), tiles_per_gauss[N] i32."""

    @staticmethod
    def forward(ctx, means: Tensor, quats: Tensor, scales: Tensor, cam: GsbCamera):
        means_c, quats_c, scales_c = f32c(means), f32c(quats), f32c(scales)
        N = means_c.shape[0]
        dev = means_c.device
        radii = torch.empty(N, dtype=torch.int32, device=dev)
        means2d = torch.empty(N, 2, dtype=torch.float32, device=dev)
        depths = torch.empty(N, dtype=torch.float32, device=dev)
        conics = torch.empty(N, 3, dtype=torch.float32, device=dev)
        comps = torch.empty(N, dtype=torch.float32, device=dev)
        tpg = torch.empty(N, dtype=torch.int32, device=dev)
        call("gsb_project_fwd", dev, C.c_int32(N), ptr(means_c), ptr(quats_c), ptr(scales_c), C.byref(cam),
                                  ptr(radii), ptr(means2d), ptr(depths), ptr(conics), ptr(comps), ptr(tpg),
                                  stream_ptr(dev))
        ctx.save_for_backward(means_c, quats_c, scales_c, radii)
        ctx.cam = cam
        ctx.mark_non_differentiable(radii, tpg)
        return means2d, depths, conics, comps, radii, tpg

    @staticmethod
    def backward(ctx, v_means2d, v_depths, v_conics, v_comps, _vr, _vt):
        means, quats, scales, radii = ctx.saved_tensors
        N = means.shape[0]
        dev = means.device
        cam = ctx.cam

        def z(v, shape):
            return torch.zeros(shape, dtype=torch.float32, device=dev) if v is None else f32c(v)

        v_means2d = z(v_means2d, (N, 2))
        v_conics = z(v_conics, (N, 3))
        v_comps_c = z(v_comps, (N,)) if cam.antialiased else None
        v_depths_c = None if v_depths is None else f32c(v_depths)
        v_means = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_quats = torch.empty(N, 4, dtype=torch.float32, device=dev)
        v_scales = torch.empty(N, 3, dtype=torch.float32, device=dev)
        call("gsb_project_bwd", dev, C.c_int32(N), ptr(means), ptr(quats), ptr(scales), C.byref(cam), ptr(radii),
                                  ptr(v_means2d), ptr(v_depths_c), ptr(v_conics), ptr(v_comps_c), ptr(v_means),
                                  ptr(v_quats), ptr(v_scales), None, None, None, C.c_int32(0), stream_ptr(dev))
        return v_means, v_quats, v_scales, None


_free_slots: list = []


def _total_slot(device: torch.device) -> Tensor:
    """A pinned host int64 the device writes M into.  Slots come from a free list and go back to it once read
    (`_release_slot`), so any number of views can be outstanding; pinned memory is addressable from every device."""
    return _free_slots.pop() if _free_slots else torch.zeros(1, dtype=torch.int64).pin_memory()


def _release_slot(slot: Tensor) -> None:
    _free_slots.append(slot)


class BinCount:
    """First half of the binning (two-stage scheme, gsb_bin2_*): the depth order of the Gaussians and the prefix sum
    of tiles-per-Gaussian in that order are queued, and M is on its way to a pinned host slot.  `total()` waits for
    it -- call it as late as possible so that other work queued in between (the shade) covers the wait."""

    def __init__(self, tiles_per_gauss: Tensor, depths: Tensor):
        dev = tiles_per_gauss.device
        self.N = N = tiles_per_gauss.shape[0]
        self.order = self.cum = None
        self._M = 0 if N == 0 else None
        if N == 0:
            return
        nbytes = C.c_size_t(0)
        call("gsb_bin2_workspace_bytes", dev, C.c_int32(N), C.c_int64(0), C.byref(nbytes))
        ws = _workspace(dev, nbytes.value)
        self.order = torch.empty(N, dtype=torch.int32, device=dev)
        self.cum = torch.empty(N, dtype=torch.int64, device=dev)
        self._slot = _total_slot(dev)
        call("gsb_bin2_count", dev, C.c_int32(N), ptr(depths), ptr(tiles_per_gauss), ptr(self.order), ptr(self.cum),
             C.c_void_p(self._slot.data_ptr()), ptr(ws), C.c_size_t(ws.numel()), stream_ptr(dev))
        self._event = torch.cuda.Event()
        self._event.record(torch.cuda.current_stream(dev))

    def total(self) -> int:
        if self._M is None:
            self._event.synchronize()     # no device->host copy: the kernel stored M straight into pinned memory
            self._M = int(self._slot[0])
            _release_slot(self._slot)
            self._slot = None
        return self._M


def bin_finish(count: BinCount, means2d: Tensor, radii: Tensor, cam: GsbCamera):
    """Second half: (tile, Gaussian) pairs emitted in depth order, stable sort by tile, per-tile offsets.
    -> (flatten_ids[M] i32 (Gaussian index), isect_offsets[1,th,tw] i32); bit-identical to `bin_sort`."""
    dev = means2d.device
    N = means2d.shape[0]
    tw = (cam.width + TILE - 1) // TILE
    th = (cam.height + TILE - 1) // TILE
    offsets = torch.empty(th * tw, dtype=torch.int32, device=dev)
    M = count.total()
    flatten_ids = torch.empty(M, dtype=torch.int32, device=dev)
    if M == 0:
        offsets.zero_()
        return flatten_ids, offsets.view(1, th, tw)
    nbytes = C.c_size_t(0)
    call("gsb_bin2_workspace_bytes", dev, C.c_int32(0), C.c_int64(M), C.byref(nbytes))
    ws = _workspace(dev, nbytes.value)
    call("gsb_bin2_sort", dev, C.c_int32(N), C.c_int64(M), ptr(means2d), ptr(radii), ptr(count.order), ptr(count.cum),
         C.byref(cam), ptr(flatten_ids), ptr(offsets), ptr(ws), C.c_size_t(ws.numel()), stream_ptr(dev))
    return flatten_ids, offsets.view(1, th, tw)


def isect_ids_from_lists(flatten_ids: Tensor, offsets: Tensor, depths: Tensor) -> Tensor:
    """The sorted 64-bit keys gsplat exposes as info['isect_ids'] (tile << 32 | bits(depth)), rebuilt from the sorted
    lists; the two-stage binning never materialises them."""
    M = flatten_ids.shape[0]
    flat = offsets.reshape(-1).long()
    counts = torch.diff(torch.cat((flat, flat.new_tensor([M]))))
    tile = torch.repeat_interleave(torch.arange(flat.shape[0], device=flat.device), counts)
    dbits = depths.detach().contiguous().view(torch.int32)[flatten_ids.long()].long() & 0xFFFFFFFF
    return (tile << 32) | dbits


def bin_sort(means2d: Tensor, radii: Tensor, depths: Tensor, tiles_per_gauss: Tensor, cam: GsbCamera,
             n_cameras: int = 1):
    """gsplat's own stage split (isect_tiles -> one radix sort on the 64-bit (camera | tile | depth) keys ->
    isect_offset_encode): -> (isect_ids[M] i64 sorted, flatten_ids[M] i32, isect_offsets[n_cameras,th,tw] i32).
    `rasterization` uses the two-stage scheme above; this path is kept as the stage-faithful reference for it."""
    dev = means2d.device
    N = means2d.shape[0]
    tw = (cam.width + TILE - 1) // TILE
    th = (cam.height + TILE - 1) // TILE
    st = stream_ptr(dev)
    offsets = torch.empty(n_cameras * th * tw, dtype=torch.int32, device=dev)
    M = 0
    if N > 0:
        nbytes = C.c_size_t(0)
        call("gsb_bin_workspace_bytes", dev, C.c_int32(N), C.c_int64(0), C.byref(nbytes))
        ws = _workspace(dev, nbytes.value)
        cum = torch.empty(N, dtype=torch.int64, device=dev)
        call("gsb_isect_scan", dev, C.c_int32(N), ptr(tiles_per_gauss), ptr(cum), ptr(ws), C.c_size_t(ws.numel()), st)
        slot = _total_slot(dev)
        call("gsb_isect_total", dev, C.c_int32(N), ptr(cum), C.c_void_p(slot.data_ptr()), st)
        torch.cuda.current_stream(dev).synchronize()
        M = int(slot[0])
        _release_slot(slot)
    if M == 0:
        offsets.zero_()
        e64 = torch.empty(0, dtype=torch.int64, device=dev)
        return e64, torch.empty(0, dtype=torch.int32, device=dev), offsets.view(n_cameras, th, tw)
    keys = torch.empty(M, dtype=torch.int64, device=dev)
    vals = torch.empty(M, dtype=torch.int32, device=dev)
    call("gsb_isect_tiles", dev, C.c_int32(N), ptr(means2d), ptr(radii), ptr(depths), ptr(cum), C.byref(cam),
                              ptr(keys), ptr(vals), st)
    keys_s = torch.empty_like(keys)
    vals_s = torch.empty_like(vals)
    call("gsb_bin_workspace_bytes", dev, C.c_int32(N), C.c_int64(M), C.byref(nbytes))
    ws = _workspace(dev, nbytes.value)
    n_tiles = tw * th
    tile_bits = int(math.floor(math.log2(n_tiles))) + 1
    cam_bits = 0 if n_cameras <= 1 else int(math.floor(math.log2(n_cameras))) + 1
    call("gsb_sort_pairs", dev, C.c_int64(M), C.c_int32(32 + tile_bits + cam_bits), ptr(keys), ptr(vals),
                             ptr(keys_s), ptr(vals_s), ptr(ws), C.c_size_t(ws.numel()), st)
    call("gsb_isect_offsets", dev, C.c_int64(M), ptr(keys_s), C.c_int32(n_cameras), C.c_int32(tw), C.c_int32(th),
                                ptr(offsets), st)
    return keys_s, vals_s, offsets.view(n_cameras, th, tw)


class _Composite(torch.autograd.Function):
    """means2d[N,2], conics[N,3], colors[N,CH], opacities[N] (+ sorted lists) -> render[H,W,CH], alphas[H,W]."""

    @staticmethod
    def forward(ctx, means2d, conics, colors, opacities, background, offsets, flatten_ids, width, height):
        means2d_c, conics_c, colors_c, opac_c = f32c(means2d), f32c(conics), f32c(colors), f32c(opacities)
        bg_c = None if background is None else f32c(background)
        dev = means2d_c.device
        N, CH = colors_c.shape
        M = flatten_ids.shape[0]
        render = torch.empty(height, width, CH, dtype=torch.float32, device=dev)
        alphas = torch.empty(height, width, dtype=torch.float32, device=dev)
        last_ids = torch.empty(height, width, dtype=torch.int32, device=dev)
        nbytes = C.c_size_t(0)
        call("gsb_composite_workspace_bytes", dev, C.c_int64(N), C.c_int64(M), C.c_int32(width), C.c_int32(height),
             C.byref(nbytes))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)   # kept alive for the backward
        call("gsb_composite_fwd", dev, C.c_int32(width), C.c_int32(height), C.c_int32(CH), C.c_int64(N),
             ptr(means2d_c), ptr(conics_c), ptr(colors_c), ptr(opac_c), C.c_int32(0), None, ptr(bg_c), ptr(offsets),
             ptr(flatten_ids),
             C.c_int64(M), ptr(render), ptr(alphas), ptr(last_ids), ptr(ws), C.c_size_t(ws.numel()), stream_ptr(dev))
        ctx.save_for_backward(colors_c, offsets, alphas, last_ids, ws)
        ctx.bg = bg_c
        ctx.dims = (width, height, CH, M, N)
        return render, alphas

    @staticmethod
    def backward(ctx, v_render, v_alphas):
        colors, offsets, alphas, last_ids, ws = ctx.saved_tensors
        width, height, CH, M, N = ctx.dims
        dev = colors.device
        v_render = torch.zeros(height, width, CH, dtype=torch.float32, device=dev) if v_render is None \
            else f32c(v_render)
        v_alphas = torch.zeros_like(alphas) if v_alphas is None else f32c(v_alphas)
        v_means2d = torch.zeros(N, 2, dtype=torch.float32, device=dev)
        v_conics = torch.zeros(N, 3, dtype=torch.float32, device=dev)
        v_colors = torch.zeros(N, CH, dtype=torch.float32, device=dev)
        v_opac = torch.zeros(N, dtype=torch.float32, device=dev)
        call("gsb_composite_bwd", dev, C.c_int32(width), C.c_int32(height), C.c_int32(CH), C.c_int64(N), ptr(colors),
             ptr(ctx.bg), ptr(offsets), C.c_int64(M), ptr(alphas), ptr(last_ids), ptr(v_render), ptr(v_alphas),
             ptr(v_means2d), ptr(v_conics), ptr(v_colors), ptr(v_opac), ptr(ws), stream_ptr(dev))
        v_bg = None
        if ctx.bg is not None and ctx.needs_input_grad[4]:
            v_bg = ((1.0 - alphas).unsqueeze(-1) * v_render).sum(dim=(0, 1))
        return v_means2d, v_conics, v_colors, v_opac, v_bg, None, None, None, None


class _LazyInfo(dict):
    """`info` dict of gsplat.rasterization; the packed=True views (gaussian_ids, packed flatten_ids, ...)
    are materialised on first access because the reference's hot path discards `info`
    (rfstudio/model/gsplat.py:357)."""

    def __init__(self, eager: dict, lazy: dict):
        super().__init__(eager)
        self._lazy = lazy

    def __missing__(self, key):
        if key in self._lazy:
            val = self._lazy[key]()
            self[key] = val
            return val
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy

    def keys(self):
        return list(dict.keys(self)) + [k for k in self._lazy if not dict.__contains__(self, k)]

    def get(self, key, default=None):
        return self[key] if key in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())


def _pad_channels(colors: Tensor) -> Tuple[Tensor, int]:
    D = colors.shape[-1]
    for ch in _SUPPORTED_CH:
        if D <= ch:
            if D == ch:
                return colors, D
            pad = colors.new_zeros(colors.shape[:-1] + (ch - D,))
            return torch.cat((colors, pad), dim=-1), D
    raise ValueError(f"rasterization: at most {_SUPPORTED_CH[-1]} colour channels are supported, got {D}")


def _check_geometry(means, quats, scales, viewmats, Ks):
    if means.dim() != 2 or means.shape[1] != 3:
        raise AssertionError(f"means must be [N,3], got {tuple(means.shape)}")
    N = means.shape[0]
    assert quats.shape == (N, 4), quats.shape
    assert scales.shape == (N, 3), scales.shape
    assert viewmats.dim() == 3 and viewmats.shape[1:] == (4, 4), viewmats.shape
    assert Ks.shape == (viewmats.shape[0], 3, 3), Ks.shape
    _require_cuda(means, "rasterization")


class Projected:
    """State between `rasterization_begin` and `rasterization_end`: per camera the projection outputs and the
    pending intersection count."""

    def __init__(self, width, height, aa, N):
        self.width, self.height, self.aa, self.N = width, height, aa, N
        self.cams = []      # (cam, means2d, depths, conics, comps, radii, tpg, BinCount)


def rasterization_begin(means: Tensor, quats: Tensor, scales: Tensor, viewmats: Tensor, Ks: Tensor, width: int,
                        height: int, *, near_plane: float = 0.01, far_plane: float = 1e10, radius_clip: float = 0.0,
                        eps2d: float = 0.3, rasterize_mode: str = "classic") -> Projected:
    """First half of `rasterization`: everything that depends on geometry only (projection, tile counts, the
    prefix sum, M on its way to the host).  Work queued between begin and end -- RenderableAttrs.splat puts the
    shade there -- runs while the host would otherwise wait for M."""
    _check_geometry(means, quats, scales, viewmats, Ks)
    assert rasterize_mode in ("classic", "antialiased"), rasterize_mode
    aa = rasterize_mode == "antialiased"
    vm_host = viewmats.detach().to("cpu", torch.float32)
    k_host = Ks.detach().to("cpu", torch.float32)
    st = Projected(width, height, aa, means.shape[0])
    for c in range(viewmats.shape[0]):
        cam = make_camera(vm_host[c], k_host[c], width, height, near_plane=near_plane, far_plane=far_plane,
                          eps2d=eps2d, radius_clip=radius_clip, antialiased=aa, camera_id=0)
        means2d, depths, conics, comps, radii, tpg = _Project.apply(means, quats, scales, cam)
        st.cams.append((cam, means2d, depths, conics, comps, radii, tpg, BinCount(tpg, depths.detach())))
    return st


def rasterization_end(st: Projected, opacities: Tensor, colors: Tensor, *, backgrounds: Optional[Tensor] = None,
                      render_mode: str = "RGB", tile_size: int = TILE) -> Tuple[Tensor, Tensor, dict]:
    """Second half of `rasterization`: binning, sort and compositing with the per-Gaussian colours."""
    N, width, height, aa = st.N, st.width, st.height, st.aa
    Cn = len(st.cams)
    assert opacities.shape == (N,), opacities.shape
    assert colors.dim() == 2 and colors.shape[0] == N, colors.shape
    assert render_mode in ("RGB", "D", "ED", "RGB+D", "RGB+ED"), render_mode
    if backgrounds is not None:
        assert backgrounds.shape[0] == Cn, backgrounds.shape
    renders, alphas_out, per_cam = [], [], []
    for c, (cam, means2d, depths, conics, comps, radii, tpg, count) in enumerate(st.cams):
        opac = opacities * comps if aa else opacities
        if render_mode in ("RGB+D", "RGB+ED"):
            feats = torch.cat((colors, depths[:, None]), dim=-1)
        elif render_mode in ("D", "ED"):
            feats = depths[:, None]
        else:
            feats = colors
        bg = None
        if backgrounds is not None:
            bg = backgrounds[c]
            if render_mode in ("D", "ED"):
                bg = bg.new_zeros(1)            # gsplat 1.4.0: depth-only modes composite over a zero background
            if render_mode in ("RGB+D", "RGB+ED"):
                bg = torch.cat((bg, bg.new_zeros(1)))
        feats_p, D = _pad_channels(feats)
        if bg is not None and feats_p.shape[1] != bg.shape[0]:
            bg = torch.cat((bg, bg.new_zeros(feats_p.shape[1] - bg.shape[0])))
        flatten_ids, offsets = bin_finish(count, means2d.detach(), radii, cam)
        render, alpha = _Composite.apply(means2d, conics, feats_p, opac, bg, offsets.view(-1), flatten_ids,
                                         width, height)
        render = render[..., :D]
        if render_mode in ("ED", "RGB+ED"):
            render = torch.cat((render[..., :-1], render[..., -1:] / alpha[..., None].clamp(min=1e-10)), dim=-1)
        renders.append(render)
        alphas_out.append(alpha[..., None])
        per_cam.append((means2d, depths, conics, comps, radii, tpg, opac, None, flatten_ids, offsets))

    render = renders[0][None] if Cn == 1 else torch.stack(renders, 0)
    alpha = alphas_out[0][None] if Cn == 1 else torch.stack(alphas_out, 0)
    # ---- info (packed=True layout), lazily materialised -------------------------------------------
    def _packed():
        cams, gids = [], []
        for c, pc in enumerate(per_cam):
            g = torch.nonzero(pc[4] > 0).squeeze(-1)
            gids.append(g)
            cams.append(torch.full_like(g, c))
        return torch.cat(cams), torch.cat(gids), gids

    cache = {}

    def packed():
        if "p" not in cache:
            cache["p"] = _packed()
        return cache["p"]

    def gather(idx):
        return lambda: torch.cat([per_cam[c][idx][g] for c, g in enumerate(packed()[2])])

    def means2d_packed():
        """gsplat's densification idiom reads `info['means2d'].grad` after `retain_grad()` (rfstudio/model/gsplat.py:
        175-181, :264-270).  The packed tensor is gathered lazily, off the path to the image, so the screen-space gradient
        is delivered by hooks on the projected means: after backward, `.grad` of the returned tensor holds it."""
        gids = packed()[2]
        out = torch.cat([per_cam[c][0].detach()[g] for c, g in enumerate(gids)])
        if any(pc[0].requires_grad for pc in per_cam):
            out.requires_grad_(True)
            base = 0
            for c, g in enumerate(gids):
                def deliver(grad, g=g, lo=base, hi=base + g.shape[0]):
                    if out.grad is None:
                        out.grad = torch.zeros_like(out)
                    out.grad[lo:hi] += grad[g]
                if per_cam[c][0].requires_grad:
                    per_cam[c][0].register_hook(deliver)
                base += g.shape[0]
        return out

    def flat_packed():
        out, base = [], 0
        for c, g in enumerate(packed()[2]):
            rank = torch.cumsum((per_cam[c][4] > 0).to(torch.int32), 0, dtype=torch.int32) - 1
            out.append(rank[per_cam[c][8].long()] + base)
            base += g.shape[0]
        return torch.cat(out)

    def isect_ids_all():
        tw, th = (width + TILE - 1) // TILE, (height + TILE - 1) // TILE
        tile_bits = int(math.floor(math.log2(tw * th))) + 1
        return torch.cat([isect_ids_from_lists(pc[8], pc[9], pc[1]) | (c << (32 + tile_bits))
                          for c, pc in enumerate(per_cam)])

    def offsets_all():
        out, base = [], 0
        for pc in per_cam:
            out.append(pc[9] + base)
            base += pc[8].shape[0]
        return torch.cat(out, 0)

    eager = dict(width=width, height=height, tile_size=tile_size, n_cameras=Cn,
                 tile_width=(width + TILE - 1) // TILE, tile_height=(height + TILE - 1) // TILE)
    lazy = dict(
        camera_ids=lambda: packed()[0],
        gaussian_ids=lambda: packed()[1],
        radii=gather(4), means2d=means2d_packed, depths=gather(1), conics=gather(2),
        opacities=gather(6), tiles_per_gauss=gather(5),
        compensations=(gather(3) if aa else (lambda: None)),
        isect_ids=isect_ids_all, flatten_ids=flat_packed, isect_offsets=offsets_all,
        # unpacked extras (this library's native layout: per-Gaussian arrays, radii==0 => culled)
        flatten_gaussian_ids=lambda: torch.cat([pc[8] for pc in per_cam]),
        radii_unpacked=lambda: torch.stack([pc[4] for pc in per_cam]),
    )
    return render, alpha, _LazyInfo(eager, lazy)


def rasterization(
    means: Tensor,
    quats: Tensor,
    scales: Tensor,
    opacities: Tensor,
    colors: Tensor,
    viewmats: Tensor,
    Ks: Tensor,
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = True,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: str = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: str = "classic",
    channel_chunk: int = 32,
    distributed: bool = False,
    camera_model: str = "pinhole",
) -> Tuple[Tensor, Tensor, dict]:
    """Same contract as gsplat 1.4.0 ``rasterization`` for the argument set the reference uses.

    means[N,3] quats[N,4](wxyz) scales[N,3] opacities[N] colors[N,D] viewmats[C,4,4] Ks[C,3,3]
    -> render[C,H,W,D(+1)], alpha[C,H,W,1], info.  Differentiable w.r.t. means, quats, scales,
    opacities, colors (and backgrounds).
    """
    _check_geometry(means, quats, scales, viewmats, Ks)
    assert render_mode in ("RGB", "D", "ED", "RGB+D", "RGB+ED"), render_mode
    assert rasterize_mode in ("classic", "antialiased"), rasterize_mode
    if tile_size != TILE:
        raise NotImplementedError("geosplatting_b200: tile_size is fixed at 16 (rfstudio/model/gsplat.py:30)")
    if sh_degree is not None:
        raise NotImplementedError("geosplatting_b200: sh_degree must be None (GeoSplatter uses sh_degree=0 -> "
                                  "colors are passed raw, rfstudio/model/gsplat.py:305-307)")
    if camera_model != "pinhole" or distributed or sparse_grad or absgrad:
        raise NotImplementedError("geosplatting_b200: only pinhole / dense-grad / single-process rasterization")
    st = rasterization_begin(means, quats, scales, viewmats, Ks, width, height, near_plane=near_plane,
                             far_plane=far_plane, radius_clip=radius_clip, eps2d=eps2d, rasterize_mode=rasterize_mode)
    return rasterization_end(st, opacities, colors, backgrounds=backgrounds, render_mode=render_mode,
                             tile_size=tile_size)
