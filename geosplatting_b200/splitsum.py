"""Host side of the split-sum prefilter (C ABI: gsb_diffuse_cubemap_*, gsb_specular_*, gsb_cubemap_mip_*).

Mirrors, with the same names and argument meaning:
    rfstudio_render_utils plugin    _splitsum/c_src/torch_bindings.cpp:112-271  -> class `render_utils`
    diffuse_cubemap / specular_cubemap   _splitsum/_wrap.py:95-157               -> same names
    __ndfBounds                      _splitsum/_wrap.py:120-135                  -> ndf_bounds
    _CubeMapMip                      graphics/_mesh/_texture.py:199-226          -> cubemap_mip
    TextureCubeMap.as_splitsum       graphics/_mesh/_texture.py:530-557          -> as_splitsum / as_envstack
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import numpy as np
import torch
from torch import Tensor

from ._lib import call, f32c, ptr, stream_ptr
from ._lib import require_cuda as _require_cuda
from .shade import EnvStack


def _bounds_device(index) -> torch.device:
    """Device of a bounds table from the `device.index` the reference passes (_wrap.py:151; None = current device)."""
    return torch.device("cuda", torch.cuda.current_device() if index is None else index)


def _spec_workspace(R: int, dev) -> Tensor:
    n = C.c_size_t(0)
    call("gsb_specular_workspace_bytes", dev, C.c_int32(R), C.byref(n))
    return torch.empty(n.value + 256, dtype=torch.uint8, device=dev)     # the library aligns its carve to 256 B


def _check_cubemap(t: Tensor, channels: int, name: str) -> None:
    # same checks as CHECK_TENSOR in torch_bindings.cpp:27-31
    _require_cuda(t, f"render_utils: {name}")
    if t.dim() != 4 or t.shape[0] != 6 or t.shape[1] != t.shape[2] or t.shape[3] != channels:
        raise RuntimeError(f"{name} must have shape [6,R,R,{channels}], got {tuple(t.shape)}")


class render_utils:
    """Function-for-function stand-in for the pybind module `rfstudio_render_utils`."""

    @staticmethod
    def diffuse_cubemap_fwd(cubemap: Tensor) -> Tensor:
        _check_cubemap(cubemap, 3, "cubemap")
        c = f32c(cubemap)
        out = torch.empty_like(c)
        call("gsb_diffuse_cubemap_fwd", c.device, C.c_int32(c.shape[1]), ptr(c), ptr(out), C.c_int32(3),
             stream_ptr(c.device))
        return out

    @staticmethod
    def diffuse_cubemap_bwd(cubemap: Tensor, grad: Tensor) -> Tensor:
        _check_cubemap(cubemap, 3, "cubemap")
        _check_cubemap(grad, 3, "grad")
        g = f32c(grad)
        out = torch.empty_like(g)
        call("gsb_diffuse_cubemap_bwd", g.device, C.c_int32(g.shape[1]), ptr(g), C.c_int32(3), ptr(out),
             stream_ptr(g.device))
        return out

    @staticmethod
    def specular_bounds(resolution: int, costheta_cutoff: float, index: int) -> Tensor:
        dev = _bounds_device(index)
        out = torch.zeros(6, resolution, resolution, 24, dtype=torch.float32, device=dev)
        call("gsb_specular_bounds", dev, C.c_int32(resolution), C.c_float(costheta_cutoff), ptr(out), stream_ptr(dev))
        return out

    @staticmethod
    def specular_cubemap_fwd(cubemap: Tensor, bounds: Tensor, roughness: float, costheta_cutoff: float) -> Tensor:
        _check_cubemap(cubemap, 3, "cubemap")
        _check_cubemap(bounds, 24, "bounds")
        c, b = f32c(cubemap), f32c(bounds)
        R = c.shape[1]
        out = torch.empty(6, R, R, 4, dtype=torch.float32, device=c.device)
        ws = _spec_workspace(R, c.device)
        call("gsb_specular_cubemap_fwd", c.device, C.c_int32(R), ptr(c), ptr(b), C.c_float(roughness),
             C.c_float(costheta_cutoff), C.c_int32(0), ptr(out), ptr(ws), stream_ptr(c.device))
        return out

    @staticmethod
    def specular_cubemap_bwd(cubemap: Tensor, bounds: Tensor, grad: Tensor, roughness: float,
                             costheta_cutoff: float) -> Tensor:
        _check_cubemap(cubemap, 3, "cubemap")
        _check_cubemap(bounds, 24, "bounds")
        _check_cubemap(grad, 4, "grad")
        b, g = f32c(bounds), f32c(grad)
        R = g.shape[1]
        out = torch.empty(6, R, R, 3, dtype=torch.float32, device=g.device)
        ws = _spec_workspace(R, g.device)
        call("gsb_specular_cubemap_bwd", g.device, C.c_int32(R), ptr(b), ptr(g), None, C.c_float(roughness),
             C.c_float(costheta_cutoff), ptr(out), ptr(ws), stream_ptr(g.device))
        return out


# ---- the two prefilter passes as differentiable operators (reference API: _splitsum/_wrap.py:95-103, :139-157) -------
class _PrefilterPass(torch.autograd.Function):
    """One node for either plugin pass.  `lobe` is None for the cosine (diffuse) pass, else (roughness, cos_cut, bounds);
    the backward of both passes is a gather over the same footprint (csrc/prefilter.cu), which needs the cotangent only
    (the passes are linear in the texels)."""

    @staticmethod
    def forward(ctx, cubemap: Tensor, lobe):
        ctx.lobe = lobe
        ctx.save_for_backward(cubemap)
        if lobe is None:
            return render_utils.diffuse_cubemap_fwd(cubemap)
        roughness, cos_cut, bounds = lobe
        return render_utils.specular_cubemap_fwd(cubemap, bounds, roughness, cos_cut)

    @staticmethod
    def backward(ctx, v_out: Tensor):
        (cubemap,) = ctx.saved_tensors
        v_out = v_out.contiguous()
        if ctx.lobe is None:
            return render_utils.diffuse_cubemap_bwd(cubemap, v_out), None
        roughness, cos_cut, bounds = ctx.lobe
        return render_utils.specular_cubemap_bwd(cubemap, bounds, v_out, roughness, cos_cut), None


def _finite_under_anomaly_mode(out: Tensor, what: str) -> Tensor:
    if torch.is_anomaly_enabled() and not bool(torch.isfinite(out).all()):
        raise AssertionError(f"Output of {what} contains inf or NaN")
    return out


def diffuse_cubemap(cubemap: Tensor) -> Tensor:
    """Cosine-weighted irradiance of a cube map [6,R,R,3] (same call as the reference's `diffuse_cubemap`)."""
    _check_cubemap(cubemap, 3, "cubemap")
    return _finite_under_anomaly_mode(_PrefilterPass.apply(cubemap, None), "diffuse_cubemap")


_ndf_bounds_cache: Dict[Tuple[int, float, float, int], Tuple[float, Tensor]] = {}


def ndf_cutoff_costheta(roughness: float, cutoff: float, n_samples: int = 1000000) -> float:
    """cos(theta) below which the GGX lobe of `roughness` holds the fraction `cutoff` of its mass, on the reference's
    grid of 1e6 uniform angles in [0, pi/2] (_wrap.py:120-132).  The table of bounds and the kernels' cone test depend
    on this number bit for bit, so the density keeps the reference's float64 expression order."""
    cos_t = np.cos(np.linspace(0.0, 0.5 * np.pi, n_samples))
    a2 = roughness ** 4
    c = np.clip(cos_t, 0.0, 1.0)
    d = (c * a2 - c) * c + 1.0
    mass = np.cumsum(a2 / (d * d * np.pi))
    first = int(np.searchsorted(mass, mass[-1] * cutoff, side="left"))      # first index with mass >= cutoff * total
    return float(cos_t[first])


def ndf_bounds(res: int, roughness: float, cutoff: float, index) -> Tuple[float, Tensor]:
    """(cos(theta_cutoff), bounds[6,res,res,24]), computed once per (res, roughness, cutoff, device)."""
    key = (res, roughness, cutoff, index)
    hit = _ndf_bounds_cache.get(key)
    if hit is None:
        cos_cut = ndf_cutoff_costheta(roughness, cutoff)
        hit = _ndf_bounds_cache[key] = (cos_cut, render_utils.specular_bounds(res, cos_cut, index))
    return hit


def specular_cubemap(cubemap: Tensor, roughness: float, cutoff: float = 0.99) -> Tensor:
    """GGX-prefiltered cube map [6,R,R,3] for one roughness (same call as the reference's `specular_cubemap`): weighted
    sum / weight sum of the plugin pass."""
    _check_cubemap(cubemap, 3, "cubemap")
    cos_cut, bounds = ndf_bounds(cubemap.shape[1], roughness, cutoff, cubemap.device.index)
    acc = _finite_under_anomaly_mode(_PrefilterPass.apply(cubemap, (roughness, cos_cut, bounds)), "specular_cubemap")
    return acc[..., :3] / acc[..., 3:]


# ---- cached plans: the GGX weights depend on (resolution, roughness, cone), not on the cube map ---------------------------
class SpecularPlan:
    """Per-level weight table of the specular prefilter (csrc/prefilter.cu, "cached plan"): built once per (resolution,
    roughness, cutoff, device), then every training step streams it forward and backward instead of re-evaluating the
    GGX lobe at every tap.  ~1-2.5 GB per level of a 512^2 cube map, 5.5 GB for the default chain -- sized for a 180 GB
    part; `GSB_PREFILTER_PLAN_GB` (default 12) caps the total, levels past the cap keep the on-the-fly kernels."""

    __slots__ = ("seg_start", "segs", "weights", "nbytes")


_plans: Dict[Tuple[int, float, float, int], "SpecularPlan"] = {}
_plan_bytes: Dict[int, int] = {}


def specular_plan(res: int, roughness: float, cutoff: float, dev: torch.device):
    """The cached plan of a level, or None (resolution not a multiple of 8, or the memory cap is reached)."""
    import os
    index = dev.index if dev.index is not None else (torch.cuda.current_device() if dev.type == "cuda" else 0)
    key = (res, roughness, cutoff, index)
    if key in _plans:
        return _plans[key]
    plan = None
    if res % 8 == 0 and 8 <= res <= 1024:
        cos_cut, bounds = ndf_bounds(res, roughness, cutoff, dev.index)
        ws = _spec_workspace(res, dev)
        st = stream_ptr(dev)
        n_pf = 6 * res * res // 32 * 6
        counts = torch.empty(n_pf, 2, dtype=torch.int32, device=dev)
        call("gsb_specular_plan_count", dev, C.c_int32(res), ptr(bounds), C.c_float(cos_cut), ptr(counts), ptr(ws), st)
        zero = torch.zeros(2, 1, dtype=torch.int64, device=dev)
        # scan along the contiguous axis (torch's scan over the outer axis of an [n, 2] tensor takes milliseconds)
        starts = torch.cat((zero, counts.t().contiguous().to(torch.int64).cumsum(1)), dim=1).t()   # [n_pf + 1, 2]
        n_segs, n_taps = (int(v) for v in starts[-1].tolist())                  # one host read, once per level
        nbytes = n_taps * 128 + n_segs * 16 + starts.numel() * 4
        cap = float(os.environ.get("GSB_PREFILTER_PLAN_GB", "12")) * 2 ** 30
        if n_taps * 32 < 2 ** 31 and _plan_bytes.get(index, 0) + nbytes <= cap:
            plan = SpecularPlan()
            plan.seg_start = starts[:, 0].to(torch.int32).contiguous()
            tap_start = starts[:, 1].to(torch.int32).contiguous()
            plan.segs = torch.empty(max(n_segs, 1), 4, dtype=torch.int32, device=dev)
            plan.weights = torch.empty(max(n_taps, 1) * 32, dtype=torch.float32, device=dev)
            call("gsb_specular_plan_fill", dev, C.c_int32(res), ptr(bounds), C.c_float(roughness), C.c_float(cos_cut),
                 ptr(plan.seg_start), ptr(tap_start), ptr(plan.segs), ptr(plan.weights), ptr(ws), st)
            plan.nbytes = nbytes
            _plan_bytes[index] = _plan_bytes.get(index, 0) + nbytes
    _plans[key] = plan
    return plan


# ---- _texture.py:199-226 --------------------------------------------------------------------------------
class _CubeMapMip(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cubemap):
        assert cubemap.shape[0] == 6 and cubemap.shape[1] == cubemap.shape[2]
        c = f32c(cubemap)
        Ro = c.shape[1] // 2
        out = torch.empty(6, Ro, Ro, c.shape[3], dtype=torch.float32, device=c.device)
        assert c.shape[3] == 3
        call("gsb_cubemap_mip_fwd", c.device, C.c_int32(Ro), ptr(c), C.c_int32(3), ptr(out), C.c_int32(3),
             stream_ptr(c.device))
        return out

    @staticmethod
    def backward(ctx, dout):
        d = f32c(dout)
        Ro = d.shape[1]
        out = torch.empty(6, 2 * Ro, 2 * Ro, 3, dtype=torch.float32, device=d.device)
        call("gsb_cubemap_mip_bwd", d.device, C.c_int32(Ro), ptr(d), ptr(out), stream_ptr(d.device))
        return out


def cubemap_mip(cubemap: Tensor) -> Tensor:
    return _CubeMapMip.apply(cubemap)


def _roughness_schedule(n_levels: int, min_roughness: float, max_roughness: float) -> List[float]:
    # _texture.py:545-548
    r = [(idx / (n_levels - 2)) * (max_roughness - min_roughness) + min_roughness for idx in range(n_levels - 1)]
    return r + [1.0]


def merge_mipmaps(mipmaps: List[Tensor]) -> Tensor:
    """_texture.py:228-244 (pure indexing; the never-read tail of channel 3 is zero here, uninitialised there)."""
    R = mipmaps[0].shape[-2]
    res = mipmaps[0].new_zeros(6, 4, R, R)
    res[:, :3] = mipmaps[0].permute(0, 3, 1, 2)
    o = 0
    for i in range(1, len(mipmaps)):
        h = R // 2
        assert mipmaps[i].shape[1:3] == (h, h)
        res[:, 3, o:o + h, o:o + h] = mipmaps[i][..., 0]
        res[:, 3, o:o + h, o + h:o + R] = mipmaps[i][..., 1]
        res[:, 3, o + h:o + R, o:o + h] = mipmaps[i][..., 2]
        o += h
        R = h
    return res


def as_splitsum(cubemap: Tensor, *, cutoff: float = 0.99, min_resolution: int = 16, min_roughness: float = 0.08,
                max_roughness: float = 0.5):
    """TextureCubeMap.as_splitsum (_texture.py:530-557) operator by operator, returning the TextureSplitSum
    fields (base [6,Rm,Rm,3], mipmaps [6,4,R,R], num_mipmaps, min_roughness, max_roughness)."""
    mips = [cubemap]
    while mips[-1].shape[1] > min_resolution:
        mips.append(cubemap_mip(mips[-1]))
    assert len(mips) > 2, "Min resolution is too large."
    base = diffuse_cubemap(mips[-1])
    rough = _roughness_schedule(len(mips), min_roughness, max_roughness)
    spec = [specular_cubemap(m, roughness=r, cutoff=cutoff) for m, r in zip(mips, rough)]
    return base, merge_mipmaps(spec), len(mips), min_roughness, max_roughness


class _PrefilterStack(torch.autograd.Function):
    """cubemap [6,R,R,3] -> env stack [T,4] in one autograd node: mip chain, per-level GGX prefilter written
    normalised straight into the stack (texel.w keeps wsum for the backward), diffuse base."""

    @staticmethod
    def forward(ctx, cubemap, cutoff, min_resolution, min_roughness, max_roughness):
        c = f32c(cubemap)
        dev = c.device
        st = stream_ptr(dev)
        R0 = c.shape[1]
        chain = [c]
        while chain[-1].shape[1] > min_resolution:
            Ro = chain[-1].shape[1] // 2
            nxt = torch.empty(6, Ro, Ro, 3, dtype=torch.float32, device=dev)
            call("gsb_cubemap_mip_fwd", dev, C.c_int32(Ro), ptr(chain[-1]), C.c_int32(3), ptr(nxt), C.c_int32(3), st)
            chain.append(nxt)
        L = len(chain)
        assert L > 2, "Min resolution is too large."
        Rb = chain[-1].shape[1]
        T = EnvStack.texels(R0, L, Rb)
        stack = torch.empty(T, 4, dtype=torch.float32, device=dev)
        rough = _roughness_schedule(L, min_roughness, max_roughness)
        ws = _spec_workspace(R0, dev)   # sized for the finest level, reused by every level (stream-ordered)
        o = 0
        cts = []
        for l in range(L):
            r = R0 >> l
            ct, bounds = ndf_bounds(r, rough[l], cutoff, dev.index)
            cts.append(ct)
            plan = specular_plan(r, rough[l], cutoff, dev)
            if plan is not None:     # stream the cached weights instead of evaluating the lobe at every tap
                call("gsb_specular_plan_fwd", dev, C.c_int32(r), ptr(chain[l]), ptr(plan.seg_start), ptr(plan.segs),
                     ptr(plan.weights), C.c_int32(1), C.c_void_p(stack.data_ptr() + o * 16), ptr(ws), st)
            else:
                call("gsb_specular_cubemap_fwd", dev, C.c_int32(r), ptr(chain[l]), ptr(bounds), C.c_float(rough[l]),
                     C.c_float(ct), C.c_int32(1), C.c_void_p(stack.data_ptr() + o * 16), ptr(ws), st)
            o += 6 * r * r
        stack[o:].zero_()
        call("gsb_diffuse_cubemap_fwd", dev, C.c_int32(Rb), ptr(chain[-1]), C.c_void_p(stack.data_ptr() + o * 16),
             C.c_int32(4), st)
        ctx.save_for_backward(stack)
        ctx.meta = (R0, L, Rb, rough, cts, cutoff)
        return stack

    @staticmethod
    def backward(ctx, v_stack):
        stack, = ctx.saved_tensors
        R0, L, Rb, rough, cts, cutoff = ctx.meta
        v = f32c(v_stack)
        dev = v.device
        st = stream_ptr(dev)
        offs, o = [], 0
        for l in range(L):
            offs.append(o)
            o += 6 * (R0 >> l) ** 2
        # gradient w.r.t. each chain level from its own prefilter
        ws = _spec_workspace(R0, dev)
        g_levels = []
        for l in range(L):
            r = R0 >> l
            _, bounds = ndf_bounds(r, rough[l], cutoff, dev.index)
            g = torch.empty(6, r, r, 3, dtype=torch.float32, device=dev)
            plan = specular_plan(r, rough[l], cutoff, dev)
            if plan is not None:
                call("gsb_specular_plan_bwd", dev, C.c_int32(r), ptr(plan.seg_start), ptr(plan.segs), ptr(plan.weights),
                     C.c_void_p(v.data_ptr() + offs[l] * 16), C.c_void_p(stack.data_ptr() + offs[l] * 16), ptr(g),
                     ptr(ws), st)
            else:
                call("gsb_specular_cubemap_bwd", dev, C.c_int32(r), ptr(bounds), C.c_void_p(v.data_ptr() + offs[l] * 16),
                     C.c_void_p(stack.data_ptr() + offs[l] * 16), C.c_float(rough[l]), C.c_float(cts[l]), ptr(g), ptr(ws),
                     st)
            g_levels.append(g)
        gb = torch.empty(6, Rb, Rb, 3, dtype=torch.float32, device=dev)
        call("gsb_diffuse_cubemap_bwd", dev, C.c_int32(Rb), C.c_void_p(v.data_ptr() + o * 16), C.c_int32(4), ptr(gb), st)
        g_levels[-1] += gb
        # back through the mip chain (coarse -> fine)
        for l in range(L - 1, 0, -1):
            Ro = R0 >> l
            up = torch.empty(6, 2 * Ro, 2 * Ro, 3, dtype=torch.float32, device=dev)
            call("gsb_cubemap_mip_bwd", dev, C.c_int32(Ro), ptr(g_levels[l]), ptr(up), st)
            g_levels[l - 1] += up
        return g_levels[0], None, None, None, None


def as_envstack(cubemap: Tensor, *, cutoff: float = 0.99, min_resolution: int = 16, min_roughness: float = 0.08,
                max_roughness: float = 0.5) -> EnvStack:
    """Same result as as_splitsum, produced directly in the native env-stack layout (no quad-tree pack, no
    per-view unpack): the fast path GeoSplatter.get_envmap -> RenderableAttrs.splat uses."""
    _require_cuda(cubemap, "as_envstack")
    data = _PrefilterStack.apply(cubemap, cutoff, min_resolution, min_roughness, max_roughness)
    R0 = cubemap.shape[1]
    L = 1
    while (R0 >> (L - 1)) > min_resolution:
        L += 1
    return EnvStack(data, R0, L, R0 >> (L - 1), min_roughness, max_roughness)
