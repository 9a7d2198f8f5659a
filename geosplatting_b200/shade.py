"""Host side of the split-sum shade and of the texture drop-ins (C ABI: gsb_shade_*, gsb_envstack_*,
gsb_texture_*).

Mirrors, for this path only:
    RenderableAttrs.splat shade block      rfstudio/model/geosplat.py:83-121       -> shade()
    TextureSplitSum.sample                 rfstudio/graphics/_mesh/_texture.py:571-613 -> splitsum_sample()
    nvdiffrast.torch.texture               geosplat.py:93, _texture.py:220,:596,:604   -> texture()
    _get_fg_lut                            rfstudio/graphics/shaders.py:22-26          -> load_fg_lut()
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from ._lib import call, f32c, ptr, stream_ptr
from ._lib import require_cuda as _require_cuda

MODES = {"pbr": 0, "diffuse": 1, "specular": 2}


class EnvStack:
    """Env map in this library's native layout: float4 texels, specular levels 0..L-1 then the diffuse base
    (see include/geosplat_b200.h).  `data` is a [T,4] tensor that may carry autograd history."""

    def __init__(self, data: Tensor, R0: int, L: int, Rb: int, min_roughness: float = 0.08,
                 max_roughness: float = 0.5):
        self.data, self.R0, self.L, self.Rb = data, int(R0), int(L), int(Rb)
        self.min_roughness, self.max_roughness = float(min_roughness), float(max_roughness)

    @staticmethod
    def texels(R0: int, L: int, Rb: int) -> int:
        n = C.c_int64(0)
        call("gsb_envstack_texels", None, C.c_int32(R0), C.c_int32(L), C.c_int32(Rb), C.byref(n))
        return int(n.value)

    @classmethod
    def from_splitsum(cls, base: Tensor, mipmaps: Tensor, num_mipmaps: int, min_roughness: float = 0.08,
                      max_roughness: float = 0.5) -> "EnvStack":
        """From the reference's TextureSplitSum fields (base [6,Rb,Rb,3], mipmaps [6,4,R0,R0] quad-tree pack,
        _texture.py:559-569).  Differentiable w.r.t. both."""
        data = _EnvPack.apply(mipmaps, base, int(num_mipmaps))
        return cls(data, mipmaps.shape[-1], int(num_mipmaps), base.shape[1], min_roughness, max_roughness)

    @classmethod
    def coerce(cls, envmap) -> "EnvStack":
        """EnvStack as is; an object with the fields of rfstudio's `TextureSplitSum` (_texture.py:559-569: base, mipmaps,
        num_mipmaps, min_roughness, max_roughness) is packed into the native layout (differentiable)."""
        if isinstance(envmap, cls):
            return envmap
        if all(hasattr(envmap, k) for k in ("base", "mipmaps", "num_mipmaps")):
            if getattr(envmap, "transform", None) is not None:
                raise NotImplementedError("TextureSplitSum.transform is not used by stage 1 (geosplat.py:784)")
            f = lambda v, d: d if v is None else (float(v.item()) if hasattr(v, "item") else float(v))   # noqa: E731
            n = envmap.num_mipmaps
            return cls.from_splitsum(envmap.base, envmap.mipmaps, int(n.item()) if hasattr(n, "item") else int(n),
                                     f(getattr(envmap, "min_roughness", None), 0.08),
                                     f(getattr(envmap, "max_roughness", None), 0.5))
        raise TypeError(f"envmap: expected EnvStack or a TextureSplitSum-like object, got {type(envmap).__name__}")

    def level_views(self) -> List[Tensor]:
        """Per-level [6,R,R,4] views (spec levels, then base)."""
        out, o = [], 0
        for l in range(self.L):
            r = self.R0 >> l
            out.append(self.data[o:o + 6 * r * r].view(6, r, r, 4))
            o += 6 * r * r
        out.append(self.data[o:o + 6 * self.Rb * self.Rb].view(6, self.Rb, self.Rb, 4))
        return out


class _EnvPack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, packed: Tensor, base: Tensor, L: int):
        packed_c, base_c = f32c(packed), f32c(base)
        dev = packed_c.device
        R0, Rb = packed_c.shape[-1], base_c.shape[1]
        assert packed_c.shape == (6, 4, R0, R0) and base_c.shape == (6, Rb, Rb, 3)
        T = EnvStack.texels(R0, L, Rb)
        stack = torch.empty(T, 4, dtype=torch.float32, device=dev)
        call("gsb_envstack_pack", dev, C.c_int32(R0), C.c_int32(L), C.c_int32(Rb), ptr(packed_c), ptr(base_c),
             ptr(stack), stream_ptr(dev))
        ctx.dims = (R0, L, Rb)
        return stack

    @staticmethod
    def backward(ctx, v_stack):
        R0, L, Rb = ctx.dims
        v_stack = f32c(v_stack)
        dev = v_stack.device
        v_packed = torch.zeros(6, 4, R0, R0, dtype=torch.float32, device=dev)
        v_base = torch.empty(6, Rb, Rb, 3, dtype=torch.float32, device=dev)
        call("gsb_envstack_unpack_grad", dev, C.c_int32(R0), C.c_int32(L), C.c_int32(Rb), ptr(v_stack),
             ptr(v_packed), ptr(v_base), stream_ptr(dev))
        return v_packed, v_base, None


_shade_ws: dict = {}


def shade_workspace(dev: torch.device, R0: int, L: int, Rb: int) -> Tensor:
    """Scratch for gsb_shade_bwd's private copies of the coarse env levels (zeroed by the call itself, so one
    buffer per device, stream and stack shape is shared by every view on that stream)."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(dev).cuda_stream, R0, L, Rb)
    ws = _shade_ws.get(key)
    if ws is None:
        n = C.c_size_t(0)
        call("gsb_shade_workspace_bytes", None, C.c_int32(R0), C.c_int32(L), C.c_int32(Rb), C.byref(n))
        ws = _shade_ws[key] = torch.empty(max(int(n.value), 16), dtype=torch.uint8, device=dev)
    return ws


class _Shade(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, normals, kd, ks, stack, cam_pos, lut, meta):
        means_c, normals_c, kd_c, ks_c, stack_c = f32c(means), f32c(normals), f32c(kd), f32c(ks), f32c(stack)
        dev = means_c.device
        N = means_c.shape[0]
        colors = torch.empty(N, 3, dtype=torch.float32, device=dev)
        cam = (C.c_float * 3)(*cam_pos)
        R0, L, Rb, min_r, max_m, emin, emax, mode = meta
        call("gsb_shade_fwd", dev, C.c_int32(N), ptr(means_c), ptr(normals_c), ptr(kd_c), ptr(ks_c), cam, ptr(lut),
             C.c_int32(lut.shape[0]), ptr(stack_c), C.c_int32(R0), C.c_int32(L), C.c_int32(Rb), C.c_float(min_r),
             C.c_float(max_m), C.c_float(emin), C.c_float(emax), C.c_int32(mode), ptr(colors), stream_ptr(dev))
        ctx.save_for_backward(means_c, normals_c, kd_c, ks_c, stack_c, lut)
        ctx.cam, ctx.meta = cam, meta
        return colors

    @staticmethod
    def backward(ctx, v_colors):
        means, normals, kd, ks, stack, lut = ctx.saved_tensors
        dev = means.device
        N = means.shape[0]
        R0, L, Rb, min_r, max_m, emin, emax, mode = ctx.meta
        v_colors = f32c(v_colors)
        v_means = torch.empty_like(means)
        v_normals = torch.empty_like(normals)
        v_kd = torch.empty_like(kd)
        v_ks = torch.empty_like(ks)
        v_stack = torch.zeros_like(stack)
        ws = shade_workspace(dev, R0, L, Rb)
        call("gsb_shade_bwd", dev, C.c_int32(N), ptr(means), ptr(normals), ptr(kd), ptr(ks), ctx.cam, ptr(lut),
             C.c_int32(lut.shape[0]), ptr(stack), C.c_int32(R0), C.c_int32(L), C.c_int32(Rb), C.c_float(min_r),
             C.c_float(max_m), C.c_float(emin), C.c_float(emax), C.c_int32(mode), ptr(v_colors), ptr(v_means),
             ptr(v_normals), ptr(v_kd), ptr(v_ks), ptr(v_stack), ptr(ws), C.c_size_t(ws.numel()), C.c_int32(0),
             stream_ptr(dev))
        return v_means, v_normals, v_kd, v_ks, v_stack, None, None, None


def shade(means: Tensor, normals: Tensor, kd: Tensor, ks: Tensor, camera_pos: Sequence[float], env: EnvStack,
          fg_lut: Tensor, *, min_roughness: float, max_metallic: float, mode: str = "pbr") -> Tensor:
    """Per-Gaussian colours [N,3] (geosplat.py:83-121 with culling=False).  `camera_pos` is c2w[:, 3] as three
    host floats; `fg_lut` is the [R,R,2] (or [1,R,R,2]) DFG table on the device."""
    if mode not in MODES:
        raise ValueError(mode)
    _require_cuda(means, "shade")
    lut = f32c(fg_lut).reshape(fg_lut.shape[-3], fg_lut.shape[-2], 2)
    cam = tuple(float(x) for x in camera_pos)
    meta = (env.R0, env.L, env.Rb, float(min_roughness), float(max_metallic), env.min_roughness, env.max_roughness,
            MODES[mode])
    return _Shade.apply(means, normals, kd, ks, env.data, cam, lut, meta)


# ------------------------------------------------------------------------------------------------------
# nvdiffrast.torch.texture drop-in (the three call shapes on the path)
# ------------------------------------------------------------------------------------------------------
def _ptr_array(tensors: Sequence[Optional[Tensor]]):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


class _TexCube(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dirs, level, n_levels, *levels):
        dirs_c = f32c(dirs)
        lv = [f32c(t) for t in levels]
        level_c = None if level is None else f32c(level)
        dev = dirs_c.device
        N, Cn, R0 = dirs_c.shape[0], lv[0].shape[-1], lv[0].shape[1]
        out = torch.empty(N, Cn, dtype=torch.float32, device=dev)
        call("gsb_texture_cube_fwd", dev, C.c_int32(N), C.c_int32(Cn), C.c_int32(n_levels), _ptr_array(lv),
             C.c_int32(R0), ptr(dirs_c), ptr(level_c), ptr(out), stream_ptr(dev))
        ctx.save_for_backward(dirs_c, *( [level_c] if level_c is not None else [] ), *lv)
        ctx.has_level = level_c is not None
        ctx.n_levels = n_levels
        return out

    @staticmethod
    def backward(ctx, v_out):
        saved = ctx.saved_tensors
        dirs = saved[0]
        level = saved[1] if ctx.has_level else None
        lv = list(saved[2:] if ctx.has_level else saved[1:])
        dev = dirs.device
        N, Cn, R0 = dirs.shape[0], lv[0].shape[-1], lv[0].shape[1]
        v_out = f32c(v_out)
        need_tex = [ctx.needs_input_grad[3 + i] for i in range(len(lv))]
        v_lv = [torch.zeros_like(t) if need else None for t, need in zip(lv, need_tex)]
        v_dirs = torch.empty_like(dirs)
        v_level = torch.empty(N, dtype=torch.float32, device=dev) if level is not None else None
        call("gsb_texture_cube_bwd", dev, C.c_int32(N), C.c_int32(Cn), C.c_int32(ctx.n_levels), _ptr_array(lv),
             C.c_int32(R0), ptr(dirs), ptr(level), ptr(v_out), _ptr_array(v_lv), ptr(v_dirs), ptr(v_level),
             stream_ptr(dev))
        return (v_dirs, v_level, None, *v_lv)


class _Tex2D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tex, uv):
        tex_c, uv_c = f32c(tex), f32c(uv)
        dev = uv_c.device
        N = uv_c.shape[0]
        out = torch.empty(N, 2, dtype=torch.float32, device=dev)
        call("gsb_texture2d_fwd", dev, C.c_int32(N), C.c_int32(tex_c.shape[1]), C.c_int32(tex_c.shape[0]), ptr(tex_c),
             ptr(uv_c), ptr(out), stream_ptr(dev))
        ctx.save_for_backward(tex_c, uv_c)
        return out

    @staticmethod
    def backward(ctx, v_out):
        tex, uv = ctx.saved_tensors
        dev = uv.device
        v_uv = torch.empty_like(uv)
        call("gsb_texture2d_bwd", dev, C.c_int32(uv.shape[0]), C.c_int32(tex.shape[1]), C.c_int32(tex.shape[0]),
             ptr(tex), ptr(uv), ptr(f32c(v_out)), ptr(v_uv), stream_ptr(dev))
        return None, v_uv


def texture(tex: Tensor, uv: Tensor, uv_da=None, mip_level_bias: Optional[Tensor] = None, mip=None,
            filter_mode: str = "auto", boundary_mode: str = "wrap", max_mip_level=None) -> Tensor:
    """`nvdiffrast.torch.texture` for the call shapes GeoSplatting's hot path uses; anything else raises.

    * tex [1,H,W,2], uv [...,2], filter 'linear', boundary 'clamp'                  -> [...,2]
    * tex [1,6,R,R,C], uv [...,3], filter 'linear', boundary 'cube'                 -> [...,C]
    * same + mip=[[1,6,R/2,R/2,C],...], mip_level_bias [...], 'linear-mipmap-linear' -> [...,C]
    """
    if uv_da is not None or max_mip_level is not None:
        raise NotImplementedError("geosplatting_b200.texture: uv_da / max_mip_level are not used on this path")
    _require_cuda(uv, "texture")
    lead = uv.shape[:-1]
    if boundary_mode == "clamp":
        if filter_mode != "linear" or tex.dim() != 4 or tex.shape[0] != 1 or tex.shape[-1] != 2:
            raise NotImplementedError("geosplatting_b200.texture: 2D mode supports [1,H,W,2], filter 'linear'")
        if tex.requires_grad:
            raise NotImplementedError("geosplatting_b200.texture: no gradient w.r.t. a 2D texture (the FG LUT is constant)")
        return _Tex2D.apply(tex[0], uv.reshape(-1, 2)).reshape(*lead, 2)
    if boundary_mode != "cube" or tex.dim() != 5 or tex.shape[:2] != (1, 6):
        raise NotImplementedError(f"geosplatting_b200.texture: unsupported boundary_mode/shape {boundary_mode} {tuple(tex.shape)}")
    d = uv.reshape(-1, 3)
    if filter_mode == "linear":
        out = _TexCube.apply(d, None, 1, tex[0])
    elif filter_mode == "linear-mipmap-linear":
        if mip is None or mip_level_bias is None:
            raise NotImplementedError("geosplatting_b200.texture: mip mode needs an explicit mip stack and mip_level_bias")
        levels = [tex[0]] + [m[0] for m in mip]
        out = _TexCube.apply(d, mip_level_bias.reshape(-1), len(levels), *levels)
    else:
        raise NotImplementedError(filter_mode)
    return out.reshape(*lead, tex.shape[-1])


def splitsum_sample(env: EnvStack, normals: Tensor, directions: Tensor, roughness: Tensor) -> Tuple[Tensor, Tensor]:
    """TextureSplitSum.sample (transform=None): (l_diffuse[...,3], l_specular[...,3]) for arbitrary normals /
    directions / roughness[...,1] (_texture.py:571-613), on the native stack (no per-call mip re-materialisation)."""
    L = env.L
    rough = roughness[..., 0]
    lo = ((rough - env.min_roughness) / (env.max_roughness - env.min_roughness)).clamp(0, 1) * (L - 2)
    hi = ((rough - env.max_roughness) / (1.0 - env.max_roughness)).clamp(0, 1) + L - 2
    level = torch.where(rough < env.max_roughness, lo, hi)
    views = env.level_views()
    lead = normals.shape[:-1]
    l_diff = _TexCube.apply(normals.reshape(-1, 3), None, 1, views[-1])[:, :3]
    l_spec = _TexCube.apply(directions.reshape(-1, 3), level.reshape(-1), L, *views[:-1])[:, :3]
    return l_diff.reshape(*lead, 3), l_spec.reshape(*lead, 3)


def load_fg_lut(path: str, device, resolution: int = 256) -> Tensor:
    """The reference's DFG table asset (rfstudio/assets/geometry/pbr/bsdf_256_256.bin, 256x256x2 fp32,
    loaded at rfstudio/graphics/shaders.py:22-26) -> [R,R,2] on `device`."""
    lut = np.fromfile(path, dtype=np.float32)
    if lut.size != resolution * resolution * 2:
        raise ValueError(f"{path}: expected {resolution * resolution * 2} floats, found {lut.size}")
    return torch.from_numpy(lut).to(device).view(resolution, resolution, 2)


_default_luts: dict = {}


def fg_lut_asset_path() -> Optional[str]:
    """Where the DFG table lives: $GSB_FG_LUT, else the asset inside an importable `rfstudio` package
    (rfstudio/assets/geometry/pbr/bsdf_256_256.bin, what shaders.py:22-26 reads)."""
    env = os.environ.get("GSB_FG_LUT")
    if env:
        return env
    try:
        import importlib.util
        spec = importlib.util.find_spec("rfstudio")
        if spec is not None and spec.submodule_search_locations:
            cand = os.path.join(list(spec.submodule_search_locations)[0], "assets", "geometry", "pbr", "bsdf_256_256.bin")
            if os.path.exists(cand):
                return cand
    except Exception:
        pass
    return None


def get_fg_lut(resolution: int = 256, device="cuda") -> Tensor:
    """`_get_fg_lut(resolution, device)` (rfstudio/graphics/shaders.py:22-26), cached per device: [R,R,2].  Raises when
    the asset cannot be found -- there is no silent substitute for the reference's table."""
    key = (resolution, str(device))
    hit = _default_luts.get(key)
    if hit is None:
        path = fg_lut_asset_path()
        if path is None:
            raise RuntimeError("geosplatting_b200: no DFG table. Pass fg_lut=, or set GSB_FG_LUT to the reference's "
                               "rfstudio/assets/geometry/pbr/bsdf_256_256.bin (or make `rfstudio` importable)")
        hit = _default_luts[key] = load_fg_lut(path, device, resolution)
    return hit


def synthetic_fg_lut(device, resolution: int = 256) -> Tensor:
    """Analytic split-sum DFG approximation (Karis 2014 mobile fit) used by benches/tests when the
    reference asset is not on the machine: same shape/range as the real table, smooth, deterministic."""
    r = (torch.arange(resolution, dtype=torch.float32) + 0.5) / resolution
    rough = r[:, None].expand(resolution, resolution)
    ndv = r[None, :].expand(resolution, resolution)
    c0 = torch.tensor([-1.0, -0.0275, -0.572, 0.022])
    c1 = torch.tensor([1.0, 0.0425, 1.04, -0.04])
    rr = rough[..., None] * c0 + c1
    a004 = torch.minimum(rr[..., 0] * rr[..., 0], torch.exp2(-9.28 * ndv)) * rr[..., 0] + rr[..., 1]
    A = -1.04 * a004 + rr[..., 2]
    B = 1.04 * a004 + rr[..., 3]
    return torch.stack((A, B), dim=-1).clamp(0, 1).contiguous().to(device)
