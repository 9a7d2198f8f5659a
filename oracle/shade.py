"""Differentiable torch restatement of the per-Gaussian split-sum shade and its neighbours.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows, line by line:
    rfstudio/model/geosplat.py:83-121     RenderableAttrs.splat shade block (modes pbr/diffuse/specular)
    rfstudio/graphics/_mesh/_texture.py:571-613  TextureSplitSum.sample (mip level formula :584-594)
    rfstudio/graphics/math.py:119-128     safe_normalize
    rfstudio/model/geosplat.py:474-476    _tone_mapping_naive
with `dr.texture` supplied by oracle/texture.py.  It is validated against the reference's own Python code
(executed in the build container with shims, scripts/make_golden.py) through tests/golden/shade_*.npz.
"""
from __future__ import annotations

from typing import Sequence

import torch
from torch import Tensor

from . import texture as T


def safe_normalize(v: Tensor) -> Tensor:
    """rfstudio/graphics/math.py:119-128."""
    lengths = v.norm(dim=-1, keepdim=True)
    fallback = torch.tensor([0.0, 0.0, 1.0], dtype=v.dtype).expand_as(v)
    return torch.where(lengths < 1e-6, fallback, v / lengths.clamp_min(1e-6))


def mip_level(roughness: Tensor, num_mipmaps: int, min_roughness: float = 0.08, max_roughness: float = 0.5) -> Tensor:
    """_texture.py:584-594."""
    lo = ((roughness - min_roughness) / (max_roughness - min_roughness)).clamp(0, 1) * (num_mipmaps - 2)
    hi = ((roughness - max_roughness) / (1.0 - max_roughness)).clamp(0, 1) + num_mipmaps - 2
    return torch.where(roughness < max_roughness, lo, hi)


def splitsum_sample(base: Tensor, mips: Sequence[Tensor], normals: Tensor, directions: Tensor, roughness: Tensor,
                    min_roughness: float = 0.08, max_roughness: float = 0.5):
    """TextureSplitSum.sample with transform=None: (l_diff[N,3], l_spec[N,3])."""
    level = mip_level(roughness, len(mips), min_roughness, max_roughness)
    l_diff = T.texture_cube_linear(base, normals)
    l_spec = T.texture_cube_mip(mips, directions, level)
    return l_diff, l_spec


def shade(means: Tensor, normals: Tensor, kd: Tensor, ks: Tensor, cam_pos: Tensor, fg_lut: Tensor, base: Tensor,
          mips: Sequence[Tensor], *, min_roughness: float = 0.1, max_metallic: float = 1.0, mode: str = "pbr",
          env_min_roughness: float = 0.08, env_max_roughness: float = 0.5) -> Tensor:
    """geosplat.py:83-121 (culling=False): colours [N,3].  fg_lut is [256,256,2] (row = roughness)."""
    roughness = ks[:, 0:1] * (1 - min_roughness) + min_roughness
    metallic = ks[:, 1:2] * max_metallic
    specular = (1.0 - metallic) * 0.04 + kd * metallic
    diffuse = kd * (1.0 - metallic)
    wo = safe_normalize(cam_pos - means)
    n_dot_v = (normals * wo).sum(-1, keepdim=True).clamp(min=1e-6)
    fg = T.texture_2d_linear_clamp(fg_lut, torch.cat((n_dot_v, roughness), dim=-1))
    inv_wi = 2 * (wo * normals).sum(-1, keepdim=True) * normals - wo
    l_diff, l_spec = splitsum_sample(base, mips, normals, inv_wi, roughness[:, 0], env_min_roughness,
                                     env_max_roughness)
    reflectance = specular * fg[:, 0:1] + fg[:, 1:2]
    if mode == "pbr":
        return diffuse + l_spec * reflectance
    if mode == "diffuse":
        return l_diff * diffuse
    if mode == "specular":
        return l_spec * reflectance
    raise ValueError(mode)


def tone_map_naive(rgba: Tensor, exposure: Tensor) -> Tensor:
    """geosplat.py:474-476: 1 - softplus_{beta=100}(1 - rgb*exposure), alpha passthrough."""
    rgb = rgba[..., :3] * exposure
    return torch.cat((1 - torch.nn.functional.softplus(1 - rgb, beta=100.0), rgba[..., 3:]), dim=-1)


def cubemap_mip_fwd(cubemap: Tensor) -> Tensor:
    """_CubeMapMip.forward (_texture.py:201-206): 2x2 box filter per face."""
    x = cubemap.permute(0, 3, 1, 2)
    x = torch.nn.functional.avg_pool2d(x, (2, 2))
    return x.permute(0, 2, 3, 1).contiguous()


def cube_texel_dirs(res: int, dtype=torch.float32) -> Tensor:
    """Unit directions of the texel centres [6,res,res,3] (_texture.py:212-219)."""
    lin = torch.linspace(-1.0 + 1.0 / res, 1.0 - 1.0 / res, res, dtype=dtype)
    gy, gx = torch.meshgrid(lin, lin, indexing="ij")
    faces = []
    for s in range(6):
        v = T._face_point(torch.full(gx.shape, s, dtype=torch.long), gx, gy)
        faces.append(v / v.norm(dim=-1, keepdim=True))
    return torch.stack(faces)


def cubemap_mip_bwd(dout: Tensor) -> Tensor:
    """_CubeMapMip.backward (_texture.py:208-226): NOT the transpose of the box filter -- a bilinear cube
    resample of 0.25*dout at the fine texel directions."""
    res = dout.shape[1] * 2
    dirs = cube_texel_dirs(res, dout.dtype)
    out = T.texture_cube_linear(dout * 0.25, dirs.reshape(-1, 3))
    return out.reshape(6, res, res, dout.shape[-1])
