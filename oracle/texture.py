"""Differentiable torch restatement of `nvdiffrast.torch.texture` for the three modes on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: nvdiffrast is installed unpinned
from git HEAD by the reference (README.md:36) and is neither vendored nor installable here; this file
restates its published algorithm (texture.cu: indexCubeMap / indexTextureLinear / wrapCubeMap /
fetchQuad, summarised in SURVEY.md Appendix D) and is anchored on the reference's call sites:
    rfstudio/model/geosplat.py:93-98          2D, linear, clamp        (FG LUT)
    rfstudio/graphics/_mesh/_texture.py:596   cube, linear             (diffuse irradiance `base`)
    rfstudio/graphics/_mesh/_texture.py:604   cube, linear-mipmap-linear with explicit mips + level bias
    rfstudio/graphics/_mesh/_texture.py:220   cube, linear             (_CubeMapMip.backward)
The face / axis convention is cross-checked against the reference's own inverse map `_cube_to_dir`
(_texture.py:178-197) in tests/test_texture_cpu.py.  Gradients come from autograd over these ops.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
from torch import Tensor


def texture_2d_linear_clamp(tex: Tensor, uv: Tensor) -> Tensor:
    """tex[H,W,C], uv[N,2] (u -> column, v -> row) -> [N,C]; clamp to edge-texel centres."""
    H, W, _ = tex.shape
    u = uv[:, 0] * W - 0.5
    v = uv[:, 1] * H - 0.5
    u = u.clamp(0.0, W - 1.0)
    v = v.clamp(0.0, H - 1.0)
    iu0 = u.detach().floor().long()
    iv0 = v.detach().floor().long()
    clamp_u = (u.detach() == 0.0) | (u.detach() == W - 1.0)
    clamp_v = (v.detach() == 0.0) | (v.detach() == H - 1.0)
    iu1 = torch.where(clamp_u, iu0, iu0 + 1)
    iv1 = torch.where(clamp_v, iv0, iv0 + 1)
    fu = (u - iu0)[:, None]
    fv = (v - iv0)[:, None]
    # zero uv-gradient where clamped (the tap pair collapses onto one texel)
    t00, t10 = tex[iv0, iu0], tex[iv0, iu1]
    t01, t11 = tex[iv1, iu0], tex[iv1, iu1]
    top = t00 + (t10 - t00) * fu
    bot = t01 + (t11 - t01) * fu
    return top + (bot - top) * fv


def cube_face_uv(d: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Direction [N,3] (any length) -> (face[N] long, u[N], v[N]) with u,v in [0,1].

    Face = arg-max |component| with ties resolved towards x, then y (nvdiffrast indexCubeMap)."""
    x, y, z = d.unbind(-1)
    ax, ay, az = x.abs(), y.abs(), z.abs()
    is_z = az > torch.maximum(ax, ay)
    is_y = (~is_z) & (ay > ax)
    is_x = ~(is_z | is_y)
    c = torch.where(is_z, z, torch.where(is_y, y, x))
    neg = c < 0
    face = torch.where(is_z, 4, torch.where(is_y, 2, 0)) + neg.long()
    m = 0.5 / c.abs()
    # in-plane coordinates per face (matches _cube_to_dir inverse, _texture.py:178-197)
    sx = torch.where(is_x, z, x)              # x-slot: z for +-x faces, else x
    sy = torch.where(is_y, z, y)              # y-slot: z for +-y faces, else y
    m0 = torch.where((face == 0) | (face == 5), -m, m)
    m1 = torch.where(face == 2, m, -m)
    u = (sx * m0 + 0.5).clamp(0.0, 1.0)
    v = (sy * m1 + 0.5).clamp(0.0, 1.0)
    return face, u, v


def _face_point(face: Tensor, gx: Tensor, gy: Tensor) -> Tensor:
    """Un-normalised point on (the extended plane of) a cube face; gx, gy in face units ([-1,1] inside)."""
    one = torch.ones_like(gx)
    pts = [
        torch.stack((one, -gy, -gx), -1), torch.stack((-one, -gy, gx), -1),
        torch.stack((gx, one, gy), -1), torch.stack((gx, -one, -gy), -1),
        torch.stack((gx, -gy, one), -1), torch.stack((-gx, -gy, -one), -1),
    ]
    out = torch.zeros_like(pts[0])
    for f in range(6):
        out = torch.where((face == f)[..., None], pts[f], out)
    return out


def wrap_cube_texel(face: Tensor, iu: Tensor, iv: Tensor, R: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Integer texel (face, iu, iv) with iu or iv possibly in {-1, R} -> (face', iu', iv', valid).

    A tap that leaves the face across one edge lands on the texel of the adjacent face that touches it
    across that edge (seamless cube filtering); a tap that leaves across two edges at once (cube corner)
    has no texel: valid=False."""
    out_u = (iu < 0) | (iu >= R)
    out_v = (iv < 0) | (iv >= R)
    valid = ~(out_u & out_v)
    gx = (2.0 * (iu.double() + 0.5) / R - 1.0)
    gy = (2.0 * (iv.double() + 0.5) / R - 1.0)
    # fold the overshoot over the cube edge: the in-plane coordinate saturates at +-1 and the overshoot
    # is taken out of the major axis
    ex = (gx.abs() - 1.0).clamp_min(0.0)
    ey = (gy.abs() - 1.0).clamp_min(0.0)
    p = _face_point(face, gx.clamp(-1.0, 1.0), gy.clamp(-1.0, 1.0))
    major = _face_point(face, torch.zeros_like(gx), torch.zeros_like(gy))  # +-unit along the face axis
    p = p - major * (ex + ey)[..., None]
    f2, u2, v2 = cube_face_uv(p)
    iu2 = (u2 * R).floor().long().clamp(0, R - 1)
    iv2 = (v2 * R).floor().long().clamp(0, R - 1)
    inside = ~(out_u | out_v)
    f2 = torch.where(inside, face, f2)
    iu2 = torch.where(inside, iu, iu2)
    iv2 = torch.where(inside, iv, iv2)
    return f2, iu2, iv2, valid


def texture_cube_linear(tex: Tensor, d: Tensor) -> Tensor:
    """tex[6,R,R,C], directions d[N,3] -> [N,C]; bilinear with cross-face wrap, 3-texel corners."""
    R = tex.shape[1]
    face, u, v = cube_face_uv(d)
    u = u * R - 0.5
    v = v * R - 0.5
    iu0 = u.detach().floor().long()
    iv0 = v.detach().floor().long()
    fu = (u - iu0)[:, None]
    fv = (v - iv0)[:, None]
    taps, valids = [], []
    for div, diu in ((0, 0), (0, 1), (1, 0), (1, 1)):  # a00 a10 a01 a11  (first index = u)
        f2, iu2, iv2, ok = wrap_cube_texel(face, iu0 + diu, iv0 + div, R)
        taps.append(tex[f2, iv2, iu2] * ok[:, None])
        valids.append(ok)
    n_valid = sum(v_.to(tex.dtype) for v_ in valids)[:, None]
    corner = n_valid < 3.5
    avg = (taps[0] + taps[1] + taps[2] + taps[3]) * 0.33333333
    taps = [torch.where(corner & (~ok)[:, None], avg, t) for t, ok in zip(taps, valids)]
    a00, a10, a01, a11 = taps
    top = a00 + (a10 - a00) * fu
    bot = a01 + (a11 - a01) * fu
    return top + (bot - top) * fv


def texture_cube_mip(mips: Sequence[Tensor], d: Tensor, level: Tensor) -> Tensor:
    """Trilinear over an explicit mip stack (mips[l] is [6,R_l,R_l,C]); `level` [N] is the mip_level_bias
    (no uv derivatives), clamped to [0, L-1]."""
    L = len(mips)
    lv = level.clamp(0.0, L - 1.0)
    l0 = lv.detach().floor().long().clamp(0, L - 1)
    l1 = (l0 + 1).clamp(max=L - 1)
    f = (lv - l0)[:, None]
    out0 = torch.zeros(d.shape[0], mips[0].shape[-1], dtype=mips[0].dtype)
    out1 = torch.zeros_like(out0)
    for l in range(L):
        s0 = l0 == l
        s1 = (l1 == l) & (l1 != l0)
        sel = s0 | s1
        if not sel.any():
            continue
        val = texture_cube_linear(mips[l], d[sel])
        full = torch.zeros_like(out0)
        full[sel] = val
        out0 = out0 + full * s0[:, None]
        out1 = out1 + full * s1[:, None]
    out1 = torch.where((l1 == l0)[:, None], out0, out1)
    return out0 + (out1 - out0) * f


def split_mipmaps(packed: Tensor, num_mipmaps: int) -> List[Tensor]:
    """Quad-tree unpack of TextureSplitSum.mipmaps [6,4,R,R] -> list of [6,R_l,R_l,3]
    (restates rfstudio/graphics/_mesh/_texture.py:247-261)."""
    out = []
    cur = packed
    for _ in range(num_mipmaps):
        R = cur.shape[-1]
        out.append(cur[:, :3].permute(0, 2, 3, 1).contiguous())
        HR = R // 2
        cur = cur[:, 3].reshape(6, 2, HR, 2, HR).transpose(-3, -2).reshape(6, 4, HR, HR)
    return out


def merge_mipmaps(mips: Sequence[Tensor]) -> Tensor:
    """Inverse of split_mipmaps (restates _texture.py:228-244; the unused quadrant stays zero)."""
    R = mips[0].shape[1]
    res = torch.zeros(6, 4, R, R, dtype=mips[0].dtype)
    res[:, :3] = mips[0].permute(0, 3, 1, 2)
    o = 0
    for i in range(1, len(mips)):
        h = R // 2
        res[:, 3, o:o + h, o:o + h] = mips[i][..., 0]
        res[:, 3, o:o + h, o + h:o + R] = mips[i][..., 1]
        res[:, 3, o + h:o + R, o:o + h] = mips[i][..., 2]
        o += h
        R = h
    return res
