"""Differentiable torch restatement of MGAdapter.make / bary2gs / rot2quat / compute_vertex_normals_(fix=True).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows rfstudio/model/geosplat.py:378-472,
rfstudio/graphics/math.py:246-278 and rfstudio/graphics/_mesh/_triangle_mesh.py:588-614; pinned on the
reference's own code through tests/golden/ref_mgadapter.npz (tests/test_golden_cpu.py).
"""
from __future__ import annotations

import torch
from torch import Tensor

from .shade import safe_normalize


def rot2quat(R: Tensor) -> Tensor:
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = R.reshape(-1, 9).unbind(-1)
    q = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], -1)
    q_abs = torch.where(q > 0, torch.sqrt(q.clamp_min(1e-30)), torch.zeros_like(q))
    cand = torch.stack([
        torch.stack([q_abs[:, 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, q_abs[:, 1] ** 2, m10 + m01, m02 + m20], -1),
        torch.stack([m02 - m20, m10 + m01, q_abs[:, 2] ** 2, m12 + m21], -1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[:, 3] ** 2], -1),
    ], -2) / (2.0 * torch.maximum(q_abs[:, :, None], q_abs.new_tensor(0.1)))
    best = q_abs.argmax(-1)
    return cand[torch.arange(R.shape[0]), best]


def vertex_normals(vertices: Tensor, faces: Tensor) -> Tensor:
    p = vertices[faces]  # [F,3,3]
    w = torch.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0], dim=-1)
    raw = torch.zeros_like(vertices).index_add(0, faces.reshape(-1), w[:, None, :].expand(-1, 3, -1).reshape(-1, 3))
    ln = raw.norm(dim=-1, keepdim=True)
    fix = torch.tensor([0.0, 0.0, 1.0], dtype=vertices.dtype)
    return torch.where(ln > 1e-10, raw / ln.clamp_min(1e-10), fix)


def _bary2gs(p0, p1, area, n, s_ratio):
    means = (p0 + p1) / 2
    mr = p1 - means
    max_s = mr.norm(dim=-1, keepdim=True).clamp(min=1e-10)
    min_s = area / 4 / max_s
    mr = mr / max_s
    scales = torch.cat(((1.6 * s_ratio * max_s).log(), (1.6 / s_ratio * min_s).log(), torch.full_like(max_s, -10.0)), -1)
    quats = rot2quat(torch.stack((mr, torch.cross(n, mr, dim=-1), n), dim=-1))
    return means, scales, quats


def make(vertices: Tensor, faces: Tensor, vnormals: Tensor):
    p0, p1, p2 = vertices[faces[:, 0]], vertices[faces[:, 1]], vertices[faces[:, 2]]
    vn0, vn1, vn2 = vnormals[faces[:, 0]], vnormals[faces[:, 1]], vnormals[faces[:, 2]]
    nrm = torch.cross(p1 - p0, p2 - p0, dim=-1)
    area = nrm.norm(dim=-1, keepdim=True).clamp(min=1e-10) / 2
    n = safe_normalize(nrm)
    offsets = n.detach() * area.detach().sqrt()
    M, S, Q, Cn = [], [], [], []
    for u, a_c, s in zip([1 / 9 - 1 / 24, 2 / 9], [1 / 4 * (1 / 3), 1 / 12 * 3], [0.5, 1.3]):
        us = [p0 * (1 - 2 * u) + (p1 + p2) * u, p1 * (1 - 2 * u) + (p2 + p0) * u, p2 * (1 - 2 * u) + (p0 + p1) * u]
        ns = [vn0 * (1 - 2 * u) + (vn1 + vn2) * u, vn1 * (1 - 2 * u) + (vn2 + vn0) * u, vn2 * (1 - 2 * u) + (vn0 + vn1) * u]
        a = area * a_c
        for e in range(3):
            m, sc, q = _bary2gs(us[e], us[(e + 1) % 3], a, n, s)
            M.append(m); S.append(sc); Q.append(q)
            Cn.append(safe_normalize((ns[e] + ns[(e + 1) % 3]) / 2))
    means, scales, quats, colors = torch.cat(M), torch.cat(S), torch.cat(Q), torch.cat(Cn)
    opac = torch.full_like(means[:, :1], 0.99).logit()
    return means, scales, quats, colors, opac, torch.cat([offsets] * 6)
