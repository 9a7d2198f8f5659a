"""GaussianField: the glue between the mesh, the hash-grid fields and the renderer.

Mirrors rfstudio/model/geosplat.py:482-674 for what GeoSplatter's stage 1 uses:
    get_patches                  :520-556   per-vertex normals and areas of the vertex-sampling warm-up
    get_gaussians_from_vertex    :558-620   one Gaussian per vertex (the first `vertex_sample_warmup` steps)
    get_gaussians_from_face      :622-674   MGAdaptor Gaussians, `shifted_means = means - offsets * sigmoid(z)`
The heavy parts run on this library's kernels (MGAdapter / compute_vertex_normals: gsb_mgadapter_*, gsb_vertex_normals_*;
the three fields: gsb_hashgrid_*); what is left here is O(N) elementwise glue in torch, once per training step.
`occ_enc` (stage 2) and activation checkpointing (`use_checkpoint`, a memory knob for 24 GB cards) are not mirrored.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from .encoding import HashEncoding, kd_field, ks_field, z_field
from .mgadapter import MGAdapter, compute_vertex_normals
from .splat import RenderableAttrs, Splats


def safe_normalize(v: Tensor) -> Tensor:
    """rfstudio/graphics/math.py:119-128."""
    lengths = v.norm(dim=-1, keepdim=True)
    return torch.where(lengths < 1e-6, torch.tensor([0, 0, 1]).to(v), v / lengths.clamp_min(1e-6))


def get_rotation_from_relative_vectors(a: Tensor, b: Tensor, *, eps: float = 1e-6) -> Tensor:
    """rfstudio/graphics/math.py:159-188: rotation taking a to b (Rodrigues); exactly opposite vectors are nudged by
    uniform noise of amplitude 0.005, as the reference does (a nondeterministic branch)."""
    a = a / a.norm(dim=-1, keepdim=True)
    b = b / b.norm(dim=-1, keepdim=True)
    c = (a * b).sum(-1)
    invalid = c < -1 + eps
    if invalid.any():
        offset = torch.where(invalid[..., None], (torch.rand(a.shape, device=a.device) - 0.5) * 0.01, 0)
        return get_rotation_from_relative_vectors(a + offset, b)
    v = torch.cross(a.expand(*c.shape, 3), b.expand(*c.shape, 3), dim=-1)
    s = v.norm(dim=-1)
    z = torch.zeros_like(v[..., 0])
    skew = torch.stack([z, -v[..., 2], v[..., 1], v[..., 2], z, -v[..., 0], -v[..., 1], v[..., 0], z], -1)
    skew = skew.view(*v.shape[:-1], 3, 3)
    factor = (1 - c) / (s ** 2 + eps)
    return torch.eye(3, device=a.device) + skew + skew @ skew * factor[..., None, None]


def rot2quat(rots: Tensor) -> Tensor:
    """rfstudio/graphics/math.py:246-278 (wxyz, best-conditioned candidate)."""
    batch = rots.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(rots.reshape(*batch, 9), dim=-1)
    q = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1)
    q_abs = torch.sqrt(q.clamp_min(0.0))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1),
    ], dim=-2)
    cand = cand / (2.0 * q_abs[..., None].max(q_abs.new_tensor(0.1)))
    pick = q_abs.argmax(dim=-1)
    return torch.gather(cand, -2, pick[..., None, None].expand(*batch, 1, 4)).squeeze(-2)


class GaussianField(nn.Module):
    """rfstudio/model/geosplat.py:482-518: the kd / ks / z fields with the reference's configuration."""

    def __init__(self, kd_enc: Optional[HashEncoding] = None, ks_enc: Optional[HashEncoding] = None,
                 z_enc: Optional[HashEncoding] = None):
        super().__init__()
        self.kd_enc = kd_enc if kd_enc is not None else kd_field()
        self.ks_enc = ks_enc if ks_enc is not None else ks_field()
        self.z_enc = z_enc if z_enc is not None else z_field()

    def get_patches(self, vertices: Tensor, faces: Tensor) -> Tuple[Tensor, Tensor]:
        """:520-556 -> (vertex normals [V,3], vertex areas [V,1]).  Normals here are the sum of UNIT face normals
        (not area-weighted as compute_vertex_normals); area = sum_faces (face cross . vertex normal) / 6."""
        F = faces.shape[0]
        idx = faces.reshape(-1)
        tri = vertices[faces]                                                     # [F,3,3]
        wfn = torch.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0], dim=-1)   # [F,3]
        unit = safe_normalize(wfn)
        normals = torch.zeros_like(vertices).index_add_(0, idx, unit[:, None, :].expand(F, 3, 3).reshape(-1, 3))
        normals = safe_normalize(normals)
        products = (wfn[:, None, :] * normals[faces]).sum(-1)                     # [F,3]
        areas = torch.zeros_like(normals[:, :1]).index_add_(0, idx, products.reshape(-1, 1))
        return normals, areas.clamp_min(1e-10) / 6

    def _attrs(self, inputs: Tensor, normals: Tensor, kd_perturb_std: float, ks_perturb_std: float,
               initial_guess: Tensor) -> RenderableAttrs:
        kd_jitter = ks_jitter = None
        if kd_perturb_std > 0:
            p = torch.normal(mean=0, std=kd_perturb_std, size=inputs.shape, device=inputs.device)
            kd_jitter = self.kd_enc((inputs + p).clamp(-1, 1))
        if ks_perturb_std > 0:
            p = torch.normal(mean=0, std=ks_perturb_std, size=inputs.shape, device=inputs.device)
            ks_jitter = (self.ks_enc((inputs + p).clamp(-1, 1)) + initial_guess).sigmoid()
        return RenderableAttrs(kd=self.kd_enc(inputs), ks=(self.ks_enc(inputs) + initial_guess).sigmoid(), normals=normals,
                               kd_jitter=kd_jitter, ks_jitter=ks_jitter)

    def get_gaussians_from_vertex(self, kd_perturb_std: float, ks_perturb_std: float, scale: float, vertices: Tensor,
                                  faces: Tensor, initial_guess: Tensor) -> Tuple[Splats, RenderableAttrs]:
        """:558-620: one disc per vertex, area/2.5, pushed inward along the normal by sigmoid(z) * sqrt(area/2.5)."""
        normals, areas = self.get_patches(vertices, faces)
        log_sqrt_areas = (areas * (1 / 2.5)).log() * 0.5
        inputs = (vertices / scale).clamp(-1, 1)
        attrs = self._attrs(inputs, normals, kd_perturb_std, ks_perturb_std, initial_guess)
        zs = self.z_enc(inputs.detach()).sigmoid()
        z_axis = torch.tensor([0, 0, 1]).to(normals)
        base_rot = get_rotation_from_relative_vectors(z_axis, normals.detach())
        scales = torch.cat((log_sqrt_areas, log_sqrt_areas, torch.empty_like(log_sqrt_areas).fill_(1e-10).log()), dim=-1)
        positions = vertices - normals * (log_sqrt_areas.detach().exp() * zs)
        V = positions.shape[0]
        return Splats(means=positions, scales=scales, quats=rot2quat(base_rot), colors=torch.empty_like(normals),
                      opacities=torch.logit(0.99 * torch.ones((V, 1), device=vertices.device))), attrs

    def get_gaussians_from_face(self, vertices: Tensor, faces: Tensor, kd_perturb_std: float, ks_perturb_std: float, *,
                                scale: float, initial_guess: Tensor) -> Tuple[Splats, RenderableAttrs, Tensor]:
        """:622-674: MGAdaptor sampling, fields at clamp(means / scale), means shifted inward by offsets * sigmoid(z)."""
        splats, offsets = MGAdapter().make(vertices, faces, compute_vertex_normals(vertices, faces))
        means = (splats.means / scale).clamp(-1, 1)
        offsets = offsets * self.z_enc(means.detach()).sigmoid()
        attrs = self._attrs(means, splats.colors, kd_perturb_std, ks_perturb_std, initial_guess)
        return (Splats(splats.means - offsets, splats.scales, splats.quats, splats.colors, splats.opacities), attrs,
                offsets)
