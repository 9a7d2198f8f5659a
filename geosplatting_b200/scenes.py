"""Synthetic scenes and cameras for tests and benches (no datasets exist in this environment).

Cameras mirror the reference dataparsers (rfstudio/data/dataparser/syn4relight_dataparser.py:43-77,
tensoir_dataparser.py:43-72, shiny_blender_dataparser.py:35-69): 800x800, fov 0.6911112 rad, eye on a
sphere of radius 4.0311 * 2/3 looking at the origin, y-up, OpenGL camera-to-world like
rfstudio/graphics/_cameras.py:33-52.  SURVEY.md section 8(d) fixes the recipes.
Everything is generated on the CPU with seeded generators and is deterministic across machines.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Tuple

import numpy as np
import torch

FOV = 0.6911112
CAM_RADIUS = 4.0311 * 2.0 / 3.0


@dataclass
class PinholeCamera:
    """One camera in the reference's `Cameras` convention (OpenGL c2w: x right, y up, z backward)."""

    c2w: np.ndarray  # [3,4] float32
    fx: float
    fy: float
    cx: float
    cy: float
    width: int
    height: int

    @property
    def position(self) -> np.ndarray:
        return self.c2w[:, 3]

    @property
    def view_matrix(self) -> np.ndarray:
        """World->camera in OpenCV axes; same algebra as rfstudio/graphics/_cameras.py:299-314."""
        R = self.c2w[:3, :3] * np.array([1.0, -1.0, -1.0], dtype=np.float32)
        T = self.c2w[:3, 3:4]
        R_inv = R.T
        T_inv = R_inv @ -T
        vm = np.zeros((4, 4), dtype=np.float32)
        vm[3, 3] = 1.0
        vm[:3, :3] = R_inv
        vm[:3, 3:4] = T_inv
        return vm

    @property
    def intrinsic_matrix(self) -> np.ndarray:
        """rfstudio/graphics/_cameras.py:289-297."""
        K = np.zeros((3, 3), dtype=np.float32)
        K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[2, 2] = self.fx, self.fy, self.cx, self.cy, 1.0
        return K


def _f(x) -> float:
    return float(x.item()) if hasattr(x, "item") else float(x)


def to_pinhole(cameras) -> list:
    """Any of the camera arguments the reference's operators take -> list of PinholeCamera.

    Accepts a PinholeCamera, a sequence of them, or an object with the fields of rfstudio's `Cameras` tensor dataclass
    (rfstudio/graphics/_cameras.py:33-52: c2w [..,3,4], fx, fy, cx, cy, width, height; batch shape () or (B,)), e.g. the
    `inputs[i:i+1]` slices GeoSplatter.render_report passes to RenderableAttrs.splat (geosplat.py:869-879).  Reading the
    scalars is a device->host copy when the tensors live on the GPU -- the reference does the same per view
    (`camera.width.item()`, rfstudio/model/gsplat.py:295)."""
    if isinstance(cameras, PinholeCamera):
        return [cameras]
    if hasattr(cameras, "c2w") and hasattr(cameras, "fx"):
        c2w = cameras.c2w.detach().to("cpu", torch.float32).reshape(-1, 3, 4).numpy()
        flat = {k: getattr(cameras, k).detach().reshape(-1).cpu() for k in ("fx", "fy", "cx", "cy", "width", "height")}
        return [PinholeCamera(np.ascontiguousarray(c2w[i]), _f(flat["fx"][i]), _f(flat["fy"][i]), _f(flat["cx"][i]),
                              _f(flat["cy"][i]), int(flat["width"][i]), int(flat["height"][i]))
                for i in range(c2w.shape[0])]
    out = []
    for c in cameras:
        out.extend(to_pinhole(c))
    return out


def look_at_camera(eye, width: int, height: int, fov: float = FOV, target=(0.0, 0.0, 0.0)) -> PinholeCamera:
    eye = np.asarray(eye, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    back = eye - target
    back /= np.linalg.norm(back)
    up = np.array([0.0, 1.0, 0.0])
    right = np.cross(up, back)
    right /= np.linalg.norm(right)
    true_up = np.cross(back, right)
    c2w = np.stack([right, true_up, back, eye], axis=1).astype(np.float32)  # [3,4]
    f = 0.5 * width / math.tan(0.5 * fov)
    return PinholeCamera(c2w, f, f, width / 2.0, height / 2.0, width, height)


def orbit_cameras(n: int, width: int, height: int, seed: int = 1, radius: float = CAM_RADIUS):
    """n cameras: azimuth uniform, elevation uniform in [0, 60] degrees."""
    g = torch.Generator().manual_seed(seed)
    az = (torch.rand(n, generator=g) * 2 * math.pi).tolist()
    el = (torch.rand(n, generator=g) * math.radians(60.0)).tolist()
    cams = []
    for a, e in zip(az, el):
        eye = (radius * math.cos(e) * math.sin(a), radius * math.sin(e), radius * math.cos(e) * math.cos(a))
        cams.append(look_at_camera(eye, width, height))
    return cams


def random_gaussians(n: int, seed: int = 0, extent: float = 0.7, scale_lo: float = 0.005, scale_hi: float = 0.05):
    """BASELINE config 1: means ~U([-extent,extent]^3), log-scales ~U(log lo, log hi), quats ~N(0,1)^4,
    opacity logit ~U(-2,4), colours ~U(0,1)^3.  Returns post-activation tensors as gsplat takes them
    (scales = exp(log-scales), opacities = sigmoid(logit)) on the CPU."""
    g = torch.Generator().manual_seed(seed)
    means = (torch.rand(n, 3, generator=g) * 2 - 1) * extent
    log_s = torch.rand(n, 3, generator=g) * (math.log(scale_hi) - math.log(scale_lo)) + math.log(scale_lo)
    quats = torch.randn(n, 4, generator=g)
    logit = torch.rand(n, generator=g) * 6 - 2
    colors = torch.rand(n, 3, generator=g)
    return dict(means=means, quats=quats, scales=log_s.exp(), opacities=torch.sigmoid(logit), colors=colors)


def surface_gaussians(n: int, seed: int = 0, radius: float = 0.6) -> dict:
    """A GeoSplatting-like workload without running FlexiCubes: n flat disc Gaussians tiling a noisy
    sphere of the given radius (what MGAdaptor emits for a closed surface: third scale exp(-10),
    in-plane scales set so neighbouring discs overlap, constant opacity 0.99, normal = disc axis).
    Used for the large bench configs; the mesh->Gaussian path has its own generator (`icosphere`)."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    bump = 1.0 + 0.08 * torch.sin(5.0 * d[:, 0:1]) * torch.cos(4.0 * d[:, 1:2]) + 0.05 * torch.sin(9.0 * d[:, 2:3])
    means = d * radius * bump
    normals = d
    # tangent frame
    helper = torch.where(d[:, 1:2].abs() < 0.9, torch.tensor([0.0, 1.0, 0.0]), torch.tensor([1.0, 0.0, 0.0]))
    t1 = torch.cross(helper.expand_as(d), d, dim=-1)
    t1 = t1 / t1.norm(dim=-1, keepdim=True)
    ang = torch.rand(n, 1, generator=g) * 2 * math.pi
    t2 = torch.cross(d, t1, dim=-1)
    u = t1 * torch.cos(ang) + t2 * torch.sin(ang)
    v = torch.cross(d, u, dim=-1)
    R = torch.stack((u, v, d), dim=-1)  # columns: major, minor, normal
    quats = rotmat_to_quat(R)
    area = 4 * math.pi * radius * radius / n
    base = math.sqrt(area)
    aniso = torch.exp((torch.rand(n, 1, generator=g) - 0.5) * 1.2)
    scales = torch.cat((1.1 * base * aniso, 1.1 * base / aniso, torch.full((n, 1), math.exp(-10.0))), dim=-1)
    opacities = torch.full((n,), 0.99)
    kd = torch.rand(n, 3, generator=g) * 0.8 + 0.1
    ks = torch.rand(n, 2, generator=g)
    return dict(means=means, quats=quats, scales=scales, opacities=opacities, normals=normals, kd=kd, ks=ks,
                colors=kd.clone())


def rotmat_to_quat(R: torch.Tensor) -> torch.Tensor:
    """wxyz quaternion of a batch of rotation matrices (numerically safe branch selection)."""
    m00, m01, m02 = R[:, 0, 0], R[:, 0, 1], R[:, 0, 2]
    m10, m11, m12 = R[:, 1, 0], R[:, 1, 1], R[:, 1, 2]
    m20, m21, m22 = R[:, 2, 0], R[:, 2, 1], R[:, 2, 2]
    q = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], -1)
    q = q.clamp_min(0).sqrt()
    cand = torch.stack([
        torch.stack([q[:, 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, q[:, 1] ** 2, m10 + m01, m02 + m20], -1),
        torch.stack([m02 - m20, m10 + m01, q[:, 2] ** 2, m12 + m21], -1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q[:, 3] ** 2], -1),
    ], -2) / (2.0 * q[:, :, None].clamp_min(0.1))
    best = q.argmax(-1)
    return cand[torch.arange(R.shape[0]), best]


def icosphere(subdivisions: int, radius: float = 0.6) -> Tuple[torch.Tensor, torch.Tensor]:
    """Closed triangle mesh (vertices [V,3] f32, indices [F,3] i64) with F = 20 * 4**subdivisions."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(subdivisions):
        edges = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
        edges_sorted = np.sort(edges, 1)
        uniq, inv = np.unique(edges_sorted, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = v.shape[0]
        v = np.concatenate([v, mid], 0)
        F = f.shape[0]
        inv = inv.reshape(-1)
        m01, m12, m20 = base + inv[:F], base + inv[F:2 * F], base + inv[2 * F:]
        f = np.concatenate([
            np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], 0)
    return torch.from_numpy((v * radius).astype(np.float32)), torch.from_numpy(f)


def cube_sphere(n: int, radius: float = 0.6, bump: float = 0.06) -> Tuple[torch.Tensor, torch.Tensor]:
    """Closed triangle mesh of a bumpy sphere from a subdivided cube: 12*n*n faces, so the MGAdaptor emits
    72*n*n Gaussians (n=83 -> 496k, n=118 -> 1.00M, n=167 -> 2.01M, n=264 -> 5.02M: the BASELINE configs).
    Vertices are shared across cube edges (welded), winding is outward."""
    lin = np.linspace(-1.0, 1.0, n + 1)
    a, b = np.meshgrid(lin, lin, indexing="ij")
    one = np.ones_like(a)
    # (point on face, such that cross(d/da, d/db) points outward)
    faces_pts = [np.stack([one, a, b], -1), np.stack([-one, b, a], -1), np.stack([b, one, a], -1),
                 np.stack([a, -one, b], -1), np.stack([a, b, one], -1), np.stack([b, a, -one], -1)]
    pts = np.concatenate([p.reshape(-1, 3) for p in faces_pts], 0)
    key = np.round((pts + 1.0) * n / 2.0).astype(np.int64)          # integer lattice coords in [0, n]
    flat = (key[:, 0] * (n + 1) + key[:, 1]) * (n + 1) + key[:, 2]
    uniq, first, inv = np.unique(flat, return_index=True, return_inverse=True)
    verts = pts[first]
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    quads = np.stack([idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]], -1).reshape(-1, 4)
    tris = np.concatenate([quads[:, [0, 1, 2]], quads[:, [0, 2, 3]]], 0)
    all_tris = np.concatenate([tris + f * (n + 1) * (n + 1) for f in range(6)], 0)
    all_tris = inv.reshape(-1)[all_tris]
    d = verts / np.linalg.norm(verts, axis=1, keepdims=True)
    r = radius * (1.0 + bump * np.sin(5.0 * d[:, 0:1]) * np.cos(4.0 * d[:, 1:2]) + 0.6 * bump * np.sin(9.0 * d[:, 2:3]))
    return torch.from_numpy((d * r).astype(np.float32)), torch.from_numpy(all_tris.astype(np.int64))
