"""Host side of FlexiCubes dual marching cubes and the SDF entropy regulariser (C ABI: gsb_fc_*; SURVEY.md section 8f
rank 3): the step that turns GeoSplatter's SDF grid into the mesh MGAdaptor samples.

Mirrors:
    FlexiCubes.from_resolution(*res, scale=...)      rfstudio/graphics/_mesh/_flexicubes.py:397-457
    FlexiCubes.dual_marching_cubes() -> (mesh, L_dev)  rfstudio/graphics/_mesh/_flexicubes.py:559-713 (+ :460-538, :727-802)
    FlexiCubes.compute_entropy()                     rfstudio/graphics/_mesh/_flexicubes.py:715-725
as GeoSplatter.get_geometry drives them (rfstudio/model/geosplat.py:751-769).  grad_func / sdf_eps (the non-differentiable
QEF variant) are not on that path and are not mirrored.

Division of labour.  The library (csrc/flexicubes.cu) does everything: per cube / per dual vertex / per quad / per grid
edge arithmetic, forward and backward, and the ORDER bookkeeping (one stable cub radix sort of the 64-bit edge keys,
three cub scans) that decides the numbering of surface edges, dual vertices, L_dev entries and quads -- which must match
the reference because MGAdaptor emits Gaussians in face order (oracle/flexicubes.py lists the orders).  This module
allocates (scratch sized by upper bounds in N) and makes seven calls per forward; two device->host reads size the
outputs: the number of surface cubes, then {E, n_quads, Q, K}.

Buffers (N surface cubes, E surface edges, Q dual vertices, K (group, edge) entries, n_quads):
    cases[F] i32        occupancy bit mask of every cube         surf_flag[F] i32  1 for surface cubes
    surf_ids[N] i32     surface cubes ascending                  case_ids[N] i32   case after ambiguity resolution
    num_vd[N] i32       dual vertices of the cube (1..4)         n_entries[N] i32  (group, edge) entries of the cube
    edge_of[N,12] i32   surface-edge id or -1                    vd_base[N] / k_base[N] i32  first dual vertex / entry
    surf_edges[E,2] i32 endpoints, ascending (v_a, v_b)          quad_entry[n_quads,4] i32 positions into vd_of
    vd_of[N,12] i32     dual vertex of every (cube, edge)        quad_vd[n_quads,4] i32 in winding order
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from ._lib import call, f32c, ptr, stream_ptr
from ._lib import require_cuda as _require_cuda

_TABLE_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "flexicubes_tables.npz")
_tables_cache: Dict[str, Dict[str, Tensor]] = {}
ENTROPY_REPLICAS = 64          # GSB_FC_ENTROPY_REPLICAS (include/geosplat_b200.h)


def _tables(dev: torch.device) -> Dict[str, Tensor]:
    """The four lookup tables of the published algorithm (cube edges, ambiguity check, dual-vertex groups and counts;
    the values _flexicubes.py:17-366 holds), int32 on `dev`."""
    key = str(dev)
    if key not in _tables_cache:
        with np.load(_TABLE_PATH) as z:
            _tables_cache[key] = {k: torch.from_numpy(z[k].astype(np.int32)).contiguous().to(dev) for k in z.files}
    return _tables_cache[key]


def _i32(shape, dev) -> Tensor:
    return torch.empty(shape, dtype=torch.int32, device=dev)


def _f32(shape, dev, zero: bool = False) -> Tensor:
    return (torch.zeros if zero else torch.empty)(shape, dtype=torch.float32, device=dev)


@dataclass
class Topology:
    """Everything integer that one SDF sign pattern determines (no gradients flow through it)."""
    N: int
    E: int
    Q: int
    K: int
    n_quads: int
    surf_ids: Tensor        # [N] (a view of an [F] buffer)
    case_ids: Tensor        # [N]
    num_vd: Tensor          # [N]
    vd_base: Tensor         # [N]
    k_base: Tensor          # [N]
    edge_of: Tensor         # [N,12]
    surf_edges: Tensor      # [12 N, 2], the first E rows are used
    quad_entry: Tensor      # [3 N, 4], the first n_quads rows are used


def _workspace(F: int, N: int, dev: torch.device) -> Tuple[Tensor, int]:
    need = C.c_size_t(0)
    call("gsb_fc_workspace_bytes", dev, C.c_int32(F), C.c_int32(N), C.byref(need))
    return torch.empty(need.value, dtype=torch.uint8, device=dev), need.value


def _topology(sdf: Tensor, cubes: Tensor, res: Tuple[int, int, int]) -> Topology:
    """_get_case_id + _identify_surf_edges (_flexicubes.py:460-538), the dual-vertex / L_dev numbering of the num_vd loop
    (:640-690) and the quad list of _triangulate (:758-771), sequenced natively: two calls, one host read each."""
    dev = sdf.device
    tb = _tables(dev)
    F, V = cubes.shape[0], sdf.shape[0]
    st = stream_ptr(dev)
    cases, flag, surf_ids = _i32(F, dev), _i32(F, dev), _i32(F, dev)
    ws, ws_bytes = _workspace(F, 0, dev)
    n_surf = C.c_int32(0)
    call("gsb_fc_surface", dev, C.c_int32(F), ptr(sdf), ptr(cubes), ptr(cases), ptr(flag), ptr(surf_ids), ptr(ws),
         C.c_size_t(ws_bytes), C.byref(n_surf), st)                                  # read 1: N
    N = n_surf.value
    if N == 0:
        raise AssertionError("FlexiCubes: the SDF has no sign change (the reference asserts N > 0, _flexicubes.py:605)")
    case_ids, num_vd, vd_base, k_base = (_i32(N, dev) for _ in range(4))
    edge_of, surf_edges, quad_entry = _i32((N, 12), dev), _i32((12 * N, 2), dev), _i32((3 * N, 4), dev)
    ws, ws_bytes = _workspace(F, N, dev)
    counts = (C.c_int32 * 4)()
    call("gsb_fc_topology", dev, C.c_int32(F), C.c_int32(N), C.c_int64(V), C.c_int32(res[0]), C.c_int32(res[1]),
         C.c_int32(res[2]), ptr(sdf), ptr(cubes), ptr(surf_ids), ptr(cases), ptr(flag), ptr(tb["check"]),
         ptr(tb["num_vd"]), ptr(tb["dmc"]), ptr(tb["cube_edges"]), ptr(case_ids), ptr(num_vd), ptr(vd_base), ptr(k_base),
         ptr(edge_of), ptr(surf_edges), ptr(quad_entry), ptr(ws), C.c_size_t(ws_bytes), counts, st)   # read 2
    E, n_quads, Q, K = (int(c) for c in counts)
    return Topology(N=N, E=E, Q=Q, K=K, n_quads=n_quads, surf_ids=surf_ids[:N], case_ids=case_ids, num_vd=num_vd,
                    vd_base=vd_base, k_base=k_base, edge_of=edge_of, surf_edges=surf_edges, quad_entry=quad_entry)


class _DualMarchingCubes(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vertices: Tensor, sdf_values: Tensor, alpha: Tensor, beta: Tensor, gamma: Tensor, cubes: Tensor,
                res: Tuple[int, int, int]):
        v, s = f32c(vertices), f32c(sdf_values).view(-1)
        a, b, g = f32c(alpha), f32c(beta), f32c(gamma).view(-1)
        dev = v.device
        tb = _tables(dev)
        st = stream_ptr(dev)
        t = _topology(s, cubes, res)
        vd, vd_gamma, l_dev = _f32((t.Q, 3), dev), _f32(t.Q, dev), _f32(t.K, dev)
        vd_of = torch.zeros(t.N, 12, dtype=torch.int32, device=dev)
        call("gsb_fc_dual_fwd", dev, C.c_int32(t.N), ptr(t.surf_ids), ptr(t.case_ids), ptr(t.num_vd), ptr(t.vd_base),
             ptr(t.k_base), ptr(tb["dmc"]), ptr(tb["cube_edges"]), ptr(t.edge_of), ptr(t.surf_edges), ptr(v), ptr(s),
             ptr(a), ptr(b), ptr(g), ptr(vd), ptr(vd_gamma), ptr(vd_of), ptr(l_dev), st)
        nq = t.n_quads
        quad = _i32((nq, 4), dev)
        call("gsb_fc_quad_gather", dev, C.c_int32(nq), ptr(t.quad_entry), ptr(vd_of), ptr(quad), st)
        centres = _f32((nq, 3), dev)
        faces = torch.empty(4 * nq, 3, dtype=torch.int64, device=dev)
        call("gsb_fc_quad_fwd", dev, C.c_int32(nq), C.c_int32(t.Q), ptr(quad), ptr(vd), ptr(vd_gamma), ptr(centres),
             ptr(faces), st)
        ctx.topo, ctx.cubes = t, cubes
        ctx.save_for_backward(v, s, a, b, g, vd, vd_gamma, quad)
        ctx.shapes = (sdf_values.shape, gamma.shape)
        ctx.mark_non_differentiable(faces)
        return torch.cat([vd, centres]), faces, l_dev

    @staticmethod
    def backward(ctx, v_mesh, _v_faces, v_l_dev):
        v, s, a, b, g, vd, vd_gamma, quad = ctx.saved_tensors
        t: Topology = ctx.topo
        dev = v.device
        tb = _tables(dev)
        st = stream_ptr(dev)
        v_mesh = _f32((t.Q + t.n_quads, 3), dev, zero=True) if v_mesh is None else f32c(v_mesh)
        v_l_dev = _f32(t.K, dev, zero=True) if v_l_dev is None else f32c(v_l_dev)
        v_vd = v_mesh[:t.Q].clone()
        v_centres = v_mesh[t.Q:].contiguous()
        v_vd_gamma = _f32(t.Q, dev, zero=True)
        call("gsb_fc_quad_bwd", dev, C.c_int32(t.n_quads), C.c_int32(t.Q), ptr(quad), ptr(vd), ptr(vd_gamma),
             ptr(v_centres), ptr(v_vd), ptr(v_vd_gamma), st)
        v_vertices, v_sdf = torch.zeros_like(v), torch.zeros_like(s)
        v_alpha, v_beta, v_gamma = torch.zeros_like(a), torch.zeros_like(b), torch.zeros_like(g)
        call("gsb_fc_dual_bwd", dev, C.c_int32(t.N), ptr(t.surf_ids), ptr(t.case_ids), ptr(t.num_vd), ptr(t.vd_base),
             ptr(t.k_base), ptr(tb["dmc"]), ptr(tb["cube_edges"]), ptr(t.edge_of), ptr(t.surf_edges), ptr(v), ptr(s),
             ptr(a), ptr(b), ptr(g), ptr(v_vd), ptr(v_vd_gamma), ptr(v_l_dev), ptr(v_vertices), ptr(v_sdf), ptr(v_alpha),
             ptr(v_beta), ptr(v_gamma), st)
        sdf_shape, gamma_shape = ctx.shapes
        return v_vertices, v_sdf.view(sdf_shape), v_alpha, v_beta, v_gamma.view(gamma_shape), None, None


class _Entropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sdf_values: Tensor, grid_edges: Tensor):
        s = f32c(sdf_values).view(-1)
        dev = s.device
        sums, partials = _f32(3, dev), _f32(3 * ENTROPY_REPLICAS, dev)
        call("gsb_fc_entropy_fwd", dev, C.c_int64(grid_edges.shape[0]), ptr(grid_edges), ptr(s), ptr(sums),
             ptr(partials), stream_ptr(dev))
        ctx.save_for_backward(s, grid_edges, sums)
        ctx.shape = sdf_values.shape
        return (sums[0] + sums[1]) / sums[2]

    @staticmethod
    def backward(ctx, v_loss):
        s, grid_edges, sums = ctx.saved_tensors
        dev = s.device
        v_sdf = torch.zeros_like(s)
        call("gsb_fc_entropy_bwd", dev, C.c_int64(grid_edges.shape[0]), ptr(grid_edges), ptr(s), ptr(sums),
             ptr(f32c(v_loss).view(1)), ptr(v_sdf), stream_ptr(dev))
        return v_sdf.view(ctx.shape), None


@dataclass
class TriangleMesh:
    """The two fields of rfstudio's TriangleMesh that this path produces (vertices [V,3] fp32, indices [F,3] int64)."""
    vertices: Tensor
    indices: Tensor
    normals: Optional[Tensor] = None

    def compute_vertex_normals(self, *, fix: bool = True) -> "TriangleMesh":
        """_triangle_mesh.py:588-614 (area-weighted, (0,0,1) for degenerate vertices)."""
        if not fix:
            raise NotImplementedError("only fix=True is on the path (geosplat.py:786)")
        from .mgadapter import compute_vertex_normals
        return TriangleMesh(self.vertices, self.indices, compute_vertex_normals(self.vertices, self.indices))


@dataclass
class FlexiCubes:
    """Same fields as the reference dataclass (_flexicubes.py:368-395); alpha / beta / gamma are the RAW parameters."""
    vertices: Tensor                      # [V,3]
    sdf_values: Tensor                    # [V,1]
    indices: Tensor                       # [F,8] int64 grid-vertex ids of every cube
    resolution: Tensor                    # [3] int64
    alpha: Optional[Tensor] = None        # [F,8]
    beta: Optional[Tensor] = None         # [F,12]
    gamma: Optional[Tensor] = None        # [F,1]
    _static: dict = field(default_factory=dict, repr=False, compare=False)   # per-grid constants, shared by replace()

    @classmethod
    def from_resolution(cls, *resolution: int, device=None, random_sdf: bool = True, scale: float = 1.0) -> "FlexiCubes":
        """Regular voxel grid over [-scale, scale]^3 with the reference's numbering (_flexicubes.py:397-457): grid
        vertices enumerated with the LAST axis fastest and stored as (i, j, k) / res, cube c at
        (c % R0, c // R0 % R1, c // (R0 R1)) with corner (dx, dy, dz) -> vertex ((z (R1+1) + y) (R0+1) + x)."""
        assert len(resolution) in (1, 3)
        r0, r1, r2 = (resolution * 3) if len(resolution) == 1 else resolution
        res = torch.tensor([r0, r1, r2], dtype=torch.int64, device=device)
        ax = [torch.arange(r + 1, device=device) for r in (r0, r1, r2)]
        coords = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3).float()
        verts = coords / res
        c = torch.arange(r0 * r1 * r2, device=device)
        origin = torch.stack([c % r0, (c // r0) % r1, c // (r1 * r0)], -1)
        corner = torch.tensor([[k & 1, (k >> 1) & 1, (k >> 2) & 1] for k in range(8)], dtype=torch.int64, device=device)
        p = origin[:, None, :] + corner
        cubes = (p[..., 2] * (1 + r1) + p[..., 1]) * (1 + r0) + p[..., 0]
        sdf = (torch.rand_like(verts[:, :1]) - 0.1) if random_sdf else torch.zeros_like(verts[:, :1])
        return cls(vertices=(2 * verts - 1) * scale, sdf_values=sdf, indices=cubes, resolution=res)

    @property
    def device(self) -> torch.device:
        return self.vertices.device

    def replace(self, **kw) -> "FlexiCubes":
        f = {k: getattr(self, k) for k in ("vertices", "sdf_values", "indices", "resolution", "alpha", "beta", "gamma")}
        f.update(kw)
        static = self._static if kw.get("indices", self.indices) is self.indices else {}
        return FlexiCubes(**f, _static=static)

    def to(self, device) -> "FlexiCubes":
        mv = lambda t: None if t is None else t.to(device)   # noqa: E731
        return FlexiCubes(mv(self.vertices), mv(self.sdf_values), mv(self.indices), mv(self.resolution), mv(self.alpha),
                          mv(self.beta), mv(self.gamma))

    # -- per-grid constants ----------------------------------------------------------------------------------------
    def _cubes_i32(self) -> Tensor:
        if "cubes" not in self._static:
            self._static["cubes"] = self.indices.int().contiguous()
            self._static["res"] = tuple(int(r) for r in self.resolution.tolist())
        return self._static["cubes"]

    def _grid_edges(self) -> Tensor:
        """Every grid edge once, [U,2] int32 (the unique() of compute_entropy, _flexicubes.py:716-717; the edge set does
        not depend on the SDF, so it is computed once per grid)."""
        if "grid_edges" not in self._static:
            ce = _tables(self.indices.device)["cube_edges"].long()
            self._static["grid_edges"] = self.indices[:, ce].view(-1, 2).unique(dim=0).int().contiguous()
        return self._static["grid_edges"]

    # -- the two operators -----------------------------------------------------------------------------------------
    def dual_marching_cubes(self, *, weight_scale: float = 0.99) -> Tuple[TriangleMesh, Tensor]:
        """(mesh, L_dev[K]); differentiable w.r.t. vertices, sdf_values, alpha, beta, gamma."""
        _require_cuda(self.vertices, "FlexiCubes.dual_marching_cubes")
        if weight_scale != 0.99:
            raise NotImplementedError("weight_scale is compiled into the kernels (0.99, the only value on the path)")
        F = self.indices.shape[0]
        dev = self.vertices.device
        # absent weights: raw zeros give alpha = beta = 1 and a uniform gamma, the reference's defaults (:607-621)
        alpha = self.alpha if self.alpha is not None else _f32((F, 8), dev, zero=True)
        beta = self.beta if self.beta is not None else _f32((F, 12), dev, zero=True)
        gamma = self.gamma if self.gamma is not None else _f32((F, 1), dev, zero=True)
        cubes = self._cubes_i32()
        verts, faces, l_dev = _DualMarchingCubes.apply(self.vertices, self.sdf_values, alpha, beta, gamma, cubes,
                                                       self._static["res"])
        return TriangleMesh(vertices=verts, indices=faces), l_dev

    def compute_entropy(self) -> Tensor:
        """Symmetric BCE between the endpoint SDF values of every sign-changing grid edge (scalar)."""
        _require_cuda(self.sdf_values, "FlexiCubes.compute_entropy")
        return _Entropy.apply(self.sdf_values, self._grid_edges())
