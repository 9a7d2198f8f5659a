"""Host side of MGAdaptor sampling, vertex normals and the tone map (C ABI: gsb_mgadapter_*,
gsb_vertex_normals_*, gsb_tonemap_*).

Mirrors:
    MGAdapter().make(mesh, normal_interpolation=True)   rfstudio/model/geosplat.py:426-472
    TriangleMesh.compute_vertex_normals(fix=True)        rfstudio/graphics/_mesh/_triangle_mesh.py:588-614
    _tone_mapping_naive(rgba, exposure)                  rfstudio/model/geosplat.py:474-476
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import Tensor

from ._lib import call, f32c, ptr, stream_ptr
from ._lib import require_cuda as _require_cuda


def _i64c(t: Tensor) -> Tensor:
    t = t.detach()
    if t.dtype != torch.int64:
        t = t.long()
    return t.contiguous()


class _VertexNormals(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vertices: Tensor, faces: Tensor):
        v, f = f32c(vertices), _i64c(faces)
        dev = v.device
        V, F = v.shape[0], f.shape[0]
        raw = torch.empty(V, 3, dtype=torch.float32, device=dev)
        normals = torch.empty(V, 3, dtype=torch.float32, device=dev)
        call("gsb_vertex_normals_fwd", dev, C.c_int32(V), C.c_int32(F), ptr(v), ptr(f), ptr(raw), ptr(normals),
             stream_ptr(dev))
        ctx.save_for_backward(v, f, raw)
        return normals

    @staticmethod
    def backward(ctx, v_normals):
        v, f, raw = ctx.saved_tensors
        dev = v.device
        V, F = v.shape[0], f.shape[0]
        scratch = torch.empty(V, 3, dtype=torch.float32, device=dev)
        v_vertices = torch.zeros(V, 3, dtype=torch.float32, device=dev)
        call("gsb_vertex_normals_bwd", dev, C.c_int32(V), C.c_int32(F), ptr(v), ptr(f), ptr(raw), ptr(f32c(v_normals)),
             ptr(scratch), ptr(v_vertices), stream_ptr(dev))
        return v_vertices, None


def compute_vertex_normals(vertices: Tensor, faces: Tensor) -> Tensor:
    """Area-weighted vertex normals with the `fix=True` fallback (0,0,1) for degenerate vertices."""
    _require_cuda(vertices, "compute_vertex_normals")
    return _VertexNormals.apply(vertices, faces)


class _MGAdapter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vertices: Tensor, vertex_normals: Optional[Tensor], faces: Tensor):
        v, f = f32c(vertices), _i64c(faces)
        vn = None if vertex_normals is None else f32c(vertex_normals)
        dev = v.device
        F = f.shape[0]
        N = 6 * F
        means = torch.empty(N, 3, dtype=torch.float32, device=dev)
        scales = torch.empty(N, 3, dtype=torch.float32, device=dev)
        quats = torch.empty(N, 4, dtype=torch.float32, device=dev)
        normals = torch.empty(N, 3, dtype=torch.float32, device=dev)
        opac = torch.empty(N, 1, dtype=torch.float32, device=dev)
        offsets = torch.empty(N, 3, dtype=torch.float32, device=dev)
        call("gsb_mgadapter_fwd", dev, C.c_int32(F), ptr(v), ptr(vn), ptr(f), ptr(means), ptr(scales), ptr(quats),
             ptr(normals), ptr(opac), ptr(offsets), stream_ptr(dev))
        ctx.save_for_backward(v, f, *([vn] if vn is not None else []))
        ctx.has_vn = vn is not None
        ctx.mark_non_differentiable(opac, offsets)
        return means, scales, quats, normals, opac, offsets

    @staticmethod
    def backward(ctx, v_means, v_scales, v_quats, v_normals, _vo, _voff):
        saved = ctx.saved_tensors
        v, f = saved[0], saved[1]
        vn = saved[2] if ctx.has_vn else None
        dev = v.device
        F = f.shape[0]
        N = 6 * F

        def z(g, shape):
            return torch.zeros(shape, dtype=torch.float32, device=dev) if g is None else f32c(g)

        v_vertices = torch.zeros_like(v)
        v_vn = torch.zeros_like(vn) if vn is not None else None
        call("gsb_mgadapter_bwd", dev, C.c_int32(F), ptr(v), ptr(vn), ptr(f), ptr(z(v_means, (N, 3))),
             ptr(z(v_scales, (N, 3))), ptr(z(v_quats, (N, 4))), ptr(z(v_normals, (N, 3))), ptr(v_vertices), ptr(v_vn),
             stream_ptr(dev))
        return v_vertices, v_vn, None


@dataclass
class MGSplats:
    """The `Splats` fields MGAdapter.make fills (rfstudio/graphics/_splats.py:17-32): log-scales, wxyz quats,
    logit opacities [N,1]; `colors` carries the per-Gaussian normal, as in the reference."""
    means: Tensor
    scales: Tensor
    quats: Tensor
    colors: Tensor
    opacities: Tensor


class MGAdapter:
    """Same constants as the reference dataclass (geosplat.py:378-388); they are compiled into the kernel."""

    def make(self, vertices, faces: Optional[Tensor] = None, vertex_normals: Optional[Tensor] = None, *,
             normal_interpolation: bool = True) -> Tuple[MGSplats, Tensor]:
        """geosplat.py:426-431: `make(mesh, normal_interpolation=True)` with a mesh object (fields vertices [V,3],
        indices [F,3], normals [V,3] -- rfstudio's TriangleMesh after compute_vertex_normals(fix=True)), or the three
        tensors spelled out."""
        if faces is None and hasattr(vertices, "vertices") and hasattr(vertices, "indices"):
            mesh = vertices
            vertices, faces, vertex_normals = mesh.vertices, mesh.indices, getattr(mesh, "normals", None)
        _require_cuda(vertices, "MGAdapter")
        if normal_interpolation and vertex_normals is None:
            raise ValueError("normal_interpolation=True needs vertex normals (mesh.compute_vertex_normals(fix=True))")
        vn = vertex_normals if normal_interpolation else None
        means, scales, quats, normals, opac, offsets = _MGAdapter.apply(vertices, vn, faces)
        return MGSplats(means, scales, quats, normals, opac), offsets


class _ToneMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgba: Tensor, exposure: Tensor):
        x, e = f32c(rgba), f32c(exposure).reshape(1)
        out = torch.empty_like(x)
        call("gsb_tonemap_fwd", x.device, C.c_int64(x.numel() // 4), ptr(x), ptr(e), ptr(out), stream_ptr(x.device))
        ctx.save_for_backward(x, e)
        ctx.eshape = exposure.shape
        return out

    @staticmethod
    def backward(ctx, v_out):
        x, e = ctx.saved_tensors
        v_rgba = torch.empty_like(x)
        v_e = torch.zeros(1, dtype=torch.float32, device=x.device)
        call("gsb_tonemap_bwd", x.device, C.c_int64(x.numel() // 4), ptr(x), ptr(e), ptr(f32c(v_out)), ptr(v_rgba),
             ptr(v_e), stream_ptr(x.device))
        return v_rgba, v_e.reshape(ctx.eshape)


def tone_mapping_naive(rgba: Tensor, exposure: Tensor) -> Tensor:
    """rgba[...,4], exposure[1] (device tensor) -> [...,4]."""
    _require_cuda(rgba, "tone_mapping_naive")
    assert rgba.shape[-1] == 4
    return _ToneMap.apply(rgba, exposure)
