"""GPU parity of MGAdaptor sampling, vertex normals and the tone map: against the fixture produced by the
reference's own MGAdapter / TriangleMesh code and against the torch oracle on a larger mesh."""
import numpy as np
import pytest
import torch

from geosplatting_b200 import scenes
from geosplatting_b200.mgadapter import MGAdapter, compute_vertex_normals, tone_mapping_naive
from oracle import mgadapter as MG
from oracle import shade as S
from tests.test_golden_cpu import load

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, tol, name=""):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    scale = max(1.0, float(np.abs(b).max()))
    err = float(np.abs(a - b).max())
    assert err <= tol * scale, (name, err, scale)


def test_against_reference_fixture():
    g = load("ref_mgadapter.npz")
    verts = torch.tensor(g["vertices"], device=DEV, requires_grad=True)
    faces = torch.tensor(g["indices"], device=DEV)
    vn = compute_vertex_normals(verts, faces)
    _close(vn, g["vertex_normals"], 1e-6, "vertex normals")
    sp, offsets = MGAdapter().make(verts, faces, vn)
    outs = dict(means=sp.means, scales=sp.scales, quats=sp.quats, colors=sp.colors, opacities=sp.opacities)
    for k, t in outs.items():
        _close(t, g[k], 2e-6, k)
    _close(offsets, g["offsets"], 2e-6, "offsets")
    loss = sum((outs[k] * torch.tensor(g["cot_" + k], device=DEV)).sum() for k in outs)
    gv, = torch.autograd.grad(loss, verts)
    assert float(np.abs(gv.cpu().numpy() - g["v_vertices"]).max()) <= 2e-4 * float(np.abs(g["v_vertices"]).max())


def test_against_oracle_on_a_noisy_icosphere():
    verts, faces = scenes.icosphere(4, radius=0.6)  # 5120 faces -> 30720 Gaussians
    gen = torch.Generator().manual_seed(5)
    verts = verts * (1.0 + 0.05 * torch.randn(verts.shape[0], 1, generator=gen))
    o_v = verts.clone().requires_grad_(True)
    o_vn = MG.vertex_normals(o_v, faces)
    o_out = MG.make(o_v, faces, o_vn)
    cots = [torch.randn(t.shape, generator=gen) for t in o_out[:4]]
    o_g, = torch.autograd.grad(sum((t * c).sum() for t, c in zip(o_out[:4], cots)), o_v)
    d_v = verts.to(DEV).requires_grad_(True)
    d_f = faces.to(DEV)
    vn = compute_vertex_normals(d_v, d_f)
    sp, offsets = MGAdapter().make(d_v, d_f, vn)
    outs = (sp.means, sp.scales, sp.quats, sp.colors)
    for name, a, b in zip(("means", "scales", "quats", "colors"), outs, o_out[:4]):
        _close(a, b, 1e-5, name)
    assert sp.means.shape[0] == 6 * faces.shape[0] and float(sp.opacities.sigmoid().mean()) == pytest.approx(0.99, abs=1e-6)
    _close(offsets, o_out[5], 1e-6, "offsets")
    g, = torch.autograd.grad(sum((t * c.to(DEV)).sum() for t, c in zip(outs, cots)), d_v)
    assert float((g.cpu() - o_g).abs().max()) <= 1e-3 * float(o_g.abs().max())


def test_flat_normals_and_degenerate_faces():
    verts = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [2, 2, 2], [2, 2, 2], [2, 2, 2]], dtype=torch.float32)
    faces = torch.tensor([[0, 1, 2], [3, 4, 5]])  # second face is a point: area clamps, normal falls back to +z
    sp, offsets = MGAdapter().make(verts.to(DEV), faces.to(DEV), None, normal_interpolation=False)
    assert torch.isfinite(sp.means).all() and torch.isfinite(sp.scales).all() and torch.isfinite(sp.quats).all()
    assert torch.allclose(sp.colors[0::2].cpu(), torch.tensor([0.0, 0.0, 1.0]).expand(6, 3))
    vn = compute_vertex_normals(verts.to(DEV), faces.to(DEV))
    assert torch.allclose(vn[3:].cpu(), torch.tensor([0.0, 0.0, 1.0]).expand(3, 3))


def test_tonemap_against_reference_fixture():
    g = load("ref_tonemap.npz")
    rgba = torch.tensor(g["rgba"], device=DEV, requires_grad=True)
    ex = torch.tensor(g["exposure"], device=DEV, requires_grad=True)
    out = tone_mapping_naive(rgba, ex)
    _close(out, g["out"], 1e-6, "tonemap")
    v_rgba, v_ex = torch.autograd.grad((out * torch.tensor(g["cot"], device=DEV)).sum(), [rgba, ex])
    _close(v_rgba, g["v_rgba"], 1e-5, "v_rgba")
    assert abs(float(v_ex) - float(g["v_exposure"].reshape(-1)[0])) <= 1e-4 * abs(float(g["v_exposure"].reshape(-1)[0]))
