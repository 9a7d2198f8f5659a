"""FlexiCubes on the GPU: geosplatting_b200/flexicubes.py + the gsb_fc_* kernels through
the C ABI against the fixtures the reference's own code produced, against oracle/flexicubes.py on fresh seeded inputs,
and -- at the BASELINE grid size -- through size-independent properties (watertight, consistently oriented mesh)."""
import numpy as np
import pytest
import torch

from geosplatting_b200 import _lib
from geosplatting_b200.mgadapter import MGAdapter
from tests import fc_cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_smooth_fixture():
    fc_cases.check_smooth_fixture(DEV)


def test_rough_fixture_with_inverted_cases():
    fc_cases.check_rough_fixture(DEV)


@pytest.mark.parametrize("res,noise", [((12, 9, 7), 0.05), ((24, 24, 24), 0.0), ((5, 11, 8), 0.2)])
def test_against_oracle(res, noise):
    _lib.CallStats.reset()
    assert fc_cases.check_against_oracle(DEV, res=res, seed=sum(res), noise=noise) > 0
    # the CUDA library did the work: native topology driver, dual fwd/bwd, quad gather/fwd/bwd, entropy fwd/bwd
    assert {k for k in _lib.CallStats.counts if k.startswith("gsb_fc_")} == {
        "gsb_fc_workspace_bytes", "gsb_fc_surface", "gsb_fc_topology", "gsb_fc_dual_fwd", "gsb_fc_dual_bwd",
        "gsb_fc_quad_gather", "gsb_fc_quad_fwd", "gsb_fc_quad_bwd", "gsb_fc_entropy_fwd", "gsb_fc_entropy_bwd"}


def test_full_size_grid_is_watertight_and_feeds_mgadapter():
    """96^3 grid (the reference trains at 96-128): every directed edge once and its opposite present, vertices near the
    SDF's zero set, finite gradients, and MGAdaptor samples 6 Gaussians per face from it -- the get_gsplat chain."""
    R = 96
    fc0, sdf, deform, w = fc_cases.sphere_case((R, R, R), 7, DEV)
    sdf = sdf.to(DEV).requires_grad_(True)
    w = w.to(DEV).requires_grad_(True)
    deform = deform.to(DEV).requires_grad_(True)
    verts = fc0.vertices + deform.tanh() * (0.5 * 0.9 / R)
    fc = fc0.replace(vertices=verts, sdf_values=sdf, alpha=w[:, :8], beta=w[:, 8:20], gamma=w[:, 20:])
    mesh, l_dev = fc.dual_marching_cubes()
    ent = fc.compute_entropy()
    f = mesh.indices
    assert int(f.min()) == 0 and int(f.max()) == mesh.vertices.shape[0] - 1
    e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    V = mesh.vertices.shape[0]
    fwd = torch.unique(e[:, 0] * V + e[:, 1])
    assert fwd.shape[0] == e.shape[0]                                     # no directed edge twice
    assert torch.equal(fwd, torch.unique(e[:, 1] * V + e[:, 0]))          # every edge has its opposite
    r = mesh.vertices.detach().norm(dim=-1)
    assert 0.40 < float(r.min()) and float(r.max()) < 0.70
    m = mesh.compute_vertex_normals(fix=True)
    splats, offsets = MGAdapter().make(m.vertices, m.indices, m.normals)
    assert splats.means.shape[0] == 6 * f.shape[0]
    loss = splats.means.square().sum() + splats.scales.sum() + l_dev.mean() * 0.5 + ent * 0.3
    grads = torch.autograd.grad(loss, [sdf, deform, w])
    assert all(bool(torch.isfinite(g).all()) for g in grads) and float(grads[0].abs().sum()) > 0
    # determinism of the topology: a second call gives the same faces
    mesh2, _ = fc.dual_marching_cubes()
    assert torch.equal(mesh2.indices, f)
    assert np.isfinite(float(ent.detach()))
