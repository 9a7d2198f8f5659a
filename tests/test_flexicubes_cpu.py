"""FlexiCubes without a GPU: the host module (geosplatting_b200/flexicubes.py) driving the REAL kernel source of
csrc/flexicubes.cu compiled for the host by tests/emu (threads run one after another; see tests/emu/cuda_runtime.h).
This checks the ordering bookkeeping, the kernels' index arithmetic and their gradient formulas against the reference's
own outputs in this GPU-less container; the CUDA build of the same source is checked by tests/test_flexicubes_gpu.py.
The emulation is wired in by monkeypatching here -- the product refuses CPU tensors (last test)."""
import pytest
import torch

from geosplatting_b200 import flexicubes as FC
from tests import fc_cases
from tests.emu import build as emu
from tests.emu.patch import route


@pytest.fixture()
def host_kernels(monkeypatch):
    route(monkeypatch, emu.build("flexicubes"), FC)


def test_kernel_source_on_host_smooth_fixture(host_kernels):
    fc_cases.check_smooth_fixture("cpu")


def test_kernel_source_on_host_rough_fixture(host_kernels):
    fc_cases.check_rough_fixture("cpu")


@pytest.mark.parametrize("res,noise", [((12, 9, 7), 0.05), ((16, 16, 16), 0.0), ((5, 11, 8), 0.2)])
def test_kernel_source_on_host_against_oracle(host_kernels, res, noise):
    assert fc_cases.check_against_oracle("cpu", res=res, seed=sum(res), noise=noise) > 0


def test_absent_weights_and_no_surface(host_kernels):
    """alpha / beta / gamma = None behave like the reference's defaults (uniform weights); an SDF without a sign change
    raises like the reference's assert."""
    fc0, sdf, _, _ = fc_cases.sphere_case((8, 8, 8), 1, "cpu")
    mesh, l_dev = fc0.replace(sdf_values=sdf).dual_marching_cubes()
    z = torch.zeros(fc0.indices.shape[0], 21)
    mesh_z, l_dev_z = fc0.replace(sdf_values=sdf, alpha=z[:, :8], beta=z[:, 8:20], gamma=z[:, 20:]).dual_marching_cubes()
    assert torch.equal(mesh.indices, mesh_z.indices) and torch.equal(mesh.vertices, mesh_z.vertices)
    assert mesh.indices.max() == mesh.vertices.shape[0] - 1 and bool(torch.isfinite(l_dev).all())
    with pytest.raises(AssertionError, match="no sign change"):
        fc0.replace(sdf_values=torch.ones_like(sdf)).dual_marching_cubes()


def test_product_has_no_cpu_path():
    fc0, sdf, _, _ = fc_cases.sphere_case((4, 4, 4), 1, "cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        fc0.replace(sdf_values=sdf).dual_marching_cubes()
    with pytest.raises(RuntimeError, match="no CPU path"):
        fc0.replace(sdf_values=sdf).compute_entropy()


def test_product_tables_are_the_reference_tables():
    """geosplatting_b200/data/flexicubes_tables.npz (written by scripts/make_golden.py section J) holds the four lookup
    tables by value; the fixtures carry the reference's own copies."""
    import numpy as np
    g = fc_cases.load("ref_flexicubes.npz")
    tb = FC._tables(torch.device("cpu"))
    assert set(tb) == {"cube_edges", "check", "dmc", "num_vd"}
    for k, v in tb.items():
        assert v.dtype == torch.int32 and np.array_equal(v.numpy(), g["tbl_" + k])


@pytest.mark.parametrize("res", [(1, 1, 1), (2, 1, 3), (3, 3, 3), (4, 2, 5)])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_tiny_random_grids_against_oracle(host_kernels, res, seed):
    """Random SDF on grids down to a single cube (no quads at all: a mesh with dual vertices and zero faces), every
    case of the ambiguity table reachable at the grid boundary: same faces, vertices, L_dev and SDF gradient."""
    import numpy as np
    from oracle import flexicubes as OF
    gen = torch.Generator().manual_seed(100 * seed + sum(res))
    fc0 = FC.FlexiCubes.from_resolution(*res, random_sdf=False, scale=1.0)
    sdf0 = torch.rand(fc0.vertices.shape[0], 1, generator=gen) - 0.5
    w0 = 0.5 * torch.randn(fc0.indices.shape[0], 21, generator=gen)
    tbl = fc_cases.tables(fc_cases.load("ref_flexicubes.npz"))
    got = {}
    for which in ("ours", "oracle"):
        sdf, w = sdf0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
        if which == "ours":
            mesh, l_dev = fc0.replace(sdf_values=sdf, alpha=w[:, :8], beta=w[:, 8:20], gamma=w[:, 20:]).dual_marching_cubes()
            mv, mf = mesh.vertices, mesh.indices
        else:
            mv, mf, l_dev = OF.dual_marching_cubes(fc0.vertices, sdf, fc0.indices, res, w[:, :8], w[:, 8:20], w[:, 20:], tbl)
        g = torch.autograd.grad(mv.square().sum() + l_dev.sum(), [sdf, w])
        got[which] = [x.detach().numpy() for x in (mf, mv, l_dev, *g)]
    a, b = got["ours"], got["oracle"]
    assert a[0].shape == b[0].shape and np.array_equal(a[0], b[0])
    for x, y in zip(a[1:], b[1:]):
        assert x.shape == y.shape and np.abs(x - y).max(initial=0.0) <= 1e-4 * max(1.0, np.abs(y).max(initial=0.0))


def test_warp_reduction_of_the_entropy_pass_under_simt(monkeypatch):
    """The same fixture with the block's threads as fibers (tests/emu SIMT mode): the warp-shuffle reduction of the
    entropy forward, compiled out in the sequential mode above, runs as written."""
    route(monkeypatch, emu.build("flexicubes", simt=True), FC)
    fc_cases.check_smooth_fixture("cpu")
    fc_cases.check_rough_fixture("cpu")


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_sdf_with_exact_zeros_and_ties(host_kernels, seed):
    """SDF quantised to {-1, -0.5, 0, 0.5, 1}: many vertices exactly on the surface (occupancy is `sdf < 0`, so zero counts
    as outside), crossings that coincide with grid vertices, dual vertices at zero distance from a crossing (the norm's
    subgradient), equal values on both sides of ambiguous faces -- same mesh and finite, matching gradients."""
    import numpy as np
    from oracle import flexicubes as OF
    gen = torch.Generator().manual_seed(seed)
    res = (5, 4, 6)
    fc0 = FC.FlexiCubes.from_resolution(*res, random_sdf=False, scale=1.0)
    sdf0 = (torch.randint(-2, 3, (fc0.vertices.shape[0], 1), generator=gen).float() * 0.5)
    w0 = 0.4 * torch.randn(fc0.indices.shape[0], 21, generator=gen)
    tbl = fc_cases.tables(fc_cases.load("ref_flexicubes.npz"))
    got = {}
    for which in ("ours", "oracle"):
        sdf, w = sdf0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
        if which == "ours":
            fc = fc0.replace(sdf_values=sdf, alpha=w[:, :8], beta=w[:, 8:20], gamma=w[:, 20:])
            mesh, l_dev = fc.dual_marching_cubes()
            mv, mf, ent = mesh.vertices, mesh.indices, fc.compute_entropy()
        else:
            mv, mf, l_dev = OF.dual_marching_cubes(fc0.vertices, sdf, fc0.indices, res, w[:, :8], w[:, 8:20], w[:, 20:], tbl)
            ent = OF.entropy(sdf, fc0.indices, tbl)
        g = torch.autograd.grad(mv.square().sum() + l_dev.sum() + ent, [sdf, w])
        got[which] = [x.detach().numpy() for x in (mf, mv, l_dev, ent, *g)]
    a, b = got["ours"], got["oracle"]
    assert np.array_equal(a[0], b[0]) and a[0].shape[0] > 0
    for x, y in zip(a[1:], b[1:]):
        assert np.isfinite(x).all() and np.isfinite(y).all()
        assert np.abs(x - y).max(initial=0.0) <= 1e-4 * max(1.0, np.abs(y).max(initial=0.0))
