"""FlexiCubes parity checks shared by the host-emulation run (tests/test_flexicubes_cpu.py) and the CUDA run
(tests/test_flexicubes_gpu.py): the host module + gsb_fc_* kernels against the fixtures the reference's own code
produced (tests/golden/ref_flexicubes*.npz) and against oracle/flexicubes.py on fresh seeded inputs."""
import os

import numpy as np
import torch

from geosplatting_b200 import flexicubes as FC
from oracle import flexicubes as OF

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# fp32 tolerances: positions / L_dev are a few dozen flops from the inputs; gradients are sums of float atomics whose
# order differs from torch's index_add (relative to the largest entry)
TOL_POS, TOL_GRAD = 2e-6, 1e-4


def load(name):
    return {k: v for k, v in np.load(os.path.join(GOLD, name), allow_pickle=False).items()}


def tables(g):
    return {k[4:]: g[k] for k in g if k.startswith("tbl_")}


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def check_smooth_fixture(dev):
    """ref_flexicubes.npz: 10^3 bumpy sphere driven exactly as GeoSplatter.get_geometry does (geosplat.py:751-769)."""
    g = load("ref_flexicubes.npz")
    R, scale = int(g["resolution"]), float(g["scale"])
    fc0 = FC.FlexiCubes.from_resolution(R, random_sdf=False, scale=scale, device=dev)
    assert np.array_equal(fc0.indices.cpu().numpy(), g["cube_indices"])
    assert np.array_equal(fc0.vertices.cpu().numpy(), g["grid_vertices"])
    sdf = torch.from_numpy(g["sdf"]).to(dev).requires_grad_(True)
    deform = torch.from_numpy(g["deform"]).to(dev).requires_grad_(True)
    weights = torch.from_numpy(g["weights"]).to(dev).requires_grad_(True)
    verts = fc0.vertices + deform.tanh() * (0.5 * scale / R)
    fc = fc0.replace(vertices=verts, sdf_values=sdf, alpha=weights[:, :8], beta=weights[:, 8:20], gamma=weights[:, 20:])
    mesh, l_dev = fc.dual_marching_cubes()
    ent = fc.compute_entropy()
    assert mesh.indices.dtype == torch.int64
    assert np.array_equal(mesh.indices.cpu().numpy(), g["mesh_indices"])          # same faces in the same ORDER
    assert np.abs(mesh.vertices.detach().cpu().numpy() - g["mesh_vertices"]).max() <= TOL_POS
    assert l_dev.shape == g["L_dev"].shape
    assert np.abs(l_dev.detach().cpu().numpy() - g["L_dev"]).max() <= TOL_POS
    assert abs(float(ent.detach()) - float(g["entropy"])) <= 2e-6
    loss = (mesh.vertices * torch.from_numpy(g["cot_vertices"]).to(dev)).sum() + l_dev.mean() * 0.5 + ent * 0.3
    v_sdf, v_def, v_w = torch.autograd.grad(loss, [sdf, deform, weights])
    for a, name in ((v_sdf, "v_sdf"), (v_def, "v_deform"), (v_w, "v_weights")):
        assert _rel(a.cpu().numpy(), g[name]) <= TOL_GRAD, (name, _rel(a.cpu().numpy(), g[name]))
    # the regulariser exactly as get_geometry assembles it
    reg = l_dev.mean() * 0.5 + weights[:, :20].abs().mean() * 0.1 + ent * 0.3
    assert bool(torch.isfinite(reg))


def check_rough_fixture(dev):
    """ref_flexicubes_rough.npz: random SDF on a 7 x 6 x 5 grid, inverted ambiguous cases, up to 4 dual vertices."""
    g = load("ref_flexicubes_rough.npz")
    r = tuple(int(x) for x in g["resolution"])
    fr = FC.FlexiCubes.from_resolution(*r, random_sdf=False, scale=1.0, device=dev)
    assert np.array_equal(fr.indices.cpu().numpy(), g["cube_indices"])
    assert np.array_equal(fr.vertices.cpu().numpy(), g["grid_vertices"])
    sdf = torch.from_numpy(g["sdf"]).to(dev).requires_grad_(True)
    w = torch.from_numpy(g["weights"]).to(dev).requires_grad_(True)
    fc = fr.replace(sdf_values=sdf, alpha=w[:, :8], beta=w[:, 8:20], gamma=w[:, 20:])
    mesh, l_dev = fc.dual_marching_cubes()
    assert np.array_equal(mesh.indices.cpu().numpy(), g["mesh_indices"])
    assert np.abs(mesh.vertices.detach().cpu().numpy() - g["mesh_vertices"]).max() <= 1e-5
    assert np.abs(l_dev.detach().cpu().numpy() - g["L_dev"]).max() <= 1e-5
    assert abs(float(fc.compute_entropy().detach()) - float(g["entropy"])) <= 2e-6
    loss = (mesh.vertices * torch.from_numpy(g["cot_vertices"]).to(dev)).sum() + l_dev.mean()
    v_sdf, v_w = torch.autograd.grad(loss, [sdf, w])
    assert _rel(v_sdf.cpu().numpy(), g["v_sdf"]) <= TOL_GRAD
    assert _rel(v_w.cpu().numpy(), g["v_weights"]) <= TOL_GRAD


def sphere_case(res, seed, dev, noise=0.0):
    """A seeded bumpy-sphere grid (optionally with SDF noise that creates ambiguous cubes) on `dev`."""
    gen = torch.Generator().manual_seed(seed)
    fc0 = FC.FlexiCubes.from_resolution(*res, random_sdf=False, scale=0.9, device=dev)
    gv = fc0.vertices.cpu()
    sdf = gv.norm(dim=-1, keepdim=True) - 0.55 + 0.08 * torch.sin(5.0 * gv[:, :1]) * torch.cos(4.0 * gv[:, 1:2])
    sdf = sdf + noise * torch.randn(sdf.shape, generator=gen)
    deform = 0.3 * torch.randn(gv.shape, generator=gen)
    weights = 0.3 * torch.randn(fc0.indices.shape[0], 21, generator=gen)
    return fc0, sdf, deform, weights


def check_against_oracle(dev, res=(12, 9, 7), seed=3, noise=0.05):
    """Fresh seeded input: same faces (order included), vertices, L_dev, entropy and gradients as oracle/flexicubes.py."""
    fc0, sdf0, deform0, weights0 = sphere_case(res, seed, dev, noise)
    tbl = tables(load("ref_flexicubes.npz"))
    out = {}
    for which in ("ours", "oracle"):
        d = dev if which == "ours" else "cpu"
        sdf = sdf0.clone().to(d).requires_grad_(True)
        deform = deform0.clone().to(d).requires_grad_(True)
        w = weights0.clone().to(d).requires_grad_(True)
        verts = fc0.vertices.to(d) + deform.tanh() * (0.5 * 0.9 / max(res))
        if which == "ours":
            fc = fc0.replace(vertices=verts, sdf_values=sdf, alpha=w[:, :8], beta=w[:, 8:20], gamma=w[:, 20:])
            mesh, l_dev = fc.dual_marching_cubes()
            mv, mf, ent = mesh.vertices, mesh.indices, fc.compute_entropy()
        else:
            mv, mf, l_dev = OF.dual_marching_cubes(verts, sdf, fc0.indices.cpu(), res, w[:, :8], w[:, 8:20], w[:, 20:],
                                                   tbl)
            ent = OF.entropy(sdf, fc0.indices.cpu(), tbl)
        cot = torch.randn(mv.shape, generator=torch.Generator().manual_seed(seed + 1)).to(d)
        loss = (mv * cot).sum() + l_dev.mean() * 0.5 + ent * 0.3
        grads = torch.autograd.grad(loss, [sdf, deform, w])
        out[which] = [x.detach().cpu().numpy() for x in (mf, mv, l_dev, ent, *grads)]
    a, b = out["ours"], out["oracle"]
    assert np.array_equal(a[0], b[0])
    assert np.abs(a[1] - b[1]).max() <= 1e-5 and np.abs(a[2] - b[2]).max() <= 1e-5 and abs(a[3] - b[3]) <= 2e-6
    for x, y, name in zip(a[4:], b[4:], ("sdf", "deform", "weights")):
        assert _rel(x, y) <= TOL_GRAD, (name, _rel(x, y))
    return a[0].shape[0]
