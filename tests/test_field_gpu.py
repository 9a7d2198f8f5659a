"""GaussianField glue against tests/golden/ref_field.npz, which scripts/make_golden.py
section I produces by exec'ing the reference's own method bodies (rfstudio/model/geosplat.py:520-674) on its own
TriangleMesh / MGAdapter / HashEncoding (torch backend) / MLP code: outputs and gradients of both sampling paths."""
import os

import numpy as np
import pytest
import torch

from geosplatting_b200 import encoding as E
from geosplatting_b200.field import GaussianField, get_rotation_from_relative_vectors, rot2quat

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load_field(g):
    def enc(tag, layers, act):
        e = E.HashEncoding(E.MLP(layers, activation=act), grad_scaling=16.0, max_res=4096, log2_hashmap_size=9)
        with torch.no_grad():
            e.hash_table.copy_(torch.from_numpy(g[f"{tag}_table"]))
            for k, w in enumerate(e.mlp.weights):
                w.copy_(torch.from_numpy(g[f"{tag}_w{k}"]))
        return e
    return GaussianField(enc("kd", [32, 32, 32, 3], "sigmoid"), enc("ks", [32, 32, 2], "none"),
                         enc("z", [32, 32, 1], "none")).to(DEV)


def _close(a, b, tol, name):
    a = a.detach().cpu().numpy()
    err = np.abs(a - b).max()
    assert err <= tol * max(1.0, np.abs(b).max()), (name, err, np.abs(b).max())


@pytest.mark.parametrize("path", ["vertex", "face"])
def test_sampling_paths_match_reference_code(path):
    g = np.load(os.path.join(GOLD, "ref_field.npz"))
    field = _load_field(g)
    verts = torch.from_numpy(g["vertices"]).to(DEV).requires_grad_(True)
    faces = torch.from_numpy(g["indices"]).to(DEV)
    guess = torch.from_numpy(g["initial_guess"]).to(DEV)
    scale = float(g["scale"])
    if path == "vertex":
        normals, areas = field.get_patches(verts, faces)
        _close(normals, g["patch_normals"], 1e-6, "patch normals")
        _close(areas, g["patch_areas"], 1e-6, "patch areas")
        sp, at = field.get_gaussians_from_vertex(0.0, 0.0, scale, verts, faces, guess)
        p = "v_"
    else:
        sp, at, offsets = field.get_gaussians_from_face(verts, faces, 0.0, 0.0, scale=scale, initial_guess=guess)
        _close(offsets, g["f_offsets"], 5e-4, "offsets")
        p = "f_"
    outs = dict(means=sp.means, scales=sp.scales, quats=sp.quats, opacities=sp.opacities, kd=at.kd, ks=at.ks,
                normals=at.normals)
    # Geometry that does not pass through a field is held to fp32 rounding.  What does (kd, ks, and the means through
    # sigmoid(z)) sees positions that differ from the reference run's by an ulp (6e-8), which the finest level
    # (resolution 4096) of this deliberately rough O(1) test table turns into ~1e-4 of feature difference.
    for k, v in outs.items():
        _close(v, g[p + k], 5e-4 if k in ("kd", "ks", "means") else 2e-6, k)
    loss = sum((v * torch.from_numpy(g["cot_" + p + k]).to(DEV)).sum() for k, v in outs.items())
    gv, gkd, gz = torch.autograd.grad(loss, [verts, field.kd_enc.hash_table, field.z_enc.hash_table])
    tag = "vertexpath" if path == "vertex" else "facepath"
    # gradients through sqrt(area)/normalisation chains of a 42-vertex mesh: 1e-4 of the max
    # d field / d x is piecewise constant per cell (O(resolution) steps on this rough table): a sample within an ulp
    # of a cell wall of some level flips its contribution, so the position gradient is compared in the L2 sense
    ref_v = g[f"{tag}_grad_vertices"]
    rel = np.linalg.norm(gv.cpu().numpy() - ref_v) / np.linalg.norm(ref_v)
    assert rel <= 1e-2, ("v_vertices", rel)
    _close(gv, ref_v, 2e-2, "v_vertices")
    _close(gkd, g[f"{tag}_grad_kd_table"], 1e-3, "v_kd_table")
    _close(gz, g[f"{tag}_grad_z_table"], 1e-3, "v_z_table")


def test_jitter_and_opposite_normals():
    g = np.load(os.path.join(GOLD, "ref_field.npz"))
    field = _load_field(g)
    verts = torch.from_numpy(g["vertices"]).to(DEV)
    faces = torch.from_numpy(g["indices"]).to(DEV)
    guess = torch.from_numpy(g["initial_guess"]).to(DEV)
    sp, at, _ = field.get_gaussians_from_face(verts, faces, 0.01, 0.02, scale=0.9, initial_guess=guess)
    assert at.kd_jitter.shape == at.kd.shape and at.ks_jitter.shape == at.ks.shape
    assert float((at.kd_jitter - at.kd).abs().max()) > 0
    # a normal exactly opposite to +z takes the reference's noise branch and still yields a proper rotation
    b = torch.tensor([[0.0, 0.0, -1.0], [0.0, 0.0, 1.0], [0.6, 0.0, 0.8]], device=DEV)
    torch.manual_seed(0)                                          # the nudge is drawn from the global generator
    R = get_rotation_from_relative_vectors(torch.tensor([0.0, 0.0, 1.0], device=DEV), b)
    z = R @ torch.tensor([0.0, 0.0, 1.0], device=DEV)
    assert bool(torch.isfinite(R).all())
    assert float((z[1:] - b[1:]).abs().max()) < 1e-5              # regular cases: exact
    # opposite vectors: the nudged source maps onto b only roughly -- the reference's formula keeps eps = 1e-6 in a
    # denominator that is |nudge|^2 ~ 1e-5 here, so the result depends on the draw (0.04 .. 0.13 over seeds)
    assert float((z[0] - b[0]).abs().max()) < 0.25
    assert float((R[1:].transpose(-1, -2) @ R[1:] - torch.eye(3, device=DEV)).abs().max()) < 1e-5
    q = rot2quat(R)
    assert float((q[1:].norm(dim=-1) - 1).abs().max()) < 1e-5 and bool(torch.isfinite(q).all())
