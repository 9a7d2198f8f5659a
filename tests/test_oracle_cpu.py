"""CPU checks of the oracle itself (no GPU): the C restatement against a dense fp64 torch restatement
with autograd, plus structural properties of the binning."""
import numpy as np
import pytest
import torch

from geosplatting_b200 import scenes
from oracle import raster as R
from oracle import raster_torch as RT
from tests.helpers import oracle_camera, to_np


def _scene(n=300, w=64, h=64, seed=0):
    g = to_np(scenes.random_gaussians(n, seed=seed, scale_lo=0.02, scale_hi=0.2))
    cam = scenes.look_at_camera((1.5, 1.0, 2.0), w, h)
    return g, oracle_camera(cam)


@pytest.mark.parametrize("mode", ["antialiased", "classic"])
def test_forward_matches_dense_fp64(mode):
    g, cam = _scene()
    render, alpha, info = R.rasterization(g["means"], g["quats"], g["scales"], g["opacities"], g["colors"], cam,
                                          rasterize_mode=mode)
    t = lambda a: torch.tensor(a, dtype=torch.float64)
    _, _, depths, _, _ = R.project_fwd(g["means"], g["quats"], g["scales"], cam, antialiased=mode == "antialiased")
    order = torch.tensor(np.argsort(depths, kind="stable"))
    img, a = RT.rasterize_dense(t(g["means"]), t(g["quats"]), t(g["scales"]), t(g["opacities"]), t(g["colors"]),
                                t(cam.viewmat), cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height,
                                antialiased=mode == "antialiased", order=order)
    ok = ~info["fragile"]
    assert ok.mean() > 0.99
    assert np.abs(img.numpy() - render)[ok].max() < 1e-5
    assert np.abs(a.numpy() - alpha[..., 0])[ok].max() < 1e-5


@pytest.mark.parametrize("mode", ["antialiased", "classic"])
def test_backward_matches_autograd(mode):
    g, cam = _scene()
    aa = mode == "antialiased"
    render, alpha, info = R.rasterization(g["means"], g["quats"], g["scales"], g["opacities"], g["colors"], cam,
                                          rasterize_mode=mode)
    rng = np.random.default_rng(1)
    H, W = cam.height, cam.width
    vr = rng.normal(size=(H, W, 3)).astype(np.float32)
    va = rng.normal(size=(H, W, 1)).astype(np.float32)
    # fragile pixels may take a different discrete branch in fp64: give them zero cotangent
    vr[info["fragile"]] = 0
    va[info["fragile"]] = 0
    grads = R.rasterization_bwd(g["means"], g["quats"], g["scales"], g["opacities"], g["colors"], cam, info, alpha,
                                vr, va, rasterize_mode=mode)
    t = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    tm, tq, ts, to, tc = t(g["means"]), t(g["quats"]), t(g["scales"]), t(g["opacities"]), t(g["colors"])
    _, _, depths, _, _ = R.project_fwd(g["means"], g["quats"], g["scales"], cam, antialiased=aa)
    order = torch.tensor(np.argsort(depths, kind="stable"))
    img, a = RT.rasterize_dense(tm, tq, ts, to, tc, torch.tensor(cam.viewmat, dtype=torch.float64), cam.fx, cam.fy,
                                cam.cx, cam.cy, W, H, antialiased=aa, order=order)
    ((img * torch.tensor(vr, dtype=torch.float64)).sum() + (a * torch.tensor(va[..., 0], dtype=torch.float64)).sum()).backward()
    for name, ours, ref in zip(["means", "quats", "scales", "opacities", "colors"], grads,
                               [tm.grad, tq.grad, ts.grad, to.grad, tc.grad]):
        ref = ref.numpy()
        scale = np.abs(ref).max()
        assert np.abs(ours - ref).max() <= 2e-4 * scale + 1e-6, name


def test_binning_structure():
    g, cam = _scene(n=500, w=100, h=70)  # ragged: not a multiple of 16
    _, _, info = R.rasterization(g["means"], g["quats"], g["scales"], g["opacities"], g["colors"], cam)
    ids, flat, off = info["isect_ids"], info["flatten_ids"], info["isect_offsets"].reshape(-1)
    assert (np.diff(ids) >= 0).all()                      # sorted by (tile, depth)
    assert off[0] == 0 and (np.diff(off) >= 0).all() and off[-1] <= len(ids)
    assert info["tiles_per_gauss"].sum() == len(ids)
    tiles = (ids >> 32)
    for t in (0, 7, len(off) - 1):
        lo, hi = off[t], (off[t + 1] if t + 1 < len(off) else len(ids))
        assert (tiles[lo:hi] == t).all()
    # depth order inside a tile, ties keep generation (Gaussian id) order
    d = info["depths"][flat]
    same_tile = tiles[1:] == tiles[:-1]
    assert (d[1:][same_tile] >= d[:-1][same_tile]).all()


def test_empty_and_culled_inputs():
    cam = oracle_camera(scenes.look_at_camera((0, 0, 2.5), 32, 32))
    z = np.zeros((0, 3), np.float32)
    render, alpha, info = R.rasterization(z, np.zeros((0, 4), np.float32), z, np.zeros(0, np.float32), z, cam)
    assert render.shape == (32, 32, 3) and not render.any() and not alpha.any()
    # one Gaussian behind the camera, one far off-screen, one visible
    means = np.array([[0, 0, 5.0], [50, 0, 0], [0, 0, 0]], np.float32)
    quats = np.tile(np.array([1, 0, 0, 0], np.float32), (3, 1))
    scales = np.full((3, 3), 0.1, np.float32)
    render, alpha, info = R.rasterization(means, quats, scales, np.full(3, 0.9, np.float32),
                                          np.ones((3, 3), np.float32), cam)
    assert list(info["gaussian_ids"]) == [2]
    assert alpha.max() > 0.5
