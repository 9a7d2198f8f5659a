"""Edge cases of the per-view path on the host build of the whole library (tests/emu SIMT mode), against the C oracle:
empty scene, everything culled, one-pixel and ragged images, splats larger than the image, saturated opacities, depth
ties, degenerate quaternions / scales.  The same cases run under AddressSanitizer in scripts/memcheck_host.sh."""
import ctypes as C

import numpy as np
import pytest
import torch

from geosplatting_b200 import _lib, scenes
from geosplatting_b200.rasterization import make_camera
from oracle import raster as OR
from tests.emu import build as emu
from tests.helpers import oracle_camera


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


@pytest.fixture(scope="module")
def lib():
    so = emu.build(*emu.all_kernel_files(), simt=True)
    so.gsb_view_bytes.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    return so


def _raster(lib, g, cam, antialiased=True):
    """projection -> two-stage binning -> compositing through the stage entry points; (render, alpha, flatten_ids, radii)."""
    N = g["means"].shape[0]
    W, H = cam.width, cam.height
    i32, i64, sz = C.c_int32, C.c_int64, C.c_size_t
    gc = make_camera(cam.view_matrix, cam.intrinsic_matrix, W, H, antialiased=antialiased)
    radii, tpg = np.zeros(N, np.int32), np.zeros(N, np.int32)
    means2d, depths = np.zeros((N, 2), np.float32), np.zeros(N, np.float32)
    conics, comps = np.zeros((N, 3), np.float32), np.zeros(N, np.float32)
    assert lib.gsb_project_fwd(i32(N), _p(g["means"]), _p(g["quats"]), _p(g["scales"]), C.byref(gc), _p(radii),
                               _p(means2d), _p(depths), _p(conics), _p(comps), _p(tpg), None) == 0
    tw, th = (W + 15) // 16, (H + 15) // 16
    nb = sz(0)
    M = 0
    order, cum, total = np.zeros(max(N, 1), np.int32), np.zeros(max(N, 1), np.int64), np.zeros(1, np.int64)
    if N:
        assert lib.gsb_bin2_workspace_bytes(i32(N), i64(0), C.byref(nb)) == 0
        ws = np.full(nb.value + 256, 0xFF, np.uint8)
        assert lib.gsb_bin2_count(i32(N), _p(depths), _p(tpg), _p(order), _p(cum), _p(total), _p(ws), sz(ws.size),
                                  None) == 0
        M = int(total[0])
    flat, off = np.zeros(M, np.int32), np.zeros(tw * th, np.int32)
    if M:
        assert lib.gsb_bin2_workspace_bytes(i32(0), i64(M), C.byref(nb)) == 0
        ws = np.full(nb.value + 256, 0xFF, np.uint8)
        assert lib.gsb_bin2_sort(i32(N), i64(M), _p(means2d), _p(radii), _p(order), _p(cum), C.byref(gc), _p(flat),
                                 _p(off), _p(ws), sz(ws.size), None) == 0, lib.gsb_last_error()
    opac = (g["opacities"] * (comps if antialiased else 1.0)).astype(np.float32)
    assert lib.gsb_composite_workspace_bytes(i64(N), i64(M), i32(W), i32(H), C.byref(nb)) == 0
    ws = np.full(nb.value + 256, 0xFF, np.uint8)
    render, alphas = np.zeros((H, W, 3), np.float32), np.zeros((H, W), np.float32)
    last = np.zeros((H, W), np.int32)
    assert lib.gsb_composite_fwd(i32(W), i32(H), i32(3), i64(N), _p(means2d), _p(conics), _p(g["colors"]), _p(opac),
                                 i32(0), None, None, _p(off), _p(flat), i64(M), _p(render), _p(alphas), _p(last), _p(ws),
                                 sz(ws.size), None) == 0, lib.gsb_last_error()
    return render, alphas, flat, radii


def _check(lib, g, cam, antialiased=True):
    g = {k: np.ascontiguousarray(v, np.float32) for k, v in g.items()}
    render, alphas, flat, radii = _raster(lib, g, cam, antialiased)
    o_render, o_alpha, info = OR.rasterization(g["means"], g["quats"], g["scales"], g["opacities"], g["colors"],
                                               oracle_camera(cam), rasterize_mode="antialiased" if antialiased else "classic")
    assert np.array_equal(radii, info["radii_unpacked"])
    assert np.array_equal(flat, info["gaussian_ids"][info["flatten_ids"]] if flat.size else flat)
    ok = ~info["fragile"]
    assert np.isfinite(render).all() and np.isfinite(alphas).all()
    if ok.any():
        assert np.abs(render - o_render)[ok].max() <= 1e-4 and np.abs(alphas - o_alpha[..., 0])[ok].max() <= 1e-4
    return flat.size


def _scene(n, seed=0, **kw):
    return {k: v.numpy() for k, v in scenes.random_gaussians(n, seed=seed, **kw).items()}


@pytest.mark.parametrize("res", [(1, 1), (17, 3), (16, 16), (33, 47)])
def test_tiny_and_ragged_images(lib, res):
    cam = scenes.orbit_cameras(1, res[0], res[1], seed=2)[0]
    assert _check(lib, _scene(300, seed=res[0], extent=0.3, scale_lo=0.02, scale_hi=0.3), cam) > 0


def test_empty_scene_and_everything_culled(lib):
    cam = scenes.orbit_cameras(1, 40, 24, seed=3)[0]
    g = _scene(8)
    assert _check(lib, {k: v[:0] for k, v in g.items()}, cam) == 0
    far = dict(g)
    far["means"] = g["means"] + 100.0 * np.asarray(cam.position, np.float32)          # all behind the camera
    assert _check(lib, far, cam) == 0


def test_splats_larger_than_the_image_and_saturated_opacities(lib):
    cam = scenes.orbit_cameras(1, 48, 32, seed=4)[0]
    g = _scene(60, seed=9, extent=0.2, scale_lo=0.5, scale_hi=3.0)
    g["opacities"] = np.where(np.arange(60) % 2 == 0, 1.0, 1e-4).astype(np.float32)      # fully opaque / below 1/255
    M = _check(lib, g, cam)
    assert M >= 30 * 6                                                                    # opaque ones cover every tile


def test_depth_ties_and_degenerate_inputs(lib):
    """Identical Gaussians (depth ties keep index order in both sorts); zero quaternions and zero scales are whatever
    the oracle says they are (same radii, same lists)."""
    cam = scenes.orbit_cameras(1, 64, 48, seed=5)[0]
    g = _scene(40, seed=11, extent=0.4, scale_lo=0.05, scale_hi=0.2)
    for k in g:
        g[k][20:] = g[k][:20]                                                             # 20 exact duplicates
    _check(lib, g, cam)
    _check(lib, g, cam, antialiased=False)
    d = _scene(30, seed=12, extent=0.4, scale_lo=0.05, scale_hi=0.2)
    d["scales"][:5] = 0.0
    d["scales"][5:10, 2] = 1e-12
    d["quats"][10:15] = 0.0
    d["quats"][15:20] *= 1e-20
    _check(lib, d, cam)


def test_random_scenes_property(lib):
    """Seeded sweep over scene size, image shape, splat size, extent and rasterize mode: radii and sorted lists
    identical to the oracle's, image within 1e-4 on stable pixels -- 24 draws."""
    rng = np.random.default_rng(2024)
    for _ in range(24):
        n = int(rng.integers(1, 400))
        W, H = int(rng.integers(1, 72)), int(rng.integers(1, 72))
        lo = float(10 ** rng.uniform(-2.5, -1.0))
        hi = lo * float(10 ** rng.uniform(0.0, 1.5))
        cam = scenes.orbit_cameras(1, W, H, seed=int(rng.integers(0, 1000)))[0]
        g = _scene(n, seed=int(rng.integers(0, 1000)), extent=float(rng.uniform(0.05, 1.5)), scale_lo=lo, scale_hi=hi)
        _check(lib, g, cam, antialiased=bool(rng.integers(0, 2)))
