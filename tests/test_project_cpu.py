"""The EWA projection kernels (csrc/project_fwd.cu, project_bwd.cu) without a GPU: the real kernel source compiled for
the host by tests/emu (threads one after another, IEEE fp32 without contraction -- the flags the CUDA build gives
project_fwd.cu) against the C oracle on BASELINE config 1 (10 k random Gaussians, one camera, 256 x 256).
The forward decides tile / bin indices, so it is held to BIT equality; the CUDA build of the same source is held to the
same oracle by tests/test_raster_gpu.py."""
import ctypes as C

import numpy as np
import pytest

from geosplatting_b200 import scenes
from geosplatting_b200.rasterization import make_camera
from oracle import raster as R
from tests.emu import build as emu
from tests.helpers import oracle_camera, rel_l2


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


@pytest.fixture(scope="module")
def libs():
    return emu.build("project_fwd"), emu.build("project_bwd")


def _scene(n, res, seed=0, **kw):
    g = {k: v.numpy() for k, v in scenes.random_gaussians(n, seed=seed, **kw).items()}
    cam = scenes.orbit_cameras(1, res[0], res[1], seed=1)[0]
    return g, cam


def _project(lib, g, cam, antialiased):
    N = g["means"].shape[0]
    gc = make_camera(cam.view_matrix, cam.intrinsic_matrix, cam.width, cam.height, antialiased=antialiased)
    radii, tpg = np.zeros(N, np.int32), np.zeros(N, np.int32)
    means2d, depths = np.zeros((N, 2), np.float32), np.zeros(N, np.float32)
    conics, comps = np.zeros((N, 3), np.float32), np.zeros(N, np.float32)
    rc = lib.gsb_project_fwd(C.c_int32(N), _p(g["means"]), _p(g["quats"]), _p(g["scales"]), C.byref(gc), _p(radii),
                             _p(means2d), _p(depths), _p(conics), _p(comps), _p(tpg), None)
    assert rc == 0, lib.gsb_last_error()
    return gc, radii, means2d, depths, conics, comps, tpg


@pytest.mark.parametrize("antialiased", [False, True])
def test_config1_projection_is_bit_exact(libs, antialiased):
    g, cam = _scene(10_000, (256, 256))
    _, radii, means2d, depths, conics, comps, tpg = _project(libs[0], g, cam, antialiased)
    o_radii, o_means2d, o_depths, o_conics, o_comps = R.project_fwd(g["means"], g["quats"], g["scales"],
                                                                    oracle_camera(cam), antialiased=antialiased)
    assert np.array_equal(radii, o_radii) and int((radii > 0).sum()) > 5000
    vis = radii > 0
    for a, b in ((means2d, o_means2d), (depths, o_depths), (conics, o_conics), (comps, o_comps)):
        assert np.array_equal(a[vis].view(np.uint32), b[vis].view(np.uint32))          # bit for bit
    o_tpg = R.bin_sort(o_means2d, o_radii, o_depths, cam.width, cam.height)[0]
    assert np.array_equal(tpg, o_tpg)                                                   # tiles touched per Gaussian


def test_ragged_resolution_big_splats_and_culling(libs):
    """Width / height not multiples of 16, splats larger than the image, Gaussians behind the camera."""
    g, cam = _scene(3000, (200, 120), seed=3, extent=3.5, scale_lo=0.01, scale_hi=1.5)
    _, radii, means2d, depths, conics, comps, tpg = _project(libs[0], g, cam, True)
    o_radii, o_means2d, o_depths, o_conics, o_comps = R.project_fwd(g["means"], g["quats"], g["scales"],
                                                                    oracle_camera(cam), antialiased=True)
    assert np.array_equal(radii, o_radii) and 0 < int((radii == 0).sum()) < 3000
    vis = radii > 0
    assert np.array_equal(means2d[vis].view(np.uint32), o_means2d[vis].view(np.uint32))
    assert np.array_equal(conics[vis].view(np.uint32), o_conics[vis].view(np.uint32))
    assert np.array_equal(tpg, R.bin_sort(o_means2d, o_radii, o_depths, cam.width, cam.height)[0])
    assert (tpg[~vis] == 0).all()


@pytest.mark.parametrize("antialiased", [False, True])
def test_projection_backward_matches_oracle(libs, antialiased):
    g, cam = _scene(4000, (256, 256), seed=2)
    N = 4000
    gc, radii, *_ = _project(libs[0], g, cam, antialiased)
    rng = np.random.default_rng(0)
    v_means2d = rng.standard_normal((N, 2)).astype(np.float32)
    v_depths = rng.standard_normal(N).astype(np.float32)
    v_conics = rng.standard_normal((N, 3)).astype(np.float32)
    v_comps = rng.standard_normal(N).astype(np.float32)
    v_means, v_quats, v_scales = (np.full((N, k), np.nan, np.float32) for k in (3, 4, 3))
    rc = libs[1].gsb_project_bwd(C.c_int32(N), _p(g["means"]), _p(g["quats"]), _p(g["scales"]), C.byref(gc), _p(radii),
                                 _p(v_means2d), _p(v_depths), _p(v_conics), _p(v_comps if antialiased else None),
                                 _p(v_means), _p(v_quats), _p(v_scales), None, None, None, C.c_int32(0), None)
    assert rc == 0, libs[1].gsb_last_error()
    o = R.project_bwd(g["means"], g["quats"], g["scales"], oracle_camera(cam), radii, v_means2d, v_depths, v_conics,
                      v_comps, antialiased=antialiased)
    for a, b, name in zip((v_means, v_quats, v_scales), o, ("means", "quats", "scales")):
        assert np.isfinite(a).all() and (a[radii == 0] == 0).all(), name
        assert rel_l2(a, b) <= 2e-5, (name, rel_l2(a, b))


# ---- tile binning (csrc/binsort.cu + the emit kernels of project_fwd.cu) on the host ------------------------------------
@pytest.fixture(scope="module")
def binlib():
    return emu.build("project_fwd", "binsort")


def _bin_both_ways(lib, gc, radii, means2d, depths, tpg):
    """(two-stage scheme gsb_bin2_*, gsplat's stage split gsb_isect_* + gsb_sort_pairs) -> flatten_ids, offsets each."""
    N = radii.shape[0]
    tw, th = (gc.width + 15) // 16, (gc.height + 15) // 16
    i32, i64, sz = C.c_int32, C.c_int64, C.c_size_t
    nb = sz(0)
    # two-stage
    assert lib.gsb_bin2_workspace_bytes(i32(N), i64(0), C.byref(nb)) == 0
    ws = np.full(nb.value, 0xFF, np.uint8)
    order, cum, total = np.zeros(N, np.int32), np.zeros(N, np.int64), np.zeros(1, np.int64)
    assert lib.gsb_bin2_count(i32(N), _p(depths), _p(tpg), _p(order), _p(cum), _p(total), _p(ws), sz(ws.size), None) == 0
    M = int(total[0])
    assert lib.gsb_bin2_workspace_bytes(i32(0), i64(M), C.byref(nb)) == 0
    ws = np.full(nb.value, 0xFF, np.uint8)
    flat2, off2 = np.zeros(M, np.int32), np.zeros(tw * th, np.int32)
    assert lib.gsb_bin2_sort(i32(N), i64(M), _p(means2d), _p(radii), _p(order), _p(cum), C.byref(gc), _p(flat2), _p(off2),
                             _p(ws), sz(ws.size), None) == 0, lib.gsb_last_error()
    # stage split
    assert lib.gsb_bin_workspace_bytes(i32(N), i64(M), C.byref(nb)) == 0
    ws = np.full(nb.value, 0xFF, np.uint8)
    cum1 = np.zeros(N, np.int64)
    assert lib.gsb_isect_scan(i32(N), _p(tpg), _p(cum1), _p(ws), sz(ws.size), None) == 0
    assert lib.gsb_isect_total(i32(N), _p(cum1), _p(total), None) == 0 and int(total[0]) == M
    keys, vals = np.zeros(M, np.int64), np.zeros(M, np.int32)
    assert lib.gsb_isect_tiles(i32(N), _p(means2d), _p(radii), _p(depths), _p(cum1), C.byref(gc), _p(keys), _p(vals),
                               None) == 0
    keys_s, vals_s, off1 = np.zeros(M, np.int64), np.zeros(M, np.int32), np.zeros(tw * th, np.int32)
    bits = 32 + int(np.floor(np.log2(tw * th))) + 1
    assert lib.gsb_sort_pairs(i64(M), i32(bits), _p(keys), _p(vals), _p(keys_s), _p(vals_s), _p(ws), sz(ws.size),
                              None) == 0
    assert lib.gsb_isect_offsets(i64(M), _p(keys_s), i32(1), i32(tw), i32(th), _p(off1), None) == 0
    return (flat2, off2), (keys_s, vals_s, off1)


@pytest.mark.parametrize("res,n,kw", [((256, 256), 10_000, {}), ((200, 120), 3000, dict(extent=3.5, scale_hi=1.5)),
                                      ((16, 16), 50, {}), ((64, 48), 0, {})])
def test_bin_indices_are_bit_exact_on_the_host(libs, binlib, res, n, kw):
    """BASELINE's "bit-exact tile/bin indices": sorted keys, Gaussian ids and per-tile offsets of both binning schemes
    equal the oracle's; ragged resolution, one-tile image and the empty scene included."""
    g, cam = _scene(max(n, 1), res, seed=5, **kw)
    if n == 0:
        g = {k: v[:0] for k, v in g.items()}
    gc, radii, means2d, depths, conics, comps, tpg = _project(libs[0], g, cam, True) if n else (
        make_camera(cam.view_matrix, cam.intrinsic_matrix, cam.width, cam.height, antialiased=True),
        np.zeros(0, np.int32), np.zeros((0, 2), np.float32), np.zeros(0, np.float32), None, None, np.zeros(0, np.int32))
    o_tpg, o_keys, o_vals, o_off = R.bin_sort(means2d, radii, depths, cam.width, cam.height)
    if n == 0:
        assert o_keys.size == 0
        return
    (flat2, off2), (keys_s, vals_s, off1) = _bin_both_ways(binlib, gc, radii, means2d, depths, tpg)
    assert np.array_equal(keys_s, o_keys) and np.array_equal(vals_s, o_vals) and np.array_equal(off1, o_off.reshape(-1))
    assert np.array_equal(flat2, o_vals) and np.array_equal(off2, o_off.reshape(-1))
    assert flat2.size == int(o_tpg.sum())


@pytest.mark.parametrize("res,n,kw,slack", [((256, 256), 10_000, {}, 1.3), ((320, 200), 6000, dict(extent=3.5, scale_hi=1.2), 1.0),
                                            ((16, 16), 50, {}, 2.0), ((64, 48), 1, {}, 1.0),
                                            ((1100, 1000), 4000, dict(extent=3.0, scale_hi=0.6), 1.1),     # 4 347 tiles: 8 warps
                                            ((2000, 1900), 4000, dict(extent=3.0, scale_hi=0.5), 1.2)])    # 14 875 tiles: 4 warps
def test_stable_tile_partition_equals_the_sort_on_the_host(res, n, kw, slack):
    """csrc/tilepart.cu (the batch driver's second binning stage: a stable partition of the depth-ordered pairs by tile
    -- chunk histograms, a scan, a warp-ranked scatter -- instead of a radix sort) against the radix-sort path
    gsb_bin2_sort on the same depth order: flatten_ids and per-tile offsets identical.  Several chunks of 16 384 pairs
    (big splats), a capacity larger than the count (the tail is never touched), one tile, one Gaussian, and the two
    smaller kernel shapes that images of more than 4 096 / 10 240 tiles select."""
    lib = emu.build("project_fwd", "binsort", "tilepart", simt=True)
    g, cam = _scene(n, res, seed=9, **kw)
    gc, radii, means2d, depths, conics, comps, tpg = _project(lib, g, cam, True)
    N = radii.shape[0]
    tw, th = (gc.width + 15) // 16, (gc.height + 15) // 16
    i32, i64, sz = C.c_int32, C.c_int64, C.c_size_t
    nb = sz(0)
    assert lib.gsb_bin2_workspace_bytes(i32(N), i64(0), C.byref(nb)) == 0
    ws = np.full(nb.value, 0xFF, np.uint8)
    order, cum, total = np.zeros(N, np.int32), np.zeros(N, np.int64), np.zeros(1, np.int64)
    assert lib.gsb_bin2_count(i32(N), _p(depths), _p(tpg), _p(order), _p(cum), _p(total), _p(ws), sz(ws.size), None) == 0
    M = int(total[0])
    assert lib.gsb_bin2_workspace_bytes(i32(0), i64(max(M, 1)), C.byref(nb)) == 0
    ws = np.full(nb.value, 0xFF, np.uint8)
    flat_ref, off_ref = np.zeros(M, np.int32), np.zeros(tw * th, np.int32)
    assert lib.gsb_bin2_sort(i32(N), i64(M), _p(means2d), _p(radii), _p(order), _p(cum), C.byref(gc), _p(flat_ref),
                             _p(off_ref), _p(ws), sz(ws.size), None) == 0, lib.gsb_last_error()
    cap = int(M * slack) + 7
    lib.gsb_tile_partition_bytes.restype = C.c_size_t
    need = lib.gsb_tile_partition_bytes(i64(cap), i32(tw * th))
    ws2 = np.full(need + 256, 0xFF, np.uint8)
    m_eff = np.asarray([min(M, cap)], np.int64)
    flat, off = np.full(cap, -7, np.int32), np.full(tw * th, -7, np.int32)
    assert lib.gsb_tile_partition_supported(i32(tw * th)) == int(tw * th <= 4096)     # the batch driver's policy
    assert lib.gsb_tile_partition_cap(i32(N), i64(cap), _p(m_eff), _p(means2d), _p(radii), _p(order), _p(cum), C.byref(gc),
                                      _p(flat), _p(off), _p(ws2), sz(ws2.size), None) == 0, lib.gsb_last_error()
    assert np.array_equal(off, off_ref)
    assert np.array_equal(flat[:M], flat_ref)
    assert (flat[M:] == -7).all()                       # nothing behind the count was written
    if n == 6000:
        assert M > 2 * 16384                            # the scatter really ran over several chunks
