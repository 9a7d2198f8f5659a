"""The compositing kernels (csrc/composite.cu: record packing, sub-list build with ballots and block barriers, LPT order,
warp-per-sub-rectangle forward, transposing backward) without a GPU: the real kernel source executed by the SIMT mode
of tests/emu (the threads of a block are fibers that rendezvous at __syncthreads and the *_sync warp primitives; shared
memory is real; only the two inline-PTX approximations are replaced by exp2f and 1/x) against the C oracle.
Same acceptance as the GPU suite (tests/test_raster_gpu.py): last contributor indices identical and render / alpha
within 1e-4 (north_star) on pixels whose discrete decisions are not within 2e-5 of flipping; gradients to 2e-4 rel. L2."""
import ctypes as C

import numpy as np
import pytest

from geosplatting_b200 import scenes
from oracle import raster as R
from tests.emu import build as emu
from tests.helpers import oracle_camera, rel_l2


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


@pytest.fixture(scope="module")
def lib():
    return emu.build("composite", simt=True)


def _inputs(n, res, seed, antialiased=True, **kw):
    g = {k: v.numpy() for k, v in scenes.random_gaussians(n, seed=seed, **kw).items()}
    cam = scenes.orbit_cameras(1, res[0], res[1], seed=1)[0]
    radii, means2d, depths, conics, comps = R.project_fwd(g["means"], g["quats"], g["scales"], oracle_camera(cam),
                                                          antialiased=antialiased)
    _, _, flatten_ids, offsets = R.bin_sort(means2d, radii, depths, cam.width, cam.height)
    opac = (g["opacities"] * comps).astype(np.float32)
    return cam, means2d, conics, g["colors"].astype(np.float32), opac, flatten_ids, offsets.reshape(-1)


def _composite(lib, W, H, means2d, conics, colors, opac, flatten_ids, offsets, background=None):
    N, M, CH = means2d.shape[0], flatten_ids.shape[0], colors.shape[1]
    nb = C.c_size_t(0)
    assert lib.gsb_composite_workspace_bytes(C.c_int64(N), C.c_int64(M), C.c_int32(W), C.c_int32(H), C.byref(nb)) == 0
    ws = np.full(nb.value + 256, 0xFF, np.uint8)
    render, alphas = np.zeros((H, W, CH), np.float32), np.zeros((H, W), np.float32)
    last_ids = np.zeros((H, W), np.int32)
    rc = lib.gsb_composite_fwd(C.c_int32(W), C.c_int32(H), C.c_int32(CH), C.c_int64(N), _p(means2d), _p(conics),
                               _p(colors), _p(opac), C.c_int32(0), None, _p(background), _p(offsets), _p(flatten_ids),
                               C.c_int64(M), _p(render), _p(alphas), _p(last_ids), _p(ws), C.c_size_t(ws.size), None)
    assert rc == 0, lib.gsb_last_error()
    return render, alphas, last_ids, ws


@pytest.mark.parametrize("res,n,kw", [((64, 48), 1500, dict(scale_lo=0.01, scale_hi=0.08)),
                                      ((40, 25), 400, dict(extent=0.5, scale_lo=0.05, scale_hi=0.4)),
                                      ((256, 256), 10_000, {})])        # BASELINE configs[0]
def test_composite_forward_and_backward_on_the_host(lib, res, n, kw):
    cam, means2d, conics, colors, opac, flatten_ids, offsets = _inputs(n, res, seed=11, **kw)
    W, H = cam.width, cam.height
    bg = np.asarray([0.2, 0.5, 0.7], np.float32)
    render, alphas, last_ids, ws = _composite(lib, W, H, means2d, conics, colors, opac, flatten_ids, offsets, bg)
    o_render, o_alphas, o_last = R.composite_fwd(means2d, conics, colors, opac, offsets, flatten_ids, W, H, background=bg)
    ok = ~R.composite_fragile(means2d, conics, opac, offsets, flatten_ids, W, H)
    assert ok.mean() > 0.97 and float(o_alphas.max()) > 0.5
    assert np.array_equal(last_ids[ok], o_last[ok])
    assert np.abs(render - o_render)[ok].max() <= 1e-4 and np.abs(alphas - o_alphas)[ok].max() <= 1e-4
    mse = float(np.mean((render.astype(np.float64) - o_render) ** 2))              # over ALL pixels (PSNR parity)
    assert mse == 0 or 10 * np.log10(1.0 / mse) >= 70.0
    # backward on the forward's own state (alphas, last_ids, workspace), cotangents zeroed on the fragile pixels
    rng = np.random.default_rng(3)
    v_render = (rng.standard_normal(render.shape) * ok[..., None]).astype(np.float32)
    v_alphas = (rng.standard_normal(alphas.shape) * ok).astype(np.float32)
    N, M = means2d.shape[0], flatten_ids.shape[0]
    v_means2d, v_conics = np.zeros((N, 2), np.float32), np.zeros((N, 3), np.float32)
    v_colors, v_opac = np.zeros((N, 3), np.float32), np.zeros(N, np.float32)
    rc = lib.gsb_composite_bwd(C.c_int32(W), C.c_int32(H), C.c_int32(3), C.c_int64(N), _p(colors), _p(bg), _p(offsets),
                               C.c_int64(M), _p(alphas), _p(last_ids), _p(v_render), _p(v_alphas), _p(v_means2d),
                               _p(v_conics), _p(v_colors), _p(v_opac), _p(ws), None)
    assert rc == 0, lib.gsb_last_error()
    o = R.composite_bwd(means2d, conics, colors, opac, offsets, flatten_ids, W, H, o_alphas, o_last, v_render, v_alphas,
                        background=bg)
    for a, b, name in zip((v_means2d, v_conics, v_colors, v_opac), o, ("means2d", "conics", "colors", "opacities")):
        assert np.isfinite(a).all() and rel_l2(a, b) <= 2e-4, (name, rel_l2(a, b))


def test_empty_lists_and_one_pixel_image(lib):
    """No intersections at all (every tile list empty) renders the background; a 1 x 1 image works."""
    z2, z3 = np.zeros((4, 2), np.float32), np.zeros((4, 3), np.float32)
    bg = np.asarray([0.1, 0.2, 0.3], np.float32)
    render, alphas, last_ids, _ = _composite(lib, 33, 17, z2, z3, z3, np.zeros(4, np.float32), np.zeros(0, np.int32),
                                             np.zeros(3 * 2, np.int32), bg)
    assert np.all(alphas == 0) and np.allclose(render, bg)
    cam, means2d, conics, colors, opac, flatten_ids, offsets = _inputs(50, (1, 1), seed=2, extent=0.1, scale_hi=0.3)
    render, alphas, _, _ = _composite(lib, 1, 1, means2d, conics, colors, opac, flatten_ids, offsets)
    o_render, o_alphas, _ = R.composite_fwd(means2d, conics, colors, opac, offsets, flatten_ids, 1, 1)
    assert np.abs(render - o_render).max() <= 1e-4 and np.abs(alphas - o_alphas).max() <= 1e-4


@pytest.mark.parametrize("CH", [1, 4, 16])
def test_wide_channel_composite_on_the_host(lib, CH):
    """depth-only (1), RGB+ED (4) and the D = 14 G-buffer padded to 16 channels through the same
    kernels -- channels beyond the third are read from `colors` through the list entry, not from the packed record."""
    cam, means2d, conics, _, opac, flatten_ids, offsets = _inputs(1200, (56, 40), seed=21, scale_lo=0.01, scale_hi=0.1)
    W, H = cam.width, cam.height
    rng = np.random.default_rng(CH)
    colors = rng.random((means2d.shape[0], CH)).astype(np.float32)
    bg = rng.random(CH).astype(np.float32)
    render, alphas, last_ids, ws = _composite(lib, W, H, means2d, conics, colors, opac, flatten_ids, offsets, bg)
    o_render, o_alphas, o_last = R.composite_fwd(means2d, conics, colors, opac, offsets, flatten_ids, W, H, background=bg)
    ok = ~R.composite_fragile(means2d, conics, opac, offsets, flatten_ids, W, H)
    assert np.array_equal(last_ids[ok], o_last[ok]) and np.abs(render - o_render)[ok].max() <= 1e-4
    v_render = (rng.standard_normal(render.shape) * ok[..., None]).astype(np.float32)
    v_alphas = (rng.standard_normal(alphas.shape) * ok).astype(np.float32)
    N, M = means2d.shape[0], flatten_ids.shape[0]
    v_means2d, v_conics = np.zeros((N, 2), np.float32), np.zeros((N, 3), np.float32)
    v_colors, v_opac = np.zeros((N, CH), np.float32), np.zeros(N, np.float32)
    rc = lib.gsb_composite_bwd(C.c_int32(W), C.c_int32(H), C.c_int32(CH), C.c_int64(N), _p(colors), _p(bg), _p(offsets),
                               C.c_int64(M), _p(alphas), _p(last_ids), _p(v_render), _p(v_alphas), _p(v_means2d),
                               _p(v_conics), _p(v_colors), _p(v_opac), _p(ws), None)
    assert rc == 0, lib.gsb_last_error()
    o = R.composite_bwd(means2d, conics, colors, opac, offsets, flatten_ids, W, H, o_alphas, o_last, v_render, v_alphas,
                        background=bg)
    for a, b, name in zip((v_means2d, v_conics, v_colors, v_opac), o, ("means2d", "conics", "colors", "opacities")):
        assert rel_l2(a, b) <= 2e-4, (name, rel_l2(a, b))


@pytest.mark.parametrize("defines", [("GSB_WPB=4", "GSB_FG=4", "GSB_BG=4", "GSB_WPB_B=1", "GSB_SEG=32"),
                                     ("GSB_WPB=1", "GSB_FG=16", "GSB_BG=16", "GSB_WPB_B=4", "GSB_SEG=16")])
def test_tuning_constants_do_not_change_results(defines):
    """The compile-time knobs scripts/tune_composite.sh sweeps on the GPU (warps per CTA, entries evaluated together in
    the forward groups and in the backward's phase A): every setting is the same function."""
    lib = emu.build("composite", simt=True, defines=defines)
    cam, means2d, conics, colors, opac, flatten_ids, offsets = _inputs(900, (48, 40), seed=31, scale_lo=0.01, scale_hi=0.1)
    W, H = cam.width, cam.height
    render, alphas, last_ids, ws = _composite(lib, W, H, means2d, conics, colors, opac, flatten_ids, offsets)
    o_render, o_alphas, o_last = R.composite_fwd(means2d, conics, colors, opac, offsets, flatten_ids, W, H)
    ok = ~R.composite_fragile(means2d, conics, opac, offsets, flatten_ids, W, H)
    assert np.array_equal(last_ids[ok], o_last[ok]) and np.abs(render - o_render)[ok].max() <= 1e-4
    rng = np.random.default_rng(5)
    v_render = (rng.standard_normal(render.shape) * ok[..., None]).astype(np.float32)
    v_alphas = (rng.standard_normal(alphas.shape) * ok).astype(np.float32)
    N, M = means2d.shape[0], flatten_ids.shape[0]
    g = [np.zeros(s, np.float32) for s in ((N, 2), (N, 3), (N, 3), N)]
    assert lib.gsb_composite_bwd(C.c_int32(W), C.c_int32(H), C.c_int32(3), C.c_int64(N), _p(colors), None, _p(offsets),
                                 C.c_int64(M), _p(alphas), _p(last_ids), _p(v_render), _p(v_alphas), *[_p(x) for x in g],
                                 _p(ws), None) == 0, lib.gsb_last_error()
    o = R.composite_bwd(means2d, conics, colors, opac, offsets, flatten_ids, W, H, o_alphas, o_last, v_render, v_alphas)
    for a, b in zip(g, o):
        assert rel_l2(a, b) <= 2e-4


@pytest.mark.parametrize("CH,seg", [(3, 16), (3, 48), (1, 32), (2, 16)])
def test_segmented_backward_equals_the_oracle(CH, seg):
    """Sub-lists longer than GSB_SEG are checkpointed by the forward and walked by the backward as independent segments
    (persistent warps drawing jobs from a queue; the host build has two 'SMs', so warps loop).  Faint Gaussians on a
    small image: every 4x4 unit's sub-list is several segments long and no pixel saturates early.  The backward is run
    TWICE on one forward (the job cursor is reset per launch)."""
    lib = emu.build("composite", simt=True, defines=("GSB_SEG=%d" % seg,))
    cam, means2d, conics, _, opac, flatten_ids, offsets = _inputs(2500, (24, 20), seed=41, extent=0.4, scale_lo=0.05,
                                                                  scale_hi=0.3)
    opac = (opac * 0.08).astype(np.float32)
    W, H = cam.width, cam.height
    rng = np.random.default_rng(CH + seg)
    colors = rng.random((means2d.shape[0], CH)).astype(np.float32)
    bg = rng.random(CH).astype(np.float32)
    render, alphas, last_ids, ws = _composite(lib, W, H, means2d, conics, colors, opac, flatten_ids, offsets, bg)
    o_render, o_alphas, o_last = R.composite_fwd(means2d, conics, colors, opac, offsets, flatten_ids, W, H, background=bg)
    ok = ~R.composite_fragile(means2d, conics, opac, offsets, flatten_ids, W, H)
    assert ok.mean() > 0.9
    assert np.array_equal(last_ids[ok], o_last[ok]) and np.abs(render - o_render)[ok].max() <= 1e-4
    # the walk really is several segments long: the deepest contributor of a pixel lies far down its tile's list
    assert int((o_last - offsets[0]).max()) > 8 * seg
    v_render = (rng.standard_normal(render.shape) * ok[..., None]).astype(np.float32)
    v_alphas = (rng.standard_normal(alphas.shape) * ok).astype(np.float32)
    N, M = means2d.shape[0], flatten_ids.shape[0]
    o = R.composite_bwd(means2d, conics, colors, opac, offsets, flatten_ids, W, H, o_alphas, o_last, v_render, v_alphas,
                        background=bg)
    for _ in range(2):
        g = [np.zeros(s, np.float32) for s in ((N, 2), (N, 3), (N, CH), N)]
        rc = lib.gsb_composite_bwd(C.c_int32(W), C.c_int32(H), C.c_int32(CH), C.c_int64(N), _p(colors), _p(bg),
                                   _p(offsets), C.c_int64(M), _p(alphas), _p(last_ids), _p(v_render), _p(v_alphas),
                                   *[_p(x) for x in g], _p(ws), None)
        assert rc == 0, lib.gsb_last_error()
        for a, b, name in zip(g, o, ("means2d", "conics", "colors", "opacities")):
            assert rel_l2(a, b) <= 2e-5, (name, rel_l2(a, b))   # 2e-6 measured; 1e-4 if a segment started from the forward's T unscaled
