"""The split-sum prefilter kernels (csrc/prefilter.cu: cone bounds, GGX-weighted gather with direction tables and
warp-cooperative loads, diffuse irradiance through dynamic shared memory, the 2x2 mip chain and its non-transpose
backward) without a GPU: the real kernel source under the SIMT mode of tests/emu, called through the C ABI, against
the C oracle (which tests/test_prefilter_gpu.py pins on the reference's own compiled plugin on the GPU box).
Acceptance as in the GPU suite: bounds within one texel, rgb / wsum 1e-3, wsum 5e-3, gradients 2e-3 of the max."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import prefilter as P
from oracle import shade as S
from tests.emu import build as emu
from tests.test_prefilter_gpu import _close, _close_spec, _cubemap


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


@pytest.fixture(scope="module")
def lib():
    return emu.build("prefilter", simt=True)


@pytest.mark.parametrize("R,rough,tables", [(16, 1.0, True), (32, 0.5, True), (32, 0.29, False), (64, 0.185, True)])
def test_specular_prefilter_on_the_host(lib, R, rough, tables):
    ct = P.ndf_cutoff_costheta(rough)
    c = _cubemap(R, 3 + R).numpy()
    i32, f32 = C.c_int32, C.c_float
    bounds = np.zeros((6, R, R, 24), np.float32)
    assert lib.gsb_specular_bounds(i32(R), f32(ct), _p(bounds), None) == 0
    b_orc = P.specular_bounds(R, ct)
    assert np.abs(bounds - b_orc).max() <= 1
    ws = None
    if tables:                      # the direction-table kernels; a NULL workspace selects the table-free kernel
        nb = C.c_size_t(0)
        assert lib.gsb_specular_workspace_bytes(i32(R), C.byref(nb)) == 0
        ws = np.full(nb.value + 256, 0xFF, np.uint8)
    out = np.zeros((6, R, R, 4), np.float32)
    assert lib.gsb_specular_cubemap_fwd(i32(R), _p(c), _p(bounds), f32(rough), f32(ct), i32(0), _p(out), _p(ws),
                                        None) == 0, lib.gsb_last_error()
    _close_spec(out, P.specular_fwd(c, b_orc, rough, ct), "spec fwd")
    g = torch.randn(6, R, R, 4, generator=torch.Generator().manual_seed(5)).numpy()
    gin = np.zeros((6, R, R, 3), np.float32)
    assert lib.gsb_specular_cubemap_bwd(i32(R), _p(bounds), _p(g), None, f32(rough), f32(ct), _p(gin), _p(ws),
                                        None) == 0, lib.gsb_last_error()
    _close(gin, P.specular_bwd(b_orc, g, rough, ct), 2e-3, "spec bwd")


def test_diffuse_irradiance_and_mip_chain_on_the_host(lib):
    R = 16
    c = _cubemap(R, 9).numpy()
    i32 = C.c_int32
    out = np.zeros_like(c)
    assert lib.gsb_diffuse_cubemap_fwd(i32(R), _p(c), _p(out), i32(3), None) == 0
    _close(out, P.diffuse_fwd(c), 1e-4, "diffuse fwd")
    g = torch.randn(6, R, R, 3, generator=torch.Generator().manual_seed(6)).numpy()
    gin = np.zeros_like(g)
    assert lib.gsb_diffuse_cubemap_bwd(i32(R), _p(g), i32(3), _p(gin), None) == 0
    _close(gin, P.diffuse_bwd(g), 1e-4, "diffuse bwd")
    fine = _cubemap(32, 4)
    down = np.zeros((6, 16, 16, 3), np.float32)
    assert lib.gsb_cubemap_mip_fwd(i32(16), _p(fine.numpy()), i32(3), _p(down), i32(3), None) == 0
    _close(down, S.cubemap_mip_fwd(fine), 1e-6, "mip fwd")
    cot = torch.randn(6, 16, 16, 3, generator=torch.Generator().manual_seed(8))
    gfine = np.zeros((6, 32, 32, 3), np.float32)
    assert lib.gsb_cubemap_mip_bwd(i32(16), _p(cot.numpy()), _p(gfine), None) == 0
    _close(gfine, S.cubemap_mip_bwd(cot), 1e-5, "mip bwd")


def _dirs_fp32(R):
    """texel_dir of csrc/prefilter.cu in numpy float32, operation for operation: [6,R,R,3]."""
    f = np.float32
    idx = np.arange(R, dtype=np.float32)
    g = f(2.0) * ((idx + f(0.5)) / f(R)) - f(1.0)
    gx, gy = np.meshgrid(g, g)                     # [y, x]
    one = np.ones_like(gx)
    faces = [(one, -gy, -gx), (-one, -gy, gx), (gx, one, gy), (gx, -one, -gy), (gx, -gy, one), (-gx, -gy, -one)]
    out = np.empty((6, R, R, 3), np.float32)
    for s, (px, py, pz) in enumerate(faces):
        l = np.sqrt(px * px + py * py + pz * pz)
        out[s] = np.stack([px / l, py / l, pz / l], -1)
    return out


@pytest.mark.parametrize("R,rough", [(16, 1.0), (16, 0.5), (32, 0.29), (32, 0.08), (64, 0.185), (64, 0.08)])
def test_bounds_are_the_exact_box_of_the_fp32_cone_test(lib, R, rough):
    """The geometric bounds search (a quadratic per texel row + exact tests at the ends) against the definition: the
    bounding box, on every face, of the texels that pass the fp32 test `texel_dir . V >= cutoff` the gather kernels
    evaluate -- every entry identical, empty faces included."""
    ct = np.float32(P.ndf_cutoff_costheta(rough))
    bounds = np.zeros((6, R, R, 24), np.float32)
    assert lib.gsb_specular_bounds(C.c_int32(R), C.c_float(ct), _p(bounds), None) == 0
    d = _dirs_fp32(R)
    flat = d.reshape(-1, 3)
    rng = np.random.default_rng(R)
    picks = rng.choice(6 * R * R, size=min(6 * R * R, 400), replace=False)
    for t in picks:
        V = flat[t]
        dots = (d[..., 0] * V[0] + d[..., 1] * V[1]) + d[..., 2] * V[2]      # dot3: left to right
        inside = dots >= ct
        want = np.empty(24, np.float32)
        for s in range(6):
            ys, xs = np.nonzero(inside[s])
            want[4 * s:4 * s + 4] = (xs.min(), xs.max(), ys.min(), ys.max()) if xs.size else (R - 1, 0, R - 1, 0)
        got = bounds.reshape(-1, 24)[t]
        assert np.array_equal(got, want), (t, got, want)


@pytest.mark.parametrize("R,rough", [(16, 1.0), (32, 0.5), (64, 0.185), (64, 0.08)])
def test_cached_plan_reproduces_the_on_the_fly_prefilter(lib, R, rough):
    """gsb_specular_plan_count / _fill build the per-(tap, lane) weight table once; gsb_specular_plan_fwd / _bwd then
    stream it.  Same traversal, same weight expression, same tap order as gsb_specular_cubemap_fwd / _bwd; the two paths
    cut a texel's sum into partial sums differently (the gather per source face, the plan into even shares of the row
    segments, so that the few patches of a coarse level keep every SM busy), so the results agree to the rounding of
    that association (sums of ~2 000 fp32 terms): 5e-6 of the largest entry (forward with and without normalisation, backward with and without the
    forward's weight sums).  With one part per patch (GSB_PLAN_PARTS_TARGET=0) and a gather that is not split either
    (R > 64 on the GPU) they are identical to the bit."""
    ct = P.ndf_cutoff_costheta(rough)
    c = _cubemap(R, 11 + R).numpy()
    i32, f32 = C.c_int32, C.c_float
    bounds = np.zeros((6, R, R, 24), np.float32)
    assert lib.gsb_specular_bounds(i32(R), f32(ct), _p(bounds), None) == 0
    nb = C.c_size_t(0)
    assert lib.gsb_specular_workspace_bytes(i32(R), C.byref(nb)) == 0
    ws = np.full(nb.value + 256, 0xFF, np.uint8)
    n_pf = 6 * R * R // 32 * 6
    counts = np.full((n_pf, 2), -1, np.int32)
    assert lib.gsb_specular_plan_count(i32(R), _p(bounds), f32(ct), _p(counts), _p(ws), None) == 0, lib.gsb_last_error()
    assert counts.min() >= 0 and counts[:, 1].sum() > 6 * R * R // 32
    seg_start = np.concatenate([[0], np.cumsum(counts[:, 0])]).astype(np.int32)
    tap_start = np.concatenate([[0], np.cumsum(counts[:, 1])]).astype(np.int32)
    segs = np.full((int(seg_start[-1]), 4), -1, np.int32)
    weights = np.full(int(tap_start[-1]) * 32, np.nan, np.float32)
    assert lib.gsb_specular_plan_fill(i32(R), _p(bounds), f32(rough), f32(ct), _p(seg_start), _p(tap_start), _p(segs),
                                      _p(weights), _p(ws), None) == 0, lib.gsb_last_error()
    assert np.isfinite(weights).all() and (weights >= 0).all() and segs[:, :3].min() >= 0
    assert int(((segs[:, 1] + 3) // 4 * 4).sum()) == int(tap_start[-1])          # segments padded to whole float4 groups
    for normalize in (0, 1):
        a, b = np.zeros((6, R, R, 4), np.float32), np.zeros((6, R, R, 4), np.float32)
        assert lib.gsb_specular_cubemap_fwd(i32(R), _p(c), _p(bounds), f32(rough), f32(ct), i32(normalize), _p(a), _p(ws),
                                            None) == 0
        assert lib.gsb_specular_plan_fwd(i32(R), _p(c), _p(seg_start), _p(segs), _p(weights), i32(normalize), _p(b),
                                         _p(ws), None) == 0, lib.gsb_last_error()
        assert np.abs(a - b).max() <= 5e-6 * np.abs(a).max(), (normalize, np.abs(a - b).max())
    g = torch.randn(6, R, R, 4, generator=torch.Generator().manual_seed(5)).numpy()
    for fwd_out in (None, b):
        ga, gb = np.zeros((6, R, R, 3), np.float32), np.zeros((6, R, R, 3), np.float32)
        assert lib.gsb_specular_cubemap_bwd(i32(R), _p(bounds), _p(g), _p(fwd_out), f32(rough), f32(ct), _p(ga), _p(ws),
                                            None) == 0
        assert lib.gsb_specular_plan_bwd(i32(R), _p(seg_start), _p(segs), _p(weights), _p(g), _p(fwd_out), _p(gb), _p(ws),
                                         None) == 0, lib.gsb_last_error()
        assert np.abs(ga - gb).max() <= 5e-6 * np.abs(ga).max(), np.abs(ga - gb).max()
