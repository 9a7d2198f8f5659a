"""numpy/ctypes front-end of oracle/raster_oracle.c -- the CPU restatement of ``gsplat.rasterization``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: gsplat 1.4.0 is an un-vendored
third-party dependency of the reference (pyproject.toml:24); this follows SURVEY.md Appendix C and is
anchored on the reference's call site rfstudio/model/gsplat.py:334-355.

`rasterization()` mirrors the `gsplat.rendering.rasterization` glue for the argument set the reference
uses: packed=True compaction (ascending Gaussian id), `opacities *= compensations` in antialiased
mode, colour gather by gaussian_ids, RGB render mode, one camera.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

EPS2D = 0.3
TILE = 16


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("raster_oracle.c", "prefilter_oracle.c")]
    stale = (not os.path.exists(so)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in srcs
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_isect_count.restype = C.c_int64
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(n))


@dataclass
class Camera:
    viewmat: np.ndarray  # [4,4] world->camera, OpenCV convention
    fx: float
    fy: float
    cx: float
    cy: float
    width: int
    height: int


def project_fwd(means, quats, scales, cam: Camera, *, near=0.01, far=1e10, eps2d=EPS2D, radius_clip=0.0,
                antialiased=True):
    means, quats, scales = _f32(means), _f32(quats), _f32(scales)
    N = means.shape[0]
    vm = _f32(cam.viewmat).reshape(16)
    radii = np.zeros(N, np.int32)
    means2d = np.zeros((N, 2), np.float32)
    depths = np.zeros(N, np.float32)
    conics = np.zeros((N, 3), np.float32)
    comps = np.zeros(N, np.float32)
    lib().orc_project_fwd(
        C.c_int(N), _p(means), _p(quats), _p(scales), _p(vm), C.c_float(cam.fx), C.c_float(cam.fy),
        C.c_float(cam.cx), C.c_float(cam.cy), C.c_int(cam.width), C.c_int(cam.height), C.c_float(near),
        C.c_float(far), C.c_float(eps2d), C.c_float(radius_clip), C.c_int(int(antialiased)), _p(radii),
        _p(means2d), _p(depths), _p(conics), _p(comps))
    return radii, means2d, depths, conics, comps


def project_bwd(means, quats, scales, cam: Camera, radii, v_means2d, v_depths, v_conics, v_comps, *,
                near=0.01, far=1e10, eps2d=EPS2D, radius_clip=0.0, antialiased=True):
    means, quats, scales = _f32(means), _f32(quats), _f32(scales)
    N = means.shape[0]
    vm = _f32(cam.viewmat).reshape(16)
    v_means = np.zeros((N, 3), np.float32)
    v_quats = np.zeros((N, 4), np.float32)
    v_scales = np.zeros((N, 3), np.float32)
    lib().orc_project_bwd(
        C.c_int(N), _p(means), _p(quats), _p(scales), _p(vm), C.c_float(cam.fx), C.c_float(cam.fy),
        C.c_float(cam.cx), C.c_float(cam.cy), C.c_int(cam.width), C.c_int(cam.height), C.c_float(near),
        C.c_float(far), C.c_float(eps2d), C.c_float(radius_clip), C.c_int(int(antialiased)),
        _p(np.ascontiguousarray(radii, np.int32)), _p(_f32(v_means2d)),
        _p(None if v_depths is None else _f32(v_depths)), _p(_f32(v_conics)), _p(_f32(v_comps)),
        _p(v_means), _p(v_quats), _p(v_scales))
    return v_means, v_quats, v_scales


def tile_grid(width: int, height: int, tile: int = TILE):
    return (width + tile - 1) // tile, (height + tile - 1) // tile


def bin_sort(means2d, radii, depths, width, height, tile=TILE, camera_id=0):
    """Packed inputs -> (tiles_per_gauss, isect_ids[M] i64, flatten_ids[M] i32, isect_offsets[th,tw] i32)."""
    means2d = _f32(means2d)
    radii = np.ascontiguousarray(radii, np.int32)
    depths = _f32(depths)
    nnz = radii.shape[0]
    tw, th = tile_grid(width, height, tile)
    tpg = np.zeros(nnz, np.int32)
    M = int(lib().orc_isect_count(C.c_int(nnz), _p(means2d), _p(radii), C.c_int(tile), C.c_int(tw),
                                  C.c_int(th), _p(tpg)))
    keys = np.zeros(M, np.int64)
    vals = np.zeros(M, np.int32)
    lib().orc_isect_tiles(C.c_int(nnz), _p(means2d), _p(radii), _p(depths), C.c_int(tile), C.c_int(tw),
                          C.c_int(th), C.c_int(camera_id), _p(keys), _p(vals))
    n_tiles = tw * th
    tile_n_bits = int(np.floor(np.log2(n_tiles))) + 1
    lib().orc_radix_sort_pairs(C.c_int64(M), C.c_int(32 + tile_n_bits), _p(keys), _p(vals))
    offsets = np.zeros(n_tiles, np.int32)
    lib().orc_isect_offsets(C.c_int64(M), _p(keys), C.c_int(1), C.c_int(tw), C.c_int(th), _p(offsets))
    return tpg, keys, vals, offsets.reshape(th, tw)


def composite_fwd(means2d, conics, colors, opacities, offsets, flatten_ids, width, height, tile=TILE,
                  background=None):
    means2d, conics, colors, opacities = _f32(means2d), _f32(conics), _f32(colors), _f32(opacities)
    CH = colors.shape[1]
    assert CH <= 64
    tw, th = tile_grid(width, height, tile)
    offsets = np.ascontiguousarray(offsets, np.int32).reshape(-1)
    flatten_ids = np.ascontiguousarray(flatten_ids, np.int32)
    M = flatten_ids.shape[0]
    render = np.zeros((height, width, CH), np.float32)
    alphas = np.zeros((height, width), np.float32)
    last_ids = np.zeros((height, width), np.int32)
    bg = None if background is None else _f32(background)
    lib().orc_composite_fwd(C.c_int(width), C.c_int(height), C.c_int(tile), C.c_int(tw), C.c_int(th),
                            C.c_int(CH), _p(means2d), _p(conics), _p(colors), _p(opacities), _p(bg),
                            _p(offsets), _p(flatten_ids), C.c_int64(M), _p(render), _p(alphas), _p(last_ids))
    return render, alphas, last_ids


def composite_fragile(means2d, conics, opacities, offsets, flatten_ids, width, height, tile=TILE, eps=2e-5):
    """uint8[H,W] mask of pixels within `eps` (relative) of flipping a discrete decision (see the C source)."""
    means2d, conics, opacities = _f32(means2d), _f32(conics), _f32(opacities)
    tw, th = tile_grid(width, height, tile)
    offsets = np.ascontiguousarray(offsets, np.int32).reshape(-1)
    flatten_ids = np.ascontiguousarray(flatten_ids, np.int32)
    out = np.zeros((height, width), np.uint8)
    lib().orc_composite_fragile(C.c_int(width), C.c_int(height), C.c_int(tile), C.c_int(tw), C.c_int(th),
                                _p(means2d), _p(conics), _p(opacities), _p(offsets), _p(flatten_ids),
                                C.c_int64(flatten_ids.shape[0]), C.c_float(eps), _p(out))
    return out.astype(bool)


def composite_bwd(means2d, conics, colors, opacities, offsets, flatten_ids, width, height, alphas, last_ids,
                  v_render, v_alphas, tile=TILE, background=None):
    means2d, conics, colors, opacities = _f32(means2d), _f32(conics), _f32(colors), _f32(opacities)
    CH = colors.shape[1]
    nnz = means2d.shape[0]
    tw, th = tile_grid(width, height, tile)
    offsets = np.ascontiguousarray(offsets, np.int32).reshape(-1)
    flatten_ids = np.ascontiguousarray(flatten_ids, np.int32)
    M = flatten_ids.shape[0]
    v_means2d = np.zeros((nnz, 2), np.float32)
    v_conics = np.zeros((nnz, 3), np.float32)
    v_colors = np.zeros((nnz, CH), np.float32)
    v_opac = np.zeros(nnz, np.float32)
    bg = None if background is None else _f32(background)
    lib().orc_composite_bwd(C.c_int(width), C.c_int(height), C.c_int(tile), C.c_int(tw), C.c_int(th),
                            C.c_int(CH), _p(means2d), _p(conics), _p(colors), _p(opacities), _p(bg),
                            _p(offsets), _p(flatten_ids), C.c_int64(M), _p(_f32(alphas)),
                            _p(np.ascontiguousarray(last_ids, np.int32)), _p(_f32(v_render)),
                            _p(_f32(v_alphas)), C.c_int(nnz), _p(v_means2d), _p(v_conics), _p(v_colors),
                            _p(v_opac))
    return v_means2d, v_conics, v_colors, v_opac


def rasterization(means, quats, scales, opacities, colors, cam: Camera, *, tile_size=TILE, near_plane=0.01,
                  far_plane=1e10, eps2d=EPS2D, radius_clip=0.0, rasterize_mode="antialiased", background=None):
    """Forward of gsplat.rasterization(packed=True, render_mode='RGB') for one camera.

    Returns (render[H,W,D], alpha[H,W,1], info) with the packed `info` fields gsplat exposes.
    """
    aa = rasterize_mode == "antialiased"
    assert rasterize_mode in ("classic", "antialiased")
    radii, means2d, depths, conics, comps = project_fwd(
        means, quats, scales, cam, near=near_plane, far=far_plane, eps2d=eps2d, radius_clip=radius_clip,
        antialiased=aa)
    gids = np.nonzero(radii > 0)[0].astype(np.int64)
    p_radii, p_m2d, p_depths, p_conics = radii[gids], means2d[gids], depths[gids], conics[gids]
    p_comps = comps[gids]
    p_opac = _f32(opacities)[gids]
    if aa:
        p_opac = p_opac * p_comps
    p_colors = _f32(colors)[gids]
    tpg, isect_ids, flatten_ids, offsets = bin_sort(p_m2d, p_radii, p_depths, cam.width, cam.height, tile_size)
    render, alphas, last_ids = composite_fwd(p_m2d, p_conics, p_colors, p_opac, offsets, flatten_ids,
                                             cam.width, cam.height, tile_size, background)
    fragile = composite_fragile(p_m2d, p_conics, p_opac, offsets, flatten_ids, cam.width, cam.height, tile_size)
    info = dict(fragile=fragile, camera_ids=np.zeros_like(gids), gaussian_ids=gids, radii=p_radii, means2d=p_m2d,
                depths=p_depths, conics=p_conics, opacities=p_opac,
                compensations=p_comps if aa else None, tiles_per_gauss=tpg, isect_ids=isect_ids,
                flatten_ids=flatten_ids, isect_offsets=offsets[None], last_ids=last_ids,
                width=cam.width, height=cam.height, tile_size=tile_size, n_cameras=1,
                radii_unpacked=radii)
    return render, alphas[..., None], info


def rasterization_bwd(means, quats, scales, opacities, colors, cam: Camera, info, alpha, v_render, v_alpha, *,
                      near_plane=0.01, far_plane=1e10, eps2d=EPS2D, radius_clip=0.0,
                      rasterize_mode="antialiased", background=None):
    """VJP of `rasterization` w.r.t. (means, quats, scales, opacities, colors)."""
    aa = rasterize_mode == "antialiased"
    gids = info["gaussian_ids"]
    N = np.asarray(means).shape[0]
    v_m2d, v_conics, v_colors_p, v_opac_p = composite_bwd(
        info["means2d"], info["conics"], _f32(colors)[gids], info["opacities"], info["isect_offsets"][0],
        info["flatten_ids"], cam.width, cam.height, np.asarray(alpha)[..., 0], info["last_ids"], v_render,
        np.asarray(v_alpha).reshape(cam.height, cam.width), info["tile_size"], background)
    D = v_colors_p.shape[1]
    v_colors = np.zeros((N, D), np.float32)
    v_colors[gids] = v_colors_p
    v_opac = np.zeros(N, np.float32)
    u_m2d = np.zeros((N, 2), np.float32); u_m2d[gids] = v_m2d
    u_con = np.zeros((N, 3), np.float32); u_con[gids] = v_conics
    u_comp = np.zeros(N, np.float32)
    if aa:
        v_opac[gids] = v_opac_p * info["compensations"]
        u_comp[gids] = v_opac_p * _f32(opacities)[gids]
    else:
        v_opac[gids] = v_opac_p
    v_means, v_quats, v_scales = project_bwd(
        means, quats, scales, cam, info["radii_unpacked"], u_m2d, None, u_con, u_comp, near=near_plane,
        far=far_plane, eps2d=eps2d, radius_clip=radius_clip, antialiased=aa)
    return v_means, v_quats, v_scales, v_opac, v_colors
