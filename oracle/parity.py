"""One training view through the composed CPU oracle, and the comparison report of a device result against it.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/ and by bench.py's `cpu_baseline` leg (which
times this very view and reports how far the CUDA path's result is from it).  The chain is the reference's per-view
chain, restated stage by stage:
    RenderableAttrs.splat shade block      rfstudio/model/geosplat.py:83-121          oracle/shade.py (torch, autograd)
    GSplatter.render_rgba -> rasterization rfstudio/model/gsplat.py:284-358            oracle/raster_oracle.c (C, OpenMP)
    _tone_mapping_naive                    rfstudio/model/geosplat.py:474-476          oracle/shade.tone_map_naive
and its backward down to the ten gradient groups the trainer consumes: means, log-scales, quats, opacity logits, kd, ks,
normals, env base, env specular levels, exposure.
"""
from __future__ import annotations

import time
from typing import Dict, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from . import raster as R
from . import shade as OS

GROUPS = ("means", "scales", "quats", "logits", "kd", "ks", "normals", "base", "mips", "exposure")


def oracle_view(g: Dict[str, Tensor], base: Tensor, mips: Sequence[Tensor], lut: Tensor, cam, exposure: Tensor,
                cot: Optional[Tensor], *, mode: str = "pbr", rasterize_mode: str = "antialiased",
                min_roughness: float = 0.1, max_metallic: float = 1.0, backward: bool = True) -> dict:
    """`g`: CPU tensors means[N,3], scales[N,3] (LOG scales), quats[N,4], opacities[N,1] (logits), kd, ks, normals.
    `cam`: scenes.PinholeCamera.  `cot`: image cotangent [H,W,4]; its entries at fragile pixels are ignored (zeroed).
    Returns image, fragile mask, binning info, timings and -- with backward -- the gradient of <image, cot> per group."""
    t0 = time.perf_counter()
    leaves = {k: g[k].detach().clone().requires_grad_(True) for k in ("means", "normals", "kd", "ks")}
    o_base = base.detach().clone().requires_grad_(True)
    o_mips = [m.detach().clone().requires_grad_(True) for m in mips]
    col = OS.shade(leaves["means"], leaves["normals"], leaves["kd"], leaves["ks"],
                   torch.from_numpy(np.asarray(cam.position, np.float32).copy()), lut, o_base, o_mips,
                   min_roughness=min_roughness, max_metallic=max_metallic, mode=mode)
    t_shade = time.perf_counter() - t0
    ocam = R.Camera(cam.view_matrix, cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height)
    logits = g["opacities"].detach().reshape(-1)
    sig = torch.sigmoid(logits)
    lin_scales = g["scales"].detach().exp()
    r_in = [np.ascontiguousarray(t.detach().numpy(), np.float32)
            for t in (g["means"], g["quats"], lin_scales, sig, col)]
    t1 = time.perf_counter()
    render, alpha, info = R.rasterization(*r_in, ocam, rasterize_mode=rasterize_mode)
    t_raster = time.perf_counter() - t1
    rgba = torch.tensor(np.concatenate([render, alpha], -1), requires_grad=True)
    ex = exposure.detach().clone().reshape(1).requires_grad_(True)
    img = OS.tone_map_naive(rgba, ex)
    out = {"image": img.detach().numpy(), "fragile": info["fragile"], "info": info, "colors": col.detach().numpy(),
           "seconds": {"shade_fwd": t_shade, "raster_fwd": t_raster}}
    if not backward:
        return out
    cot = cot.detach().clone()
    cot[torch.from_numpy(info["fragile"])] = 0
    out["cot"] = cot
    t2 = time.perf_counter()
    v_rgba, v_ex = torch.autograd.grad((img * cot).sum(), [rgba, ex])
    rg = R.rasterization_bwd(*r_in, ocam, info, alpha, np.ascontiguousarray(v_rgba[..., :3].numpy()),
                             np.ascontiguousarray(v_rgba[..., 3:].numpy()), rasterize_mode=rasterize_mode)
    out["seconds"]["raster_bwd"] = time.perf_counter() - t2
    v_means_r, v_quats, v_scales, v_opac, v_colors = [torch.from_numpy(x) for x in rg]
    t3 = time.perf_counter()
    order = [leaves["means"], leaves["normals"], leaves["kd"], leaves["ks"], o_base, *o_mips]
    sg = torch.autograd.grad(col, order, grad_outputs=v_colors, allow_unused=True)
    out["seconds"]["shade_bwd"] = time.perf_counter() - t3
    z = lambda a, like: torch.zeros_like(like) if a is None else a   # noqa: E731
    out["grads"] = {
        "means": v_means_r + z(sg[0], leaves["means"]), "scales": v_scales * lin_scales, "quats": v_quats,
        "logits": (v_opac * sig * (1 - sig)).reshape(g["opacities"].shape),
        "kd": z(sg[2], leaves["kd"]), "ks": z(sg[3], leaves["ks"]), "normals": z(sg[1], leaves["normals"]),
        "base": z(sg[4], o_base), "mips": [z(a, m) for a, m in zip(sg[5:], o_mips)], "exposure": v_ex}
    return out


def _flat(x) -> np.ndarray:
    if isinstance(x, (list, tuple)):
        return np.concatenate([_flat(t) for t in x]) if len(x) else np.zeros(0)
    if isinstance(x, Tensor):
        x = x.detach().cpu().numpy()
    return np.asarray(x, np.float64).reshape(-1)


def compare(o: dict, image: np.ndarray, flatten_ids: Optional[np.ndarray] = None,
            isect_offsets: Optional[np.ndarray] = None, grads: Optional[dict] = None) -> dict:
    """Report of a device result against `oracle_view`'s: per-pixel L-inf over the non-fragile pixels (and its 99.9 %
    quantile), PSNR over ALL pixels, the fragile fraction, identity of the tile lists, and per gradient group the
    relative L2 error and the L-inf error as a fraction of the group's largest oracle entry."""
    ok = ~o["fragile"]
    d = np.abs(np.asarray(image, np.float64) - o["image"]).max(-1)
    mse = float(np.mean((np.asarray(image, np.float64) - o["image"]) ** 2))
    rep = {"linf": float(d[ok].max()) if ok.any() else 0.0,
           "linf_q999": float(np.quantile(d[ok], 0.999)) if ok.any() else 0.0,
           "linf_fragile": float(d[~ok].max()) if (~ok).any() else 0.0,
           "fragile_pixels": int((~ok).sum()), "fragile_frac": float((~ok).mean()),
           "psnr_db_all_pixels": float("inf") if mse == 0 else float(10.0 * np.log10(1.0 / mse)),
           "intersections": int(o["info"]["flatten_ids"].shape[0])}
    if flatten_ids is not None:
        rep["ids_equal"] = bool(np.array_equal(np.asarray(flatten_ids), o["info"]["flatten_ids"]) and
                                (isect_offsets is None or
                                 np.array_equal(np.asarray(isect_offsets).reshape(-1),
                                                o["info"]["isect_offsets"].reshape(-1))))
    if grads is not None:
        rep["grads"] = {}
        for k in GROUPS:
            if k not in grads or grads[k] is None:
                continue
            a, b = _flat(grads[k]), _flat(o["grads"][k])
            scale = float(np.abs(b).max()) if b.size else 0.0
            rep["grads"][k] = {"rel_l2": float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)),
                               "linf_over_max": float(np.abs(a - b).max() / max(scale, 1e-30)) if b.size else 0.0,
                               "max": scale}
    return rep
