"""Workload statistics of the composite kernels on the bench scene (CPU, oracle restatement of projection and
binning): tile-list lengths and (tile, 8x4 sub-rectangle) sub-list lengths.  Diagnostic only."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from geosplatting_b200 import scenes
from oracle import mgadapter as OM, raster as R

n = int(sys.argv[1]) if len(sys.argv) > 1 else 118
res = int(sys.argv[2]) if len(sys.argv) > 2 else 800
verts, faces = scenes.cube_sphere(n)
means, scales, quats, colors, opac, _ = OM.make(verts, faces, OM.vertex_normals(verts, faces))
cam = scenes.orbit_cameras(8, res, res, seed=1)[0]
ocam = R.Camera(cam.view_matrix, cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height)
out = R.project_fwd(means.numpy(), quats.numpy(), np.exp(scales.numpy()), ocam, antialiased=True)
print({k: (v.shape, v.dtype) for k, v in out.items()} if isinstance(out, dict) else [getattr(o, 'shape', o) for o in out])
radii, m2d, depths, conics, comps = out
op = (1.0 / (1.0 + np.exp(-opac.numpy()[:, 0]))) * comps
tpg, keys, vals, offsets = R.bin_sort(m2d, radii, depths, res, res)
M = len(vals)
off = np.append(offsets.reshape(-1), M)
tl = np.diff(off)
print(f"N={len(radii)} M={M} tiles={len(tl)} nonempty={int((tl>0).sum())} tile list: mean(nonempty)={tl[tl>0].mean():.0f} "
      f"p50={np.percentile(tl[tl>0],50):.0f} p90={np.percentile(tl[tl>0],90):.0f} p99={np.percentile(tl,99):.0f} max={tl.max()}")
# alpha >= 1/255 extents (composite.cu alpha_extent)
t = 255.0 * op
tau2 = 2.0 * np.log(np.maximum(t, 1.0 + 1e-9))
det = conics[:, 0] * conics[:, 2] - conics[:, 1] ** 2
hx = np.sqrt(tau2 * conics[:, 2] / det) * 1.0005 + 0.02
hy = np.sqrt(tau2 * conics[:, 0] / det) * 1.0005 + 0.02
hx[t <= 1] = -1e30
print(f"extent half-widths: mean hx={hx[t>1].mean():.2f} hy={hy[t>1].mean():.2f}  radius mean={radii.mean():.2f}  opac*comp mean={op.mean():.3f}")
tile_of = np.repeat(np.arange(len(tl)), tl)
g = vals
tw = (res + 15) // 16
tx, ty = tile_of % tw, tile_of // tw
for (sw, sh) in ((8, 4), (8, 8), (16, 4), (4, 4), (16, 8), (16, 16)):
    tot = 0
    mx = 0
    lens = []
    for sy in range(16 // sh):
        for sx in range(16 // sw):
            rcx = tx * 16 + sx * sw + 0.5 * sw
            rcy = ty * 16 + sy * sh + 0.5 * sh
            hit = (np.abs(m2d[g, 0] - rcx) <= hx[g] + 0.5 * (sw - 1)) & (np.abs(m2d[g, 1] - rcy) <= hy[g] + 0.5 * (sh - 1))
            c = np.bincount(tile_of[hit], minlength=len(tl))
            lens.append(c)
            tot += hit.sum()
    lens = np.stack(lens, 1).reshape(-1)
    print(f"sub-rect {sw}x{sh}: entries={tot} ({tot/M:.2f} x M) pixel-evals={tot*sw*sh/1e6:.0f} M  sub-list max={lens.max()} "
          f"p99={np.percentile(lens,99):.0f} mean(nonempty)={lens[lens>0].mean():.0f}")

# ---- how much tighter than the AABB test is the exact ellipse-vs-rectangle test (min of the quadratic form over the
#      rectangle of pixel centres <= tau)?  8x4 sub-rectangles.
a_, b_, c_ = conics[:, 0], conics[:, 1], conics[:, 2]
tau = np.log(np.maximum(255.0 * op, 1.0 + 1e-9))
tot_aabb = tot_exact = 0
sw, sh = 8, 4
for sy in range(16 // sh):
    for sx in range(16 // sw):
        rcx = tx * 16 + sx * sw + 0.5 * sw
        rcy = ty * 16 + sy * sh + 0.5 * sh
        hit = (np.abs(m2d[g, 0] - rcx) <= hx[g] + 0.5 * (sw - 1)) & (np.abs(m2d[g, 1] - rcy) <= hy[g] + 0.5 * (sh - 1))
        gi = g[hit]
        lx = (rcx[hit] - 0.5 * (sw - 1)) - m2d[gi, 0]; ux = (rcx[hit] + 0.5 * (sw - 1)) - m2d[gi, 0]
        ly = (rcy[hit] - 0.5 * (sh - 1)) - m2d[gi, 1]; uy = (rcy[hit] + 0.5 * (sh - 1)) - m2d[gi, 1]
        A, B, Cc = a_[gi], b_[gi], c_[gi]
        def q(dx, dy): return 0.5 * (A * dx * dx + Cc * dy * dy) + B * dx * dy
        inside = (lx <= 0) & (ux >= 0) & (ly <= 0) & (uy >= 0)
        cands = []
        for X in (lx, ux):
            dy = np.clip(-B * X / Cc, ly, uy); cands.append(q(X, dy))
        for Y in (ly, uy):
            dx = np.clip(-B * Y / A, lx, ux); cands.append(q(dx, Y))
        qmin = np.where(inside, 0.0, np.minimum.reduce(cands))
        tot_aabb += hit.sum(); tot_exact += (qmin <= tau[gi]).sum()
print(f"8x4: AABB test keeps {tot_aabb} entries, exact ellipse-vs-rectangle test keeps {tot_exact} ({tot_exact / tot_aabb:.3f})")
