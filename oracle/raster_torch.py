"""Differentiable dense torch restatement of the rasterizer (any dtype) -- used ONLY to check the
hand-derived backward of oracle/raster_oracle.c (and hence of the CUDA path) with autograd on tiny cases.

TEST INFRASTRUCTURE ONLY.  Same semantics as oracle/raster.py, written as
O(P*N) tensor algebra: every pixel evaluates every Gaussian, masks restate the tile binning, the
alpha threshold and the transmittance stop.
"""
from __future__ import annotations

import torch

ALPHA_CLAMP = 0.999
ALPHA_MIN = 1.0 / 255.0
T_STOP = 1e-4


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y),
    ], dim=-1).reshape(q.shape[:-1] + (3, 3))


def project(means, quats, scales, viewmat, fx, fy, cx, cy, W, H, *, near=0.01, far=1e10, eps2d=0.3,
            antialiased=True):
    R = quat_to_rotmat(quats)
    M = R * scales[:, None, :]
    S = M @ M.transpose(-1, -2)
    Rcw, t = viewmat[:3, :3], viewmat[:3, 3]
    pc = means @ Rcw.T + t
    Sc = Rcw @ S @ Rcw.T
    x, y, z = pc.unbind(-1)
    tanx, tany = 0.5 * W / fx, 0.5 * H / fy
    lxp, lxn = (W - cx) / fx + 0.3 * tanx, cx / fx + 0.3 * tanx
    lyp, lyn = (H - cy) / fy + 0.3 * tany, cy / fy + 0.3 * tany
    rz = 1.0 / z
    tx = z * torch.clamp(x * rz, min=-lxn, max=lxp)
    ty = z * torch.clamp(y * rz, min=-lyn, max=lyp)
    zero = torch.zeros_like(z)
    J = torch.stack([fx * rz, zero, -fx * tx * rz * rz, zero, fy * rz, -fy * ty * rz * rz], -1).reshape(-1, 2, 3)
    C2 = J @ Sc @ J.transpose(-1, -2)
    mean2d = torch.stack([fx * x * rz + cx, fy * y * rz + cy], -1)
    c00, c01, c11 = C2[:, 0, 0], C2[:, 0, 1], C2[:, 1, 1]
    det0 = c00 * c11 - c01 * c01
    c00 = c00 + eps2d
    c11 = c11 + eps2d
    det = c00 * c11 - c01 * c01
    comp = torch.sqrt(torch.clamp(det0 / det, min=0.0)) if antialiased else torch.ones_like(det)
    conic = torch.stack([c11 / det, -c01 / det, c00 / det], -1)
    b = 0.5 * (c00 + c11)
    radius = torch.ceil(3.0 * torch.sqrt(b + torch.sqrt(torch.clamp(b * b - det, min=0.01))))
    valid = (z >= near) & (z <= far) & (det > 0) & (radius > 0)
    valid &= (mean2d[:, 0] + radius > 0) & (mean2d[:, 0] - radius < W)
    valid &= (mean2d[:, 1] + radius > 0) & (mean2d[:, 1] - radius < H)
    return mean2d, z, conic, comp, radius.detach(), valid


def rasterize_dense(means, quats, scales, opacities, colors, viewmat, fx, fy, cx, cy, W, H, *, tile=16,
                    antialiased=True, eps2d=0.3, order=None):
    """Returns (render[H,W,D], alpha[H,W]).  `order` (optional, LongTensor) fixes the depth order so
    that ties are resolved like the stable radix sort of the fp32 path."""
    mean2d, depth, conic, comp, radius, valid = project(
        means, quats, scales, viewmat, fx, fy, cx, cy, W, H, eps2d=eps2d, antialiased=antialiased)
    opac = opacities * comp if antialiased else opacities
    if order is None:
        order = torch.argsort(depth.detach().float(), stable=True)
    m2, cn, op, col, rad, val = mean2d[order], conic[order], opac[order], colors[order], radius[order], valid[order]
    tw, th = (W + tile - 1) // tile, (H + tile - 1) // tile
    # tile rectangle of every Gaussian (float32 arithmetic of the bin step is restated in fp32)
    m2f, rf = m2.detach().float(), rad.float()
    tx0 = torch.clamp(torch.floor((m2f[:, 0] - rf) / tile), 0, tw)
    tx1 = torch.clamp(torch.ceil((m2f[:, 0] + rf) / tile), 0, tw)
    ty0 = torch.clamp(torch.floor((m2f[:, 1] - rf) / tile), 0, th)
    ty1 = torch.clamp(torch.ceil((m2f[:, 1] + rf) / tile), 0, th)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    px = (xs.reshape(-1).to(means.dtype) + 0.5)
    py = (ys.reshape(-1).to(means.dtype) + 0.5)
    ptx = (xs.reshape(-1) // tile).float()
    pty = (ys.reshape(-1) // tile).float()
    in_tile = (ptx[:, None] >= tx0[None]) & (ptx[:, None] < tx1[None]) & (pty[:, None] >= ty0[None]) & (pty[:, None] < ty1[None])
    dx = m2[None, :, 0] - px[:, None]
    dy = m2[None, :, 1] - py[:, None]
    sigma = 0.5 * (cn[None, :, 0] * dx * dx + cn[None, :, 2] * dy * dy) + cn[None, :, 1] * dx * dy
    vis = torch.exp(-sigma)
    alpha = torch.clamp(op[None] * vis, max=ALPHA_CLAMP)
    ok = in_tile & val[None] & (sigma >= 0) & (alpha >= ALPHA_MIN)
    alpha = torch.where(ok, alpha, torch.zeros_like(alpha))
    Tn = torch.cumprod(1 - alpha, dim=1)  # transmittance AFTER each Gaussian
    stopped = (Tn <= T_STOP) & ok
    live = torch.cumsum(stopped.to(torch.int32), dim=1) == 0  # Gaussian that trips the stop is excluded
    alpha = torch.where(live, alpha, torch.zeros_like(alpha))
    Tn = torch.cumprod(1 - alpha, dim=1)
    Tb = torch.cat([torch.ones_like(Tn[:, :1]), Tn[:, :-1]], dim=1)
    w = alpha * Tb
    render = w @ col
    a_out = 1 - Tn[:, -1]
    return render.reshape(H, W, -1), a_out.reshape(H, W)
