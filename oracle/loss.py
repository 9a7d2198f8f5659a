"""CPU restatement (torch, autograd) of the per-view training loss that consumes the hot path's image
(SURVEY.md section 8f rank 2).  TEST INFRASTRUCTURE ONLY.

Follows rfstudio/trainer/geosplat_trainer.py:171-180 (random-background composite of the rendered and the ground-truth
image, SSIML1Loss, 5 x mask MSE) and rfstudio/loss/photometric_loss.py:72-112 (SSIML1Loss: 0.2 * (1 - SSIM) + 0.8 * L1).
SSIM itself is `torchmetrics.functional.image.structural_similarity_index_measure` with its defaults (Gaussian window,
sigma 1.5, kernel size 11, k1 0.01, k2 0.03, reflect padding, padded border cropped before the mean) and
data_range 1.0 **[3P-recalled: torchmetrics is an unpinned third-party dependency, absent here -> PARITY UNPINNED]**.
Because the border of width 5 is cropped, the reflect padding never reaches the result: SSIM is the mean of the valid
convolution's map over the interior (H-10) x (W-10) pixels.  The variance clamp at 0 is torchmetrics >= 1.0 behaviour.
"""
from __future__ import annotations

import torch
from torch import Tensor

K1, K2, SIGMA, KSIZE = 0.01, 0.03, 1.5, 11


def gaussian_window() -> Tensor:
    dist = torch.arange((1 - KSIZE) / 2, (1 + KSIZE) / 2, 1.0)
    g = torch.exp(-((dist / SIGMA) ** 2) / 2)
    return g / g.sum()


def ssim(x: Tensor, y: Tensor, data_range: float = 1.0) -> Tensor:
    """x, y [H,W,3] -> scalar."""
    c1, c2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    g = gaussian_window().to(x)
    k2d = (g[:, None] * g[None, :])[None, None].expand(3, 1, KSIZE, KSIZE)
    xs, ys = x.permute(2, 0, 1)[None], y.permute(2, 0, 1)[None]
    stack = torch.cat((xs, ys, xs * xs, ys * ys, xs * ys), dim=0)                # [5,3,H,W]
    out = torch.nn.functional.conv2d(stack, k2d, groups=3)                       # valid: [5,3,H-10,W-10]
    mu_x, mu_y = out[0], out[1]
    s_xx = torch.clamp(out[2] - mu_x * mu_x, min=0.0)
    s_yy = torch.clamp(out[3] - mu_y * mu_y, min=0.0)
    s_xy = out[4] - mu_x * mu_y
    m = ((2 * mu_x * mu_y + c1) * (2 * s_xy + c2)) / ((mu_x * mu_x + mu_y * mu_y + c1) * (s_xx + s_yy + c2))
    return m.mean()


def view_loss(rgba: Tensor, gt_rgba: Tensor, bg: Tensor, ssim_lambda: float = 0.2, mask_coeff: float = 5.0):
    """geosplat_trainer.py:171-180.  rgba [H,W,4] rendered (tone-mapped linear rgb + alpha), gt_rgba [H,W,4] ground
    truth in linear rgb + mask, bg [H,W,3] the random background.  -> (loss, ssim_term, l1_term, mask_term)."""
    mask = gt_rgba[..., 3:]
    img1 = rgba[..., :3] + (1 - rgba[..., 3:]) * bg
    img2 = gt_rgba[..., :3] * mask + (1 - mask) * bg
    ssim_loss = 1 - ssim(img2, img1)
    l1 = (img1 - img2).abs().mean()
    mk = (mask - rgba[..., 3:]).square().mean()
    return ssim_loss * ssim_lambda + l1 * (1 - ssim_lambda) + mask_coeff * mk, ssim_loss, l1, mk
