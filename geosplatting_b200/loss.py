"""Per-view training loss on the rendered image, over the C ABI (gsb_loss_fwd / gsb_loss_bwd).

Mirrors the per-view body of GeoSplatTrainer.step (rfstudio/trainer/geosplat_trainer.py:171-180):

    train_bg_color = torch.rand_like(pbra_item[..., :3])
    img1 = pbra_item[..., :3] + (1 - pbra_item[..., 3:]) * train_bg_color
    img2 = gt_pbra_item[..., :3] * mask + (1 - mask) * train_bg_color
    loss = SSIML1Loss()._impl(img1, img2) [+ 5 * (mask - pbra_item[..., 3:]).square().mean()]

with SSIML1Loss = 0.2 * (1 - SSIM) + 0.8 * L1 (rfstudio/loss/photometric_loss.py:100-112).  Two kernels per view instead
of ~40 torch kernels; the backward writes the [H,W,4] image cotangent the splat backward consumes.  No CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from ._lib import call, f32c, ptr, stream_ptr
from ._lib import require_cuda as _require_cuda


class _ViewLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgba: Tensor, gt_rgba: Tensor, bg: Tensor, ssim_lambda: float, mask_coeff: float):
        r, g, b = f32c(rgba), f32c(gt_rgba), f32c(bg)
        dev = r.device
        H, W = r.shape[0], r.shape[1]
        if H <= 10 or W <= 10:
            raise ValueError("view_loss: the SSIM window needs images larger than 10 x 10")
        sums = torch.empty(4, dtype=torch.float32, device=dev)
        maps = torch.empty(9, H, W, dtype=torch.float32, device=dev)
        call("gsb_loss_fwd", dev, C.c_int32(H), C.c_int32(W), ptr(r), ptr(g), ptr(b), C.c_float(ssim_lambda),
             C.c_float(mask_coeff), ptr(sums), ptr(maps), stream_ptr(dev))
        ctx.save_for_backward(r, g, b, maps)
        ctx.misc = (H, W, float(ssim_lambda), float(mask_coeff))
        terms = sums[:3]
        ctx.mark_non_differentiable(terms)
        return sums[3], terms

    @staticmethod
    def backward(ctx, v_loss, _v_sums):
        r, g, b, maps = ctx.saved_tensors
        H, W, lam, mc = ctx.misc
        dev = r.device
        v_rgba = torch.empty_like(r)
        vl = f32c(v_loss).reshape(1)
        call("gsb_loss_bwd", dev, C.c_int32(H), C.c_int32(W), ptr(r), ptr(g), ptr(b), ptr(maps), C.c_float(lam),
             C.c_float(mc), ptr(vl), ptr(v_rgba), stream_ptr(dev))
        return v_rgba, None, None, None, None


def view_loss(rgba: Tensor, gt_rgba: Tensor, train_bg_color: Optional[Tensor] = None, *, ssim_lambda: float = 0.2,
              use_mask_loss: bool = True, mask_coeff: float = 5.0, return_terms: bool = False):
    """geosplat_trainer.py:171-180 for one view.  rgba [H,W,4]: the rendered image (RenderableAttrs.splat output);
    gt_rgba [H,W,4]: ground truth in LINEAR rgb + mask (`gt_rgba.srgb2rgb()`); train_bg_color [H,W,3] (default:
    torch.rand_like, as the reference draws it).  Differentiable w.r.t. `rgba`.  -> scalar loss
    (with return_terms: also the three raw sums: SSIM map, |img1 - img2|, (mask - alpha)^2)."""
    _require_cuda(rgba, "view_loss")
    assert rgba.shape[-1] == 4 and gt_rgba.shape == rgba.shape and rgba.dim() == 3
    if train_bg_color is None:
        train_bg_color = torch.rand_like(rgba[..., :3])
    assert train_bg_color.shape == rgba.shape[:2] + (3,)
    loss, sums = _ViewLoss.apply(rgba, gt_rgba, train_bg_color, float(ssim_lambda),
                                 float(mask_coeff) if use_mask_loss else 0.0)
    return (loss, sums) if return_terms else loss
