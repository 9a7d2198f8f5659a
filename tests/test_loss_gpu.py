"""The fused per-view loss (gsb_loss_fwd / gsb_loss_bwd) against oracle/loss.py -- a torch
restatement of geosplat_trainer.py:171-180 + SSIML1Loss + torchmetrics' SSIM (third-party, absent here: PARITY
UNPINNED, see the oracle's header).  Loss 1e-5, image cotangent 1e-4 of its max; ragged sizes; SSIM(x, x) = 1."""
import numpy as np
import pytest
import torch

from geosplatting_b200.loss import view_loss
from oracle import loss as OL

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _images(H, W, seed, smooth=True):
    g = torch.Generator().manual_seed(seed)
    def img(c):
        t = torch.rand(1, c, H // 4 + 2, W // 4 + 2, generator=g)
        t = torch.nn.functional.interpolate(t, size=(H, W), mode="bilinear", align_corners=False)[0].permute(1, 2, 0)
        return (t + 0.05 * torch.rand(H, W, c, generator=g)).clamp(0, 1) if smooth else torch.rand(H, W, c, generator=g)
    rgba, gt = img(4), img(4)
    gt[..., 3] = (gt[..., 3] > 0.5).float()                      # binary mask like a dataset's alpha
    rgba[: H // 8] = 0.0                                          # empty rows: exactly zero variance windows
    return rgba.contiguous(), gt.contiguous(), torch.rand(H, W, 3, generator=g)


@pytest.mark.parametrize("H,W,seed", [(800, 800, 0), (123, 77, 1), (16, 40, 2), (11, 11, 3)])
def test_loss_and_cotangent_match_oracle(H, W, seed):
    rgba, gt, bg = _images(H, W, seed)
    o = rgba.clone().requires_grad_(True)
    o_loss, o_ssim, o_l1, o_mask = OL.view_loss(o, gt, bg)
    (o_grad,) = torch.autograd.grad(o_loss * 1.7, [o])
    d = rgba.to(DEV).requires_grad_(True)
    loss, sums = view_loss(d, gt.to(DEV), bg.to(DEV), return_terms=True)
    assert abs(float(loss) - float(o_loss)) <= 1e-5 * max(1.0, abs(float(o_loss)))
    assert abs(float(sums[1]) / (3 * H * W) - float(o_l1)) <= 1e-6
    assert abs(float(sums[2]) / (H * W) - float(o_mask)) <= 1e-6
    assert abs(1 - float(sums[0]) / (3 * (H - 10) * (W - 10)) - float(o_ssim)) <= 1e-5
    (grad,) = torch.autograd.grad(loss * 1.7, [d])
    err = float((grad.cpu() - o_grad).abs().max())
    assert err <= 1e-4 * float(o_grad.abs().max()), (err, float(o_grad.abs().max()))


def test_identity_and_options():
    rgba, gt, bg = _images(96, 64, 5)
    gt = gt.clone()
    gt[..., 3] = 1.0
    same = torch.cat((gt[..., :3], torch.ones(96, 64, 1)), -1).to(DEV).requires_grad_(True)
    loss, sums = view_loss(same, gt.to(DEV), bg.to(DEV), return_terms=True)
    assert abs(float(loss)) <= 1e-6                              # SSIM(x, x) = 1, L1 = 0, mask term = 0
    assert abs(float(sums[0]) / (3 * 86 * 54) - 1.0) <= 1e-6
    l_nomask = view_loss(rgba.to(DEV), gt.to(DEV), bg.to(DEV), use_mask_loss=False)
    l_mask = view_loss(rgba.to(DEV), gt.to(DEV), bg.to(DEV))
    assert float(l_mask) > float(l_nomask)
    assert view_loss(rgba.to(DEV), gt.to(DEV)).dim() == 0        # draws its own background like the reference
    with pytest.raises(RuntimeError, match="no CPU path"):
        view_loss(rgba, gt, bg)
    with pytest.raises(ValueError):
        view_loss(torch.zeros(8, 8, 4, device=DEV), torch.zeros(8, 8, 4, device=DEV), torch.zeros(8, 8, 3, device=DEV))
