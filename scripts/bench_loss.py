"""Fused per-view loss vs the same loss in torch ops (oracle code on the GPU), 800x800, CUDA events."""
import sys; sys.path.insert(0, ".")
import torch
from geosplatting_b200.loss import view_loss
from oracle import loss as OL      # diagnostic script: the torch restatement doubles as the "what the reference runs" arm
dev = "cuda:0"
H = W = 800
rgba = torch.rand(H, W, 4, device=dev, requires_grad=True); gt = torch.rand(H, W, 4, device=dev); bg = torch.rand(H, W, 3, device=dev)
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def fused():
    torch.autograd.grad(view_loss(rgba, gt, bg), [rgba])
def torch_ops():
    torch.autograd.grad(OL.view_loss(rgba, gt, bg)[0], [rgba])
print(f"fused loss fwd+bwd {timed(fused):.3f} ms   torch ops {timed(torch_ops):.3f} ms")
