"""The fused field MLP alone at the hot path's size (1 002 528 points x 32 features): forward and forward + backward of
the three field heads, CUDA events, L2 flushed between iterations.  GSB_LIB_PATH selects a tuning build."""
import json, os, sys
sys.path.insert(0, ".")
import torch
from geosplatting_b200 import encoding as E

dev = torch.device("cuda:0")
N = 1_002_528
x = torch.randn(N, 32, device=dev).requires_grad_(True)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


out = {"lib": os.environ.get("GSB_LIB_PATH", "default")}
for name, layers, act in (("kd", [32, 32, 32, 3], "sigmoid"), ("ks", [32, 32, 2], "none"), ("z", [32, 32, 1], "none")):
    mlp = E.MLP(layers, activation=act).to(dev)
    y = mlp(x)
    cot = torch.randn_like(y)
    with torch.no_grad():
        t_f = timed(lambda: mlp(x))
    t_fb = timed(lambda: torch.autograd.grad(mlp(x), [x] + list(mlp.weights), grad_outputs=cot))
    gfma = N * (len(layers) - 2) * 1024 / 1e9
    out[name] = {"fwd_ms": round(t_f, 4), "fwd_bwd_ms": round(t_fb, 4), "bwd_ms": round(t_fb - t_f, 4),
                 "bwd_tflops_fp32": round(2 * 3 * gfma / ((t_fb - t_f) * 1e-3) / 1e3, 1)}
print(json.dumps(out))
