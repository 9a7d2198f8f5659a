"""GPU-box micro-benchmark: per-level time of the specular prefilter (forward and backward)."""
import sys
sys.path.insert(0, ".")
import torch
from geosplatting_b200 import splitsum as SS

dev = torch.device("cuda:0")
levels = [(512, 0.08), (256, 0.185), (128, 0.29), (64, 0.395), (32, 0.5), (16, 1.0)]
for R, rough in levels:
    c = torch.rand(6, R, R, 3, device=dev) + 0.1
    ct, b = SS.ndf_bounds(R, rough, 0.99, 0)
    g = torch.rand(6, R, R, 4, device=dev)
    for name, fn in (("fwd", lambda: SS.render_utils.specular_cubemap_fwd(c, b, rough, ct)),
                     ("bwd", lambda: SS.render_utils.specular_cubemap_bwd(c, b, g, rough, ct))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            fn()
        e.record()
        torch.cuda.synchronize()
        bb = b.view(-1, 6, 4)
        taps = ((bb[..., 1] - bb[..., 0] + 1).clamp_min(0) * (bb[..., 3] - bb[..., 2] + 1).clamp_min(0)).sum().item()
        print(f"R={R:4d} rough={rough:.3f} {name}: {s.elapsed_time(e) / 5:.3f} ms   AABB taps {taps:.3e}")
