"""GPU-box diagnostic: this library's specular prefilter vs the reference plugin (oracle/_ref)."""
import sys
sys.path.insert(0, ".")
import torch
from geosplatting_b200 import splitsum as SS
from oracle import prefilter as P

plugin = P.load_reference_plugin()
DEV = "cuda:0"
for R, rough in [(64, 0.29), (128, 0.29), (256, 0.185), (512, 0.08)]:
    ct = SS.ndf_cutoff_costheta(rough, 0.99)
    g = torch.Generator().manual_seed(11)
    c = torch.exp(torch.randn(6, R, R, 3, generator=g)).clamp_min(1e-2).to(DEV)
    b_ref = plugin.specular_bounds(R, ct, 0)
    b = SS.render_utils.specular_bounds(R, ct, 0)
    a = SS.render_utils.specular_cubemap_fwd(c, b_ref, rough, ct)
    r = plugin.specular_cubemap_fwd(c, b_ref, rough, ct)
    ra, rr = a[..., :3] / a[..., 3:], r[..., :3] / r[..., 3:]
    gg = torch.randn(6, R, R, 4, generator=g).to(DEV)
    ga = SS.render_utils.specular_cubemap_bwd(c, b, gg, rough, ct)
    gr = plugin.specular_cubemap_bwd(c, b_ref, gg, rough, ct)
    print(R, rough, "bounds diff", (b - b_ref).abs().max().item(),
          "rgb relerr %.2e" % ((ra - rr).abs().max() / rr.abs().max()).item(),
          "wsum relerr %.2e" % ((a[..., 3] - r[..., 3]).abs() / r[..., 3]).max().item(),
          "bwd relerr %.2e" % ((ga - gr).abs().max() / gr.abs().max()).item())
