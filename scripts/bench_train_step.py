"""One WHOLE stage-1 training step of BASELINE config 3 through model.GeoSplatter (everything the reference's trainer does
per step): FlexiCubes mesh + regularisers (R^3 SDF grid) -> vertex normals + MGAdaptor -> kd / ks / z hash-grid fields
-> split-sum prefilter of the 6 x 512^2 cube map -> 8 views 800 x 800 (shade, rasterize, tone map) -> per-view
SSIM / L1 / mask loss -> backward to all eight parameter groups -> Adam.  CUDA events, ms per step.
    python scripts/bench_train_step.py [R=140]          -> gpurun_out/train_step.json"""
import json
import os
import sys

sys.path.insert(0, ".")
import torch

from geosplatting_b200 import _lib, scenes
from geosplatting_b200.model import GeoSplatter
from geosplatting_b200.shade import synthetic_fg_lut

R = int(sys.argv[1]) if len(sys.argv) > 1 else 140
dev = "cuda:0"
torch.manual_seed(0)
m = GeoSplatter(resolution=R, light_resolution=512, scale=0.9, fg_lut=synthetic_fg_lut(torch.device("cpu"))).to(dev)
gv = m.geometric_repr.vertices.to(dev)
with torch.no_grad():
    m.sdf_params.copy_(gv.norm(dim=-1, keepdim=True) - 0.6 + 0.06 * torch.sin(5.0 * gv[:, :1]) * torch.cos(4.0 * gv[:, 1:2]))
    m.cubemap.copy_(torch.exp(torch.randn_like(m.cubemap)).clamp_min(1e-2))
m.train()
m.sdf_weight, m.light_weight = 0.2, 2e-3
m.kd_regualr_perturb_std = m.ks_regualr_perturb_std = 0.01          # the jitter regularisers of the trainer's schedule
m.kd_grad_weight, m.ks_grad_weight = 0.03, 0.001
m.cubemap.register_hook(lambda g: g * 64)
cams = scenes.orbit_cameras(8, 800, 800, seed=1)
gen = torch.Generator().manual_seed(1)
gt = []
for _ in cams:
    img = torch.rand(800, 800, 4, generator=gen)
    img[..., 3] = (img[..., 3] > 0.5).float()
    gt.append(img.to(dev))
opt = torch.optim.Adam([
    {"params": [m.sdf_params, m.deform_params, m.weight_params], "lr": 1e-3},
    {"params": list(m.field.parameters()), "lr": 1e-2},
    {"params": [m.cubemap], "lr": 1e-2}, {"params": [m.exposure_params], "lr": 5e-3}], eps=1e-15, fused=True)
stats = {}


def step():
    opt.zero_grad(set_to_none=True)
    loss, metrics = m.training_loss(cams, gt)
    loss.backward()
    opt.step()
    stats.update(gaussians=int(metrics["#gaussians"]), loss=float(metrics["loss"]))


for _ in range(5):          # the first steps size the speculative capacities and fill the allocator's cache
    step()
torch.cuda.synchronize()
n = 8
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
for a, b in evs:
    a.record()
    step()
    b.record()
torch.cuda.synchronize()
per_step = sorted(a.elapsed_time(b) for a, b in evs)
ms = per_step[n // 2]       # median; min / max reported beside it
_lib.CallStats.reset(timing=True)
step()
per = {k: round(v, 3) for k, (c, v) in _lib.CallStats.durations_ms().items() if v > 0.05}
launches = _lib.CallStats.launches()
_lib.CallStats.reset()
out = {"what": "whole stage-1 training step through model.GeoSplatter.training_loss + Adam, 8 views 800x800", "resolution": R,
       **stats, "ms_per_step": round(ms, 3), "ms_per_step_min_max": [round(per_step[0], 3), round(per_step[-1], 3)],
       "views_per_s": round(8 / (ms / 1e3), 2), "steps_timed": n,
       "entry_point_ms_one_step": per, "gpu_launches_one_step": launches,
       "peak_memory_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/train_step.json", "w"), indent=1)
