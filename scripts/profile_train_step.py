"""Where the whole training step of BASELINE config 3 goes, kernel by kernel (torch.profiler / CUPTI; no nsys in the
image): one profiled step after warm-up -> profiles/rNN_train_step_kernels.json (device time per kernel name, grouped:
this library's kernels, cub / radix sort, torch elementwise + reductions, cuBLAS GEMMs, memcpy / memset) plus the
wall-clock of the step and the device-idle share.
    python scripts/profile_train_step.py [R=140] [out.json]"""
import json
import os
import sys

sys.path.insert(0, ".")
import torch
from torch.profiler import ProfilerActivity, profile

from geosplatting_b200 import scenes
from geosplatting_b200.model import GeoSplatter
from geosplatting_b200.shade import synthetic_fg_lut

R = int(sys.argv[1]) if len(sys.argv) > 1 else 140
OUT = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/train_step_kernels.json"
dev = "cuda:0"
torch.manual_seed(0)
m = GeoSplatter(resolution=R, light_resolution=512, scale=0.9, fg_lut=synthetic_fg_lut(torch.device("cpu"))).to(dev)
gv = m.geometric_repr.vertices.to(dev)
with torch.no_grad():
    m.sdf_params.copy_(gv.norm(dim=-1, keepdim=True) - 0.6 + 0.06 * torch.sin(5.0 * gv[:, :1]) * torch.cos(4.0 * gv[:, 1:2]))
    m.cubemap.copy_(torch.exp(torch.randn_like(m.cubemap)).clamp_min(1e-2))
m.train()
m.sdf_weight, m.light_weight = 0.2, 2e-3
m.kd_regualr_perturb_std = m.ks_regualr_perturb_std = 0.01
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


def step():
    opt.zero_grad(set_to_none=True)
    loss, metrics = m.training_loss(cams, gt)
    loss.backward()
    opt.step()


for _ in range(4):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    step()
b.record()
torch.cuda.synchronize()
ms_step = a.elapsed_time(b) / 5

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()

OURS = ("mlp_fwd", "mlp_bwd", "composite", "sublists", "lpt_order", "pack_records", "shade", "project", "tonemap", "isect", "iota", "tile_offsets",
        "publish_total", "grad_sum", "specular", "diffuse", "cubemap", "dir_table", "prep_source", "hashgrid", "mgadapter",
        "vertex_normals", "fc_", "loss_", "envstack", "texture")
groups = {"library kernels": {}, "cub / radix sort": {}, "torch elementwise / reduce / index": {}, "cuBLAS / GEMM": {},
          "memcpy / memset": {}}
t_min, t_max, busy = None, None, []
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    name, dur = ev.name, ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
    lo = name.lower()
    if "memcpy" in lo or "memset" in lo:
        g = "memcpy / memset"
    elif any(k in lo for k in OURS) and "at::native" not in name:
        g = "library kernels"
    elif "cub" in lo or "radix" in lo or "devicescan" in lo:
        g = "cub / radix sort"
    elif "gemm" in lo or "cutlass" in lo or "sm90" in lo or "sm100" in lo or "gemv" in lo:
        g = "cuBLAS / GEMM"
    else:
        g = "torch elementwise / reduce / index"
    clean = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    short = clean.split("(")[0].split("<")[0][-70:] or clean[:70]
    c, t = groups[g].get(short, (0, 0.0))
    groups[g][short] = (c + 1, t + dur)
    tr = ev.time_range
    busy.append((tr.start, tr.end))
busy.sort()
union, cur_s, cur_e = 0.0, None, None
for s, e in busy:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            union += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
if cur_e is not None:
    union += cur_e - cur_s
span = (busy[-1][1] - busy[0][0]) if busy else 0.0
out = {"what": "one profiled training step (model.GeoSplatter.training_loss + backward + Adam, 8 views 800x800, "
               f"FlexiCubes R={R}); device time per kernel, summed over the streams (so the groups add up to more than "
               "the step when streams overlap)",
       "ms_per_step_unprofiled": round(ms_step, 3), "profiled_span_ms": round(span / 1e3, 3),
       "device_busy_ms (union of kernel intervals)": round(union / 1e3, 3),
       "device_idle_ms_inside_span": round((span - union) / 1e3, 3), "groups": {}}
for g, d in groups.items():
    tot = sum(t for _, t in d.values())
    top = sorted(d.items(), key=lambda kv: -kv[1][1])[:14]
    out["groups"][g] = {"total_ms": round(tot / 1e3, 3), "launches": sum(c for c, _ in d.values()),
                        "top": [{"kernel": k, "launches": c, "ms": round(t / 1e3, 3)} for k, (c, t) in top]}
# which torch operators own the elementwise time (self device time, grouped by input shapes)
ops = []
for ka in prof.key_averages(group_by_input_shape=True):
    t = getattr(ka, "self_device_time_total", None)
    if t is None:
        t = ka.self_cuda_time_total
    if t > 0 and (ka.key.startswith("aten::") or "Optimizer" in ka.key):
        ops.append((t, ka.key, ka.count, str(ka.input_shapes)[:120]))
ops.sort(reverse=True)
out["torch_ops_by_self_device_time"] = [{"op": k, "calls": c, "ms": round(t / 1e3, 3), "shapes": sh} for t, k, c, sh in ops[:30]]
os.makedirs(os.path.dirname(OUT) or ".", exist_ok=True)
json.dump(out, open(OUT, "w"), indent=1)
print(json.dumps({k: (v if not isinstance(v, dict) else {g: x["total_ms"] for g, x in v.items()}) for k, v in out.items()}))
