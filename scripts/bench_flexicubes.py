"""FlexiCubes get_geometry (mesh extraction + regulariser, forward and backward) on the GPU at training grid sizes:
total ms per call (CUDA events) and ms per C-ABI entry point; the CPU oracle (numpy topology + torch arithmetic, the
reference's op sequence restated) timed beside it at the smallest size.  Writes gpurun_out/fc_bench.json."""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import torch

from geosplatting_b200 import _lib
from geosplatting_b200.flexicubes import FlexiCubes

dev = "cuda:0"
out = {}
for R in (64, 96, 128):
    fc0 = FlexiCubes.from_resolution(R, random_sdf=False, scale=0.9, device=dev)
    gv = fc0.vertices
    sdf = (gv.norm(dim=-1, keepdim=True) - 0.6 + 0.06 * torch.sin(5.0 * gv[:, :1]) * torch.cos(4.0 * gv[:, 1:2]))
    sdf = sdf.clone().requires_grad_(True)
    deform = torch.zeros_like(gv).requires_grad_(True)
    w = (0.1 * torch.randn(fc0.indices.shape[0], 21, device=dev)).requires_grad_(True)
    stats = {}

    def step():
        fc = fc0.replace(vertices=fc0.vertices + deform.tanh() * (0.5 * 0.9 / R), sdf_values=sdf, alpha=w[:, :8],
                         beta=w[:, 8:20], gamma=w[:, 20:])
        mesh, l_dev = fc.dual_marching_cubes()
        reg = l_dev.mean() * 0.5 + w[:, :20].abs().mean() * 0.1 + fc.compute_entropy() * 0.3
        stats.update(faces=mesh.indices.shape[0], vertices=mesh.vertices.shape[0], K=l_dev.shape[0])
        torch.autograd.grad(mesh.vertices.square().sum() + reg, [sdf, deform, w])

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n = 10
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        step()
    b.record()
    torch.cuda.synchronize()
    total = a.elapsed_time(b) / n
    _lib.CallStats.reset(timing=True)
    for _ in range(n):
        step()
    per = {k: round(ms / n, 4) for k, (c, ms) in _lib.CallStats.durations_ms().items()}
    _lib.CallStats.reset()
    F, V = fc0.indices.shape[0], gv.shape[0]
    # gsb_fc_surface / gsb_fc_topology end with a device->host read: their ms include the wait for it
    out[R] = {"cubes": F, "grid_vertices": V, **stats, "ms_per_call_fwd_bwd": round(total, 3), "entry_point_ms": per,
              "entry_point_ms_sum": round(sum(per.values()), 4)}
    print(R, json.dumps(out[R]))

# the CPU oracle beside it (bounded: one grid size)
from oracle import flexicubes as OF
import numpy as np
tb = {k: v.cpu().numpy().astype(np.int64) for k, v in __import__("geosplatting_b200.flexicubes", fromlist=["_tables"])._tables(torch.device("cpu")).items()}
R = 64
fc0 = FlexiCubes.from_resolution(R, random_sdf=False, scale=0.9)
gv = fc0.vertices
sdf = (gv.norm(dim=-1, keepdim=True) - 0.6 + 0.06 * torch.sin(5.0 * gv[:, :1]) * torch.cos(4.0 * gv[:, 1:2])).requires_grad_(True)
w = (0.1 * torch.randn(fc0.indices.shape[0], 21)).requires_grad_(True)
t0 = time.time()
mv, mf, l_dev = OF.dual_marching_cubes(gv, sdf, fc0.indices, (R, R, R), w[:, :8], w[:, 8:20], w[:, 20:], tb)
ent = OF.entropy(sdf, fc0.indices, tb)
torch.autograd.grad(mv.square().sum() + l_dev.mean() + ent, [sdf, w])
out["cpu_oracle_R64_ms"] = round((time.time() - t0) * 1e3, 1)
out["cpu_threads"] = torch.get_num_threads()
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fc_bench.json", "w"), indent=1)
