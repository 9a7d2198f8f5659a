"""Hash-grid field timings at the hot path's size (N = 1 002 528 points, GaussianField configs): CUDA events, L2 flushed
between iterations; algorithmic bytes per point: 12 B in + 128 B of features out (+ 1 KB of L2-resident table gathers)."""
import json, sys
sys.path.insert(0, ".")
import torch
from geosplatting_b200 import encoding as E

dev = torch.device("cuda:0")
N = 1_002_528
x = (torch.rand(N, 3, device=dev) * 2 - 1).requires_grad_(True)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6650.0) if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0
out = {}
for name, mk in (("kd", E.kd_field), ("ks", E.ks_field), ("z", E.z_field)):
    enc = mk().to(dev)
    def timed(fn, n=10):
        fn(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(n):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / n
    feats = enc.encode(x)
    cot = torch.randn_like(feats)
    t_enc = timed(lambda: enc.encode(x))
    t_enc_bwd = timed(lambda: torch.autograd.grad(enc.encode(x), [x, enc.hash_table], grad_outputs=cot))
    y = enc(x)
    coty = torch.randn_like(y)
    t_all = timed(lambda: torch.autograd.grad(enc(x), [x, enc.hash_table] + list(enc.mlp.weights), grad_outputs=coty))
    out[name] = {"encode_fwd_ms": round(t_enc, 4), "encode_fwd_gbs": round(N * 140 / t_enc / 1e6, 1),
                 "encode_fwd_frac_of_measured_hbm": round(N * 140 / t_enc / 1e6 / peak, 3),
                 "encode_fwd_bwd_ms": round(t_enc_bwd, 4), "field_fwd_bwd_ms": round(t_all, 4)}
# the same kd field on mesh-ordered points (the bench scene's MGAdaptor means): neighbouring threads share cells
from geosplatting_b200 import scenes
from geosplatting_b200.mgadapter import MGAdapter, compute_vertex_normals
with torch.no_grad():
    verts, faces = scenes.cube_sphere(118)
    vd, fd = verts.to(dev), faces.to(dev)
    sp, _ = MGAdapter().make(vd, fd, compute_vertex_normals(vd, fd))
xm = (sp.means / 0.9).clamp(-1, 1).detach().requires_grad_(True)
enc = E.kd_field().to(dev)
cotm = torch.randn(xm.shape[0], 32, device=dev)
out["kd_mesh_ordered_points"] = {
    "encode_fwd_ms": round(timed(lambda: enc.encode(xm)), 4),
    "encode_fwd_bwd_ms": round(timed(lambda: torch.autograd.grad(enc.encode(xm), [xm, enc.hash_table], grad_outputs=cotm)), 4)}
print(json.dumps({"N": N, "fields": out}))
