"""GPU-box micro-benchmark: per-level time of the specular prefilter, forward and backward, on-the-fly kernels
(gsb_specular_cubemap_*) against the cached plan (gsb_specular_plan_*), with the plan's size and the HBM rate its
streaming reaches; then the whole as_envstack forward + backward."""
import ctypes as C
import json
import sys

sys.path.insert(0, ".")
import torch

from geosplatting_b200 import splitsum as SS
from geosplatting_b200._lib import call, ptr, stream_ptr

dev = torch.device("cuda:0")
levels = [(512, 0.08), (256, 0.185), (128, 0.29), (64, 0.395), (32, 0.5), (16, 1.0)]
out = {"levels": []}


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for R, rough in levels:
    c = torch.rand(6, R, R, 3, device=dev) + 0.1
    ct, b = SS.ndf_bounds(R, rough, 0.99, 0)
    g = torch.rand(6, R, R, 4, device=dev)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    plan = SS.specular_plan(R, rough, 0.99, dev)
    torch.cuda.synchronize()
    build_ms = (time.perf_counter() - t0) * 1e3
    ws = SS._spec_workspace(R, dev)
    o4, g3 = torch.empty(6, R, R, 4, device=dev), torch.empty(6, R, R, 3, device=dev)
    st = stream_ptr(dev)
    row = {"R": R, "roughness": rough,
           "fly_fwd_ms": round(timed(lambda: SS.render_utils.specular_cubemap_fwd(c, b, rough, ct)), 4),
           "fly_bwd_ms": round(timed(lambda: SS.render_utils.specular_cubemap_bwd(c, b, g, rough, ct)), 4)}
    if plan is not None:
        f = lambda: call("gsb_specular_plan_fwd", dev, C.c_int32(R), ptr(c), ptr(plan.seg_start), ptr(plan.segs),   # noqa: E731
                         ptr(plan.weights), C.c_int32(0), ptr(o4), ptr(ws), st)
        bw = lambda: call("gsb_specular_plan_bwd", dev, C.c_int32(R), ptr(plan.seg_start), ptr(plan.segs),        # noqa: E731
                          ptr(plan.weights), ptr(g), None, ptr(g3), ptr(ws), st)
        fm, bm = timed(f), timed(bw)
        row.update(plan_fwd_ms=round(fm, 4), plan_bwd_ms=round(bm, 4), plan_mb=round(plan.nbytes / 2 ** 20, 1),
                   plan_build_ms=round(build_ms, 2), plan_stream_gbs=round(plan.nbytes / (fm * 1e-3) / 1e9, 1))
        ref_ = SS.render_utils.specular_cubemap_fwd(c, b, rough, ct)
        assert float((o4 - ref_).abs().max()) <= 1e-5 * float(ref_.abs().max())   # small levels add through atomics
    out["levels"].append(row)
    print(row)
cube = (torch.rand(6, 512, 512, 3, device=dev) + 0.1).requires_grad_(True)
cot = torch.rand(SS.as_envstack(cube).data.shape, device=dev)
out["as_envstack_fwd_bwd_ms"] = round(timed(lambda: torch.autograd.grad(SS.as_envstack(cube).data, cube, grad_outputs=cot)), 3)
out["plan_total_gb"] = round(sum(r.get("plan_mb", 0) for r in out["levels"]) / 1024, 2)
print(json.dumps(out))
