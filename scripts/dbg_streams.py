import sys, time; sys.path.insert(0, ".")
sys.argv = [sys.argv[0]]
import torch
import bench
from geosplatting_b200 import scenes, splitsum
from geosplatting_b200.fused import splat_view, splat_views
from geosplatting_b200.mgadapter import MGAdapter, compute_vertex_normals
from geosplatting_b200.shade import EnvStack, synthetic_fg_lut
from tests.test_golden_cpu import synthetic_fg_lut as lut_np
DEV = "cuda:0"
names = ("means", "scales", "quats", "opacities", "kd", "ks", "normals")
# ---- timing on the bench scene
a = bench.parse_args()
sc = bench.build_scene_host(a)
dev = torch.device(DEV)
cams = scenes.orbit_cameras(8, a.res, a.res, seed=1)
lut = synthetic_fg_lut(dev)
with torch.no_grad():
    vd, fd = sc["verts"].to(dev), sc["faces"].to(dev)
    sp, _ = MGAdapter().make(vd, fd, compute_vertex_normals(vd, fd))
    env0 = splitsum.as_envstack(sc["cubemap"].to(dev))
p = {"means": sp.means, "scales": sp.scales, "quats": sp.quats, "opacities": sp.opacities, "kd": sc["kd"].to(dev),
     "ks": sc["ks"].to(dev), "normals": sp.colors}
p = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
env_data = env0.data.detach().clone().requires_grad_(True)
env = EnvStack(env_data, env0.R0, env0.L, env0.Rb, env0.min_roughness, env0.max_roughness)
exposure = torch.ones(1, device=dev, requires_grad=True)
v_img = torch.randn(a.res, a.res, 4, device=dev)
gi = [p[k] for k in bench.PARAM_NAMES] + [env_data, exposure]
args = [p[k] for k in names]
kw = dict(envmap=env, fg_lut=lut, min_roughness=0.1, max_metallic=1.0)
def batch(ns, fwd_only=False):
    imgs = splat_views(*args, cams, exposures=exposure, n_streams=ns, **kw)
    if not fwd_only:
        torch.autograd.grad(imgs, gi, grad_outputs=[v_img] * 8)
def seq():
    for c in cams:
        img = splat_view(*args, c, exposure=exposure, **kw)
        torch.autograd.grad(img, gi, grad_outputs=v_img)
def timeit(f, n=4):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n / 8, (time.perf_counter() - t0) / n / 8 * 1e3
print("per view ms (device, wall): sequential", timeit(seq))
for ns in (1, 2, 3):
    print(f"  batch n_streams={ns}", timeit(lambda: batch(ns)), " fwd only", timeit(lambda: batch(ns, True)))
print("mem GB", torch.cuda.max_memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def batch_flush():
    flush.zero_(); batch(3)
print("  batch n_streams=3 + flush", timeit(batch_flush))
