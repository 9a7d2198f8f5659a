"""Where does the host spend its time per view?  Diagnostic: CPU enqueue time of each phase vs device time."""
import sys, time
sys.path.insert(0, ".")
sys.argv = [sys.argv[0]]
import torch
import bench
from geosplatting_b200 import _lib, scenes, splitsum
Rz = sys.modules['geosplatting_b200.rasterization']
from geosplatting_b200.mgadapter import MGAdapter, compute_vertex_normals
from geosplatting_b200.shade import EnvStack, synthetic_fg_lut
from geosplatting_b200.splat import GSplatter, RenderableAttrs, Splats

a = bench.parse_args()
dev = torch.device("cuda", 0)
sc = bench.build_scene_host(a)
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
marks = {}
orig_total = Rz.BinCount.total
def total(self):
    marks["t_wait0"] = time.perf_counter()
    r = orig_total(self)
    marks["t_wait1"] = time.perf_counter()
    return r
Rz.BinCount.total = total
gi = [p[k] for k in bench.PARAM_NAMES] + [env_data, exposure]
def step(i):
    t0 = time.perf_counter()
    gs = GSplatter(gaussians=Splats(p["means"], p["scales"], p["quats"], p["normals"], p["opacities"]), rasterize_mode="antialiased")
    attrs = RenderableAttrs(kd=p["kd"], ks=p["ks"], normals=p["normals"])
    img = attrs.splat(gs, [cams[i % 8]], exposure=exposure, envmap=env, fg_lut=lut, min_roughness=0.1, max_metallic=1.0)
    t1 = time.perf_counter()
    g = torch.autograd.grad(img, gi, grad_outputs=v_img)
    t2 = time.perf_counter()
    return t0, t1, t2
for i in range(5): step(i)
torch.cuda.synchronize()
acc = [0.0] * 5
n = 40
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w0 = time.perf_counter(); e0.record()
for i in range(n):
    t0, t1, t2 = step(i)
    acc[0] += marks["t_wait0"] - t0; acc[1] += marks["t_wait1"] - marks["t_wait0"]; acc[2] += t1 - marks["t_wait1"]; acc[3] += t2 - t1
e1.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
print(f"per view: host fwd before wait {acc[0]/n*1e3:.3f} ms | wait for M {acc[1]/n*1e3:.3f} | host fwd after wait {acc[2]/n*1e3:.3f} | "
      f"host bwd {acc[3]/n*1e3:.3f} | wall {(w1-w0)/n*1e3:.3f} | device (events) {e0.elapsed_time(e1)/n:.3f}")
