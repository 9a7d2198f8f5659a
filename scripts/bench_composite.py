#!/usr/bin/env python
"""Per-entry-point timings of one view of the bench scene (1 002 528 MGAdaptor Gaussians, 800x800), one view at a time on
one stream, L2 flushed between views: what bench.py's instrumented pass measures, without the rest of bench.py.
    python scripts/bench_composite.py [--iters 20] [--mesh-n 118] [--res 800]
GSB_LIB_PATH selects a tuning build of the library."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from geosplatting_b200 import _lib, scenes, splitsum  # noqa: E402
from geosplatting_b200.fused import splat_view  # noqa: E402
from geosplatting_b200.mgadapter import MGAdapter, compute_vertex_normals  # noqa: E402
from geosplatting_b200.shade import EnvStack, synthetic_fg_lut  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--mesh-n", type=int, default=118)
    ap.add_argument("--res", type=int, default=800)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    verts, faces = scenes.cube_sphere(a.mesh_n)
    gen = torch.Generator().manual_seed(0)
    with torch.no_grad():
        vd, fd = verts.to(dev), faces.to(dev)
        sp, _ = MGAdapter().make(vd, fd, compute_vertex_normals(vd, fd))
        N = sp.means.shape[0]
        kd = (torch.rand(N, 3, generator=gen) * 0.8 + 0.1).to(dev)
        ks = torch.rand(N, 2, generator=gen).to(dev)
        cube = torch.exp(torch.randn(6, 512, 512, 3, generator=gen)).clamp_min(1e-2).to(dev)
        env0 = splitsum.as_envstack(cube)
    p = [t.detach().clone().requires_grad_(True) for t in (sp.means, sp.scales, sp.quats, sp.opacities, kd, ks, sp.colors)]
    env = EnvStack(env0.data.detach().clone().requires_grad_(True), env0.R0, env0.L, env0.Rb)
    ex = torch.ones(1, device=dev, requires_grad=True)
    lut = synthetic_fg_lut(dev)
    cams = scenes.orbit_cameras(8, a.res, a.res, seed=1)
    cot = torch.randn(a.res, a.res, 4, generator=gen).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step(i):
        img = splat_view(*p, cams[i % 8], exposure=ex, envmap=env, fg_lut=lut, min_roughness=0.1, max_metallic=1.0,
                         native=False)
        torch.autograd.grad(img, p + [env.data, ex], grad_outputs=cot)

    for i in range(3):
        step(i)
    _lib.CallStats.reset(timing=True)
    torch.cuda.synchronize()
    for i in range(a.iters):
        flush.zero_()
        step(i)
    d = _lib.CallStats.durations_ms()
    out = {k: round(ms / c, 4) for k, (c, ms) in d.items() if c and ms / c > 0.004}
    out["sum_ms_per_view"] = round(sum(ms / c * (c / a.iters) for k, (c, ms) in d.items() if c), 4)
    out["lib"] = os.environ.get("GSB_LIB_PATH", "default")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
