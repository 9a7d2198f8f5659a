#!/usr/bin/env python
"""bench.py -- training views/sec (fwd+bwd) of the splat + shade hot path on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the CPU oracle on the box's host cores

A "step" is one training view: shade -> project -> bin/sort -> composite forward, then the backward of
all of them for a fixed random image cotangent.  Views are sharded over ranks (one process per GPU,
no data-path collective inside a view; the per-step gradient all-reduce is measured by
`--allreduce`), so `scaling` is "weak": every rank renders its own views of the same Gaussian set.
Prints ONE JSON line (contract in the task statement; keys documented in DESIGN.md section 7).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "training views/sec (fwd+bwd) @1M Gaussians 800x800"
UNIT = "views/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--res", type=int, default=800)
    ap.add_argument("--views", type=int, default=8, help="distinct cameras cycled through per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--allreduce", action="store_true", help="all-reduce the per-Gaussian gradients every --views steps")
    return ap.parse_args()


def workload_name(a):
    return (f"surface-disc Gaussians N={a.gaussians}, {a.res}x{a.res}, antialiased, fwd+bwd per view, "
            f"{a.views} orbit cameras/rank (dataparser intrinsics)")


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
            rows = [r.strip().split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = sorted(float(r[1]) for r in rows if len(r) >= 9)
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = max(float(r[2]) for r in rows if len(r) >= 9)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for k, nm in enumerate(names):
                    if any("Active" in r[5 + k] and "Not" not in r[5 + k] for r in rows if len(r) >= 9):
                        out["reasons"].append(nm)
                out["samples"] = len(sm)
        except Exception as e:  # pragma: no cover
            out["error"] = str(e)
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


# ------------------------------------------------------------------------------------------------
# algorithmic bytes (SURVEY.md section 8d / DESIGN.md section 4)
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(N, Nv, M, P, T, key_bits):
    p = (key_bits + 7) // 8
    per_kernel = {
        "gsb_project_fwd": 44 * N + 40 * Nv,
        "gsb_isect_tiles": 16 * Nv + 12 * M,
        "gsb_sort_pairs": 24 * p * M,
        "gsb_isect_offsets": 8 * M + 4 * T,
        "gsb_composite_fwd": 40 * M + 20 * P,
        "gsb_composite_bwd": 76 * M + 24 * P,
        "gsb_project_bwd": 120 * Nv + 44 * N,
        "gsb_shade_fwd": 56 * N,
        "gsb_shade_bwd": 112 * N,
    }
    return per_kernel


# ------------------------------------------------------------------------------------------------
# this repository's arm
# ------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist

    from geosplatting_b200 import _lib, scenes
    from geosplatting_b200.rasterization import rasterization

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.load()

    g = scenes.surface_gaussians(a.gaussians, seed=0)
    W = H = a.res
    cams = scenes.orbit_cameras(a.views * world, W, H, seed=1)[rank::world]
    host = {k: g[k].contiguous().pin_memory() for k in ("means", "quats", "scales", "opacities", "colors")}
    gen = torch.Generator().manual_seed(1234 + rank)
    v_img_host = torch.randn(H, W, 4, generator=gen).pin_memory()
    params = {k: v.to(dev).requires_grad_(True) for k, v in host.items()}
    v_img = v_img_host.to(dev)
    vms = [torch.from_numpy(c.view_matrix)[None] for c in cams]  # host tensors: no D2H sync to read them
    Ks = [torch.from_numpy(c.intrinsic_matrix)[None] for c in cams]
    names = ("means", "quats", "scales", "opacities", "colors")
    stats = {}

    def step(i, p):
        c = i % len(cams)
        render, alpha, info = rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"],
                                            vms[c], Ks[c], W, H, packed=True, rasterize_mode="antialiased")
        rgba = torch.cat((render[0], alpha[0]), dim=-1)
        grads = torch.autograd.grad(rgba, [p[k] for k in names], grad_outputs=v_img)
        stats["info"] = info
        return rgba, grads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # inputs (76 MB) + lists (~100 MB) exceed nothing like L2=126 MB on their own, so flush L2 between
    # iterations by writing a 256 MB buffer (outside the per-kernel event pairs, inside the step loop).
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    for i in range(a.warmup):
        step(i, params)
    barrier()

    # ---- timed region: EXACTLY K steps, device-timed, per-kernel events on ---------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.CallStats.reset(timing=True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    for i in range(a.steps):
        flush.zero_()
        ev[i][0].record()
        step(i, params)
        ev[i][1].record()
    barrier()
    step_ms = [s.elapsed_time(e) for s, e in ev]
    total_ms = sum(step_ms)
    durations = _lib.CallStats.durations_ms()
    launches = _lib.CallStats.launches()
    clocks = sampler.stop() if rank == 0 else None
    _lib.CallStats.reset(timing=False)

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = a.steps * world / (total_ms_max / 1e3)

    # ---- end-to-end through the public API with HOST buffers ----------------------------------------
    e2e = None
    if not a.no_e2e:
        out_host = torch.empty(H, W, 4).pin_memory()
        grad_host = {k: torch.empty_like(host[k]).pin_memory() for k in names}
        n_e2e = max(3, min(a.steps, 10))

        def e2e_step(i):
            p = {k: host[k].to(dev, non_blocking=True).requires_grad_(True) for k in names}
            vimg = v_img_host.to(dev, non_blocking=True)
            c = i % len(cams)
            render, alpha, _ = rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"],
                                             vms[c], Ks[c], W, H, packed=True, rasterize_mode="antialiased")
            rgba = torch.cat((render[0], alpha[0]), dim=-1)
            grads = torch.autograd.grad(rgba, [p[k] for k in names], grad_outputs=vimg)
            out_host.copy_(rgba.detach(), non_blocking=True)
            for k, gk in zip(names, grads):
                grad_host[k].copy_(gk, non_blocking=True)

        for i in range(2):
            e2e_step(i)
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(n_e2e):
            e2e_step(i)
        e.record()
        barrier()
        te = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = sum(host[k].numel() * 4 for k in names) + v_img_host.numel() * 4
        d2h = out_host.numel() * 4 + sum(grad_host[k].numel() * 4 for k in names)
        e2e = {"value": n_e2e * world / (float(te.item()) / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": n_e2e}

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    info = stats["info"]
    M = int(info["flatten_gaussian_ids"].shape[0])
    Nv = int(info["gaussian_ids"].shape[0])
    tw, th = (W + 15) // 16, (H + 15) // 16
    key_bits = 32 + (tw * th).bit_length()
    alg = algorithmic_bytes(a.gaussians, Nv, M, W * H, tw * th, key_bits)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    per_kernel = {}
    for k, (calls, ms) in durations.items():
        if calls == 0 or ms <= 0:
            continue
        avg_ms = ms / calls
        per_kernel[k] = {"calls": calls, "avg_ms": round(avg_ms, 4), "share": round(ms / total_ms, 4)}
        if k in alg:
            gbs = alg[k] / (avg_ms * 1e-3) / 1e9
            per_kernel[k].update({"alg_bytes": alg[k], "gbs": round(gbs, 1), "frac": round(gbs / peak, 4)})
    dom = max((k for k in per_kernel if k in alg), key=lambda k: per_kernel[k]["share"])
    evals = 256.0 * M
    roofline = {"bound": "hbm", "kernel": dom, "achieved": per_kernel[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": per_kernel[dom]["frac"], "traffic": None, "peak_source": peak_kind,
                "alg_bytes_per_launch": alg[dom], "avg_ms": per_kernel[dom]["avg_ms"],
                "note": "composite kernels are FP32/MUFU-issue bound, not HBM bound (DESIGN.md section 4); "
                        "pixel x Gaussian evaluation upper bound per launch = 256*M",
                "pix_gauss_evals_upper": evals}

    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": round(total_ms_max / a.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "gaussians": a.gaussians, "visible": Nv, "intersections": M,
                       "resolution": [W, H], "views_per_rank": len(cams), "l2": "flushed between steps (256 MB write)",
                       "parallelism": f"views sharded over {world} rank(s)"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": per_kernel,
            "impl": "b200",
        }
    return out, rank, world


# ------------------------------------------------------------------------------------------------
# CPU oracle arm (also used for cpu_baseline)
# ------------------------------------------------------------------------------------------------
def run_oracle(a, steps, warmup, budget_s=150.0):
    """fwd+bwd of one view per step with the CPU oracle on all host threads."""
    import numpy as np

    from geosplatting_b200 import scenes
    from oracle import raster as R

    g = scenes.surface_gaussians(a.gaussians, seed=0)
    gn = {k: g[k].numpy() for k in ("means", "quats", "scales", "opacities", "colors")}
    W = H = a.res
    cams = scenes.orbit_cameras(a.views, W, H, seed=1)
    rng = np.random.default_rng(1234)
    vr = rng.normal(size=(H, W, 3)).astype(np.float32)
    va = rng.normal(size=(H, W, 1)).astype(np.float32)
    cores = R.num_threads()

    def step(i):
        c = cams[i % len(cams)]
        ocam = R.Camera(c.view_matrix, c.fx, c.fy, c.cx, c.cy, c.width, c.height)
        render, alpha, info = R.rasterization(gn["means"], gn["quats"], gn["scales"], gn["opacities"], gn["colors"],
                                              ocam, rasterize_mode="antialiased")
        R.rasterization_bwd(gn["means"], gn["quats"], gn["scales"], gn["opacities"], gn["colors"], ocam, info, alpha,
                            vr, va, rasterize_mode="antialiased")

    t0 = time.perf_counter()
    step(0)
    t_first = time.perf_counter() - t0
    # bound the run: never more than budget_s of CPU work in total
    max_steps = max(1, int(budget_s / max(t_first, 1e-3)))
    warm = min(warmup, max(0, max_steps // 4))
    timed = max(1, min(steps, max_steps - warm))
    for i in range(1, warm):
        step(i)
    t0 = time.perf_counter()
    for i in range(timed):
        step(i)
    dt = time.perf_counter() - t0
    return {"value": timed / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{timed} full view(s) fwd+bwd of the same workload (C oracle, OpenMP {cores} threads; "
                      f"raster only until the shade oracle lands)", "ms_per_step": dt / timed * 1e3,
            "steps_timed": timed}


def main():
    a = parse_args()
    if a.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return
        cb = run_oracle(a, a.steps, a.warmup)
        out = {"metric": METRIC, "value": round(cb["value"], 5), "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": round(cb["ms_per_step"], 2), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": workload_name(a), "gaussians": a.gaussians, "resolution": [a.res, a.res]},
               "impl": "reference", "cpu_baseline": cb,
               "e2e": {"value": round(cb["value"], 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out))
        return
    out, rank, world = run_b200(a)
    if rank == 0:
        if not a.no_cpu_baseline and world == 1:
            cb = run_oracle(a, steps=2, warmup=1, budget_s=25.0)
            out["cpu_baseline"] = {k: (round(v, 5) if isinstance(v, float) else v) for k, v in cb.items()}
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
