#!/usr/bin/env python
"""bench.py -- training views/sec (fwd+bwd) of the splat + PBR-shade hot path on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the CPU oracle on the box's host cores

A "step" is one training view of BASELINE.json's metric configuration (1M Gaussians, 800x800): EWA projection -> depth
order / intersection count -> split-sum shade -> tile binning -> alpha composite -> tone map, then the backward of all of
them for a fixed random image cotangent (gradients to means / log-scales / quats / opacities / kd / ks / normals /
env-map texels / exposure).  The Gaussians come from the MGAdaptor kernel on a 167k-face mesh.  The K views run the way
the reference's trainer runs them -- batches of `--views` (8): all forwards, one backward over the batch -- through
fused.splat_views (one autograd node per batch, views software-pipelined over `--streams` CUDA streams, three native
calls per view).  Views are sharded over ranks (one process per GPU, no data-path collective inside a view; once per
batch the packed per-Gaussian gradients are summed with ONE asynchronous NCCL all-reduce, inside the timed region), so
`scaling` is "weak".  Prints ONE JSON line; keys are documented in DESIGN.md section 7.
"""
from __future__ import annotations

import argparse
import gc
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
PARAM_NAMES = ("means", "scales", "quats", "opacities", "kd", "ks", "normals")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=80)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mesh-n", type=int, default=118, help="cube-sphere subdivision: 72*n*n Gaussians (118 -> 1.0M)")
    ap.add_argument("--res", type=int, default=800)
    ap.add_argument("--light-res", type=int, default=512, help="env cube-map resolution (GeoSplatter.light_resolution)")
    ap.add_argument("--views", type=int, default=8, help="batch: distinct cameras per rank, forwarded then back-propagated together (the trainer's batch of 8); one gradient all-reduce per batch when N > 1")
    ap.add_argument("--streams", type=int, default=4, help="CUDA streams the views of a batch are spread over")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of BASELINE configs 2 / 4 / 5")
    ap.add_argument("--full-step", action="store_true", help="also time a full train step (a1-a12, B=8 views)")
    ap.add_argument("--no-train-step", action="store_true",
                    help="skip the whole-training-step probe of BASELINE config 3 (model.GeoSplatter + loss + Adam)")
    ap.add_argument("--fc-res", type=int, default=0,
                    help="with --full-step: the mesh comes from FlexiCubes on an R^3 SDF grid every step (8f rank 3)")
    return ap.parse_args()


def n_gaussians(a):
    return 72 * a.mesh_n * a.mesh_n


def workload_name(a):
    return (f"MGAdaptor Gaussians of a {12 * a.mesh_n ** 2}-face bumpy sphere (N={n_gaussians(a)}), {a.res}x{a.res}, "
            f"split-sum PBR shade (env cube {a.light_res}^2, 6 mips) + antialiased raster + tone map, fwd+bwd per view, "
            f"{a.views} orbit cameras/rank with the dataparser intrinsics")


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            time.sleep(0.3)
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
            rows = [r for r in rows if len(r) >= 9]
            sm = sorted(float(r[1]) for r in rows)
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = max(float(r[2]) for r in rows)
                out["power_w_max"] = max(float(r[3]) for r in rows)
                for k, nm in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
                    if any(r[5 + k].strip() == "Active" for r in rows):
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


def algorithmic_bytes(N, Nv, M, P, T, key_bits):
    """Per-launch algorithmic bytes."""
    p = (key_bits + 7) // 8
    tp = ((T.bit_length()) + 7) // 8          # radix passes over the tile id alone (two-stage binning)
    return {
        # two-stage binning: 4 passes over N (key,index) pairs + ordered scan | emission + tp passes over M pairs + offsets
        "gsb_bin2_count": 16 * 4 * N + 16 * N, "gsb_bin2_sort": 12 * N + 12 * Nv + 8 * M + 16 * tp * M + 4 * M + 4 * T,
        "gsb_tonemap_planar_fwd": 32 * P, "gsb_tonemap_planar_bwd": 44 * P,
        "gsb_shade_fwd": 56 * N, "gsb_shade_bwd": 68 * N + 44 * N,
        "gsb_project_fwd": 44 * N + 40 * Nv, "gsb_project_bwd": 120 * Nv + 44 * N,
        "gsb_isect_tiles": 16 * Nv + 12 * M, "gsb_sort_pairs": 24 * p * M, "gsb_isect_offsets": 8 * M + 4 * T,
        "gsb_composite_fwd": 40 * M + 20 * P, "gsb_composite_bwd": 76 * M + 24 * P,
        "gsb_tonemap_fwd": 32 * P, "gsb_tonemap_bwd": 48 * P,
    }


# ------------------------------------------------------------------------------------------------
def build_scene_host(a):
    """CPU tensors of the scene: mesh, per-Gaussian material attributes, env cube map, cameras."""
    import torch

    from geosplatting_b200 import scenes
    verts, faces = scenes.cube_sphere(a.mesh_n)
    N = 6 * faces.shape[0]
    g = torch.Generator().manual_seed(0)
    kd = torch.rand(N, 3, generator=g) * 0.8 + 0.1
    ks = torch.rand(N, 2, generator=g)
    cubemap = torch.exp(torch.randn(6, a.light_res, a.light_res, 3, generator=g)).clamp_min(1e-2)
    return dict(verts=verts, faces=faces, kd=kd, ks=ks, cubemap=cubemap)


def run_b200(a):
    import torch
    import torch.distributed as dist

    from geosplatting_b200 import _lib, scenes, splitsum
    from geosplatting_b200.mgadapter import MGAdapter, compute_vertex_normals
    from geosplatting_b200.parallel import FlatAllReduce, shard_views
    from geosplatting_b200.shade import EnvStack, synthetic_fg_lut
    from geosplatting_b200.splat import GSplatter, RenderableAttrs, Splats

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    sc = build_scene_host(a)
    W = H = a.res
    cams = shard_views(scenes.orbit_cameras(a.views * world, W, H, seed=1), rank, world)
    lut = synthetic_fg_lut(dev)
    with torch.no_grad():
        vd, fd = sc["verts"].to(dev), sc["faces"].to(dev)
        sp, _ = MGAdapter().make(vd, fd, compute_vertex_normals(vd, fd))
        env0 = splitsum.as_envstack(sc["cubemap"].to(dev))
    host = {"means": sp.means, "scales": sp.scales, "quats": sp.quats, "opacities": sp.opacities,
            "kd": sc["kd"], "ks": sc["ks"], "normals": sp.colors}
    host = {k: v.detach().cpu().contiguous().pin_memory() for k, v in host.items()}
    N = host["means"].shape[0]
    params = {k: v.to(dev).requires_grad_(True) for k, v in host.items()}
    env_data = env0.data.detach().clone().requires_grad_(True)
    env = EnvStack(env_data, env0.R0, env0.L, env0.Rb, env0.min_roughness, env0.max_roughness)
    exposure = torch.ones(1, device=dev, requires_grad=True)
    gen = torch.Generator().manual_seed(1234 + rank)
    v_img_host = torch.randn(H, W, 4, generator=gen).pin_memory()
    v_img = v_img_host.to(dev)

    from geosplatting_b200.fused import splat_view, splat_views

    def render(p, env_, ex, cam):
        gs = GSplatter(gaussians=Splats(p["means"], p["scales"], p["quats"], p["normals"], p["opacities"]),
                       rasterize_mode="antialiased")
        attrs = RenderableAttrs(kd=p["kd"], ks=p["ks"], normals=p["normals"])
        return attrs.splat(gs, [cam], exposure=ex, envmap=env_, fg_lut=lut, min_roughness=0.1, max_metallic=1.0)

    grad_inputs = [params[k] for k in PARAM_NAMES] + [env_data, exposure]
    B = max(1, a.views)                     # the reference's batch: 8 views per step of the trainer
    # the trainer's loss is the mean over the views of the (global) batch: that 1 / (views x ranks) rides in the image
    # cotangent, so the summed gradients need no scaling pass
    v_img_batch = v_img / float(B * world) if world > 1 else v_img
    pending = {"work": None, "nbytes": 0}

    def run_views(i0, n):
        """n consecutive views (steps) the way GeoSplatter.render_report + the trainer run a batch: all forwards,
        then one backward over the sum (geosplat.py:869-879, geosplat_trainer.py:171-180) -- here with the views
        spread over two CUDA streams (fused.splat_views) so that neighbouring views overlap on the device."""
        cs = [cams[(i0 + j) % len(cams)] for j in range(n)]
        imgs = splat_views(params["means"], params["scales"], params["quats"], params["opacities"], params["kd"],
                           params["ks"], params["normals"], cs, exposures=exposure, envmap=env, fg_lut=lut,
                           min_roughness=0.1, max_metallic=1.0, n_streams=a.streams)
        grads = torch.autograd.grad(imgs, grad_inputs, grad_outputs=[v_img_batch] * n)
        if world > 1:
            if pending["work"] is not None:
                pending["work"].wait()         # the previous batch's collective
            # ONE all-reduce, in place on the flat buffer the batched backward wrote (no pack); asynchronous, it
            # overlaps the next batch
            pending["work"] = FlatAllReduce(list(grads), async_op=True)
            pending["nbytes"] = pending["work"].nbytes
        return imgs, grads

    batch_marks = []

    def run_steps(k, marks=None):
        i = 0
        while i < k:
            n = min(B, k - i)
            if marks is not None:
                f0 = torch.cuda.Event(enable_timing=True); f0.record()
            flush.zero_()                      # L2 flush between batches (a view's own working set is ~4x L2)
            if marks is not None:
                e0 = torch.cuda.Event(enable_timing=True); e0.record()   # also closes the flush's own window (f0, e0)
                h0 = time.perf_counter()
            run_views(i, n)
            if marks is not None:
                e1 = torch.cuda.Event(enable_timing=True); e1.record()
                marks.append((n, e0, e1, time.perf_counter() - h0, f0))
            i += n

    def step(i):
        """One view at a time on the current stream, every kernel group launched from Python (native=False) so that
        each C-ABI entry point can be bracketed by CUDA events: the instrumented pass."""
        img = splat_view(params["means"], params["scales"], params["quats"], params["opacities"], params["kd"],
                         params["ks"], params["normals"], cams[i % len(cams)], exposure=exposure, envmap=env, fg_lut=lut,
                         min_roughness=0.1, max_metallic=1.0, native=False)
        return img, torch.autograd.grad(img, grad_inputs, grad_outputs=v_img)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # One view touches ~400 MB (Gaussian state 76 MB, projection 48 MB, records 48 MB, lists ~60 MB, gradients
    # 72 MB, env stack 34 MB x2) against a 126 MB L2, and two views are in flight at a time, so no view finds its
    # inputs cached; on top of that a 256 MB write flushes L2 between batches.
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    n_warm = max(a.warmup, 2 * B)           # at least two full batches: the allocator must have seen a batch's peak
    def wait_collective():
        if pending["work"] is not None:
            pending["work"].wait()
            pending["work"] = None

    run_steps(n_warm)
    wait_collective()
    barrier()

    # ---- timed region: EXACTLY K steps (views), device-timed ----------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    gc.collect()
    run_steps(B)                               # the sampler's start-up left the GPU idle for 0.3 s: clocks back up
    wait_collective()
    _lib.CallStats.reset(timing=False)         # launch counters only ...
    import geosplatting_b200.fused as fused_mod
    fused_mod.PROBES = []                      # ... plus one event pair per view around the compositing backward
    barrier()
    wall0 = time.perf_counter()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    seg0 = torch.cuda.memory_stats(dev).get("segment.all.allocated", 0)
    # CPython's cyclic collector walks every live object when its generation-2 threshold trips (~30 ms with torch
    # loaded): a stop-the-world pause longer than four views.  Nothing in a view creates reference cycles.
    gc.disable()
    main_trace = os.environ.get("GSB_TRACE")   # diagnosis: chrome trace of the timed region (the profiler slows it down)
    if main_trace:
        from torch.profiler import ProfilerActivity, profile
        main_prof = profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA])
        main_prof.__enter__()
    t_begin.record()
    run_steps(a.steps, batch_marks)
    wait_collective()                          # the last collective completes inside the timed region
    t_end.record()
    barrier()
    if main_trace:
        main_prof.__exit__(None, None, None)
        main_prof.export_chrome_trace(main_trace)
    gc.enable()
    wall = time.perf_counter() - wall0
    batches = {"views": [m[0] for m in batch_marks], "device_ms": [round(m[1].elapsed_time(m[2]), 3) for m in batch_marks],
               "host_enqueue_ms": [round(m[3] * 1e3, 3) for m in batch_marks],
               "cuda_mallocs_in_region": torch.cuda.memory_stats(dev).get("segment.all.allocated", 0) - seg0}
    batches["l2_flush_ms_in_region"] = None
    launches = _lib.CallStats.launches()
    n_batches_timed = max(1, len(batch_marks))
    batches["host_in_library_ms_per_batch"] = {k: round(v * 1e3 / n_batches_timed, 4)
                                               for k, v in _lib.CallStats.host_s.items() if v * 1e3 / n_batches_timed > 0.005}
    batches["note"] = ("host_enqueue_ms is wall time around a batch's forward + backward as the host sees it; it includes the "
                       "backward's look at the forward's intersection counts (an event wait that ends when the device, one "
                       "batch behind, finishes that forward) -- the time the host spends enqueueing is host_in_library_ms_per_batch")
    probes, fused_mod.PROBES = fused_mod.PROBES, None
    live_ms = [a_.elapsed_time(b_) for a_, b_ in probes]
    live_bwd_ms = sum(live_ms) / max(1, len(live_ms))
    clocks = sampler.stop() if rank == 0 else None
    # the L2-flush writes are not part of a step: each one runs alone on the caller's stream between two batches (the
    # previous batch has been joined, the next not yet forked) inside its own event window, which is taken off the region
    torch.cuda.synchronize()
    flush_ms = sum(m[4].elapsed_time(m[1]) for m in batch_marks)
    total_ms = t_begin.elapsed_time(t_end) - flush_ms
    batches["l2_flush_ms_in_region"] = round(flush_ms, 3)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = a.steps * world / (total_ms_max / 1e3)

    # ---- instrumented pass: the same K views one at a time on one stream, every entry point timed -------------
    # (kernels of overlapping views share the SMs, so their isolated durations have to be measured sequentially)
    for i in range(2):
        step(i)
    _lib.CallStats.reset(timing=True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    torch.cuda.synchronize()
    for i in range(a.steps):
        flush.zero_()
        ev[i][0].record()
        step(i)
        ev[i][1].record()
    torch.cuda.synchronize()
    seq_ms = sum(s_.elapsed_time(e_) for s_, e_ in ev)
    durations = _lib.CallStats.durations_ms()
    _lib.CallStats.reset(timing=False)

    # ---- end to end through the public API with HOST buffers ------------------------------------------------
    # The unit the trainer hands the path is a BATCH: one Gaussian state, B cameras, B image cotangents in; B images and
    # the batch's summed gradients out (rfstudio/trainer/geosplat_trainer.py:171-180).  Per batch, inside the timed region:
    # pinned host Gaussian state (once) + B cotangent images -> device, fused.splat_views forward + backward, B images +
    # every gradient -> pinned host.  Double-buffered over three streams: the upload of batch k+1 and the download of
    # batch k-1 overlap the compute of batch k (PCIe is full duplex).
    e2e = None
    if not a.no_e2e:
        # at least 6 batches whatever --steps says: the region ends when the LAST batch's results have landed in host
        # memory, and that drain (190 MB) is 30 % of a two-batch region but not of a trainer's steady state; the number
        # of views timed is reported as e2e.steps
        n_b = max(6, min(12, a.steps // B))
        _E2E_DEBUG = os.environ.get("GSB_E2E_DEBUG", "")   # 'noup' / 'nodown': diagnosis only (an e2e without its copies is not an e2e)
        cot_host = [torch.randn(H, W, 4, generator=gen).pin_memory() for _ in range(B)]
        img_host = [[torch.empty(H, W, 4).pin_memory() for _ in range(B)] for _ in range(2)]
        grad_host = [{k: torch.empty_like(host[k]).pin_memory() for k in PARAM_NAMES} for _ in range(2)]
        ex_host = torch.empty(1).pin_memory()
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        s_cmp = torch.cuda.current_stream(dev)
        dev_in = [{k: torch.empty_like(host[k], device=dev) for k in PARAM_NAMES} for _ in range(2)]
        dev_cot = [[torch.empty(H, W, 4, device=dev) for _ in range(B)] for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]

        def upload(i):
            b_ = i % 2
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_free[b_])            # the compute that last read this buffer set is done
                if _E2E_DEBUG != "noup":
                    for k in PARAM_NAMES:
                        dev_in[b_][k].copy_(host[k], non_blocking=True)
                    for j in range(B):
                        dev_cot[b_][j].copy_(cot_host[j], non_blocking=True)
                ev_in[b_].record(s_in)

        def e2e_batch(i):
            b_ = i % 2
            upload(i + 1)
            s_cmp.wait_event(ev_in[b_])
            s_cmp.wait_event(ev_out[b_])                # the download that last used these host-bound tensors is done
            p = {k: dev_in[b_][k].detach().requires_grad_(True) for k in PARAM_NAMES}
            cs = [cams[(i * B + j) % len(cams)] for j in range(B)]
            imgs = splat_views(p["means"], p["scales"], p["quats"], p["opacities"], p["kd"], p["ks"], p["normals"], cs,
                               exposures=exposure, envmap=env, fg_lut=lut, min_roughness=0.1, max_metallic=1.0,
                               n_streams=a.streams)
            grads = torch.autograd.grad(imgs, [p[k] for k in PARAM_NAMES] + [exposure], grad_outputs=dev_cot[b_])
            ev_free[b_].record(s_cmp)
            done = torch.cuda.Event()
            done.record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                for j, im in enumerate(imgs):
                    im_d = im.detach()
                    im_d.record_stream(s_out)
                    if _E2E_DEBUG != "nodown":
                        img_host[b_][j].copy_(im_d, non_blocking=True)
                for k, gk in zip(PARAM_NAMES, grads):
                    gk.record_stream(s_out)
                    if _E2E_DEBUG != "nodown":
                        grad_host[b_][k].copy_(gk, non_blocking=True)
                grads[-1].record_stream(s_out)
                ex_host.copy_(grads[-1], non_blocking=True)
                ev_out[b_].record(s_out)

        for b0 in range(2):
            ev_free[b0].record(s_cmp)
            ev_out[b0].record(s_cmp)
        upload(0)
        for i in range(2):
            e2e_batch(i)
        barrier()
        upload(2)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gc.collect()
        gc.disable()
        trace = os.environ.get("GSB_E2E_TRACE")            # diagnosis: chrome trace of the e2e region (slows it down)
        if trace:
            from torch.profiler import ProfilerActivity, profile
            prof = profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA])
            prof.__enter__()
        s.record()
        for i in range(2, 2 + n_b):
            e2e_batch(i)
        s_cmp.wait_stream(s_out)                          # the last batch's results have landed in host memory
        e.record()
        if trace:
            torch.cuda.synchronize()
            prof.__exit__(None, None, None)
            prof.export_chrome_trace(trace)
        barrier()
        gc.enable()
        te = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        state_b = sum(host[k].numel() * 4 for k in PARAM_NAMES)
        h2d = (state_b + B * H * W * 16) / B
        d2h = (B * H * W * 16 + sum(grad_host[0][k].numel() * 4 for k in PARAM_NAMES) + 4) / B
        e2e = {"value": round(n_b * B * world / (float(te.item()) / 1e3), 3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": n_b * B, "views_per_batch": B,
               "what": "per batch of 8 views (the trainer's unit): pinned host Gaussian state (76 MB, once) + 8 image "
                       "cotangents -> device, fused.splat_views forward + backward, 8 images + all per-Gaussian gradients "
                       "of the batch -> pinned host; bytes are per view (batch bytes / 8); uploads / downloads of "
                       "neighbouring batches overlap the compute on separate streams (double-buffered)"}

    # ---- optional: one full train step (a1-a12, B = 8 views) ------------------------------------------------
    full = None
    if a.full_step:
        from geosplatting_b200.field import GaussianField
        torch.manual_seed(0)
        field = GaussianField().to(dev)
        vd = sc["verts"].to(dev).requires_grad_(True)
        cube = sc["cubemap"].to(dev).requires_grad_(True)
        guess = torch.zeros(2, device=dev)
        field_params = [t for t in field.parameters()]
        fc_grid = None
        if a.fc_res > 0:
            # GeoSplatter.get_geometry (geosplat.py:751-769): SDF grid -> FlexiCubes mesh + regulariser, every step
            from geosplatting_b200.flexicubes import FlexiCubes
            R_ = a.fc_res
            fc_grid = FlexiCubes.from_resolution(R_, random_sdf=False, scale=0.9, device=dev)
            gv_ = fc_grid.vertices
            sdf_p = (gv_.norm(dim=-1, keepdim=True) - 0.6 + 0.06 * torch.sin(5.0 * gv_[:, :1]) * torch.cos(4.0 * gv_[:, 1:2]))
            sdf_p = sdf_p.clone().requires_grad_(True)
            deform_p = torch.zeros_like(gv_).requires_grad_(True)
            weight_p = torch.zeros(fc_grid.indices.shape[0], 21, device=dev).requires_grad_(True)
            fc_stats = {}

        def geometry():
            vertices = fc_grid.vertices + deform_p.tanh() * (0.5 * 0.9 / a.fc_res)
            fc_ = fc_grid.replace(vertices=vertices, sdf_values=sdf_p, alpha=weight_p[:, :8], beta=weight_p[:, 8:20],
                                  gamma=weight_p[:, 20:])
            mesh_, l_dev_ = fc_.dual_marching_cubes()
            reg_ = l_dev_.mean() * 0.5 + weight_p[:, :20].abs().mean() * 0.1 + fc_.compute_entropy() * 0.3
            fc_stats.update(mesh_vertices=mesh_.vertices.shape[0], faces=mesh_.indices.shape[0],
                            gaussians=6 * mesh_.indices.shape[0])
            return mesh_, reg_

        def full_step():
            if fc_grid is not None:
                mesh_, reg_ = geometry()
                spl, at, _ = field.get_gaussians_from_face(mesh_.vertices, mesh_.indices, 0.0, 0.0, scale=0.9,
                                                           initial_guess=guess)
                e_ = splitsum.as_envstack(cube)
                imgs = splat_views(spl.means, spl.scales, spl.quats, spl.opacities, at.kd, at.ks, at.normals, cams[:8],
                                   exposures=exposure, envmap=e_, fg_lut=lut, min_roughness=0.1, max_metallic=1.0,
                                   n_streams=a.streams)
                torch.autograd.grad(list(imgs) + [reg_], [sdf_p, deform_p, weight_p, cube, exposure] + field_params,
                                    grad_outputs=[v_img] * len(imgs) + [torch.ones_like(reg_)])
                return
            _full_step_fixed_mesh()

        def _full_step_fixed_mesh():
            """One training step of stage 1 without the optimiser and the loss (a random image cotangent stands in):
            vertex normals + MGAdaptor + the three hash-grid fields (a1-a3), split-sum prefilter (a4/a5), the batch of
            8 views (a7-a12), and the backward of all of it down to vertices, tables, MLP weights, cube map, exposure."""
            spl, at, _ = field.get_gaussians_from_face(vd, fd, 0.0, 0.0, scale=0.9, initial_guess=guess)
            e_ = splitsum.as_envstack(cube)
            imgs = splat_views(spl.means, spl.scales, spl.quats, spl.opacities, at.kd, at.ks, at.normals, cams[:8],
                               exposures=exposure, envmap=e_, fg_lut=lut, min_roughness=0.1, max_metallic=1.0,
                               n_streams=a.streams)
            torch.autograd.grad(imgs, [vd, cube, exposure] + field_params, grad_outputs=[v_img] * len(imgs))

        for _ in range(2):
            full_step()
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_full = 3
        _lib.CallStats.reset(timing=True)
        gc.collect()
        gc.disable()
        s.record()
        for _ in range(n_full):
            full_step()
        e.record()
        barrier()
        gc.enable()
        fk = {k: round(ms_ / n_full, 3) for k, (c_, ms_) in _lib.CallStats.durations_ms().items() if ms_ / n_full > 0.05}
        _lib.CallStats.reset(timing=False)
        ms = s.elapsed_time(e) / n_full
        full = {"ms_per_step": round(ms, 3), "views_per_step": min(8, len(cams)),
                "views_per_s": round(min(8, len(cams)) / (ms / 1e3), 2),
                "what": "vertex normals + MGAdaptor + kd/ks/z hash-grid fields + prefilter + 8 views, forward and "
                        "backward, one rank (entry points of overlapping streams share the SMs: their ms are elapsed, "
                        "not isolated)",
                "entry_point_ms_per_step": fk}
        if fc_grid is not None:
            full["flexicubes"] = dict(resolution=a.fc_res, **fc_stats)
            full["what"] = "FlexiCubes mesh extraction + regulariser (SDF grid -> mesh, every step) + " + full["what"]

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------
    M, Nv = stats_last_view(params, cams[0], W, H)
    tw, th = (W + 15) // 16, (H + 15) // 16
    key_bits = 32 + (tw * th).bit_length()
    alg = algorithmic_bytes(N, Nv, M, W * H, tw * th, key_bits)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    per_kernel = {}
    for k, (calls, ms) in durations.items():
        if calls == 0 or ms <= 0:
            continue
        avg_ms = ms / calls
        per_kernel[k] = {"calls": calls, "avg_ms": round(avg_ms, 4), "share": round(ms / seq_ms, 4)}
        if k in alg:
            gbs = alg[k] / (avg_ms * 1e-3) / 1e9
            per_kernel[k].update({"alg_bytes": alg[k], "gbs": round(gbs, 1), "frac": round(gbs / peak, 4)})
    dom = max((k for k in per_kernel if k in alg), key=lambda k: per_kernel[k]["share"])
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    live = (dom == "gsb_composite_bwd" and live_bwd_ms > 0)
    dom_ms = live_bwd_ms if live else per_kernel[dom]["avg_ms"]
    dom_gbs = alg[dom] / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(dom_gbs, 1), "peak": peak, "unit": "GB/s",
                "frac": round(dom_gbs / peak, 4), "traffic": traffic, "peak_source": peak_kind,
                "alg_bytes_per_launch": alg[dom], "avg_ms": round(dom_ms, 4),
                "avg_ms_alone": per_kernel[dom]["avg_ms"], "frac_alone": per_kernel[dom]["frac"],
                "launches_timed": len(live_ms) if live else per_kernel[dom]["calls"],
                "note": "avg_ms / achieved / frac: CUDA events recorded by the batch driver around the compositing "
                        "backward of every view INSIDE the timed region (gsb_batch_backward probe events), where the "
                        "kernels of up to four views share the SMs; avg_ms_alone / frac_alone: the same entry point in "
                        "the instrumented pass (one view at a time).  The kernel is NOT HBM bound (DESIGN.md section 4): "
                        "every sub-list entry is evaluated at the 16 pixel centres of a 4x4 unit out of shared memory; "
                        "ncu: L1 / shared-memory data pipe 78-85 % busy, issue 50-56 %, DRAM < 3 %",
                "pix_gauss_evals_upper_per_launch": 256.0 * M}

    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "warmup_views_run": n_warm, "ms_per_step": round(total_ms_max / a.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "gaussians": N, "visible": Nv, "intersections": M,
                       "resolution": [W, H], "views_per_rank": len(cams),
                       "batch": f"{B} views forwarded, then back-propagated together, spread over {a.streams} CUDA streams",
                       "l2": "per-view working set ~400 MB (inputs 110 MB) vs 126 MB L2; 256 MB flush write between batches",
                       "parallelism": f"views sharded over {world} rank(s)" +
                                      (f", 1 NCCL all-reduce of {pending['nbytes']} B per {B} views, in place on the backward's flat "
                                       f"gradient buffer" if world > 1 else "")},
            "sequential_ms_per_view": round(seq_ms / a.steps, 4), "batches": batches,
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": per_kernel,
            "kernels_note": "instrumented pass: one view at a time on one stream through the per-view operators, every "
                            "C-ABI entry point bracketed by CUDA events.  Two differences from the timed batch path: its "
                            "second binning stage is the radix sort (gsb_bin2_sort) where the batch driver runs the "
                            "one-pass tile partition (csrc/tilepart.cu, ~0.08 ms instead of ~0.115), and it packs the "
                            "compositing records in gsb_composite_fwd where the batch driver's shade forward writes them",
            "wall_s_timed_region": round(wall, 3), "impl": "b200",
        }
        if full:
            out["full_step"] = full
    # ---- parity of the headline configuration: the view the CPU baseline leg times, compared with the CUDA path's result
    if world == 1 and out is not None and not a.no_cpu_baseline:
        try:
            out["parity"] = parity_block(host, env0, lut, cams[0], dev, a.streams)
        except Exception as exc:
            out["parity"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    # ---- the other BASELINE.json configurations, short device-timed runs (N = 1: configs 2, 4, 5 on one GPU; N > 1: the
    #      strong-scaling form of config 4 and the 64-view batch of config 5 sharded over the ranks)
    if out is not None or world > 1:
        extra = None
        if not a.no_configs:
            try:
                extra = other_configs(a, rank, world, dev)
            except Exception as exc:
                extra = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        if out is not None:
            out["configs"] = extra
    # ---- extra: the e2e of a TRAINING view as the reference's trainer drives the path:
    #      the Gaussians stay on the device, the host supplies a view's ground-truth image (pinned, H2D inside the timed
    #      region) and reads back the batch's loss scalar.  Guarded: it can never cost the headline line.
    if world == 1 and out is not None and not a.no_e2e:
        try:
            from geosplatting_b200.loss import view_loss
            n_b = max(2, min(6, a.steps // B))
            gt_host = [torch.rand(H, W, 4, generator=gen).pin_memory() for _ in range(B)]
            copy_stream = torch.cuda.Stream(dev)

            def train_batch():
                cs = [cams[j % len(cams)] for j in range(B)]
                with torch.cuda.stream(copy_stream):
                    gts = [g.to(dev, non_blocking=True) for g in gt_host]
                done = torch.cuda.Event()
                done.record(copy_stream)
                imgs = splat_views(params["means"], params["scales"], params["quats"], params["opacities"], params["kd"],
                                   params["ks"], params["normals"], cs, exposures=exposure, envmap=env, fg_lut=lut,
                                   min_roughness=0.1, max_metallic=1.0, n_streams=a.streams)
                torch.cuda.current_stream(dev).wait_event(done)
                loss = sum(view_loss(i_, g_) for i_, g_ in zip(imgs, gts)) / B
                torch.autograd.grad(loss, grad_inputs)
                return float(loss)                          # the D2H read of the batch's result

            for _ in range(2):
                train_batch()
            barrier()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            for _ in range(n_b):
                last = train_batch()
            e_.record()
            barrier()
            ms_b = s_.elapsed_time(e_) / n_b
            out["e2e_training_view"] = {
                "value": round(B * world / (ms_b / 1e3), 3), "unit": UNIT, "h2d_bytes_per_step": H * W * 4 * 4,
                "d2h_bytes_per_step": 4.0 / B, "batches_timed": n_b, "loss": round(last, 5),
                "what": "per view: ground-truth image (pinned host) -> device, splat fwd, per-view SSIM/L1/mask loss, "
                        "backward to the per-Gaussian state, env texels and exposure; per batch of 8: the loss scalar "
                        "-> host.  Gaussian state resident on the device, as in the reference's trainer"}
        except Exception as exc:
            out["e2e_training_view"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    return out, rank, world


def parity_block(host, env0, lut, cam, dev, n_streams):
    """One view of the bench workload through the product path (fused.splat_views -> the batch driver) against the
    composed CPU oracle on identical inputs (oracle/parity.py): image L-inf on the non-fragile pixels, fragile fraction,
    identity of the tile lists, and the relative L2 error of every gradient group."""
    import time as _t

    import torch

    from geosplatting_b200.fused import splat_views
    from geosplatting_b200.rasterization import rasterization
    from geosplatting_b200.shade import EnvStack
    from oracle import parity as OP
    t0 = _t.perf_counter()
    g = {k: v.detach().cpu() for k, v in host.items()}
    lv = [x.detach().cpu() for x in env0.level_views()]
    base, mips = lv[-1][..., :3].contiguous(), [x[..., :3].contiguous() for x in lv[:-1]]
    exposure = torch.tensor([1.0])
    cot = torch.randn(cam.height, cam.width, 4, generator=torch.Generator().manual_seed(7))
    o = OP.oracle_view(g, base, mips, lut.detach().cpu(), cam, exposure, cot)
    order = ("means", "scales", "quats", "opacities", "kd", "ks", "normals")
    p = {k: g[k].to(dev).requires_grad_(True) for k in order}
    leaf = env0.data.detach().clone().requires_grad_(True)
    env = EnvStack(leaf, env0.R0, env0.L, env0.Rb, env0.min_roughness, env0.max_roughness)
    ex = exposure.to(dev).requires_grad_(True)
    (img,) = splat_views(*[p[k] for k in order], [cam], exposures=ex, envmap=env, fg_lut=lut, min_roughness=0.1,
                         max_metallic=1.0, n_streams=n_streams)
    gr = torch.autograd.grad(img, [p[k] for k in order] + [leaf, ex], grad_outputs=o["cot"].to(dev))
    lvg = EnvStack(gr[7], env0.R0, env0.L, env0.Rb).level_views()
    grads = {"means": gr[0], "scales": gr[1], "quats": gr[2], "logits": gr[3], "kd": gr[4], "ks": gr[5], "normals": gr[6],
             "base": lvg[-1][..., :3], "mips": [x[..., :3] for x in lvg[:-1]], "exposure": gr[8]}
    with torch.no_grad():
        vm = torch.from_numpy(cam.view_matrix)[None].to(dev)
        K = torch.from_numpy(cam.intrinsic_matrix)[None].to(dev)
        _, _, info = rasterization(p["means"], p["quats"], p["scales"].exp(), torch.sigmoid(p["opacities"])[:, 0],
                                   p["normals"], vm, K, cam.width, cam.height, rasterize_mode="antialiased")
    rep = OP.compare(o, img.detach().cpu().numpy(), info["flatten_ids"].cpu().numpy(), info["isect_offsets"].cpu().numpy(),
                     grads)
    return {"view": "camera 0 of the timed workload, identical Gaussians / env levels on both sides",
            "linf": rep["linf"], "linf_tolerance": 1e-4, "fragile_frac": rep["fragile_frac"],
            "fragile_pixels": rep["fragile_pixels"], "psnr_db_all_pixels": rep["psnr_db_all_pixels"],
            "ids_equal": rep["ids_equal"], "intersections": rep["intersections"],
            "grad_rel_l2": {k: round(v["rel_l2"], 8) for k, v in rep["grads"].items()},
            "oracle_seconds": {k: round(v, 3) for k, v in o["seconds"].items()},
            "seconds": round(_t.perf_counter() - t0, 2)}


def other_configs(a, rank, world, dev):
    """Device-timed views/s of the BASELINE.json configurations the headline does not cover, same protocol (batches of
    views through fused.splat_views + one backward, L2 flushed between batches, CUDA events, max over ranks):
      N = 1 : config 2 (~0.5 M Gaussians, 800^2), config 4 (~2 M, 800^2), config 5 (~5 M, 1600^2), 8 views each;
      N > 1 : config 4 as STRONG scaling (the trainer's 8 views in total, 8 / N per rank, one all-reduce) and config 5 with
              its 64-view batch sharded over the ranks."""
    import torch
    import torch.distributed as dist

    from geosplatting_b200 import scenes, splitsum
    from geosplatting_b200.fused import splat_views
    from geosplatting_b200.mgadapter import MGAdapter, compute_vertex_normals
    from geosplatting_b200.parallel import FlatAllReduce, shard_views
    from geosplatting_b200.shade import EnvStack, synthetic_fg_lut
    lut = synthetic_fg_lut(dev)
    gen = torch.Generator().manual_seed(0)
    cube = torch.exp(torch.randn(6, a.light_res, a.light_res, 3, generator=gen)).clamp_min(1e-2).to(dev)
    with torch.no_grad():
        env0 = splitsum.as_envstack(cube)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    order = ("means", "scales", "quats", "opacities", "kd", "ks", "normals")

    def run(mesh_n, res, global_views, n_batches):
        verts, faces = scenes.cube_sphere(mesh_n)
        with torch.no_grad():
            vd, fd = verts.to(dev), faces.to(dev)
            sp, _ = MGAdapter().make(vd, fd, compute_vertex_normals(vd, fd))
        N = sp.means.shape[0]
        g = torch.Generator().manual_seed(1)
        p = {"means": sp.means, "scales": sp.scales, "quats": sp.quats, "opacities": sp.opacities,
             "kd": (torch.rand(N, 3, generator=g) * 0.8 + 0.1).to(dev), "ks": torch.rand(N, 2, generator=g).to(dev),
             "normals": sp.colors}
        p = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
        leaf = env0.data.detach().clone().requires_grad_(True)
        env = EnvStack(leaf, env0.R0, env0.L, env0.Rb, env0.min_roughness, env0.max_roughness)
        ex = torch.ones(1, device=dev, requires_grad=True)
        cams = shard_views(scenes.orbit_cameras(global_views, res, res, seed=1), rank, world)
        cot = torch.randn(res, res, 4, generator=g).to(dev) / float(global_views)
        inputs = [p[k] for k in order] + [leaf, ex]
        pend = [None]

        def batch():
            imgs = splat_views(*[p[k] for k in order], cams, exposures=ex, envmap=env, fg_lut=lut, min_roughness=0.1,
                               max_metallic=1.0, n_streams=a.streams)
            grads = torch.autograd.grad(imgs, inputs, grad_outputs=[cot] * len(cams))
            if world > 1:
                if pend[0] is not None:
                    pend[0].wait()
                pend[0] = FlatAllReduce(list(grads), async_op=True)

        for _ in range(2):
            batch()
        if pend[0] is not None:
            pend[0].wait()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(n_batches):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            batch()
            if pend[0] is not None:
                pend[0].wait()                 # strong scaling: the step ends when the summed gradients are there
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        t = torch.tensor([ms / n_batches], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_batch = float(t.item())
        del p, leaf, env, sp
        torch.cuda.empty_cache()
        return {"gaussians": N, "resolution": [res, res], "views_global": global_views, "views_per_rank": len(cams),
                "ms_per_batch": round(ms_batch, 3), "views_per_s": round(global_views / (ms_batch / 1e3), 2),
                "batches_timed": n_batches}

    def wide_channels(mesh_n, res):
        """the D = 14 G-buffer splat of the later stages (geosplat.py:276-295) through the drop-in
        `rasterization()` (padded to 16 channels), forward + backward per view, next to D = 3 through the same operator."""
        from geosplatting_b200.rasterization import rasterization
        verts, faces = scenes.cube_sphere(mesh_n)
        with torch.no_grad():
            vd, fd = verts.to(dev), faces.to(dev)
            sp, _ = MGAdapter().make(vd, fd, compute_vertex_normals(vd, fd))
        N = sp.means.shape[0]
        cam = scenes.orbit_cameras(1, res, res, seed=1)[0]
        vm = torch.from_numpy(cam.view_matrix)[None].to(dev)
        K = torch.from_numpy(cam.intrinsic_matrix)[None].to(dev)
        geo = [t.detach().clone().requires_grad_(True) for t in (sp.means, sp.quats)]
        scales, opac = sp.scales.exp().detach(), torch.sigmoid(sp.opacities)[:, 0].detach()
        res_ = {}
        for D in (3, 14):
            feats = torch.rand(N, D, device=dev, requires_grad=True)
            cot = torch.randn(1, res, res, D, device=dev)

            def one():
                r, _, _ = rasterization(geo[0], geo[1], scales, opac, feats, vm, K, res, res, rasterize_mode="antialiased")
                torch.autograd.grad(r, [feats] + geo, grad_outputs=cot)

            for _ in range(3):
                one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                one()
            e1.record()
            torch.cuda.synchronize()
            res_[f"D{D}_ms_per_view"] = round(e0.elapsed_time(e1) / 5, 3)
        res_.update(gaussians=N, resolution=[res, res], what="rasterization() fwd+bwd, one view at a time, staged operator")
        return res_

    out = {}
    if world == 1:
        out["wide_channel_raster_1M_800"] = wide_channels(118, 800)
        out["config2_500k_800"] = run(83, 800, 8, 4)
        out["config4_2M_800"] = run(167, 800, 8, 3)
        out["config5_5M_1600"] = run(264, 1600, 8, 2)
    else:
        r = run(167, 800, 8, 4)
        r["scaling"] = "strong: the trainer's batch of 8 views in total, one all-reduce per batch inside the timed window"
        out["config4_2M_800_strong"] = r
        r = run(264, 1600, 64, 2)
        r["scaling"] = "the 64-view batch of config 5 sharded over the ranks"
        out["config5_5M_1600_64views"] = r
    return out


def train_step_probe(R=140, n=7):
    """BASELINE configs[2] ("1M Gaussians + FlexiCubes MGAdaptor, 800x800, full train step") through the reference-facing
    model: FlexiCubes mesh + regularisers on an R^3 SDF grid -> vertex normals + MGAdaptor -> kd / ks / z hash-grid
    fields (jitter regularisers on) -> split-sum prefilter of a 6 x 512^2 cube map -> 8 views 800 x 800 -> per-view
    SSIM / L1 / mask loss -> backward to all eight parameter groups -> Adam.  CUDA events around each of n steps after 5
    warm-up steps; the MEDIAN is reported with min / max beside it (the mesh, and with it every buffer size, changes from
    step to step: an occasional step pays for allocator growth).  A reported extra, outside the timed region of the headline metric (same code as scripts/bench_train_step.py)."""
    import torch

    from geosplatting_b200 import _lib, scenes
    from geosplatting_b200.model import GeoSplatter
    from geosplatting_b200.shade import synthetic_fg_lut
    dev = torch.device("cuda", torch.cuda.current_device())
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
    stats = {}

    def step():
        opt.zero_grad(set_to_none=True)
        loss, metrics = m.training_loss(cams, gt)
        loss.backward()
        opt.step()
        stats.update(gaussians=int(metrics["#gaussians"]), loss=round(float(metrics["loss"]), 5))

    gc.collect()
    torch.cuda.empty_cache()               # the headline phases left tens of GB cached in other size classes
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    _lib.CallStats.reset()
    per_step = []
    gc.disable()
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step()
        e.record()
        per_step.append((s, e))
    torch.cuda.synchronize()
    gc.enable()
    per_step = sorted(a_.elapsed_time(b_) for a_, b_ in per_step)
    ms = per_step[n // 2]
    launches = _lib.CallStats.launches() // n
    _lib.CallStats.reset()
    return {"what": "whole stage-1 training step of BASELINE config 3: model.GeoSplatter.training_loss (FlexiCubes, fields, "
                    "prefilter, 8 views 800x800, per-view loss) + backward + Adam", "flexicubes_resolution": R, **stats,
            "ms_per_step": round(ms, 3), "statistic": "median", "ms_per_step_min_max": [round(per_step[0], 3), round(per_step[-1], 3)],
            "views_per_s": round(8 / (ms / 1e3), 2), "steps_timed": n, "gpu_launches_per_step": launches, "peak_memory_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}


def stats_last_view(params, cam, W, H):
    """(M, Nv) of one view, read back outside the timed region."""
    import torch

    from geosplatting_b200.rasterization import rasterization
    with torch.no_grad():
        _, _, info = rasterization(params["means"], params["quats"], params["scales"].exp(),
                                   torch.sigmoid(params["opacities"]).squeeze(-1), params["normals"],
                                   torch.from_numpy(cam.view_matrix)[None], torch.from_numpy(cam.intrinsic_matrix)[None],
                                   W, H, rasterize_mode="antialiased")
        return int(info["flatten_gaussian_ids"].shape[0]), int(info["gaussian_ids"].shape[0])


# ------------------------------------------------------------------------------------------------
def run_oracle(a, steps, warmup, budget_s=150.0):
    """fwd+bwd of one view per step with the CPU oracle (torch shade restatement + C rasterizer, all host threads).
    The Gaussians come from the torch MGAdaptor restatement, the env map from torch mip + C prefilter, both OUTSIDE
    the timed region (they are per-step work, the metric is per view)."""
    import numpy as np
    import torch

    from geosplatting_b200 import scenes
    from geosplatting_b200.shade import synthetic_fg_lut
    from oracle import mgadapter as OMG
    from oracle import raster as R
    from oracle import shade as OS

    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: override it explicitly)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    R.set_num_threads(cores)
    cores = R.num_threads()
    torch.set_num_threads(cores)
    sc = build_scene_host(a)
    W = H = a.res
    cams = scenes.orbit_cameras(a.views, W, H, seed=1)
    with torch.no_grad():
        vn = OMG.vertex_normals(sc["verts"], sc["faces"])
        means, ls, quats, normals, opac, _ = OMG.make(sc["verts"], sc["faces"], vn)
    # env levels: the timed unit (a view) only samples them; use the mip chain of the cube map as the stack
    mips = [sc["cubemap"]]
    while mips[-1].shape[1] > 16:
        mips.append(OS.cubemap_mip_fwd(mips[-1]))
    base = mips[-1]
    lut = synthetic_fg_lut("cpu")
    rng = np.random.default_rng(1234)
    cot = torch.from_numpy(rng.normal(size=(H, W, 4)).astype(np.float32))
    scales = ls.exp().numpy()
    op = torch.sigmoid(opac)[:, 0].numpy()

    def step(i):
        c = cams[i % len(cams)]
        m = means.clone().requires_grad_(True)
        n_ = normals.clone().requires_grad_(True)
        kd = sc["kd"].clone().requires_grad_(True)
        ks = sc["ks"].clone().requires_grad_(True)
        col = OS.shade(m, n_, kd, ks, torch.from_numpy(c.position.copy()), lut, base, mips, min_roughness=0.1,
                       max_metallic=1.0, mode="pbr")
        ocam = R.Camera(c.view_matrix, c.fx, c.fy, c.cx, c.cy, c.width, c.height)
        coln = col.detach().numpy()
        render, alpha, info = R.rasterization(means.numpy(), quats.numpy(), scales, op, coln, ocam,
                                              rasterize_mode="antialiased")
        rgba = torch.from_numpy(np.concatenate([render, alpha], -1)).requires_grad_(True)
        ex = torch.ones(1, requires_grad=True)
        img = OS.tone_map_naive(rgba, ex)
        v_rgba, _ = torch.autograd.grad((img * cot).sum(), [rgba, ex])
        g = R.rasterization_bwd(means.numpy(), quats.numpy(), scales, op, coln, ocam, info, alpha,
                                v_rgba[..., :3].contiguous().numpy(), v_rgba[..., 3:].contiguous().numpy(),
                                rasterize_mode="antialiased")
        torch.autograd.grad(col, [m, n_, kd, ks], grad_outputs=torch.from_numpy(g[4]))

    t0 = time.perf_counter()
    step(0)
    t_first = time.perf_counter() - t0
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
            "sample": f"{timed} full view(s) fwd+bwd of the same workload: torch restatement of the shade "
                      f"(autograd backward) + C restatement of the rasterizer (OpenMP), {cores} host threads",
            "ms_per_step": dt / timed * 1e3, "steps_timed": timed}


def main():
    a = parse_args()
    if a.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        cb = run_oracle(a, a.steps, a.warmup)
        out = {"metric": METRIC, "value": round(cb["value"], 5), "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": round(cb["ms_per_step"], 2), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": workload_name(a), "gaussians": n_gaussians(a), "resolution": [a.res, a.res]},
               "impl": "reference", "cpu_baseline": {k: (round(v, 5) if isinstance(v, float) else v) for k, v in cb.items()},
               "e2e": {"value": round(cb["value"], 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out))
        return
    out, rank, world = run_b200(a)
    if rank == 0 and world == 1 and not a.no_train_step:
        try:
            out["train_step"] = train_step_probe()
        except Exception as exc:      # the probe is an extra: it must never cost the headline line
            out["train_step"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
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
