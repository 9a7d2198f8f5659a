"""One autograd node for a whole batch of views of RenderableAttrs.splat (rfstudio/model/geosplat.py:53-132 with
culling=False; the per-view loop of GeoSplatter.render_report, geosplat.py:869-879): activations -> projection ->
intersection count -> split-sum shade -> binning/sort -> compositing -> tone map per view, and the reverse chain of
every view in one backward.

Same C-ABI kernels as the stage-by-stage operators (rasterization.py, shade.py, mgadapter.py) and the same results;
what it removes is host work, glue kernels and idle SMs:
  * one autograd node instead of ~15 per view; no torch.cat / stack / slice copies of the image
    (gsb_tonemap_planar_*); sigmoid(opacity) * compensation folded into the record packing and its chain rule into
    gsb_project_bwd; cached camera structs; exp(log-scales) once per batch;
  * the views of a batch are spread round-robin over CUDA streams and software-pipelined: a view is PREPARED
    (projection, count, shade) one stream-round before it is finished, so the host's wait for its intersection count
    never leaves the device idle, and binning / compositing -- forward and backward -- of neighbouring views overlap:
    the composite kernels end in a long, poorly occupied tail (one warp per sub-list) that other views fill;
  * the backward kernels of the views on one stream ADD into that stream's gradient buffer (accumulate flag of
    gsb_project_bwd / gsb_shade_bwd), so a batch costs one zero-fill and one add per stream instead of one add per
    view and tensor in the autograd engine.
At ~1.2 ms of device time per view the host would otherwise be the bottleneck (scripts/host_overhead.py).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, NamedTuple, Optional, Sequence

import torch
from torch import Tensor

from ._lib import CallStats, GsbCamera, GsbViewConfig, call, f32c, ptr, stream_ptr
from ._lib import require_cuda as _require_cuda
from .rasterization import BinCount, _release_slot, _total_slot, bin_finish, make_camera
from .scenes import PinholeCamera, to_pinhole
from .shade import MODES, EnvStack, get_fg_lut, shade_workspace


class ViewMeta(NamedTuple):
    R0: int
    L: int
    Rb: int
    env_min_roughness: float
    env_max_roughness: float
    min_roughness: float
    max_metallic: float
    mode: int
    antialiased: bool
    naive_tonemap: int


def _camera_struct(camera: PinholeCamera, antialiased: bool):
    """(gsb_camera, cam_pos float[3]) for a camera, cached on the camera object (cameras are immutable here, as
    the reference's `Cameras` tensors are)."""
    cache = camera.__dict__.setdefault("_gsb_cache", {})
    hit = cache.get(antialiased)
    if hit is None:
        cam = make_camera(camera.view_matrix, camera.intrinsic_matrix, camera.width, camera.height, near_plane=0.01,
                          far_plane=1e10, antialiased=antialiased)
        pos = (C.c_float * 3)(*[float(x) for x in camera.c2w[:, 3]])
        hit = cache[antialiased] = (cam, pos)
    return hit


_side_streams: dict = {}


def _streams(dev: torch.device, n: int) -> List[torch.cuda.Stream]:
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    pool = _side_streams.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(dev))
    return pool[:n]


class _Shared:
    """Per-batch inputs in kernel form (contiguous fp32, activated scales)."""
    __slots__ = ("means", "quats", "scales", "logits", "kd", "ks", "normals", "env", "lut", "meta", "N", "dev")


class _View:
    """State of one view between prepare, finish and backward."""
    __slots__ = ("cam", "cam_pos", "exposure", "means2d", "depths", "conics", "comps", "radii", "count", "colors",
                 "offsets", "M", "render", "alphas", "last_ids", "ws", "stream")


def _prepare(sh: _Shared, camera: PinholeCamera, exposure: Tensor) -> _View:
    """Everything of a view that does not need M on the host: projection, intersection count in flight, shade."""
    v = _View()
    dev, N, meta = sh.dev, sh.N, sh.meta
    v.cam, v.cam_pos = _camera_struct(camera, meta.antialiased)
    v.exposure = f32c(exposure).reshape(1)
    st = stream_ptr(dev)
    fbuf = torch.empty(7 * N, dtype=torch.float32, device=dev)
    v.means2d, v.depths = fbuf[:2 * N].view(N, 2), fbuf[2 * N:3 * N]
    v.conics, v.comps = fbuf[3 * N:6 * N].view(N, 3), fbuf[6 * N:]
    ibuf = torch.empty(2 * N, dtype=torch.int32, device=dev)
    v.radii, tpg = ibuf[:N], ibuf[N:]
    call("gsb_project_fwd", dev, C.c_int32(N), ptr(sh.means), ptr(sh.quats), ptr(sh.scales), C.byref(v.cam),
         ptr(v.radii), ptr(v.means2d), ptr(v.depths), ptr(v.conics), ptr(v.comps), ptr(tpg), st)
    v.count = BinCount(tpg, v.depths)                         # M is on its way to the host ...
    v.colors = torch.empty(N, 3, dtype=torch.float32, device=dev)
    call("gsb_shade_fwd", dev, C.c_int32(N), ptr(sh.means), ptr(sh.normals), ptr(sh.kd), ptr(sh.ks), v.cam_pos,
         ptr(sh.lut), C.c_int32(sh.lut.shape[0]), ptr(sh.env), C.c_int32(meta.R0), C.c_int32(meta.L), C.c_int32(meta.Rb),
         C.c_float(meta.min_roughness), C.c_float(meta.max_metallic), C.c_float(meta.env_min_roughness),
         C.c_float(meta.env_max_roughness), C.c_int32(meta.mode), ptr(v.colors), st)
    return v


def _finish(sh: _Shared, v: _View, out: Tensor) -> Tensor:
    """Binning, compositing, tone map of a prepared view (the host waits for M here)."""
    dev, N, meta = sh.dev, sh.N, sh.meta
    W, H = v.cam.width, v.cam.height
    st = stream_ptr(dev)
    flatten_ids, v.offsets = bin_finish(v.count, v.means2d, v.radii, v.cam)   # ... and is awaited only here
    v.M = M = flatten_ids.shape[0]
    v.count = None
    v.render = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
    v.alphas = torch.empty(H, W, dtype=torch.float32, device=dev)
    v.last_ids = torch.empty(H, W, dtype=torch.int32, device=dev)
    nbytes = C.c_size_t(0)
    call("gsb_composite_workspace_bytes", dev, C.c_int64(N), C.c_int64(M), C.c_int32(W), C.c_int32(H), C.byref(nbytes))
    v.ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)          # kept alive for the backward
    call("gsb_composite_fwd", dev, C.c_int32(W), C.c_int32(H), C.c_int32(3), C.c_int64(N), ptr(v.means2d),
         ptr(v.conics), ptr(v.colors), ptr(sh.logits), C.c_int32(1), ptr(v.comps) if meta.antialiased else None, None,
         ptr(v.offsets), ptr(flatten_ids), C.c_int64(M), ptr(v.render), ptr(v.alphas), ptr(v.last_ids), ptr(v.ws),
         C.c_size_t(v.ws.numel()), st)
    call("gsb_tonemap_planar_fwd", dev, C.c_int64(H * W), ptr(v.render), ptr(v.alphas), ptr(v.exposure),
         C.c_int32(meta.naive_tonemap), ptr(out), st)
    v.means2d = v.depths = v.conics = v.comps = None          # only radii / colors / lists are needed again
    return out


# ---- native orchestration: three C-ABI calls per view (csrc/view.cu) -----------------------------------------------
def _u8(n: int, dev) -> Tensor:
    return torch.empty(n, dtype=torch.uint8, device=dev)


def _view_config(sh: _Shared, cam) -> GsbViewConfig:
    m = sh.meta
    return GsbViewConfig(sh.N, cam.width, cam.height, sh.lut.shape[0], m.R0, m.L, m.Rb, m.min_roughness, m.max_metallic,
                         m.env_min_roughness, m.env_max_roughness, m.mode, m.naive_tonemap)


class _NativeView:
    __slots__ = ("cam", "cam_pos", "cfg", "sizes", "exposure", "keep1", "tmp1", "keep2", "slot", "event", "M", "stream")


def _prepare_native(sh: _Shared, camera: PinholeCamera, exposure: Tensor, cfgs: dict) -> _NativeView:
    v = _NativeView()
    dev = sh.dev
    v.cam, v.cam_pos = _camera_struct(camera, sh.meta.antialiased)
    v.exposure = f32c(exposure).reshape(1)
    key = (v.cam.width, v.cam.height)
    hit = cfgs.get(key)
    if hit is None:
        cfg = _view_config(sh, v.cam)
        sizes = (C.c_size_t * 5)()
        call("gsb_view_bytes", dev, C.addressof(cfg), 0, C.addressof(sizes))
        hit = cfgs[key] = (cfg, tuple(sizes))
    v.cfg, v.sizes = hit
    v.keep1, v.tmp1 = _u8(v.sizes[0], dev), _u8(v.sizes[1], dev)
    v.slot = _total_slot(dev)
    stream = torch.cuda.current_stream(dev)
    call("gsb_view_prepare", dev, C.addressof(v.cfg), C.addressof(v.cam), C.addressof(v.cam_pos), sh.means.data_ptr(),
         sh.quats.data_ptr(), sh.scales.data_ptr(), sh.normals.data_ptr(), sh.kd.data_ptr(), sh.ks.data_ptr(),
         sh.lut.data_ptr(), sh.env.data_ptr(), v.keep1.data_ptr(), v.tmp1.data_ptr(), v.slot.data_ptr(),
         stream.cuda_stream)
    v.event = torch.cuda.Event()
    v.event.record(stream)
    return v


def _finish_native(sh: _Shared, v: _NativeView, out: Tensor) -> Tensor:
    dev = sh.dev
    v.event.synchronize()                                     # M was stored straight into pinned memory
    v.M = M = int(v.slot[0])
    _release_slot(v.slot)
    v.event = v.slot = None
    sizes = (C.c_size_t * 5)()
    call("gsb_view_bytes", dev, C.addressof(v.cfg), M, C.addressof(sizes))
    v.keep2 = _u8(sizes[2], dev)
    tmp2 = _u8(sizes[3], dev)
    call("gsb_view_finish", dev, C.addressof(v.cfg), C.addressof(v.cam), M, sh.logits.data_ptr(), v.exposure.data_ptr(),
         v.keep1.data_ptr(), v.tmp1.data_ptr(), v.keep2.data_ptr(), tmp2.data_ptr(), out.data_ptr(),
         torch.cuda.current_stream(dev).cuda_stream)
    v.tmp1 = None
    return out


# bench.py sets this to a list to time the compositing backward of every view live, inside the running batches:
# (start, stop) CUDA-event pairs are appended to it
PROBES = None


def _view_backward_native(sh: _Shared, v: _NativeView, v_out: Tensor, g: "_Grads", v_exp: Tensor) -> None:
    dev = sh.dev
    v_out = f32c(v_out)
    tmp3 = _u8(v.sizes[4], dev)
    p0 = p1 = 0
    if PROBES is not None:
        stream = torch.cuda.current_stream(dev)
        pair = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        for e in pair:
            e.record(stream)          # materialises the cudaEvent_t; the driver re-records it around the stage
        PROBES.append(pair)
        p0, p1 = pair[0].cuda_event, pair[1].cuda_event
    call("gsb_view_backward", dev, C.addressof(v.cfg), C.addressof(v.cam), C.addressof(v.cam_pos), v.M,
         sh.means.data_ptr(), sh.quats.data_ptr(), sh.scales.data_ptr(), sh.logits.data_ptr(), sh.normals.data_ptr(),
         sh.kd.data_ptr(), sh.ks.data_ptr(), sh.lut.data_ptr(), sh.env.data_ptr(), v.exposure.data_ptr(),
         v.keep1.data_ptr(), v.keep2.data_ptr(), tmp3.data_ptr(), v_out.data_ptr(), g.means.data_ptr(),
         g.quats.data_ptr(), g.scales.data_ptr(), g.logits.data_ptr(), g.normals.data_ptr(), g.kd.data_ptr(),
         g.ks.data_ptr(), g.env.data_ptr(), v_exp.data_ptr(), p0 or None, p1 or None,
         torch.cuda.current_stream(dev).cuda_stream)


class _Grads:
    """One stream's gradient buffer: [env 4T | quats 4N | ks 2N | means 3N | scales 3N | logits N | normals 3N | kd 3N]
    (the float4 / float2 consumers first, so that every view is 16-byte aligned)."""

    def __init__(self, N: int, T: int, dev):
        self.flat = torch.zeros(4 * T + 19 * N, dtype=torch.float32, device=dev)
        o = 0

        def take(n):
            nonlocal o
            t = self.flat[o:o + n]
            o += n
            return t

        self.env, self.quats, self.ks = take(4 * T), take(4 * N), take(2 * N)
        self.means, self.scales, self.logits = take(3 * N), take(3 * N), take(N)
        self.normals, self.kd = take(3 * N), take(3 * N)


def _view_backward(sh: _Shared, v: _View, v_out: Tensor, g: _Grads, v_exp: Tensor) -> None:
    dev, N, meta = sh.dev, sh.N, sh.meta
    W, H = v.cam.width, v.cam.height
    st = stream_ptr(dev)
    v_out = f32c(v_out)
    v_render = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
    v_alphas = torch.empty(H, W, dtype=torch.float32, device=dev)
    call("gsb_tonemap_planar_bwd", dev, C.c_int64(H * W), ptr(v.render), ptr(v.exposure), C.c_int32(meta.naive_tonemap),
         ptr(v_out), ptr(v_render), ptr(v_alphas), ptr(v_exp), st)
    acc = torch.zeros(9 * N, dtype=torch.float32, device=dev)           # this view's four atomic accumulators
    v_means2d, v_conics = acc[:2 * N], acc[2 * N:5 * N]
    v_colors, v_opac = acc[5 * N:8 * N], acc[8 * N:]
    call("gsb_composite_bwd", dev, C.c_int32(W), C.c_int32(H), C.c_int32(3), C.c_int64(N), ptr(v.colors), None,
         ptr(v.offsets), C.c_int64(v.M), ptr(v.alphas), ptr(v.last_ids), ptr(v_render), ptr(v_alphas), ptr(v_means2d),
         ptr(v_conics), ptr(v_colors), ptr(v_opac), ptr(v.ws), st)
    call("gsb_project_bwd", dev, C.c_int32(N), ptr(sh.means), ptr(sh.quats), ptr(sh.scales), C.byref(v.cam),
         ptr(v.radii), ptr(v_means2d), None, ptr(v_conics), None, ptr(g.means), ptr(g.quats), ptr(g.scales),
         ptr(sh.logits), ptr(v_opac), ptr(g.logits), C.c_int32(1), st)
    sws = shade_workspace(dev, meta.R0, meta.L, meta.Rb)
    call("gsb_shade_bwd", dev, C.c_int32(N), ptr(sh.means), ptr(sh.normals), ptr(sh.kd), ptr(sh.ks), v.cam_pos,
         ptr(sh.lut), C.c_int32(sh.lut.shape[0]), ptr(sh.env), C.c_int32(meta.R0), C.c_int32(meta.L), C.c_int32(meta.Rb),
         C.c_float(meta.min_roughness), C.c_float(meta.max_metallic), C.c_float(meta.env_min_roughness),
         C.c_float(meta.env_max_roughness), C.c_int32(meta.mode), ptr(v_colors), ptr(g.means), ptr(g.normals),
         ptr(g.kd), ptr(g.ks), ptr(g.env), ptr(sws), C.c_size_t(sws.numel()), C.c_int32(1), st)


# ---- batch driver: ONE C-ABI call per batch each way, nothing waits for a view's intersection count (csrc/view.cu) -----
_caps: dict = {}          # (device, size class of N, W, H) -> capacity of the tile lists (intersections per view)
_pinned_pool: list = []   # recycled pinned int64 buffers the device publishes the raw counts to
CAP_MARGIN = 1.25
CAP_QUANTUM = 1 << 16


def _size_class(N: int) -> int:
    """Scenes within ~12 % of each other share a capacity: a trainer whose mesh (and with it the Gaussian count) changes a
    little every step (FlexiCubes re-extraction, geosplat.py:735-760) must not re-probe -- and wait for -- every batch."""
    return int(math.log(max(N, 1)) / math.log(1.125))


def _round_cap(m: int) -> int:
    return max(CAP_QUANTUM, (int(m * CAP_MARGIN) + CAP_QUANTUM - 1) // CAP_QUANTUM * CAP_QUANTUM)


def _pinned_counts(n: int) -> Tensor:
    for i, t in enumerate(_pinned_pool):
        if t.numel() >= n:
            return _pinned_pool.pop(i)
    return torch.empty(max(n, 64), dtype=torch.int64).pin_memory()


class CapacityExceeded(RuntimeError):
    """A view produced more tile intersections than the batch's arenas were carved for (the farthest ones were
    dropped).  The capacity has been raised: run the step again."""


class _Batch:
    """State of one batch between forward and backward."""
    __slots__ = ("cfg", "cams", "cam_pos", "n", "streams", "stream_ptrs", "cap", "keep", "totals", "event", "ex_all",
                 "ex_stride", "key", "checked")


def _batch_arrays(sh: _Shared, cameras, side):
    n = len(cameras)
    cams = (GsbCamera * n)()
    pos = (C.c_float * (3 * n))()
    for i, c in enumerate(cameras):
        cam, p = _camera_struct(c, sh.meta.antialiased)
        C.memmove(C.addressof(cams) + i * C.sizeof(GsbCamera), C.addressof(cam), C.sizeof(GsbCamera))
        pos[3 * i], pos[3 * i + 1], pos[3 * i + 2] = p[0], p[1], p[2]
    main = torch.cuda.current_stream(sh.dev)
    streams = list(side) if side else [main]
    ptrs = (C.c_void_p * len(streams))(*[st.cuda_stream for st in streams])
    return cams, pos, streams, ptrs, main


def _exposure_array(exposures, dev):
    """(device float array, stride): stride 0 when every view shares one exposure tensor."""
    first = exposures[0]
    if all(e is first for e in exposures) and first.numel() == 1:
        return f32c(first).reshape(1), 0
    return torch.cat([f32c(e).reshape(1) for e in exposures]), 1


def _batch_forward(sh: _Shared, cameras, exposures, side) -> tuple:
    dev, N, n = sh.dev, sh.N, len(cameras)
    b = _Batch()
    b.n = n
    b.cams, b.cam_pos, b.streams, b.stream_ptrs, main = _batch_arrays(sh, cameras, side)
    W, H = b.cams[0].width, b.cams[0].height
    b.cfg = _view_config(sh, b.cams[0])
    b.ex_all, b.ex_stride = _exposure_array(exposures, dev)
    b.key = (dev.index if dev.index is not None else torch.cuda.current_device(), _size_class(N), W, H)
    ns = len(b.streams)
    out = torch.empty(n, H, W, 4, dtype=torch.float32, device=dev)
    probing = b.key not in _caps
    cap = _caps.get(b.key) or _round_cap(4 * N + 4 * ((W + 15) // 16) * ((H + 15) // 16))
    while True:
        sizes = (C.c_size_t * 2)()
        call("gsb_batch_bytes", dev, C.addressof(b.cfg), n, ns, cap, C.addressof(sizes))
        b.keep, scratch = _u8(sizes[0], dev), _u8(sizes[1], dev)
        b.totals = _pinned_counts(n)
        call("gsb_batch_forward", dev, C.addressof(b.cfg), n, C.addressof(b.cams), C.addressof(b.cam_pos),
             sh.means.data_ptr(), sh.quats.data_ptr(), sh.scales.data_ptr(), sh.logits.data_ptr(), sh.normals.data_ptr(),
             sh.kd.data_ptr(), sh.ks.data_ptr(), sh.lut.data_ptr(), sh.env.data_ptr(), b.ex_all.data_ptr(), b.ex_stride,
             b.keep.data_ptr(), scratch.data_ptr(), cap, b.totals.data_ptr(), out.data_ptr(), C.addressof(b.stream_ptrs),
             ns, main.cuda_stream)
        CallStats.counts["batch_view_forward"] = CallStats.counts.get("batch_view_forward", 0) + n
        b.cap = cap
        b.event = torch.cuda.Event()
        b.event.record(main)
        b.checked = False
        if not probing:
            return b, out
        # first batch of this (scene size, resolution): the capacity was a guess -- wait once, size it from the counts
        b.event.synchronize()
        m_max = int(b.totals[:n].max())
        _caps[b.key] = _round_cap(m_max)
        if m_max <= cap:
            b.checked = True
            _pinned_pool.append(b.totals)
            b.totals = None
            return b, out
        cap = _caps[b.key]


def _batch_check(b: _Batch) -> None:
    """Did every view of the forward fit its capacity?  (The forward has long finished when autograd gets here.)"""
    if b.checked:
        return
    b.event.synchronize()
    m_max = int(b.totals[:b.n].max())
    _pinned_pool.append(b.totals)
    b.totals, b.checked = None, True
    if m_max * 1.08 > _caps.get(b.key, 0):
        _caps[b.key] = _round_cap(m_max)                     # less than 8 % of headroom left: the scene grew, follow it
                                                             # (a bump re-sizes the arenas -- keep it rare)
    if m_max > b.cap:
        raise CapacityExceeded(f"a view of this batch has {m_max} tile intersections, the batch was carved for {b.cap}; "
                               "the capacity has been raised -- run the step again")


def _batch_backward(sh: _Shared, b: _Batch, v_outs, T: int):
    dev, N, n = sh.dev, sh.N, b.n
    main = torch.cuda.current_stream(dev)
    ns = len(b.streams)
    nf = C.c_int64(0)
    call("gsb_batch_grad_floats", dev, C.addressof(b.cfg), n, T, C.byref(nf))
    used = min(ns, n)
    bufs = torch.zeros(used, nf.value, dtype=torch.float32, device=dev)
    sizes = (C.c_size_t * 2)()
    call("gsb_batch_bytes", dev, C.addressof(b.cfg), n, ns, b.cap, C.addressof(sizes))
    scratch = _u8(sizes[1], dev)
    keepalive = [None if v is None else f32c(v) for v in v_outs]
    vptrs = (C.c_void_p * n)(*[None if v is None else v.data_ptr() for v in keepalive])
    gptrs = (C.c_void_p * ns)(*[bufs[min(i, used - 1)].data_ptr() for i in range(ns)])
    probes = None
    if PROBES is not None:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * n)]
        for e in evs:
            e.record(main)            # materialises the cudaEvent_t; the driver re-records it around the stage
        probes = (C.c_void_p * (2 * n))(*[e.cuda_event for e in evs])
        PROBES.extend((evs[2 * i], evs[2 * i + 1]) for i in range(n) if keepalive[i] is not None)
    call("gsb_batch_backward", dev, C.addressof(b.cfg), n, C.addressof(b.cams), C.addressof(b.cam_pos),
         sh.means.data_ptr(), sh.quats.data_ptr(), sh.scales.data_ptr(), sh.logits.data_ptr(), sh.normals.data_ptr(),
         sh.kd.data_ptr(), sh.ks.data_ptr(), sh.lut.data_ptr(), sh.env.data_ptr(), b.ex_all.data_ptr(), b.ex_stride,
         b.keep.data_ptr(), scratch.data_ptr(), b.cap, C.addressof(vptrs), T, C.addressof(gptrs), C.c_float(1.0),
         C.addressof(b.stream_ptrs), ns, None if probes is None else C.addressof(probes), main.cuda_stream)
    CallStats.counts["batch_view_backward"] = CallStats.counts.get("batch_view_backward", 0) + sum(
        v is not None for v in keepalive)
    # the kernels are memory-safe under an overflow (they read min(M, capacity)); checking AFTER the backward has been
    # enqueued keeps the device busy while the host looks at the forward's counts -- and no gradient leaves if it failed
    _batch_check(b)
    flat = bufs[0]
    g = _Grads.__new__(_Grads)
    g.flat = flat
    o = 0
    for name, k in (("env", 4 * T), ("quats", 4 * N), ("ks", 2 * N), ("means", 3 * N), ("scales", 3 * N),
                    ("logits", N), ("normals", 3 * N), ("kd", 3 * N)):
        setattr(g, name, flat[o:o + k])
        o += k
    return g, flat[o:o + n + 1]       # per-view exposure gradients + the spare slot


class _SplatBatch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, log_scales, quats, logits, kd, ks, normals, env_data, lut, meta, cameras, n_streams,
                native, *exposures):
        sh = _Shared()
        sh.means, sh.quats, sh.kd, sh.ks, sh.normals = f32c(means), f32c(quats), f32c(kd), f32c(ks), f32c(normals)
        sh.env, sh.lut, sh.meta = f32c(env_data), lut, meta
        sh.dev = dev = sh.means.device
        sh.N = N = sh.means.shape[0]
        sh.logits = f32c(logits).reshape(N)
        sh.scales = f32c(log_scales).exp()                    # rfstudio/model/gsplat.py:337, once per batch
        main = torch.cuda.current_stream(dev)
        side = _streams(dev, n_streams) if n_streams > 1 and len(cameras) > 1 else []
        n = len(cameras)
        ctx.batch = None
        if native == "batch" and len({(c.width, c.height) for c in cameras}) == 1:
            ctx.batch, out = _batch_forward(sh, cameras, exposures, side)
            ctx.sh, ctx.views, ctx.side, ctx.native = sh, None, side, native
            ctx.save_for_backward(sh.means, sh.quats, sh.logits, sh.kd, sh.ks, sh.normals, sh.env)
            ctx.shapes = (tuple(logits.shape), [tuple(e.shape) for e in exposures], env_data.shape[0])
            return tuple(out[i] for i in range(n))
        # what outlives the side streams' work (the images here, the gradient buffers in the backward) is allocated
        # on the CALLER's stream before the side streams wait for it: the caching allocator then recycles it in plain
        # stream order, with no record_stream() events that would delay reuse and grow the pool
        outs = [torch.empty(c.height, c.width, 4, dtype=torch.float32, device=dev) for c in cameras]
        for s in side:
            s.wait_stream(main)
        where = [side[i % len(side)] if side else main for i in range(n)]
        # Software pipeline over the views: one view per stream is prepared up front; after view i is finished (the
        # only point where the host waits, for M_i), view i + n_streams is prepared on the same stream.  The host never
        # waits on an idle device, and at any time the streams are in different phases, so the big compositing
        # kernels of one view share the SMs with the small sort / scan kernels of its neighbours.
        views, cfgs = [None] * n, {}

        def prep(i):
            with torch.cuda.stream(where[i]):
                v = _prepare_native(sh, cameras[i], exposures[i], cfgs) if native else _prepare(sh, cameras[i], exposures[i])
                v.stream = where[i]
                views[i] = v

        ahead = max(2, len(side))
        for i in range(min(ahead, n)):
            prep(i)
        for i in range(n):
            with torch.cuda.stream(where[i]):
                (_finish_native if native else _finish)(sh, views[i], outs[i])
            if i + ahead < n:
                prep(i + ahead)
        for s in side:
            main.wait_stream(s)
        ctx.sh, ctx.views, ctx.side, ctx.native = sh, views, side, native
        # the kernels of the backward read the inputs again: registering them lets autograd's version counters catch
        # an in-place update (an optimiser step) between this forward and its backward
        ctx.save_for_backward(sh.means, sh.quats, sh.logits, sh.kd, sh.ks, sh.normals, sh.env)
        ctx.shapes = (tuple(logits.shape), [tuple(e.shape) for e in exposures], env_data.shape[0])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *v_outs):
        sh, views, side = ctx.sh, ctx.views, ctx.side
        _ = ctx.saved_tensors                                  # raises if an input was modified in place since the forward
        logits_shape, exposure_shapes, T = ctx.shapes
        dev, N = sh.dev, sh.N
        if ctx.batch is not None:
            if all(v is None for v in v_outs):
                return (None,) * (13 + len(v_outs))
            total, v_exp = _batch_backward(sh, ctx.batch, v_outs, T)
            total.scales.mul_(sh.scales.reshape(-1))           # d exp(s) / d s
            n = len(v_outs)
            if ctx.batch.ex_stride == 0:
                # one exposure tensor shared by every view: its gradient is the sum over the views, kept INSIDE the flat
                # buffer (spare slot) so that a data-parallel caller all-reduces one contiguous buffer and nothing else
                torch.sum(v_exp[:n], dim=0, keepdim=True, out=v_exp[n:n + 1])
                ex_grads = [v_exp[n:n + 1].reshape(exposure_shapes[0])] + [None] * (n - 1)
            else:
                ex_grads = [None if v is None else v_exp[i].reshape(shp)
                            for i, (v, shp) in enumerate(zip(v_outs, exposure_shapes))]
            return (total.means.view(N, 3), total.scales.view(N, 3), total.quats.view(N, 4),
                    total.logits.view(logits_shape), total.kd.view(N, 3), total.ks.view(N, 2), total.normals.view(N, 3),
                    total.env.view(T, 4), None, None, None, None, None, *ex_grads)
        main = torch.cuda.current_stream(dev)
        v_exps = []
        # one zero-filled gradient buffer per stream in use, every view on that stream adds into it (allocated on the
        # caller's stream, see forward)
        bufs = {v.stream: None for v, v_out in zip(views, v_outs) if v_out is not None}
        for k in bufs:
            bufs[k] = _Grads(N, T, dev)
        v_exp_all = torch.zeros(len(views), dtype=torch.float32, device=dev)
        for s in side:
            s.wait_stream(main)
        for i, (v, v_out) in enumerate(zip(views, v_outs)):
            if v_out is None:
                v_exps.append(None)
                continue
            with torch.cuda.stream(v.stream):
                g = bufs[v.stream]
                v_exp = v_exp_all[i:i + 1]
                (_view_backward_native if ctx.native else _view_backward)(sh, v, v_out, g, v_exp)
                v_exps.append(v_exp)
        for s in side:
            main.wait_stream(s)
        if not bufs:
            return (None,) * (13 + len(views))
        gs = list(bufs.values())
        total = gs[0]
        for g in gs[1:]:
            total.flat.add_(g.flat)                            # one add per extra stream for the whole batch
        total.scales.mul_(sh.scales.reshape(-1))               # d exp(s) / d s
        return (total.means.view(N, 3), total.scales.view(N, 3), total.quats.view(N, 4), total.logits.view(logits_shape),
                total.kd.view(N, 3), total.ks.view(N, 2), total.normals.view(N, 3), total.env.view(T, 4), None, None,
                None, None, None, *[None if e is None else e.reshape(shp) for e, shp in zip(v_exps, exposure_shapes)])


def _meta_and_lut(envmap: EnvStack, fg_lut: Tensor, min_roughness, max_metallic, mode, tone_type, rasterize_mode):
    if mode not in MODES:
        raise ValueError(mode)
    if tone_type not in ("naive", "none"):
        raise ValueError(tone_type)
    if rasterize_mode not in ("antialiased", "classic"):
        raise ValueError(f"Unknown rasterize_mode: {rasterize_mode}")
    lut = f32c(fg_lut).reshape(fg_lut.shape[-3], fg_lut.shape[-2], 2)
    meta = ViewMeta(envmap.R0, envmap.L, envmap.Rb, envmap.min_roughness, envmap.max_roughness, float(min_roughness),
                    float(max_metallic), MODES[mode], rasterize_mode == "antialiased", int(tone_type == "naive"))
    return meta, lut


def splat_views(means: Tensor, log_scales: Tensor, quats: Tensor, opacity_logits: Tensor, kd: Tensor, ks: Tensor,
                normals: Tensor, cameras, *, exposures, envmap, fg_lut: Optional[Tensor] = None,
                min_roughness: float, max_metallic: float, mode: str = "pbr", tone_type: str = "naive",
                rasterize_mode: str = "antialiased", n_streams: int = 4, native="batch") -> List[Tensor]:
    """The per-view loop of GeoSplatter.render_report for a batch of cameras: list of [H,W,4] tone-mapped RGBA images,
    ready on the caller's stream, differentiable w.r.t. every tensor argument and `envmap.data`.

    `exposures`: one tensor shared by all views or a sequence of one per view.  `n_streams` <= 1 keeps everything on
    the caller's stream.  `native`: "batch" (default) = the whole batch is ONE C-ABI call each way (gsb_batch_forward /
    gsb_batch_backward, csrc/view.cu): nothing on the host waits for a view's intersection count, the tile lists are
    carved for a capacity learned from the first batch (+25 %) and an overflow raises CapacityExceeded in the backward;
    True = three C-ABI calls per view with exact sizes (the host waits for each count, views software-pipelined over the
    streams); False = every kernel group called from Python (what per-kernel instrumentation needs)."""
    _require_cuda(means, "splat_views")
    envmap = EnvStack.coerce(envmap)                 # the reference's TextureSplitSum is accepted as is
    if fg_lut is None:
        fg_lut = get_fg_lut(256, means.device)      # `_get_fg_lut(256, device)`: the reference's asset
    meta, lut = _meta_and_lut(envmap, fg_lut, min_roughness, max_metallic, mode, tone_type, rasterize_mode)
    cameras = to_pinhole(cameras)                    # PinholeCameras or the reference's `Cameras[B]`
    ex = [exposures] * len(cameras) if isinstance(exposures, Tensor) else list(exposures)
    assert len(ex) == len(cameras)
    if not cameras:
        return []
    return list(_SplatBatch.apply(means, log_scales, quats, opacity_logits, kd, ks, normals, envmap.data, lut, meta,
                                  cameras, int(n_streams), native if native == "batch" else bool(native), *ex))


def splat_view(means: Tensor, log_scales: Tensor, quats: Tensor, opacity_logits: Tensor, kd: Tensor, ks: Tensor,
               normals: Tensor, camera, *, exposure: Tensor, envmap, fg_lut: Optional[Tensor] = None,
               min_roughness: float, max_metallic: float, mode: str = "pbr", tone_type: str = "naive",
               rasterize_mode: str = "antialiased", native="batch") -> Tensor:
    """[H,W,4] tone-mapped RGBA of one view on the caller's stream (a batch of one)."""
    return splat_views(means, log_scales, quats, opacity_logits, kd, ks, normals, [camera], exposures=exposure,
                       envmap=envmap, fg_lut=fg_lut, min_roughness=min_roughness, max_metallic=max_metallic, mode=mode,
                       tone_type=tone_type, rasterize_mode=rasterize_mode, n_streams=1, native=native)[0]
