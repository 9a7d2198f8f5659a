"""ctypes binding of libgeosplat_b200.so (the C ABI declared in include/geosplat_b200.h).

There is NO fallback: if the CUDA library is missing or a call fails, this raises.  PyTorch is only the
owner of device memory and streams here; every pointer handed to the library is a `tensor.data_ptr()`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSB_LIB_PATH") or os.path.join(_HERE, "lib", "libgeosplat_b200.so")   # override: tuning builds
CSRC_DIR = os.path.join(_HERE, "csrc")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "geosplat_b200.h")

_lib: Optional[C.CDLL] = None


class GsbViewConfig(C.Structure):
    """Mirror of `struct gsb_view_config` (include/geosplat_b200.h)."""

    _fields_ = [
        ("N", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("lut_res", C.c_int32), ("R0", C.c_int32), ("L", C.c_int32), ("Rb", C.c_int32),
        ("min_roughness", C.c_float), ("max_metallic", C.c_float),
        ("env_min_roughness", C.c_float), ("env_max_roughness", C.c_float),
        ("mode", C.c_int32), ("naive_tonemap", C.c_int32),
    ]


class GsbCamera(C.Structure):
    """Mirror of `struct gsb_camera` (include/geosplat_b200.h)."""

    _fields_ = [
        ("viewmat", C.c_float * 16),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("width", C.c_int32), ("height", C.c_int32),
        ("near_plane", C.c_float), ("far_plane", C.c_float),
        ("eps2d", C.c_float), ("radius_clip", C.c_float),
        ("antialiased", C.c_int32), ("camera_id", C.c_int32),
    ]


def build(verbose: bool = False) -> str:
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC_DIR, "-j8"], stdout=out)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"geosplatting_b200: {LIB_PATH} is missing. Build it with `python -c 'import "
                "__graft_entry__ as g; g.build()'` (or `make -C geosplatting_b200/csrc`). "
                "There is no CPU or PyTorch fallback for this path.")
        _lib = declare(C.CDLL(LIB_PATH))
    return _lib


def declare(lib: C.CDLL) -> C.CDLL:
    """Prototypes of the entry points that are called with raw integers (tensor.data_ptr()) rather than ctypes objects:
    the pointer widths have to be declared."""
    lib.gsb_last_error.restype = C.c_char_p
    vp, i64 = C.c_void_p, C.c_int64
    lib.gsb_view_bytes.argtypes = [vp, i64, vp]
    lib.gsb_view_prepare.argtypes = [vp] * 15
    lib.gsb_view_finish.argtypes = [vp, vp, i64] + [vp] * 8
    lib.gsb_view_backward.argtypes = [vp, vp, vp, i64] + [vp] * 26
    i32 = C.c_int32
    lib.gsb_batch_bytes.argtypes = [vp, i32, i32, i64, vp]
    lib.gsb_batch_grad_floats.argtypes = [vp, i32, i64, vp]
    lib.gsb_batch_forward.argtypes = [vp, i32, vp, vp] + [vp] * 10 + [i32, vp, vp, i64, vp, vp, vp, i32, vp]
    lib.gsb_batch_backward.argtypes = ([vp, i32, vp, vp] + [vp] * 10 + [i32, vp, vp, i64, vp, i64, vp, C.c_float, vp,
                                       i32, vp, vp])
    return lib


class CallStats:
    """Per-entry-point launch counters and (optionally) CUDA-event timings on the launching stream.

    `bench.py` turns `timing` on to get the live duration of each kernel group inside the timed region;
    the counters are always on (they are what `gpu_launches` reports).  KERNELS maps a C-ABI entry point
    to the number of kernels written in this repository that one call launches (cub's internal kernels
    for scan / radix sort are not counted as ours).
    """

    KERNELS = {"gsb_isect_scan": 0, "gsb_sort_pairs": 0, "gsb_bin_workspace_bytes": 0, "gsb_envstack_texels": 0,
               "gsb_composite_workspace_bytes": 0, "gsb_specular_workspace_bytes": 0, "gsb_specular_cubemap_fwd": 3,
               "gsb_specular_cubemap_bwd": 3, "gsb_specular_plan_count": 2, "gsb_specular_plan_fill": 2,
               "gsb_specular_plan_fwd": 3, "gsb_specular_plan_bwd": 3, "gsb_composite_fwd": 4, "gsb_composite_bwd": 2, "gsb_shade_bwd": 2,
               "gsb_shade_workspace_bytes": 0, "gsb_vertex_normals_fwd": 2, "gsb_vertex_normals_bwd": 2,
               "gsb_bin2_workspace_bytes": 0, "gsb_bin2_count": 2, "gsb_bin2_sort": 2,
               # the native per-view driver: prepare = project + iota + total + shade; finish = emit + offsets + pack +
               # build + order + composite + tone map; backward = tone map + order + composite + project + shade + sum
               "gsb_view_bytes": 0, "gsb_view_prepare": 4, "gsb_view_finish": 7, "gsb_view_backward": 6,
               # the batch driver: per view forward = prepare (project, iota, total, publish, shade) + finish (emit,
               # offsets, pack, build, order, composite, tone map); per view backward = 6; + 1 gradient sum per batch
               "gsb_batch_bytes": 0, "gsb_batch_grad_floats": 0, "gsb_batch_forward": 0, "gsb_batch_backward": 1,
               "batch_view_forward": 12, "batch_view_backward": 6,
               # FlexiCubes: surface = classify (+ cub select); topology = resolve, class flags, numbering, edge keys,
               # edge flags, edge assignment (+ cub sort and two scans)
               "gsb_fc_workspace_bytes": 0, "gsb_fc_surface": 1, "gsb_fc_topology": 6, "gsb_fc_entropy_fwd": 2}
    timing = False        # False, True (every entry point) or a set of entry-point names
    counts: dict = {}
    events: dict = {}
    host_s: dict = {}     # seconds the host spent INSIDE each entry point (enqueueing; none of them waits for the device)

    @classmethod
    def reset(cls, timing=False) -> None:
        cls.timing = timing
        cls.counts = {}
        cls.events = {}
        cls.host_s = {}

    @classmethod
    def launches(cls) -> int:
        return sum(n * cls.KERNELS.get(k, 1) for k, n in cls.counts.items())

    @classmethod
    def durations_ms(cls) -> dict:
        """name -> (calls, total_ms); synchronises."""
        torch.cuda.synchronize()
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in cls.events.items()}


def call(name: str, dev: torch.device, *args) -> None:
    """Invoke C-ABI entry point `name` on `dev`'s current stream; raise on a non-zero return code."""
    fn = getattr(load(), name)
    CallStats.counts[name] = CallStats.counts.get(name, 0) + 1
    # tensors on a GPU that is not the current one: launch with that device current (streams and kernels are per device)
    if dev is not None and dev.type == "cuda" and dev.index is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):
            return call(name, dev, *args)
    if CallStats.timing is True or (CallStats.timing and name in CallStats.timing):
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.current_stream(dev))
        rc = fn(*args)
        b.record(torch.cuda.current_stream(dev))
        CallStats.events.setdefault(name, []).append((a, b))
    else:
        t0 = time.perf_counter()
        rc = fn(*args)
        CallStats.host_s[name] = CallStats.host_s.get(name, 0.0) + (time.perf_counter() - t0)
    check(rc, name)


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gsb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t: Optional[torch.Tensor]):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("geosplatting_b200: expected a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("geosplatting_b200: expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


def require_cuda(t: torch.Tensor, what: str) -> None:
    """The product has no CPU path: every public operator refuses CPU tensors through this one check (modules bind it
    as `_require_cuda`, which is also the single point tests/emu patches to run kernel source compiled for the host)."""
    if not t.is_cuda:
        raise RuntimeError(f"geosplatting_b200.{what} needs CUDA tensors; there is no CPU path")


def stream_ptr(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def f32c(t: torch.Tensor) -> torch.Tensor:
    """Contiguous fp32 view/copy (detached)."""
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
