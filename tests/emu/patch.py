"""TEST INFRASTRUCTURE ONLY: route a host module of geosplatting_b200 (its `ptr`, `stream_ptr`, `_require_cuda`) and
`_lib.load` to a host-compiled kernel library for the duration of one test (pytest's monkeypatch undoes it)."""
import ctypes as C

import torch

from geosplatting_b200 import _lib


def host_ptr(t):
    if t is None:
        return None
    assert t.is_contiguous() and t.device.type == "cpu"
    return C.c_void_p(t.data_ptr())


def _poisoned_empty(real_empty):
    """torch.empty that hands out NaN / -1 / 0xFF instead of whatever the allocator held (usually fresh zero pages on
    the CPU): a kernel that reads a buffer the host module allocated with torch.empty before writing it shows up."""
    def empty(*args, **kwargs):
        t = real_empty(*args, **kwargs)
        if t.numel():
            if t.is_floating_point():
                t.fill_(float("nan"))
            elif t.dtype == torch.bool:
                t.fill_(True)
            else:
                t.fill_(255 if t.dtype == torch.uint8 else -1)
        return t
    return empty


def route(monkeypatch, so, *modules):
    monkeypatch.setattr(_lib, "load", lambda: so)
    monkeypatch.setattr(torch, "empty", _poisoned_empty(torch.empty))
    for m in modules:
        monkeypatch.setattr(m, "ptr", host_ptr)
        monkeypatch.setattr(m, "stream_ptr", lambda dev: None)
        monkeypatch.setattr(m, "_require_cuda", lambda t, what: None)
    monkeypatch.setattr(_lib.CallStats, "counts", dict(_lib.CallStats.counts))
