"""TEST INFRASTRUCTURE ONLY: route a host module of geosplatting_b200 (its `ptr`, `stream_ptr`, `_require_cuda`) and
`_lib.load` to a host-compiled kernel library for the duration of one test (pytest's monkeypatch undoes it)."""
import ctypes as C

from geosplatting_b200 import _lib


def host_ptr(t):
    if t is None:
        return None
    assert t.is_contiguous() and t.device.type == "cpu"
    return C.c_void_p(t.data_ptr())


def route(monkeypatch, so, *modules):
    monkeypatch.setattr(_lib, "load", lambda: so)
    for m in modules:
        monkeypatch.setattr(m, "ptr", host_ptr)
        monkeypatch.setattr(m, "stream_ptr", lambda dev: None)
        monkeypatch.setattr(m, "_require_cuda", lambda t, what: None)
    monkeypatch.setattr(_lib.CallStats, "counts", dict(_lib.CallStats.counts))
