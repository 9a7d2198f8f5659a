"""No-GPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and the product package never touches the oracle."""
import ctypes
import os
import re

import pytest

from geosplatting_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(_lib.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.gsb_version.restype = ctypes.c_int
    assert lib.gsb_version() >= 100


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    rc = lib.gsb_bin_workspace_bytes(ctypes.c_int32(-1), ctypes.c_int64(0), None)
    assert rc == -1
    assert b"invalid argument" in lib.gsb_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "geosplatting_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "liboracle" in src \
                        or re.search(r'#include\s+".*oracle', src):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_cpu_tensors_are_rejected():
    import torch
    from geosplatting_b200 import rasterization
    z = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        rasterization(z, torch.zeros(4, 4), z, torch.zeros(4), z, torch.eye(4)[None], torch.eye(3)[None], 32, 32)
