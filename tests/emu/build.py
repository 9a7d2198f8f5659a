"""TEST INFRASTRUCTURE ONLY: compile a streaming kernel file of geosplatting_b200/csrc for the HOST (g++ over
tests/emu/cuda_runtime.h) so that CPU tests can run the real kernel source against the golden fixtures.  See the header
of tests/emu/cuda_runtime.h for what this is and is not; the product never loads the result."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "geosplatting_b200", "csrc")
OUT = os.path.join(HERE, "_build")

_LAUNCH = re.compile(r"(\w+(?:<\w+>)?)<<<(.+?), (\d+), 0, \(cudaStream_t\)stream>>>\(")


def build(name: str) -> C.CDLL:
    """`name`.cu -> tests/emu/_build/lib`name`_emu.so (rebuilt when the source is newer)."""
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(CSRC, name + ".cu")
    cpp = os.path.join(OUT, name + "_emu.cpp")
    lib = os.path.join(OUT, f"lib{name}_emu.so")
    deps = [src, os.path.join(CSRC, "gsb_common.cuh"), os.path.join(HERE, "cuda_runtime.h"), __file__]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        text = open(src).read()
        text, n = _LAUNCH.subn(r"gsb_emu::launch(\2, \3, [](auto... a) { \1(a...); })(", text)
        assert n > 0 and "<<<" not in text, "a launch site the emulation rewrite does not understand"
        with open(cpp, "w") as f:
            f.write(text)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", HERE, "-I", CSRC,
                               "-o", lib, cpp, "-x", "c++", os.path.join(CSRC, "error.cu")])
    so = C.CDLL(lib)
    so.gsb_last_error.restype = C.c_char_p
    return so
