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

_LAUNCH = re.compile(r"(\w+(?:<[\w, ]+>)?)<<<(.+?), (\d+), 0, (?:\(cudaStream_t\)stream|st)>>>\(")


def build(*names: str) -> C.CDLL:
    """`names`.cu (compiled together) -> tests/emu/_build/lib<names>_emu.so (rebuilt when a source is newer)."""
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(CSRC, n + ".cu") for n in names]
    lib = os.path.join(OUT, "lib" + "_".join(names) + "_emu.so")
    deps = srcs + [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith(".cuh")] + [__file__] + \
        [os.path.join(d, f) for d, _, fs in os.walk(HERE) if "_build" not in d for f in fs if f.endswith((".h", ".cuh"))]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        cpps = []
        for n, src in zip(names, srcs):
            text = open(src).read()
            text, k = _LAUNCH.subn(r"gsb_emu::launch(\2, \3, [](auto... a) { \1(a...); })(", text)
            assert k > 0 and "<<<" not in text, f"{n}.cu: a launch site the emulation rewrite does not understand"
            cpps.append(os.path.join(OUT, n + "_emu.cpp"))
            with open(cpps[-1], "w") as f:
                f.write(text)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", HERE, "-I", CSRC,
                               "-o", lib, *cpps, "-x", "c++", os.path.join(CSRC, "error.cu")])
    so = C.CDLL(lib)
    so.gsb_last_error.restype = C.c_char_p
    return so
