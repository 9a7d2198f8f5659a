"""TEST INFRASTRUCTURE ONLY: compile a streaming kernel file of geosplatting_b200/csrc for the HOST (g++ over
tests/emu/cuda_runtime.h) so that CPU tests can run the real kernel source against the golden fixtures.  See the header
of tests/emu/cuda_runtime.h for what this is and is not; the product never loads the result."""
from __future__ import annotations

import ctypes as C
import os
import re
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "geosplatting_b200", "csrc")
OUT = os.path.join(HERE, "_build")

_DYN_SMEM = re.compile(r"extern __shared__ (\w+) (\w+)\[\];")


def _split_top_level(text: str):
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def _rewrite(text: str, name: str) -> str:
    """`kernel<...><<<grid, block, smem, stream>>>(` -> `gsb_emu::launch(grid, block, smem, [](auto... a) { kernel<...>(a...); })(`
    and `extern __shared__ T x[];` -> a pointer into the emulator's dynamic shared memory."""
    out, pos = "", 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            break
        j = text.index(">>>", i)
        assert text[j + 3] == "(", f"{name}.cu: launch without an argument list"
        # kernel name (with optional template arguments) right before <<<
        k = i
        if text[k - 1] == ">":
            depth = 0
            while True:
                k -= 1
                depth += {">": 1, "<": -1}.get(text[k], 0)
                if depth == 0:
                    break
        while text[k - 1].isalnum() or text[k - 1] == "_":
            k -= 1
        kernel = text[k:i]
        cfg = _split_top_level(text[i + 3:j])
        assert len(cfg) == 4, f"{name}.cu: launch configuration {cfg}"
        out += text[pos:k] + f"gsb_emu::launch({cfg[0]}, {cfg[1]}, {cfg[2]}, [](auto... a) {{ {kernel}(a...); }})"
        pos = j + 3
    out += text[pos:]
    return _DYN_SMEM.sub(r"\1 *\2 = static_cast<\1 *>(gsb_emu::dyn_smem());", out)


def all_kernel_files():
    """Every .cu of the library (error.cu is always linked in)."""
    return sorted(f[:-3] for f in os.listdir(CSRC) if f.endswith(".cu") and f != "error.cu")


def build(*names: str, simt: bool = False, defines: tuple = ()) -> C.CDLL:
    """`names`.cu (compiled together) -> tests/emu/_build/lib<names>_emu[_simt].so (rebuilt when a source is newer).
    simt: run the threads of a block as fibers with real barriers, warp collectives and shared memory (simt.h).
    defines: extra -D macros (the kernels' tuning constants, e.g. "GSB_FG=4").
    GSB_EMU_SANITIZE=1 in the environment builds the libraries with AddressSanitizer + UBSan (the process
    must run with libasan preloaded: tests/test_kernels_asan_cpu.py does that in a subprocess)."""
    if shutil.which("g++") is None:      # the image has it; a box without a host compiler skips these tests
        import pytest
        pytest.skip("g++ not available: the kernel source cannot be compiled for the host")
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE if n == "selftest" else CSRC, n + ".cu") for n in names]     # selftest.cu lives here
    sanitize = bool(os.environ.get("GSB_EMU_SANITIZE"))      # simt.h tells ASan about its fiber switches
    tag = ("_simt" if simt else "") + ("_asan" if sanitize else "") + "".join("_" + d.replace("=", "") for d in defines)
    stem = "_".join(names) if len(names) <= 3 else f"all{len(names)}"
    lib = os.path.join(OUT, f"lib{stem}_emu{tag}.so")
    deps = srcs + [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith(".cuh")] + [__file__] + \
        [os.path.join(d, f) for d, _, fs in os.walk(HERE) if "_build" not in d for f in fs if f.endswith((".h", ".cuh"))]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        cpps = []
        for n, src in zip(names, srcs):
            cpps.append(os.path.join(OUT, f"{n}_emu{tag}.cpp"))
            with open(cpps[-1], "w") as f:
                f.write(_rewrite(open(src).read(), n))
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", HERE, "-I", CSRC,
                               *(["-DGSB_EMU_SIMT"] if simt else []), *[f"-D{d}" for d in defines],
                               *(["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-g"] if sanitize else []),
                               "-o", lib, *cpps, "-x", "c++",
                               os.path.join(CSRC, "error.cu")])
    so = C.CDLL(lib)
    so.gsb_last_error.restype = C.c_char_p
    return so
