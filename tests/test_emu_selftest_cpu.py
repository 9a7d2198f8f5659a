"""The SIMT emulation itself (tests/emu/simt.h) on kernels with known answers: shuffles, votes, disjoint half-warp
collectives running at the same time, a block reduction where threads exit before the barrier, dynamic shared memory on a
2-D grid, a write / barrier / read-neighbour pattern -- in the default, reversed and random resume orders.  A green
parity test over the product kernels must not be an artefact of the emulator."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _run_all():
    from tests.emu import build as emu
    lib = emu.build("selftest", simt=True)
    rng = np.random.default_rng(0)
    # warp scan
    n = 1000
    x = rng.integers(-5, 9, n).astype(np.int32)
    out = np.zeros(n, np.int32)
    lib.emu_warp_scan(C.c_int(n), _p(x), _p(out), None)
    ref = np.concatenate([np.cumsum(x[i:i + 32]) for i in range(0, n, 32)]).astype(np.int32)
    assert np.array_equal(out, ref)
    # votes and xor shuffle
    n = 96 * 5
    x = rng.integers(-3, 50, n).astype(np.int32)
    ballot, any_odd = np.zeros(n, np.uint32), np.zeros(n, np.int32)
    all_pos, partner = np.zeros(n, np.int32), np.zeros(n, np.int32)
    lib.emu_votes(C.c_int(n), _p(x), _p(ballot), _p(any_odd), _p(all_pos), _p(partner), None)
    for w in range(0, n, 32):
        v = x[w:w + 32]
        b = sum(1 << l for l in range(32) if v[l] & 1)
        assert np.all(ballot[w:w + 32] == b) and np.all(any_odd[w:w + 32] == int(b != 0))
        assert np.all(all_pos[w:w + 32] == int(np.all(v > 0)))
        assert np.array_equal(partner[w:w + 32], v[np.arange(32) ^ 5])
    # disjoint half-warp collectives
    n = 64 * 7
    x = rng.integers(0, 100, n).astype(np.int32)
    out = np.zeros(n, np.int32)
    lib.emu_half_warp(C.c_int(n), _p(x), _p(out), None)
    for w in range(0, n, 32):
        assert np.all(out[w:w + 16] == x[w:w + 16].sum()) and np.all(out[w + 16:w + 32] == 2 * x[w + 16:w + 32].max())
    # block sum with early exits
    n = 256 * 3 + 77
    f = rng.random(n).astype(np.float32)
    sums = np.zeros(4, np.float32)
    lib.emu_block_sum(C.c_int(n), _p(f), _p(sums), None)
    assert np.allclose(sums, [f[i:i + 256].sum() for i in range(0, n, 256)], rtol=1e-5)
    # dynamic shared memory, 2-D grid
    rows, cols = 5, 150
    a = rng.integers(0, 1000, (rows, cols)).astype(np.int32)
    out = np.zeros_like(a)
    lib.emu_reverse_rows(C.c_int(rows), C.c_int(cols), _p(a), _p(out), None)
    ref = np.concatenate([a[:, i:i + 64][:, ::-1] for i in range(0, cols, 64)], axis=1)
    assert np.array_equal(out, ref)
    # write / barrier / read neighbours
    out = np.zeros(3 * 256, np.int32)
    lib.emu_neighbour(C.c_int(3), _p(out), None)
    t = np.arange(256)
    for b in range(3):
        assert np.array_equal(out[b * 256:(b + 1) * 256], ((t + 1) % 256) * 3 + ((t + 255) % 256) * 3 + 2 * b)


def test_emulator_semantics_default_order():
    _run_all()


@pytest.mark.parametrize("order", ["reverse", "random:3"])
def test_emulator_semantics_other_orders(order):
    code = "import sys; sys.path.insert(0, '.'); import tests.test_emu_selftest_cpu as t; t._run_all(); print('selftest-ok')"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, GSB_EMU_ORDER=order),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "selftest-ok" in r.stdout, (r.stdout[-1500:], r.stderr[-1500:])


@pytest.mark.parametrize("which,needle", [("shuffle", "not a participant"), ("barrier", "cannot make progress")])
def test_emulator_refuses_undefined_behaviour(which, needle):
    """Negative controls: a shuffle that names an exited lane, and a block whose threads wait at different barriers for
    ever, abort with a diagnosis instead of producing numbers."""
    call = {"shuffle": "x = np.ones(70, np.float32); o = np.zeros(70, np.float32); "
                       "lib.emu_bad_shuffle(C.c_int(70), p(x), p(o), None)",
            "barrier": "o = np.zeros(64, np.int32); lib.emu_bad_barrier(p(o), None)"}[which]
    code = ("import sys, ctypes as C, numpy as np; sys.path.insert(0, '.')\n"
            "from tests.emu import build as emu\n"
            "lib = emu.build('selftest', simt=True)\n"
            "p = lambda a: C.c_void_p(a.ctypes.data)\n" + call + "\nprint('survived')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "survived" not in r.stdout and needle in r.stderr, (r.stdout[-500:], r.stderr[-800:])
