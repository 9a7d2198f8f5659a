"""Race fuzzing of the block-level kernels on the host: tests/emu's SIMT scheduler resumes the fibers of a block in
forward order by default; code whose only inter-thread ordering comes from __syncthreads and the *_sync warp primitives
must give the same results in ANY order.  The compositing, loss and FlexiCubes / hash-grid / MGAdaptor warp-reduction
tests are re-run with the order reversed and with a fresh random permutation every round (GSB_EMU_ORDER): a missing
__syncwarp or __syncthreads around shared memory would make them fail."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("order", ["reverse", "random:7"])
def test_results_do_not_depend_on_the_thread_schedule(order):
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", "tests/test_composite_cpu.py",
                        "tests/test_loss_cpu.py", "tests/test_flexicubes_cpu.py", "tests/test_fields_cpu.py",
                        "-k", "(composite and not 10000) or loss or simt"],
                       cwd=ROOT, env=dict(os.environ, GSB_EMU_ORDER=order), capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert " passed" in r.stdout and "no tests ran" not in r.stdout
