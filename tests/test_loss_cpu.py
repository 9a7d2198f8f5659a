"""The fused per-view loss kernels (csrc/loss.cu: separable 11 x 11 SSIM windows through shared memory, block
reductions) without a GPU: the real kernel source under the SIMT mode of tests/emu, driven by the real host module, on
the GPU suite's own test body (tests/test_loss_gpu.py) at its ragged sizes."""
import pytest

import tests.test_loss_gpu as G
from geosplatting_b200 import loss as L
from tests.emu import build as emu
from tests.emu.patch import route


@pytest.fixture()
def host_loss(monkeypatch):
    route(monkeypatch, emu.build("loss", simt=True), L)
    monkeypatch.setattr(G, "DEV", "cpu")


@pytest.mark.parametrize("H,W,seed", [(123, 77, 1), (16, 40, 2), (11, 11, 3)])
def test_loss_kernel_source_on_host_matches_oracle(host_loss, H, W, seed):
    G.test_loss_and_cotangent_match_oracle(H, W, seed)
