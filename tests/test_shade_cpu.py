"""The per-Gaussian split-sum shade, the env-stack packing and the texture drop-ins (csrc/shade.cu, texture.cu) without a
GPU: the real kernel source compiled for the host by tests/emu, driven by the real host module (geosplatting_b200/
shade.py), against the fixtures the reference's own code produced and the torch oracle.  The test bodies are the GPU
suite's own (tests/test_shade_gpu.py) with the device switched; the emulation is monkeypatched in here -- the product
refuses CPU tensors."""
import ctypes as C

import pytest
import torch

import tests.test_shade_gpu as G
from geosplatting_b200 import _lib
from geosplatting_b200 import shade as SH
from tests.emu import build as emu
from tests.emu.patch import route


@pytest.fixture()
def host_kernels(monkeypatch):
    route(monkeypatch, emu.build("shade", "texture"), SH)

    def workspace(dev, R0, L, Rb):          # shade_workspace keys its cache on the CUDA stream
        n = C.c_size_t(0)
        _lib.call("gsb_shade_workspace_bytes", None, C.c_int32(R0), C.c_int32(L), C.c_int32(Rb), C.byref(n))
        return torch.empty(max(int(n.value), 16), dtype=torch.uint8)

    monkeypatch.setattr(SH, "shade_workspace", workspace)
    monkeypatch.setattr(G, "DEV", "cpu")


@pytest.mark.parametrize("mode", ["pbr", "diffuse", "specular"])
def test_shade_kernel_source_on_host_reference_fixture(host_kernels, mode):
    G.test_shade_against_reference_fixture(mode)


def test_shade_on_the_reference_fg_lut_on_host(host_kernels, tmp_path):
    G.test_shade_on_the_reference_fg_lut_against_reference_fixture(tmp_path)


def test_splitsum_sample_on_host_reference_fixture(host_kernels):
    G.test_splitsum_sample_against_reference_fixture()


@pytest.mark.parametrize("mode", ["pbr", "specular"])
def test_shade_kernel_source_on_host_against_oracle(host_kernels, mode):
    G.test_shade_against_oracle_large(mode)


def test_texture_dropins_on_host_against_oracle(host_kernels):
    G.test_texture_dropin_against_oracle()
