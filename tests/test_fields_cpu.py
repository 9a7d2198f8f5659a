"""The hash-grid fields and MGAdaptor sampling / vertex normals / tone map (csrc/hashgrid.cu, mgadapter.cu) without a GPU:
the real kernel source compiled for the host by tests/emu, driven by the real host modules, on the GPU suite's own test
bodies (tests/test_encoding_gpu.py, tests/test_mgadapter_gpu.py) with the device switched -- the hash encoding is held
to BIT equality with the reference's torch backend here too.  The warp-level reductions (d/dx of the encoding, the
exposure gradient) take their atomics route under emulation."""
import pytest

import tests.test_encoding_gpu as GE
import tests.test_mgadapter_gpu as GM
from geosplatting_b200 import encoding as E
from geosplatting_b200 import mgadapter as M
from tests.emu import build as emu
from tests.emu.patch import route


@pytest.fixture()
def host_hashgrid(monkeypatch):
    route(monkeypatch, emu.build("hashgrid"), E)
    monkeypatch.setattr(GE, "DEV", "cpu")


@pytest.fixture()
def host_mgadapter(monkeypatch):
    route(monkeypatch, emu.build("mgadapter"), M)
    monkeypatch.setattr(GM, "DEV", "cpu")


@pytest.mark.parametrize("tag,layers,act", [("kd", [32, 32, 32, 3], "sigmoid"), ("ks", [32, 32, 2], "none"),
                                            ("z", [32, 32, 1], "none")])
def test_hash_grid_kernel_source_on_host_matches_reference_code(host_hashgrid, tag, layers, act):
    GE.test_fields_match_reference_code(tag, layers, act)


def test_mgadapter_kernel_source_on_host_reference_fixture(host_mgadapter):
    GM.test_against_reference_fixture()


def test_mgadapter_kernel_source_on_host_against_oracle(host_mgadapter):
    GM.test_against_oracle_on_a_noisy_icosphere()


def test_mgadapter_degenerate_faces_and_tonemap_on_host(host_mgadapter):
    GM.test_flat_normals_and_degenerate_faces()
    GM.test_tonemap_against_reference_fixture()


def test_warp_reductions_under_simt(monkeypatch):
    """SIMT mode of tests/emu (threads of a block as fibers): the segmented 16-lane butterfly that sums d/dx of the hash
    encoding over its levels and the block reduction of the exposure gradient run as written."""
    route(monkeypatch, emu.build("hashgrid", simt=True), E)
    monkeypatch.setattr(GE, "DEV", "cpu")
    GE.test_fields_match_reference_code("ks", [32, 32, 2], "none")
    route(monkeypatch, emu.build("mgadapter", simt=True), M)
    monkeypatch.setattr(GM, "DEV", "cpu")
    GM.test_tonemap_against_reference_fixture()
