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
    # the fields call the fused MLP kernel, which stages through shared memory: SIMT mode for both files
    route(monkeypatch, emu.build("hashgrid", "mlp", simt=True), E)
    monkeypatch.setattr(GE, "DEV", "cpu")


@pytest.fixture()
def host_mgadapter(monkeypatch):
    route(monkeypatch, emu.build("mgadapter"), M)
    monkeypatch.setattr(GM, "DEV", "cpu")


@pytest.mark.parametrize("tag,layers,act", [("kd", [32, 32, 32, 3], "sigmoid"), ("ks", [32, 32, 2], "none"),
                                            ("z", [32, 32, 1], "none")])
def test_hash_grid_kernel_source_on_host_matches_reference_code(host_hashgrid, tag, layers, act):
    GE.test_fields_match_reference_code(tag, layers, act)


def test_warp_aggregated_table_gradient_on_host(host_hashgrid):
    GE.test_clustered_points_table_gradient_matches_oracle()


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
    route(monkeypatch, emu.build("hashgrid", "mlp", simt=True), E)
    monkeypatch.setattr(GE, "DEV", "cpu")
    GE.test_fields_match_reference_code("ks", [32, 32, 2], "none")
    route(monkeypatch, emu.build("mgadapter", simt=True), M)
    monkeypatch.setattr(GM, "DEV", "cpu")
    GM.test_tonemap_against_reference_fixture()


@pytest.mark.parametrize("defines", [(), ("GSB_MLP_PPL=4", "GSB_MLP_BWD_WARPS=3", "GSB_MLP_FWD_WARPS=5")])
@pytest.mark.parametrize("layers,act,n", [([32, 32, 32, 3], "sigmoid", 133), ([32, 32, 2], "none", 64),
                                          ([32, 32, 1], "none", 7), ([32, 32, 32, 4], "none", 300)])
def test_fused_mlp_kernel_source_on_host_against_torch(monkeypatch, layers, act, n, defines):
    """csrc/mlp.cu (all layers + activations in one kernel as register-tiled products over transposed shared-memory
    tiles; backward = recompute + chain + the three weight gradients; 8 or 4 points per lane) under the SIMT emulation, against torch.nn.functional.linear / relu / sigmoid:
    values 1e-6, gradients 1e-5 of their largest entry; with and without the reference's input rounding; ragged N."""
    import torch
    route(monkeypatch, emu.build("mlp", simt=True, defines=defines), E)
    gen = torch.Generator().manual_seed(len(layers) * 100 + n)
    mlp = E.MLP(layers, activation=act)
    x = torch.randn(n, 32, generator=gen)
    cot = torch.randn(n, layers[-1], generator=gen)
    for scale in (0.0, 16.0):
        xs = x.clone().requires_grad_(True)
        y = mlp(xs, ref_round_scale=scale)
        assert y.shape == (n, layers[-1])
        grads = torch.autograd.grad((y * cot).sum(), [xs] + mlp.weights)
        xo = x.clone().requires_grad_(True)
        h = xo * scale + xo.detach() * (1.0 - scale) if scale else xo          # encoding.py:239-240
        ws = [w.detach().clone().requires_grad_(True) for w in mlp.weights]
        for i, w in enumerate(ws):
            h = torch.nn.functional.linear(h, w)
            h = torch.relu(h) if i < len(ws) - 1 else (torch.sigmoid(h) if act == "sigmoid" else h)
        ref = torch.autograd.grad((h * cot).sum(), [xo] + ws)
        assert float((y - h).abs().max()) <= 1e-6 * max(1.0, float(h.abs().max()))
        for k, (a, b) in enumerate(zip(grads, ref)):
            if k == 0 and scale:
                b = b / scale          # the straight-through factor lives in the hash-grid backward, not here
            assert float((a - b).abs().max()) <= 1e-5 * max(float(b.abs().max()), 1e-6), (k, scale)
    # leading batch dimensions and an input that needs no gradient
    y3 = mlp(torch.randn(2, 5, 32, generator=gen))
    assert y3.shape == (2, 5, layers[-1])
    assert E.fused_mlp_supported([32, 32, 3], "sigmoid") and not E.fused_mlp_supported([32, 64, 3], "none")
    assert not E.fused_mlp_supported([32, 32, 3], "tanh")
