"""GPU parity of the hash-grid fields through the reference-facing `HashEncoding` / `MLP`
mirrors (C ABI: gsb_hashgrid_fwd / gsb_hashgrid_bwd).

* against tests/golden/ref_encoding.npz -- produced by the reference's OWN HashEncoding (torch backend) and MLP code
  (scripts/make_golden.py section H): features bit-identical, outputs 1e-6, every gradient 1e-5 of its max;
* against the CPU oracle at GaussianField's real configuration (2^18 entries x 16 levels, max_res 4096);
* edge cases: out-of-range / lattice-exact coordinates, empty input, features_per_level != 2, CPU tensors.
"""
import os

import numpy as np
import pytest
import torch

from geosplatting_b200 import encoding as E
from oracle import encoding as OE

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _field(tag, layers, act, log2, g):
    enc = E.HashEncoding(E.MLP(layers, activation=act), grad_scaling=16.0, max_res=4096, log2_hashmap_size=log2).to(DEV)
    with torch.no_grad():
        enc.hash_table.copy_(torch.from_numpy(g[f"{tag}_table"]))
        for k, w in enumerate(enc.mlp.weights):
            w.copy_(torch.from_numpy(g[f"{tag}_w{k}"]))
    return enc


@pytest.mark.parametrize("tag,layers,act", [("kd", [32, 32, 32, 3], "sigmoid"), ("ks", [32, 32, 2], "none"),
                                            ("z", [32, 32, 1], "none")])
def test_fields_match_reference_code(tag, layers, act):
    g = np.load(os.path.join(GOLD, "ref_encoding.npz"))
    enc = _field(tag, layers, act, int(g[f"{tag}_log2"]), g)
    assert np.array_equal(np.asarray(enc.scalings, np.float32), g[f"{tag}_scalings"])
    x = torch.from_numpy(g[f"{tag}_x"]).to(DEV).requires_grad_(True)
    feats = enc.encode(x)
    assert np.array_equal(feats.detach().cpu().numpy(), g[f"{tag}_feats"])         # bit-identical features
    y = enc(x)
    # the MLP is fp32 GEMMs on cuBLAS here and on the CPU BLAS in the reference run: 1e-6 of the output scale
    assert np.abs(y.detach().cpu().numpy() - g[f"{tag}_y"]).max() <= 1e-6 * max(1.0, np.abs(g[f"{tag}_y"]).max())
    grads = torch.autograd.grad((y * torch.from_numpy(g[f"{tag}_cot"]).to(DEV)).sum(),
                                [x, enc.hash_table] + list(enc.mlp.weights))
    ref_vx = g[f"{tag}_v_x"]
    assert np.abs(grads[0].cpu().numpy() - ref_vx).max() <= 1e-5 * np.abs(ref_vx).max()
    vt = np.zeros_like(g[f"{tag}_table"])
    vt[g[f"{tag}_v_table_idx"]] = g[f"{tag}_v_table_val"]
    assert np.abs(grads[1].cpu().numpy() - vt).max() <= 1e-5 * np.abs(vt).max()
    for k in range(len(layers) - 1):
        ref = g[f"{tag}_v_w{k}"]
        assert np.abs(grads[2 + k].cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()


def test_full_size_config_matches_oracle():
    """GaussianField.kd_enc as configured (geosplat.py:485-495): 2^18 x 16 x 2 table, 60 000 points."""
    torch.manual_seed(3)
    enc = E.kd_field()
    with torch.no_grad():
        enc.hash_table.mul_(1000.0)
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(60_000, 3, generator=gen) * 2 - 1
    cot = torch.randn(60_000, 3, generator=gen)
    table, ws = enc.hash_table.detach().clone(), [w.detach().clone() for w in enc.mlp.weights]
    # oracle (CPU)
    ox, ot = x.clone().requires_grad_(True), table.clone().requires_grad_(True)
    ows = [w.clone().requires_grad_(True) for w in ws]
    scal = OE.level_scalings(16, 16, 4096)
    o_feats = OE.hash_encode(ox, ot, scal, 18)
    oy = OE.field(ox, ot, ows, scal, 18, "sigmoid", 16.0)
    og = torch.autograd.grad((oy * cot).sum(), [ox, ot] + ows)
    # this library
    enc = enc.to(DEV)
    dx = x.to(DEV).requires_grad_(True)
    feats = enc.encode(dx)
    assert torch.equal(feats.cpu(), o_feats.detach())
    y = enc(dx)
    assert float((y.cpu() - oy.detach()).abs().max()) <= 1e-6 * max(1.0, float(oy.abs().max()))
    dg = torch.autograd.grad((y * cot.to(DEV)).sum(), [dx, enc.hash_table] + list(enc.mlp.weights))
    for a, b in zip(dg, og):
        assert float((a.cpu() - b).abs().max()) <= 2e-5 * float(b.abs().max())


def test_clustered_points_table_gradient_matches_oracle():
    """Points in MESH ORDER (32 consecutive points are neighbours on a surface, as MGAdaptor emits them): at the coarse levels
    the 32 points of a warp fall into a handful of cells and the backward sums them inside the warp before it touches the
    table (hashgrid.cu: warp_grouped_add); finer levels scatter per lane.  Ragged count (not a multiple of 32); table and
    position gradients against the CPU oracle."""
    torch.manual_seed(4)
    enc = E.HashEncoding(E.MLP([32, 32, 2]), log2_hashmap_size=14, max_res=2048, grad_scaling=None).to("cpu")
    with torch.no_grad():
        enc.hash_table.mul_(1000.0)
    gen = torch.Generator().manual_seed(8)
    n = 32 * 40 + 13
    centres = torch.nn.functional.normalize(torch.randn(n // 32 + 1, 3, generator=gen), dim=-1) * 0.6
    x = centres.repeat_interleave(32, 0)[:n] + 4e-3 * torch.randn(n, 3, generator=gen)   # patches of 32 neighbours
    cot = torch.randn(n, 32, generator=gen)
    scal = OE.level_scalings(16, 16, 2048)
    ox, ot = x.clone().requires_grad_(True), enc.hash_table.detach().clone().requires_grad_(True)
    o_feats = OE.hash_encode(ox, ot, scal, 14)
    og = torch.autograd.grad((o_feats * cot).sum(), [ox, ot])
    enc = enc.to(DEV)
    dx = x.to(DEV).requires_grad_(True)
    feats = enc.encode(dx)
    assert torch.equal(feats.cpu(), o_feats.detach())
    dg = torch.autograd.grad((feats * cot.to(DEV)).sum(), [dx, enc.hash_table])
    for a, b in zip(dg, og):
        assert float((a.cpu() - b).abs().max()) <= 2e-5 * float(b.abs().max())
    # the coarse levels really are shared inside a warp: 32 consecutive points, at most a few distinct cells at level 0,
    # (nearly) all distinct at the finest level
    cells = lambda lvl: [len(torch.unique(torch.floor((x[i:i + 32] * 0.5 + 0.5) * scal[lvl]), dim=0)) for i in range(0, n, 32)]
    assert max(cells(0)) <= 6 and min(cells(15)[:-1]) >= 16, (cells(0), cells(15))


def test_edge_cases():
    enc = E.HashEncoding(E.MLP([32, 16, 2]), log2_hashmap_size=8, max_res=512, grad_scaling=None).to(DEV)
    with torch.no_grad():
        enc.hash_table.mul_(1000.0)
    scal = OE.level_scalings(16, 16, 512)
    x = torch.tensor([[-1.0, -1, -1], [1, 1, 1], [1.5, -2.0, 0.3], [0, 0, 0], [-0.5, 0.125, 0.99999994]])
    ref = OE.hash_encode(x, enc.hash_table.detach().cpu(), scal, 8)
    assert torch.equal(enc.encode(x.to(DEV)).cpu(), ref)                            # outside [-1,1]: hashes of negative cells too
    assert enc(torch.zeros(0, 3, device=DEV)).shape == (0, 2)
    assert enc.encode(torch.zeros(4, 5, 3, device=DEV)).shape == (4, 5, 32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        enc.encode(torch.zeros(3, 3))
    with pytest.raises(NotImplementedError):
        E.HashEncoding(E.MLP([64, 2]), features_per_level=4)
    with pytest.raises(NotImplementedError):
        E.MLP([32, 32, 3], bias=True)
    # gradient only w.r.t. the table (positions detached, as z_enc is called at geosplat.py:644)
    xd = (torch.rand(500, 3) * 2 - 1).to(DEV)
    (g,) = torch.autograd.grad(enc.encode(xd).sum(), [enc.hash_table])
    assert bool(torch.isfinite(g).all()) and float(g.abs().sum()) > 0
