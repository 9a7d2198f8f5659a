"""CPU restatement (torch, any dtype, autograd) of the hash-grid fields that produce kd / ks / z for the hot path:
HashEncoding with the reference's own `torch` backend + the bias-free MLP behind it.

TEST INFRASTRUCTURE ONLY -- never imported by the product package.  Follows
    rfstudio/model/components/encoding.py:124-138 (__setup__: growth factor, per-level scalings, table offsets),
    :164-180 (hash_fn), :182-229 (pytorch_fwd: corners, offsets, trilinear blend), :231-241 (__call__: grad scaling),
    rfstudio/nn/mlp.py:125-145 (MLP.__call__: ReLU between layers, `activation` after the last).
Pinned on tests/golden/ref_encoding.npz, which scripts/make_golden.py produces by running those very methods of the
reference (tests/test_golden_cpu.py::test_encoding_oracle_matches_reference_code).

The reference's default backend is tinycudann (unpinned git HEAD, absent here; a different hash/level layout);
`backend='torch'` is the reference's own documented fallback with the same interface and is the contract here.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import torch
from torch import Tensor

PRIMES = (1, 2654435761, 805459861)


def level_scalings(num_levels: int = 16, min_res: int = 16, max_res: int = 4096) -> Tensor:
    """encoding.py:129-135: floor(min_res * growth ** level), growth = exp((ln max - ln min) / (L - 1))."""
    levels = torch.arange(num_levels)
    growth = math.exp((math.log(max_res) - math.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
    return torch.floor(min_res * growth ** levels)


def hash_index(ijk: Tensor, log2_hashmap_size: int) -> Tensor:
    """encoding.py:164-180 (without the per-level offset): ijk [..., L, 3] integer -> [..., L] in [0, 2^log2)."""
    v = ijk.to(torch.int64) * torch.tensor(PRIMES, dtype=torch.int64)
    h = torch.bitwise_xor(torch.bitwise_xor(v[..., 0], v[..., 1]), v[..., 2])
    return h % (2 ** log2_hashmap_size)


def hash_encode(x: Tensor, table: Tensor, scalings: Tensor, log2_hashmap_size: int) -> Tensor:
    """encoding.py:182-229.  x [..., 3] in [-1, 1], table [L * 2^log2, F] -> [..., L * F]."""
    L = scalings.shape[0]
    T = 2 ** log2_hashmap_size
    offs = torch.arange(L) * T
    t = x[..., None, :] * 0.5 + 0.5
    scaled = t * scalings.view(-1, 1).to(t.dtype)
    c = torch.ceil(scaled).to(torch.int32)
    f = torch.floor(scaled).to(torch.int32)
    o = scaled - f

    def pick(ix, iy, iz):
        ijk = torch.stack([(c if ix else f)[..., 0], (c if iy else f)[..., 1], (c if iz else f)[..., 2]], dim=-1)
        return table[hash_index(ijk, log2_hashmap_size) + offs]

    f0, f1, f2, f3 = pick(1, 1, 1), pick(1, 0, 1), pick(0, 0, 1), pick(0, 1, 1)
    f4, f5, f6, f7 = pick(1, 1, 0), pick(1, 0, 0), pick(0, 0, 0), pick(0, 1, 0)
    ox, oy, oz = o[..., 0:1], o[..., 1:2], o[..., 2:3]
    f03 = f0 * ox + f3 * (1 - ox)
    f12 = f1 * ox + f2 * (1 - ox)
    f56 = f5 * ox + f6 * (1 - ox)
    f47 = f4 * ox + f7 * (1 - ox)
    f0312 = f03 * oy + f12 * (1 - oy)
    f4756 = f47 * oy + f56 * (1 - oy)
    out = f0312 * oz + f4756 * (1 - oz)
    return out.flatten(-2)


def mlp(feats: Tensor, weights: Sequence[Tensor], activation: str = "none") -> Tensor:
    """rfstudio/nn/mlp.py:125-145 with bias=False, no skip connections."""
    x = feats
    for i, w in enumerate(weights):
        x = torch.nn.functional.linear(x, w)
        if i < len(weights) - 1:
            x = torch.relu(x)
        elif activation == "sigmoid":
            x = x.sigmoid()
        elif activation != "none":
            raise ValueError(activation)
    return x


def field(x: Tensor, table: Tensor, weights: List[Tensor], scalings: Tensor, log2_hashmap_size: int, activation: str,
          grad_scaling: float = 16.0) -> Tensor:
    """encoding.py:231-241: values unchanged, gradient to the table scaled by `grad_scaling`."""
    if grad_scaling is not None:
        x = x * (1 / grad_scaling) + x.detach() * (1 - 1 / grad_scaling)
    feats = hash_encode(x, table, scalings, log2_hashmap_size)
    if grad_scaling is not None:
        feats = feats * grad_scaling + feats.detach() * (1 - grad_scaling)
    return mlp(feats, weights, activation)
