"""Hash-grid fields that produce kd / ks / z for the hot path, over the C ABI
(gsb_hashgrid_fwd / gsb_hashgrid_bwd).

Mirrors, with the same field names and argument meaning:
    HashEncoding          rfstudio/model/components/encoding.py:96-241   (backend='torch' semantics)
    MLP                   rfstudio/nn/mlp.py:27-145                       (bias / skip connections / weight norm are not
                                                                          used by GeoSplatting's fields and raise)
    GaussianField configs rfstudio/model/geosplat.py:485-518              -> kd_field(), ks_field(), z_field()
The gather / scatter over the table is the hand-written part (L2-bound random access); the MLP is three bias-free
GEMMs on cuBLAS through torch.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence

import torch
from torch import Tensor, nn

from ._lib import call, f32c, ptr, stream_ptr
from ._lib import require_cuda as _require_cuda


def level_scalings(num_levels: int, min_res: int, max_res: int) -> List[float]:
    """encoding.py:129-135: floor(min_res * growth ** level) (computed like the reference, in fp32 torch ops)."""
    levels = torch.arange(num_levels)
    growth = math.exp((math.log(max_res) - math.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
    return [float(v) for v in torch.floor(min_res * growth ** levels)]


class _HashGrid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, table: Tensor, scalings, log2_T: int, table_grad_scale: float,
                straight_through: bool = False):
        x_c, table_c = f32c(x), f32c(table)
        if straight_through and table_grad_scale != 1.0:
            # encoding.py:232-233 evaluates the grid at x * (1/s) + x.detach() * (1 - 1/s): the same point up to an ulp,
            # and an ulp decides the cell when x sits on a cell boundary of some level -- take the reference's point
            inv = 1.0 / table_grad_scale
            x_c = x_c * inv + x_c * (1.0 - inv)
        dev = x_c.device
        N, L, F = x_c.shape[0], len(scalings), table_c.shape[1]
        feats = torch.empty(N, L * F, dtype=torch.float32, device=dev)
        sc = (C.c_float * L)(*scalings)
        call("gsb_hashgrid_fwd", dev, C.c_int64(N), ptr(x_c), ptr(table_c), C.c_int32(L), C.c_int32(F), C.c_int32(log2_T),
             sc, ptr(feats), stream_ptr(dev))
        ctx.save_for_backward(x_c, table_c)
        ctx.misc = (sc, L, F, log2_T, float(table_grad_scale))
        return feats

    @staticmethod
    def backward(ctx, v_feats):
        x, table = ctx.saved_tensors
        sc, L, F, log2_T, gscale = ctx.misc
        dev = x.device
        N = x.shape[0]
        v_table = torch.zeros_like(table) if ctx.needs_input_grad[1] else None
        v_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        if v_table is None and v_x is None:
            return None, None, None, None, None, None
        call("gsb_hashgrid_bwd", dev, C.c_int64(N), ptr(x), ptr(table), C.c_int32(L), C.c_int32(F), C.c_int32(log2_T), sc,
             ptr(f32c(v_feats)), C.c_float(gscale), ptr(v_table), ptr(v_x), stream_ptr(dev))
        return v_x, v_table, None, None, None, None


class _ReferenceRounding(torch.autograd.Function):
    """encoding.py:239-240 hands the MLP `feats * s + feats.detach() * (1 - s)`: the features up to an fp32 rounding, and a
    rounding decides which side of a ReLU kink a hidden unit is on.  Same values here; the gradient factor s lives in the
    hash-grid backward kernel, so this node passes gradients through unchanged."""

    @staticmethod
    def forward(ctx, feats: Tensor, s: float):
        return feats * s + feats * (1.0 - s)

    @staticmethod
    def backward(ctx, v):
        return v, None


_ACT = {"none": 0, "sigmoid": 1}


def fused_mlp_supported(layers: Sequence[int], activation: str) -> bool:
    """The shapes of GeoSplatting's fields (geosplat.py:485-518): 32 -> 32 [-> 32] -> {1..4}, `none` / `sigmoid`."""
    return (len(layers) in (3, 4) and all(w == 32 for w in layers[:-1]) and 1 <= layers[-1] <= 4
            and activation in _ACT)


class _FusedMLP(torch.autograd.Function):
    """All layers, ReLUs and the output activation in one kernel each way (csrc/mlp.cu)."""

    @staticmethod
    def forward(ctx, x: Tensor, activation: str, ref_round_scale: float, *weights: Tensor):
        x_c = f32c(x).reshape(-1, 32)
        ws = [f32c(w) for w in weights]
        dev = x_c.device
        N, dout, nh = x_c.shape[0], ws[-1].shape[0], len(ws) - 1
        y = torch.empty(N, dout, dtype=torch.float32, device=dev)
        call("gsb_mlp_fwd", dev, C.c_int64(N), ptr(x_c), ptr(ws[0]), ptr(ws[1]) if nh == 2 else None, ptr(ws[-1]),
             C.c_int32(nh), C.c_int32(dout), C.c_int32(_ACT[activation]), C.c_float(ref_round_scale), ptr(y),
             stream_ptr(dev))
        ctx.save_for_backward(x_c, *ws)
        ctx.misc = (activation, float(ref_round_scale), x.shape)
        return y.view(*x.shape[:-1], dout)

    @staticmethod
    def backward(ctx, v_y: Tensor):
        x_c, *ws = ctx.saved_tensors
        activation, rr, x_shape = ctx.misc
        dev = x_c.device
        N, dout, nh = x_c.shape[0], ws[-1].shape[0], len(ws) - 1
        v_x = torch.empty_like(x_c) if ctx.needs_input_grad[0] else None
        v_ws = [torch.empty_like(w) for w in ws]
        call("gsb_mlp_bwd", dev, C.c_int64(N), ptr(x_c), ptr(ws[0]), ptr(ws[1]) if nh == 2 else None, ptr(ws[-1]),
             C.c_int32(nh), C.c_int32(dout), C.c_int32(_ACT[activation]), C.c_float(rr), ptr(f32c(v_y).reshape(N, dout)),
             ptr(v_x), ptr(v_ws[0]), ptr(v_ws[1]) if nh == 2 else None, ptr(v_ws[-1]), stream_ptr(dev))
        return (None if v_x is None else v_x.view(x_shape), None, None, *v_ws)


class MLP(nn.Module):
    """rfstudio/nn/mlp.py: `layers` [in, hidden..., out] (in may be -1 = inferred), ReLU between layers, `activation`
    after the last, kaiming-uniform weights, no bias."""

    def __init__(self, layers: Sequence[int], activation: str = "none", bias: bool = False,
                 initialization: str = "kaiming-uniform", skip_connections: Sequence[int] = (), weight_norm: bool = False,
                 in_dim: Optional[int] = None):
        super().__init__()
        if bias or skip_connections or weight_norm:
            raise NotImplementedError("geosplatting_b200.MLP: GeoSplatting's fields use bias=False, no skip connections, "
                                      "no weight norm (rfstudio/model/geosplat.py:485-518)")
        if activation not in ("none", "sigmoid", "relu", "tanh", "softplus"):
            raise ValueError(activation)
        if initialization != "kaiming-uniform":
            raise NotImplementedError(initialization)
        layers = list(layers)
        if layers[0] == -1:
            if in_dim is None:
                raise ValueError("MLP(layers=[-1, ...]) needs in_dim")
            layers[0] = in_dim
        self.layers, self.activation = layers, activation
        # parameter names are the reference's (mlp.py:46-57: `nn_layers.<i>.weight`), so a state_dict of this module loads
        # into the reference's MLP and back (geosplat.py:848 exports ks_enc.state_dict(); geosplat_mc.py:73 loads it)
        self.nn_layers = nn.ModuleList(nn.Linear(i, o, bias=False) for i, o in zip(layers[:-1], layers[1:]))
        for layer in self.nn_layers:
            nn.init.kaiming_uniform_(layer.weight, nonlinearity="relu")   # mlp.py:99-100

    @property
    def weights(self) -> List[nn.Parameter]:
        return [layer.weight for layer in self.nn_layers]

    def forward(self, x: Tensor, ref_round_scale: float = 0.0) -> Tensor:
        """`ref_round_scale` s != 0: the input is taken as x * s + x * (1 - s), the value (up to an fp32 rounding the
        same as x) that HashEncoding.__call__ hands the MLP (encoding.py:239-240)."""
        if fused_mlp_supported(self.layers, self.activation) and x.shape[-1] == 32 and not x.device.type == "meta":
            _require_cuda(x, "MLP")
            return _FusedMLP.apply(x, self.activation, float(ref_round_scale), *self.weights)
        if ref_round_scale:
            x = _ReferenceRounding.apply(x, float(ref_round_scale))
        n = len(self.nn_layers)
        for i, layer in enumerate(self.nn_layers):
            x = layer(x)
            if i < n - 1:
                x = torch.relu(x)
        return {"none": lambda t: t, "sigmoid": torch.sigmoid, "relu": torch.relu, "tanh": torch.tanh,
                "softplus": torch.nn.functional.softplus}[self.activation](x)


class _Table(nn.Module):
    """The reference's ParameterModule (rfstudio/nn/module.py:1478-1492): one parameter called `params`."""

    def __init__(self, t: Tensor):
        super().__init__()
        self.params = nn.Parameter(t)


class HashEncoding(nn.Module):
    """rfstudio/model/components/encoding.py:96-241 (backend='torch' semantics, interpolation 'linear')."""

    def __init__(self, mlp: MLP, num_levels: int = 16, min_res: int = 16, max_res: int = 1024,
                 log2_hashmap_size: int = 19, features_per_level: int = 2, hash_init_scale: float = 0.001,
                 interpolation: str = "linear", grad_scaling: Optional[float] = None):
        super().__init__()
        assert grad_scaling is None or grad_scaling > 0                                    # encoding.py:126
        if interpolation != "linear":
            raise NotImplementedError(f"interpolation '{interpolation}' is not supported")  # as the torch backend, :139-143
        if features_per_level != 2:
            raise NotImplementedError("geosplatting_b200.HashEncoding: features_per_level must be 2")
        self.mlp, self.num_levels, self.min_res, self.max_res = mlp, num_levels, min_res, max_res
        self.log2_hashmap_size, self.features_per_level, self.grad_scaling = log2_hashmap_size, features_per_level, grad_scaling
        self.hash_table_size = 2 ** log2_hashmap_size
        self.scalings = level_scalings(num_levels, min_res, max_res)
        # registered as `encoder.params`, the reference's key for either backend (encoding.py:144-148: ParameterModule;
        # tcnn.Encoding.params)
        self.encoder = _Table((torch.rand(self.hash_table_size * num_levels, features_per_level) * 2 - 1)
                              * hash_init_scale)                                            # encoding.py:144-147

    @property
    def hash_table(self) -> nn.Parameter:
        return self.encoder.params

    def encode(self, in_tensor: Tensor, straight_through: bool = False) -> Tensor:
        """pytorch_fwd (encoding.py:182-229): [..., 3] in [-1, 1] -> [..., num_levels * features_per_level].
        `straight_through`: evaluate at the point `__call__` (encoding.py:231-233) hands to pytorch_fwd."""
        _require_cuda(in_tensor, "HashEncoding")
        assert in_tensor.shape[-1] == 3
        flat = in_tensor.reshape(-1, 3)
        feats = _HashGrid.apply(flat, self.hash_table, self.scalings, self.log2_hashmap_size,
                                1.0 if self.grad_scaling is None else float(self.grad_scaling), straight_through)
        return feats.view(*in_tensor.shape[:-1], self.num_levels * self.features_per_level)

    def forward(self, in_tensor: Tensor) -> Tensor:
        """encoding.py:231-241.  The two straight-through expressions of the reference (`x/s + x.detach()(1-1/s)`,
        `f*s + f.detach()(1-s)`) leave values unchanged and multiply the gradient that reaches the table by s while the
        one that reaches x stays as it is: that factor is applied inside the backward kernel."""
        feats = self.encode(in_tensor, straight_through=True)
        return self.mlp(feats, ref_round_scale=0.0 if self.grad_scaling is None else float(self.grad_scaling))


class TcnnEncoding(nn.Module):
    """Stand-in for `tinycudann.Encoding(n_input_dims=3, encoding_config={"otype": "HashGrid", ...})` as the reference
    constructs and calls it (rfstudio/model/components/encoding.py:150-163, :235-236): one parameter `params`, called
    with points in [0, 1]^3, returns [N, n_levels * n_features_per_level].

    Semantics are those of the reference's own `backend='torch'` (hash, level resolutions, trilinear weights; fp32
    storage and output), NOT tinycudann's (different level layout, fp16 tables): a tcnn checkpoint does not load."""

    def __init__(self, n_input_dims: int, encoding_config: dict, seed: Optional[int] = None, dtype=None):
        super().__init__()
        c = dict(encoding_config)
        if n_input_dims != 3 or c.get("otype", "HashGrid") != "HashGrid":
            raise NotImplementedError("TcnnEncoding: only the 3-D HashGrid encoding of GeoSplatting's fields")
        if str(c.get("interpolation", "Linear")).lower() != "linear":
            raise NotImplementedError(f"interpolation {c.get('interpolation')!r} is not supported")
        self.n_levels, self.n_features = int(c.get("n_levels", 16)), int(c.get("n_features_per_level", 2))
        if self.n_features != 2:
            raise NotImplementedError("TcnnEncoding: n_features_per_level must be 2")
        self.log2_hashmap_size = int(c.get("log2_hashmap_size", 19))
        base, growth = int(c.get("base_resolution", 16)), float(c.get("per_level_scale", 2.0))
        self.scalings = [float(v) for v in torch.floor(base * growth ** torch.arange(self.n_levels))]
        self.n_output_dims = self.n_levels * self.n_features
        gen = None if seed is None else torch.Generator().manual_seed(seed)
        self.params = nn.Parameter((torch.rand((2 ** self.log2_hashmap_size) * self.n_levels, self.n_features,
                                               generator=gen) * 2 - 1) * 1e-3)

    def forward(self, x01: Tensor) -> Tensor:
        _require_cuda(x01, "TcnnEncoding")
        assert x01.shape[-1] == 3
        return _HashGrid.apply(x01.reshape(-1, 3) * 2.0 - 1.0, self.params, self.scalings, self.log2_hashmap_size, 1.0)


def kd_field() -> HashEncoding:
    """GaussianField.kd_enc (geosplat.py:485-495)."""
    return HashEncoding(MLP([32, 32, 32, 3], activation="sigmoid"), grad_scaling=16.0, max_res=4096, log2_hashmap_size=18)


def ks_field() -> HashEncoding:
    """GaussianField.ks_enc (geosplat.py:497-507)."""
    return HashEncoding(MLP([32, 32, 2], activation="none"), grad_scaling=16.0, max_res=4096, log2_hashmap_size=18)


def z_field() -> HashEncoding:
    """GaussianField.z_enc (geosplat.py:508-518)."""
    return HashEncoding(MLP([32, 32, 1], activation="none"), grad_scaling=16.0, max_res=4096, log2_hashmap_size=18)
