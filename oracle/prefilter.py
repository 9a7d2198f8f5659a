"""numpy/ctypes front-end of oracle/prefilter_oracle.c (CPU restatement of the reference's
`rfstudio_render_utils` plugin) plus a loader for the REAL plugin built from the reference sources
(oracle/_ref/rfstudio_render_utils.so, see oracle/build_ref.sh).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import importlib.util
import os

import numpy as np

from .raster import _f32, _p, lib

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "rfstudio_render_utils.so")


def ndf_cutoff_costheta(roughness: float, cutoff: float = 0.99) -> float:
    """rfstudio/graphics/_mesh/_splitsum/_wrap.py:120-132."""
    n = 1000000
    costheta = np.cos(np.linspace(0, np.pi / 2.0, n))
    a2 = roughness ** 4
    c = np.clip(costheta, 0.0, 1.0)
    d = (c * a2 - c) * c + 1.0
    D = np.cumsum(a2 / (d * d * np.pi))
    return float(costheta[np.argmax(D >= D[-1] * cutoff)])


def diffuse_fwd(cubemap):
    c = _f32(cubemap)
    out = np.zeros_like(c)
    lib().orc_diffuse_cubemap_fwd(C.c_int(c.shape[1]), _p(c), _p(out))
    return out


def diffuse_bwd(grad):
    g = _f32(grad)
    out = np.zeros_like(g)
    lib().orc_diffuse_cubemap_bwd(C.c_int(g.shape[1]), _p(g), _p(out))
    return out


def specular_bounds(R, costheta_cutoff):
    out = np.zeros((6, R, R, 24), np.float32)
    lib().orc_specular_bounds(C.c_int(R), C.c_float(costheta_cutoff), _p(out))
    return out


def specular_fwd(cubemap, bounds, roughness, costheta_cutoff):
    c, b = _f32(cubemap), _f32(bounds)
    R = c.shape[1]
    out = np.zeros((6, R, R, 4), np.float32)
    lib().orc_specular_cubemap_fwd(C.c_int(R), _p(c), _p(b), C.c_float(roughness), C.c_float(costheta_cutoff), _p(out))
    return out


def specular_bwd(bounds, grad4, roughness, costheta_cutoff):
    b, g = _f32(bounds), _f32(grad4)
    R = g.shape[1]
    out = np.zeros((6, R, R, 3), np.float32)
    lib().orc_specular_cubemap_bwd(C.c_int(R), _p(b), _p(g), C.c_float(roughness), C.c_float(costheta_cutoff), _p(out))
    return out


def load_reference_plugin():
    """The reference's own compiled plugin (CUDA; GPU box only).  Returns None when it was not built."""
    if not os.path.exists(REF_SO):
        return None
    import torch  # noqa: F401  (its shared libraries must be loaded first)
    spec = importlib.util.spec_from_file_location("rfstudio_render_utils", REF_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def specular_fwd_f64(cubemap, bounds, roughness, costheta_cutoff, stride=1):
    """Double-precision yardstick for output texels 0, stride, 2*stride, ... -> [n,5] float64
    (sum w*rgb, sum w, fragile)."""
    c, b = _f32(cubemap), _f32(bounds)
    R = c.shape[1]
    n = (6 * R * R + stride - 1) // stride
    out = np.zeros((n, 5), np.float64)
    lib().orc_specular_cubemap_fwd_f64(C.c_int(R), _p(c), _p(b), C.c_float(roughness), C.c_float(costheta_cutoff),
                                       C.c_int(stride), _p(out))
    return out
