"""geosplatting_b200 -- the B200 (sm_100a) splat + PBR-shade hot path of GeoSplatting.

Python host code over a C-ABI CUDA library (include/geosplat_b200.h).  Public operators mirror the
reference's interfaces for this path:

    rasterization(...)                  <- gsplat.rasterization   (rfstudio/model/gsplat.py:334-355)

There is no CPU, PyTorch-eager or Triton fallback anywhere in this package.
"""
from . import _lib
from .rasterization import rasterization

__all__ = ["rasterization", "_lib"]
