"""geosplatting_b200 -- the B200 (sm_100a) splat + PBR-shade hot path of GeoSplatting.

Python host code over a C-ABI CUDA library (include/geosplat_b200.h).  Public operators mirror the
reference's interfaces for this path:

    rasterization(...)                       <- gsplat.rasterization            (rfstudio/model/gsplat.py:334-355)
    shade.texture(...)                       <- nvdiffrast.torch.texture        (geosplat.py:93, _texture.py:220,:596,:604)
    splitsum.render_utils / as_splitsum      <- rfstudio_render_utils plugin, TextureCubeMap.as_splitsum
    splat.RenderableAttrs / GSplatter        <- rfstudio/model/geosplat.py:43-132, rfstudio/model/gsplat.py:284-358
    fused.splat_views(...)                   <- the per-view loop of GeoSplatter.render_report (geosplat.py:869-879)
    mgadapter.MGAdapter / compute_vertex_normals
    encoding.HashEncoding / MLP, field.GaussianField   <- the kd / ks / z fields and their glue (geosplat.py:482-674)
    flexicubes.FlexiCubes                    <- FlexiCubes.dual_marching_cubes / compute_entropy (_flexicubes.py:368-802)
    model.GeoSplatter                        <- GeoSplatter stage 1 (geosplat.py:676-927) + GeoSplatTrainer.step's loss
    loss.view_loss(...)                      <- the per-view loss of GeoSplatTrainer.step (geosplat_trainer.py:171-180)
    parallel.shard_views / GradientBucket    <- view sharding + one all-reduce per batch (new: the reference is single-GPU)

There is no CPU, PyTorch-eager or Triton fallback anywhere in this package.
"""
from . import _lib
from .rasterization import rasterization

__all__ = ["rasterization", "_lib"]
