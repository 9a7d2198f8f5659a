"""GeoSplatter stage 1 over this library: the caller that strings the rows of SURVEY.md section 8 together.

Mirrors (same names, argument meaning and defaults; host orchestration only -- every tensor op of substance is a call
into the C-ABI library through the operator mirrors of this package):

    GeoSplatter.__setup__ / get_geometry / get_background_color / get_envmap / get_gsplat / render_report
                                                     rfstudio/model/geosplat.py:676-927
    GeoSplatTrainer.step (loss assembly only)         rfstudio/trainer/geosplat_trainer.py:146-182  -> training_loss()

What it drives per step: FlexiCubes mesh + regulariser (flexicubes.py) -> vertex normals + MGAdaptor + kd / ks / z
hash-grid fields (field.py) -> split-sum prefilter (splitsum.py) -> the batch of views (fused.splat_views: shade,
rasterize, tone map) -> per-view loss (loss.py).  `smooth_type` 'grad' / 'tv' (extra render_rgb passes per view, off in
every shipped stage-1 recipe until their weights are raised) are not mirrored; 'jitter' (the default) is.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor, nn

from . import splitsum
from .field import GaussianField
from .flexicubes import FlexiCubes, TriangleMesh
from .fused import splat_views
from .loss import view_loss
from .scenes import PinholeCamera
from .splat import GSplatter, RenderableAttrs

_GUESS = {"outdoor": (0.0, 0.0), "diffuse": (0.0, -3.0), "hybrid": (-3.0, -3.0), "specular": (-3.0, 0.0),
          "glossy": (-3.0, 0.0)}                                                # geosplat.py:727-740


def srgb2rgb(rgba: Tensor) -> Tensor:
    """RGBAImages.srgb2rgb (rfstudio/graphics/_images.py:287-300): sRGB -> linear on the colour channels."""
    c = rgba[..., :3]
    lin = torch.where(c <= 0.04045, c / 12.92, torch.pow((c.clamp_min(0.04045) + 0.055) / 1.055, 2.4))
    return torch.cat((lin, rgba[..., 3:]), dim=-1)


class GeoSplatter(nn.Module):
    def __init__(self, *, background_color: str = "random", resolution: int = 32, light_resolution: int = 512,
                 field: Optional[GaussianField] = None, gaussian_limits_hard: int = 1500000,
                 gaussian_limits_soft: int = 1000000, scale: float = 1.05, min_roughness: float = 0.1,
                 max_metallic: float = 1.0, smooth_type: str = "jitter", initial_guess: str = "hybrid",
                 fg_lut: Optional[Tensor] = None, n_streams: int = 4):
        super().__init__()
        if smooth_type != "jitter":
            raise NotImplementedError("smooth_type 'grad' / 'tv' are not mirrored (module docstring)")
        if initial_guess not in _GUESS:
            raise ValueError(initial_guess)
        self.background_color, self.resolution, self.light_resolution = background_color, resolution, light_resolution
        self.field = field if field is not None else GaussianField()
        self.gaussian_limits_hard, self.gaussian_limits_soft = gaussian_limits_hard, gaussian_limits_soft
        self.scale, self.min_roughness, self.max_metallic = scale, min_roughness, max_metallic
        self.smooth_type, self.initial_guess, self.n_streams = smooth_type, initial_guess, n_streams
        # __setup__ (geosplat.py:703-749)
        self.last_num_gaussians = 0
        self.exposure_params = nn.Parameter(torch.zeros(1))
        self.geometric_repr = FlexiCubes.from_resolution(resolution, scale=scale)
        self.deform_params = nn.Parameter(torch.zeros_like(self.geometric_repr.vertices))
        self.sdf_params = nn.Parameter(self.geometric_repr.sdf_values.clone())
        self.weight_params = nn.Parameter(torch.zeros(self.geometric_repr.indices.shape[0], 21))
        self.sdf_weight = self.occ_weight = self.light_weight = 0.0
        self.kd_grad_weight = self.ks_grad_weight = 0.0
        self.kd_regualr_perturb_std = self.ks_regualr_perturb_std = 0.0          # (sic: the reference's spelling)
        self.sample_method = "face"
        self.initial_guess_bias = nn.Parameter(torch.tensor(_GUESS[initial_guess]), requires_grad=False)
        self.cubemap = nn.Parameter(torch.full((6, light_resolution, light_resolution, 3), 0.5))
        self.register_buffer("fg_lut", fg_lut, persistent=False)

    @property
    def device(self) -> torch.device:
        return self.exposure_params.device

    @property
    def minimal_memory(self) -> bool:
        return self.last_num_gaussians > self.gaussian_limits_hard and self.training

    @property
    def save_memory(self) -> bool:
        return self.last_num_gaussians > self.gaussian_limits_soft and self.training

    def get_geometry(self) -> Tuple[TriangleMesh, Tensor]:
        """geosplat.py:751-769."""
        if self.geometric_repr.device != self.device:
            self.geometric_repr = self.geometric_repr.to(self.device)
        vertices = self.geometric_repr.vertices + self.deform_params.tanh() * (0.5 * self.scale / self.resolution)
        flexicubes = self.geometric_repr.replace(vertices=vertices, sdf_values=self.sdf_params,
                                                 alpha=self.weight_params[:, :8], beta=self.weight_params[:, 8:20],
                                                 gamma=self.weight_params[:, 20:])
        mesh, L_dev = flexicubes.dual_marching_cubes()
        reg_loss = torch.add(L_dev.mean() * 0.5 + self.weight_params[:, :20].abs().mean() * 0.1,
                             flexicubes.compute_entropy() * self.sdf_weight)
        return mesh, reg_loss

    def get_background_color(self) -> Tensor:
        """geosplat.py:771-778."""
        if self.background_color == "black":
            return torch.zeros(3)
        if self.background_color == "white":
            return torch.ones(3)
        if self.training:
            return torch.rand(3)
        return torch.tensor([0.1490, 0.1647, 0.2157])

    def get_envmap(self) -> Tuple[splitsum.EnvStack, Tensor]:
        """geosplat.py:780-785: white-balance regulariser + the split-sum prefilter of the cube map."""
        white = self.cubemap.mean(-1, keepdim=True)
        return splitsum.as_envstack(self.cubemap), (self.cubemap - white).abs().mean()

    def get_gsplat(self, sampling: str):
        """geosplat.py:787-831 -> (mesh, gsplat, attrs, regularisation, offsets)."""
        mesh, reg = self.get_geometry()
        gsplat = GSplatter(background_color=self.background_color, rasterize_mode="antialiased")
        kd_std = self.kd_regualr_perturb_std if self.smooth_type == "jitter" else 0
        ks_std = self.ks_regualr_perturb_std if self.smooth_type == "jitter" else 0
        if sampling == "face":
            self.last_num_gaussians = mesh.indices.shape[0] * 6
            splats, attrs, offsets = self.field.get_gaussians_from_face(
                mesh.vertices, mesh.indices, kd_std, ks_std, scale=self.scale, initial_guess=self.initial_guess_bias)
        elif sampling == "vertex":
            self.last_num_gaussians = mesh.vertices.shape[0]
            splats, attrs = self.field.get_gaussians_from_vertex(kd_std, ks_std, self.scale, mesh.vertices, mesh.indices,
                                                                 self.initial_guess_bias)
            offsets = None
        else:
            raise ValueError(sampling)
        gsplat.gaussians = splats
        if kd_std > 0 and self.kd_grad_weight > 0:
            reg = reg + self.kd_grad_weight * (attrs.kd_jitter - attrs.kd).abs().mean()
        if ks_std > 0 and self.ks_grad_weight > 0:
            reg = reg + self.ks_grad_weight * (attrs.ks_jitter - attrs.ks).abs().mean()
        if self.occ_weight > 0 and attrs.occ is not None:
            reg = reg + self.occ_weight * attrs.occ.abs().mean()
        attrs = RenderableAttrs(kd=attrs.kd, ks=attrs.ks, normals=attrs.normals, occ=attrs.occ)
        return mesh, gsplat, attrs, reg, offsets

    def render_report(self, inputs: Sequence[PinholeCamera], *, indices=None, gt_outputs=None
                      ) -> Tuple[List[Tensor], int, Tensor]:
        """geosplat.py:856-927 -> (tone-mapped linear RGBA images [H,W,4] per camera, #gaussians, regularisation).
        The per-view loop is one batched node (fused.splat_views)."""
        if self.fg_lut is None:
            raise RuntimeError("GeoSplatter needs the DFG table: pass fg_lut= (shade.load_fg_lut of the reference's "
                               "bsdf_256_256.bin, or shade.synthetic_fg_lut for synthetic runs)")
        mesh, gsplat, attrs, regularization, _ = self.get_gsplat(self.sample_method)
        envmap, light_reg = self.get_envmap()
        g = gsplat.gaussians
        images = splat_views(g.means, g.scales, g.quats, g.opacities, attrs.kd, attrs.ks, attrs.normals, list(inputs),
                             exposures=self.exposure_params.exp(), envmap=envmap, fg_lut=self.fg_lut,
                             min_roughness=self.min_roughness, max_metallic=self.max_metallic, n_streams=self.n_streams)
        return images, g.means.shape[0], regularization + light_reg * self.light_weight

    def training_loss(self, inputs: Sequence[PinholeCamera], gt_rgba: Sequence[Tensor], *, use_mask_loss: bool = True
                      ) -> Tuple[Tensor, dict]:
        """GeoSplatTrainer.step (geosplat_trainer.py:146-182): mean per-view loss against sRGB ground truth + the
        model's regularisation.  gt_rgba: [H,W,4] sRGB + mask per camera."""
        images, num_gaussians, reg_loss = self.render_report(inputs, gt_outputs=gt_rgba)
        losses = [view_loss(img, srgb2rgb(gt), use_mask_loss=use_mask_loss) for img, gt in zip(images, gt_rgba)]
        loss = sum(losses) / len(losses)
        metrics = {"loss": loss.detach(), "#gaussians": num_gaussians, "regularization": reg_loss.detach(),
                   "exposure": self.exposure_params.detach().mean().exp()}
        return loss + reg_loss, metrics
