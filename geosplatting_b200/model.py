"""GeoSplatter stage 1 over this library: the caller that strings the rows of SURVEY.md section 8 together.

Mirrors (same names, argument meaning and defaults; host orchestration only -- every tensor op of substance is a call
into the C-ABI library through the operator mirrors of this package):

    GeoSplatter.__setup__ / get_geometry / get_background_color / get_envmap / get_gsplat / render_report
                                                     rfstudio/model/geosplat.py:676-927
    GeoSplatTrainer.step (loss assembly only)         rfstudio/trainer/geosplat_trainer.py:146-182  -> training_loss()

What it drives per step: FlexiCubes mesh + regulariser (flexicubes.py) -> vertex normals + MGAdaptor + kd / ks / z
hash-grid fields (field.py) -> split-sum prefilter (splitsum.py) -> the batch of views (fused.splat_views: shade,
rasterize, tone map) -> per-view loss (loss.py).  `smooth_type` 'grad' / 'tv' (extra render_rgb passes per view, off in
every shipped stage-1 recipe until their weights are raised) and the normal smoothness term run through
GSplatter.render_rgb like the reference's; their Sobel filter restates kornia.filters.spatial_gradient (third-party,
absent here: PARITY UNPINNED for that filter, |.| is taken so its sign conventions do not matter).
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


def spatial_gradient(x: Tensor) -> Tensor:
    """kornia.filters.spatial_gradient(x, mode='sobel', order=1, normalized=True): [B,C,H,W] -> [B,C,2,H,W] (d/dx, d/dy),
    3x3 Sobel kernels divided by 8, replicate padding."""
    B, Cn, H, W = x.shape
    kx = x.new_tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]]) / 8.0
    k = torch.stack((kx, kx.t()))[:, None]                                         # [2,1,3,3]
    y = torch.nn.functional.conv2d(torch.nn.functional.pad(x.reshape(B * Cn, 1, H, W), (1, 1, 1, 1), mode="replicate"), k)
    return y.view(B, Cn, 2, H, W)


def _edge_aware(rendered: Tensor, gt_rgb: Tensor) -> Tensor:
    """geosplat.py:885-888: |grad(render)| * exp(-|grad(gt)|), summed over channels, averaged."""
    a = spatial_gradient(rendered[None].permute(0, 3, 1, 2))[0].abs()
    b = (-spatial_gradient(gt_rgb[None].permute(0, 3, 1, 2))[0].abs()).exp()
    return (a * b).sum(1).mean()


def _tv(rendered: Tensor) -> Tensor:
    """geosplat.py:907-910."""
    return (rendered[1:, :] - rendered[:-1, :]).square().mean() + (rendered[:, 1:] - rendered[:, :-1]).square().mean()


class _FlexiWeights(torch.autograd.Function):
    """weights[:, :8], weights[:, 8:20], weights[:, 20:] and weights[:, :20].abs().mean() (geosplat.py:756-766) as ONE
    autograd node.  The values are those of the four expressions; the point is the backward: sliced one by one, autograd
    materialises a zero-filled [R^3, 21] tensor per slice, copies the slice's gradient into it and adds the four up --
    1.6 ms of elementwise passes per step on the 2.7 M x 21 weights of R = 140.  Here it is one concatenation plus one
    signed add."""

    @staticmethod
    def forward(ctx, w: Tensor):
        ctx.save_for_backward(w)
        return w[:, :8].contiguous(), w[:, 8:20].contiguous(), w[:, 20:].contiguous(), w[:, :20].abs().mean()

    @staticmethod
    def backward(ctx, g_alpha, g_beta, g_gamma, g_mean):
        (w,) = ctx.saved_tensors
        F = w.shape[0]
        parts = [g if g is not None else w.new_zeros(F, n) for g, n in ((g_alpha, 8), (g_beta, 12), (g_gamma, 1))]
        grad = torch.cat(parts, dim=1)
        if g_mean is not None:
            s = torch.sign(w)
            s[:, 20].zero_()
            grad.addcmul_(s, g_mean / (F * 20))
        return grad


class GeoSplatter(nn.Module):
    def __init__(self, *, background_color: str = "random", resolution: int = 32, light_resolution: int = 512,
                 field: Optional[GaussianField] = None, gaussian_limits_hard: int = 1500000,
                 gaussian_limits_soft: int = 1000000, scale: float = 1.05, min_roughness: float = 0.1,
                 max_metallic: float = 1.0, smooth_type: str = "jitter", initial_guess: str = "hybrid",
                 fg_lut: Optional[Tensor] = None, n_streams: int = 4):
        super().__init__()
        if smooth_type not in ("jitter", "grad", "tv"):
            raise ValueError(smooth_type)
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
        self.kd_grad_weight = self.ks_grad_weight = self.normal_grad_weight = 0.0
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

    def parameter_groups(self) -> dict:
        """The eight groups GeoSplatTrainer.setup builds one Adam optimiser for (geosplat_trainer.py:84-142; the first
        five are what `as_module(field_name=...)` exposes, geosplat.py:929-942)."""
        return {"deforms": [self.deform_params], "sdfs": [self.sdf_params], "weights": [self.weight_params],
                "light": [self.cubemap], "exposure": [self.exposure_params],
                "kd": list(self.field.kd_enc.parameters()), "ks": list(self.field.ks_enc.parameters()),
                "z": list(self.field.z_enc.parameters())}

    @torch.no_grad()
    def export_model(self, path) -> None:
        """geosplat.py:838-854: the attribute dictionary the reference's relighting tools load."""
        torch.save({"geom_scale": self.scale, "resolution": self.resolution, "min_roughness": self.min_roughness,
                    "max_metallic": self.max_metallic, "exposure": self.exposure_params, "cubemap": self.cubemap,
                    "deforms": self.deform_params, "weights": self.weight_params, "sdfs": self.sdf_params,
                    "ks_enc": self.field.ks_enc.state_dict(), "initial_guess": self.initial_guess_bias}, path)

    def get_geometry(self) -> Tuple[TriangleMesh, Tensor]:
        """geosplat.py:751-769."""
        if self.geometric_repr.device != self.device:
            self.geometric_repr = self.geometric_repr.to(self.device)
        vertices = self.geometric_repr.vertices + self.deform_params.tanh() * (0.5 * self.scale / self.resolution)
        alpha, beta, gamma, w_abs_mean = _FlexiWeights.apply(self.weight_params)
        flexicubes = self.geometric_repr.replace(vertices=vertices, sdf_values=self.sdf_params,
                                                 alpha=alpha, beta=beta, gamma=gamma)
        mesh, L_dev = flexicubes.dual_marching_cubes()
        reg_loss = torch.add(L_dev.mean() * 0.5 + w_abs_mean * 0.1, flexicubes.compute_entropy() * self.sdf_weight)
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
        regularization = regularization + self._smoothness(gsplat, attrs, list(inputs), gt_outputs)
        return images, g.means.shape[0], regularization + light_reg * self.light_weight

    def _smoothness(self, gsplat: GSplatter, attrs: RenderableAttrs, cameras, gt_outputs) -> Tensor:
        """geosplat.py:881-922: image-space smoothness of kd / ks / normals, rendered as plain colours (render_rgb)."""
        reg = torch.zeros((), device=self.device)
        B = len(cameras)
        passes = []
        if self.smooth_type in ("grad", "tv") and self.kd_grad_weight > 0:
            passes.append((attrs.kd, self.kd_grad_weight, self.smooth_type))
        if self.smooth_type in ("grad", "tv") and self.ks_grad_weight > 0:
            passes.append((torch.cat((torch.zeros_like(attrs.ks[..., :1]), attrs.ks), dim=-1), self.ks_grad_weight,
                           self.smooth_type))
        if self.normal_grad_weight > 0:
            passes.append((attrs.normals * 0.5 + 0.5, self.normal_grad_weight, "grad"))
        if not passes:
            return reg
        gsplat.training = self.training
        gt_rgb = None
        if any(kind == "grad" for _, _, kind in passes):
            if gt_outputs is None:
                raise ValueError("the edge-aware smoothness terms need gt_outputs")
            bg = self.get_background_color().to(self.device)
            gt_rgb = [gt[..., 3:] * gt[..., :3] + bg * (1 - gt[..., 3:]) for gt in gt_outputs]     # RGBAImages.blend
        for i, camera in enumerate(cameras):
            for colors, weight, kind in passes:
                gsplat.gaussians.replace_(colors=colors)
                rendered = gsplat.render_rgb(camera)
                term = _edge_aware(rendered, gt_rgb[i]) if kind == "grad" else _tv(rendered)
                reg = reg + term * weight / B
        return reg

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
