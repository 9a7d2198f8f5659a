"""The reference-facing operators of the hot path, over the C-ABI kernels.

Mirrors (same names, argument meaning, side effects and errors):
    Splats                      rfstudio/graphics/_splats.py:17-32       (fields only; densify/cull are out of scope)
    GSplatter.render_rgba       rfstudio/model/gsplat.py:284-358
    RenderableAttrs.splat       rfstudio/model/geosplat.py:53-132
    GeoSplatter.render_report   rfstudio/model/geosplat.py:856-879 (the per-view loop) -> render_views()
Cameras are `scenes.PinholeCamera` (same conventions as rfstudio/graphics/_cameras.py: OpenGL c2w,
`view_matrix` / `intrinsic_matrix` properties), passed one per view as the reference does (`cameras.shape == (1,)`).
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import List, Optional, Sequence, Tuple, Union

import torch
from torch import Tensor

from .fused import splat_view
from .mgadapter import tone_mapping_naive
from .rasterization import Projected, rasterization_begin, rasterization_end
from .scenes import PinholeCamera, to_pinhole
from .shade import EnvStack, get_fg_lut, shade


@dataclass
class Splats:
    means: Tensor       # [N,3]
    scales: Tensor      # [N,3] log-scales
    quats: Tensor       # [N,4] wxyz, un-normalised
    colors: Tensor      # [N,3]
    opacities: Tensor   # [N,1] logits

    @property
    def shape(self) -> Tuple[int]:
        return (self.means.shape[0],)

    def replace_(self, **kw) -> "Splats":
        for k, v in kw.items():
            setattr(self, k, v)
        return self

    def __getitem__(self, mask) -> "Splats":
        if mask is Ellipsis:
            return replace(self)
        return Splats(self.means[mask], self.scales[mask], self.quats[mask], self.colors[mask], self.opacities[mask])


def _single_camera(cameras) -> PinholeCamera:
    """`Cameras[1]` of the reference (rfstudio Cameras with shape (1,), gsplat.py:293 / geosplat.py:66), a PinholeCamera or
    a one-element sequence of either."""
    if isinstance(cameras, PinholeCamera):
        return cameras
    cams = to_pinhole(cameras)
    assert len(cams) == 1  # gsplat.py:293 / geosplat.py:66
    return cams[0]


@dataclass
class GSplatter:
    """rfstudio/model/gsplat.py:20-58; only what GeoSplatter uses: sh_degree=0 (colours raw), block_width=16,
    rasterize_mode in {'classic','antialiased'}."""

    gaussians: Optional[Splats] = None
    sh_degree: int = 0
    block_width: int = 16
    background_color: str = "random"
    rasterize_mode: str = "classic"
    training: bool = True

    def begin_render(self, inputs) -> Projected:
        """Geometry half of `render_rgba`: activations, projection and the intersection count (the colours are not
        needed yet, so RenderableAttrs.splat shades while the host waits for the count)."""
        camera = _single_camera(inputs)
        if self.rasterize_mode not in ("antialiased", "classic"):
            raise ValueError(f"Unknown rasterize_mode: {self.rasterize_mode}")
        if self.sh_degree != 0:
            raise NotImplementedError("geosplatting_b200.GSplatter: sh_degree must be 0 (GeoSplatter, geosplat.py:794)")
        g = self.gaussians
        return rasterization_begin(
            means=g.means, quats=g.quats, scales=g.scales.exp(), viewmats=torch.from_numpy(camera.view_matrix)[None],
            Ks=torch.from_numpy(camera.intrinsic_matrix)[None], width=camera.width, height=camera.height,
            near_plane=0.01, far_plane=1e10, rasterize_mode=self.rasterize_mode)

    def finish_render(self, st: Projected) -> Tensor:
        g = self.gaussians
        render, alpha, _ = rasterization_end(st, torch.sigmoid(g.opacities).squeeze(-1), g.colors,
                                             render_mode="RGB", tile_size=self.block_width)
        return torch.cat((render[..., :3], alpha), dim=-1).squeeze(0)

    def render_rgba(self, inputs) -> Tensor:
        """-> [H,W,4] (linear RGB + alpha), differentiable w.r.t. the Gaussians (gsplat.py:284-358: `rasterization`
        with packed=True, tile 16, near 0.01, far 1e10, render_mode 'RGB', sh_degree None, dense gradients)."""
        return self.finish_render(self.begin_render(inputs))

    def get_background_color(self) -> Tensor:
        """gsplat.py:100-107."""
        if self.background_color == "black":
            return torch.zeros(3)
        if self.background_color == "white":
            return torch.ones(3)
        if self.training:
            return torch.rand(3)
        return torch.tensor([0.1490, 0.1647, 0.2157])

    def render_rgb(self, inputs) -> Tensor:
        """-> [H,W,3]: the colours composited over `get_background_color()` (gsplat.py:187-282; the blend is the
        reference's `render + (1 - alpha) * background`, done on the image, not inside the rasterizer)."""
        st = self.begin_render(inputs)
        g = self.gaussians
        render, alpha, _ = rasterization_end(st, torch.sigmoid(g.opacities).squeeze(-1), g.colors, render_mode="RGB",
                                             tile_size=self.block_width)
        background = self.get_background_color().to(render.device)
        return (render[..., :3] + (1 - alpha) * background).squeeze(0)

    def render_depth(self, inputs) -> Tensor:
        """-> [H,W,2]: expected depth (accumulated depth / alpha) and alpha (gsplat.py:112-186, render_mode 'ED')."""
        st = self.begin_render(inputs)
        g = self.gaussians
        render, alpha, _ = rasterization_end(st, torch.sigmoid(g.opacities).squeeze(-1), g.colors.detach(),
                                             render_mode="ED", tile_size=self.block_width)
        return torch.cat((render, alpha), dim=-1).squeeze(0)


@dataclass
class RenderableAttrs:
    """rfstudio/model/geosplat.py:43-51.  `occ` and the jittered copies feed regularisers / later stages, not `splat`."""

    kd: Tensor        # [N,3]
    ks: Tensor        # [N,2]
    normals: Tensor   # [N,3]
    occ: Optional[Tensor] = None
    kd_jitter: Optional[Tensor] = None
    ks_jitter: Optional[Tensor] = None

    def splat(self, gsplat: GSplatter, cameras, *, exposure: Tensor, envmap, min_roughness: float, max_metallic: float,
              mode: str = "pbr", tone_type: str = "naive", culling: bool = False, fg_lut: Optional[Tensor] = None,
              fused: bool = True) -> Tensor:
        """geosplat.py:53-65, same keyword arguments.  `cameras`: the reference's `Cameras[1]` (any object with its tensor
        fields) or a PinholeCamera; `envmap`: the reference's `TextureSplitSum` (base / mipmaps / num_mipmaps ...) or this
        library's EnvStack (splitsum.as_envstack(cubemap): no quad-tree pack per step); `fg_lut`: defaults to
        `_get_fg_lut(256, device)`, i.e. the reference's asset (shade.get_fg_lut).

        With culling=False and tone_type 'naive' / 'none' (what GeoSplatter.render_report uses) the whole view is
        one autograd node over the C-ABI kernels (fused.splat_view); `fused=False`, culling or 'aces' take the
        stage-by-stage operators below -- same kernels, same results, more host work."""
        camera = _single_camera(cameras)
        envmap = EnvStack.coerce(envmap)
        if fg_lut is None:
            fg_lut = get_fg_lut(256, self.kd.device)
        if fused and not culling and tone_type in ("naive", "none") and exposure.numel() == 1:
            if gsplat.sh_degree != 0:
                raise NotImplementedError("geosplatting_b200.GSplatter: sh_degree must be 0 (GeoSplatter, geosplat.py:794)")
            g = gsplat.gaussians
            return splat_view(g.means, g.scales, g.quats, g.opacities, self.kd, self.ks, self.normals, camera,
                              exposure=exposure, envmap=envmap, fg_lut=fg_lut, min_roughness=min_roughness,
                              max_metallic=max_metallic, mode=mode, tone_type=tone_type,
                              rasterize_mode=gsplat.rasterize_mode)
        cam_pos = camera.c2w[:, 3]
        if culling:
            with torch.no_grad():
                d = torch.as_tensor(cam_pos, device=self.normals.device) - gsplat.gaussians.means
                wo = torch.nn.functional.normalize(d, dim=-1)
                mask = (self.normals * wo).sum(-1) > 0.0
            if not mask.any():
                raise ValueError("No valid splat found.")
        else:
            mask = ...
        origin = gsplat.gaussians
        gsplat.gaussians = origin[mask]
        try:
            normals, kd, ks = (self.normals, self.kd, self.ks) if mask is Ellipsis else \
                (self.normals[mask], self.kd[mask], self.ks[mask])
            st = gsplat.begin_render(camera)          # projection + intersection count are queued first ...
            colors = shade(gsplat.gaussians.means, normals, kd, ks, [float(x) for x in cam_pos], envmap, fg_lut,
                           min_roughness=min_roughness, max_metallic=max_metallic, mode=mode)
            gsplat.gaussians.replace_(colors=colors)
            rgba = gsplat.finish_render(st)           # ... so the shade covers the host's wait for the count
            if tone_type == "none":
                out = torch.cat((rgba[..., :3] * exposure, rgba[..., 3:]), dim=-1)
            elif tone_type == "naive":
                out = tone_mapping_naive(rgba, exposure)
            elif tone_type == "aces":
                rgb = rgba[..., :3] * exposure
                out = torch.cat(((rgb * (2.51 * rgb + 0.03)) / (rgb * (2.43 * rgb + 0.59) + 0.14), rgba[..., 3:]), dim=-1)
            else:
                raise ValueError(tone_type)
        finally:
            gsplat.gaussians = origin   # geosplat.py:131 restores the caller's Gaussians
        return out


def render_views(attrs: RenderableAttrs, gsplat: GSplatter, cameras: Sequence[PinholeCamera], *, exposures: Tensor,
                 envmap: EnvStack, fg_lut: Tensor, min_roughness: float = 0.1, max_metallic: float = 1.0,
                 mode: str = "pbr") -> List[Tensor]:
    """The per-view loop of GeoSplatter.render_report (geosplat.py:869-879)."""
    return [attrs.splat(gsplat, [cam], exposure=exposures[i:i + 1], envmap=envmap, fg_lut=fg_lut,
                        min_roughness=min_roughness, max_metallic=max_metallic, mode=mode)
            for i, cam in enumerate(cameras)]
