"""GeoSplatter stage 1 end to end on the GPU (geosplatting_b200/model.py): SDF grid -> FlexiCubes mesh -> MGAdaptor
Gaussians -> hash-grid materials -> split-sum shade -> rasterize -> loss, and back to every parameter group the
reference's trainer optimises (rfstudio/trainer/geosplat_trainer.py:66-142)."""
import pytest
import torch

from geosplatting_b200 import scenes
from geosplatting_b200.model import GeoSplatter, srgb2rgb
from geosplatting_b200.shade import synthetic_fg_lut
from geosplatting_b200.splat import RenderableAttrs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rgb2srgb(rgba):
    c = rgba[..., :3]
    s = torch.where(c <= 0.0031308, c * 12.92, torch.clamp(c, min=0.0031308).pow(1.0 / 2.4) * 1.055 - 0.055)
    return torch.cat((s, rgba[..., 3:]), -1)


def make_model(radius, seed, light=0.5):
    torch.manual_seed(seed)
    m = GeoSplatter(resolution=16, light_resolution=64, scale=0.9, fg_lut=synthetic_fg_lut(torch.device("cpu")),
                    background_color="white").to(DEV)
    gv = m.geometric_repr.vertices.to(DEV)
    with torch.no_grad():
        m.sdf_params.copy_(gv.norm(dim=-1, keepdim=True) - radius)
        m.cubemap.fill_(light)
    return m


def test_render_report_matches_the_staged_operators_and_reaches_every_parameter():
    m = make_model(0.55, 0)
    m.sdf_weight, m.light_weight = 0.2, 2e-3
    cams = scenes.orbit_cameras(2, 64, 64, seed=3)
    images, n, reg = m.render_report(cams)
    assert len(images) == 2 and images[0].shape == (64, 64, 4) and n == m.last_num_gaussians > 1000
    assert all(bool(torch.isfinite(i).all()) for i in images) and float(images[0][..., 3].detach().max()) > 0.9
    # the same view through the stage-by-stage operators (RenderableAttrs.splat, geosplat.py:53-132)
    _, gsplat, attrs, _, _ = m.get_gsplat("face")
    env, _ = m.get_envmap()
    staged = RenderableAttrs(kd=attrs.kd, ks=attrs.ks, normals=attrs.normals).splat(
        gsplat, [cams[0]], exposure=m.exposure_params.exp(), envmap=env, fg_lut=m.fg_lut, min_roughness=m.min_roughness,
        max_metallic=m.max_metallic, fused=False)
    assert float((staged - images[0]).abs().max()) <= 1e-5      # the prefilter's weight sums are atomics: not bit-stable
    gt = [rgb2srgb(torch.rand(64, 64, 4, device=DEV)) for _ in cams]
    loss, metrics = m.training_loss(cams, gt)
    loss.backward()
    groups = {"sdf": m.sdf_params, "deform": m.deform_params, "weights": m.weight_params, "light": m.cubemap,
              "exposure": m.exposure_params, "kd": next(m.field.kd_enc.parameters()),
              "ks": next(m.field.ks_enc.parameters()), "z": next(m.field.z_enc.parameters())}
    for name, p in groups.items():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()) and float(p.grad.abs().sum()) > 0, name
    assert metrics["#gaussians"] == n and float(metrics["loss"]) > 0
    assert float((srgb2rgb(gt[0]) - gt[0]).abs().max()) > 0           # the ground truth went through the sRGB decode


def test_vertex_sampling_warmup_path_and_jitter_regularisers():
    """The first `vertex_sample_warmup` steps sample one disc per vertex (geosplat_trainer.py:215-216); the jitter
    regularisers (geosplat.py:824-827) switch on with their weights."""
    m = make_model(0.5, 1)
    m.sample_method = "vertex"
    m.kd_regualr_perturb_std = m.ks_regualr_perturb_std = 0.1
    m.kd_grad_weight, m.ks_grad_weight = 0.03, 0.001
    cams = scenes.orbit_cameras(1, 48, 48, seed=4)
    images, n, reg = m.render_report(cams)
    assert n == m.last_num_gaussians and bool(torch.isfinite(images[0]).all()) and float(images[0][..., 3].max()) > 0.5
    m.kd_grad_weight = m.ks_grad_weight = 0.0
    _, _, reg0 = m.render_report(cams)
    assert float(reg) > float(reg0)


def test_a_few_adam_steps_reduce_the_loss():
    """Fit a darker, smaller sphere to images of a brighter, larger one with the trainer's optimiser groups: the loss
    goes down and the exposure / light move up -- gradients have the right sign through the whole chain."""
    cams = scenes.orbit_cameras(4, 64, 64, seed=5)
    with torch.no_grad():
        target = make_model(0.62, 2, light=0.9)
        target.exposure_params.fill_(0.3)
        gt = [rgb2srgb(i.clamp(0, 1)) for i in target.render_report(cams)[0]]
    m = make_model(0.5, 3, light=0.4)
    m.train()
    m.cubemap.register_hook(lambda g: g * 64)                          # geosplat_trainer.py:69
    opt = torch.optim.Adam([
        {"params": [m.sdf_params, m.deform_params, m.weight_params], "lr": 3e-3},
        {"params": list(m.field.parameters()), "lr": 1e-2},
        {"params": [m.cubemap], "lr": 1e-2}, {"params": [m.exposure_params], "lr": 5e-3}], eps=1e-15)
    history = []
    for step in range(30):
        m.sdf_weight = 0.2
        m.light_weight = 2e-3
        opt.zero_grad(set_to_none=True)
        loss, metrics = m.training_loss(cams, gt)
        loss.backward()
        opt.step()
        history.append(float(metrics["loss"]))
    assert all(h == h for h in history)
    assert sum(history[-5:]) / 5 < 0.9 * sum(history[:5]) / 5, history
    assert float(m.exposure_params) > 0 and float(m.cubemap.mean()) > 0.4


@pytest.mark.parametrize("smooth_type", ["tv", "grad"])
def test_image_space_smoothness_terms(smooth_type):
    """smooth_type 'tv' / 'grad' and the normal term (geosplat.py:881-922) render kd / ks / normals as plain colours
    through GSplatter.render_rgb: the terms add to the regularisation and reach the kd field and the geometry."""
    m = make_model(0.55, 4)
    m.smooth_type = smooth_type
    cams = scenes.orbit_cameras(2, 64, 64, seed=6)
    gt = [torch.rand(64, 64, 4, device=DEV) for _ in cams]
    _, _, reg0 = m.render_report(cams, gt_outputs=gt)
    m.kd_grad_weight, m.ks_grad_weight, m.normal_grad_weight = 0.03, 0.001, 0.5
    _, _, reg = m.render_report(cams, gt_outputs=gt)
    assert float(reg.detach()) > float(reg0.detach())
    grads = torch.autograd.grad(reg, [next(m.field.kd_enc.parameters()), m.sdf_params])
    assert all(bool(torch.isfinite(g).all()) and float(g.abs().sum()) > 0 for g in grads)


def test_gsplatter_render_rgb_and_depth():
    """GSplatter.render_rgb (gsplat.py:187-282) = render_rgba blended over the background; render_depth (:112-186)
    = expected depth + alpha, inside the scene's depth range where covered."""
    m = make_model(0.55, 5)
    cam = scenes.orbit_cameras(1, 64, 64, seed=7)[0]
    _, gsplat, attrs, _, _ = m.get_gsplat("face")
    gsplat.gaussians.replace_(colors=attrs.kd.detach())
    gsplat.training = False
    rgba = gsplat.render_rgba(cam)
    rgb = gsplat.render_rgb(cam)
    bg = gsplat.get_background_color().to(DEV)
    assert rgb.shape == (64, 64, 3) and float((rgb - (rgba[..., :3] + (1 - rgba[..., 3:]) * bg)).abs().max()) <= 1e-6
    d = gsplat.render_depth(cam)
    assert d.shape == (64, 64, 2) and float((d[..., 1] - rgba[..., 3]).abs().max()) <= 1e-6
    covered = d[..., 1] > 0.99
    dist = float(torch.tensor(cam.position).norm())
    assert bool(covered.any()) and float(d[..., 0][covered].min()) > dist - 0.7 and float(d[..., 0][covered].max()) < dist + 0.1
