"""Host-only behaviour of model.GeoSplatter (no kernel runs): construction mirrors GeoSplatter.__setup__
(rfstudio/model/geosplat.py:703-749), parameter groups mirror the trainer's optimisers, export_model writes the
reference's attribute dictionary, the image-space filter helpers restate their definitions, and the product path
refuses to run without CUDA."""
import numpy as np
import pytest
import torch

from geosplatting_b200.model import GeoSplatter, _edge_aware, _tv, spatial_gradient, srgb2rgb
from geosplatting_b200.shade import synthetic_fg_lut


def _model(**kw):
    torch.manual_seed(0)
    return GeoSplatter(resolution=6, light_resolution=16, scale=1.05, fg_lut=synthetic_fg_lut(torch.device("cpu")), **kw)


def test_setup_shapes_and_initial_values():
    m = _model()
    V, F = 7 ** 3, 6 ** 3
    assert m.deform_params.shape == (V, 3) and float(m.deform_params.detach().abs().max()) == 0
    assert m.sdf_params.shape == (V, 1) and -0.1 <= float(m.sdf_params.min()) and float(m.sdf_params.max()) <= 0.9
    assert m.weight_params.shape == (F, 21) and m.cubemap.shape == (6, 16, 16, 3) and float(m.cubemap.mean()) == 0.5
    assert m.exposure_params.shape == (1,) and m.sample_method == "face" and m.last_num_gaussians == 0
    assert float(m.geometric_repr.vertices.abs().max()) == pytest.approx(1.05)
    assert not m.initial_guess_bias.requires_grad
    for guess, bias in (("outdoor", (0, 0)), ("diffuse", (0, -3)), ("hybrid", (-3, -3)), ("specular", (-3, 0)),
                        ("glossy", (-3, 0))):
        assert tuple(_model(initial_guess=guess).initial_guess_bias.tolist()) == bias      # geosplat.py:727-740
    with pytest.raises(ValueError):
        _model(initial_guess="shiny")
    with pytest.raises(ValueError):
        _model(smooth_type="sobel")


def test_background_colours_and_memory_switches():
    m = _model(background_color="white")
    assert torch.equal(m.get_background_color(), torch.ones(3))
    m = _model(background_color="black")
    assert torch.equal(m.get_background_color(), torch.zeros(3))
    m = _model()
    m.eval()
    assert torch.allclose(m.get_background_color(), torch.tensor([0.1490, 0.1647, 0.2157]))
    m.train()
    m.last_num_gaussians = 1_200_000
    assert m.save_memory and not m.minimal_memory
    m.last_num_gaussians = 1_600_000
    assert m.minimal_memory
    m.eval()
    assert not m.save_memory


def test_parameter_groups_cover_every_trainable_parameter(tmp_path):
    m = _model()
    groups = m.parameter_groups()
    assert set(groups) == {"deforms", "sdfs", "weights", "light", "exposure", "kd", "ks", "z"}
    grouped = {id(p) for ps in groups.values() for p in ps}
    trainable = {id(p) for p in m.parameters() if p.requires_grad}
    assert grouped == trainable
    m.export_model(tmp_path / "model.pt")
    d = torch.load(tmp_path / "model.pt", weights_only=False)
    assert set(d) == {"geom_scale", "resolution", "min_roughness", "max_metallic", "exposure", "cubemap", "deforms",
                      "weights", "sdfs", "ks_enc", "initial_guess"}
    assert d["resolution"] == 6 and torch.equal(d["sdfs"], m.sdf_params)


def test_image_space_helpers():
    x = torch.arange(20.0).view(1, 1, 4, 5)                     # ramp: d/dx = 1, d/dy = 5 in the interior
    g = spatial_gradient(x)
    assert g.shape == (1, 1, 2, 4, 5)
    assert torch.allclose(g[0, 0, 0, 1:3, 1:4], torch.ones(2, 3)) and torch.allclose(g[0, 0, 1, 1:3, 1:4], 5 * torch.ones(2, 3))
    flat = torch.full((8, 9, 3), 0.3)
    assert float(_tv(flat)) == 0 and float(_edge_aware(flat, torch.rand(8, 9, 3))) == 0
    img = torch.rand(8, 9, 3)
    assert float(_edge_aware(img, flat)) > float(_edge_aware(img, 50 * torch.rand(8, 9, 3)))   # gt edges excuse render edges
    c = torch.tensor([[[0.0, 0.04045, 0.5, 1.0]]])
    lin = srgb2rgb(torch.cat((c[..., :3], c[..., 3:]), -1))
    assert np.allclose(lin[0, 0, :3].numpy(), [0.0, 0.04045 / 12.92, ((0.5 + 0.055) / 1.055) ** 2.4], atol=1e-7)
    assert float(lin[0, 0, 3]) == 1.0                            # alpha untouched


def test_no_cpu_path_and_missing_lut():
    m = _model()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.get_geometry()
    m2 = GeoSplatter(resolution=4, light_resolution=16)
    with pytest.raises(RuntimeError, match="DFG table"):
        m2.render_report([])


def test_flexicubes_weight_node_equals_the_sliced_expressions():
    """model._FlexiWeights (one autograd node for the three weight slices and the |w| regulariser of geosplat.py:756-766)
    against the plain sliced expressions: same values, same gradient, also when a slice is unused."""
    import torch

    from geosplatting_b200.model import _FlexiWeights
    gen = torch.Generator().manual_seed(3)
    w0 = torch.randn(57, 21, generator=gen)
    ca, cb, cg = torch.randn(57, 8, generator=gen), torch.randn(57, 12, generator=gen), torch.randn(57, 1, generator=gen)
    for use_gamma in (True, False):
        w = w0.clone().requires_grad_(True)
        a, b, g, m = _FlexiWeights.apply(w)
        loss = (a * ca).sum() + (b * cb).sum() + m * 0.1 + ((g * cg).sum() if use_gamma else 0.0)
        (gw,) = torch.autograd.grad(loss, w)
        r = w0.clone().requires_grad_(True)
        ref = (r[:, :8] * ca).sum() + (r[:, 8:20] * cb).sum() + r[:, :20].abs().mean() * 0.1 + \
            ((r[:, 20:] * cg).sum() if use_gamma else 0.0)
        (gr,) = torch.autograd.grad(ref, r)
        assert torch.equal(a, w0[:, :8]) and torch.equal(b, w0[:, 8:20]) and torch.equal(g, w0[:, 20:])
        assert float((loss - ref).abs()) <= 1e-5 * float(ref.abs())
        assert float((gw - gr).abs().max()) <= 1e-6 * float(gr.abs().max())
