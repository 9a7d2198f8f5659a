"""Split-sum prefilter parity on the GPU box.  Three-way: the C oracle (oracle/prefilter_oracle.c), the
REAL reference plugin built from the reference's sources (oracle/_ref, the strongest oracle available) and
this repository's CUDA path, through the plugin-compatible interface `splitsum.render_utils` and the
autograd wrappers.  fp32; tolerance 1e-4 relative to the value scale (north_star)."""
import numpy as np
import pytest
import torch

from geosplatting_b200 import splitsum as SS
from geosplatting_b200.shade import EnvStack
from oracle import prefilter as P
from oracle import shade as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, tol, name=""):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    scale = max(1e-6, float(np.abs(b).max()))
    err = float(np.abs(a - b).max())
    assert err <= tol * scale, (name, err, scale)


def _close_spec(a, b, name, tol_rgb=1e-3, tol_w=5e-3):
    """Specular prefilter outputs [6,R,R,4] = (sum w*rgb, sum w).  The reference's GGX weight
    alpha^2 / (pi * (1 - c^2 (1 - alpha^2))^2) cancels catastrophically near c = 1 (relative error of one weight
    ~ 2 * 6e-8 / alpha^2 in fp32: 1e-4 at roughness 0.185, 3e-3 at roughness 0.08), so two correct fp32
    builds (FMA contraction on/off) disagree per weight by that much.  What `specular_cubemap` returns is the
    RATIO rgb/wsum, in which the common weight error cancels; compare that, and wsum separately."""
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    ra, rb = a[..., :3] / a[..., 3:], b[..., :3] / b[..., 3:]
    err = float(np.abs(ra - rb).max())
    assert err <= tol_rgb * float(np.abs(rb).max()), (name, "rgb/wsum", err)
    werr = float((np.abs(a[..., 3] - b[..., 3]) / b[..., 3]).max())
    assert werr <= tol_w, (name, "wsum", werr)


def _cubemap(R, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.exp(torch.randn(6, R, R, 3, generator=g)).clamp_min(1e-2)


@pytest.fixture(scope="module")
def plugin():
    mod = P.load_reference_plugin()
    if mod is None:
        pytest.skip("oracle/_ref/rfstudio_render_utils.so not built (oracle/build_ref.sh needs /root/reference)")
    return mod


def test_oracle_matches_reference_plugin(plugin):
    """Pins the C restatement on the reference's own compiled kernels."""
    c16 = _cubemap(16, 0)
    _close(P.diffuse_fwd(c16.numpy()), plugin.diffuse_cubemap_fwd(c16.to(DEV)), 2e-5, "diffuse fwd")
    g16 = torch.randn(6, 16, 16, 3, generator=torch.Generator().manual_seed(1))
    _close(P.diffuse_bwd(g16.numpy()), plugin.diffuse_cubemap_bwd(c16.to(DEV), g16.to(DEV)), 2e-5, "diffuse bwd")
    for R, rough in ((32, 0.3), (16, 1.0), (64, 0.185)):
        ct = P.ndf_cutoff_costheta(rough)
        c = _cubemap(R, R)
        b_ref = plugin.specular_bounds(R, ct, 0)
        b_orc = P.specular_bounds(R, ct)
        assert np.abs(b_orc - b_ref.cpu().numpy()).max() <= 1, "bounds differ by more than one texel"
        _close_spec(P.specular_fwd(c.numpy(), b_orc, rough, ct),
                    plugin.specular_cubemap_fwd(c.to(DEV), b_ref, rough, ct), f"spec fwd {R}")
        g = torch.randn(6, R, R, 4, generator=torch.Generator().manual_seed(2))
        _close(P.specular_bwd(b_orc, g.numpy(), rough, ct),
               plugin.specular_cubemap_bwd(c.to(DEV), b_ref, g.to(DEV), rough, ct), 2e-3, f"spec bwd {R}")


@pytest.mark.parametrize("R,rough", [(16, 1.0), (32, 0.5), (64, 0.29), (64, 0.185)])
def test_product_matches_oracle(R, rough):
    ct = P.ndf_cutoff_costheta(rough)
    assert abs(ct - SS.ndf_cutoff_costheta(rough, 0.99)) < 1e-12
    c = _cubemap(R, 3 + R)
    b_orc = P.specular_bounds(R, ct)
    b = SS.render_utils.specular_bounds(R, ct, 0)
    assert np.abs(b.cpu().numpy() - b_orc).max() <= 1
    out = SS.render_utils.specular_cubemap_fwd(c.to(DEV), b, rough, ct)
    _close_spec(out, P.specular_fwd(c.numpy(), b_orc, rough, ct), "spec fwd")
    g = torch.randn(6, R, R, 4, generator=torch.Generator().manual_seed(5))
    gin = SS.render_utils.specular_cubemap_bwd(c.to(DEV), b, g.to(DEV), rough, ct)
    _close(gin, P.specular_bwd(b_orc, g.numpy(), rough, ct), 2e-3, "spec bwd")
    if R <= 32:
        _close(SS.render_utils.diffuse_cubemap_fwd(c.to(DEV)), P.diffuse_fwd(c.numpy()), 1e-4, "diffuse fwd")
        _close(SS.render_utils.diffuse_cubemap_bwd(c.to(DEV), g[..., :3].contiguous().to(DEV)),
               P.diffuse_bwd(g[..., :3].numpy()), 1e-4, "diffuse bwd")


@pytest.mark.parametrize("R,rough", [(128, 0.29), (256, 0.185), (512, 0.08)])
def test_product_vs_reference_plugin_and_fp64_at_full_size(plugin, R, rough):
    """The reference's own kernels at the resolutions GeoSplatter uses (light_resolution=512 -> 6 levels), with
    a double-precision evaluation of the same sums as yardstick: this library must be within 1e-4 of the fp64
    value and at least as close to it as the reference plugin is (whose fp32 GGX weight is ill-conditioned)."""
    ct = SS.ndf_cutoff_costheta(rough, 0.99)
    c = _cubemap(R, 11)
    cd = c.to(DEV)
    b_ref = plugin.specular_bounds(R, ct, 0)
    b = SS.render_utils.specular_bounds(R, ct, 0)
    assert float((b - b_ref).abs().max()) <= 1
    ours = SS.render_utils.specular_cubemap_fwd(cd, b_ref, rough, ct).cpu().numpy().reshape(-1, 4)
    theirs = plugin.specular_cubemap_fwd(cd, b_ref, rough, ct).cpu().numpy().reshape(-1, 4)
    stride = 97 if R >= 256 else 13
    truth = P.specular_fwd_f64(c.numpy(), b_ref.cpu().numpy(), rough, ct, stride)
    ok = truth[:, 4] == 0          # texels whose cone membership cannot flip between two fp32 builds
    assert ok.sum() > 200
    t_rgb = truth[:, :3] / truth[:, 3:4]
    scale = float(np.abs(t_rgb).max())

    def errs(x):
        x = x[::stride].astype(np.float64)
        return (float(np.abs(x[:, :3] / x[:, 3:] - t_rgb)[ok].max()) / scale,
                float((np.abs(x[:, 3] - truth[:, 3]) / truth[:, 3])[ok].max()))

    e_ours, e_ref = errs(ours), errs(theirs)
    assert e_ours[0] <= 1e-4 and e_ours[1] <= 1e-4, (e_ours, e_ref)
    assert e_ours[0] <= max(e_ref[0], 2e-5) and e_ours[1] <= max(e_ref[1], 2e-5), (e_ours, e_ref)
    # and the two fp32 implementations agree to the conditioning of the reference's formula
    tol = 8.0 * 6e-8 / rough ** 4 + 2e-4
    _close_spec(ours.reshape(6, R, R, 4), theirs.reshape(6, R, R, 4), "spec fwd", tol_rgb=4 * tol, tol_w=4 * tol)
    g = torch.randn(6, R, R, 4, generator=torch.Generator().manual_seed(7)).to(DEV)
    _close(SS.render_utils.specular_cubemap_bwd(cd, b, g, rough, ct),
           plugin.specular_cubemap_bwd(cd, b_ref, g, rough, ct), 4 * tol, "spec bwd")


def test_mip_chain_operator():
    c = _cubemap(32, 4)
    x = c.to(DEV).requires_grad_(True)
    down = SS.cubemap_mip(x)
    _close(down, S.cubemap_mip_fwd(c), 1e-6, "mip fwd")
    cot = torch.randn(6, 16, 16, 3, generator=torch.Generator().manual_seed(8))
    g, = torch.autograd.grad((down * cot.to(DEV)).sum(), x)
    _close(g, S.cubemap_mip_bwd(cot), 1e-5, "mip bwd")


def test_as_splitsum_and_envstack_agree_and_are_finite():
    """tests/graphics/test_splitsum.py:7-13 of the reference (isfinite smoke on a random 64^2 cube map), plus:
    the fused env-stack path equals the operator-by-operator path, forward and backward."""
    c = _cubemap(64, 9).to(DEV)
    x1 = c.clone().requires_grad_(True)
    base, packed, L, rmin, rmax = SS.as_splitsum(x1)
    assert torch.isfinite(base).all() and torch.isfinite(packed).all() and L == 3
    env1 = EnvStack.from_splitsum(base, packed, L, rmin, rmax)
    x2 = c.clone().requires_grad_(True)
    env2 = SS.as_envstack(x2)
    assert (env2.R0, env2.L, env2.Rb) == (64, 3, 16)
    _close(env2.data[:, :3], env1.data[:, :3], 1e-5, "stack")
    cot = torch.randn(env1.data.shape, generator=torch.Generator().manual_seed(10)).to(DEV)
    g1, = torch.autograd.grad((env1.data * cot).sum(), x1)
    cot2 = cot.clone()
    cot2[:, 3] = 0
    g2, = torch.autograd.grad((env2.data * cot2).sum(), x2)
    _close(g2, g1, 2e-4, "cubemap grad")
    assert torch.isfinite(g2).all()
