"""TEST INFRASTRUCTURE: run the REFERENCE'S OWN, UNMODIFIED caller code over this library's drop-ins and compare with
this library's own operators on identical parameters.  Executed as a subprocess by tests/test_reference_caller_cpu.py
(the sys.modules surgery below must not leak into the pytest process); prints one JSON line.

    python tests/ref_caller_main.py            # needs /root/reference (this container only -- not the GPU box)

Wiring = INTEGRATION.md section 1, literally:
    gsplat.rasterization                       -> geosplatting_b200.rasterization                      (seam B1)
    nvdiffrast.torch.texture                   -> geosplatting_b200.shade.texture                      (seam B2)
    _splitsum._wrap._cached_plugin             -> geosplatting_b200.splitsum.render_utils              (seam B3)
    tinycudann.Encoding                        -> geosplatting_b200.encoding.TcnnEncoding              (fields, 8f rank 1)
and then the reference's own code runs: GeoSplatter.render_report (rfstudio/model/geosplat.py:856-927) with
get_gsplat / get_geometry / get_envmap, FlexiCubes.dual_marching_cubes, GaussianField.get_gaussians_from_face,
MGAdapter.make, TextureCubeMap.as_splitsum + _CubeMapMip, RenderableAttrs.splat, TextureSplitSum.sample,
GSplatter.render_rgba (rfstudio/model/gsplat.py:284-358), _tone_mapping_naive, _get_fg_lut (the real asset).

No GPU here: the kernels are the shipped .cu sources compiled for the host (tests/emu, SIMT mode), reached through the
real host modules and the C ABI.  Work-arounds that do NOT touch reference code:
  * Python 3.12 rejects the reference's dataclass instance defaults (geosplat.py:482-518, :683): dataclasses'
    "set __hash__ to None" action is switched off while rfstudio is imported;
  * rfstudio/model/__init__.py imports every model (OptiX stages included): the package is pre-registered with its
    path so that only gsplat.py / geosplat.py load;
  * the reference asserts `cubemap.is_cuda` in _wrap.diffuse_cubemap / specular_cubemap (:96, :146): the two wrappers
    are re-stated here around the reference's OWN autograd Functions and its own __ndfBounds, minus that assert;
  * absent third-party modules (open3d, kornia, sklearn, ...) are stubbed by scripts/ref_shim.py.
"""
import dataclasses
import importlib
import importlib.machinery
import json
import os
import sys
import types
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import torch  # noqa: E402

REF = "/root/reference"
LUT_PATH = f"{REF}/rfstudio/assets/geometry/pbr/bsdf_256_256.bin"


class _Stream:
    cuda_stream = 0

    def synchronize(self):
        pass

    def wait_event(self, e):
        pass

    def wait_stream(self, s):
        pass


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def synchronize(self):
        pass


def route_library_to_host():
    """Every host module of the package -> the host build of the kernels (what tests/emu/patch.route does per test)."""
    import geosplatting_b200 as gb
    from geosplatting_b200 import _lib
    from tests.emu import build as emu
    from tests.emu.patch import host_ptr
    so = _lib.declare(emu.build(*emu.all_kernel_files(), simt=True))
    _lib.load = lambda: so
    mods = [importlib.import_module(f"geosplatting_b200.{n}") for n in
            ("rasterization", "shade", "splitsum", "encoding", "mgadapter", "flexicubes", "fused", "loss", "splat")]
    for m in mods:
        for name, val in (("ptr", host_ptr), ("stream_ptr", lambda dev: None), ("_require_cuda", lambda t, what: None)):
            if hasattr(m, name):
                setattr(m, name, val)
    torch.cuda.current_stream = lambda device=None: _Stream()
    torch.cuda.current_device = lambda: 0
    torch.cuda.Event = _Event
    rz = importlib.import_module("geosplatting_b200.rasterization")
    rz._total_slot = lambda device: torch.zeros(1, dtype=torch.int64)
    fu = importlib.import_module("geosplatting_b200.fused")
    fu._total_slot = rz._total_slot
    fu._pinned_counts = lambda n: torch.zeros(max(n, 64), dtype=torch.int64)     # pinned memory needs a CUDA runtime
    importlib.import_module("geosplatting_b200.splitsum")._bounds_device = lambda index: torch.device("cpu")
    return gb


def import_reference(gb):
    import ref_shim
    from geosplatting_b200 import encoding, shade, splitsum
    ref_shim.install(shade.texture, gb.rasterization)                       # B2, B1
    sys.modules["tinycudann"].Encoding = encoding.TcnnEncoding
    sys.modules["tinycudann"].free_temporary_memory = lambda: None
    saved = dict(dataclasses._hash_action)
    for k, v in saved.items():
        if v is dataclasses._hash_set_none:
            dataclasses._hash_action[k] = None
    try:
        del sys.modules["rfstudio.graphics._mesh._splitsum"]                # ref_shim's inert stand-in: use the real one
        import rfstudio  # noqa: F401
        import rfstudio.graphics._mesh._splitsum._wrap as wrap
        wrap._cached_plugin = splitsum.render_utils                         # B3
        # the package's two public wrappers minus `assert cubemap.is_cuda` (see module docstring)
        nd = wrap.__dict__["__ndfBounds"]
        cache = {}

        def diffuse_cubemap(cubemap, use_python=False):
            return wrap._diffuse_cubemap_func.apply(cubemap)

        def specular_cubemap(cubemap, roughness, cutoff=0.99, use_python=False):
            key = (cubemap.shape[1], roughness, cutoff, cubemap.device.index)
            if key not in cache:
                cache[key] = nd(*key)
            out = wrap._specular_cubemap.apply(cubemap, roughness, *cache[key])
            return out[..., 0:3] / out[..., 3:]

        pk = sys.modules["rfstudio.graphics._mesh._splitsum"]
        pk.diffuse_cubemap, pk.specular_cubemap = diffuse_cubemap, specular_cubemap
        import rfstudio.graphics._mesh._texture as RT                       # _texture.py:22-25 imported them by name
        RT._diffuse_prefilter_cubemap, RT._specular_prefilter_cubemap = diffuse_cubemap, specular_cubemap
        pkg = types.ModuleType("rfstudio.model")
        pkg.__path__ = [f"{REF}/rfstudio/model"]
        pkg.__spec__ = importlib.machinery.ModuleSpec("rfstudio.model", None, is_package=True)
        sys.modules["rfstudio.model"] = pkg
        import rfstudio.model.geosplat as GEO
    finally:
        dataclasses._hash_action.update(saved)
    return GEO


def setup_recursively(x):
    """rfstudio/engine/task.py:243-250 (`_setup`): fields first, then the component itself."""
    from rfstudio.nn import Module as RModule
    if dataclasses.is_dataclass(x):
        for f in dataclasses.fields(x):
            setup_recursively(getattr(x, f.name, None))
    if isinstance(x, RModule):
        x.__setup__()


def rel(a, b):
    a, b = a.detach().double().reshape(-1), b.detach().double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    R, LIGHT, LOG2T, W, H = 12, 64, 14, 48, 40
    gb = route_library_to_host()
    GEO = import_reference(gb)
    import rfstudio.graphics as G
    from rfstudio.model.components.encoding import HashEncoding as RefHash
    from rfstudio.nn import MLP as RefMLP

    from geosplatting_b200 import scenes, shade
    from geosplatting_b200 import encoding as E
    from geosplatting_b200.field import GaussianField
    from geosplatting_b200.model import GeoSplatter

    def ref_enc(layers, act):
        return RefHash(mlp=RefMLP(layers=layers, activation=act, bias=False, initialization="kaiming-uniform"),
                       grad_scaling=16.0, max_res=4096, log2_hashmap_size=LOG2T)

    torch.manual_seed(0)
    ref = GEO.GeoSplatter(resolution=R, light_resolution=LIGHT, scale=0.9, background_color="black",
                          field=GEO.GaussianField(kd_enc=ref_enc([-1, 32, 32, 3], "sigmoid"),
                                                  ks_enc=ref_enc([-1, 32, 2], "none"), z_enc=ref_enc([-1, 32, 1], "none")))
    setup_recursively(ref)
    ref.to(torch.device("cpu"))
    ref.train()
    gv = ref.geometric_repr.vertices
    gen = torch.Generator().manual_seed(3)
    with torch.no_grad():
        ref.sdf_params.copy_(gv.norm(dim=-1, keepdim=True) - 0.55 + 0.05 * torch.sin(6.0 * gv[:, :1]))
        ref.deform_params.copy_(0.3 * torch.randn(gv.shape, generator=gen))
        ref.weight_params.copy_(0.2 * torch.randn(ref.weight_params.shape, generator=gen))
        ref.cubemap.copy_(torch.rand(ref.cubemap.shape, generator=gen) + 0.2)
        ref.exposure_params.fill_(0.15)
        for e in (ref.field.kd_enc, ref.field.ks_enc, ref.field.z_enc):
            e.encoder.params.mul_(300.0)                                     # off the 1e-3 init: the fields must matter
    cl = scenes.orbit_cameras(2, W, H, seed=7)
    cams = G.Cameras(c2w=torch.from_numpy(np.stack([c.c2w for c in cl])), fx=torch.tensor([c.fx for c in cl]),
                     fy=torch.tensor([c.fy for c in cl]), cx=torch.tensor([c.cx for c in cl]),
                     cy=torch.tensor([c.cy for c in cl]), width=torch.tensor([W] * 2), height=torch.tensor([H] * 2),
                     near=torch.tensor([0.01] * 2), far=torch.tensor([100.0] * 2))
    cots = [torch.randn(H, W, 4, generator=gen) for _ in cl]

    # ---------------------------------------------------- the reference's model over the drop-ins
    images, n_ref, reg = ref.render_report(cams, indices=None, gt_outputs=None)
    ref_imgs = [im for im in images]
    loss = sum((im * c).sum() for im, c in zip(ref_imgs, cots)) + reg
    ref_params = {"sdfs": ref.sdf_params, "deforms": ref.deform_params, "weights": ref.weight_params,
                  "light": ref.cubemap, "exposure": ref.exposure_params,
                  "kd_table": ref.field.kd_enc.encoder.params, "ks_table": ref.field.ks_enc.encoder.params,
                  "z_table": ref.field.z_enc.encoder.params,
                  "kd_mlp0": ref.field.kd_enc.mlp.nn_layers[0].weight, "ks_mlp1": ref.field.ks_enc.mlp.nn_layers[1].weight}
    ref_grads = dict(zip(ref_params, torch.autograd.grad(loss, list(ref_params.values()))))

    # ---------------------------------------------------- this library's model on the same parameters
    def own_enc(layers, act):
        return E.HashEncoding(E.MLP(layers, activation=act), grad_scaling=16.0, max_res=4096, log2_hashmap_size=LOG2T)

    own = GeoSplatter(resolution=R, light_resolution=LIGHT, scale=0.9, background_color="black", n_streams=1,
                      field=GaussianField(own_enc([32, 32, 32, 3], "sigmoid"), own_enc([32, 32, 2], "none"),
                                          own_enc([32, 32, 1], "none")),
                      fg_lut=shade.load_fg_lut(LUT_PATH, "cpu"))
    own.train()
    # the exported state of the reference's ks field loads into this library's module key for key
    missing = own.field.ks_enc.load_state_dict(ref.field.ks_enc.state_dict(), strict=True)
    keys_equal = sorted(own.field.ks_enc.state_dict().keys()) == sorted(ref.field.ks_enc.state_dict().keys())
    with torch.no_grad():
        for a, b in ((own.sdf_params, ref.sdf_params), (own.deform_params, ref.deform_params),
                     (own.weight_params, ref.weight_params), (own.cubemap, ref.cubemap),
                     (own.exposure_params, ref.exposure_params)):
            a.copy_(b)
        own.field.kd_enc.load_state_dict(ref.field.kd_enc.state_dict())
        own.field.z_enc.load_state_dict(ref.field.z_enc.state_dict())
    own_imgs, n_own, own_reg = own.render_report(cl)
    loss2 = sum((im * c).sum() for im, c in zip(own_imgs, cots)) + own_reg
    own_params = {"sdfs": own.sdf_params, "deforms": own.deform_params, "weights": own.weight_params,
                  "light": own.cubemap, "exposure": own.exposure_params,
                  "kd_table": own.field.kd_enc.hash_table, "ks_table": own.field.ks_enc.hash_table,
                  "z_table": own.field.z_enc.hash_table, "kd_mlp0": own.field.kd_enc.mlp.weights[0],
                  "ks_mlp1": own.field.ks_enc.mlp.weights[1]}
    own_grads = dict(zip(own_params, torch.autograd.grad(loss2, list(own_params.values()))))

    # ---------------------------------------------------- B4: this library's operators called with the REFERENCE'S
    # argument types and signature (geosplat.py:53-65): Cameras[1], TextureSplitSum, no fg_lut; MGAdapter.make(mesh)
    from geosplatting_b200.mgadapter import MGAdapter as OwnAdapter
    from geosplatting_b200.splat import GSplatter as OwnSplatter
    from geosplatting_b200.splat import RenderableAttrs as OwnAttrs
    with torch.no_grad():
        mesh, r_gsplat, r_attrs, _, _ = ref.get_gsplat("face")
        r_env, _ = ref.get_envmap()
        exposures = ref.exposure_params.exp().expand(2)
        b4 = []
        for i in range(2):
            want = r_attrs.splat(r_gsplat, cams[i:i + 1], envmap=r_env, exposure=exposures[i],
                                 min_roughness=ref.min_roughness, max_metallic=ref.max_metallic)
            got = OwnAttrs(kd=r_attrs.kd, ks=r_attrs.ks, normals=r_attrs.normals).splat(
                OwnSplatter(gaussians=r_gsplat.gaussians, rasterize_mode="antialiased"), cams[i:i + 1], envmap=r_env,
                exposure=exposures[i], min_roughness=ref.min_roughness, max_metallic=ref.max_metallic)
            b4.append(float((want - got).abs().max()))
        nmesh = mesh.compute_vertex_normals(fix=True)
        r_splats, r_off = GEO.MGAdapter().make(nmesh)
        o_splats, o_off = OwnAdapter().make(nmesh)
        mg = max(float((getattr(r_splats, k) - getattr(o_splats, k)).abs().max())
                 for k in ("means", "scales", "quats", "colors", "opacities"))
        mg = max(mg, float((r_off - o_off).abs().max()))

    out = {"b4_reference_types_image_linf": b4, "b4_mgadapter_mesh_linf": mg, "gaussians_ref": int(n_ref), "gaussians_own": int(n_own), "ks_state_dict_keys_equal": bool(keys_equal),
           "ks_state_dict_missing": [list(missing.missing_keys), list(missing.unexpected_keys)],
           "image_linf": [float((a - b).abs().max()) for a, b in zip(ref_imgs, own_imgs)],
           "coverage": [float(a[..., 3].mean()) for a in ref_imgs],
           "image_mean_rgb": [float(a[..., :3].mean()) for a in ref_imgs],
           "reg": [float(reg), float(own_reg)],
           "grad_rel_l2": {k: rel(own_grads[k], ref_grads[k]) for k in ref_grads},
           "grad_max": {k: float(ref_grads[k].abs().max()) for k in ref_grads}}
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
