#!/usr/bin/env python
"""Generates tests/golden/ref_*.npz by running the REFERENCE'S OWN Python code for the hot path
(/root/reference, read-only, this container only) on seeded inputs.

    python scripts/make_golden.py

The third-party kernels the reference calls (`nvdiffrast.torch.texture`) are supplied by this
repository's CPU restatement (oracle/texture.py), so the fixtures pin:
  * the reference's own algebra (MGAdapter.make, compute_vertex_normals, RenderableAttrs.splat shade block,
    TextureSplitSum.sample mip-level rule, _merge/_split_mipmaps, _CubeMapMip fwd/bwd, tone mapping,
    Cameras.view_matrix / intrinsic_matrix, rot2quat, safe_normalize) -- bit-for-bit what the reference
    computes on the CPU in fp32;
  * NOT nvdiffrast / gsplat themselves (absent here: "parity unpinned").
Gradients are torch autograd through the same code.  Fixtures are small (< 250 KB each) and committed.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import ref_shim  # noqa: E402
from geosplatting_b200 import scenes  # noqa: E402
from oracle import texture as T  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def dr_texture(tex, uv, uv_da=None, mip_level_bias=None, mip=None, filter_mode="auto", boundary_mode="wrap",
               max_mip_level=None):
    """nvdiffrast.torch.texture stand-in for the three modes the hot path uses (oracle/texture.py)."""
    assert uv_da is None and max_mip_level is None
    if boundary_mode == "clamp":
        assert filter_mode == "linear" and tex.shape[0] == 1
        out = T.texture_2d_linear_clamp(tex[0], uv.reshape(-1, 2))
        return out.reshape(*uv.shape[:-1], tex.shape[-1])
    assert boundary_mode == "cube" and tex.shape[:2] == (1, 6)
    d = uv.reshape(-1, 3)
    if filter_mode == "linear":
        out = T.texture_cube_linear(tex[0], d)
    else:
        assert filter_mode == "linear-mipmap-linear"
        out = T.texture_cube_mip([tex[0]] + [m[0] for m in mip], d, mip_level_bias.reshape(-1))
    return out.reshape(*uv.shape[:-1], tex.shape[-1])


def synthetic_fg_lut():
    """Deterministic smooth stand-in for the split-sum DFG LUT [256,256,2] (the real asset is 512 KB)."""
    i = np.arange(256, dtype=np.float64)[:, None]
    j = np.arange(256, dtype=np.float64)[None, :]
    a = 0.5 + 0.45 * np.sin(0.021 * i + 0.013 * j)
    b = 0.5 + 0.45 * np.cos(0.017 * i - 0.011 * j)
    return np.stack([a, b], -1).astype(np.float32)


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                 for k, v in arrays.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KB")


def main():
    ref_shim.install(dr_texture)
    import rfstudio.graphics as G
    from rfstudio.graphics import math as RM
    from rfstudio.graphics._mesh import _texture as RT
    from rfstudio.graphics import shaders as RS
    ns = ref_shim.load_geosplat_head()
    torch.manual_seed(0)

    # ---- A. MGAdapter.make + compute_vertex_normals(fix=True)  (geosplat.py:390-472, _triangle_mesh.py:588-614)
    verts, faces = scenes.icosphere(1, radius=0.6)  # 42 vertices, 80 faces
    g = torch.Generator().manual_seed(3)
    verts = verts * (1.0 + 0.15 * torch.randn(verts.shape[0], 1, generator=g))
    verts = verts.clone().requires_grad_(True)
    mesh = G.TriangleMesh(vertices=verts, indices=faces).compute_vertex_normals(fix=True)
    splats, offsets = ns["MGAdapter"]().make(mesh)
    outs = dict(means=splats.means, scales=splats.scales, quats=splats.quats, colors=splats.colors,
                opacities=splats.opacities)
    cot = {k: torch.randn(v.shape, generator=g) for k, v in outs.items()}
    loss = sum((outs[k] * cot[k]).sum() for k in outs)
    v_verts, = torch.autograd.grad(loss, verts)
    save("ref_mgadapter.npz", vertices=verts, indices=faces, vertex_normals=mesh.normals, offsets=offsets,
         v_vertices=v_verts, **outs, **{"cot_" + k: v for k, v in cot.items()})

    # ---- B. RenderableAttrs.splat shade block (geosplat.py:83-121) for the three modes
    N = 600
    lut = torch.from_numpy(synthetic_fg_lut())
    ns["_get_fg_lut"] = lambda resolution, device: lut.view(1, 256, 256, 2)
    sg = scenes.surface_gaussians(N, seed=5)
    means = sg["means"].clone().requires_grad_(True)
    normals = torch.nn.functional.normalize(sg["normals"] + 0.2 * torch.randn(N, 3, generator=g), dim=-1)
    normals = normals.clone().requires_grad_(True)
    kd = sg["kd"].clone().requires_grad_(True)
    ks = sg["ks"].clone()
    ks[:40, 0] = torch.linspace(0, 1, 40)  # sweep roughness over both mip-level branches
    ks = ks.requires_grad_(True)
    base = torch.exp(0.5 * torch.randn(6, 16, 16, 3, generator=g)).requires_grad_(True)
    mips = [torch.exp(0.7 * torch.randn(6, r, r, 3, generator=g)) for r in (32, 16, 8, 4, 2)]
    packed = RT._merge_mipmaps(mips)
    packed[:, 3, 30:, 30:] = 0  # _merge_mipmaps leaves this never-read tail uninitialised (torch.empty_like)
    packed = packed.clone().requires_grad_(True)
    envmap = G.TextureSplitSum(base=base, mipmaps=packed, num_mipmaps=torch.tensor([5]),
                               min_roughness=torch.tensor([0.08]), max_roughness=torch.tensor([0.5]), transform=None)
    cam = scenes.orbit_cameras(1, 64, 64, seed=2)[0]
    cams = G.Cameras(c2w=torch.from_numpy(cam.c2w)[None], fx=torch.tensor([cam.fx]), fy=torch.tensor([cam.fy]),
                     cx=torch.tensor([cam.cx]), cy=torch.tensor([cam.cy]), width=torch.tensor([64]),
                     height=torch.tensor([64]), near=torch.tensor([0.01]), far=torch.tensor([100.0]))

    class _Gaussians:
        def __init__(self, m):
            self.means = m
            self.colors = None

        def __getitem__(self, mask):
            return self

        def replace_(self, colors):
            self.colors = colors

    class _FakeGSplatter:
        def __init__(self, m):
            self.gaussians = _Gaussians(m)

        def render_rgba(self, cameras):
            class _R:
                def item(_self):
                    return torch.zeros(4, 4, 4)
            return _R()

    attrs = ns["RenderableAttrs"](kd=kd, ks=ks, normals=normals, occ=None, kd_jitter=None, ks_jitter=None)
    shade_out = {}
    for mode in ("pbr", "diffuse", "specular"):
        fake = _FakeGSplatter(means)
        gauss = fake.gaussians
        attrs.splat(fake, cams, exposure=torch.ones(1), envmap=envmap, min_roughness=0.1, max_metallic=1.0,
                    mode=mode, tone_type="none")
        colors = gauss.colors
        cotc = torch.randn(N, 3, generator=g)
        grads = torch.autograd.grad((colors * cotc).sum(), [means, normals, kd, ks, base, packed], allow_unused=True)
        shade_out[f"colors_{mode}"] = colors
        shade_out[f"cot_{mode}"] = cotc
        for nm, gr in zip(("means", "normals", "kd", "ks", "base", "packed"), grads):
            shade_out[f"v_{nm}_{mode}"] = torch.zeros(1) if gr is None else gr
    # the packed env-map gradient is sparse; store it as float16-free sparse triplets to stay small
    for mode in ("pbr", "diffuse", "specular"):
        gp = shade_out.pop(f"v_packed_{mode}")
        if gp.numel() > 1:
            idx = torch.nonzero(gp.reshape(-1)).reshape(-1)
            shade_out[f"v_packed_idx_{mode}"] = idx.to(torch.int32)
            shade_out[f"v_packed_val_{mode}"] = gp.reshape(-1)[idx]
    save("ref_shade.npz", means=means, normals=normals, kd=kd, ks=ks, base=base,
         packed=packed.detach(), num_mipmaps=5,
         cam_pos=torch.from_numpy(cam.c2w[:, 3].copy()), **shade_out)

    # ---- B2. the same shade block on the REFERENCE'S OWN FG LUT (shaders.py:22-26 loads the asset): pbr mode, its own
    #          generator so that the fixtures above and below do not move
    g2 = torch.Generator().manual_seed(11)
    ns["_get_fg_lut"] = RS._get_fg_lut
    N2 = 2000
    sg2 = scenes.surface_gaussians(N2, seed=6)
    leaves2 = [sg2["means"].clone().requires_grad_(True),
               torch.nn.functional.normalize(sg2["normals"] + 0.3 * torch.randn(N2, 3, generator=g2), dim=-1).requires_grad_(True),
               torch.rand(N2, 3, generator=g2).requires_grad_(True), torch.rand(N2, 2, generator=g2).requires_grad_(True)]
    attrs2 = ns["RenderableAttrs"](kd=leaves2[2], ks=leaves2[3], normals=leaves2[1], occ=None, kd_jitter=None,
                                   ks_jitter=None)
    fake2 = _FakeGSplatter(leaves2[0])
    gauss2 = fake2.gaussians
    attrs2.splat(fake2, cams, exposure=torch.ones(1), envmap=envmap, min_roughness=0.1, max_metallic=1.0, mode="pbr",
                 tone_type="none")
    cot2 = torch.randn(N2, 3, generator=g2)
    gr2 = torch.autograd.grad((gauss2.colors * cot2).sum(), leaves2)
    save("ref_shade_real_lut.npz", means=leaves2[0], normals=leaves2[1], kd=leaves2[2], ks=leaves2[3], colors=gauss2.colors,
         cot=cot2, v_means=gr2[0], v_normals=gr2[1], v_kd=gr2[2], v_ks=gr2[3], base=base, packed=packed.detach(),
         num_mipmaps=5, cam_pos=torch.from_numpy(cam.c2w[:, 3].copy()))
    ns["_get_fg_lut"] = lambda resolution, device: lut.view(1, 256, 256, 2)

    # ---- C. tone mapping (geosplat.py:474-476)
    rgba = (torch.rand(32, 32, 4, generator=g) * 1.6).requires_grad_(True)
    exposure = torch.tensor([1.3], requires_grad=True)
    out = ns["_tone_mapping_naive"](rgba, exposure)
    cott = torch.randn(32, 32, 4, generator=g)
    v_rgba, v_exp = torch.autograd.grad((out * cott).sum(), [rgba, exposure])
    save("ref_tonemap.npz", rgba=rgba, exposure=exposure, out=out, cot=cott, v_rgba=v_rgba, v_exposure=v_exp)

    # ---- D. split-sum plumbing: merge/split, sample(), _CubeMapMip fwd/bwd (_texture.py:199-261,:571-613)
    small = [torch.rand(6, r, r, 3, generator=g) for r in (16, 8, 4)]
    merged = RT._merge_mipmaps(small)
    merged[:, 3, 12:, 12:] = 0
    split = RT._split_mipmaps(merged, num_mipmaps=3)
    assert all(torch.equal(a, b) for a, b in zip(small, split))
    nd = torch.randn(300, 3, generator=g)
    dd = torch.randn(300, 3, generator=g)
    rough = torch.rand(300, 1, generator=g)
    env_small = G.TextureSplitSum(base=base.detach(), mipmaps=torch.nan_to_num(merged), num_mipmaps=torch.tensor([3]),
                                  min_roughness=torch.tensor([0.08]), max_roughness=torch.tensor([0.5]),
                                  transform=None)
    l_diff, l_spec = env_small.sample(normals=nd[None, None], directions=dd[None, None], roughness=rough[None, None])
    cube = torch.rand(6, 8, 8, 3, generator=g).requires_grad_(True)
    down = RT._CubeMapMip.apply(cube)
    cotd = torch.randn(6, 4, 4, 3, generator=g)
    v_cube, = torch.autograd.grad((down * cotd).sum(), cube)
    save("ref_splitsum.npz", mip0=small[0], mip1=small[1], mip2=small[2], merged=torch.nan_to_num(merged),
         normals=nd, directions=dd, roughness=rough, base=base.detach(), l_diff=l_diff.reshape(-1, 3),
         l_spec=l_spec.reshape(-1, 3), cube=cube, down=down, cot_down=cotd, v_cube=v_cube)

    # ---- E. cameras (_cameras.py:289-314)
    cl = scenes.orbit_cameras(4, 800, 800, seed=7)
    cams4 = G.Cameras(c2w=torch.from_numpy(np.stack([c.c2w for c in cl])), fx=torch.tensor([c.fx for c in cl]),
                      fy=torch.tensor([c.fy for c in cl]), cx=torch.tensor([c.cx for c in cl]),
                      cy=torch.tensor([c.cy for c in cl]), width=torch.tensor([800] * 4),
                      height=torch.tensor([800] * 4), near=torch.tensor([0.01] * 4), far=torch.tensor([100.0] * 4))
    save("ref_cameras.npz", c2w=cams4.c2w, fx=cams4.fx, fy=cams4.fy, cx=cams4.cx, cy=cams4.cy,
         view_matrix=cams4.view_matrix, intrinsic_matrix=cams4.intrinsic_matrix)

    # ---- F. math helpers (math.py:119-128, :246-278)
    rots = RM.quat2rot(torch.nn.functional.normalize(torch.randn(200, 4, generator=g), dim=-1))
    vecs = torch.randn(50, 3, generator=g)
    vecs[:3] = 0
    save("ref_math.npz", rots=rots, quats=RM.rot2quat(rots), vecs=vecs, safe_normalized=RM.safe_normalize(vecs))

    # ---- G. the reference's FG LUT asset: hash + a 32x32 subsample (shaders.py:22-26)
    path = "/root/reference/rfstudio/assets/geometry/pbr/bsdf_256_256.bin"
    raw = open(path, "rb").read()
    full = np.frombuffer(raw, dtype=np.float32).reshape(256, 256, 2)
    save("ref_fg_lut_sub.npz", sha256=hashlib.sha256(raw).hexdigest(), sub=full[::8, ::8].copy(),
         corners=np.stack([full[0, 0], full[0, 255], full[255, 0], full[255, 255]]))
    # the whole table (a published split-sum DFG table, 512 KB of data, no code): the GPU box has no reference tree,
    # and the shade kernels must be exercised on the reference's bytes, not only on a synthetic table
    save("ref_fg_lut.npz", lut=full, sha256=hashlib.sha256(raw).hexdigest())

    # ---- H. HashEncoding (torch backend) + MLP: the fields that produce kd / ks / z (SURVEY section 8f rank 1;
    #         rfstudio/model/components/encoding.py:124-241, rfstudio/nn/mlp.py:125-145, configs geosplat.py:485-518).
    #         The reference's Module framework does not initialise under this Python, so its METHODS are run on plain
    #         namespaces: __setup__ (scalings, offsets, table init), hash_fn, pytorch_fwd, __call__ (grad scaling) and
    #         MLP.__call__ are the reference's own code; only the nn.Linear containers are built here.
    import functools
    import importlib.machinery
    import types
    for name, path_ in (("rfstudio.model", "/root/reference/rfstudio/model"),
                        ("rfstudio.model.components", "/root/reference/rfstudio/model/components")):
        m = types.ModuleType(name)
        m.__path__ = [path_]
        m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
        sys.modules[name] = m
    import rfstudio.model.components.encoding as RE
    from rfstudio.nn import mlp as RMLP

    class _Holder:
        @staticmethod
        def from_tensor(t):
            return types.SimpleNamespace(params=t.clone().requires_grad_(True))

    RE.ParameterModule = _Holder
    enc_out = {}
    for tag, layers, act, log2 in (("kd", [32, 32, 32, 3], "sigmoid", 10), ("ks", [32, 32, 2], "none", 12),
                                   ("z", [32, 32, 1], "none", 9)):
        torch.manual_seed({"kd": 11, "ks": 12, "z": 13}[tag])
        e = types.SimpleNamespace(num_levels=16, min_res=16, max_res=4096, log2_hashmap_size=log2, features_per_level=2,
                                  hash_init_scale=0.001, backend="torch", interpolation="linear", grad_scaling=16.0)
        RE.HashEncoding.__setup__(e)
        e.hash_fn = functools.partial(RE.HashEncoding.hash_fn, e)
        e.pytorch_fwd = functools.partial(RE.HashEncoding.pytorch_fwd, e)
        with torch.no_grad():   # a 1e-3 table gives nearly constant outputs: use O(1) features for a meaningful check
            e.hash_table.mul_(1000.0)
        lin = [torch.nn.Linear(i, o, bias=False) for i, o in zip(layers[:-1], layers[1:])]
        for l in lin:
            torch.nn.init.kaiming_uniform_(l.weight, nonlinearity="relu")           # rfstudio/nn/mlp.py:99-100
        mlp_ns = types.SimpleNamespace(nn_layers=lin, skip_connection_set=set(), activation=act,
                                       initialize_weights=lambda d: None)
        e.mlp = functools.partial(RMLP.MLP.__call__, mlp_ns)
        x = (torch.rand(1500, 3, generator=g) * 2 - 1)
        x[:8] = torch.tensor([[-1.0, -1, -1], [1, 1, 1], [0, 0, 0], [1, -1, 0.5], [0.25, 0.5, -0.75], [-1, 1, 1],
                              [0.999999, -0.999999, 0], [0.0625, 0.0625, 0.0625]])      # cell corners / exact lattice points
        x = x.requires_grad_(True)
        feats = e.pytorch_fwd(x)
        y = RE.HashEncoding.__call__(e, x)
        coty = torch.randn(y.shape, generator=g)
        grads = torch.autograd.grad((y * coty).sum(), [x, e.hash_table] + [l.weight for l in lin])
        idx = torch.nonzero(grads[1].abs().sum(-1)).reshape(-1)
        enc_out.update({f"{tag}_x": x, f"{tag}_table": e.hash_table, f"{tag}_feats": feats, f"{tag}_y": y, f"{tag}_cot": coty,
                        f"{tag}_v_x": grads[0], f"{tag}_v_table_idx": idx.to(torch.int32), f"{tag}_v_table_val": grads[1][idx],
                        f"{tag}_scalings": e.scalings, f"{tag}_log2": log2})
        for k, (l, gw) in enumerate(zip(lin, grads[2:])):
            enc_out[f"{tag}_w{k}"] = l.weight
            enc_out[f"{tag}_v_w{k}"] = gw
    save("ref_encoding.npz", **enc_out)

    # ---- I. GaussianField.get_patches / get_gaussians_from_vertex / get_gaussians_from_face (SURVEY 8a row a3;
    #         geosplat.py:520-674).  The class itself cannot be created under Python 3.12 (mutable dataclass default at
    #         :482), so the three method bodies are cut out of the source by line range and exec'd as plain functions in
    #         the namespace of the module head; `self` is a namespace holding the encoders of section H's kind.
    import ast
    import textwrap
    src_lines = open("/root/reference/rfstudio/model/geosplat.py").read().split("\n")
    tree = ast.parse("\n".join(src_lines))
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "GaussianField")
    fns = {}
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("get_patches", "get_gaussians_from_vertex",
                                                           "get_gaussians_from_face"):
            code = textwrap.dedent("\n".join(src_lines[fn.lineno - 1:fn.end_lineno]))
            exec(compile(code, "/root/reference/rfstudio/model/geosplat.py", "exec"), ns)
            fns[fn.name] = ns[fn.name]

    def make_enc(layers, act, log2, seed):
        torch.manual_seed(seed)
        e = types.SimpleNamespace(num_levels=16, min_res=16, max_res=4096, log2_hashmap_size=log2, features_per_level=2,
                                  hash_init_scale=0.001, backend="torch", interpolation="linear", grad_scaling=16.0)
        RE.HashEncoding.__setup__(e)
        e.hash_fn = functools.partial(RE.HashEncoding.hash_fn, e)
        e.pytorch_fwd = functools.partial(RE.HashEncoding.pytorch_fwd, e)
        with torch.no_grad():
            e.hash_table.mul_(1000.0)
        lin = [torch.nn.Linear(i, o, bias=False) for i, o in zip(layers[:-1], layers[1:])]
        for l in lin:
            torch.nn.init.kaiming_uniform_(l.weight, nonlinearity="relu")
        mlp_ns = types.SimpleNamespace(nn_layers=lin, skip_connection_set=set(), activation=act,
                                       initialize_weights=lambda d: None)
        e.mlp = functools.partial(RMLP.MLP.__call__, mlp_ns)
        return (lambda x: RE.HashEncoding.__call__(e, x)), e.hash_table, [l.weight for l in lin]

    kd_f, kd_t, kd_w = make_enc([32, 32, 32, 3], "sigmoid", 9, 21)
    ks_f, ks_t, ks_w = make_enc([32, 32, 2], "none", 9, 22)
    z_f, z_t, z_w = make_enc([32, 32, 1], "none", 9, 23)
    field = types.SimpleNamespace(kd_enc=kd_f, ks_enc=ks_f, z_enc=z_f, occ_enc=None, device=torch.device("cpu"))
    field.get_patches = functools.partial(fns["get_patches"], field)
    verts2, faces2 = scenes.icosphere(1, radius=0.6)
    verts2 = (verts2 * (1.0 + 0.1 * torch.randn(verts2.shape[0], 1, generator=g))).requires_grad_(True)
    guess = torch.tensor([0.3, -0.2])
    fout = {"vertices": verts2, "indices": faces2, "initial_guess": guess, "scale": 0.9}
    for k, (t_, w_) in (("kd", (kd_t, kd_w)), ("ks", (ks_t, ks_w)), ("z", (z_t, z_w))):
        fout[f"{k}_table"] = t_
        for i, w in enumerate(w_):
            fout[f"{k}_w{i}"] = w
    mesh2 = G.TriangleMesh(vertices=verts2, indices=faces2)
    pts, areas = fns["get_patches"](field, mesh2)
    fout.update(patch_normals=pts.normals, patch_areas=areas)
    sp_v, at_v = fns["get_gaussians_from_vertex"](field, 0.0, 0.0, 0.9, mesh2, guess)
    outs_v = dict(v_means=sp_v.means, v_scales=sp_v.scales, v_quats=sp_v.quats, v_opacities=sp_v.opacities, v_kd=at_v.kd,
                  v_ks=at_v.ks, v_normals=at_v.normals)
    cot_v = {k: torch.randn(v.shape, generator=g) for k, v in outs_v.items()}
    gv = torch.autograd.grad(sum((outs_v[k] * cot_v[k]).sum() for k in outs_v), [verts2, kd_t, z_t], allow_unused=True)
    fout.update(outs_v)
    fout.update({"cot_" + k: v for k, v in cot_v.items()})
    fout.update(vertexpath_grad_vertices=gv[0], vertexpath_grad_kd_table=gv[1], vertexpath_grad_z_table=gv[2])
    sp_f, at_f, off_f = fns["get_gaussians_from_face"](field, mesh2, 0.0, 0.0, scale=0.9, initial_guess=guess)
    outs_f = dict(f_means=sp_f.means, f_scales=sp_f.scales, f_quats=sp_f.quats, f_opacities=sp_f.opacities, f_kd=at_f.kd,
                  f_ks=at_f.ks, f_normals=at_f.normals)
    cot_f = {k: torch.randn(v.shape, generator=g) for k, v in outs_f.items()}
    gf = torch.autograd.grad(sum((outs_f[k] * cot_f[k]).sum() for k in outs_f), [verts2, kd_t, z_t], allow_unused=True)
    fout.update(outs_f)
    fout.update({"cot_" + k: v for k, v in cot_f.items()})
    fout.update(f_offsets=off_f, facepath_grad_vertices=gf[0], facepath_grad_kd_table=gf[1], facepath_grad_z_table=gf[2])
    save("ref_field.npz", **fout)

    # ---- J. FlexiCubes.dual_marching_cubes + compute_entropy (_flexicubes.py:559-725), as
    #         GeoSplatter.get_geometry drives it (geosplat.py:751-769): fixtures for flexicubes.py + csrc/flexicubes.cu, so
    #         that a CUDA implementation can be held to the reference's own outputs (vertex / face ORDER included).
    R = 10
    fc0 = G.FlexiCubes.from_resolution(R, random_sdf=False, scale=0.9)
    gv = fc0.vertices
    # a bumpy sphere's SDF (the same family bench.py meshes come from) on the grid vertices
    sdf = (gv.norm(dim=-1, keepdim=True) - 0.55 + 0.08 * torch.sin(5.0 * gv[:, :1]) * torch.cos(4.0 * gv[:, 1:2]))
    sdf = sdf.clone().requires_grad_(True)
    deform = (0.3 * torch.randn(gv.shape, generator=g)).requires_grad_(True)
    weights = (0.2 * torch.randn(fc0.indices.shape[0], 21, generator=g)).requires_grad_(True)
    verts_fc = gv + deform.tanh() * (0.5 * 0.9 / R)                                   # geosplat.py:756
    fc = fc0.replace(vertices=verts_fc, sdf_values=sdf, alpha=weights[:, :8], beta=weights[:, 8:20],
                     gamma=weights[:, 20:])
    mesh_fc, L_dev = fc.dual_marching_cubes()
    entropy = fc.compute_entropy()
    cot_v = torch.randn(mesh_fc.vertices.shape, generator=g)
    loss_fc = (mesh_fc.vertices * cot_v).sum() + L_dev.mean() * 0.5 + entropy * 0.3
    g_sdf, g_def, g_w = torch.autograd.grad(loss_fc, [sdf, deform, weights])
    from rfstudio.graphics._mesh import _flexicubes as RFC
    cpu = torch.device("cpu")
    # the four lookup tables are DATA of the algorithm (published with the FlexiCubes paper): stored in the fixture
    # by value, read from the reference at generation time
    tables = dict(tbl_cube_edges=RFC._get_cube_edges(cpu), tbl_check=RFC._get_check_table(cpu),
                  tbl_dmc=RFC._get_dmc_table(cpu), tbl_num_vd=RFC._get_num_vd_table(cpu))
    # the same four tables, int32, as the product's data file (geosplatting_b200/data/flexicubes_tables.npz)
    np.savez_compressed(os.path.join(os.path.dirname(OUT), "..", "geosplatting_b200", "data", "flexicubes_tables.npz"),
                        **{k[4:]: v.numpy().astype(np.int32) for k, v in tables.items()})
    save("ref_flexicubes.npz", **tables, resolution=R, scale=0.9, grid_vertices=gv, cube_indices=fc0.indices, sdf=sdf,
         deform=deform, weights=weights, mesh_vertices=mesh_fc.vertices, mesh_indices=mesh_fc.indices, L_dev=L_dev,
         entropy=entropy, cot_vertices=cot_v, v_sdf=g_sdf, v_deform=g_def, v_weights=g_w)

    # a rough random SDF on a 7 x 6 x 5 grid: every topology case incl. the ambiguous ones that get inverted
    # (_flexicubes.py:472-505), non-cubic resolution, cubes with 2-4 dual vertices
    fr = G.FlexiCubes.from_resolution(7, 6, 5, random_sdf=False, scale=1.0)
    sdf_r = (torch.rand(fr.vertices.shape[0], 1, generator=g) - 0.45).requires_grad_(True)
    w_r = (0.5 * torch.randn(fr.indices.shape[0], 21, generator=g)).requires_grad_(True)
    fcr = fr.replace(sdf_values=sdf_r, alpha=w_r[:, :8], beta=w_r[:, 8:20], gamma=w_r[:, 20:])
    occ_r = (sdf_r < 0)[fr.indices.flatten()].view(-1, 8)
    surf_r = (occ_r.sum(-1) > 0) & (occ_r.sum(-1) < 8)
    plain_case = (occ_r[surf_r] * torch.pow(2, torch.arange(8))).sum(-1)
    resolved_case = fcr._get_case_id(occ_r, surf_r)
    mesh_r, L_r = fcr.dual_marching_cubes()
    cot_r = torch.randn(mesh_r.vertices.shape, generator=g)
    gr = torch.autograd.grad((mesh_r.vertices * cot_r).sum() + L_r.mean(), [sdf_r, w_r])
    save("ref_flexicubes_rough.npz", **tables, resolution=torch.tensor([7, 6, 5]), grid_vertices=fr.vertices,
         cube_indices=fr.indices, sdf=sdf_r, weights=w_r, mesh_vertices=mesh_r.vertices, mesh_indices=mesh_r.indices,
         L_dev=L_r, entropy=fcr.compute_entropy(), cot_vertices=cot_r, v_sdf=gr[0], v_weights=gr[1],
         n_inverted=(plain_case != resolved_case).sum())


if __name__ == "__main__":
    main()
