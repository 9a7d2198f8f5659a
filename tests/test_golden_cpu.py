"""The oracle restatements (oracle/shade.py, oracle/texture.py) against fixtures produced by the
REFERENCE'S OWN Python code (scripts/make_golden.py, run in the build container).  No GPU."""
import os

import numpy as np
import pytest
import torch

from geosplatting_b200 import scenes
from oracle import shade as S
from oracle import texture as T

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return {k: v for k, v in np.load(os.path.join(GOLD, name), allow_pickle=False).items()}


def synthetic_fg_lut():
    i = np.arange(256, dtype=np.float64)[:, None]
    j = np.arange(256, dtype=np.float64)[None, :]
    a = 0.5 + 0.45 * np.sin(0.021 * i + 0.013 * j)
    b = 0.5 + 0.45 * np.cos(0.017 * i - 0.011 * j)
    return np.stack([a, b], -1).astype(np.float32)


@pytest.mark.parametrize("mode", ["pbr", "diffuse", "specular"])
def test_shade_matches_reference_code(mode):
    g = load("ref_shade.npz")
    t = lambda k: torch.tensor(g[k], requires_grad=True)
    means, normals, kd, ks, base, packed = t("means"), t("normals"), t("kd"), t("ks"), t("base"), t("packed")
    mips = T.split_mipmaps(packed, int(g["num_mipmaps"]))
    colors = S.shade(means, normals, kd, ks, torch.tensor(g["cam_pos"]), torch.from_numpy(synthetic_fg_lut()), base,
                     mips, min_roughness=0.1, max_metallic=1.0, mode=mode)
    assert np.abs(colors.detach().numpy() - g[f"colors_{mode}"]).max() < 2e-6
    grads = torch.autograd.grad((colors * torch.tensor(g[f"cot_{mode}"])).sum(), [means, normals, kd, ks, base, packed],
                                allow_unused=True)
    for nm, gr in zip(("means", "normals", "kd", "ks", "base"), grads):
        ref = g[f"v_{nm}_{mode}"]
        if ref.size == 1:
            assert gr is None or float(gr.abs().max()) == 0
            continue
        assert np.abs(gr.numpy() - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), nm
    if f"v_packed_idx_{mode}" in g:
        dense = np.zeros(packed.numel(), np.float32)
        dense[g[f"v_packed_idx_{mode}"]] = g[f"v_packed_val_{mode}"]
        assert np.abs(grads[5].numpy().reshape(-1) - dense).max() <= 1e-4 * max(1.0, np.abs(dense).max())


def test_tonemap_matches_reference_code():
    g = load("ref_tonemap.npz")
    rgba = torch.tensor(g["rgba"], requires_grad=True)
    ex = torch.tensor(g["exposure"], requires_grad=True)
    out = S.tone_map_naive(rgba, ex)
    assert np.abs(out.detach().numpy() - g["out"]).max() < 1e-6
    v_rgba, v_ex = torch.autograd.grad((out * torch.tensor(g["cot"])).sum(), [rgba, ex])
    assert np.abs(v_rgba.numpy() - g["v_rgba"]).max() < 1e-5
    assert abs(float(v_ex) - float(g["v_exposure"])) < 1e-3 * abs(float(g["v_exposure"]))


def test_splitsum_plumbing_matches_reference_code():
    g = load("ref_splitsum.npz")
    mips = [torch.tensor(g[f"mip{i}"]) for i in range(3)]
    assert np.array_equal(T.merge_mipmaps(mips).numpy(), g["merged"])
    back = T.split_mipmaps(torch.tensor(g["merged"]), 3)
    assert all(torch.equal(a, b) for a, b in zip(mips, back))
    l_diff, l_spec = S.splitsum_sample(torch.tensor(g["base"]), mips, torch.tensor(g["normals"]),
                                       torch.tensor(g["directions"]), torch.tensor(g["roughness"])[:, 0])
    assert np.abs(l_diff.numpy() - g["l_diff"]).max() < 1e-6
    assert np.abs(l_spec.numpy() - g["l_spec"]).max() < 1e-6
    cube = torch.tensor(g["cube"])
    assert np.abs(S.cubemap_mip_fwd(cube).numpy() - g["down"]).max() < 1e-7
    assert np.abs(S.cubemap_mip_bwd(torch.tensor(g["cot_down"])).numpy() - g["v_cube"]).max() < 1e-6


def test_cameras_match_reference_code():
    g = load("ref_cameras.npz")
    for i in range(g["c2w"].shape[0]):
        cam = scenes.PinholeCamera(g["c2w"][i], float(g["fx"][i]), float(g["fy"][i]), float(g["cx"][i]),
                                   float(g["cy"][i]), 800, 800)
        assert np.abs(cam.view_matrix - g["view_matrix"][i]).max() < 1e-6
        assert np.array_equal(cam.intrinsic_matrix, g["intrinsic_matrix"][i])


def test_math_helpers_match_reference_code():
    g = load("ref_math.npz")
    q = scenes.rotmat_to_quat(torch.tensor(g["rots"])).numpy()
    assert np.abs(q - g["quats"]).max() < 1e-6
    assert np.abs(S.safe_normalize(torch.tensor(g["vecs"])).numpy() - g["safe_normalized"]).max() < 1e-7


def test_cube_face_convention_inverts_the_reference_map():
    """oracle/texture.cube_face_uv must invert `_cube_to_dir` (_texture.py:178-197): the direction of
    texel (s, y, x) must come back as face s with (u, v) at that texel's centre."""
    R = 8
    dirs = S.cube_texel_dirs(R).reshape(-1, 3)
    face, u, v = T.cube_face_uv(dirs)
    exp_face = torch.arange(6).repeat_interleave(R * R)
    assert torch.equal(face, exp_face)
    ys, xs = torch.meshgrid(torch.arange(R), torch.arange(R), indexing="ij")
    assert torch.allclose(u.reshape(6, R, R), ((xs + 0.5) / R).expand(6, R, R), atol=1e-6)
    assert torch.allclose(v.reshape(6, R, R), ((ys + 0.5) / R).expand(6, R, R), atol=1e-6)


def test_mgadapter_oracle_matches_reference_code():
    from oracle import mgadapter as MG
    g = load("ref_mgadapter.npz")
    verts = torch.tensor(g["vertices"], requires_grad=True)
    faces = torch.tensor(g["indices"])
    vn = MG.vertex_normals(verts, faces)
    assert np.abs(vn.detach().numpy() - g["vertex_normals"]).max() < 1e-6
    means, scales, quats, colors, opac, offsets = MG.make(verts, faces, vn)
    for name, t in (("means", means), ("scales", scales), ("quats", quats), ("colors", colors), ("opacities", opac),
                    ("offsets", offsets)):
        assert np.abs(t.detach().numpy() - g[name]).max() < 2e-6, name
    loss = sum((t * torch.tensor(g["cot_" + k])).sum() for k, t in
               (("means", means), ("scales", scales), ("quats", quats), ("colors", colors), ("opacities", opac)))
    gv, = torch.autograd.grad(loss, verts)
    assert np.abs(gv.numpy() - g["v_vertices"]).max() <= 1e-4 * np.abs(g["v_vertices"]).max()


def test_encoding_oracle_matches_reference_code():
    """oracle/encoding.py against the reference's own HashEncoding (torch backend) + MLP code
    (tests/golden/ref_encoding.npz, scripts/make_golden.py section H): features, outputs and every gradient."""
    from oracle import encoding as OE
    g = load("ref_encoding.npz")
    for tag, act, n_w in (("kd", "sigmoid", 3), ("ks", "none", 2), ("z", "none", 2)):
        log2 = int(g[f"{tag}_log2"])
        scal = OE.level_scalings(16, 16, 4096)
        assert np.array_equal(scal.numpy(), g[f"{tag}_scalings"])
        x = torch.from_numpy(g[f"{tag}_x"]).requires_grad_(True)
        table = torch.from_numpy(g[f"{tag}_table"]).requires_grad_(True)
        ws = [torch.from_numpy(g[f"{tag}_w{k}"]).requires_grad_(True) for k in range(n_w)]
        feats = OE.hash_encode(x, table, scal, log2)
        assert np.array_equal(feats.detach().numpy(), g[f"{tag}_feats"])          # same op order: bit-identical
        y = OE.field(x, table, ws, scal, log2, act, grad_scaling=16.0)
        assert np.abs(y.detach().numpy() - g[f"{tag}_y"]).max() <= 1e-6
        grads = torch.autograd.grad((y * torch.from_numpy(g[f"{tag}_cot"])).sum(), [x, table] + ws)
        assert np.abs(grads[0].numpy() - g[f"{tag}_v_x"]).max() <= 1e-5 * np.abs(g[f"{tag}_v_x"]).max()
        vt = np.zeros_like(g[f"{tag}_table"])
        vt[g[f"{tag}_v_table_idx"]] = g[f"{tag}_v_table_val"]
        assert np.abs(grads[1].numpy() - vt).max() <= 1e-5 * np.abs(vt).max()
        for k in range(n_w):
            assert np.abs(grads[2 + k].numpy() - g[f"{tag}_v_w{k}"]).max() <= 1e-5 * np.abs(g[f"{tag}_v_w{k}"]).max()


def test_field_host_glue_matches_reference_code():
    """The torch-only pieces of geosplatting_b200.field (host glue: get_patches, rot2quat, the rotation helper, the level
    scalings) against the reference's own code (ref_field.npz / ref_math.npz / ref_encoding.npz); they need no GPU."""
    from geosplatting_b200 import encoding as E
    from geosplatting_b200.field import GaussianField, get_rotation_from_relative_vectors, rot2quat, safe_normalize
    g = load("ref_field.npz")
    verts, faces = torch.from_numpy(g["vertices"]), torch.from_numpy(g["indices"])
    normals, areas = GaussianField.get_patches(None, verts, faces)
    assert np.abs(normals.numpy() - g["patch_normals"]).max() <= 1e-6
    assert np.abs(areas.numpy() - g["patch_areas"]).max() <= 1e-6 * np.abs(g["patch_areas"]).max()
    # the vertex path's quaternions are rot2quat(rotation from +z to the patch normal)
    R = get_rotation_from_relative_vectors(torch.tensor([0.0, 0.0, 1.0]), normals)
    assert np.abs(rot2quat(R).numpy() - g["v_quats"]).max() <= 2e-6
    m = load("ref_math.npz")
    assert np.abs(rot2quat(torch.from_numpy(m["rots"])).numpy() - m["quats"]).max() <= 1e-6
    assert np.abs(safe_normalize(torch.from_numpy(m["vecs"])).numpy() - m["safe_normalized"]).max() <= 1e-6
    e = load("ref_encoding.npz")
    assert np.array_equal(np.asarray(E.level_scalings(16, 16, 4096), np.float32), e["kd_scalings"])
    with pytest.raises(RuntimeError, match="no CPU path"):
        E.HashEncoding(E.MLP([32, 3]), log2_hashmap_size=4).encode(torch.zeros(2, 3))


def test_flexicubes_fixture_is_a_closed_surface():
    """tests/golden/ref_flexicubes.npz (scripts/make_golden.py section J: the reference's own FlexiCubes code on a
    10^3 grid, as GeoSplatter.get_geometry drives it) is the fixture the SURVEY 8f rank-3 row (flexicubes.py) is held
    to.  Sanity of the fixture itself: a watertight, consistently oriented surface near the SDF's zero level set, finite
    gradients -- and it feeds the MGAdaptor oracle (6 Gaussians per face)."""
    from oracle import mgadapter as OMG
    g = load("ref_flexicubes.npz")
    v, f = g["mesh_vertices"], g["mesh_indices"]
    assert f.min() == 0 and f.max() == v.shape[0] - 1 and f.shape[1] == 3
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    directed = {(int(a), int(b)) for a, b in e}
    assert len(directed) == e.shape[0]                                   # no directed edge twice: consistent orientation
    assert all((b, a) in directed for a, b in directed)                   # every edge has its opposite: watertight
    r = np.linalg.norm(v, axis=1)
    assert 0.35 < r.min() and r.max() < 0.75                              # bumpy sphere of radius ~0.55
    for k in ("v_sdf", "v_deform", "v_weights", "L_dev"):
        assert np.isfinite(g[k]).all()
    assert float(np.abs(g["v_sdf"]).sum()) > 0 and float(np.abs(g["v_weights"]).sum()) > 0
    vt, ft = torch.from_numpy(v), torch.from_numpy(f)
    means, scales, quats, colors, opac, offsets = OMG.make(vt, ft, OMG.vertex_normals(vt, ft))
    assert means.shape[0] == 6 * f.shape[0] and bool(torch.isfinite(scales).all())


def test_flexicubes_oracle_matches_reference_code():
    """oracle/flexicubes.py against the reference's own FlexiCubes code (ref_flexicubes.npz, scripts/make_golden.py
    section J): identical topology (vertex and face ORDER included), vertices / L_dev / entropy to fp32 rounding, and the
    gradients w.r.t. SDF values, grid deformation and the alpha / beta / gamma weights."""
    from oracle import flexicubes as OF
    g = load("ref_flexicubes.npz")
    tbl = {k[4:]: g[k] for k in g if k.startswith("tbl_")}
    R, scale = int(g["resolution"]), float(g["scale"])
    sdf = torch.from_numpy(g["sdf"]).requires_grad_(True)
    deform = torch.from_numpy(g["deform"]).requires_grad_(True)
    weights = torch.from_numpy(g["weights"]).requires_grad_(True)
    cubes = torch.from_numpy(g["cube_indices"])
    verts = torch.from_numpy(g["grid_vertices"]) + deform.tanh() * (0.5 * scale / R)       # geosplat.py:756
    mv, mf, l_dev = OF.dual_marching_cubes(verts, sdf, cubes, (R, R, R), weights[:, :8], weights[:, 8:20],
                                           weights[:, 20:], tbl)
    assert np.array_equal(mf.numpy(), g["mesh_indices"])                 # same faces in the same order
    assert np.abs(mv.detach().numpy() - g["mesh_vertices"]).max() <= 1e-6
    assert l_dev.shape == g["L_dev"].shape and np.abs(l_dev.detach().numpy() - g["L_dev"]).max() <= 1e-6
    ent = OF.entropy(sdf, cubes, tbl)
    assert abs(float(ent) - float(g["entropy"])) <= 1e-6
    loss = (mv * torch.from_numpy(g["cot_vertices"])).sum() + l_dev.mean() * 0.5 + ent * 0.3
    v_sdf, v_def, v_w = torch.autograd.grad(loss, [sdf, deform, weights])
    for a, b, name in ((v_sdf, g["v_sdf"], "sdf"), (v_def, g["v_deform"], "deform"), (v_w, g["v_weights"], "weights")):
        assert np.abs(a.numpy() - b).max() <= 1e-4 * np.abs(b).max(), (name, np.abs(a.numpy() - b).max(), np.abs(b).max())


def test_flexicubes_oracle_rough_sdf_with_inverted_cases():
    """Random SDF on a non-cubic 7 x 6 x 5 grid (ref_flexicubes_rough.npz): six ambiguous configurations are inverted by
    the reference, cubes emit up to four dual vertices -- same faces in the same order, same vertices and gradients."""
    from oracle import flexicubes as OF
    g = load("ref_flexicubes_rough.npz")
    assert int(g["n_inverted"]) > 0
    tbl = {k[4:]: g[k] for k in g if k.startswith("tbl_")}
    sdf = torch.from_numpy(g["sdf"]).requires_grad_(True)
    weights = torch.from_numpy(g["weights"]).requires_grad_(True)
    cubes = torch.from_numpy(g["cube_indices"])
    mv, mf, l_dev = OF.dual_marching_cubes(torch.from_numpy(g["grid_vertices"]), sdf, cubes, tuple(g["resolution"]),
                                           weights[:, :8], weights[:, 8:20], weights[:, 20:], tbl)
    assert np.array_equal(mf.numpy(), g["mesh_indices"])
    assert np.abs(mv.detach().numpy() - g["mesh_vertices"]).max() <= 1e-5
    assert np.abs(l_dev.detach().numpy() - g["L_dev"]).max() <= 1e-5
    assert abs(float(OF.entropy(sdf, cubes, tbl).detach()) - float(g["entropy"])) <= 1e-6
    v_sdf, v_w = torch.autograd.grad((mv * torch.from_numpy(g["cot_vertices"])).sum() + l_dev.mean(), [sdf, weights])
    assert np.abs(v_sdf.numpy() - g["v_sdf"]).max() <= 1e-4 * np.abs(g["v_sdf"]).max()
    assert np.abs(v_w.numpy() - g["v_weights"]).max() <= 1e-4 * np.abs(g["v_weights"]).max()


def test_fg_lut_asset_loader_reads_the_reference_bytes(tmp_path):
    """`shade.load_fg_lut` on the reference's asset = the committed copy of the table = the
    sha256 / subsample / corner values recorded from /root/reference (ref_fg_lut_sub.npz)."""
    import hashlib
    import os

    from geosplatting_b200.shade import load_fg_lut
    full, sub = load("ref_fg_lut.npz"), load("ref_fg_lut_sub.npz")
    lut = full["lut"]
    assert lut.shape == (256, 256, 2) and lut.dtype == np.float32
    assert hashlib.sha256(np.ascontiguousarray(lut).tobytes()).hexdigest() == str(sub["sha256"])
    assert np.array_equal(lut[::8, ::8], sub["sub"])
    assert np.array_equal(np.stack([lut[0, 0], lut[0, 255], lut[255, 0], lut[255, 255]]), sub["corners"])
    assert abs(lut[0, 0, 0] - 0.00973) < 1e-4 and abs(lut[0, 0, 1] - 0.99025) < 1e-4
    p = tmp_path / "lut.bin"
    p.write_bytes(np.ascontiguousarray(lut).tobytes())
    assert np.array_equal(load_fg_lut(str(p), "cpu").numpy(), lut)
    (tmp_path / "short.bin").write_bytes(b"\0" * 100)
    with pytest.raises(ValueError):
        load_fg_lut(str(tmp_path / "short.bin"), "cpu")
    asset = "/root/reference/rfstudio/assets/geometry/pbr/bsdf_256_256.bin"
    if os.path.exists(asset):                     # this container only: the GPU box has no reference tree
        assert np.array_equal(load_fg_lut(asset, "cpu").numpy(), lut)
