"""CPU restatement of FlexiCubes dual marching cubes + the SDF entropy regulariser, the step that produces the mesh
MGAdaptor samples.  TEST INFRASTRUCTURE ONLY: the checker of geosplatting_b200/flexicubes.py
+ csrc/flexicubes.cu (tests/test_flexicubes_cpu.py, tests/test_flexicubes_gpu.py); never imported by the product.

Follows rfstudio/graphics/_mesh/_flexicubes.py:460-506 (ambiguity resolution), :508-538 (surface edges),
:559-713 (dual vertices, L_dev), :715-725 (entropy), :727-802 (regulariser, triangulation) as GeoSplatter.get_geometry
drives them (rfstudio/model/geosplat.py:751-769: deformed grid vertices, alpha / beta / gamma from one [F,21] parameter).
Written around the structure a GPU implementation needs, not around the reference's op sequence: the topology (integer
work: cases, edge ranks, dual-vertex numbering, quads) is explicit numpy with every output ORDER stated, the
differentiable arithmetic is torch so that autograd yields the gradients to compare.  The four lookup tables (cube edges,
ambiguity check, dual-vertex groups, dual-vertex counts) are data of the published algorithm and are passed in; the
fixture stores them by value.  Pinned on the fixture by tests/test_golden_cpu.py::test_flexicubes_oracle_matches_reference_code.

Output orders (all of them matter: MGAdaptor emits Gaussians in face order):
  surface cubes        ascending cube index
  surface edges        ascending (v_a, v_b) of the grid edges with a sign change
  dual vertices        for k = 1..4: the surface cubes that emit k dual vertices, ascending, k vertices each
  L_dev entries        same loop, per cube group-major then the group's edges in table order
  quads                surface edges shared by 4 cubes, ascending; those whose first endpoint has sdf > 0 first
  mesh vertices        dual vertices, then one centre per quad;  faces: 4 per quad, (q0,q1,c) (q1,q2,c) (q2,q3,c) (q3,q0,c)
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
from torch import Tensor

WEIGHT_SCALE = 0.99


def _topology(sdf: np.ndarray, cubes: np.ndarray, res: Tuple[int, int, int], tbl: Dict[str, np.ndarray]):
    occ = sdf.reshape(-1) < 0
    n_occ = occ[cubes].sum(1)
    surf = np.nonzero((n_occ > 0) & (n_occ < 8))[0]                     # surface cubes, ascending
    case = (occ[cubes[surf]] * (1 << np.arange(8))).sum(1)

    # ambiguous configurations: flip to the alternative case when the neighbour across the ambiguous face is ambiguous too
    cfg = tbl["check"][case]                                             # [N,5]: flag, offset(3), alternative case
    flagged = np.zeros(tuple(res), dtype=bool)
    pos = np.stack(np.unravel_index(surf, tuple(res)), 1)                # the reference's nonzero() order of the cube volume
    amb = cfg[:, 0] == 1
    flagged[tuple(pos[amb].T)] = True
    nb = pos + cfg[:, 1:4]
    inside = amb & (nb >= 0).all(1) & (nb < np.asarray(res)).all(1)
    flip = np.zeros_like(amb)
    flip[inside] = flagged[tuple(nb[inside].T)]
    case = np.where(flip, cfg[:, 4], case)

    # grid edges of the surface cubes; rank of the sign-changing ones in ascending (v_a, v_b)
    ce = tbl["cube_edges"].reshape(12, 2)
    ends = cubes[surf][:, ce]                                            # [N,12,2]
    V = sdf.shape[0]
    key = ends[..., 0].astype(np.int64) * V + ends[..., 1]
    uniq, inv, cnt = np.unique(key.reshape(-1), return_inverse=True, return_counts=True)
    ua, ub = uniq // V, uniq % V
    crossing = occ[ua] != occ[ub]
    rank = np.full(uniq.shape[0], -1, np.int64)
    rank[crossing] = np.arange(int(crossing.sum()))
    edge_of = rank[inv].reshape(-1, 12)                                  # [N,12] surface-edge id or -1
    shared = cnt[inv].reshape(-1, 12)                                    # number of surface cubes around the edge
    surf_edges = np.stack([ua[crossing], ub[crossing]], 1)               # [E,2]

    # dual vertices and their edge groups
    num_vd = tbl["num_vd"][case]
    grp_cube, grp_edge, grp_vd, vd_cube = [], [], [], []
    total = 0
    for k in range(1, 5):
        for c in np.nonzero(num_vd == k)[0]:
            for j in range(k):
                for e in tbl["dmc"][case[c], j]:
                    if e != -1:
                        grp_cube.append(c); grp_edge.append(int(e)); grp_vd.append(total)
                vd_cube.append(c)
                total += 1
    grp_cube, grp_edge, grp_vd = np.asarray(grp_cube), np.asarray(grp_edge), np.asarray(grp_vd)
    vd_of = np.zeros((surf.shape[0], 12), np.int64)
    vd_of[grp_cube, grp_edge] = grp_vd

    # quads around the surface edges that four surface cubes share
    sel = (shared == 4) & (edge_of >= 0)
    e_id, v_id = edge_of[sel], vd_of[sel]                                # (cube, local edge) order
    order = np.argsort(e_id, kind="stable")
    quad_edge = e_id[order].reshape(-1, 4)[:, 0]
    quad = v_id[order].reshape(-1, 4)
    first_positive = sdf.reshape(-1)[surf_edges[quad_edge, 0]] > 0
    quad = np.concatenate([quad[first_positive][:, [0, 1, 3, 2]], quad[~first_positive][:, [2, 3, 1, 0]]])
    return dict(surf=surf, surf_edges=surf_edges, edge_of=edge_of, grp_cube=grp_cube, grp_edge=grp_edge, grp_vd=grp_vd,
                vd_cube=np.asarray(vd_cube), n_vd=total, quad=quad, cube_edges=ce)


def _crossing(s: Tensor, x: Tensor) -> Tensor:
    """Zero crossing of the linear interpolant: s [...,2], x [...,2,3]."""
    w_b = s[..., 0] / (s[..., 0] - s[..., 1])
    return x[..., 1, :] * w_b[..., None] + x[..., 0, :] * (1 - w_b[..., None])


def dual_marching_cubes(vertices: Tensor, sdf: Tensor, cubes: Tensor, res, alpha: Tensor, beta: Tensor, gamma: Tensor,
                        tbl: Dict[str, np.ndarray]):
    """vertices [V,3] (already deformed), sdf [V,1], cubes [F,8], alpha [F,8], beta [F,12], gamma [F,1] (raw
    parameters) -> (mesh vertices [Q+quads,3], faces [4*quads,3] int64, L_dev [K])."""
    t = _topology(sdf.detach().numpy(), cubes.numpy(), tuple(int(r) for r in res), tbl)
    surf = torch.from_numpy(t["surf"])
    a_act = alpha[surf].tanh() * WEIGHT_SCALE + 1
    b_act = beta[surf].tanh() * WEIGHT_SCALE + 1
    g_act = gamma[surf].sigmoid() * WEIGHT_SCALE + (1 - WEIGHT_SCALE) / 2
    gc, ge, gv = (torch.from_numpy(t[k]) for k in ("grp_cube", "grp_edge", "grp_vd"))
    ends = torch.from_numpy(t["surf_edges"][t["edge_of"][t["grp_cube"], t["grp_edge"]]])        # [K,2] grid vertices
    x, s = vertices[ends], sdf[ends, 0]                                                       # [K,2,3], [K,2]
    corners = torch.from_numpy(t["cube_edges"][t["grp_edge"]])                                 # [K,2] local corners
    ue = _crossing(s * a_act[gc[:, None], corners], x)
    w = b_act[gc, ge][:, None]
    Q = t["n_vd"]
    vd = torch.zeros(Q, 3).index_add_(0, gv, ue * w) / torch.zeros(Q, 1).index_add_(0, gv, w)
    # L_dev: absolute deviation of every edge's plain zero crossing from the mean distance to its dual vertex
    dist = (_crossing(s, x) - vd[gv]).norm(dim=-1)
    n_edges = torch.zeros(Q).index_add_(0, gv, torch.ones_like(dist))
    l_dev = (dist - (torch.zeros(Q).index_add_(0, gv, dist) / n_edges)[gv]).abs()
    quad = torch.from_numpy(t["quad"])
    vg = g_act[torch.from_numpy(t["vd_cube"]), 0][quad]                                        # [quads,4]
    g02, g13 = vg[:, 0] * vg[:, 2], vg[:, 1] * vg[:, 3]
    vq = vd[quad]
    centre = ((vq[:, 0] + vq[:, 2]) / 2 * g02[:, None] + (vq[:, 1] + vq[:, 3]) / 2 * g13[:, None]) / \
             ((g02 + g13) + 1e-8)[:, None]
    c_idx = torch.arange(quad.shape[0]) + Q
    faces = torch.stack([torch.stack([quad[:, i], quad[:, (i + 1) % 4], c_idx], -1) for i in range(4)], 1).reshape(-1, 3)
    return torch.cat([vd, centre]), faces, l_dev


def entropy(sdf: Tensor, cubes: Tensor, tbl: Dict[str, np.ndarray]) -> Tensor:
    """_flexicubes.py:715-725: symmetric BCE between the two endpoint values of every sign-changing grid edge."""
    ce = tbl["cube_edges"].reshape(12, 2)
    ends = cubes.numpy()[:, ce].reshape(-1, 2)
    uniq = np.unique(ends, axis=0)
    s = sdf[:, 0]
    occ = (s < 0).numpy()
    e = torch.from_numpy(uniq[occ[uniq[:, 0]] != occ[uniq[:, 1]]])
    a, b = s[e[:, 0]], s[e[:, 1]]
    bce = torch.nn.functional.binary_cross_entropy_with_logits
    return bce(a, (b > 0).float()) + bce(b, (a > 0).float())
