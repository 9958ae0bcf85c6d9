"""Per-window orchestration of the CPU oracle — the order of operations of the reference's
`run_same` (src/same.py:972-1197) restated over plain arrays.  TEST INFRASTRUCTURE ONLY
(see oracle/same_oracle.c); never imported by `same_b200/`.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import Delaunay

from . import oracle as O


def window_pipeline(a_xy, r_xy, a_prob, r_prob, a_type, r_type, a_size, r_size, *, radius, knn,
                    dist_ct_coeff=1.0, min_angle_deg=15, ignore_same_type_triangles=True,
                    ignore_knn_if_matched=False, max_matches=1, ref_metacell_match_multiplier=None,
                    tri_global=None, a_vid=None, brute=False):
    """Returns a dict of arrays describing the model `run_same` would build.

    keepA/keepR are rows of the INPUT frames that survive (after KNN compaction, utils.py:733-742,
    and — with a precomputed triangulation — unconstrained-node removal, same.py:1056-1085)."""
    a_xy = np.asarray(a_xy, np.float64).reshape(-1, 2)
    r_xy = np.asarray(r_xy, np.float64).reshape(-1, 2)
    out = {}
    keepA, keepR, pairs = O.find_knn_within_radius(a_xy, r_xy, radius, knn, brute=brute)      # same.py:972-979
    if ignore_knn_if_matched:
        pairs = O.knn_priority(pairs, np.asarray(a_type)[keepA], np.asarray(r_type)[keepR])
    out["n_pairs_knn"] = len(pairs)
    if len(pairs) == 0:
        raise ValueError("No valid_pairs after KNN filtering. Increase radius and/or knn.")     # same.py:1002-1003
    axy = a_xy[keepA]
    using_precomputed = tri_global is not None
    if not using_precomputed:
        tri = Delaunay(axy).simplices.astype(np.int32)                                          # same.py:1023
    else:
        tri, _ = O.remap_triangles(tri_global, np.asarray(a_vid)[keepA])                        # same.py:1028-1031
    kept_src, unc, band = O.filter_triangles(axy, tri, radius, min_angle_deg, np.asarray(a_type)[keepA],
                                             ignore_same_type_triangles)                         # same.py:1034-1053
    tri = tri[kept_src]
    out["n_band"] = band
    if using_precomputed and len(unc):                                                           # same.py:1056-1085
        un = np.zeros(len(keepA), bool)
        un[unc] = True
        pairs = pairs[~un[pairs[:, 0]]]
        old_to_new = np.cumsum(~un) - 1
        pairs = np.column_stack([old_to_new[pairs[:, 0]], pairs[:, 1]]).astype(np.int32)
        tri = tri[~un[tri].any(axis=1)] if len(tri) else tri
        tri = old_to_new[tri].astype(np.int32) if len(tri) else tri
        keepA = keepA[~un]
        axy = a_xy[keepA]
    out.update(keepA=keepA, keepR=keepR, pairs=pairs, tri=tri.reshape(-1, 3))
    rxy = r_xy[keepR]
    tt = O.tri_tables(axy, np.asarray(a_size, np.float64)[keepA], tri)
    out.update(weight=tt["weight"], sign=tt["sign"], bounds=tt["bounds"], argv=tt["argv"])
    out["cost"] = O.pair_cost(pairs, axy, rxy, np.asarray(a_prob, np.float64)[keepA], np.asarray(r_prob, np.float64)[keepR],
                              dist_ct_coeff)
    gnode, gptr, gidx = O.group_pairs(pairs, 1, len(keepR))
    rs = np.asarray(r_size, np.float64)[keepR]
    has_mc = bool((rs > 1).any())                                                                # helpers.py:121
    mult = ref_metacell_match_multiplier
    if has_mc and mult is None:
        mult = int(rs.max())
    limit = np.where(has_mc & (rs[gnode] > 1), (mult if mult is not None else 1) * max_matches, max_matches)
    out.update(ref_group_node=gnode, ref_group_ptr=gptr, ref_group_idx=gidx, ref_group_limit=limit)
    anode, aptr, aidx = O.group_pairs(pairs, 0, len(keepA))
    out.update(al_group_node=anode, al_group_ptr=aptr, al_group_idx=aidx)
    return out


def constraints_from_groups(res, n_x):
    """The constraint list add_basic_constraints_optimized emits (helpers.py:130-158), as
    (names, sense, rhs, ptr, idx, val) with variables numbered x[0..P), penalty[P..P+Nr), no_match[...]."""
    P = n_x
    nr, na = len(res["keepR"]), len(res["keepA"])
    names, sense, rhs, ptr, idx, val = [], [], [], [0], [], []

    def emit(nm, s, r, ids, vals):
        names.append(nm); sense.append(s); rhs.append(r)
        idx.extend(ids); val.extend(vals); ptr.append(len(idx))

    gn, gp, gi = res["ref_group_node"], res["ref_group_ptr"], res["ref_group_idx"]
    an, ap, ai = res["al_group_node"], res["al_group_ptr"], res["al_group_idx"]
    for g, j in enumerate(gn):
        m = gi[gp[g]:gp[g + 1]]
        emit(f"max_matches_{j}", "<=", float(res["ref_group_limit"][g]), m.tolist(), [1.0] * len(m))
    for g, i in enumerate(an):
        m = ai[ap[g]:ap[g + 1]]
        emit(f"one_match_{i}", "<=", 1.0, m.tolist(), [1.0] * len(m))
    for g, j in enumerate(gn):
        m = gi[gp[g]:gp[g + 1]]
        emit(f"penalty_{j}", "<=", 1.0, m.tolist() + [P + int(j)], [1.0] * len(m) + [-1.0])
    for g, i in enumerate(an):
        m = ai[ap[g]:ap[g + 1]]
        emit(f"no_match_{i}", "==", 1.0, m.tolist() + [P + nr + int(i)], [1.0] * len(m) + [1.0])
    return (np.asarray(names), np.asarray(sense), np.asarray(rhs), np.asarray(ptr, np.int64), np.asarray(idx, np.int64),
            np.asarray(val))
