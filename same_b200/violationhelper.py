"""`verify_spatial_preservation` / `print_violation_report` with the reference's signatures
(src/violationhelper.py:1-166).  The per-triangle order tests run on the GPU (k_postsolve); this module
only folds the resulting bit masks into the reference's report dictionary."""
from __future__ import annotations


import numpy as np

from . import _lib as L

_PAIRS = ((0, 1), (0, 2), (1, 2))


def postsolve_arrays(tri, a_xy, r_xy, match_j, device=None):
    """Stateless GPU call (same_postsolve_arrays) -> mask, area_before, area_after, flipped."""
    tri = np.ascontiguousarray(tri, dtype=np.int32).reshape(-1, 3)
    a_xy = np.ascontiguousarray(a_xy, dtype=np.float64).reshape(-1, 2)
    r_xy = np.ascontiguousarray(r_xy, dtype=np.float64).reshape(-1, 2)
    mj = np.ascontiguousarray(match_j, dtype=np.int32)
    t = len(tri)
    mask, ab, aa, fl = np.zeros(t, np.int32), np.zeros(t), np.zeros(t), np.zeros(t, np.uint8)
    from .device import default_device
    device = default_device() if device is None else device
    L.check(L.load().same_postsolve_arrays(device, t, L.ptr(tri), len(a_xy), L.ptr(a_xy), len(r_xy), L.ptr(r_xy), L.ptr(mj),
                                           L.ptr(mask), L.ptr(ab), L.ptr(aa), L.ptr(fl)))
    return mask, ab, aa, fl.astype(bool)


def violations_from_mask(mask, tri, match_j, a_xy, r_xy, tri_order, n_triangle_info):
    """Fold per-triangle masks (bits 0-2 x-order, 3-5 y-order violations of vertex pairs (0,1),(0,2),(1,2);
    bits 8-10 vertex matched) into the dictionary of src/violationhelper.py:24-134.  `tri_order` = iteration order
    of the reference's `triangle_info` dict."""
    mask = np.asarray(mask)
    tri = np.asarray(tri).reshape(-1, 3)
    matched = (mask >> 8) & 7
    n_matched = ((matched & 1) + ((matched >> 1) & 1) + ((matched >> 2) & 1))
    order = np.asarray(tri_order, dtype=np.int64)
    total_comparisons = int(np.where(n_matched == 3, 3, np.where(n_matched == 2, 1, 0))[order].sum()) if len(order) else 0
    hot = order[(mask[order] & 63) != 0] if len(order) else order
    # one record per (violated triangle in `tri_order` order, vertex pair q, axis): the reference appends the x record of a pair,
    # then its y record, pair by pair (src/violationhelper.py:62-121); the lists below keep that order per axis
    bits = (mask[hot, None] >> np.arange(6)) & 1                 # [n_hot, 6]: x bits of pairs 0..2, then y bits
    pu = np.array([0, 0, 1]); pw = np.array([1, 2, 2])
    v1_all, v2_all = tri[hot][:, pu], tri[hot][:, pw]            # [n_hot, 3]
    match_j = np.asarray(match_j)

    def records(axis, name):
        sel = bits[:, 3 * axis:3 * axis + 3] != 0
        rows, q = np.nonzero(sel)                                # row-major: triangle order, then pair order
        t = hot[rows]
        v1, v2 = v1_all[rows, q], v2_all[rows, q]
        j1, j2 = match_j[v1], match_j[v2]
        o1, o2, m1, m2 = a_xy[v1, axis], a_xy[v2, axis], r_xy[j1, axis], r_xy[j2, axis]
        ok, mk = f"orig_{name}", f"matched_{name}"
        recs = [{"triangle_idx": tt, "point1": {"aligned_idx": a1, "ref_idx": b1, ok: c1, mk: d1},
                 "point2": {"aligned_idx": a2, "ref_idx": b2, ok: c2, mk: d2}}
                for tt, a1, b1, c1, d1, a2, b2, c2, d2 in zip(t.tolist(), v1.tolist(), j1.tolist(), o1, m1, v2.tolist(), j2.tolist(), o2, m2)]
        return recs, v1, v2
    xv, xa, xb = records(0, "x")
    yv, ya, yb = records(1, "y")
    total_violations = len(xv) + len(yv)
    violated = int(len(hot))
    # (the reference collects these in sets and returns list(set): iteration order of a set of small ints, reproduced by
    # inserting in the same order — triangle by triangle, pair by pair, x before y)
    tri_set = set(hot.tolist())
    pt_set = set()
    if total_violations:
        sel_any = bits[:, :3] | bits[:, 3:]
        rows, q = np.nonzero(sel_any)
        seq = np.stack([v1_all[rows, q], v2_all[rows, q]], axis=1).reshape(-1)
        pt_set = set()
        for v in seq.tolist():
            pt_set.add(v)
    summary = {"total_triangles": int(n_triangle_info), "violated_triangles": violated, "total_comparisons": total_comparisons,
               "total_violations": total_violations}
    summary["percent_triangles_violated"] = violated / summary["total_triangles"] * 100 if summary["total_triangles"] > 0 else 0
    summary["percent_violations"] = total_violations / total_comparisons * 100 if total_comparisons > 0 else 0
    return {"x_order_violations": xv, "y_order_violations": yv, "triangles_with_violations": list(tri_set),
            "points_with_violations": list(pt_set), "violation_summary": summary}


def verify_spatial_preservation(aligned_df, ref_df, matches_df, triangle_info, tolerance=1e-6):
    """For every triangle with >= 2 matched vertices, does each vertex pair keep its x (and y) order after
    matching?  Same report dictionary as the reference (`tolerance` is unused there too)."""
    a_xy = aligned_df[["X", "Y"]].to_numpy(dtype=np.float64)
    r_xy = ref_df[["X", "Y"]].to_numpy(dtype=np.float64)
    match_j = np.full(len(aligned_df), -1, np.int32)
    if len(matches_df):
        match_j[matches_df["aligned_idx"].to_numpy(dtype=np.int64)] = matches_df["ref_idx"].to_numpy(dtype=np.int32)  # later rows win
    order = list(triangle_info.keys())
    n_tri = (max(order) + 1) if order else 0
    tri = np.zeros((n_tri, 3), np.int32)
    for s, info in triangle_info.items():
        tri[s] = np.asarray(info["vertices"], dtype=np.int32)
    mask, _, _, _ = postsolve_arrays(tri, a_xy, r_xy, match_j)
    return violations_from_mask(mask, tri, match_j, a_xy, r_xy, order, len(triangle_info))


def print_violation_report(violations):
    s = violations["violation_summary"]
    print("\nSpatial Preservation Violation Report")
    print("=====================================")
    print(f"Total triangles analyzed: {s['total_triangles']}")
    print(f"Triangles with violations: {s['violated_triangles']} ({s['percent_triangles_violated']:.2f}%)")
    print(f"Total position comparisons: {s['total_comparisons']}")
    print(f"Total violations found: {s['total_violations']} ({s['percent_violations']:.2f}%)")
    print(f"Number of points involved in violations: {len(violations['points_with_violations'])}")
