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
    xv, yv = [], []
    tri_set, pt_set = set(), set()
    matched = (mask >> 8) & 7
    n_matched = ((matched & 1) + ((matched >> 1) & 1) + ((matched >> 2) & 1))
    total_comparisons = int(np.where(n_matched == 3, 3, np.where(n_matched == 2, 1, 0))[tri_order].sum()) if len(tri_order) else 0
    total_violations = 0
    violated = 0
    hot = [int(t) for t in tri_order if mask[t] & 63]
    for t in hot:
        m = int(mask[t])
        v = tri[t]
        for q, (u, w) in enumerate(_PAIRS):
            v1, v2 = int(v[u]), int(v[w])
            if m & (1 << q):
                j1, j2 = int(match_j[v1]), int(match_j[v2])
                xv.append({"triangle_idx": t,
                           "point1": {"aligned_idx": v1, "ref_idx": j1, "orig_x": a_xy[v1, 0], "matched_x": r_xy[j1, 0]},
                           "point2": {"aligned_idx": v2, "ref_idx": j2, "orig_x": a_xy[v2, 0], "matched_x": r_xy[j2, 0]}})
                pt_set.update([v1, v2])
                total_violations += 1
            if m & (1 << (3 + q)):
                j1, j2 = int(match_j[v1]), int(match_j[v2])
                yv.append({"triangle_idx": t,
                           "point1": {"aligned_idx": v1, "ref_idx": j1, "orig_y": a_xy[v1, 1], "matched_y": r_xy[j1, 1]},
                           "point2": {"aligned_idx": v2, "ref_idx": j2, "orig_y": a_xy[v2, 1], "matched_y": r_xy[j2, 1]}})
                pt_set.update([v1, v2])
                total_violations += 1
        tri_set.add(t)
        violated += 1
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
