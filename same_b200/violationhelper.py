"""`verify_spatial_preservation` / `print_violation_report` with the reference's signatures
(src/violationhelper.py:1-166).  The per-triangle order tests run on the GPU (k_postsolve); this module
only folds the resulting bit masks into the reference's report dictionary."""
from __future__ import annotations


from collections.abc import Sequence

import numpy as np

from . import _lib as L

_PAIRS = ((0, 1), (0, 2), (1, 2))


class LazyRecords(Sequence):
    """The reference's list of per-violation dictionaries (src/violationhelper.py:76-121) as a read-only SEQUENCE whose elements are
    built from flat arrays the first time they are read (iteration, indexing, comparison, pickling ...): a solution with tens of
    thousands of order violations does not pay for tens of thousands of nested dicts unless somebody looks at them.  `len()` is
    free; it compares equal to the plain list, `list(x)` gives the plain list, and it pickles (`np.save(var_out)`) AS a plain list.
    Deliberately not a `list` subclass: C code that reads a list's storage directly (pandas, numpy) would see it empty.
    SAME_B200_EAGER_REPORT=1 makes run_same return plain lists / dicts instead."""

    def __init__(self, n, build, unordered=None):
        self._n, self._build, self._items = int(n), build, None
        self.unordered = unordered      # the same elements as an array in arbitrary order, when the caller has them (no fill needed)

    def _fill(self):
        if self._items is None:
            self._items = list(self._build())
            self._build = None
            assert len(self._items) == self._n
        return self._items

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        return self._fill()[i]

    def __iter__(self):
        return iter(self._fill())

    def __eq__(self, other):
        return self._fill() == (other._fill() if isinstance(other, LazyRecords) else other)

    __hash__ = None

    def __repr__(self):
        return repr(self._fill())

    def __add__(self, other):
        return self._fill() + list(other)

    def __radd__(self, other):
        return list(other) + self._fill()

    def __reduce__(self):
        return (list, (self._fill(),))


def eager_report():
    import os
    return os.environ.get("SAME_B200_EAGER_REPORT", "0") not in ("", "0")


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
    of the reference's `triangle_info` dict: a sequence, or a zero-argument callable that returns it (and then covers every
    triangle) — in that case only the counts are computed now and everything whose ORDER depends on it (the record lists, the
    two `list(set)` results) is a LazyRecords that asks for the order when it is first read."""
    mask = np.asarray(mask)
    tri = np.asarray(tri).reshape(-1, 3)
    match_j = np.asarray(match_j)
    lazy = callable(tri_order) and not eager_report()
    matched = (mask >> 8) & 7
    n_matched = ((matched & 1) + ((matched >> 1) & 1) + ((matched >> 2) & 1))
    comparisons = np.where(n_matched == 3, 3, np.where(n_matched == 2, 1, 0))
    pu = np.array([0, 0, 1]); pw = np.array([1, 2, 2])
    state = {}

    def ordered():
        """Everything that needs the order of the violated triangles, once."""
        if not state:
            order = np.asarray(tri_order() if callable(tri_order) else tri_order, dtype=np.int64)
            hot = order[(mask[order] & 63) != 0] if len(order) else order
            # one record per (violated triangle in `tri_order` order, vertex pair q, axis): the reference appends the x record of a
            # pair, then its y record, pair by pair (src/violationhelper.py:62-121); the lists keep that order per axis
            bits = (mask[hot, None] >> np.arange(6)) & 1                 # [n_hot, 6]: x bits of pairs 0..2, then y bits
            state.update(order=order, hot=hot, bits=bits, v1=tri[hot][:, pu], v2=tri[hot][:, pw])
        return state

    def build_records(axis, name):
        st = ordered()
        rows, q = np.nonzero(st["bits"][:, 3 * axis:3 * axis + 3] != 0)   # row-major: triangle order, then pair order
        t = st["hot"][rows]
        v1, v2 = st["v1"][rows, q], st["v2"][rows, q]
        j1, j2 = match_j[v1], match_j[v2]
        o1, o2, m1, m2 = a_xy[v1, axis], a_xy[v2, axis], r_xy[j1, axis], r_xy[j2, axis]
        ok, mk = f"orig_{name}", f"matched_{name}"
        return [{"triangle_idx": tt, "point1": {"aligned_idx": a1, "ref_idx": b1, ok: c1, mk: d1},
                 "point2": {"aligned_idx": a2, "ref_idx": b2, ok: c2, mk: d2}}
                for tt, a1, b1, c1, d1, a2, b2, c2, d2 in zip(t.tolist(), v1.tolist(), j1.tolist(), o1, m1, v2.tolist(), j2.tolist(), o2, m2)]

    # (the reference collects these two in sets and returns list(set): the iteration order of a set of small ints, reproduced by
    # inserting in the same order — triangle by triangle, pair by pair)
    def build_tri_list():
        return list(set(ordered()["hot"].tolist()))

    def build_pt_list():
        st = ordered()
        rows, q = np.nonzero(st["bits"][:, :3] | st["bits"][:, 3:])
        pts = set()
        for v in np.stack([st["v1"][rows, q], st["v2"][rows, q]], axis=1).reshape(-1).tolist():
            pts.add(v)
        return list(pts)

    if lazy:
        all_bits = (mask[:, None] >> np.arange(6)) & 1
        n_x, n_y = int(all_bits[:, :3].sum()), int(all_bits[:, 3:].sum())
        hot_unordered = np.flatnonzero((mask & 63) != 0)
        any_pair = (all_bits[:, :3] | all_bits[:, 3:]) != 0
        rows, q = np.nonzero(any_pair)
        pts_unordered = np.unique(np.concatenate([tri[rows, pu[q]], tri[rows, pw[q]]])) if len(rows) else np.zeros(0, np.int64)
        total_comparisons, violated = int(comparisons.sum()), int(len(hot_unordered))
        xv, yv = LazyRecords(n_x, lambda: build_records(0, "x")), LazyRecords(n_y, lambda: build_records(1, "y"))
        tri_list = LazyRecords(violated, build_tri_list, unordered=hot_unordered)
        pt_list = LazyRecords(len(pts_unordered), build_pt_list, unordered=pts_unordered)
    else:
        st = ordered()
        total_comparisons = int(comparisons[st["order"]].sum()) if len(st["order"]) else 0
        violated = int(len(st["hot"]))
        xv, yv = build_records(0, "x"), build_records(1, "y")
        tri_list, pt_list = build_tri_list(), (build_pt_list() if len(xv) + len(yv) else [])
    total_violations = len(xv) + len(yv)
    summary = {"total_triangles": int(n_triangle_info), "violated_triangles": violated, "total_comparisons": total_comparisons,
               "total_violations": total_violations}
    summary["percent_triangles_violated"] = violated / summary["total_triangles"] * 100 if summary["total_triangles"] > 0 else 0
    summary["percent_violations"] = total_violations / total_comparisons * 100 if total_comparisons > 0 else 0
    return {"x_order_violations": xv, "y_order_violations": yv, "triangles_with_violations": tri_list,
            "points_with_violations": pt_list, "violation_summary": summary}


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
