"""Host-side mirror of the reference's helper functions (src/helpers.py), array work on the GPU.

Same names, argument meaning and return types as the reference so call sites and tests read alike:
`filter_triangles_by_radius` (:233-395), `precompute_triangle_info` (:184-210), `calculate_signed_area`
(:73-77), `get_unprocessed_windows` (:21-70), `load_matching_results` (:667-689).
"""
from __future__ import annotations

import os
from collections.abc import Mapping

import numpy as np
import pandas as pd

from . import _lib as L
from .device import Section


def calculate_signed_area(p1, p2, p3):
    """0.5 * (x1(y2-y3) + x2(y3-y1) + x3(y1-y2))  (src/helpers.py:73-77) — scalar convenience, not a hot path."""
    x1, y1 = p1
    x2, y2 = p2
    x3, y3 = p3
    return 0.5 * (x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2))


# ---------------------------------------------------------------------------------------------------------
# exact-predicate diagnostic (north_star "with an exact-predicate fallback"; SURVEY.md §7 hard part 2: report only)
# ---------------------------------------------------------------------------------------------------------
def naive_orientation_sign(a, b, c):
    """sign((bx-ax)*(cy-ay) - (by-ay)*(cx-ax)) one float64 operation at a time — the reference's arithmetic
    (src/same.py:658, :1146) and what the kernels evaluate."""
    ax, ay, bx, by, cx, cy = (np.float64(v) for v in (a[0], a[1], b[0], b[1], c[0], c[1]))
    d = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)
    return int(d > 0) - int(d < 0)


def exact_orientation_sign(a, b, c):
    """Sign of the same determinant in exact rational arithmetic (every float64 is a rational number)."""
    from fractions import Fraction as F
    ax, ay, bx, by, cx, cy = (F(float(v)) for v in (a[0], a[1], b[0], b[1], c[0], c[1]))
    d = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)
    return int(d > 0) - int(d < 0)


def exact_predicate_check(batch, w, which, a_xy, r_xy, match_j=None):
    """Window `w`: of the triangles whose naive orientation the device filter could not certify (`WindowBatch.uncertain`), how
    many have a naive sign that differs from the exact one?  which = 0: source signs on the window's aligned coordinates `a_xy`
    (kept rows); which = 1: the last separation call, on `r_xy[match_j]`.  The sets the path computes are NEVER changed by this:
    bit-exact parity with the reference means reproducing its rounding.  -> dict(uncertain, listed, naive_differs_from_exact)"""
    n, idx = batch.uncertain(which)
    out = dict(uncertain=0, listed=0, naive_differs_from_exact=0, triangles=[])
    if n == 0:
        return out
    t_off = batch.offsets(L.TRI)
    mine = idx[(idx >= t_off[w]) & (idx < t_off[w + 1])]
    out["uncertain"] = int(len(mine)) if n <= len(idx) else int(n)       # (beyond the device's list capacity only the total is known)
    out["listed"] = int(len(mine))
    tri_w = batch.get_window(L.TRI, w) if len(mine) > 4 else None
    for t in mine:
        v = tri_w[int(t - t_off[w])] if tri_w is not None else batch.get(L.TRI, int(t), int(t) + 1)[0]
        if which == 0:
            p = a_xy[v]
        else:
            j = np.asarray(match_j)[v]
            if (j < 0).any():
                continue
            p = r_xy[j]
        if naive_orientation_sign(p[0], p[1], p[2]) != exact_orientation_sign(p[0], p[1], p[2]):
            out["naive_differs_from_exact"] += 1
            out["triangles"].append(int(t - t_off[w]))
    return out


# ---------------------------------------------------------------------------------------------------------
# guard band: the GPU decides every triangle whose side / angle is not within a few ulps of a threshold; the
# (normally empty) in-band list is re-decided here with the very numpy expressions the reference evaluates
# (BLAS dot / libm arccos are host-dependent at the last ulp; SURVEY.md §7 hard part 1, App. A.4)
# ---------------------------------------------------------------------------------------------------------
def _reference_class(p1, p2, p3, radius, min_angle_deg, same_type, ignore_same_type):
    def angle(a, b, c):          # angle at b (src/helpers.py:278-288)
        v1, v2 = a - b, c - b
        n1, n2 = np.linalg.norm(v1), np.linalg.norm(v2)
        if n1 == 0 or n2 == 0:
            return 0
        return np.degrees(np.arccos(np.clip(np.dot(v1, v2) / (n1 * n2), -1, 1)))
    s1, s2, s3 = np.linalg.norm(p2 - p1), np.linalg.norm(p3 - p2), np.linalg.norm(p1 - p3)
    if max(s1, s2, s3) >= radius:
        return L.TRI_DROP_RADIUS
    if min_angle_deg is not None and min(angle(p2, p1, p3), angle(p1, p2, p3), angle(p1, p3, p2)) < min_angle_deg:
        return L.TRI_DROP_ANGLE
    return L.TRI_SAME_TYPE if (ignore_same_type and same_type) else L.TRI_KEEP


def redecide_band(batch, section_xy, section_type, radius, min_angle_deg, ignore_same_type):
    """Re-classify the guard-band triangles of `batch` on the host and push the classes back (same_batch_tri_override)."""
    band = np.sort(batch.get(L.TRI_BAND))
    if len(band) == 0:
        return 0
    tin_off, ka_off = batch.offsets(L.TRI_IN), batch.offsets(L.KEEP_A)
    keepA = batch.get(L.KEEP_A)
    cls = np.empty(len(band), np.uint8)
    for n, t in enumerate(band):
        w = int(np.searchsorted(tin_off, t, side="right") - 1)
        v = batch.get(L.TRI_IN, int(t), int(t) + 1)[0]
        rows = keepA[ka_off[w] + v]
        p = section_xy[rows]
        same = section_type is not None and len(set(section_type[rows].tolist())) == 1
        cls[n] = _reference_class(p[0], p[1], p[2], radius, min_angle_deg, same, ignore_same_type)
    batch.tri_override(band, cls)
    return len(band)


def filter_triangles_by_radius(points, triangles, radius, aligned_df=None, ignore_same_type_triangles=False,
                               ensure_min_triangle_per_node=True, remove_unconstrained_nodes=False, min_angle_deg=15):
    """Keep triangles whose longest side is < radius, whose smallest angle is >= min_angle_deg and (optionally)
    whose vertices are not all of one `cell_type`; nodes that would lose every triangle get their smallest-perimeter
    same-type triangle back (appended in node order).  Returns a list of triangles (rows of the input, input
    order first) and, with `remove_unconstrained_nodes`, the set of nodes with no radius+angle-valid triangle."""
    points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
    tri_in = np.asarray(triangles)
    tri = np.ascontiguousarray(tri_in, dtype=np.int32).reshape(-1, 3)
    n = len(points)
    types = None
    use_types = ignore_same_type_triangles and aligned_df is not None
    if use_types:
        types, _ = pd.factorize(aligned_df["cell_type"].to_numpy(), use_na_sentinel=False)
        types = types.astype(np.int32)
    with Section(points, points, np.zeros((n, 0)), np.zeros((n, 0)), types, types) as sec, sec.batch() as b:
        b.candidates(0.0, 1)                 # every point keeps itself -> row indices stay 0..n-1
        if b.length(L.KEEP_A) != n:          # duplicate coordinates collapse nothing (each finds itself), but be explicit
            raise RuntimeError("internal: identity candidates did not keep every row")
        b.triangles_set(tri, [0, len(tri)])
        if b.tri_classify(radius, min_angle_deg, use_types) > 0:
            redecide_band(b, points, types, radius, min_angle_deg, use_types)
        b.tri_finalize(use_types, ensure_min_triangle_per_node, remove_unconstrained=False)
        src = b.get(L.TRI_SRC)
        unc = b.get(L.UNCONSTRAINED)
    filtered = [tri_in[k] for k in src]       # the reference returns the input's own rows (helpers.py:342,381)
    if remove_unconstrained_nodes:
        return filtered, set(int(v) for v in unc)
    return filtered


class LazyDict(Mapping):
    """A read-only MAPPING whose items are produced by `build()` the first time anything but its length is asked for (same idea
    as violationhelper.LazyRecords): equal to the plain dict, `dict(x)` gives it, and it pickles AS a plain dict."""

    def __init__(self, n, build):
        self._n, self._build, self._items = int(n), build, None

    def _fill(self):
        if self._items is None:
            self._items = dict(self._build())
            self._build = None
        return self._items

    def __len__(self):
        return self._n

    def __getitem__(self, k):
        return self._fill()[k]

    def __iter__(self):
        return iter(self._fill())

    def __repr__(self):
        return repr(self._fill())

    def __reduce__(self):
        return (dict, (self._fill(),))


def triangle_info_order(n_nodes, aligned_simplex_map):
    """Key order of the reference's triangle_info dict (src/helpers.py:190-195): first appearance while walking nodes 0..n-1 and
    each node's simplex SET in its iteration order."""
    seen, order = set(), []
    for ip in range(n_nodes):
        for s in aligned_simplex_map[ip]:
            if s not in seen:
                seen.add(s)
                order.append(s)
    return order


def precompute_triangle_info(aligned_df, aligned_delaunay, aligned_simplex_map, bounds=None, argv=None, order=None, lazy=False, n_entries=None):
    """dict[simplex] -> vertices, bounds, arg-min/max vertices (src/helpers.py:184-210).  Key order follows the
    reference: first appearance while walking nodes 0..n-1 and each node's simplex set.  `bounds`/`argv` are the
    GPU tables (TRI_BOUNDS / TRI_ARGV); when omitted they are computed here with numpy.  `order` = a precomputed
    `triangle_info_order` (or, with `lazy=True`, a callable returning it); `lazy=True` returns a mapping that builds its entries
    on first access (`n_entries` = its length when the order is not known yet: every triangle has an entry)."""
    tri = np.asarray(aligned_delaunay).reshape(-1, 3)
    if bounds is None or argv is None:
        xy = aligned_df[["X", "Y"]].to_numpy(dtype=np.float64)
        px, py = xy[tri, 0], xy[tri, 1]
        bounds = np.stack([px.min(1), px.max(1), py.min(1), py.max(1)], axis=1)
        pick = lambda m: tri[np.arange(len(tri)), m.argmax(1)]
        argv = np.stack([pick(px == bounds[:, 1:2]), pick(px == bounds[:, 0:1]), pick(py == bounds[:, 3:4]), pick(py == bounds[:, 2:3])], axis=1)
    bounds, argv = np.asarray(bounds), np.asarray(argv)
    if order is None:
        order = triangle_info_order(len(aligned_df), aligned_simplex_map)

    def build():
        info = {}
        for s in (order() if callable(order) else order):
            b, a = bounds[s], argv[s]
            info[s] = {"vertices": aligned_delaunay[s],
                       "bounds": {"min_x": b[0], "max_x": b[1], "min_y": b[2], "max_y": b[3]},
                       "max_x_vertex": a[0], "min_x_vertex": a[1], "max_y_vertex": a[2], "min_y_vertex": a[3]}
        return info
    if not lazy:
        return build()
    return LazyDict(n_entries if callable(order) else len(order), build)


def get_unprocessed_windows(moving_df, output_name, x_windows, y_windows, window_size, overlap, cell_id_col="Cell_Num_Old",
                            counts=None):
    """Resume support (src/helpers.py:21-70): windows that contain moving cells minus windows already present in
    `output_name` (`window_id -> (id % nx, id // nx)`).  `counts[(i, j)]` may carry GPU-computed cell counts."""
    all_windows = set()
    xs, ys = moving_df["X"].to_numpy(), moving_df["Y"].to_numpy()
    for i, x in enumerate(x_windows):
        for j, y in enumerate(y_windows):
            if counts is not None:
                n = counts[(i, j)]
            else:
                n = int(((xs >= x) & (xs < x + window_size) & (ys >= y) & (ys < y + window_size)).sum())
            if n > 0:
                all_windows.add((i, j))
    try:
        existing = pd.read_csv(output_name)
    except FileNotFoundError:
        return all_windows, None
    done = set()
    if "window_id" in existing.columns:
        done = {(int(w) % len(x_windows), int(w) // len(x_windows)) for w in existing["window_id"].unique()}
    return all_windows - done, existing


def load_matching_results(outprefix):
    """(var_out, aligned_df, ref_df, matches_df) of one saved window directory (src/helpers.py:667-689)."""
    var_out = np.load(os.path.join(outprefix, "var_out.npy"), allow_pickle=True).item()
    aligned_df = pd.read_csv(os.path.join(outprefix, "aligned_df.csv"))
    ref_df = pd.read_csv(os.path.join(outprefix, "ref_df.csv"))
    matches_df = pd.read_csv(os.path.join(outprefix, "matches_df.csv"))
    return var_out, aligned_df, ref_df, matches_df
