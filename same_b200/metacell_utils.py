"""Metacell pre/post-processing with the reference's API (src/metacell_utils.py): `MetaCell`,
`greedy_triangle_collapse`, `unpack_metacell_matches`.

These sit OUTSIDE the per-window GPU hot path (Qhull runs once per collapse iteration on
the host either way).  They are host code, written from the reference's documented behaviour with numpy-vectorised
triangle scoring; the collapse loop keeps the reference's greedy order (perimeter ascending, non-overlapping batch).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import numpy as np
import pandas as pd
from scipy.spatial import Delaunay


@dataclass
class MetaCell:
    """Collapse result + metadata; same fields and helpers as the reference dataclass (src/metacell_utils.py:25-157)."""
    original_df: pd.DataFrame
    params: Dict[str, Any]
    x_col: str
    y_col: str
    cell_type_col: str
    original_idx_col: str
    metacell_idx_col: str
    original_delaunay: np.ndarray      # (n, 3) in original-ID space
    metacell_df: pd.DataFrame
    metacell_delaunay: np.ndarray      # (n, 3) rows of metacell_df

    def metacell_members(self, metacell_idx: int) -> List[Any]:
        return list(self.metacell_df.iloc[int(metacell_idx)]["members"])

    def original_delaunay_to_row_indices(self, triangles: Optional[np.ndarray] = None, *, on_missing: str = "drop") -> np.ndarray:
        tri = self.original_delaunay if triangles is None else np.asarray(triangles)
        if tri.size == 0:
            return np.array([], dtype=int).reshape(0, 3)
        if tri.ndim != 2 or tri.shape[1] != 3:
            raise ValueError(f"triangles must have shape (n, 3); got {tri.shape}")
        ids = self.original_df[self.original_idx_col].to_numpy()
        lut = pd.Series(np.arange(len(ids)), index=ids)
        lut = lut[~lut.index.duplicated(keep="last")]
        pos = lut.reindex(tri.reshape(-1)).to_numpy()
        missing = np.isnan(pos)
        if missing.any() and on_missing == "error":
            bad = list(dict.fromkeys(tri.reshape(-1)[missing].tolist()))[:10]
            raise KeyError(f"Found triangle vertices not in original_df[{self.original_idx_col}]: {bad}")
        pos = np.where(missing, -1, pos).astype(int).reshape(tri.shape)
        return pos[(pos >= 0).all(axis=1)]

    def original_delaunay_to_pos(self, triangles: Optional[np.ndarray] = None, *, on_missing: str = "drop") -> np.ndarray:
        return self.original_delaunay_to_row_indices(triangles=triangles, on_missing=on_missing)

    def original_delaunay_to_xy(self, triangles: Optional[np.ndarray] = None, *, on_missing: str = "drop") -> np.ndarray:
        pos = self.original_delaunay_to_row_indices(triangles=triangles, on_missing=on_missing)
        if pos.size == 0:
            return np.array([], dtype=float).reshape(0, 3, 2)
        return self.original_df[[self.x_col, self.y_col]].to_numpy(dtype=float)[pos]

    def metacell_delaunay_to_xy(self) -> np.ndarray:
        tri = np.asarray(self.metacell_delaunay)
        if tri.size == 0:
            return np.array([], dtype=float).reshape(0, 3, 2)
        return self.metacell_df[[self.x_col, self.y_col]].to_numpy(dtype=float)[tri.astype(int)]

    def to_summary_dict(self) -> Dict[str, Any]:
        return {"n_original": int(len(self.original_df)), "n_metacells": int(len(self.metacell_df)), "params": dict(self.params),
                "x_col": self.x_col, "y_col": self.y_col, "cell_type_col": self.cell_type_col,
                "original_idx_col": self.original_idx_col, "metacell_idx_col": self.metacell_idx_col,
                "n_original_triangles": int(getattr(self.original_delaunay, "shape", [0])[0]),
                "n_metacell_triangles": int(getattr(self.metacell_delaunay, "shape", [0])[0])}


def _reference_valid(p1, p2, p3, r_max, min_angle_deg):
    """is_triangle_valid of the reference, expression for expression (src/metacell_utils.py:233-262): 1-D np.linalg.norm / np.dot."""
    def angle(a, b, c):          # angle at b
        v1, v2 = a - b, c - b
        cos_angle = np.dot(v1, v2) / (np.linalg.norm(v1) * np.linalg.norm(v2))
        return np.degrees(np.arccos(np.clip(cos_angle, -1, 1)))
    if r_max is not None:
        if max(np.linalg.norm(p2 - p1), np.linalg.norm(p3 - p2), np.linalg.norm(p1 - p3)) > r_max:
            return False
    if min_angle_deg is not None:
        with np.errstate(invalid="ignore", divide="ignore"):
            if min(angle(p2, p1, p3), angle(p1, p2, p3), angle(p1, p3, p2)) < min_angle_deg:
                return False
    return True


def _valid_triangles(coords, tri, r_max, min_angle_deg):
    """Edge length <= r_max (note: '>' drops, src/metacell_utils.py:251) and min angle >= min_angle_deg (:257-261).

    Decided with array operations; the reference evaluates the same quantities through 1-D `np.linalg.norm` / `np.dot` (BLAS
    ddot, fused multiply-add), which can differ from the array expressions in the last bit, so triangles whose longest edge or
    smallest angle lies within a relative 1e-9 of a threshold are re-decided with the reference's own scalar expressions."""
    if len(tri) == 0:
        return np.zeros(0, bool)
    p = coords[tri]                                    # (T, 3, 2)
    ok = np.ones(len(tri), bool)
    band = np.zeros(len(tri), bool)
    e = [np.linalg.norm(p[:, (k + 1) % 3] - p[:, k], axis=1) for k in range(3)]
    if r_max is not None:
        longest = np.maximum(np.maximum(e[0], e[1]), e[2])
        ok &= ~(longest > r_max)
        band |= np.abs(longest - r_max) <= 1e-9 * abs(r_max)
    if min_angle_deg is not None:
        angs = []
        for k in range(3):
            v1, v2 = p[:, (k + 1) % 3] - p[:, k], p[:, (k + 2) % 3] - p[:, k]
            with np.errstate(invalid="ignore", divide="ignore"):
                c = np.einsum("ij,ij->i", v1, v2) / (np.linalg.norm(v1, axis=1) * np.linalg.norm(v2, axis=1))
            angs.append(np.degrees(np.arccos(np.clip(c, -1, 1))))
        mn = np.minimum(np.minimum(angs[0], angs[1]), angs[2])
        ok &= ~(mn < min_angle_deg)                    # NaN (degenerate) compares False -> kept, as in the reference
        with np.errstate(invalid="ignore"):
            band |= np.abs(mn - min_angle_deg) <= 1e-9 * max(abs(min_angle_deg), 1.0)
    for t in np.flatnonzero(band):
        ok[t] = _reference_valid(p[t, 0], p[t, 1], p[t, 2], r_max, min_angle_deg)
    return ok


def _filter_triangles(coords, tri, r_max, min_angle_deg, use_alpha_shape, alpha):
    ok = _valid_triangles(coords, tri, r_max, min_angle_deg)
    if use_alpha_shape:
        try:
            from alphashape import alphashape
            from shapely.geometry import Polygon
            shape = alphashape([tuple(c) for c in coords], alpha)
            ok &= np.array([shape.contains(Polygon(coords[t])) for t in tri], dtype=bool)
        except ImportError:
            print("Warning: alphashape not available, skipping alpha shape filtering")
    out = tri[ok]
    return out if len(out) else np.array([]).reshape(0, 3)


def greedy_triangle_collapse(aligned_df, max_metacell_size=3, max_iterations=1000, r_max=None, min_angle_deg=10, use_alpha_shape=False,
                             alpha=0.05, *, original_idx_col: str = "Cell_Num_Old", metacell_idx_col: str = "metacell_id",
                             x_col: str = "X", y_col: str = "Y", cell_type_col: str = "cell_type", return_object: bool = False):
    """Iteratively collapse same-type Delaunay triangles into metacells (src/metacell_utils.py:160-561).

    Returns `(metacell_df, metacell_delaunay)` or a `MetaCell` (`return_object=True`).  With `max_metacell_size=1`
    nothing collapses and the call only produces the filtered triangulation `run_same` reuses."""
    required = [x_col, y_col, cell_type_col, original_idx_col]
    missing = [c for c in required if c not in aligned_df.columns]
    if missing:
        raise ValueError(f"Input dataframe missing required columns: {missing}")
    aligned_df = aligned_df.copy()
    if aligned_df[original_idx_col].duplicated().any():
        dups = aligned_df.loc[aligned_df[original_idx_col].duplicated(), original_idx_col].head(5).tolist()
        raise ValueError(f"'{original_idx_col}' must be unique per original cell. Found duplicates (examples): {dups}")
    id_index = pd.Index(aligned_df[original_idx_col])
    # numeric columns whose member means run on the GPU: float64 / integer / bool without NaN (pandas sums exactly these in float64,
    # pairwise); anything else keeps pandas' own per-group mean
    gpu_cols = [c for c in aligned_df.columns
                if aligned_df[c].dtype.kind in "iub" or (aligned_df[c].dtype == np.float64 and not aligned_df[c].isna().any())]
    gpu_col_at = {c: k for k, c in enumerate(gpu_cols)}
    V = aligned_df[gpu_cols].to_numpy(dtype=np.float64) if gpu_cols else np.zeros((len(aligned_df), 0))

    coords0 = aligned_df[[x_col, y_col]].to_numpy()
    if len(coords0) >= 4:
        pos = _filter_triangles(coords0, Delaunay(coords0).simplices, r_max, min_angle_deg, use_alpha_shape, alpha)
    else:
        pos = np.array([], dtype=int).reshape(0, 3)
    ids = aligned_df[original_idx_col].to_numpy()
    original_delaunay = np.array([], dtype=ids.dtype).reshape(0, 3) if pos.size == 0 else ids[pos.astype(int)]

    id_columns = ["Cell_Num", "Cell_Num_Old", "cell_id", "Cell_ID", "ID", "id"]
    id_cols_present = [c for c in aligned_df.columns if c in id_columns]
    if original_idx_col not in id_cols_present:
        id_cols_present.append(original_idx_col)
    if metacell_idx_col in aligned_df.columns and metacell_idx_col not in id_cols_present:
        id_cols_present.append(metacell_idx_col)
    extra = [c for c in aligned_df.columns if c not in [x_col, y_col, cell_type_col] + id_cols_present]
    mdf = pd.DataFrame({x_col: aligned_df[x_col].to_numpy(), y_col: aligned_df[y_col].to_numpy(),
                        cell_type_col: aligned_df[cell_type_col].to_numpy(), "size": 1})
    mdf["members"] = [[v] for v in aligned_df[original_idx_col].tolist()]
    for c in extra:
        mdf[c] = aligned_df[c].to_numpy()
    mdf[metacell_idx_col] = range(len(mdf))

    for _ in range(max_iterations):
        coords = mdf[[x_col, y_col]].values
        if len(coords) < 4:
            break
        tri = _filter_triangles(coords, Delaunay(coords).simplices, r_max, min_angle_deg, use_alpha_shape, alpha)
        if len(tri) == 0:
            break
        tri = tri.astype(int)
        # candidate test, perimeter and batch selection on the GPU (src/metacell_utils.py:388-433, csrc/greedy.cu): same type,
        # merged size within the limit, candidates in ascending (perimeter, position) order — list.sort is stable — each taken
        # iff none of its vertices is used yet
        from .device import collapse_select
        sizes = mdf["size"].to_numpy()
        tot = sizes[tri[:, 0]] + sizes[tri[:, 1]] + sizes[tri[:, 2]]
        if not (tot <= max_metacell_size).any():            # nothing can merge (always the case for max_metacell_size=1): no device work
            break
        codes = pd.factorize(mdf[cell_type_col])[0]
        sel, perim = collapse_select(coords, codes, sizes, tri, max_metacell_size)
        chosen = np.flatnonzero(sel)
        if len(chosen) == 0:
            break
        batch = chosen[np.argsort(perim[chosen], kind="stable")].tolist()   # merged metacells are appended in selection order
        # merge step (src/metacell_utils.py:436-489): members of the three vertices concatenated, true means over the ORIGINAL
        # member cells for every numeric column — on the GPU in pandas' summation order (same_segment_mean) —, first vertex's
        # value for the others; merged metacells are appended in selection order
        tb = tri[batch]
        mem = mdf["members"].tolist()
        members = [mem[a] + mem[b] + mem[c] for a, b, c in tb.tolist()]
        ptr = np.r_[0, np.cumsum([len(m) for m in members])].astype(np.int64)
        pos = id_index.get_indexer([v for m in members for v in m])
        skip = [x_col, y_col, cell_type_col, "size", "members", metacell_idx_col] + id_cols_present
        other = [c for c in mdf.columns if c not in skip]
        mean_cols = [x_col, y_col] + [c for c in other if pd.api.types.is_numeric_dtype(mdf[c]) and c in aligned_df.columns]
        on_gpu = [c for c in mean_cols if c in gpu_col_at]
        means = {}
        if on_gpu:
            from .device import segment_mean
            got = segment_mean(V[:, [gpu_col_at[c] for c in on_gpu]], ptr, pos)
            means = {c: got[:, k] for k, c in enumerate(on_gpu)}
        for c in mean_cols:
            if c not in means:      # NaNs present or a dtype pandas sums differently (float32, object numbers): pandas itself, per group
                ser = aligned_df[c].reset_index(drop=True)
                means[c] = np.array([ser.iloc[pos[ptr[g]:ptr[g + 1]]].mean() for g in range(len(members))])
        new = {x_col: means[x_col], y_col: means[y_col], cell_type_col: mdf[cell_type_col].to_numpy()[tb[:, 0]], "size": tot[batch],
               "members": pd.Series(members, dtype=object)}
        for c in other:
            if c in means:
                new[c] = means[c]
            elif pd.api.types.is_numeric_dtype(mdf[c]):      # numeric column that the original frame does not have: size-weighted average
                new[c] = [np.average([mdf.iloc[i][c] for i in v], weights=[mdf.iloc[i]["size"] for i in v]) for v in tb.tolist()]
            else:
                new[c] = mdf[c].to_numpy()[tb[:, 0]]
        mdf = mdf.drop(tb.ravel().tolist()).reset_index(drop=True)
        mdf = pd.concat([mdf, pd.DataFrame(new)], ignore_index=True)
        mdf[metacell_idx_col] = range(len(mdf))

    final_coords = mdf[[x_col, y_col]].values
    if len(final_coords) >= 4:
        final = _filter_triangles(final_coords, Delaunay(final_coords).simplices, r_max, min_angle_deg, use_alpha_shape, alpha)
    else:
        final = np.array([]).reshape(0, 3)
    if return_object:
        params = {"max_metacell_size": max_metacell_size, "max_iterations": max_iterations, "r_max": r_max,
                  "min_angle_deg": min_angle_deg, "use_alpha_shape": use_alpha_shape, "alpha": alpha}
        return MetaCell(original_df=aligned_df, params=params, x_col=x_col, y_col=y_col, cell_type_col=cell_type_col,
                        original_idx_col=original_idx_col, metacell_idx_col=metacell_idx_col, original_delaunay=original_delaunay,
                        metacell_df=mdf, metacell_delaunay=final)
    return mdf, final


def unpack_metacell_matches(metacell_matches, metacell_aligned_df, metacell_ref_df, aligned_df=None, ref_df=None,
                            strategy="distribute", aligned_original_idx_col: Optional[str] = None,
                            ref_original_idx_col: Optional[str] = None, x_col: str = "X", y_col: str = "Y"):
    """Metacell-level matches -> individual cells (src/metacell_utils.py:564-766): columns `Aligned_cell_id`, `Ref_cell_id`.
    Reads `Aligned_metacell_id` / `Ref_metacell_id`; 'distribute' = round-robin over ref members, 'nearest' = Hungarian on
    member coordinates (ref members tiled when there are more aligned than ref members)."""
    from scipy.optimize import linear_sum_assignment
    from scipy.spatial.distance import cdist
    a_idx = r_idx = None
    if aligned_df is not None and aligned_original_idx_col is not None:
        if aligned_original_idx_col not in aligned_df.columns:
            raise ValueError(f"aligned_df missing aligned_original_idx_col='{aligned_original_idx_col}'")
        a_idx = aligned_df.set_index(aligned_original_idx_col, drop=False)
    if ref_df is not None and ref_original_idx_col is not None:
        if ref_original_idx_col not in ref_df.columns:
            raise ValueError(f"ref_df missing ref_original_idx_col='{ref_original_idx_col}'")
        r_idx = ref_df.set_index(ref_original_idx_col, drop=False)
    ref_has_mc = "members" in metacell_ref_df.columns and metacell_ref_df["members"].apply(lambda v: isinstance(v, list)).any()
    if ref_has_mc and strategy == "nearest" and (aligned_df is None or ref_df is None):
        raise ValueError("When ref has metacells and strategy='nearest', must provide both aligned_df and ref_df "
                         "for nearest neighbor unpacking.")
    if strategy == "nearest" and aligned_df is None:
        raise ValueError("strategy='nearest' requires aligned_df parameter")
    if strategy not in ("distribute", "nearest"):
        if ref_has_mc and len(metacell_matches):
            raise ValueError(f"Unknown strategy: {strategy}")
    out = []
    a_members = metacell_aligned_df["members"].tolist()
    r_members = metacell_ref_df["members"].tolist() if ref_has_mc else None
    for ma, mr in zip(metacell_matches["Aligned_metacell_id"].tolist(), metacell_matches["Ref_metacell_id"].tolist()):
        am = a_members[int(ma)]
        if not ref_has_mc:
            if strategy in ("distribute", "nearest"):
                out.extend({"Aligned_cell_id": m, "Ref_cell_id": mr} for m in am)
            continue
        rm = r_members[int(mr)]
        if strategy == "distribute":
            out.extend({"Aligned_cell_id": m, "Ref_cell_id": rm[i % len(rm)]} for i, m in enumerate(am))
        else:
            ac = (a_idx if a_idx is not None else aligned_df).loc[am, [x_col, y_col]].values
            rc = (r_idx if r_idx is not None else ref_df).loc[rm, [x_col, y_col]].values
            d = cdist(ac, rc)
            if len(am) > len(rm):
                d = np.tile(d, (1, int(np.ceil(len(am) / len(rm)))))
            rows, cols = linear_sum_assignment(d)
            out.extend({"Aligned_cell_id": am[i], "Ref_cell_id": rm[j % len(rm)]} for i, j in zip(rows, cols))
    return pd.DataFrame(out)
