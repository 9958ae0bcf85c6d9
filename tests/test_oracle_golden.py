"""Pin the CPU oracle (oracle/same_oracle.c) to the reference's own outputs.

The golden .npz files were produced by running the unmodified reference
(tests/golden/gen_golden.py).  Integer/index outputs must be bit-exact, costs are
compared bit-exactly too (the oracle reproduces the reference's operation order;
the contract in BASELINE.json is 1e-5 relative)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pipeline as OP
from tests.util import (GOLDEN_CASES, STAGE_CASES, golden_params, incumbent_rule, joint_type_codes, load_golden)

SINGLE = [c for c in GOLDEN_CASES if c not in ("tiles4_sliding", "sparse_merge")]


@pytest.mark.parametrize("case", STAGE_CASES)
@pytest.mark.parametrize("brute", [False, True])
def test_a1_candidates(case, brute):
    g = load_golden(case)
    o = golden_params(g, "optim")
    keepA, keepR, pairs = O.find_knn_within_radius(g["aligned_xy"], g["ref_xy"], float(o["radius"]), int(o["knn"]), brute=brute)
    assert np.array_equal(keepA, g["knn_keepA"])
    assert np.array_equal(keepR, g["knn_keepR"])
    assert np.array_equal(pairs, g["knn_pairs"])


@pytest.mark.parametrize("case", STAGE_CASES)
def test_a2_priority(case):
    g = load_golden(case)
    tA, tR = joint_type_codes(g)
    out = O.knn_priority(g["knn_pairs"], tA[g["knn_keepA"]], tR[g["knn_keepR"]])
    assert np.array_equal(out, g["prio_pairs"])


@pytest.mark.parametrize("case", STAGE_CASES)
def test_a5_remap(case):
    g = load_golden(case)
    out, src = O.remap_triangles(g["remap_tri_global"], g["remap_vid_all"][g["remap_rows"]])
    assert np.array_equal(out, g["remap_out"])
    assert np.all(np.diff(src) > 0)


@pytest.mark.parametrize("case", STAGE_CASES)
@pytest.mark.parametrize("variant", ["same0", "same1", "tight"])
def test_a6_filter(case, variant):
    g = load_golden(case)
    o = golden_params(g, "optim")
    tA, _ = joint_type_codes(g)
    pts = g["aligned_xy"][g["knn_keepA"]]
    ty = tA[g["knn_keepA"]]
    radius = float(g["filt_tight_radius"]) if variant == "tight" else float(o["radius"])
    same = variant != "same0"
    kept, unc, band = O.filter_triangles(pts, g["delaunay"], radius, float(o["min_angle_deg"]), ty, same)
    if variant != "tight":  # the tight radius is a quantile of the side lengths and may equal one of them
        assert band == 0, "golden case sits inside the threshold guard band"
    assert np.array_equal(g["delaunay"][kept], g[f"filt_{variant}"])
    assert np.array_equal(unc, g[f"filt_{variant}_unc"])


def _pipeline(g, w=0):
    o = golden_params(g, "optim")
    tA, tR = joint_type_codes(g)
    tri_global = g["mc_aligned_delaunay"] if "mc_aligned_delaunay" in g else None
    vid = g["mc_aligned_metacell_id"] if tri_global is not None else None
    a_size = g["aligned_size"] if "aligned_size" in g else np.ones(len(g["aligned_xy"]))
    r_size = g["ref_size"] if "ref_size" in g else np.ones(len(g["ref_xy"]))
    mult = o.get("ref_metacell_match_multiplier")
    res = OP.window_pipeline(
        g["aligned_xy"], g["ref_xy"], g["aligned_prob"], g["ref_prob"], tA, tR, a_size, r_size,
        radius=float(o["radius"]), knn=int(o["knn"]), dist_ct_coeff=float(o.get("dist_ct_coeff", 1)),
        min_angle_deg=float(o.get("min_angle_deg", 15)),
        ignore_same_type_triangles=bool(o.get("ignore_same_type_triangles", True)),
        ignore_knn_if_matched=bool(o.get("ignore_knn_if_matched", False)), max_matches=int(o.get("max_matches", 1)),
        ref_metacell_match_multiplier=None if mult is None else int(mult), tri_global=tri_global, a_vid=vid)
    return o, res


@pytest.mark.parametrize("case", SINGLE)
def test_model_arrays(case):
    """a1-a9 composed as run_same composes them: pairs, costs, triangles, signs, weights, constraints."""
    g = load_golden(case)
    o, res = _pipeline(g)
    assert np.array_equal(res["pairs"], g["w0_pairs"])
    assert np.array_equal(res["cost"], g["w0_cost"]), np.abs(res["cost"] - g["w0_cost"]).max()
    assert np.array_equal(res["tri"], g["w0_tri"])
    assert np.array_equal(res["sign"].astype(np.float64), g["w0_source_signs"])
    dp = float(o.get("delaunay_penalty", 5))
    assert np.array_equal(dp * res["weight"], g["w0_obj_q"])
    names, sense, rhs, ptr, idx, val = OP.constraints_from_groups(res, len(res["pairs"]))
    assert np.array_equal(names, g["w0_con_name"])
    assert np.array_equal(sense, g["w0_con_sense"])
    assert np.array_equal(rhs, g["w0_con_rhs"])
    assert np.array_equal(ptr, g["w0_con_ptr"])
    assert np.array_equal(idx, g["w0_con_idx"])
    assert np.array_equal(val, g["w0_con_val"])


@pytest.mark.parametrize("case", SINGLE)
def test_a10_lazy_cuts(case):
    g = load_golden(case)
    o, res = _pipeline(g)
    gp = golden_params(g, "gurobi")
    x = incumbent_rule(res["pairs"], int(g["seed"]))
    assert np.array_equal(x, g["w0_x_sol"])
    cuts = O.lazy_cuts(x, res["pairs"], res["tri"], res["sign"], g["ref_xy"][res["keepR"]], len(res["keepA"]),
                       gp.get("lazy_allowed_flip_fraction", 0.05), gp.get("lazy_max_cuts_per_incumbent", 1000),
                       gp.get("lazy_max_cuts"))
    assert np.array_equal(cuts, g["w0_cuts"])
    if "vo_lazy_cuts_added" in g:
        assert len(cuts) == int(g["vo_lazy_cuts_added"])


@pytest.mark.parametrize("case", SINGLE)
def test_a11_a12_postsolve(case):
    g = load_golden(case)
    if "vo_areas_before" not in g:
        pytest.skip("sliding-window golden: var_out is not returned by the driver")
    o, res = _pipeline(g)
    x = g["w0_x_sol"]
    mj, _ = O.matching_from_x(x, res["pairs"], len(res["keepA"]))
    ps = O.postsolve(res["tri"], g["aligned_xy"][res["keepA"]], g["ref_xy"][res["keepR"]], mj)
    assert np.array_equal(ps["area_before"], g["vo_areas_before"])
    assert np.array_equal(np.isnan(ps["area_after"]), np.isnan(g["vo_areas_after"]))
    ok = ~np.isnan(ps["area_after"])
    assert np.array_equal(ps["area_after"][ok], g["vo_areas_after"][ok])
    assert np.array_equal(np.flatnonzero(ps["flipped"]), g["vo_flipped"])
    m = ps["mask"]
    matched = np.stack([(m >> 8) & 1, (m >> 9) & 1, (m >> 10) & 1], axis=1).astype(bool)
    assert np.array_equal(matched, g["vo_matched_vertices"])
    # violation summary (violationhelper.py:54-121)
    tri = res["tri"]
    nm = matched.sum(axis=1)
    comparisons = np.where(nm == 3, 3, np.where(nm == 2, 1, 0)).sum()
    xv = np.stack([(m >> q) & 1 for q in range(3)], axis=1)
    yv = np.stack([(m >> (3 + q)) & 1 for q in range(3)], axis=1)
    viol_tri = np.flatnonzero((m & 63) != 0)
    assert np.array_equal(viol_tri, g["vo_tri_with_viol"])
    assert np.array_equal([len(tri), len(viol_tri), comparisons, xv.sum() + yv.sum()], g["vo_summary"])
    P = np.array([[0, 1], [0, 2], [1, 2]])
    pts = set()
    for t in viol_tri:
        for q in range(3):
            if xv[t, q] or yv[t, q]:
                pts.update(tri[t, P[q]].tolist())
    assert np.array_equal(sorted(pts), g["vo_pts_with_viol"])
    # x-violation records as a set of (t, v1, v2)
    gx = {tuple(r) for r in g["vo_xviol"].tolist()}
    ox = {(int(t), int(tri[t, P[q, 0]]), int(tri[t, P[q, 1]])) for t in viol_tri for q in range(3) if xv[t, q]}
    assert gx == ox
    # triangle info (helpers.py:184-210)
    assert np.array_equal(res["argv"], g["vo_tri_info"])
    assert np.array_equal(res["bounds"], g["vo_tri_bounds"])


def test_simulated_st_known_answer():
    """the saved run of examples/simulated_st has 1152 = 144*8 candidate pairs."""
    g = load_golden("simulated_st")
    assert len(g["knn_pairs"]) == 1152
    keepA, keepR, pairs = O.find_knn_within_radius(g["aligned_xy"], g["ref_xy"], 3.0, 8)
    assert len(pairs) == 1152 and len(keepA) == 144 and len(keepR) == 144


def test_grid_equals_brute_random():
    rng = np.random.default_rng(0)
    for n, r, k in [(500, 0.07, 8), (800, 0.2, 3), (300, 0.01, 5), (50, 2.0, 64)]:
        a, b = rng.uniform(size=(n, 2)), rng.uniform(size=(n + 17, 2))
        c1, n1 = O.knn_candidates(a, b, r, k)
        c2, n2 = O.knn_candidates(a, b, r, k, brute=True)
        assert np.array_equal(c1, c2) and np.array_equal(n1, n2)


def test_empty_inputs():
    c, n = O.knn_candidates(np.zeros((3, 2)), np.zeros((0, 2)), 1.0, 4)
    assert n.sum() == 0 and (c == -1).all()
    keepA, keepR, pairs = O.knn_compact(c, n, 0)
    assert len(keepA) == 0 and len(keepR) == 0 and len(pairs) == 0
    v, ck = O.separation(np.zeros((0, 3), np.int32), np.zeros(0, np.int8), np.zeros(1, np.int32), np.zeros((1, 2)))
    assert len(v) == 0 and ck == 0


# ---- greedy MIP start, metacell collapse -------------------------------------------------
def _next_golden(name):
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "next", name), allow_pickle=True)


def test_oracle_mip_start_equals_reference():
    """oracle_greedy_select reproduces compute_mip_start_pairs(init_method='greedy') (src/init_helpers.py:110-132): same
    chosen pairs in the same order, same unmatched set — including the tie-heavy rounded-cost case."""
    g = _next_golden("mip_start.npz")
    for case in g["cases"]:
        pairs, cost, sizes, pen = g[f"{case}__pairs"], g[f"{case}__cost"], g[f"{case}__sizes"], float(g[f"{case}__penalty"])
        na, nr = int(pairs[:, 0].max()) + 1, int(pairs[:, 1].max()) + 1
        chosen, unmatched = O.mip_start_greedy(pairs, cost, na, nr, sizes, pen)
        assert np.array_equal(chosen, g[f"{case}__chosen"]), case
        assert np.array_equal(unmatched, g[f"{case}__unmatched"]), case


def test_oracle_collapse_equals_reference(monkeypatch):
    """oracle_collapse_score + oracle_greedy_select (candidate test, perimeter in the reference's fma arithmetic, ordered disjoint
    selection) reproduce greedy_triangle_collapse of the unmodified reference: plugged into the host loop of
    same_b200.metacell_utils in place of the CUDA call, the metacell frames equal the recorded ones bit for bit — on the seeded
    sections and on the lattice-like ISS heart sections, where perimeters tie to the last bit."""
    import same_b200
    from same_b200 import datagen, device
    from tests.util import golden_frame
    monkeypatch.setattr(device, "collapse_select", lambda xy, tc, sz, tri, ms, device=0: O.collapse_select(xy, tc, sz, tri, ms))
    monkeypatch.setattr(device, "segment_mean", lambda v, ptr, pos, device=0: O.segment_mean(v, ptr, pos))
    g = _next_golden("collapse.npz")
    for case in g["cases"]:
        tiles, seed, ms, r_max, ang = g[f"{case}__params"]
        ref, qry, ct = datagen.make_section_pair(n_tiles=int(tiles), seed=int(seed))
        mc = same_b200.greedy_triangle_collapse(qry, max_metacell_size=int(ms), r_max=float(r_max), min_angle_deg=None if ang < 0 else float(ang),
                                                return_object=True)
        assert np.array_equal(mc.metacell_df[["X", "Y"]].to_numpy(), g[f"{case}__xy"]), case
        assert np.array_equal(mc.metacell_df[ct].to_numpy(), g[f"{case}__prob"]), case
        assert np.array_equal(np.asarray([m for ms_ in mc.metacell_df["members"] for m in ms_]), g[f"{case}__members_flat"]), case
        assert np.array_equal(np.asarray(mc.metacell_delaunay).reshape(-1, 3), g[f"{case}__delaunay"]), case
    h = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "next", "heart_mc10.npz")))
    for tag, df in (("mca", golden_frame(h, "aligned")), ("mcr", golden_frame(h, "ref"))):
        mc = same_b200.greedy_triangle_collapse(df, cell_type_col="cell_type", original_idx_col="Cell_Num", return_object=True, max_metacell_size=10,
                                                r_max=50.0, min_angle_deg=15.0, use_alpha_shape=False)
        assert np.array_equal(mc.metacell_df[["X", "Y"]].to_numpy(), h[f"{tag}_xy"]), tag
        assert np.array_equal(mc.metacell_df["size"].to_numpy(), h[f"{tag}_size"]), tag
        assert np.array_equal(mc.metacell_df[[str(c) for c in h["commonCT"]]].to_numpy(), h[f"{tag}_prob"]), tag
        assert np.array_equal(np.asarray(mc.metacell_delaunay, dtype=np.int64).reshape(-1, 3), h[f"{tag}_delaunay"]), tag


def test_oracle_segment_mean_equals_pandas_mean():
    """oracle_segment_mean == `rows[col].mean()` as the reference computes member means (src/metacell_utils.py:446-474): pandas
    nanmean = numpy pairwise sum / count, for group sizes on both sides of numpy's 8- and 128-element thresholds."""
    import pandas as pd
    rng = np.random.default_rng(1)
    V = rng.uniform(-1, 1, (4000, 3)) * 10.0 ** rng.integers(-3, 4, (4000, 3))
    sizes = np.r_[rng.integers(1, 20, 300), [7, 8, 9, 15, 16, 17, 127, 128, 129, 136, 200, 257, 1000]]
    ptr = np.r_[0, np.cumsum(sizes)]
    pos = rng.integers(0, len(V), ptr[-1]).astype(np.int32)
    out = O.segment_mean(V, ptr, pos)
    df = pd.DataFrame(V, columns=list("abc"))
    for g in range(len(sizes)):
        rows = df.iloc[pos[ptr[g]:ptr[g + 1]]]
        assert [rows[c].mean() for c in "abc"] == out[g].tolist(), (g, sizes[g])
    ints = rng.integers(0, 1000, 4000)                       # integer columns are summed in float64 too
    assert O.segment_mean(ints.astype(np.float64), ptr[:50], pos[:ptr[49]])[:, 0].tolist() == \
        [pd.Series(ints).iloc[pos[ptr[g]:ptr[g + 1]]].mean() for g in range(49)]
