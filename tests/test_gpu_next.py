"""GPU parity of the SURVEY §8(f) rows: ordered greedy selection (csrc/greedy.cu) behind the greedy MIP start
(src/init_helpers.py:110-132) and the batch selection of greedy_triangle_collapse (src/metacell_utils.py:423-433), against
the C oracle and the golden records of the unmodified reference (tests/golden/next/)."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
NEXT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "next")


def _golden(name):
    return np.load(os.path.join(NEXT, name), allow_pickle=True)


@pytest.mark.parametrize("degree", [1, 2, 3])
@pytest.mark.parametrize("n,n_nodes,levels", [(0, 5, 0), (1, 3, 0), (5000, 900, 0), (200000, 60000, 0), (50000, 20000, 7)])
def test_greedy_select_vs_oracle(degree, n, n_nodes, levels):
    """Random conflict hypergraphs; `levels` > 0 quantises the keys so that most comparisons are ties (index order decides)."""
    from same_b200.device import greedy_select
    rng = np.random.default_rng(n + degree)
    nodes = rng.integers(0, n_nodes, size=(n, degree)).astype(np.int32)
    key = rng.uniform(0, 100, n)
    if levels:
        key = np.floor(key / 100 * levels)
    key[rng.uniform(size=n) < 0.01] = -0.0          # -0.0 == +0.0 in the reference's sort
    key[rng.uniform(size=n) < 0.01] = 0.0
    eligible = rng.uniform(size=n) < 0.8
    for el in (None, eligible):
        got, rounds = greedy_select(nodes, key, n_nodes, el, return_rounds=True)
        want, _ = O.greedy_select(nodes, key, n_nodes, el)
        assert np.array_equal(got, want)
        assert rounds <= 64


def test_greedy_select_monotone_chain_uses_sequential_tail():
    """A path whose keys increase along it defeats the parallel rounds (one item per round): the single-thread tail finishes it."""
    from same_b200.device import greedy_select
    n = 400
    nodes = np.stack([np.arange(n), np.arange(1, n + 1)], axis=1).astype(np.int32)
    key = np.arange(n, dtype=np.float64)
    got, rounds = greedy_select(nodes, key, n + 1, return_rounds=True)
    want, _ = O.greedy_select(nodes, key, n + 1)
    assert rounds == 64 and np.array_equal(got, want) and got[::2].all() and not got[1::2].any()


def test_init_helpers_greedy_vs_reference_golden():
    """same_b200.init_helpers.compute_mip_start_pairs == the reference's, element for element (chosen order, unmatched set)."""
    from same_b200 import init_helpers
    g = _golden("mip_start.npz")
    for case in g["cases"]:
        pairs, cost, sizes, pen = g[f"{case}__pairs"], g[f"{case}__cost"], g[f"{case}__sizes"], float(g[f"{case}__penalty"])
        na, nr = int(pairs[:, 0].max()) + 1, int(pairs[:, 1].max()) + 1
        chosen, unmatched = init_helpers.compute_mip_start_pairs(valid_pairs=[tuple(p) for p in pairs.tolist()], costs=cost.tolist(), n_aligned=na,
                                                                 n_ref=nr, aligned_sizes=sizes, no_match_penalty=pen, max_matches=1,
                                                                 init_method="greedy", verbose=False)
        assert np.array_equal(np.asarray(chosen, dtype=np.int64).reshape(-1, 3), g[f"{case}__chosen"]), case
        assert sorted(unmatched) == g[f"{case}__unmatched"].tolist(), case
        v = init_helpers.mip_start_vectors(valid_pairs=pairs, costs=cost, n_aligned=na, n_ref=nr, aligned_sizes=sizes, no_match_penalty=pen,
                                           max_matches=1, init_method="greedy", verbose=False)
        assert np.array_equal(np.flatnonzero(v[0]), np.sort(g[f"{case}__chosen"][:, 2])) and np.array_equal(np.flatnonzero(v[1]), g[f"{case}__unmatched"])
    with pytest.raises(ValueError):
        init_helpers.compute_mip_start_pairs(valid_pairs=[(0, 0)], costs=[1.0], n_aligned=1, n_ref=1, aligned_sizes=np.ones(1), no_match_penalty=1.0,
                                             max_matches=2, init_method="hungarian")
    with pytest.raises(ValueError):
        init_helpers.compute_mip_start_pairs(valid_pairs=[(0, 0)], costs=[1.0], n_aligned=1, n_ref=1, aligned_sizes=np.ones(1), no_match_penalty=1.0,
                                             max_matches=1, init_method="nope")


def test_init_helpers_hungarian_small():
    from same_b200 import init_helpers
    pairs = [(0, 0), (0, 1), (1, 0), (1, 1), (2, 1)]
    cost = [1.0, 5.0, 2.0, 1.5, 50.0]
    chosen, unmatched = init_helpers.compute_mip_start_pairs(valid_pairs=pairs, costs=cost, n_aligned=3, n_ref=2, aligned_sizes=np.ones(3),
                                                             no_match_penalty=10.0, max_matches=1, init_method="hungarian", verbose=False)
    assert sorted(chosen) == [(0, 0, 0), (1, 1, 3)] and unmatched == {2}


@pytest.mark.parametrize("penalty", [100.0, 1.2])
def test_batch_mip_start_vs_oracle(penalty):
    """All windows of a batch at once (device-resident pairs / costs / sizes) == the oracle window by window."""
    from same_b200 import _lib as L
    from same_b200 import datagen
    from same_b200.device import Section
    ref, qry, ct = datagen.make_section_pair(n_tiles=9, n_types=3, seed=77)
    a_xy, r_xy = qry[["X", "Y"]].to_numpy(), ref[["X", "Y"]].to_numpy()
    rng = np.random.default_rng(1)
    a_size = rng.integers(1, 4, len(qry)).astype(np.float64)
    step = 10.0
    rects = np.array([[x, x + 14.0, y, y + 14.0] for x in np.arange(0, 30, step) for y in np.arange(0, 30, step)])
    with Section(a_xy, r_xy, qry[ct].to_numpy(), ref[ct].to_numpy(), a_size=a_size) as sec, sec.batch(rects) as b:
        b.candidates(1.0, 6, False, 1.0)
        rounds = b.mip_start(penalty)
        assert 0 < rounds <= 64
        sx, su = b.get(L.START_X), b.get(L.START_UNMATCHED)
        po, ko, ro = b.offsets(L.PAIRS), b.offsets(L.KEEP_A), b.offsets(L.KEEP_R)
        pairs, cost, keepA = b.get(L.PAIRS), b.get(L.COST), b.get(L.KEEP_A)
        n_sel = 0
        for w in range(len(rects)):
            p, c = pairs[po[w]:po[w + 1]], cost[po[w]:po[w + 1]]
            na, nr = int(ko[w + 1] - ko[w]), int(ro[w + 1] - ro[w])
            if len(p) == 0:
                continue
            chosen, unmatched = O.mip_start_greedy(p, c, na, nr, a_size[keepA[ko[w]:ko[w + 1]]], penalty)
            assert np.array_equal(np.flatnonzero(sx[po[w]:po[w + 1]]), np.sort(chosen[:, 2])), w
            assert np.array_equal(np.flatnonzero(su[ko[w]:ko[w + 1]]), unmatched), w
            n_sel += len(chosen)
        assert n_sel > 0 and (penalty > 50 or su.sum() > 0.2 * len(su))


def test_collapse_vs_reference_golden():
    """greedy_triangle_collapse with real merging == the unmodified reference: coordinates, averaged probability columns, sizes,
    member lists, metacell order and both triangulations, bit for bit."""
    import same_b200
    from same_b200 import datagen
    g = _golden("collapse.npz")
    for case in g["cases"]:
        tiles, seed, ms, r_max, ang = g[f"{case}__params"]
        ref, qry, ct = datagen.make_section_pair(n_tiles=int(tiles), seed=int(seed))
        mc = same_b200.greedy_triangle_collapse(qry, max_metacell_size=int(ms), r_max=float(r_max), min_angle_deg=None if ang < 0 else float(ang),
                                                return_object=True)
        mdf = mc.metacell_df
        assert np.array_equal(mdf[["X", "Y"]].to_numpy(), g[f"{case}__xy"]), case
        assert np.array_equal(mdf["size"].to_numpy(), g[f"{case}__size"]), case
        assert np.array_equal(mdf["cell_type"].astype(str).to_numpy(), g[f"{case}__type"]), case
        assert np.array_equal(mdf[ct].to_numpy(), g[f"{case}__prob"]), case
        assert np.array_equal(np.asarray([m for ms_ in mdf["members"] for m in ms_]), g[f"{case}__members_flat"]), case
        assert np.array_equal(np.r_[0, np.cumsum([len(m) for m in mdf["members"]])], g[f"{case}__members_ptr"]), case
        assert np.array_equal(mdf["metacell_id"].to_numpy(), g[f"{case}__metacell_id"]), case
        assert np.array_equal(np.asarray(mc.metacell_delaunay).reshape(-1, 3), g[f"{case}__delaunay"]), case
        assert np.array_equal(np.asarray(mc.original_delaunay).reshape(-1, 3), g[f"{case}__original_delaunay"]), case


def test_metacell_collapse_and_unpack_roundtrip():
    import same_b200
    from same_b200 import datagen
    ref, qry, ct = datagen.make_section_pair(n_tiles=1, seed=3)
    mdf, tri = same_b200.greedy_triangle_collapse(qry, max_metacell_size=3, r_max=1.5, min_angle_deg=10)
    assert mdf["size"].sum() == len(qry) and mdf["size"].max() <= 3 and len(mdf) < len(qry)
    members = sorted(m for ms in mdf["members"] for m in ms)
    assert members == sorted(qry["Cell_Num_Old"].tolist())
    big = mdf[mdf["size"] == 3].iloc[0]
    sub = qry.set_index("Cell_Num_Old").loc[big["members"]]
    assert np.isclose(big["X"], sub["X"].mean()) and np.isclose(big["c1"], sub["c1"].mean()) and len(set(sub["cell_type"])) == 1
    matches = pd.DataFrame({"Aligned_metacell_id": [0, int(big["metacell_id"])], "Ref_metacell_id": [5, 7]})
    ind = same_b200.unpack_metacell_matches(matches, mdf, ref)
    assert len(ind) == len(mdf.iloc[0]["members"]) + 3 and set(ind["Ref_cell_id"]) == {5, 7}




def test_unpack_metacell_matches_vs_reference_golden():
    """unpack_metacell_matches == the unmodified reference (src/metacell_utils.py:564-766) row for row: 'distribute' and 'nearest',
    metacells on the aligned side only and on both sides.  (The metacells come from greedy_triangle_collapse, hence a GPU test.)"""
    import same_b200
    from same_b200 import datagen
    g = _golden("unpack.npz")
    ref, qry, ct = datagen.make_section_pair(n_tiles=2, seed=31)
    mc_a = same_b200.greedy_triangle_collapse(qry, max_metacell_size=4, r_max=1.5, min_angle_deg=10, return_object=True)
    mc_r = same_b200.greedy_triangle_collapse(ref, max_metacell_size=3, r_max=1.5, min_angle_deg=10, return_object=True)
    for case in g["cases"]:
        m = pd.DataFrame({"Aligned_metacell_id": g["match_a"], "Ref_metacell_id": g[f"{case}__match_r"]})
        ref_side = mc_r.metacell_df if str(case).startswith("both") else ref
        out = same_b200.unpack_metacell_matches(m, mc_a.metacell_df, ref_side, aligned_df=qry, ref_df=ref, strategy=str(case).split("_")[-1],
                                                aligned_original_idx_col="Cell_Num_Old", ref_original_idx_col="Cell_Num_Old")
        assert np.array_equal(out["Aligned_cell_id"].to_numpy(np.int64), g[f"{case}__aligned"]), case
        assert np.array_equal(out["Ref_cell_id"].to_numpy(np.int64), g[f"{case}__ref"]), case


@pytest.mark.parametrize("lattice", [False, True])
def test_collapse_select_vs_oracle(lattice):
    """same_collapse_select == oracle: candidate flags, perimeters bit for bit (fma(dy,dy,dx*dx) per side), selection.  The lattice
    variant puts the points on an integer grid with a few fractional centroids, so most perimeters tie exactly or to the last bit."""
    from scipy.spatial import Delaunay
    from same_b200.device import collapse_select
    rng = np.random.default_rng(3 + lattice)
    n = 6000
    if lattice:
        gx, gy = np.meshgrid(np.arange(80.0) * 17.0, np.arange(75.0) * 17.0)
        xy = np.stack([gx.ravel(), gy.ravel()], axis=1)[:n]
        k = rng.choice(n, 600, replace=False)
        xy[k] += rng.integers(-5, 6, size=(600, 2)) / 3.0
    else:
        xy = rng.uniform(0, 1000, size=(n, 2))
    tri = Delaunay(xy).simplices.astype(np.int32)
    types = rng.integers(0, 3, n).astype(np.int32)
    types[: n // 2] = 0                                    # a large same-type region so that many triangles compete
    sizes = rng.integers(1, 5, n).astype(np.float64)
    sel, per = collapse_select(xy, types, sizes, tri, 8)
    want_sel, want_per = O.collapse_select(xy, types, sizes, tri, 8)
    assert np.array_equal(per, want_per)
    assert np.array_equal(sel, want_sel) and sel.sum() > 100
    if lattice:
        assert len(np.unique(per)) < 0.2 * len(per)


def test_segment_mean_vs_oracle():
    """same_segment_mean == oracle (== pandas mean, tests/test_oracle_golden.py) bit for bit, incl. groups beyond 8 and 128 members."""
    from same_b200.device import segment_mean
    rng = np.random.default_rng(2)
    V = rng.uniform(-1, 1, (20000, 5)) * 10.0 ** rng.integers(-3, 4, (20000, 5))
    sizes = np.r_[rng.integers(1, 12, 5000), [127, 128, 129, 300, 1000, 5000]]
    ptr = np.r_[0, np.cumsum(sizes)]
    pos = rng.integers(0, len(V), ptr[-1]).astype(np.int32)
    assert np.array_equal(segment_mean(V, ptr, pos), O.segment_mean(V, ptr, pos))
    assert segment_mean(V, np.array([0]), np.zeros(0, np.int32)).shape == (0, 5)
