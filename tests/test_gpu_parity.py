"""GPU parity: the CUDA library (through its C-ABI) against the golden vectors of the reference
and against the CPU oracle on seeded inputs.  Index/structure outputs are compared bit-exactly;
costs are compared bit-exactly as well (the contract is 1e-5 relative, BASELINE.json)."""
import numpy as np
import pytest
from scipy.spatial import Delaunay

from oracle import oracle as O
from oracle import pipeline as OP
from tests.util import (GOLDEN_CASES, STAGE_CASES, golden_params, incumbent_rule, joint_type_codes, load_golden)

pytestmark = pytest.mark.gpu

SINGLE = [c for c in GOLDEN_CASES if c not in ("tiles4_sliding", "sparse_merge")]


def _section(g, with_sizes=True):
    from same_b200.device import Section
    tA, tR = joint_type_codes(g)
    a_size = g["aligned_size"] if "aligned_size" in g else None
    r_size = g["ref_size"] if "ref_size" in g else None
    return Section(g["aligned_xy"], g["ref_xy"], g["aligned_prob"], g["ref_prob"], tA, tR, a_size, r_size)


def _run_window_pipeline(g):
    """The product pipeline for one unbounded window, as run_same drives it."""
    from same_b200 import _lib as L
    o = golden_params(g, "optim")
    sec = _section(g)
    b = sec.batch()
    b.candidates(float(o["radius"]), int(o["knn"]), bool(o.get("ignore_knn_if_matched", False)), float(o.get("dist_ct_coeff", 1)))
    pre = "mc_aligned_delaunay" in g
    if pre:
        sec.set_triangles(g["mc_aligned_delaunay"], g["mc_aligned_metacell_id"])
        b.triangles_remap()
    else:
        keepA = b.get(L.KEEP_A)
        tri = Delaunay(g["aligned_xy"][keepA]).simplices.astype(np.int32)
        b.triangles_set(tri, [0, len(tri)])
    same = bool(o.get("ignore_same_type_triangles", True))
    nb = b.tri_classify(float(o["radius"]), float(o.get("min_angle_deg", 15)), same)
    assert nb == 0
    b.tri_finalize(same, True, remove_unconstrained=pre)
    mult = o.get("ref_metacell_match_multiplier")
    b.groups(int(o.get("max_matches", 1)), None if mult is None else int(mult))
    return o, sec, b


@pytest.mark.parametrize("case", STAGE_CASES)
def test_a1_candidates_vs_golden(case):
    from same_b200 import _lib as L
    g = load_golden(case)
    o = golden_params(g, "optim")
    with _section(g) as sec, sec.batch() as b:
        b.candidates(float(o["radius"]), int(o["knn"]))
        assert np.array_equal(b.get(L.KEEP_A), g["knn_keepA"])
        assert np.array_equal(b.get(L.KEEP_R), g["knn_keepR"])
        assert np.array_equal(b.get(L.PAIRS), g["knn_pairs"])


@pytest.mark.parametrize("case", STAGE_CASES)
def test_a2_priority_vs_golden(case):
    from same_b200 import _lib as L
    g = load_golden(case)
    o = golden_params(g, "optim")
    with _section(g) as sec, sec.batch() as b:
        b.candidates(float(o["radius"]), int(o["knn"]), priority=True)
        assert np.array_equal(b.get(L.PAIRS), g["prio_pairs"])
        assert np.array_equal(b.get(L.KEEP_A), g["knn_keepA"])


@pytest.mark.parametrize("case", STAGE_CASES)
def test_a5_remap_vs_golden(case):
    """Window = a random subset of rows: emulate with a section whose aligned frame is that subset."""
    from same_b200 import _lib as L
    from same_b200.device import Section
    g = load_golden(case)
    rows = g["remap_rows"]
    pts = g["aligned_xy"][g["knn_keepA"]][rows]
    n = len(pts)
    # every aligned row gets itself as a candidate so that nothing is dropped by KNN compaction
    with Section(pts, pts, np.zeros((n, 1)), np.zeros((n, 1))) as sec, sec.batch() as b:
        b.candidates(1e-12, 1)
        assert np.array_equal(b.get(L.KEEP_A), np.arange(n))
        sec.set_triangles(g["remap_tri_global"], g["remap_vid_all"][rows])
        b.triangles_remap()
        assert np.array_equal(b.get(L.TRI_IN), g["remap_out"])


@pytest.mark.parametrize("case", STAGE_CASES)
@pytest.mark.parametrize("variant", ["same0", "same1", "tight"])
def test_a6_filter_vs_golden(case, variant):
    from same_b200 import _lib as L
    from same_b200.device import Section
    g = load_golden(case)
    o = golden_params(g, "optim")
    tA, _ = joint_type_codes(g)
    pts = g["aligned_xy"][g["knn_keepA"]]
    ty = tA[g["knn_keepA"]]
    n = len(pts)
    radius = float(g["filt_tight_radius"]) if variant == "tight" else float(o["radius"])
    same = variant != "same0"
    with Section(pts, pts, np.zeros((n, 1)), np.zeros((n, 1)), ty, ty) as sec, sec.batch() as b:
        b.candidates(1e-12, 1)
        b.triangles_set(g["delaunay"], [0, len(g["delaunay"])])
        b.tri_classify(radius, float(o["min_angle_deg"]), same)
        b.tri_finalize(same, True, remove_unconstrained=False)
        assert np.array_equal(b.get(L.TRI), g[f"filt_{variant}"])
        assert np.array_equal(b.get(L.UNCONSTRAINED), g[f"filt_{variant}_unc"])
        assert np.array_equal(g["delaunay"][b.get(L.TRI_SRC)], g[f"filt_{variant}"])


@pytest.mark.parametrize("case", SINGLE)
def test_model_arrays_vs_golden(case):
    g = load_golden(case)
    o, sec, b = _run_window_pipeline(g)
    m = b.window_model(0)
    assert np.array_equal(m["pairs"], g["w0_pairs"])
    assert np.array_equal(m["cost"], g["w0_cost"]), np.abs(m["cost"] - g["w0_cost"]).max()
    assert np.array_equal(m["tri"], g["w0_tri"])
    assert np.array_equal(m["sign"].astype(np.float64), g["w0_source_signs"])
    assert np.array_equal(float(o.get("delaunay_penalty", 5)) * m["weight"], g["w0_obj_q"])
    # constraints in the reference's creation order (helpers.py:130-158)
    res = dict(keepA=m["keepA"], keepR=m["keepR"], ref_group_node=m["ref_group_node"], ref_group_ptr=m["ref_group_ptr"],
               ref_group_idx=m["ref_group_idx"], ref_group_limit=m["ref_group_limit"])
    na = len(m["keepA"])
    res["al_group_node"] = np.flatnonzero(np.diff(m["row_ptr"]) > 0).astype(np.int32)
    res["al_group_ptr"] = np.r_[m["row_ptr"][:-1][np.diff(m["row_ptr"]) > 0], m["row_ptr"][-1]]
    res["al_group_idx"] = np.arange(len(m["pairs"]), dtype=np.int32)
    names, sense, rhs, ptr, idx, val = OP.constraints_from_groups(res, len(m["pairs"]))
    for got, key in ((names, "con_name"), (sense, "con_sense"), (rhs, "con_rhs"), (ptr, "con_ptr"), (idx, "con_idx"), (val, "con_val")):
        assert np.array_equal(got, g[f"w0_{key}"]), key
    b.close(); sec.close()


@pytest.mark.parametrize("case", SINGLE)
def test_a10_separation_vs_golden(case):
    g = load_golden(case)
    o, sec, b = _run_window_pipeline(g)
    gp = golden_params(g, "gurobi")
    x = g["w0_x_sol"]
    cap = gp.get("lazy_max_cuts_per_incumbent", 1000) or 1000
    nv, nc, cuts = b.separation(x, cap=int(cap))
    allowed = gp.get("lazy_allowed_flip_fraction", 0.05)
    want = g["w0_cuts"]
    if nc[0] == 0 or nv[0] == 0 or (allowed is not None and nv[0] / float(nc[0]) <= allowed):
        assert len(want) == 0
    else:
        k = min(int(nv[0]), int(cap))
        assert np.array_equal(cuts[0, :k], want)
    # and against the oracle's raw counts
    m = b.window_model(0)
    mj, _ = O.matching_from_x(x, m["pairs"], len(m["keepA"]))
    viol, checked = O.separation(m["tri"], m["sign"], mj, g["ref_xy"][m["keepR"]])
    assert nv[0] == len(viol) and nc[0] == checked
    b.close(); sec.close()


@pytest.mark.parametrize("case", SINGLE)
def test_a11_a12_postsolve_vs_golden(case):
    from same_b200 import _lib as L
    g = load_golden(case)
    if "vo_areas_before" not in g:
        pytest.skip("sliding-window golden has no var_out")
    o, sec, b = _run_window_pipeline(g)
    b.postsolve(g["w0_x_sol"])
    ab, aa, fl, mask = b.get(L.AREA_BEFORE), b.get(L.AREA_AFTER), b.get(L.FLIPPED), b.get(L.TRI_MASK)
    assert np.array_equal(ab, g["vo_areas_before"])
    assert np.array_equal(np.isnan(aa), np.isnan(g["vo_areas_after"]))
    ok = ~np.isnan(aa)
    assert np.array_equal(aa[ok], g["vo_areas_after"][ok])
    assert np.array_equal(np.flatnonzero(fl), g["vo_flipped"])
    assert np.array_equal(np.flatnonzero((mask & 63) != 0), g["vo_tri_with_viol"])
    matched = np.stack([(mask >> 8) & 1, (mask >> 9) & 1, (mask >> 10) & 1], axis=1).astype(bool)
    assert np.array_equal(matched, g["vo_matched_vertices"])
    assert np.array_equal(b.get(L.TRI_ARGV), g["vo_tri_info"])
    assert np.array_equal(b.get(L.TRI_BOUNDS), g["vo_tri_bounds"])
    b.close(); sec.close()


# ---------------------------------------------------------------------------------------------------
# seeded sections vs the oracle, batched windows
# ---------------------------------------------------------------------------------------------------
def _oracle_window(a_xy, r_xy, a_prob, r_prob, tA, tR, sA, sR, rect, **kw):
    ra = O.subset(a_xy, *rect)
    rr = O.subset(r_xy, *rect)
    res = OP.window_pipeline(a_xy[ra], r_xy[rr], a_prob[ra], r_prob[rr], tA[ra], tR[rr], sA[ra], sR[rr], **kw)
    res["rowsA"], res["rowsR"] = ra, rr
    return res


@pytest.mark.parametrize("n_tiles,knn,k_types,seed", [(9, 8, 3, 0), (16, 5, 8, 1), (6, 20, 3, 2), (4, 40, 3, 3)])
def test_batched_windows_vs_oracle(n_tiles, knn, k_types, seed):
    from same_b200 import _lib as L
    from same_b200 import datagen
    from same_b200.device import Section
    ref, qry, ct = datagen.make_section_pair(n_tiles=n_tiles, n_types=k_types, seed=seed)
    a_xy, r_xy = qry[["X", "Y"]].to_numpy(), ref[["X", "Y"]].to_numpy()
    a_prob, r_prob = qry[ct].to_numpy(), ref[ct].to_numpy()
    lut = {c: i for i, c in enumerate(ct)}
    tA = qry["cell_type"].map(lut).to_numpy(np.int32)
    tR = ref["cell_type"].map(lut).to_numpy(np.int32)
    rng = np.random.default_rng(seed)
    sA = np.where(rng.uniform(size=len(qry)) < 0.1, 3.0, 1.0)
    sR = np.where(rng.uniform(size=len(ref)) < 0.1, 2.0, 1.0)
    ext = max(a_xy.max(), r_xy.max()) + 1
    step, ws = ext / 3 + 0.01, ext / 3 + 2.0
    rects = np.array([[i * step, i * step + ws, j * step, j * step + ws] for i in range(3) for j in range(3)])
    radius = 1.1
    kw = dict(radius=radius, knn=knn, dist_ct_coeff=1.7, min_angle_deg=15, ignore_same_type_triangles=True, max_matches=2)
    with Section(a_xy, r_xy, a_prob, r_prob, tA, tR, sA, sR) as sec, sec.batch(rects) as b:
        b.candidates(radius, knn, False, 1.7)
        # fresh Delaunay per window on the host, as run_same does (same.py:1023)
        tris, off = [], [0]
        ka_off = b.offsets(L.KEEP_A)
        keepA = b.get(L.KEEP_A)
        for w in range(len(rects)):
            rows = keepA[ka_off[w]:ka_off[w + 1]]
            t = Delaunay(a_xy[rows]).simplices.astype(np.int32) if len(rows) >= 4 else np.zeros((0, 3), np.int32)
            tris.append(t)
            off.append(off[-1] + len(t))
        b.triangles_set(np.concatenate(tris), off)
        assert b.tri_classify(radius, 15.0, True) == 0
        b.tri_finalize(True, True, False)
        b.groups(2, None)
        xs = []
        for w in range(len(rects)):
            m = b.window_model(w)
            if len(m["pairs"]) == 0:
                assert len(O.find_knn_within_radius(a_xy[O.subset(a_xy, *rects[w])], r_xy[O.subset(r_xy, *rects[w])], radius, knn)[2]) == 0
                xs.append(np.zeros(0))
                continue
            res = _oracle_window(a_xy, r_xy, a_prob, r_prob, tA, tR, sA, sR, rects[w], **kw)
            assert np.array_equal(b.get_window(L.WIN_A, w), res["rowsA"])
            assert np.array_equal(m["keepA"], res["rowsA"][res["keepA"]])
            assert np.array_equal(m["keepR"], res["rowsR"][res["keepR"]])
            assert np.array_equal(m["pairs"], res["pairs"])
            assert np.array_equal(m["cost"], res["cost"])
            assert np.array_equal(m["tri"], res["tri"])
            assert np.array_equal(m["sign"], res["sign"])
            assert np.array_equal(m["weight"], res["weight"])
            assert np.array_equal(m["ref_group_node"], res["ref_group_node"])
            assert np.array_equal(m["ref_group_ptr"], res["ref_group_ptr"])
            assert np.array_equal(m["ref_group_idx"], res["ref_group_idx"])
            assert np.array_equal(m["ref_group_limit"], res["ref_group_limit"])
            xs.append(incumbent_rule(m["pairs"], seed * 100 + w, p_match=0.95))
        x = np.concatenate(xs)
        nv, nc, cuts = b.separation(x, cap=50)
        b.postsolve(x)
        for w in range(len(rects)):
            m = b.window_model(w)
            mj, mp = O.matching_from_x(xs[w], m["pairs"], len(m["keepA"]))
            viol, checked = O.separation(m["tri"], m["sign"], mj, r_xy[m["keepR"]])
            assert nv[w] == len(viol) and nc[w] == checked
            k = min(len(viol), 50)
            want = np.column_stack([mp[m["tri"][viol[:k], 0]], mp[m["tri"][viol[:k], 1]], mp[m["tri"][viol[:k], 2]], viol[:k]])
            assert np.array_equal(cuts[w, :k], want)
            ps = O.postsolve(m["tri"], a_xy[m["keepA"]], r_xy[m["keepR"]], mj)
            assert np.array_equal(b.get_window(L.TRI_MASK, w), ps["mask"])
            assert np.array_equal(b.get_window(L.FLIPPED, w).astype(bool), ps["flipped"])
            assert np.array_equal(b.get_window(L.AREA_BEFORE, w), ps["area_before"])
            # single-window separation call (the MIPSOL callback form) agrees with the batched one
            nv1, nc1, cuts1 = b.separation(xs[w], w, w + 1, cap=50)
            assert nv1[0] == nv[w] and nc1[0] == nc[w] and np.array_equal(cuts1[0, :k], cuts[w, :k])


@pytest.mark.parametrize("step", [11.0, 5.0])
def test_precomputed_triangulation_batched_vs_oracle(step):
    """Global triangulation in id space remapped into overlapping windows + unconstrained-node removal.  step 11: a row lies in at
    most four windows (the remap's one-sector row table); step 5: in up to nine (the general walk over the row's window list)."""
    from same_b200 import _lib as L
    from same_b200 import datagen
    from same_b200.device import Section
    ref, qry, ct = datagen.make_section_pair(n_tiles=9, n_types=3, seed=7)
    a_xy, r_xy = qry[["X", "Y"]].to_numpy(), ref[["X", "Y"]].to_numpy()
    a_prob, r_prob = qry[ct].to_numpy(), ref[ct].to_numpy()
    lut = {c: i for i, c in enumerate(ct)}
    tA = qry["cell_type"].map(lut).to_numpy(np.int32)
    tR = ref["cell_type"].map(lut).to_numpy(np.int32)
    sA, sR = np.ones(len(qry)), np.ones(len(ref))
    rng = np.random.default_rng(3)
    vid = (rng.permutation(len(qry)) * 5 + 11).astype(np.int64)
    tri_g = Delaunay(a_xy).simplices
    # keep only short triangles globally (what greedy_triangle_collapse's filter would do) so that windows
    # cut through the mesh and leave unconstrained nodes at their borders
    side = np.linalg.norm(a_xy[tri_g] - a_xy[np.roll(tri_g, 1, axis=1)], axis=2).max(axis=1)
    tri_g = tri_g[side < 0.9]
    tri_vid = vid[tri_g]
    ext = max(a_xy.max(), r_xy.max()) + 1
    rects = np.array([[x0, x0 + 14.0, y0, y0 + 14.0] for x0 in np.arange(0, ext, step) for y0 in np.arange(0, ext, step)])
    radius, knn = 1.0, 6
    kw = dict(radius=radius, knn=knn, dist_ct_coeff=1.0, min_angle_deg=12, ignore_same_type_triangles=True, max_matches=1)
    with Section(a_xy, r_xy, a_prob, r_prob, tA, tR, sA, sR) as sec, sec.batch(rects) as b:
        sec.set_triangles(tri_vid, vid)
        b.candidates(radius, knn)
        b.triangles_remap()
        assert b.tri_classify(radius, 12.0, True) == 0
        b.tri_finalize(True, True, remove_unconstrained=True)
        b.groups(1, None)
        n_unc = 0
        for w in range(len(rects)):
            ra, rr = O.subset(a_xy, *rects[w]), O.subset(r_xy, *rects[w])
            if len(ra) == 0 or len(rr) == 0:
                continue
            res = OP.window_pipeline(a_xy[ra], r_xy[rr], a_prob[ra], r_prob[rr], tA[ra], tR[rr], sA[ra], sR[rr],
                                     tri_global=tri_vid, a_vid=vid[ra], **kw)
            m = b.window_model(w)
            assert np.array_equal(m["keepA"], ra[res["keepA"]])
            assert np.array_equal(m["pairs"], res["pairs"])
            assert np.array_equal(m["cost"], res["cost"])
            assert np.array_equal(m["tri"], res["tri"])
            assert np.array_equal(m["sign"], res["sign"])
            assert np.array_equal(m["ref_group_node"], res["ref_group_node"])
            assert np.array_equal(m["ref_group_idx"], res["ref_group_idx"])
            n_unc += len(m["unconstrained"])
        assert n_unc > 0, "test should exercise unconstrained-node removal"


@pytest.mark.parametrize("knn", [1, 4, 9, 17, 33, 64])
def test_knn_capacities_vs_oracle(knn):
    from same_b200 import _lib as L
    from same_b200.device import Section
    rng = np.random.default_rng(knn)
    a, r = rng.uniform(0, 50, size=(3000, 2)), rng.uniform(0, 50, size=(3500, 2))
    with Section(a, r, np.zeros((3000, 1)), np.zeros((3500, 1))) as sec, sec.batch() as b:
        b.candidates(4.0, knn)
        keepA, keepR, pairs = O.find_knn_within_radius(a, r, 4.0, knn)
        assert np.array_equal(b.get(L.KEEP_A), keepA)
        assert np.array_equal(b.get(L.KEEP_R), keepR)
        assert np.array_equal(b.get(L.PAIRS), pairs)


def test_ties_on_exact_grid():
    """Duplicate distances (regular grid on both sides): our defined order is (d2, ref index)."""
    from same_b200 import _lib as L
    from same_b200.device import Section
    gx, gy = np.meshgrid(np.arange(30.0), np.arange(30.0))
    pts = np.stack([gx.ravel(), gy.ravel()], axis=1)
    with Section(pts, pts[::-1].copy(), np.zeros((900, 1)), np.zeros((900, 1))) as sec, sec.batch() as b:
        b.candidates(2.0, 8)
        keepA, keepR, pairs = O.find_knn_within_radius(pts, pts[::-1].copy(), 2.0, 8, brute=True)
        assert np.array_equal(b.get(L.PAIRS), pairs)


@pytest.mark.parametrize("knn", [7, 8, 9])
def test_near_ties_inside_one_distance_bucket(knn):
    """The search ranks on the high word of d2 first: neighbours whose d2 differ only in the low word (here by one ulp
    of the coordinate) must still come out in exact (d2, ref index) order — at the k-th boundary (exact fallback),
    inside the list (re-rank of the survivors) and just outside it."""
    from same_b200 import _lib as L
    from same_b200.device import Section
    rng = np.random.default_rng(5)
    near = rng.uniform(-1.5, 1.5, size=(6, 2))
    x1 = 3.0
    x2 = np.nextafter(x1, 4.0)
    x3 = np.nextafter(x2, 4.0)
    # stored farthest first so that memory order and index order both disagree with distance order
    twins = np.array([[x3, 0.0], [0.0, x2], [-x1, 0.0]])
    far = rng.uniform(4.0, 6.0, size=(20, 2)) * rng.choice([-1.0, 1.0], size=(20, 2))
    r = np.concatenate([twins, near, far])
    a = np.concatenate([np.zeros((1, 2)), rng.uniform(-3, 3, size=(40, 2))])
    with Section(a, r, np.zeros((len(a), 1)), np.zeros((len(r), 1))) as sec, sec.batch() as b:
        b.candidates(8.0, knn)
        keepA, keepR, pairs = O.find_knn_within_radius(a, r, 8.0, knn, brute=True)
        assert np.array_equal(b.get(L.KEEP_A), keepA) and np.array_equal(b.get(L.KEEP_R), keepR)
        got = b.get(L.PAIRS)
        assert np.array_equal(got, pairs)
        # row 0 really has the three twins at ranks 7, 8, 9 (0-based 6, 7, 8)
        kr = b.get(L.KEEP_R)
        mine = kr[got[got[:, 0] == 0][:, 1]]
        assert mine[6:knn].tolist() == [2, 1, 0][:max(0, knn - 6)]


def test_edge_cases():
    from same_b200 import _lib as L
    from same_b200.device import Section
    # nothing within radius -> zero pairs, empty keeps
    a, r = np.zeros((5, 2)), np.full((4, 2), 100.0)
    with Section(a, r, np.zeros((5, 2)), np.zeros((4, 2))) as sec, sec.batch() as b:
        b.candidates(1.0, 3)
        assert b.length(L.PAIRS) == 0 and b.length(L.KEEP_A) == 0 and b.length(L.KEEP_R) == 0
    # radius boundary is inclusive on d2 <= r*r; 3-4-5 triangle
    a = np.array([[0.0, 0.0]])
    r = np.array([[3.0, 4.0], [0.0, 5.0], [5.0, 0.0], [3.0000000000000004, 4.0], [0.0, 0.0]])
    with Section(a, r, np.zeros((1, 1)), np.zeros((5, 1))) as sec, sec.batch() as b:
        b.candidates(5.0, 8)
        assert sorted(b.get(L.KEEP_R).tolist()) == [0, 1, 2, 4]
    # empty windows inside a batch
    a = np.random.default_rng(0).uniform(0, 10, (200, 2))
    rects = np.array([[0, 5, 0, 5], [50, 60, 50, 60], [5, 10, 5, 10]], dtype=float)
    with Section(a, a + 0.01, np.zeros((200, 1)), np.zeros((200, 1))) as sec, sec.batch(rects) as b:
        b.candidates(0.5, 4)
        off = b.offsets(L.PAIRS)
        assert off[1] == off[2] and off[3] > off[2]
    # stage order is enforced
    with Section(a, a, np.zeros((200, 1)), np.zeros((200, 1))) as sec, sec.batch() as b:
        with pytest.raises(L.SameError):
            b.tri_finalize(True)
        with pytest.raises(L.SameError):
            b.candidates(1.0, 1000)


def test_batched_windows_priority_vs_oracle():
    """Cell-type priority KNN (src/knn_utils.py:31-65) for several overlapping windows in one batch: the claim of a reference cell
    is per window (the same cell seen from two windows is claimed independently), pairs / costs / groups equal the oracle's."""
    from same_b200 import _lib as L
    from same_b200 import datagen
    from same_b200.device import Section
    ref, qry, ct = datagen.make_section_pair(n_tiles=9, n_types=3, seed=21)
    a_xy, r_xy = qry[["X", "Y"]].to_numpy(), ref[["X", "Y"]].to_numpy()
    a_prob, r_prob = qry[ct].to_numpy(), ref[ct].to_numpy()
    lut = {c: i for i, c in enumerate(ct)}
    tA, tR = qry["cell_type"].map(lut).to_numpy(np.int32), ref["cell_type"].map(lut).to_numpy(np.int32)
    sA, sR = np.ones(len(qry)), np.ones(len(ref))
    ext = max(a_xy.max(), r_xy.max()) + 1
    step, ws = ext / 3 + 0.01, ext / 3 + 3.0
    rects = np.array([[i * step, i * step + ws, j * step, j * step + ws] for i in range(3) for j in range(3)])
    kw = dict(radius=1.3, knn=6, dist_ct_coeff=2.5, min_angle_deg=15, ignore_same_type_triangles=True, ignore_knn_if_matched=True, max_matches=1)
    n_single = 0
    with Section(a_xy, r_xy, a_prob, r_prob, tA, tR, sA, sR) as sec, sec.batch(rects) as b:
        b.candidates(1.3, 6, True, 2.5)
        b.groups(1, None)
        po, ko, ro = b.offsets(L.PAIRS), b.offsets(L.KEEP_A), b.offsets(L.KEEP_R)
        for w in range(len(rects)):
            res = _oracle_window(a_xy, r_xy, a_prob, r_prob, tA, tR, sA, sR, rects[w], **kw)   # runs the priority rule (pipeline.py)
            assert np.array_equal(b.get(L.KEEP_A, int(ko[w]), int(ko[w + 1])), res["rowsA"][res["keepA"]])
            assert np.array_equal(b.get(L.KEEP_R, int(ro[w]), int(ro[w + 1])), res["rowsR"][res["keepR"]])
            assert np.array_equal(b.get(L.PAIRS, int(po[w]), int(po[w + 1])), res["pairs"])
            assert np.array_equal(b.get(L.COST, int(po[w]), int(po[w + 1])), res["cost"])
            m = b.window_model(w)
            assert np.array_equal(m["ref_group_node"], res["ref_group_node"]) and np.array_equal(m["ref_group_idx"], res["ref_group_idx"])
            n_single += int((np.bincount(res["pairs"][:, 0]) == 1).sum())
    assert n_single > 50, "the priority rule should have collapsed many rows to their single claimed pair"


def test_page_locked_inputs_may_be_reused_after_wait_uploads():
    """same_section_create reads page-locked buffers asynchronously (auxiliary upload stream); after same_section_wait_uploads the
    caller may overwrite them — the section's results must not change."""
    from same_b200 import _lib as L
    from same_b200 import datagen
    from same_b200.device import Section, pinned_empty
    ref, qry, ct = datagen.make_section_pair(n_tiles=9, n_types=3, seed=31)
    src = dict(a_xy=qry[["X", "Y"]].to_numpy(), r_xy=ref[["X", "Y"]].to_numpy(), a_prob=qry[ct].to_numpy(), r_prob=ref[ct].to_numpy())
    pins = {}
    for k, v in src.items():
        pins[k] = pinned_empty(v.shape, np.float64)
        pins[k][...] = v
    with Section(pins["a_xy"], pins["r_xy"], pins["a_prob"], pins["r_prob"]) as sec:
        sec.wait_uploads()
        for v in pins.values():
            v[...] = -1.0e9                                  # scribble over the host buffers
        with sec.batch() as b:
            b.candidates(1.2, 6, False, 1.5)
            keepA, keepR, pairs = O.find_knn_within_radius(src["a_xy"], src["r_xy"], 1.2, 6)
            cost = O.pair_cost(pairs, src["a_xy"][keepA], src["r_xy"][keepR], src["a_prob"][keepA], src["r_prob"][keepR], 1.5)
            assert np.array_equal(b.get(L.PAIRS), pairs) and np.array_equal(b.get(L.COST), cost)


@pytest.mark.parametrize("n_types,knn", [(0, 4), (1, 3), (2, 8), (3, 5), (4, 8), (6, 7), (7, 8), (13, 6)])
def test_pair_cost_over_type_counts(n_types, knn):
    """The pair-cost kernel gathers packed row records [x, y, p_0..p_{K-1}, padding]: every record width (specialised 1-4
    double2 words, general above) and both power-of-two and other knn against the oracle, bit-exact (src/same.py:1183-1188)."""
    from same_b200 import _lib as L
    from same_b200.device import Section
    rng = np.random.default_rng(100 + n_types)
    a, r = rng.uniform(0, 30, size=(2500, 2)), rng.uniform(0, 30, size=(2700, 2))
    pa = rng.dirichlet(np.ones(max(n_types, 1)), size=2500)[:, :n_types]
    pr = rng.dirichlet(np.ones(max(n_types, 1)), size=2700)[:, :n_types]
    rects = np.array([[0, 17.0, 0, 31.0], [13.0, 31.0, 0, 31.0], [100.0, 101.0, 0, 1.0]])   # two overlapping windows and an empty one
    with Section(a, r, pa, pr) as sec, sec.batch(rects) as b:
        b.candidates(2.0, knn, False, 0.7)
        for w in range(2):
            ra, rr = O.subset(a, *rects[w]), O.subset(r, *rects[w])
            keepA, keepR, pairs = O.find_knn_within_radius(a[ra], r[rr], 2.0, knn)
            cost = O.pair_cost(pairs, a[ra][keepA], r[rr][keepR], pa[ra][keepA], pr[rr][keepR], 0.7)
            assert np.array_equal(b.get_window(L.PAIRS, w), pairs)
            assert np.array_equal(b.get_window(L.COST, w), cost)
