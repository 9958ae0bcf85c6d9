"""Drop-in parity of the public API on the GPU: `run_same` / `sliding_window_matching` of same_b200 against the golden
records of the UNMODIFIED reference (tests/golden/gen_golden.py).  Both sides talk to the same recording fake
`gurobipy` (oracle/ref_loader.py): every variable, objective coefficient, constraint (names, order, members) and lazy
cut is compared, then the returned matches frame column by column."""
import os
import sys

import numpy as np
import pandas as pd
import pytest

from tests.util import GOLDEN_CASES, golden_frame, golden_params, incumbent_rule, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fake_gurobi():
    from oracle import ref_loader
    saved = sys.modules.get("gurobipy")
    before = set(sys.modules)
    ref_loader.install_stubs()
    Model = ref_loader.Model
    if not hasattr(Model, "_orig_optimize"):
        Model._orig_optimize = Model.optimize

    def optimize(self, callback=None):   # same rule as gen_golden.patched_optimize
        Model._orig_optimize(self, callback)
        for terms, sense, rhs in self.lazy:
            for k, v in terms.items():
                if v < 0:
                    self.vars[k].x = 1.0
    Model.optimize = optimize
    ref_loader.MODELS.clear()
    yield ref_loader
    Model.optimize = Model._orig_optimize
    ref_loader.INCUMBENT_FN = None
    for name in set(sys.modules) - before:       # every stub module the fixture brought in goes away again
        if name.split(".")[0] in ("gurobipy", "scanpy", "alphashape", "matplotlib", "shapely"):
            sys.modules.pop(name, None)
    if saved is not None:
        sys.modules["gurobipy"] = saved
    else:
        sys.modules.pop("gurobipy", None)


def _incumbent_fn(seed):
    def fn(model):
        rp = np.asarray(model._row_ptr)
        rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        return incumbent_rule(np.column_stack([rows, rows]), seed)
    return fn


def _record(model):
    xs = [v.index for v in model.vars if v.VarName.startswith("x[")]
    qs = [v.index for v in model.vars if v.VarName.startswith("q_tri[")]
    pen = [v.index for v in model.vars if v.VarName.startswith("penalty[")]
    nom = [v.index for v in model.vars if v.VarName.startswith("no_match[")]
    obj = model.objective[0].terms
    out = dict(cost=np.array([obj.get(i, 0.0) for i in xs]), obj_q=np.array([obj.get(i, 0.0) for i in qs]),
               obj_penalty=np.array([obj.get(i, 0.0) for i in pen]), obj_no_match=np.array([obj.get(i, 0.0) for i in nom]),
               n_vars=len(model.vars))
    cname, csense, crhs, cptr, cidx, cval = [], [], [], [0], [], []
    for nm, sense, terms, rhs in model.constrs:
        cname.append(nm); csense.append(sense); crhs.append(rhs)
        cidx.extend(terms.keys()); cval.extend(terms.values()); cptr.append(len(cidx))
    out.update(con_name=np.asarray(cname), con_sense=np.asarray(csense), con_rhs=np.asarray(crhs, dtype=float), con_ptr=np.asarray(cptr),
               con_idx=np.asarray(cidx), con_val=np.asarray(cval, dtype=float))
    cuts = []
    for terms, sense, rhs in model.lazy:
        p = [k for k, v in terms.items() if v > 0]
        q = [k for k, v in terms.items() if v < 0]
        cuts.append([xs.index(p[0]), xs.index(p[1]), xs.index(p[2]), qs.index(q[0])])
    out["cuts"] = np.asarray(cuts, dtype=np.int64).reshape(-1, 4)
    return out


def _compare_models(ref_loader, g):
    assert len(ref_loader.MODELS) == int(g["n_models"])
    for w, model in enumerate(ref_loader.MODELS):
        rec = _record(model)
        for k, v in rec.items():
            want = g[f"w{w}_{k}"]
            if k == "n_vars":
                assert v == int(want), (w, k)
            else:
                assert np.array_equal(v, want), (w, k)


def _compare_matches(got: pd.DataFrame, g):
    cols = [str(c) for c in g["matches_columns"]]
    assert list(got.columns) == cols
    for c in cols:
        want = g[f"matches__{c}"]
        have = got[c].to_numpy()
        if c == "run_time":
            continue
        if have.dtype == object:
            have = have.astype("U")
        assert np.array_equal(have, want), c


def _mc_params(g):
    return dict(max_metacell_size=1, r_max=5, min_angle_deg=5, use_alpha_shape=False, alpha=None)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_public_api_vs_reference(case, fake_gurobi, tmp_path):
    import same_b200
    g = load_golden(case)
    ref_df, al_df = golden_frame(g, "ref"), golden_frame(g, "aligned")
    ct = [str(c) for c in g["commonCT"]]
    optim, gurobi = golden_params(g, "optim"), golden_params(g, "gurobi")
    id_col = str(g["id_col"])
    fake_gurobi.INCUMBENT_FN = _incumbent_fn(int(g["seed"]))
    sliding = case in ("fig2_script", "tiles4_sliding", "sparse_merge")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        if case == "fig2_script":
            mc_al = same_b200.greedy_triangle_collapse(al_df, cell_type_col="cell_type", original_idx_col=id_col, return_object=True, **_mc_params(g))
            mc_rf = same_b200.greedy_triangle_collapse(ref_df, cell_type_col="cell_type", original_idx_col=id_col, return_object=True, **_mc_params(g))
            assert np.array_equal(np.asarray(mc_al.metacell_delaunay, dtype=np.int64), g["mc_aligned_delaunay"])
            got = same_b200.sliding_window_matching(mc_rf, mc_al, commonCT=ct, outprefix=str(tmp_path / "out"), optim_params=dict(optim),
                                                    gurobi_params=dict(gurobi))
        elif sliding:
            got = same_b200.sliding_window_matching(ref_df, al_df, commonCT=ct, outprefix=str(tmp_path / "out"), optim_params=dict(optim),
                                                    gurobi_params=dict(gurobi))
        else:
            got, var_out = same_b200.run_same(ref_df, al_df, ct, outprefix=None, optim_params=dict(optim), gurobi_params=dict(gurobi))
    finally:
        os.chdir(cwd)
    _compare_models(fake_gurobi, g)
    _compare_matches(got, g)
    if not sliding:
        td = var_out["triangle_data"]
        T = len(td["triangles"])
        assert np.array_equal(np.array([td["areas_before"][t] for t in range(T)]), g["vo_areas_before"])
        assert np.array_equal(np.asarray(td["flipped_triangles"], dtype=np.int64), g["vo_flipped"])
        assert np.array_equal(np.asarray(list(td["triangle_info"].keys()), dtype=np.int64), g["vo_tri_info_order"])
        vio = var_out["violations"]
        s = vio["violation_summary"]
        assert [s["total_triangles"], s["violated_triangles"], s["total_comparisons"], s["total_violations"]] == g["vo_summary"].tolist()
        assert np.allclose([s["percent_triangles_violated"], s["percent_violations"]], g["vo_percent"], rtol=0, atol=0)
        xv = np.asarray([[d["triangle_idx"], d["point1"]["aligned_idx"], d["point2"]["aligned_idx"]] for d in vio["x_order_violations"]],
                        dtype=np.int64).reshape(-1, 3)
        yv = np.asarray([[d["triangle_idx"], d["point1"]["aligned_idx"], d["point2"]["aligned_idx"]] for d in vio["y_order_violations"]],
                        dtype=np.int64).reshape(-1, 3)
        assert np.array_equal(xv, g["vo_xviol"]) and np.array_equal(yv, g["vo_yviol"])
        assert sorted(vio["triangles_with_violations"]) == g["vo_tri_with_viol"].tolist()
        assert sorted(int(p) for p in vio["points_with_violations"]) == g["vo_pts_with_viol"].tolist()
        assert var_out["lazy_cuts_added"] == int(g["vo_lazy_cuts_added"])
        assert sorted(var_out["violation_penalty_comparison"]["points_both"]) == g["vo_points_both"].tolist()


def test_sliding_resume_from_checkpoint(fake_gurobi, tmp_path):
    """CSV checkpoint/resume (same.py:497-515, helpers.py:21-70): a second call skips windows already in matchedDF.csv."""
    import same_b200
    g = load_golden("tiles4_sliding")
    ref_df, al_df = golden_frame(g, "ref"), golden_frame(g, "aligned")
    ct = [str(c) for c in g["commonCT"]]
    optim, gurobi = golden_params(g, "optim"), golden_params(g, "gurobi")
    fake_gurobi.INCUMBENT_FN = _incumbent_fn(int(g["seed"]))
    out = str(tmp_path / "out")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        full = same_b200.sliding_window_matching(ref_df, al_df, commonCT=ct, outprefix=out, optim_params=dict(optim), gurobi_params=dict(gurobi))
        n_models = len(fake_gurobi.MODELS)
        saved = pd.read_csv(os.path.join(out, "matchedDF.csv"))
        keep_ids = sorted(saved["window_id"].unique())[:4]
        saved[saved["window_id"].isin(keep_ids)].to_csv(os.path.join(out, "matchedDF.csv"), index=False)
        fake_gurobi.MODELS.clear()
        again = same_b200.sliding_window_matching(ref_df, al_df, commonCT=ct, outprefix=out, optim_params=dict(optim), gurobi_params=dict(gurobi))
    finally:
        os.chdir(cwd)
    assert len(fake_gurobi.MODELS) == n_models - len(keep_ids)
    assert len(again) == len(full)
    key = ["window_id", "aligned_idx", "ref_idx"]
    a = full.sort_values(key).reset_index(drop=True)[key]
    b = again.sort_values(key).reset_index(drop=True)[key]
    assert a.equals(b.astype(a.dtypes.to_dict()))


def test_window_shard_union_equals_full(fake_gurobi, tmp_path):
    """Multi-GPU contract: the union of the rank shards equals the single-process result."""
    import same_b200
    g = load_golden("tiles4_sliding")
    ref_df, al_df = golden_frame(g, "ref"), golden_frame(g, "aligned")
    ct = [str(c) for c in g["commonCT"]]
    optim, gurobi = golden_params(g, "optim"), golden_params(g, "gurobi")
    fake_gurobi.INCUMBENT_FN = _incumbent_fn(int(g["seed"]))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        full = same_b200.sliding_window_matching(ref_df, al_df, commonCT=ct, optim_params=dict(optim), gurobi_params=dict(gurobi))
        parts = [same_b200.sliding_window_matching(ref_df, al_df, commonCT=ct, optim_params=dict(optim), gurobi_params=dict(gurobi),
                                                   window_shard=(r, 3)) for r in range(3)]
    finally:
        os.chdir(cwd)
    merged = pd.concat(parts, ignore_index=True)
    assert merged.drop(columns=["run_time"]).equals(full.drop(columns=["run_time"]))


def test_mirror_functions(fake_gurobi):
    """The helper functions keep the reference's signatures and results."""
    from same_b200.helpers import filter_triangles_by_radius
    from same_b200.knn_utils import find_knn_with_cell_type_priority
    from same_b200.same import _remap_triangles_by_vertex_ids
    from same_b200.utils import find_knn_within_radius
    g = load_golden("fig2_direct")
    o = golden_params(g, "optim")
    ref_df, al_df = golden_frame(g, "ref"), golden_frame(g, "aligned")
    a1, r1, pairs = find_knn_within_radius(al_df, ref_df, radius=o["radius"], knn=int(o["knn"]))
    assert np.array_equal(pairs, g["knn_pairs"]) and len(a1) == len(g["knn_keepA"]) and list(a1.index) == list(range(len(a1)))
    a2, r2, pp = find_knn_with_cell_type_priority(al_df, ref_df, o["radius"], knn=int(o["knn"]))
    assert np.array_equal(np.asarray(pp), g["prio_pairs"])
    pts = a1[["X", "Y"]].values
    kept, unc = filter_triangles_by_radius(pts, g["delaunay"], float(g["filt_tight_radius"]), aligned_df=a1, ignore_same_type_triangles=True,
                                           remove_unconstrained_nodes=True, min_angle_deg=o["min_angle_deg"])
    assert np.array_equal(np.asarray(kept).reshape(-1, 3), g["filt_tight"]) and sorted(unc) == g["filt_tight_unc"].tolist()
    out = _remap_triangles_by_vertex_ids(g["remap_tri_global"], g["remap_vid_all"][g["remap_rows"]])
    assert np.array_equal(out, g["remap_out"])


def test_errors(fake_gurobi):
    import same_b200
    g = load_golden("fig2_direct")
    ref_df, al_df = golden_frame(g, "ref"), golden_frame(g, "aligned")
    ct = [str(c) for c in g["commonCT"]]
    with pytest.raises(ValueError, match="No valid_pairs"):
        same_b200.run_same(ref_df, al_df, ct, optim_params=dict(radius=1e-9, knn=3, cell_id_col="cell_idx"))
    bad = ref_df.copy()
    bad["cell_type"] = "zzz"
    with pytest.raises(ValueError, match="Cell type categories differ"):
        same_b200.sliding_window_matching(bad, al_df, commonCT=ct, optim_params=dict(cell_id_col="cell_idx"))
    with pytest.raises(ValueError, match="not in aligned_df"):
        same_b200.run_same(ref_df, al_df, ct, aligned_delaunay=np.zeros((1, 3), int), aligned_delaunay_vertex_col="nope",
                           optim_params=dict(cell_id_col="cell_idx"))
    with pytest.raises(ValueError, match="shape"):
        same_b200.run_same(ref_df, al_df, ct, aligned_delaunay=np.zeros((4, 2), int), optim_params=dict(cell_id_col="cell_idx", radius=1.2))


@pytest.mark.parametrize("case", ["fig2_direct", "tiles4_sliding"])
def test_mip_start_through_public_api(case, fake_gurobi, tmp_path):
    """gurobi_params['init_method']='greedy' (src/same.py:1199-1215): the `.Start` values on x[...] and no_match[...] of every
    window's model equal the ones the unmodified reference sets (tests/golden/next/run_same_start.npz)."""
    import same_b200
    gs = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "next", "run_same_start.npz"))
    g = load_golden(case)
    ref_df, al_df = golden_frame(g, "ref"), golden_frame(g, "aligned")
    ct = [str(c) for c in g["commonCT"]]
    optim, gurobi = golden_params(g, "optim"), golden_params(g, "gurobi")
    gurobi["init_method"] = "greedy"
    optim["no_match_penalty"] = float(gs["no_match_penalty"])
    fake_gurobi.INCUMBENT_FN = _incumbent_fn(int(g["seed"]))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        if case == "tiles4_sliding":
            same_b200.sliding_window_matching(ref_df, al_df, commonCT=ct, outprefix=str(tmp_path / "out"), optim_params=dict(optim),
                                              gurobi_params=dict(gurobi))
        else:
            same_b200.run_same(ref_df, al_df, ct, outprefix=None, optim_params=dict(optim), gurobi_params=dict(gurobi))
    finally:
        os.chdir(cwd)
    assert len(fake_gurobi.MODELS) == int(gs[f"{case}__n_models"])
    for w, m in enumerate(fake_gurobi.MODELS):
        sx = np.asarray([np.nan if v.Start is None else v.Start for v in m.vars if v.VarName.startswith("x[")], dtype=np.float64)
        sn = np.asarray([np.nan if v.Start is None else v.Start for v in m.vars if v.VarName.startswith("no_match[")], dtype=np.float64)
        assert np.array_equal(sx, gs[f"{case}__w{w}_start_x"]), (case, w)
        assert np.array_equal(sn, gs[f"{case}__w{w}_start_no_match"]), (case, w)


def test_heart_metacells_vs_reference(fake_gurobi, tmp_path):
    """BASELINE configs[2]: the ISS heart sections (3,801 / 3,184 spots, K=8) collapsed to metacells of up to 10 cells and matched
    through sliding_window_matching with the paper script's parameters (examples/heart/run_same.sh): metacell frames, every
    window's model (pairs, costs, constraints, cuts) and the matches frame equal the unmodified reference's record."""
    import same_b200
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "next", "heart_mc10.npz")
    g = dict(np.load(path, allow_pickle=False))
    ref_df, al_df = golden_frame(g, "ref"), golden_frame(g, "aligned")
    ct = [str(c) for c in g["commonCT"]]
    optim, gurobi = golden_params(g, "optim"), golden_params(g, "gurobi")
    mcp = dict(max_metacell_size=int(g["mc_max_metacell_size"]), r_max=float(g["mc_r_max"]), min_angle_deg=float(g["mc_min_angle_deg"]),
               use_alpha_shape=False)
    mc_al = same_b200.greedy_triangle_collapse(al_df, cell_type_col="cell_type", original_idx_col="Cell_Num", return_object=True, **mcp)
    mc_rf = same_b200.greedy_triangle_collapse(ref_df, cell_type_col="cell_type", original_idx_col="Cell_Num", return_object=True, **mcp)
    for tag, mc in (("mca", mc_al), ("mcr", mc_rf)):
        mdf = mc.metacell_df
        assert np.array_equal(mdf[["X", "Y"]].to_numpy(), g[f"{tag}_xy"]), tag
        assert np.array_equal(mdf["size"].to_numpy(), g[f"{tag}_size"]), tag
        assert np.array_equal(mdf[ct].to_numpy(), g[f"{tag}_prob"]), tag
        assert np.array_equal(mdf["cell_type"].astype(str).to_numpy(), g[f"{tag}_type"]), tag
        assert np.array_equal(np.asarray([m for ms_ in mdf["members"] for m in ms_]), g[f"{tag}_members_flat"]), tag
        assert np.array_equal(np.asarray(mc.metacell_delaunay, dtype=np.int64).reshape(-1, 3), g[f"{tag}_delaunay"]), tag
    assert mc_al.metacell_df["size"].max() > 3
    fake_gurobi.INCUMBENT_FN = _incumbent_fn(int(g["seed"]))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        got = same_b200.sliding_window_matching(mc_rf, mc_al, commonCT=ct, outprefix=str(tmp_path / "out"), optim_params=dict(optim),
                                                gurobi_params=dict(gurobi))
    finally:
        os.chdir(cwd)
    _compare_models(fake_gurobi, g)
    _compare_matches(got, g)


def test_tongue_sections_vs_reference(fake_gurobi, tmp_path):
    """The shipped tongue sections (3,608 MERFISH reference / 4,671 protein query cells, K=5, fractional probabilities, UUID-string
    ids on one side) through greedy_triangle_collapse(MS=1) + sliding_window_matching with the paper script's parameters
    (examples/tongue/run_same.sh): every window's model and the matches frame equal the unmodified reference's record."""
    import same_b200
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "next", "tongue.npz")
    if not os.path.exists(path):
        pytest.skip("tongue fixture not generated")
    g = dict(np.load(path, allow_pickle=False))
    ref_df, al_df = golden_frame(g, "ref"), golden_frame(g, "aligned")
    ct = [str(c) for c in g["commonCT"]]
    optim, gurobi = golden_params(g, "optim"), golden_params(g, "gurobi")
    mcp = dict(max_metacell_size=1, r_max=300, min_angle_deg=15, use_alpha_shape=False)
    mc_al = same_b200.greedy_triangle_collapse(al_df, cell_type_col="cell_type", original_idx_col="Cell_Num", return_object=True, **mcp)
    mc_rf = same_b200.greedy_triangle_collapse(ref_df, cell_type_col="cell_type", original_idx_col="Cell_Num", return_object=True, **mcp)
    assert np.array_equal(np.asarray(mc_al.metacell_delaunay, dtype=np.int64), g["mc_aligned_delaunay"])
    fake_gurobi.INCUMBENT_FN = _incumbent_fn(int(g["seed"]))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        got = same_b200.sliding_window_matching(mc_rf, mc_al, commonCT=ct, outprefix=str(tmp_path / "out"), optim_params=dict(optim),
                                                gurobi_params=dict(gurobi))
    finally:
        os.chdir(cwd)
    assert int(g["n_models"]) >= 1
    _compare_models(fake_gurobi, g)
    _compare_matches(got, g)


def test_highs_cut_loop_same_solution_on_both_pipelines():
    """A real MIP solver in the loop (scipy/HiGHS: solve -> separate -> add cuts -> re-solve, capped at 24 cuts): `run_same` on the
    GPU path and the same loop over the CPU oracle's arrays with the oracle's separation reach the same cuts and the same
    solution vector — "final matches identical given identical candidates and costs" exercised with an actual solver."""
    import same_b200
    from oracle import oracle as O
    from same_b200.solver import HighsCutLoopBackend
    from tests.test_host_logic import _spec_from_oracle
    from tests.test_oracle_golden import _pipeline
    g = load_golden("simulated_st")
    ref_df, al_df = golden_frame(g, "ref"), golden_frame(g, "aligned")
    ct = [str(c) for c in g["commonCT"]]
    optim = golden_params(g, "optim")
    gurobi = {"time_limit": 120, "mip_gap": 0.0, "lazy_allowed_flip_fraction": 0.0, "lazy_max_cuts_per_incumbent": 12, "lazy_max_cuts": 24}
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        got, var_out = same_b200.run_same(ref_df, al_df, ct, outprefix=None, optim_params=dict(optim), gurobi_params=dict(gurobi), solver="highs")
    _, spec = _spec_from_oracle("simulated_st")
    o, res = _pipeline(g)
    r_xy = g["ref_xy"][res["keepR"]]
    cpu = HighsCutLoopBackend().solve(spec, lambda x, n: O.lazy_cuts(x, res["pairs"], res["tri"], res["sign"], r_xy, len(res["keepA"]), 0.0, 12, 24, n),
                                      dict(gurobi))
    assert cpu.status == "optimal" and cpu.cuts_added == var_out["lazy_cuts_added"] == 24
    assert np.array_equal(np.asarray(var_out["x"]) > 0.5, cpu.x > 0.5)
    sel = np.flatnonzero(cpu.x > 0.5)
    assert np.array_equal(got["aligned_idx"].to_numpy(), res["pairs"][sel, 0]) and np.array_equal(got["ref_idx"].to_numpy(), res["pairs"][sel, 1])


@pytest.mark.gpu
def test_pair_j_and_asynchronous_downloads_equal_the_blocking_path():
    """SAME_ARR_PAIR_J is PAIRS[:, 1]; ROW_PTR + PAIR_J rebuild valid_pairs; same_batch_get_many_async + same_batch_sync deliver the
    same bytes as the blocking call; CandidateStream (sections overlapped on several streams) returns what a plain batch returns."""
    from same_b200 import _lib as L
    from same_b200 import datagen
    from same_b200.device import CandidateStream, Section, pairs_from_rows
    ref, qry, ct = datagen.make_section_pair(n_tiles=9, n_types=3, seed=5)
    lut = {c: i for i, c in enumerate(ct)}
    D = dict(a_xy=qry[["X", "Y"]].to_numpy(), r_xy=ref[["X", "Y"]].to_numpy(), a_prob=qry[ct].to_numpy(), r_prob=ref[ct].to_numpy(),
             a_type=qry["cell_type"].map(lut).to_numpy(np.int32), r_type=ref["cell_type"].map(lut).to_numpy(np.int32))
    rects = np.array([[x, x + 20.0, y, y + 20.0] for x in (0.0, 17.5) for y in (0.0, 17.5)])
    frames = (D["a_xy"], D["r_xy"], D["a_prob"], D["r_prob"], D["a_type"], D["r_type"])
    with Section(*frames) as sec, sec.batch(rects) as b:
        b.candidates(1.0, 8, False, 1.0)
        want = b.get_many([L.KEEP_A, L.KEEP_R, L.ROW_PTR, L.PAIRS, L.COST, L.PAIR_J], pinned=False)
        assert np.array_equal(want[L.PAIR_J], want[L.PAIRS][:, 1])
        off_p, off_a = b.offsets(L.PAIRS), b.offsets(L.KEEP_A)
        for w in range(len(rects)):
            rp = want[L.ROW_PTR][off_a[w]:off_a[w + 1] + 1]
            assert np.array_equal(pairs_from_rows(rp, want[L.PAIR_J][off_p[w]:off_p[w + 1]]), want[L.PAIRS][off_p[w]:off_p[w + 1]])
        got = b.get_many([L.PAIR_J, L.COST, L.ROW_PTR], wait=False)
        b.sync()
        for k in got:
            assert np.array_equal(got[k], want[k])
        with pytest.raises(ValueError):
            b.get_many([L.COST], pinned=False, wait=False)
        j16 = b.get(L.PAIR_J16)                                  # two bytes per pair: the index is window-local
        assert j16.dtype == np.uint16 and np.array_equal(j16.astype(np.int32), want[L.PAIR_J])
    with CandidateStream(1.0, 8, depth=2, j16=True) as cs:
        out = cs.submit(frames, rects).result()
        assert out[L.PAIR_J].dtype == np.uint16 and np.array_equal(out[L.PAIR_J].astype(np.int32), want[L.PAIR_J])
        assert np.array_equal(out[L.COST], want[L.COST]) and np.array_equal(out[L.ROW_PTR], want[L.ROW_PTR])
    with CandidateStream(1.0, 8, depth=2) as cs:
        prev, n_done = None, 0
        for _ in range(5):                                       # two sections in flight at any time
            h = cs.submit(frames, rects)
            if prev is not None:
                out = prev.result()
                n_done += 1
                for k in (L.KEEP_A, L.KEEP_R, L.ROW_PTR, L.PAIR_J, L.COST):
                    assert np.array_equal(out[k], want[k])
                assert np.array_equal(out["offsets"][L.PAIRS], off_p)
            prev = h
        h2 = cs.submit(frames, rects)
        with pytest.raises(RuntimeError):
            cs.submit(frames, rects)                             # a third outstanding section is refused
        assert len(prev.result()[L.COST]) == len(want[L.COST]) and len(h2.result()[L.PAIR_J]) == len(want[L.PAIR_J])


@pytest.mark.gpu
def test_exact_predicate_diagnostic_on_near_collinear_triangles():
    """Triangles built to be almost degenerate (c = a + t (b - a), rounded): wherever the naive fp64 orientation sign (the
    reference's arithmetic, which the kernels reproduce) differs from the exact sign, the device filter must have listed the
    triangle, and the host's rational check must count exactly those — for the source signs and for a separation call.  The
    computed signs themselves stay the naive ones."""
    from same_b200 import _lib as L
    from same_b200 import helpers as H
    from same_b200.device import Section
    rng = np.random.default_rng(7)
    n_tri = 4000
    a, b_ = rng.uniform(0, 100, (n_tri, 2)), rng.uniform(0, 100, (n_tri, 2))
    c = a + rng.uniform(0.05, 0.95, (n_tri, 1)) * (b_ - a)
    c[::7] += rng.uniform(-1, 1, (len(c[::7]), 2))          # some ordinary triangles in between
    pts = np.concatenate([a, b_, c])
    tri = np.column_stack([np.arange(n_tri), np.arange(n_tri) + n_tri, np.arange(n_tri) + 2 * n_tri]).astype(np.int32)
    prob = np.full((len(pts), 2), 50.0)
    naive = np.array([H.naive_orientation_sign(pts[t[0]], pts[t[1]], pts[t[2]]) for t in tri])
    exact = np.array([H.exact_orientation_sign(pts[t[0]], pts[t[1]], pts[t[2]]) for t in tri])
    differs = np.flatnonzero(naive != exact)
    assert len(differs) > 20, "the construction should produce sign disagreements"
    with Section(pts, pts, prob, prob) as sec, sec.batch() as b:
        b.candidates(1e-6, 1, False, 1.0)                    # every cell pairs with its twin only
        assert np.array_equal(b.get(L.KEEP_A), np.arange(len(pts))) and np.array_equal(b.get(L.PAIRS)[:, 1], np.arange(len(pts)))
        b.triangles_set(tri, [0, n_tri])
        b.tri_classify(1e9, None, False)
        b.tri_finalize(False, True, False)
        assert np.array_equal(b.get(L.TRI), tri)
        assert np.array_equal(b.get(L.TRI_SIGN), naive)      # the path keeps the reference's naive signs
        n0, listed0 = b.uncertain(0)
        assert n0 == len(listed0) and set(differs.tolist()) <= set(listed0.tolist())
        assert not (listed0 % 7 == 0).any()                  # the ordinary triangles pass the filter
        chk0 = H.exact_predicate_check(b, 0, 0, pts, pts)
        assert chk0["naive_differs_from_exact"] == len(differs) and chk0["triangles"] == differs.tolist()
        x = np.ones(len(pts))
        nv, nc, cuts = b.separation(x, cap=10)
        assert nv[0] == 0 and nc[0] == int((naive != 0).sum())          # identity mapping: nothing flips
        n1, listed1 = b.uncertain(1)
        assert np.array_equal(listed1, listed0) and b.uncertain(1, cap=0)[0] == n1
        chk1 = H.exact_predicate_check(b, 0, 1, pts, pts, b.get(L.MATCH_J))
        assert chk1["naive_differs_from_exact"] == len(differs)


@pytest.mark.gpu
def test_incidence_csr_equals_aligned_simplex_map():
    """SAME_ARR_NODE_TRI_PTR / LEN / IDX (a8): every kept aligned node's triangles, ascending — the reference's aligned_simplex_map
    (src/same.py:1096-1099) as CSR, per window of a batch, also after the unconstrained-node removal renumbered the nodes."""
    from scipy.spatial import Delaunay
    from same_b200 import _lib as L
    from same_b200 import datagen
    from same_b200.device import Section
    ref, qry, ct = datagen.make_section_pair(n_tiles=9, n_types=3, seed=11)
    lut = {c: i for i, c in enumerate(ct)}
    a_xy, r_xy = qry[["X", "Y"]].to_numpy(), ref[["X", "Y"]].to_numpy()
    rects = np.array([[x, x + 22.0, y, y + 22.0] for x in (0.0, 16.0) for y in (0.0, 16.0)])
    with Section(a_xy, r_xy, qry[ct].to_numpy(), ref[ct].to_numpy(), qry["cell_type"].map(lut).to_numpy(np.int32),
                 ref["cell_type"].map(lut).to_numpy(np.int32)) as sec:
        sec.set_triangles(Delaunay(a_xy).simplices.astype(np.int64), None)
        with sec.batch(rects) as b:
            b.candidates(1.0, 8, False, 1.0)
            b.triangles_remap()
            b.tri_classify(0.8, 25.0, True)
            b.tri_finalize(True, True, True)
            ptr, ln, idx = b.get(L.NODE_TRI_PTR), b.get(L.NODE_TRI_LEN), b.get(L.NODE_TRI_IDX)
            ka, to = b.offsets(L.KEEP_A), b.offsets(L.TRI)
            assert len(ptr) == ka[-1] + 1 and len(idx) == 3 * to[-1] and np.array_equal(b.offsets(L.NODE_TRI_IDX), 3 * to)
            total = 0
            for w in range(len(rects)):
                tri = b.get_window(L.TRI, w)
                n = int(ka[w + 1] - ka[w])
                want = {i: set() for i in range(n)}
                for t, simplex in enumerate(tri.tolist()):
                    for v in simplex:
                        want[v].add(t)
                for i in range(n):
                    got = idx[ptr[ka[w] + i]:ptr[ka[w] + i] + ln[ka[w] + i]]
                    assert (np.diff(got) > 0).all() and set(got.tolist()) == want[i]
                    total += len(got)
            assert total == 3 * to[-1] == ptr[-1]


@pytest.mark.gpu
def test_missing_cell_type_raises_like_the_reference_and_timings_are_recorded(tmp_path):
    """Without a `cell_type` column the reference fails with KeyError whenever it has to read it (src/helpers.py:329,
    src/knn_utils.py:37); nothing is silently treated as "one type".  A successful run leaves machine-readable stage timers
    (var_out['timings'], window_<id>/timings.json)."""
    import json
    import same_b200
    from same_b200 import datagen
    from same_b200.solver import IncumbentBackend
    ref, qry, ct = datagen.make_section_pair(n_tiles=1, n_types=3, seed=2)
    optim = dict(radius=1.0, knn=6, max_matches=1, min_angle_deg=10, cell_id_col="Cell_Num_Old")

    def inc(spec):
        rp = np.asarray(spec.row_ptr, dtype=np.int64)
        rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        return incumbent_rule(np.column_stack([rows, rows]), 1)
    no_type = qry.drop(columns=["cell_type"])
    with pytest.raises(KeyError):
        same_b200.run_same(ref, no_type, list(ct), optim_params=dict(optim, ignore_same_type_triangles=True), solver=IncumbentBackend(inc))
    with pytest.raises(KeyError):
        same_b200.run_same(ref, no_type, list(ct), optim_params=dict(optim, ignore_same_type_triangles=False, ignore_knn_if_matched=True),
                           solver=IncumbentBackend(inc))
    out, var_out = same_b200.run_same(ref, no_type, list(ct), outprefix=str(tmp_path / "w"),
                                      optim_params=dict(optim, ignore_same_type_triangles=False), solver=IncumbentBackend(inc))
    assert len(out) > 0
    t = var_out["timings"]
    assert {"fetch_model_arrays_s", "prepare_model_s", "solve_s", "post_solve_analysis_s", "n_pairs", "n_triangles", "separation_calls"} <= set(t)
    assert json.load(open(tmp_path / "w" / "timings.json")) == t
    epc = var_out["exact_predicate_check"]
    assert epc["separation_calls"] == 1 and epc["source_signs"]["naive_differs_from_exact"] == 0


@pytest.mark.gpu
def test_candidate_stream_errors_and_pool_accounting():
    """A section that cannot be built surfaces its error in result() and frees its slot; the memory-pool calls behave."""
    from same_b200 import _lib as L
    from same_b200 import datagen
    from same_b200.device import CandidateStream
    ref, qry, ct = datagen.make_section_pair(n_tiles=1, n_types=3, seed=3)
    lut = {c: i for i, c in enumerate(ct)}
    good = (qry[["X", "Y"]].to_numpy(), ref[["X", "Y"]].to_numpy(), qry[ct].to_numpy(), ref[ct].to_numpy(),
            qry["cell_type"].map(lut).to_numpy(np.int32), ref["cell_type"].map(lut).to_numpy(np.int32))
    bad = (good[0], good[1], good[2], good[3][:, :2], good[4], good[5])          # probability blocks of different widths
    with CandidateStream(1.0, 8, depth=2) as cs:
        h_bad, h_good = cs.submit(bad), cs.submit(good)
        with pytest.raises(ValueError):
            h_bad.result()
        out = h_good.result()
        assert len(out[L.PAIR_J]) == len(out[L.COST]) > 0
        h2 = cs.submit(good)                                                     # both slots are free again
        assert np.array_equal(h2.result()[L.PAIR_J], out[L.PAIR_J])
        r0, u0 = L.mempool_stats(0)
        cs.reserve(8 << 20)
        r1, u1 = L.mempool_stats(0)
        assert r1 >= r0 and r1 >= 8 << 20 and u1 <= r1
    with pytest.raises(L.SameError):
        L.mempool_reserve(0, -1)
