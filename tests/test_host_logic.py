"""CPU tests of the host-side logic: parameter dictionaries, the sliding-window walk, metacell triangulation,
sharding + the two small collectives (gloo, world_size 2)."""
import os
import socket
import sys

import numpy as np
import pandas as pd
import pytest

from tests.util import golden_frame, golden_params, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_param_defaults_are_the_reference_contract():
    from same_b200 import init_gurobi_params, init_optim_params
    assert init_optim_params() == {
        "window_size": 1000, "overlap": 250, "min_cells_per_window": 10, "max_matches": 1, "ref_metacell_match_multiplier": None,
        "radius": 250, "penalty_coeff": 100, "no_match_penalty": 100, "delaunay_penalty": 5, "dist_ct_coeff": 1, "knn": 8,
        "cell_id_col": "Cell_Num_Old", "hard_spatial_constraints": False, "ignore_same_type_triangles": True,
        "ignore_knn_if_matched": False, "lazy_constraints": True, "min_angle_deg": 15}
    assert init_gurobi_params() == {
        "time_limit": 7200, "mip_gap": 0.05, "mip_focus": 2, "cuts": 2, "heuristics": 0.1, "init_method": None, "init_big_m": 1e9,
        "init_hungarian_max_n": 5000, "lazy_max_cuts": None, "lazy_allowed_flip_fraction": 0.05, "lazy_max_cuts_per_incumbent": 1000}
    assert init_optim_params(radius=5, knn=3)["radius"] == 5 and init_gurobi_params(time_limit=1)["time_limit"] == 1


def test_public_names():
    import same_b200
    for n in ["init_gurobi_params", "init_optim_params", "sliding_window_matching", "run_same", "merge_window_matches_unique_ref",
              "MetaCell", "greedy_triangle_collapse", "unpack_metacell_matches"]:
        assert n in same_b200.__all__ and callable(getattr(same_b200, n))


@pytest.mark.parametrize("case", ["tiles4_sliding", "sparse_merge", "fig2_script"])
def test_window_walk_matches_reference(case):
    """enumerate_windows reproduces the reference's grid walk incl. the small-window merge rule: the windows it runs are
    exactly the models the reference built, and every window_id in the reference's output is one of ours."""
    from same_b200 import windows as WN
    g = load_golden(case)
    o = golden_params(g, "optim")
    rxy, axy = g["ref_xy"], g["aligned_xy"]
    x_min, x_max = min(rxy[:, 0].min(), axy[:, 0].min()), max(rxy[:, 0].max(), axy[:, 0].max())
    y_min, y_max = min(rxy[:, 1].min(), axy[:, 1].min()), max(rxy[:, 1].max(), axy[:, 1].max())
    xw, yw = WN.window_grid(x_min, x_max, y_min, y_max, int(o["window_size"]), int(o["overlap"]))
    count = WN.numpy_counter(rxy, axy)
    wins = WN.enumerate_windows(xw, yw, int(o["window_size"]), int(o["overlap"]), int(o["min_cells_per_window"]), count,
                                (int(x_min), int(x_max), int(y_min), int(y_max)))
    runnable = [w for w in wins if w.run]
    assert len(runnable) == int(g["n_models"])
    ids = [w.window_id for w in runnable]
    assert set(np.unique(g["matches__window_id"]).tolist()) <= set(ids)
    # the batched GPU counter sees the same rectangles
    rects, keys = WN.candidate_rects(xw, yw, int(o["window_size"]))
    cr = np.array([count(tuple(r))[0] for r in rects])
    ca = np.array([count(tuple(r))[1] for r in rects])
    wins2 = WN.enumerate_windows(xw, yw, int(o["window_size"]), int(o["overlap"]), int(o["min_cells_per_window"]),
                                 WN.table_counter(keys, cr, ca), (int(x_min), int(x_max), int(y_min), int(y_max)))
    assert [(w.i, w.j, w.rect, w.window_id, w.central, w.run) for w in wins] == [(w.i, w.j, w.rect, w.window_id, w.central, w.run) for w in wins2]
    # per-window pair counts of the reference = number of x variables of each model; sizes sanity
    for k, w in enumerate(runnable):
        assert w.n_ref >= int(o["min_cells_per_window"]) and w.n_moving >= int(o["min_cells_per_window"])


def test_sparse_merge_case_really_merges():
    from same_b200 import windows as WN
    g = load_golden("sparse_merge")
    o = golden_params(g, "optim")
    rxy, axy = g["ref_xy"], g["aligned_xy"]
    x_min, x_max = min(rxy[:, 0].min(), axy[:, 0].min()), max(rxy[:, 0].max(), axy[:, 0].max())
    y_min, y_max = min(rxy[:, 1].min(), axy[:, 1].min()), max(rxy[:, 1].max(), axy[:, 1].max())
    xw, yw = WN.window_grid(x_min, x_max, y_min, y_max, int(o["window_size"]), int(o["overlap"]))
    wins = WN.enumerate_windows(xw, yw, int(o["window_size"]), int(o["overlap"]), int(o["min_cells_per_window"]), WN.numpy_counter(rxy, axy),
                                (int(x_min), int(x_max), int(y_min), int(y_max)))
    ws = int(o["window_size"])
    assert any((w.rect[1] - w.rect[0] > ws) or (w.rect[3] - w.rect[2] > ws) for w in wins), "fixture should trigger the merge rule"


def test_metacell_size1_triangulation_matches_reference():
    """greedy_triangle_collapse(max_metacell_size=1) only filters the Delaunay triangulation (examples/synthetic/run_same.sh:85-95)."""
    import same_b200
    g = load_golden("fig2_script")
    al = golden_frame(g, "aligned")
    mc = same_b200.greedy_triangle_collapse(al, cell_type_col="cell_type", original_idx_col=str(g["id_col"]), return_object=True,
                                            max_metacell_size=1, r_max=5, min_angle_deg=5, use_alpha_shape=False, alpha=None)
    assert np.array_equal(np.asarray(mc.metacell_delaunay, dtype=np.int64), g["mc_aligned_delaunay"])
    assert np.array_equal(mc.metacell_df["metacell_id"].to_numpy(), g["mc_aligned_metacell_id"])
    assert list(mc.metacell_df["size"].unique()) == [1] and mc.metacell_idx_col == "metacell_id"
    assert mc.metacell_members(3) == [al[str(g["id_col"])].iloc[3]]


def test_merge_window_matches_unique_ref():
    import same_b200
    a = pd.DataFrame({"window_id": [0, 0], "Aligned_Cell_Num_Old": [1, 2], "Ref_Cell_Num_Old": [10, 10], "X": 0.0, "Y": 0.0,
                      "filtered_violation": [False, False]})
    b = pd.DataFrame({"window_id": [1, 1], "Aligned_Cell_Num_Old": [2, 3], "Ref_Cell_Num_Old": [11, 11], "X": 0.0, "Y": 0.0,
                      "filtered_violation": [True, False]})
    out = same_b200.merge_window_matches_unique_ref([a, b])
    assert out["Aligned_Cell_Num_Old"].is_unique and out["Ref_Cell_Num_Old"].is_unique and len(out) == 2
    assert same_b200.merge_window_matches_unique_ref([]).empty


def test_shard_windows_partitions():
    from same_b200.windows import shard_windows
    for n in (0, 1, 7, 49, 392):
        for world in (1, 2, 3, 8):
            blocks = [shard_windows(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_merge_window_matches_equals_reference_under_fixed_hash_seed():
    """merge_window_matches_unique_ref (src/helpers.py:692-815) row for row against the unmodified reference.  The reference's own
    result depends on PYTHONHASHSEED (networkx walks a set of string labels), so both the record
    (tests/golden/next/gen_golden_merge.py) and this comparison run in a process with PYTHONHASHSEED=0."""
    import subprocess
    code = (
        "import sys, numpy as np, importlib.util\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        f"spec = importlib.util.spec_from_file_location('gm', {os.path.join(ROOT, 'tests', 'golden', 'next', 'gen_golden_merge.py')!r})\n"
        "gm = importlib.util.module_from_spec(spec); spec.loader.exec_module(gm)\n"
        "from same_b200.merge import merge_window_matches_unique_ref\n"
        "out = merge_window_matches_unique_ref([f.copy() for f in gm.make_input()])\n"
        f"g = np.load({os.path.join(ROOT, 'tests', 'golden', 'next', 'merge.npz')!r})\n"
        "assert len(out) == len(g['out_aligned'])\n"
        "assert np.array_equal(out['window_id'].to_numpy(), g['out_window'])\n"
        "assert np.array_equal(out['Aligned_Cell_Num_Old'].to_numpy(), g['out_aligned'])\n"
        "assert np.array_equal(out['Ref_Cell_Num_Old'].to_numpy(), g['out_ref'])\n"
        "assert np.array_equal(out['X'].to_numpy(), g['out_x']) and np.array_equal(out['filtered_violation'].to_numpy(bool), g['out_fv'])\n"
        "assert out['Aligned_Cell_Num_Old'].is_unique and out['Ref_Cell_Num_Old'].is_unique\n"
        "print('merge parity ok', len(out))\n")
    env = dict(os.environ, PYTHONHASHSEED="0")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "merge parity ok" in r.stdout, r.stdout + r.stderr


# ---- gloo, world_size 2 ---------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        from same_b200 import sharding as S
        from same_b200.windows import shard_windows
        # each rank owns a strip of 100 units; rows below strip_lo + 30 are the border band its predecessor needs
        rng = np.random.default_rng(rank)
        n = 50 + 10 * rank
        xy = np.column_stack([rng.uniform(0, 100, n), rng.uniform(100 * rank, 100 * (rank + 1), n)])
        val = np.arange(n, dtype=np.float64)[:, None] + 1000 * rank
        halo, info = S.exchange_halo({"xy": xy, "val": val, "y": xy[:, 1:2]}, "y", 100 * rank + 30.0)
        # row-sharded upload: every rank holds the same array, moves 1/world of it and gathers the rest
        full = np.random.default_rng(99).uniform(0, 1, (101, 3))
        codes = (np.arange(37) % 5).astype(np.int32)
        got_full, got_codes = S.allgather_rows(full).numpy(), S.allgather_rows(codes).numpy()
        assert np.array_equal(got_full, full) and np.array_equal(got_codes[:, 0], codes) and got_codes.dtype == np.int32
        lo, hi = shard_windows(7, world, rank)
        local = pd.DataFrame({"window_id": list(range(lo, hi)), "rank": rank})
        merged = S.gather_matches(local)
        q.put((rank, halo["xy"], halo["val"], info, merged, xy, val))
    finally:
        dist.destroy_process_group()


def test_gloo_halo_exchange_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r = q.get(timeout=120)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # rank 0 received exactly rank 1's border band; the last rank receives nothing
    xy1, val1 = res[1][5], res[1][6]
    band = xy1[:, 1] < 130.0
    assert np.array_equal(res[0][1], xy1[band]) and np.array_equal(res[0][2], val1[band])
    assert len(res[1][1]) == 0 and res[0][3]["rows"] == int(band.sum())
    for r in (0, 1):
        assert res[r][4]["window_id"].tolist() == list(range(7))
        assert res[r][4]["rank"].tolist() == [0, 0, 0, 0, 1, 1, 1]


# ---- the host model builder fed from the CSR/CSC arrays ------------------------------------------
def _spec_from_oracle(case):
    """ModelSpec of one golden window, its arrays taken from the CPU oracle (the same arrays the GPU batch returns)."""
    from same_b200.solver import ModelSpec
    from tests.test_oracle_golden import _pipeline
    from tests.util import load_golden
    g = load_golden(case)
    o, res = _pipeline(g)
    P, na = len(res["pairs"]), len(res["keepA"])
    row_ptr = np.searchsorted(res["pairs"][:, 0], np.arange(na + 1)).astype(np.int64)
    a_size = (g["aligned_size"] if "aligned_size" in g else np.ones(len(g["aligned_xy"])))[res["keepA"]]
    spec = ModelSpec(n_pairs=P, n_ref=len(res["keepR"]), n_aligned=na, n_tri=len(res["tri"]), cost=res["cost"], row_ptr=row_ptr,
                     ref_group_node=res["ref_group_node"], ref_group_ptr=res["ref_group_ptr"], ref_group_idx=res["ref_group_idx"],
                     ref_group_limit=res["ref_group_limit"], aligned_size=a_size, tri_weight=res["weight"],
                     penalty_coeff=float(o.get("penalty_coeff", 100)), no_match_penalty=float(o.get("no_match_penalty", 100)),
                     delaunay_penalty=float(o.get("delaunay_penalty", 5)))
    return g, spec


@pytest.mark.parametrize("case", ["fig2_direct", "fig2_priority", "fig2_script", "simulated_st", "uniform_k8"])
def test_model_matrices_equal_reference_model(case):
    """model_matrices assembles, with array operations only, exactly the constraint list and objective the reference builds
    row by row (src/helpers.py:130-158, src/same.py:1191-1197), as recorded from the unmodified reference."""
    from same_b200.solver import model_matrices
    g, spec = _spec_from_oracle(case)
    mm = model_matrices(spec)
    for key, ref in (("names", "con_name"), ("sense", "con_sense"), ("rhs", "con_rhs"), ("ptr", "con_ptr"), ("idx", "con_idx"), ("val", "con_val")):
        assert np.array_equal(mm[key], g[f"w0_{ref}"]), key
    assert mm["n_vars"] == int(g["w0_n_vars"])
    P, nr, na = spec.n_pairs, spec.n_ref, spec.n_aligned
    assert np.array_equal(mm["obj"][:P], g["w0_cost"])
    assert np.array_equal(mm["obj"][P:P + nr], g["w0_obj_penalty"])
    assert np.array_equal(mm["obj"][P + nr:P + nr + na], g["w0_obj_no_match"])
    assert np.array_equal(mm["obj"][P + nr + na:], g["w0_obj_q"])


@pytest.mark.parametrize("case", ["fig2_direct", "simulated_st", "uniform_k8"])
@pytest.mark.parametrize("backend", ["gurobi", "gurobi_rows"])
def test_gurobi_builders_record_the_reference_model(case, backend, tmp_path, monkeypatch):
    """Both Gurobi model builders — the default matrix builder (addMVar + addMConstr) and the row-by-row one — executed against the
    recording fake gurobipy: variables, objective, every constraint (name, order, members, sense, right-hand side), the lazy cuts
    of one callback and the MIP start land in the model exactly as the unmodified reference recorded them."""
    import sys
    from oracle import ref_loader
    from same_b200.solver import get_backend
    from tests.test_gpu_api import _record
    saved = sys.modules.get("gurobipy")
    ref_loader.install_stubs()
    ref_loader.MODELS.clear()
    monkeypatch.chdir(tmp_path)
    try:
        g, spec = _spec_from_oracle(case)
        P = spec.n_pairs
        ref_loader.INCUMBENT_FN = lambda model: (np.arange(P) % 3 == 0).astype(float)
        cuts = np.array([[0, 1, 2, 0], [3, 4, 5, min(1, spec.n_tri - 1)]], dtype=np.int64)
        seen = {}

        def separate(x_vals, cuts_so_far):
            seen["x"] = np.asarray(x_vals).copy()
            return cuts
        start = (np.r_[1.0, np.zeros(P - 1)], np.ones(spec.n_aligned))
        res = get_backend(backend).solve(spec, separate, {"time_limit": 10, "mip_gap": 0.05}, start=start)
        model = ref_loader.MODELS[-1]
        rec = _record(model)
        for k in ("cost", "obj_q", "obj_penalty", "obj_no_match", "con_name", "con_sense", "con_rhs", "con_ptr", "con_idx", "con_val"):
            assert np.array_equal(rec[k], g[f"w0_{k}"]), (backend, k)
        assert rec["n_vars"] == int(g["w0_n_vars"])
        assert np.array_equal(rec["cuts"], cuts) and res.cuts_added == 2 and res.status == "optimal"
        assert np.array_equal(seen["x"], (np.arange(P) % 3 == 0).astype(float)) and np.array_equal(res.x, seen["x"])
        xs = [v for v in model.vars if v.VarName.startswith("x[")]
        nm = [v for v in model.vars if v.VarName.startswith("no_match[")]
        assert [v.Start for v in xs] == start[0].tolist() and [v.Start for v in nm] == start[1].tolist()
        assert [v.VarName for v in model.vars[:2]] == ["x[0]", "x[1]"] and model.vars[-1].VarName == f"q_tri[{spec.n_tri - 1}]"
    finally:
        ref_loader.INCUMBENT_FN = None
        if saved is not None:
            sys.modules["gurobipy"] = saved
        else:
            sys.modules.pop("gurobipy", None)


def test_highs_backend_solves_from_model_matrices():
    """The HiGHS stand-in consumes the same matrix; on the 144-cell known-answer fixture it returns a feasible one-to-one matching."""
    from same_b200.solver import HighsCutLoopBackend, model_matrices
    g, spec = _spec_from_oracle("simulated_st")
    res = HighsCutLoopBackend().solve(spec, None, {"time_limit": 60, "mip_gap": 0.05})
    assert res.status in ("optimal", "time_limit")
    mm = model_matrices(spec)
    from scipy.sparse import csr_matrix
    A = csr_matrix((mm["val"], mm["idx"], mm["ptr"]), shape=(len(mm["rhs"]), mm["n_vars"]))
    full = np.concatenate([res.x, res.penalty, res.no_match, res.q])
    lhs = A @ full
    eq = mm["sense"] == "=="
    assert np.all(lhs[~eq] <= mm["rhs"][~eq] + 1e-6) and np.allclose(lhs[eq], mm["rhs"][eq], atol=1e-6)


def test_metacell_filter_thresholds_follow_reference_expressions():
    """Triangles exactly on the r_max / min-angle thresholds (a 3-4-5 lattice: edges of exactly r_max, a right isosceles triangle
    against min_angle_deg=45) are decided by the reference's scalar expressions (src/metacell_utils.py:233-262)."""
    from same_b200.metacell_utils import _reference_valid, _valid_triangles
    coords = np.array([[0.0, 0.0], [30.0, 40.0], [30.0, 0.0], [0.0, 40.0], [10.0, 0.0], [0.0, 10.0], [7.3, 9.1], [50.0, 0.0]])
    tri = np.array([[0, 1, 2], [0, 3, 1], [0, 4, 5], [4, 6, 5], [0, 7, 1]])
    for r_max, ang in ((50.0, None), (50.0, 45.0), (49.99999999999999, 30.0), (None, 45.0), (50.00000000000001, 36.86989764584402)):
        got = _valid_triangles(coords, tri, r_max, ang)
        want = np.array([_reference_valid(coords[a], coords[b], coords[c], r_max, ang) for a, b, c in tri])
        assert np.array_equal(got, want), (r_max, ang)
    assert _valid_triangles(coords, tri, 50.0, None)[0] and not _valid_triangles(coords, tri, 49.99999999999999, None)[0]


def test_default_device_follows_local_rank(monkeypatch):
    """One process per GPU: the device of a process is SAME_B200_DEVICE, else LOCAL_RANK (torchrun), else 0."""
    from same_b200.device import default_device
    monkeypatch.delenv("SAME_B200_DEVICE", raising=False)
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    assert default_device() == 0
    monkeypatch.setenv("LOCAL_RANK", "3")
    assert default_device() == 3
    monkeypatch.setenv("SAME_B200_DEVICE", "5")
    assert default_device() == 5
    monkeypatch.setenv("SAME_B200_DEVICE", "junk")
    assert default_device() == 3


def test_host_threads_follow_the_environment(monkeypatch):
    """run_same triangulates the windows on `host_threads()` threads: SAME_B200_HOST_THREADS, else the cores of the process."""
    from same_b200.same import host_threads
    monkeypatch.setenv("SAME_B200_HOST_THREADS", "3")
    assert host_threads() == 3
    monkeypatch.setenv("SAME_B200_HOST_THREADS", "0")
    assert host_threads() == 1
    monkeypatch.delenv("SAME_B200_HOST_THREADS")
    assert 1 <= host_threads() <= (os.cpu_count() or 1)
