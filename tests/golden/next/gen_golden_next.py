"""Golden vectors for the SURVEY §8(f) rows (MIP start, metacell collapse), produced by the UNMODIFIED reference.

    python tests/golden/next/gen_golden_next.py       # writes tests/golden/next/{mip_start,collapse}.npz

Same rules as gen_golden.py: runs only in the build container (imports /root/reference under oracle/ref_loader.py's
stubs); the committed .npz files are what the tests read.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.dirname(HERE)                      # the run_same fixtures these cases start from
ROOT = os.path.dirname(os.path.dirname(GOLD))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from same_b200 import datagen  # noqa: E402

REF = ref_loader.load_reference()
import src.init_helpers as rinit  # noqa: E402
import src.metacell_utils as rmc  # noqa: E402


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


def mip_start_cases():
    """compute_mip_start_pairs(init_method='greedy') (src/init_helpers.py:110-132) on recorded candidate pairs + costs."""
    rec = {}
    cases = []
    for name in ("fig2_direct", "uniform_k8", "sparse_merge"):
        g = np.load(os.path.join(GOLD, f"{name}.npz"), allow_pickle=True)
        pairs, cost = g["w0_pairs"].astype(np.int64), g["w0_cost"].astype(np.float64)
        cases.append((name, pairs, cost))
    # many equal costs: order among ties is the stable sort's (pair index)
    g = np.load(os.path.join(GOLD, "fig2_direct.npz"), allow_pickle=True)
    cases.append(("fig2_rounded", g["w0_pairs"].astype(np.int64), np.round(g["w0_cost"].astype(np.float64) / 20.0)))
    names = []
    for name, pairs, cost in cases:
        na, nr = int(pairs[:, 0].max()) + 1, int(pairs[:, 1].max()) + 1
        rng = np.random.default_rng(len(pairs))
        sizes = rng.integers(1, 4, na).astype(np.float64)
        row_min = np.full(na, np.inf)
        np.minimum.at(row_min, pairs[:, 0], cost)
        for tag, pen in (("p100", 100.0), ("plow", float(np.median(row_min) / 1.5) + 1e-9)):   # plow: about half of the rows prefer no match
            with quiet():
                chosen, unmatched = rinit.compute_mip_start_pairs(valid_pairs=[tuple(p) for p in pairs.tolist()], costs=cost.tolist(), n_aligned=na,
                                                                  n_ref=nr, aligned_sizes=sizes, no_match_penalty=pen, max_matches=1,
                                                                  init_method="greedy", verbose=False)
            key = f"{name}_{tag}"
            names.append(key)
            rec[f"{key}__pairs"], rec[f"{key}__cost"], rec[f"{key}__sizes"], rec[f"{key}__penalty"] = pairs, cost, sizes, np.float64(pen)
            rec[f"{key}__chosen"] = np.asarray(chosen, dtype=np.int64).reshape(-1, 3)
            rec[f"{key}__unmatched"] = np.asarray(sorted(unmatched), dtype=np.int64)
            print(key, len(pairs), "pairs ->", len(chosen), "chosen,", len(unmatched), "unmatched")
    rec["cases"] = np.asarray(names)
    np.savez_compressed(os.path.join(HERE, "mip_start.npz"), **rec)


def collapse_cases():
    """greedy_triangle_collapse with real collapsing (src/metacell_utils.py:160-561)."""
    rec, names = {}, []
    for name, tiles, seed, ms, r_max, ang in (("t2_ms3", 2, 21, 3, 1.5, 10), ("t3_ms10", 3, 22, 10, 2.0, 15), ("t1_ms5_noangle", 1, 23, 5, 1.2, None)):
        ref, qry, ct = datagen.make_section_pair(n_tiles=tiles, seed=seed)
        with quiet():
            mc = rmc.greedy_triangle_collapse(qry, max_metacell_size=ms, r_max=r_max, min_angle_deg=ang, return_object=True)
        mdf = mc.metacell_df
        names.append(name)
        rec[f"{name}__params"] = np.asarray([tiles, seed, ms, r_max, -1.0 if ang is None else ang], dtype=np.float64)
        rec[f"{name}__xy"] = mdf[["X", "Y"]].to_numpy(np.float64)
        rec[f"{name}__size"] = mdf["size"].to_numpy(np.int64)
        rec[f"{name}__type"] = np.asarray(mdf["cell_type"].astype(str).tolist(), dtype="U")
        rec[f"{name}__prob"] = mdf[ct].to_numpy(np.float64)
        rec[f"{name}__members_flat"] = np.asarray([m for ms_ in mdf["members"] for m in ms_], dtype=np.int64)
        rec[f"{name}__members_ptr"] = np.r_[0, np.cumsum([len(m) for m in mdf["members"]])].astype(np.int64)
        rec[f"{name}__metacell_id"] = mdf["metacell_id"].to_numpy(np.int64)
        rec[f"{name}__delaunay"] = np.asarray(mc.metacell_delaunay, dtype=np.int64).reshape(-1, 3)
        rec[f"{name}__original_delaunay"] = np.asarray(mc.original_delaunay, dtype=np.int64).reshape(-1, 3)
        print(name, len(qry), "cells ->", len(mdf), "metacells, max size", int(mdf["size"].max()))
    rec["cases"] = np.asarray(names)
    np.savez_compressed(os.path.join(HERE, "collapse.npz"), **rec)


def run_same_start_cases():
    """The whole reference pipeline with gurobi_params['init_method']='greedy' (src/same.py:1199-1215): the `.Start` values it
    leaves on x[...] and no_match[...] of every window's model, recorded from the fake gurobipy."""
    import tempfile
    from tests.golden import gen_golden as GG          # installs the recording fake + incumbent rule used by the main fixtures
    from tests.util import golden_frame, golden_params, load_golden
    rec, names = {}, []
    for case, sliding in (("fig2_direct", False), ("tiles4_sliding", True)):
        g = load_golden(case)
        ref_df, al_df = golden_frame(g, "ref"), golden_frame(g, "aligned")
        ct = [str(c) for c in g["commonCT"]]
        optim, gurobi = golden_params(g, "optim"), golden_params(g, "gurobi")
        gurobi["init_method"] = "greedy"
        optim["no_match_penalty"] = 25.0               # low enough that some rows prefer to stay unmatched
        ref_loader.MODELS.clear()
        ref_loader.INCUMBENT_FN = GG.make_incumbent_fn(int(g["seed"]))
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as td:
            os.chdir(td)
            try:
                with quiet():
                    if sliding:
                        REF.sliding_window_matching(ref_df, al_df, commonCT=ct, outprefix=os.path.join(td, "out"), optim_params=dict(optim),
                                                    gurobi_params=dict(gurobi))
                    else:
                        REF.run_same(ref_df, al_df, ct, outprefix=None, optim_params=dict(optim), gurobi_params=dict(gurobi))
            finally:
                os.chdir(cwd)
        names.append(case)
        rec[f"{case}__n_models"] = np.int64(len(ref_loader.MODELS))
        for w, m in enumerate(ref_loader.MODELS):
            xs = [v for v in m.vars if v.VarName.startswith("x[")]
            nm = [v for v in m.vars if v.VarName.startswith("no_match[")]
            rec[f"{case}__w{w}_start_x"] = np.asarray([np.nan if v.Start is None else v.Start for v in xs], dtype=np.float64)
            rec[f"{case}__w{w}_start_no_match"] = np.asarray([np.nan if v.Start is None else v.Start for v in nm], dtype=np.float64)
            print(case, "window", w, int(np.nansum(rec[f"{case}__w{w}_start_x"])), "matches,", int(np.nansum(rec[f"{case}__w{w}_start_no_match"])), "unmatched")
    rec["cases"] = np.asarray(names)
    rec["no_match_penalty"] = np.float64(25.0)
    np.savez_compressed(os.path.join(HERE, "run_same_start.npz"), **rec)


def unpack_cases():
    """unpack_metacell_matches (src/metacell_utils.py:564-766): both strategies, metacells on the aligned side only and on both sides."""
    import pandas as pd
    rec, names = {}, []
    ref, qry, ct = datagen.make_section_pair(n_tiles=2, seed=31)
    with quiet():
        mc_a = rmc.greedy_triangle_collapse(qry, max_metacell_size=4, r_max=1.5, min_angle_deg=10, return_object=True)
        mc_r = rmc.greedy_triangle_collapse(ref, max_metacell_size=3, r_max=1.5, min_angle_deg=10, return_object=True)
    rng = np.random.default_rng(31)
    n = min(len(mc_a.metacell_df), len(mc_r.metacell_df), 400)
    big_a = np.argsort(-mc_a.metacell_df["size"].to_numpy(), kind="stable")[:n]          # favour merged metacells
    matches = pd.DataFrame({"Aligned_metacell_id": big_a, "Ref_metacell_id": rng.permutation(len(mc_r.metacell_df))[:n]})
    rec["match_a"], rec["match_r"] = matches["Aligned_metacell_id"].to_numpy(np.int64), matches["Ref_metacell_id"].to_numpy(np.int64)
    for name, ref_side, strategy in (("both_distribute", mc_r.metacell_df, "distribute"), ("both_nearest", mc_r.metacell_df, "nearest"),
                                     ("aligned_only_distribute", ref, "distribute"), ("aligned_only_nearest", ref, "nearest")):
        m = matches if ref_side is not ref else pd.DataFrame({"Aligned_metacell_id": big_a, "Ref_metacell_id": rng.permutation(len(ref))[:n]})
        with quiet():
            out = rmc.unpack_metacell_matches(m, mc_a.metacell_df, ref_side, aligned_df=qry, ref_df=ref, strategy=strategy,
                                              aligned_original_idx_col="Cell_Num_Old", ref_original_idx_col="Cell_Num_Old")
        names.append(name)
        rec[f"{name}__match_r"] = m["Ref_metacell_id"].to_numpy(np.int64)
        rec[f"{name}__aligned"] = out["Aligned_cell_id"].to_numpy(np.int64)
        rec[f"{name}__ref"] = out["Ref_cell_id"].to_numpy(np.int64)
        print("unpack", name, len(m), "metacell matches ->", len(out), "cell matches")
    rec["cases"] = np.asarray(names)
    np.savez_compressed(os.path.join(HERE, "unpack.npz"), **rec)


def heart_case():
    """BASELINE configs[2]: the shipped ISS heart serial sections (examples/heart/data, 3,801 / 3,184 spots, K=8) with
    greedy_triangle_collapse metacells (max_metacell_size=10) through sliding_window_matching, parameters of
    examples/heart/run_same.sh:40-133 (MS=10).  Recorded with the same machinery as the main fixtures."""
    import pandas as pd
    from tests.golden import gen_golden as GG
    d = os.path.join(ref_loader.REFERENCE_ROOT, "examples", "heart", "data")
    cts = ['Smooth muscle cells', 'Fibroblast', 'Atrial cardiomyocytes', 'Cardiomyocytes', 'Endothelium', 'Epicardium',
           'Schwan progenitors', 'Ventricular cardiomyocytes']
    frames = []
    for f in ("refAD_valis.csv", "queryAD_valis.csv"):
        df = pd.read_csv(os.path.join(d, f))
        df = df.rename(columns={f"{c}_percentage": c for c in cts})
        df["X"], df["Y"] = df["New_X"].astype(float), df["New_Y"].astype(float)   # the script's spot_x + 75 gives 0 triangles at r_max=50
        df["cell_type"] = df[cts].idxmax(axis=1)
        frames.append(df[["X", "Y", "Cell_Num", "cell_type"] + cts].copy())
    ref, qry = frames
    ms = 10
    optim = dict(window_size=4000, overlap=100, min_cells_per_window=30, max_matches=1, radius=50, knn=8, no_match_penalty=10000,
                 penalty_coeff=100, dist_ct_coeff=1, delaunay_penalty=10, cell_id_col="metacell_id", ref_metacell_match_multiplier=ms,
                 ignore_same_type_triangles=True, lazy_constraints=True, min_angle_deg=15)
    gurobi = dict(mip_gap=0.05, lazy_allowed_flip_fraction=0.05, time_limit=7200)
    mcp = dict(max_metacell_size=ms, r_max=50, min_angle_deg=15, use_alpha_shape=False)
    rec = {}
    rec.update(GG.frame_arrays(ref, cts, "Cell_Num", "ref"))
    rec.update(GG.frame_arrays(qry, cts, "Cell_Num", "aligned"))
    rec["commonCT"], rec["id_col"], rec["seed"] = np.asarray(cts, dtype="U"), np.asarray("Cell_Num"), np.int64(15)
    for k, v in optim.items():
        rec[f"optim_{k}"] = np.asarray(np.nan if v is None else v)
    for k, v in gurobi.items():
        rec[f"gurobi_{k}"] = np.asarray(np.nan if v is None else v)
    for k, v in mcp.items():
        rec[f"mc_{k}"] = np.asarray(v)
    with quiet():
        mc_al = rmc.greedy_triangle_collapse(qry, cell_type_col="cell_type", original_idx_col="Cell_Num", return_object=True, **mcp)
        mc_rf = rmc.greedy_triangle_collapse(ref, cell_type_col="cell_type", original_idx_col="Cell_Num", return_object=True, **mcp)
    for tag, mc in (("mca", mc_al), ("mcr", mc_rf)):
        mdf = mc.metacell_df
        rec[f"{tag}_xy"], rec[f"{tag}_size"] = mdf[["X", "Y"]].to_numpy(np.float64), mdf["size"].to_numpy(np.int64)
        rec[f"{tag}_prob"] = mdf[cts].to_numpy(np.float64)
        rec[f"{tag}_type"] = np.asarray(mdf["cell_type"].astype(str).tolist(), dtype="U")
        rec[f"{tag}_members_flat"] = np.asarray([m for ms_ in mdf["members"] for m in ms_], dtype=np.int64)
        rec[f"{tag}_delaunay"] = np.asarray(mc.metacell_delaunay, dtype=np.int64).reshape(-1, 3)
    print("heart: metacells", len(mc_al.metacell_df), "/", len(mc_rf.metacell_df), "of", len(qry), "/", len(ref), flush=True)
    ref_loader.MODELS.clear()
    ref_loader.INCUMBENT_FN = GG.make_incumbent_fn(15)
    import tempfile
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            with quiet():
                matches = REF.sliding_window_matching(mc_rf, mc_al, commonCT=list(cts), outprefix=os.path.join(td, "out"), optim_params=dict(optim),
                                                      gurobi_params=dict(gurobi))
        finally:
            os.chdir(cwd)
    rec["n_models"] = np.int64(len(ref_loader.MODELS))
    for w, m in enumerate(ref_loader.MODELS):
        for k, v in GG.record_model(m, cts).items():
            rec[f"w{w}_{k}"] = v
    rec["matches_columns"] = np.asarray(list(matches.columns), dtype="U")
    for col in matches.columns:
        v = matches[col].to_numpy()
        rec[f"matches__{col}"] = np.asarray(v.tolist(), dtype="U") if v.dtype.kind in "OUT" else v
    np.savez_compressed(os.path.join(HERE, "heart_mc10.npz"), **rec)
    print("heart:", len(ref_loader.MODELS), "window model(s),", len(matches), "matches,", sum(len(m.lazy) for m in ref_loader.MODELS), "cuts")


def tongue_case():
    """The shipped tongue sections (examples/tongue/data: 3,608 MERFISH reference / 4,671 protein query cells, K=5, fractional
    probabilities x100, an int64 id column on one side and UUID strings on the other) through greedy_triangle_collapse(MS=1) +
    sliding_window_matching with the parameters of examples/tongue/run_same.sh:22-47.  Real cross-modality data: costs with
    non-integer probabilities, several windows with 300-unit overlaps."""
    import pandas as pd
    import shutil
    from tests.golden import gen_golden as GG
    d = os.path.join(ref_loader.REFERENCE_ROOT, "examples", "tongue", "data")
    cts = ['Endothelial cells', 'Epithelial cells', 'Fibroblasts', 'Lymphoid cells', 'Myeloid cells']
    frames = []
    for f in ("mer_df.csv", "prot_df.csv"):
        df = pd.read_csv(os.path.join(d, f), index_col=0)
        df["X"], df["Y"] = df["transformed_x"], df["transformed_y"]
        df[cts] = df[cts] * 100
        df["cell_type"] = df[cts].idxmax(axis=1)
        frames.append(df[["X", "Y", "Cell_Num", "cell_type"] + cts].copy())
    ref, qry = frames
    optim = dict(window_size=4000, overlap=300, min_cells_per_window=30, max_matches=1, radius=300, knn=8, no_match_penalty=10000,
                 penalty_coeff=100, dist_ct_coeff=1, delaunay_penalty=10, cell_id_col="metacell_id", ref_metacell_match_multiplier=1,
                 lazy_constraints=True, min_angle_deg=15)
    gurobi = dict(mip_gap=0.05, lazy_allowed_flip_fraction=0.05)
    GG.run_case("tongue", ref, qry, cts, optim, gurobi, "Cell_Num", seed=16, use_metacell=True, sliding=True, stage=False,
                mc_params=dict(max_metacell_size=1, r_max=300, min_angle_deg=15, use_alpha_shape=False))
    shutil.move(os.path.join(GOLD, "tongue.npz"), os.path.join(HERE, "tongue.npz"))


if __name__ == "__main__":
    if "--tongue" in sys.argv:
        tongue_case()
        sys.exit(0)
    if "--heart" in sys.argv:
        heart_case()
        sys.exit(0)
    unpack_cases()
    mip_start_cases()
    collapse_cases()
    run_same_start_cases()
