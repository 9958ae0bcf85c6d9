#!/usr/bin/env python
"""Wall-clock of the UNMODIFIED reference's run_same on BASELINE configs[1] (the ~10 k-cell synthetic section), solver excluded.

Build container only (imports /root/reference under oracle/ref_loader.py's stubs; the recording fake gurobipy "solves" by a
seeded incumbent and fires ONE MIPSOL callback, so what is timed is everything around the MIP: KNN, costs, triangulation,
filtering, model construction calls, one lazy separation, post-solve analysis).  Writes profiles/reference_cpu_c2.json.

    python tools/time_reference_c2.py [n_tiles]
"""
import contextlib, io, json, os, sys, tempfile, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import ref_loader
from same_b200 import datagen
from tests.golden import gen_golden as GG      # installs the fake solver's incumbent rule

REF = ref_loader.load_reference()
n_tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 25
ref, qry, ct = datagen.make_section_pair(n_tiles=n_tiles, n_types=3, seed=1)
optim = dict(radius=5 * 0.2, knn=8, max_matches=2, min_angle_deg=5, cell_id_col="Cell_Num_Old", dist_ct_coeff=1, ignore_same_type_triangles=False,
             delaunay_penalty=10, no_match_penalty=10000, penalty_coeff=100, lazy_constraints=True)
gurobi = dict(mip_gap=0.025, lazy_allowed_flip_fraction=0.0, time_limit=7200, mip_focus=2)
ref_loader.MODELS.clear()
ref_loader.INCUMBENT_FN = GG.make_incumbent_fn(1)
cwd = os.getcwd()
with tempfile.TemporaryDirectory() as td:
    os.chdir(td)
    try:
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            matches, var_out = REF.run_same(ref, qry, list(ct), outprefix=None, optim_params=dict(optim), gurobi_params=dict(gurobi))
        dt = time.perf_counter() - t0
    finally:
        os.chdir(cwd)
m = ref_loader.MODELS[0]
out = dict(config=f"BASELINE configs[1]: same_b200.datagen section, {n_tiles} tiles, seed 1 ({len(ref)} ref / {len(qry)} query cells, K=3)",
           optim_params=optim, seconds=dt, pairs=len(m._valid_pairs), triangles=len(m._aligned_delaunay), cuts=len(m.lazy), matches=len(matches),
           cpu=os.cpu_count(), where="build container (no GPU), single Python thread — the reference has no other mode",
           note="solver replaced by the recording fake (seeded incumbent, one MIPSOL callback): the time is the reference's own code around the MIP")
print(json.dumps(out))
json.dump(out, open(os.path.join(ROOT, "profiles", f"reference_cpu_c2_{n_tiles}tiles.json"), "w"), indent=1)
