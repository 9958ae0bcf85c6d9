"""Timing of greedy_triangle_collapse on a LUAD-shape frame (BASELINE configs[4]: ~94 K cells over 13,000^2 units, K=5, MS=3)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
import same_b200
rng = np.random.default_rng(4)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 94000
xy = rng.uniform(0, 13000, (n, 2)); prob = rng.dirichlet(np.full(5, 0.3), n) * 100
ct = ["t%d" % k for k in range(5)]
df = pd.DataFrame({"X": xy[:, 0], "Y": xy[:, 1], "Cell_Num_Old": np.arange(n)})
for k, c in enumerate(ct):
    df[c] = prob[:, k]
df["cell_type"] = np.asarray(ct)[prob.argmax(1)]
t0 = time.time()
mc = same_b200.greedy_triangle_collapse(df, max_metacell_size=3, r_max=250, min_angle_deg=15, return_object=True)
print("LUAD-shape collapse: %d cells -> %d metacells in %.1f s" % (n, len(mc.metacell_df), time.time() - t0))
