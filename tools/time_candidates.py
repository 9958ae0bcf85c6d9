"""Dev helper: per-kernel times of the candidate stage only (subset + bin + search + compaction + emission) on the bench
workload.  python tools/time_candidates.py [tiles] [reps]   (SAME_B200_TILE_DBG / SAME_B200_BIN_TARGET are honoured)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from same_b200 import _lib as L
from same_b200.device import Section

tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
W = bench.make_workload(tiles, 0, 1)
rects, _ = bench.window_rects(W, 0, 1)
with Section(W["a_xy"], W["r_xy"], W["a_prob"], W["r_prob"], W["a_type"], W["r_type"]) as sec:
    for it in range(reps + 2):
        if it == 2:
            L.profile_enable(True)
        with sec.batch(rects) as b:
            b.candidates(bench.RADIUS, bench.KNN, False, 1.0)
            b.sync()
            n_pairs = b.offsets(L.PAIRS)[-1]
    rep = L.profile_report()
    L.profile_enable(False)
tot = 0.0
for name, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:44s} {ms / n * 1000:8.1f} us x{n / reps:.0f}")
    tot += ms / reps
print(f"pairs {n_pairs}  sum of kernels per pass {tot * 1000:.1f} us  tag={os.environ.get('SAME_B200_TILE_DBG', '0')}")
