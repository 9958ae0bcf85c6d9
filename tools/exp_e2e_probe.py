import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from same_b200.device import Section, _PINNED_POOL
from same_b200 import _lib as L
tiles = int(sys.argv[1])
W = bench.make_workload(tiles, 0, 1)
rects, grid = bench.window_rects(W, 0, 1)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
P = {k: pin(W[k]) for k in ("a_xy", "r_xy", "a_prob", "r_prob", "a_type", "r_type")}
for it in range(5):
    t0 = time.perf_counter()
    s2 = Section(P["a_xy"], P["r_xy"], P["a_prob"], P["r_prob"], P["a_type"], P["r_type"], device=0, stream=st.cuda_stream)
    t1 = time.perf_counter()
    b = s2.batch(rects)
    t2 = time.perf_counter()
    b.candidates(bench.RADIUS, bench.KNN, False, 1.0); b.sync()
    t3 = time.perf_counter()
    n = [b.length(w) for w in (L.KEEP_A, L.KEEP_R, L.PAIRS, L.COST, L.ROW_PTR)]
    t4 = time.perf_counter()
    got = b.get_many([L.KEEP_A, L.KEEP_R, L.PAIRS, L.COST, L.ROW_PTR])
    t5 = time.perf_counter()
    b.close(); s2.close(); del got
    t6 = time.perf_counter()
    print("iter %d: section %.2f batch %.2f cand %.2f len %.2f get_many %.2f close %.2f ms  pool %s" % (it, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3, (t6-t5)*1e3, {k: len(v) for k, v in _PINNED_POOL.items()}))
