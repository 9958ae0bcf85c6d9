"""Experiment: do two half-batches on two streams overlap usefully?  (two Sections = two copies of the frames, each with its own stream)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from same_b200.device import Section
from same_b200 import _lib as L

class A: tiles=2500
W = bench.make_workload(2500, 0, 1)
rects, grid = bench.window_rects(W, 0, 1)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
mk = lambda st: Section(W["a_xy"], W["r_xy"], W["a_prob"], W["r_prob"], W["a_type"], W["r_type"], device=0, stream=st.cuda_stream)
secA, secB, secC = mk(s1), mk(s2), mk(s1)
half = len(rects) // 2
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
def run(two):
    ts = []
    for it in range(8):
        flush.fill_(1); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if two:
            b1 = secA.batch(rects[:half]); b2 = secB.batch(rects[half:])
            b1.candidates(bench.RADIUS, bench.KNN, False, 1.0); b2.candidates(bench.RADIUS, bench.KNN, False, 1.0)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            b1.close(); b2.close()
        else:
            b = secC.batch(rects); b.candidates(bench.RADIUS, bench.KNN, False, 1.0)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            b.close()
        ts.append((t1 - t0) * 1e3)
    return min(ts[2:]), np.median(ts[2:])
print("one batch, one stream   (wall ms, incl. host):", run(False))
print("two halves, two streams (wall ms, incl. host):", run(True))
print("one batch, one stream   (wall ms, incl. host):", run(False))
