"""Dev helper: where the time of CandidateStream goes (host timestamps per phase)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from same_b200 import _lib as L
from same_b200.device import Section, CandidateStream

W = bench.make_workload(2500, 0, 1)
rects, _ = bench.window_rects(W, 0, 1)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
frames = tuple(pin(W[k]) for k in ("a_xy", "r_xy", "a_prob", "r_prob", "a_type", "r_type"))
ARR = [L.KEEP_A, L.KEEP_R, L.ROW_PTR, L.PAIR_J, L.COST]
import ctypes as C
STREAMS = []
for _ in range(2):
    p_ = C.c_void_p(); L.check(L.load().same_stream_create(0, C.byref(p_))); STREAMS.append(p_.value)
KK = [0]
def once(stream=None):
    stream = STREAMS[KK[0] % 2]; KK[0] += 1
    t = [time.perf_counter()]
    s2 = Section(*frames, device=0, stream=stream); t.append(time.perf_counter())
    b = s2.batch(rects); t.append(time.perf_counter())
    b.candidates(bench.RADIUS, bench.KNN, False, 1.0); t.append(time.perf_counter())
    got = b.get_many(ARR, wait=False); t.append(time.perf_counter())
    return s2, b, got, t
def finish(s2, b):
    t0 = time.perf_counter(); b.sync(); t1 = time.perf_counter(); b.close(); t15 = time.perf_counter(); s2.close(); t2 = time.perf_counter()
    return t1 - t0, t15 - t1, t2 - t15
from same_b200 import device as DV
with CandidateStream(bench.RADIUS, bench.KNN, device=0) as cs:
    for rep in range(3):
        torch.cuda.synchronize(); T0 = time.perf_counter(); prev = None; rows = []
        for k in range(16):
            t0 = time.perf_counter()
            h = cs.submit(frames, rects)
            if prev is not None: prev.result()
            prev = h
            rows.append((round((time.perf_counter() - t0) * 1e3, 1), DV.PINNED_ALLOCS[0], L.mempool_stats(0)[0] >> 20))
        prev.result(); torch.cuda.synchronize()
        print("CandidateStream ms/section", round((time.perf_counter() - T0) * 1e3 / 16, 2), "(ms, pinned allocs, pool MiB):", rows)
