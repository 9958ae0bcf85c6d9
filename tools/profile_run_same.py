"""Dev helper: cProfile of the public run_same (IncumbentBackend) on a configs[1]-sized section: python tools/profile_run_same.py [tiles]"""
import cProfile, contextlib, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 25
bench.run_same_wall(tiles)     # warm
t0 = time.perf_counter(); r = bench.run_same_wall(tiles); print("run_same_wall", r["seconds"], "s; total for 3 reps", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); bench.run_same_wall(tiles); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
