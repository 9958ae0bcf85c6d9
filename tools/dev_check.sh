# dev helper: GPU parity tests, then a short bench summary
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 5 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_dev.json 2>gpurun_out/bench_dev.err || tail -5 gpurun_out/bench_dev.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_dev.json").read().strip().splitlines()[-1])
k=d["roofline_kernels"]
print("value %.2f Gp/s  tri %.2f G/s  full %.3f ms  launches/step %d" % (d["value"]/1e9, d["triangle_checks"]["value"]/1e9, d["full_pass_ms"], d["gpu_launches"]/d["steps"]))
print({a:round(b,3) for a,b in d["stage_ms"].items()})
if "e2e" in d: print("e2e ms", round(d["e2e"]["ms_per_step"],2))
for n,v in list(k.items())[:14]: print("  %-45s %5.1f x %.4f ms" % (n, v["launches_per_step"], v["avg_ms"]))
PY
