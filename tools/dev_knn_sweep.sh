python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for t in ${TARGETS:-12}; do
SAME_B200_BIN_TARGET=$t python bench.py --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_knn_$t.json 2>gpurun_out/bench_knn_$t.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_knn_$t.json").read().strip().splitlines()[-1])
k=d["roofline_kernels"]
print($t, "cand_ms", round(d["stage_ms"]["candidates"],3), "knn", round(k["k_knn<8>"]["avg_ms"],4), "full", round(d["full_pass_ms"],3))
PY
done
