"""Print the headline numbers and the per-kernel table of a bench.py JSON line: python tools/show_bench.py FILE [N]"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 18
print("value %.3f G pairs/s  ms/step %.4f  tri checks %.3f G/s" % (d["value"] / 1e9, d["ms_per_step"], d.get("triangle_checks", {}).get("value", 0) / 1e9))
print("stage_ms", {k: round(v, 4) for k, v in d.get("stage_ms", {}).items()}, "full_pass_ms", round(d.get("full_pass_ms", 0), 4))
e = d.get("e2e", {})
print("e2e %.3f G pairs/s %.3f ms  full_path %.3f ms" % (e.get("value", 0) / 1e9, e.get("ms_per_step", 0), e.get("full_path", {}).get("ms_per_step", 0)))
r = d.get("roofline", {})
print("roofline", {k: r[k] for k in r if k not in ("note",)})
ks = d.get("roofline_kernels", {})
for k, v in sorted(ks.items(), key=lambda kv: -kv[1]["avg_ms"] * kv[1]["launches_per_step"])[:n]:
    print(f"  {k:42s} {v['avg_ms'] * 1000:8.1f} us x{v['launches_per_step']:.0f}")
for k in d:
    if k.startswith("knn") or k in ("search_stats",):
        print(k, d[k])
