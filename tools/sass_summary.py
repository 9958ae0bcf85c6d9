"""ptxas resource table + SASS instruction mix of the top kernels -> profiles/<name>.md (runs on the build container, no GPU).
   python tools/sass_summary.py profiles/r2_sass_ptxas_summary.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "same_b200", "csrc", "_obj")
TOP = ["k_knn<8>", "k_emit_pairs<3>", "k_row_table", "k_pack_records", "k_compact_frames", "k_bin_scatter", "k_subset_count", "k_separation", "k_match_rows", "k_tri_classify",
       "k_remap_count", "k_compact_pairs", "k_group_fill", "k_postsolve", "k_tri_tables"]
out = ["# ptxas resources and SASS instruction mix of the hot kernels (round 2)\n",
       "`nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -Xptxas -v`; SASS from `cuobjdump -sass` of the objects in",
       "`same_b200/csrc/_obj/`.  None of these kernels is a dense contraction (SURVEY.md §0.3: L1 cost over K <= 8 columns, 2-D geometry), so",
       "no `UTC*MMA` / `LDTM` appears by design; the memory path is plain `LDG`/`STG` (+ `LDS`/`STS`, `ATOMG`/`RED`), the arithmetic FP64",
       "(`DADD`/`DMUL`/`DSETP`, no `DFMA` except where the reference's BLAS is an fma) and integer select/min-max.\n",
       "| kernel | registers | shared B | stack B | spill st/ld B |", "|---|---:|---:|---:|---:|"]
res = {}
for f in sorted(os.listdir(OBJ)):
    if not f.endswith(".ptxas.txt"):
        continue
    txt = open(os.path.join(OBJ, f)).read()
    for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("same::", "")
        res[short] = (m.group(1), int(m.group(5)), int(m.group(7) or 0), int(m.group(2)), int(m.group(3)), int(m.group(4)), f.replace(".ptxas.txt", ".o"))
for k in TOP:
    if k in res:
        _, regs, smem, stack, st, ld, _ = res[k]
        out.append(f"| `{k}` | {regs} | {smem} | {stack} | {st} / {ld} |")
out.append("\n## SASS instruction mix (static counts, top mnemonics)\n")
for k in TOP[:8]:
    if k not in res:
        continue
    mangled, obj = res[k][0], os.path.join(OBJ, res[k][6])
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, obj], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1).split(".")[0]] += 1
    tot = sum(ops.values())
    out.append(f"* `{k}`: {tot} instructions — " + ", ".join(f"{o} {n}" for o, n in ops.most_common(14)))
open(sys.argv[1], "w").write("\n".join(out) + "\n")
print("\n".join(out))
