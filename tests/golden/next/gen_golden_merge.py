"""Golden record of the reference's merge_window_matches_unique_ref (src/helpers.py:692-815).

Its result depends on PYTHONHASHSEED: networkx's Hopcroft-Karp walks `set(align_nodes)` — a set of STRINGS — and string hashes are
salted per process.  So the record is made, and compared, in a subprocess with PYTHONHASHSEED=0:

    PYTHONHASHSEED=0 python tests/golden/next/gen_golden_merge.py     (build container only: imports /root/reference)

Input: three overlapping "windows" of matches with conflicting reference ids, duplicate (aligned, ref) pairs across windows with
different filtered_violation / window_id, NaN violations.  Output: tests/golden/next/merge.npz (input frame + the reference's rows)."""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)


def make_input(seed=3):
    rng = np.random.default_rng(seed)
    frames = []
    for w in range(3):
        n = 60
        a = rng.choice(np.arange(40 * w, 40 * w + 90), n, replace=False)
        r = rng.integers(30 * w, 30 * w + 70, n)
        fv = rng.choice([True, False, np.nan], n, p=[0.2, 0.7, 0.1]).astype(object)
        frames.append(pd.DataFrame({"window_id": w, "Aligned_Cell_Num_Old": a, "Ref_Cell_Num_Old": r, "X": rng.uniform(0, 1, n),
                                    "Y": rng.uniform(0, 1, n), "filtered_violation": fv}))
    # the same (aligned, ref) pair reported by two windows
    dup = frames[0].iloc[:10].copy()
    dup["window_id"] = 2
    dup["filtered_violation"] = False
    frames[2] = pd.concat([frames[2], dup], ignore_index=True)
    return frames


if __name__ == "__main__":
    assert os.environ.get("PYTHONHASHSEED") == "0", "run with PYTHONHASHSEED=0"
    from oracle import ref_loader
    ref = ref_loader.load_reference()
    frames = make_input()
    out = ref.merge_window_matches_unique_ref([f.copy() for f in frames], cell_id_col="Cell_Num_Old")
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "merge.npz"),
                        out_window=out["window_id"].to_numpy(np.int64), out_aligned=out["Aligned_Cell_Num_Old"].to_numpy(np.int64),
                        out_ref=out["Ref_Cell_Num_Old"].to_numpy(np.int64), out_x=out["X"].to_numpy(np.float64),
                        out_fv=out["filtered_violation"].to_numpy(bool))
    print("merge.npz:", len(out), "rows")
