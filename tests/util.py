"""Shared helpers for the parity tests (golden loading, frame rebuilding, incumbent rule)."""
from __future__ import annotations

import glob
import os

import numpy as np
import pandas as pd

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
STAGE_CASES = [c for c in GOLDEN_CASES if c not in ("fig2_priority", "sparse_merge")]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False))


def golden_frame(g, prefix):
    """Rebuild the DataFrame the reference saw (X, Y, id, cell_type, probability columns[, size])."""
    ct = [str(c) for c in g["commonCT"]]
    id_col = str(g["id_col"])
    df = pd.DataFrame({"X": g[f"{prefix}_xy"][:, 0], "Y": g[f"{prefix}_xy"][:, 1]})
    df[id_col] = g[f"{prefix}_id"]
    df["cell_type"] = g[f"{prefix}_type_names"][g[f"{prefix}_type_code"]].astype(object)
    for k, c in enumerate(ct):
        df[c] = g[f"{prefix}_prob"][:, k]
    if f"{prefix}_size" in g:
        df["size"] = g[f"{prefix}_size"]
    df.index = g[f"{prefix}_index"]
    return df


def golden_params(g, kind):
    out = {}
    for k, v in g.items():
        if k.startswith(kind + "_"):
            v = v.item() if v.shape == () else v
            if isinstance(v, float) and np.isnan(v):
                v = None
            if isinstance(v, (np.str_,)):
                v = str(v)
            out[k[len(kind) + 1:]] = v
    return out


def joint_type_codes(g):
    """Type codes comparable across the two frames."""
    names = sorted(set(g["ref_type_names"].tolist()) | set(g["aligned_type_names"].tolist()))
    lut = {n: i for i, n in enumerate(names)}
    ra = np.array([lut[n] for n in g["ref_type_names"][g["ref_type_code"]]], dtype=np.int32)
    al = np.array([lut[n] for n in g["aligned_type_names"][g["aligned_type_code"]]], dtype=np.int32)
    return al, ra


def incumbent_rule(pairs, seed, p_match=0.9):
    """Same rule as tests/golden/gen_golden.py::incumbent_rule (kept textually identical)."""
    pairs = np.asarray(pairs).reshape(-1, 2)
    rng = np.random.default_rng(seed)
    x = np.zeros(len(pairs))
    if len(pairs) == 0:
        return x
    i = pairs[:, 0]
    starts = np.flatnonzero(np.r_[True, i[1:] != i[:-1]])
    counts = np.diff(np.r_[starts, len(i)])
    u = rng.uniform(size=len(starts))
    pick = rng.integers(0, 1 << 30, size=len(starts)) % counts
    sel = starts + pick
    x[sel[u < p_match]] = 1.0
    return x
