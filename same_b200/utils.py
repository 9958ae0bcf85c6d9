"""`find_knn_within_radius` with the reference's signature (src/utils.py:709-742), computed on the GPU."""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .device import Section


def _knn(aligned_df, ref_df, radius, knn, priority):
    a_xy = np.ascontiguousarray(aligned_df[["X", "Y"]].to_numpy(dtype=np.float64))
    r_xy = np.ascontiguousarray(ref_df[["X", "Y"]].to_numpy(dtype=np.float64))
    a_type = r_type = None
    if priority:
        from .frames import joint_type_codes
        a_type, r_type = joint_type_codes(aligned_df["cell_type"].to_numpy(), ref_df["cell_type"].to_numpy())
    na, nr = len(a_xy), len(r_xy)
    with Section(a_xy, r_xy, np.zeros((na, 0)), np.zeros((nr, 0)), a_type, r_type) as sec, sec.batch() as b:
        b.candidates(radius, knn, priority, 1.0)
        keepA, keepR, pairs = b.get(L.KEEP_A), b.get(L.KEEP_R), b.get(L.PAIRS)
    new_aligned = aligned_df.iloc[keepA].reset_index(drop=True)     # utils.py:736
    new_ref = ref_df.iloc[keepR].reset_index(drop=True)             # utils.py:737
    return new_aligned, new_ref, pairs.astype(np.int64)


def find_knn_within_radius(aligned_df, ref_df, radius=25, knn=5):
    """Per aligned point the <= knn nearest reference points with d^2 <= radius^2; frames compacted to the
    rows that occur and re-indexed; pairs [P, 2] ordered by aligned row then distance (ties: reference index)."""
    return _knn(aligned_df, ref_df, radius, knn, False)
