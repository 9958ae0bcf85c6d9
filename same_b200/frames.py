"""DataFrame contract -> flat arrays for the C-ABI (reference README.md:95-101, src/same.py:934-970)."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import pandas as pd


def joint_type_codes(aligned_types, ref_types):
    """Integer codes of `cell_type`, one code space for both frames (string equality in the reference,
    src/knn_utils.py:55, src/helpers.py:328-330)."""
    a = pd.Series(np.asarray(aligned_types, dtype=object))
    r = pd.Series(np.asarray(ref_types, dtype=object))
    codes, _ = pd.factorize(pd.concat([a, r], ignore_index=True), use_na_sentinel=False)
    return codes[: len(a)].astype(np.int32), codes[len(a):].astype(np.int32)


def frame_arrays(df: pd.DataFrame, commonCT: Sequence[str]):
    xy = np.ascontiguousarray(df[["X", "Y"]].to_numpy(dtype=np.float64))
    prob = np.ascontiguousarray(df[list(commonCT)].to_numpy(dtype=np.float64)) if len(commonCT) else np.zeros((len(df), 0))
    size = df["size"].to_numpy(dtype=np.float64) if "size" in df.columns else None
    return xy, prob, size


def build_section(aligned_df: pd.DataFrame, ref_df: pd.DataFrame, commonCT: Sequence[str], device=None, stream=None):
    """Upload both frames (same_section_create)."""
    from .device import Section
    a_xy, a_prob, a_size = frame_arrays(aligned_df, commonCT)
    r_xy, r_prob, r_size = frame_arrays(ref_df, commonCT)
    if "cell_type" in aligned_df.columns and "cell_type" in ref_df.columns:
        a_type, r_type = joint_type_codes(aligned_df["cell_type"].to_numpy(), ref_df["cell_type"].to_numpy())
    else:
        a_type = r_type = None
    return Section(a_xy, r_xy, a_prob, r_prob, a_type, r_type, a_size, r_size, device=device, stream=stream)
