"""`find_knn_with_cell_type_priority` with the reference's signature (src/knn_utils.py:5-78), on the GPU."""
from __future__ import annotations

from .utils import _knn


def find_knn_with_cell_type_priority(aligned_df, ref_df, radius, knn=5):
    """KNN candidates, then in aligned-row order: a row whose nearest reference cell has the same `cell_type`
    and was not claimed by an earlier row keeps only that pair.  Returns `(aligned_df, ref_df, pairs)`; the
    frames are the KNN-compacted ones (not re-compacted after pruning, as in the reference); pairs come back as
    a list of `(i, j)` tuples like the reference's `filtered_pairs`."""
    a, r, pairs = _knn(aligned_df, ref_df, radius, knn, True)
    return a, r, [(int(i), int(j)) for i, j in pairs]
