"""MIP-start heuristics (mirror of src/init_helpers.py).

`greedy` runs on the GPU: the reference's loop — pairs in ascending (cost, pair index) order, a pair is chosen iff its aligned
row prefers a match and neither endpoint is taken — is the fixed point the `same_greedy_select` kernels compute in parallel
rounds (csrc/greedy.cu).  `hungarian` stays the reference's dense scipy assignment on the host (it is capped at
`init_hungarian_max_n` = 2000 nodes by the reference itself).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Set, Tuple

import numpy as np


def compute_mip_start_pairs(*, valid_pairs: Sequence[Tuple[int, int]], costs: Sequence[float], n_aligned: int, n_ref: int,
                            aligned_sizes: np.ndarray, no_match_penalty: float, max_matches: int, init_method: str,
                            init_big_m: float = 1e9, init_hungarian_max_n: int = 2000, verbose: bool = True,
                            device=None) -> Tuple[List[Tuple[int, int, int]], Set[int]]:
    """-> (chosen (aligned_i, ref_j, var_idx) in selection order, set of unmatched aligned indices)  (src/init_helpers.py:46-177)"""
    method = str(init_method).lower()
    if method not in {"greedy", "hungarian"}:
        raise ValueError(f"Unknown init_method={init_method!r}. Use 'greedy' or 'hungarian'.")
    if method == "hungarian" and max_matches != 1:
        raise ValueError("init_method='hungarian' requires max_matches == 1.")
    if len(valid_pairs) != len(costs):
        raise ValueError("valid_pairs and costs must have the same length.")
    costs_arr = np.asarray(costs, dtype=float)
    pairs = np.asarray(valid_pairs, dtype=np.int64).reshape(-1, 2)
    unmatched_cost = float(no_match_penalty) * np.asarray(aligned_sizes, dtype=float)

    if method == "greedy":
        from .device import greedy_select
        best = np.full(n_aligned, np.inf)
        if len(pairs):
            np.minimum.at(best, pairs[:, 0], costs_arr)
        prefer = best < unmatched_cost                                                  # :118-122
        nodes = np.stack([pairs[:, 0], n_aligned + pairs[:, 1]], axis=1) if len(pairs) else np.zeros((0, 2), np.int64)
        sel = greedy_select(nodes, costs_arr, n_aligned + n_ref, eligible=prefer[pairs[:, 0]] if len(pairs) else None, device=device)
        idx = np.flatnonzero(sel)
        idx = idx[np.argsort(costs_arr[idx], kind="stable")]                            # the order the reference appends in (:124-130)
        chosen = [(int(pairs[k, 0]), int(pairs[k, 1]), int(k)) for k in idx]
        used = np.zeros(n_aligned, bool)
        used[pairs[idx, 0]] = True
        return chosen, set(np.flatnonzero(~used).tolist())

    chosen_pairs: List[Tuple[int, int, int]] = []
    chosen_unmatched: Set[int] = set()
    if (n_aligned + n_ref) > int(init_hungarian_max_n):                                 # :134-141
        if verbose:
            print(f"Skipping Hungarian init: n_aligned+n_ref={n_aligned + n_ref} > init_hungarian_max_n={init_hungarian_max_n}")
        return [], set()
    from scipy.optimize import linear_sum_assignment
    cost_mat = np.full((n_aligned, n_ref + n_aligned), float(init_big_m), dtype=float)  # :150-154
    for idx, (i, j) in enumerate(pairs.tolist()):
        cost_mat[i, j] = float(costs_arr[idx])
    cost_mat[np.arange(n_aligned), n_ref + np.arange(n_aligned)] = unmatched_cost
    row_ind, col_ind = linear_sum_assignment(cost_mat)
    used_ref: Set[int] = set()
    pair_to_var_idx = {(int(i), int(j)): idx for idx, (i, j) in enumerate(pairs.tolist())}
    for i, col in zip(row_ind.tolist(), col_ind.tolist()):
        if col < n_ref and cost_mat[i, col] < float(init_big_m) * 0.5:
            if col in used_ref:
                continue
            used_ref.add(col)
            var_idx = pair_to_var_idx.get((i, col))
            if var_idx is not None:
                chosen_pairs.append((i, col, int(var_idx)))
        else:
            chosen_unmatched.add(i)
    return chosen_pairs, chosen_unmatched


def mip_start_vectors(*, valid_pairs, costs, n_aligned, n_ref, aligned_sizes, no_match_penalty, max_matches, init_method: Optional[str],
                      init_big_m: float = 1e9, init_hungarian_max_n: int = 2000, verbose: bool = True):
    """The `.Start` values apply_mip_start would set, as arrays: (x_start [P], no_match_start [n_aligned]) or None when skipped."""
    if init_method is None:
        return None
    chosen, unmatched = compute_mip_start_pairs(valid_pairs=valid_pairs, costs=costs, n_aligned=n_aligned, n_ref=n_ref, aligned_sizes=aligned_sizes,
                                                no_match_penalty=no_match_penalty, max_matches=max_matches, init_method=init_method,
                                                init_big_m=init_big_m, init_hungarian_max_n=init_hungarian_max_n, verbose=verbose)
    if not chosen and not unmatched:
        return None
    x0 = np.zeros(len(valid_pairs))
    nm = np.zeros(n_aligned)
    nm[list(unmatched)] = 1.0
    for i, _j, k in chosen:
        x0[k] = 1.0
        nm[i] = 0.0
    if verbose:
        print(f"Initialized MIP start ({str(init_method).lower()}): {len(chosen)} matches, {len(unmatched)} unmatched")
    return x0, nm


def apply_mip_start(*, x_vars, no_match_vars, valid_pairs, costs, n_aligned, n_ref, aligned_sizes, no_match_penalty, max_matches,
                    init_method: Optional[str], init_big_m: float = 1e9, init_hungarian_max_n: int = 2000, verbose: bool = True) -> None:
    """Set `.Start` on the solver's variables (src/init_helpers.py:180-246)."""
    v = mip_start_vectors(valid_pairs=valid_pairs, costs=costs, n_aligned=n_aligned, n_ref=n_ref, aligned_sizes=aligned_sizes,
                          no_match_penalty=no_match_penalty, max_matches=max_matches, init_method=init_method, init_big_m=init_big_m,
                          init_hungarian_max_n=init_hungarian_max_n, verbose=verbose)
    if v is None:
        return
    x0, nm = v
    for k in range(len(x0)):
        x_vars[k].Start = float(x0[k])
    for i in range(n_aligned):
        no_match_vars[i].Start = float(nm[i])
