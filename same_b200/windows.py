"""Sliding-window grid of `sliding_window_matching` (reference src/same.py:481-593), as host logic.

The reference walks the grid sequentially with a small-window merge rule whose side effects
(`i += 1` inside the `j` loop, `window_id` computed after the increments) are part of the
contract.  `enumerate_windows` reproduces that walk but only needs
*cell counts* of candidate rectangles, which the GPU supplies for all rectangles in one launch
(`Section.count_rects`), so the walk itself touches no cell data.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Set, Tuple

import numpy as np


@dataclass
class Window:
    i: int                      # grid column/row AFTER merge increments (used for window_id)
    j: int
    x: int                      # x_windows[i_at_entry], y_windows[j_at_entry]
    y: int
    rect: Tuple[float, float, float, float]          # x_min, x_max, y_min, y_max (half-open, same.py:293-295)
    window_id: int              # len(x_windows) * j + i   (same.py:583)
    central: Tuple[float, float, float, float]       # x_min_filter, x_max_filter, y_min_filter, y_max_filter (same.py:571-582)
    n_ref: int = 0
    n_moving: int = 0
    run: bool = True            # False: too few cells even after merging -> run_same is skipped (same.py:546)


def window_grid(x_min, x_max, y_min, y_max, window_size, overlap):
    """x_windows / y_windows exactly as same.py:483-488 (int() truncation of the bbox)."""
    step = window_size - overlap
    return list(range(int(x_min), int(x_max), step)), list(range(int(y_min), int(y_max), step))


def candidate_rects(x_windows, y_windows, window_size):
    """Every rectangle the merge rule can ask a count for: for each (i, j) the base window, the window
    widened to the right, widened right+down, and widened down only (when there is no right neighbour)."""
    rects, keys = [], {}
    nx, ny = len(x_windows), len(y_windows)
    for i in range(nx):
        for j in range(ny):
            x, y = x_windows[i], y_windows[j]
            xr = x_windows[i + 1] + window_size if i + 1 < nx else None
            yd = y_windows[j + 1] + window_size if j + 1 < ny else None
            for xm in (x + window_size, xr):
                for ym in (y + window_size, yd):
                    if xm is None or ym is None:
                        continue
                    k = (x, xm, y, ym)
                    if k not in keys:
                        keys[k] = len(rects)
                        rects.append(k)
    return np.asarray(rects, dtype=np.float64).reshape(-1, 4), keys


def enumerate_windows(x_windows: Sequence[int], y_windows: Sequence[int], window_size, overlap, min_cells,
                      count: Callable[[Tuple[float, float, float, float]], Tuple[int, int]],
                      bbox_int: Tuple[int, int, int, int],
                      windows_to_process: Optional[Set[Tuple[int, int]]] = None) -> List[Window]:
    """The `while i / while j` walk of same.py:509-593.  `count(rect) -> (n_ref, n_moving)`.
    `bbox_int` = (int(x_min), int(x_max), int(y_min), int(y_max))."""
    ix_min, ix_max, iy_min, iy_max = bbox_int
    nx, ny = len(x_windows), len(y_windows)
    out: List[Window] = []
    i = 0
    while i < nx:
        j = 0
        while j < ny:
            if windows_to_process is not None and (i, j) not in windows_to_process:
                j += 1
                continue
            x, y = x_windows[i], y_windows[j]
            x_lo, x_hi, y_lo, y_hi = x, x + window_size, y, y + window_size
            n_ref, n_mov = count((x_lo, x_hi, y_lo, y_hi))
            if n_ref < min_cells or n_mov < min_cells:
                if i + 1 < nx:
                    x_hi = x_windows[i + 1] + window_size
                    n_ref, n_mov = count((x_lo, x_hi, y_lo, y_hi))
                    if n_ref >= min_cells and n_mov >= min_cells:
                        i += 1
                if (n_ref < min_cells or n_mov < min_cells) and j + 1 < ny:
                    y_hi = y_windows[j + 1] + window_size
                    n_ref, n_mov = count((x_lo, x_hi, y_lo, y_hi))
                    if n_ref >= min_cells and n_mov >= min_cells:
                        j += 1
            ok = n_ref >= min_cells and n_mov >= min_cells
            is_left, is_right = x == ix_min, x_hi >= ix_max
            is_top, is_bottom = y == iy_min, y_hi >= iy_max
            central = (x_lo if is_left else x_lo + overlap / 2, x_hi if is_right else x_hi - overlap / 2,
                       y_lo if is_top else y_lo + overlap / 2, y_hi if is_bottom else y_hi - overlap / 2)
            out.append(Window(i=i, j=j, x=x, y=y, rect=(float(x_lo), float(x_hi), float(y_lo), float(y_hi)),
                              window_id=nx * j + i, central=central, n_ref=int(n_ref), n_moving=int(n_mov), run=ok))
            j += 1
        i += 1
    return out


def numpy_counter(ref_xy: np.ndarray, mov_xy: np.ndarray):
    """Host counter for tests / tiny inputs (the product path counts on the GPU)."""
    def count(rect):
        x0, x1, y0, y1 = rect
        f = lambda p: int(((p[:, 0] >= x0) & (p[:, 0] < x1) & (p[:, 1] >= y0) & (p[:, 1] < y1)).sum())
        return f(ref_xy), f(mov_xy)
    return count


def table_counter(keys: dict, cnt_ref: np.ndarray, cnt_mov: np.ndarray):
    """Counter backed by one batched GPU count over `candidate_rects`."""
    def count(rect):
        k = keys[tuple(rect)]
        return int(cnt_ref[k]), int(cnt_mov[k])
    return count


def shard_windows(n_windows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of the (column-major) window list owned by `rank`."""
    base, rem = divmod(n_windows, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
