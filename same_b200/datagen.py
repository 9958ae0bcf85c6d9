"""Scaled synthetic sections for tests and benchmarks.

A from-scratch generator in the spirit of the reference's 4-quadrant benchmark
(`src/synthetic_datagen.py:530-643`, semantics summarised in SURVEY.md App. B):
one *tile* is four 10x10 jittered grids on 5x5-unit squares plus an 11-cell ring
(411 reference cells); classes follow a 0.6-unit checkerboard; the query section
is the reference pushed through a smooth deformation with ~9.5 % of the cells
dropped (372 query cells per tile) and independently re-drawn soft one-hot
probabilities (0.85-0.95 on the true class, x100 like the shipped CSVs).
Tiles repeat on a 12.5-unit pitch; `scale` stretches coordinates to tissue
units.  Everything is vectorised numpy and seeded (`np.random.default_rng`), so
a 2,500-tile (~1 M cell) section generates in about a second.

This is NOT the reference generator and makes no attempt to be bit-identical to
it; the shipped Fig-2 CSVs stay the tile-0 fixture (tests/golden/).
"""
from __future__ import annotations

import numpy as np
import pandas as pd

TILE_PITCH = 12.5
_QUAD_ORIGINS = np.array([[1.0, 7.25], [7.25, 7.25], [7.25, 1.0], [1.0, 1.0]])


def _tile_ref(rng: np.random.Generator, n_tiles: int):
    """XY of the reference cells of `n_tiles` tiles in tile-local units -> (n_tiles, 411, 2)."""
    g = (np.arange(10) + 0.5) * 0.5
    gx, gy = np.meshgrid(g, g, indexing="ij")
    grid = np.stack([gx.ravel(), gy.ravel()], axis=1)  # (100, 2) inside a 5x5 square
    quads = grid[None, :, :] + _QUAD_ORIGINS[:, None, :]  # (4, 100, 2)
    base = quads.reshape(1, 400, 2) + rng.normal(0.0, 0.08, size=(n_tiles, 400, 2))
    ang = (np.arange(11) / 11.0) * 2 * np.pi
    ring = np.stack([6.625 + 0.45 * np.cos(ang), 6.625 + 0.45 * np.sin(ang)], axis=1)
    ring = ring[None] + rng.normal(0.0, 0.03, size=(n_tiles, 11, 2))
    return np.concatenate([base, ring], axis=1)


def _soft_one_hot(rng, cls, k):
    n = len(cls)
    hot = rng.uniform(0.85, 0.95, size=n)
    rest = rng.uniform(0.0, 1.0, size=(n, k))
    rest[np.arange(n), cls] = 0.0
    s = rest.sum(axis=1, keepdims=True)
    s[s == 0] = 1.0
    p = rest / s * (1.0 - hot)[:, None]
    p[np.arange(n), cls] = hot
    return p * 100.0


def make_section_pair(n_tiles: int = 1, n_types: int = 3, seed: int = 0, scale: float = 1.0,
                      tiles_per_row: int | None = None, drop_frac: float = 0.095,
                      id_col: str = "Cell_Num_Old"):
    """Return `(ref_df, query_df, commonCT)` obeying the reference's DataFrame contract
    (`README.md:95-101`): `X`, `Y`, `<id_col>`, `cell_type`, one probability column per type."""
    rng = np.random.default_rng(seed)
    if tiles_per_row is None:
        tiles_per_row = int(np.ceil(np.sqrt(n_tiles)))
    t = np.arange(n_tiles)
    origin = np.stack([(t % tiles_per_row) * TILE_PITCH, (t // tiles_per_row) * TILE_PITCH], axis=1)
    ref_xy = (_tile_ref(rng, n_tiles) + origin[:, None, :]).reshape(-1, 2)
    cls = (np.floor(ref_xy[:, 0] / 0.6) + np.floor(ref_xy[:, 1] / 0.6)).astype(np.int64) % n_types
    names = [f"c{i + 1}" for i in range(n_types)]

    # query = smooth deformation of ref + small jitter, a fraction dropped, shuffled row order
    span = max(1.0, tiles_per_row * TILE_PITCH)
    ph = rng.uniform(0, 2 * np.pi, size=4)
    dx = 0.18 * np.sin(2 * np.pi * ref_xy[:, 1] / 6.1 + ph[0]) + 0.10 * np.sin(2 * np.pi * ref_xy[:, 0] / span + ph[1])
    dy = 0.18 * np.cos(2 * np.pi * ref_xy[:, 0] / 5.3 + ph[2]) + 0.10 * np.sin(2 * np.pi * ref_xy[:, 1] / span + ph[3])
    q_xy = ref_xy + np.stack([dx, dy], axis=1) + rng.normal(0.0, 0.04, size=ref_xy.shape)
    keep = rng.uniform(size=len(q_xy)) >= drop_frac
    q_xy, q_cls = q_xy[keep], cls[keep]
    perm = rng.permutation(len(q_xy))
    q_xy, q_cls = q_xy[perm], q_cls[perm]

    def frame(xy, c):
        df = pd.DataFrame({"X": xy[:, 0] * scale, "Y": xy[:, 1] * scale})
        df[id_col] = np.arange(len(df))
        df["cell_type"] = np.asarray(names, dtype=object)[c]
        p = _soft_one_hot(rng, c, n_types)
        for i, nm in enumerate(names):
            df[nm] = p[:, i]
        return df

    return frame(ref_xy, cls), frame(q_xy, q_cls), names


def make_uniform_pair(n_ref: int, n_query: int, extent: float, n_types: int = 5, seed: int = 4,
                      alpha: float = 0.3, id_col: str = "Cell_Num_Old"):
    """LUAD-shaped section (BASELINE.json config 5): uniform-jittered cells over `extent`^2,
    Dirichlet(alpha) x 100 probabilities, `cell_type = argmax`."""
    rng = np.random.default_rng(seed)
    names = [f"t{i}" for i in range(n_types)]

    def frame(n):
        xy = rng.uniform(0, extent, size=(n, 2))
        p = rng.dirichlet(np.full(n_types, alpha), size=n) * 100.0
        df = pd.DataFrame({"X": xy[:, 0], "Y": xy[:, 1]})
        df[id_col] = np.arange(n)
        df["cell_type"] = np.asarray(names, dtype=object)[p.argmax(axis=1)]
        for i, nm in enumerate(names):
            df[nm] = p[:, i]
        return df

    return frame(n_ref), frame(n_query), names
