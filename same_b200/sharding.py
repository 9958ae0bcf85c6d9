"""One process per GPU: window sharding and the only collectives on the path.

Windows are independent units, so a section shards by contiguous blocks of its window list with NO data-path
collective.  Two small exchanges exist around it, both through `torch.distributed` (NCCL on GPUs, gloo in the CPU
tests):
  * `exchange_halo` — when the *cells* are spatially partitioned across ranks (each rank loaded one strip of the
    section), the cells of the next strip that a rank's last window row reaches into are all-gathered once;
  * `gather_matches` — per-rank result frames are gathered in window order.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import pandas as pd


def _dist():
    import torch.distributed as dist
    return dist


def rank_world(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def exchange_halo(frames: Dict[str, np.ndarray], y_key: str, y_limit: float, device=None, group=None):
    """All-gather the rows of every rank whose `frames[y_key]` (a column vector) is < that rank's `y_limit`, and return
    for THIS rank the rows contributed by rank+1 (empty for the last rank) as a dict of arrays.

    `frames` maps names to [N, c] float64 arrays sharing N.  Padded to the largest contribution so a plain
    `all_gather` works on both NCCL and gloo; total traffic is a few MB per rank (border strips only)."""
    import torch
    dist = _dist()
    rank, world = rank_world(group)
    names = list(frames)
    cols = [np.asarray(frames[n], dtype=np.float64).reshape(len(frames[y_key]), -1) for n in names]
    widths = [c.shape[1] for c in cols]
    mask = np.asarray(frames[y_key]).reshape(-1) < y_limit
    pack = np.concatenate([c[mask] for c in cols], axis=1) if cols else np.zeros((0, 0))
    if world == 1:
        return {n: np.zeros((0, w)) for n, w in zip(names, widths)}, dict(bytes=0, rows=0)
    dev = device if device is not None else "cpu"
    cnt = torch.tensor([len(pack)], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    cnts = [int(c.item()) for c in cnts]
    mx = max(max(cnts), 1)
    buf = torch.zeros((mx, pack.shape[1]), dtype=torch.float64, device=dev)
    if len(pack):
        buf[: len(pack)] = torch.from_numpy(np.ascontiguousarray(pack)).to(dev)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    nxt = rank + 1
    got = outs[nxt][: cnts[nxt]].cpu().numpy() if nxt < world else np.zeros((0, pack.shape[1]))
    res, c0 = {}, 0
    for n, w in zip(names, widths):
        res[n] = got[:, c0:c0 + w]
        c0 += w
    return res, dict(bytes=int(buf.numel() * 8 * world), rows=int(len(got)))


def gather_matches(local: pd.DataFrame, group=None) -> Optional[pd.DataFrame]:
    """Concatenate the per-rank match frames in rank order (= window order, shards are contiguous) on every rank."""
    dist = _dist()
    rank, world = rank_world(group)
    if world == 1:
        return local
    parts: List[Optional[pd.DataFrame]] = [None] * world
    dist.all_gather_object(parts, local, group=group)
    parts = [p for p in parts if p is not None and len(p)]
    return pd.concat(parts, ignore_index=True) if parts else pd.DataFrame()


def distributed_sliding_window_matching(ref, moving, commonCT=None, group=None, **kw):
    """`sliding_window_matching` across the ranks of `group`: every rank holds (a replica of) both frames — at
    <= 80 MB per million cells replication is cheaper than partitioning — runs its contiguous block of
    windows on its own GPU and the result frames are gathered.  Parity: identical to the single-process result."""
    from .same import sliding_window_matching
    rank, world = rank_world(group)
    local = sliding_window_matching(ref, moving, commonCT=commonCT, window_shard=(rank, world), **kw)
    return gather_matches(local, group=group)
