"""One process per GPU: window sharding and the only collectives on the path.

Windows are independent units, so a section shards by contiguous blocks of its window list with NO data-path
collective.  Two small exchanges exist around it, both through `torch.distributed` (NCCL on GPUs, gloo in the CPU
tests):
  * `exchange_halo` — when the *cells* are spatially partitioned across ranks (each rank loaded one strip of the
    section), the cells of the next strip that a rank's last window row reaches into are sent to it once, neighbour to
    neighbour;
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


def exchange_halo(frames: Dict[str, np.ndarray], y_key, y_limit: float, device=None, group=None):
    """Neighbour exchange of border cells: every rank sends the rows with `frames[y_key] < y_limit` (the band of its strip that
    the windows of the PREVIOUS strip reach into) to rank-1 and receives rank+1's band (empty for the last rank).  `y_key` is
    the name of a single-column frame or `(name, column)` to select on one column of a wider frame (e.g. `("xy", 1)`).

    `frames` maps names to arrays sharing their first dimension, in their native dtypes (float64 coordinates and probabilities,
    int32 type codes ...): the selected rows are packed column by column into one byte buffer, so nothing is widened or sent
    twice.  Point-to-point (`batch_isend_irecv`: NCCL over NVLink between GPUs, gloo in the CPU tests): the traffic is the band
    itself, once, whatever the number of ranks.  With `device` set the packing, the transfer and the unpacking run on that GPU
    and the host sees one upload of the band and one download of the received rows.
    -> ({name: received rows}, {"bytes": bytes received, "rows": rows received})"""
    import torch
    dist = _dist()
    rank, world = rank_world(group)
    names = list(frames)
    y_name, y_col = y_key if isinstance(y_key, tuple) else (y_key, 0)
    n = len(np.asarray(frames[y_name]))
    cols = [np.ascontiguousarray(np.asarray(frames[k]).reshape(n, -1)) for k in names]
    shapes = [(c.shape[1], c.dtype) for c in cols]
    empty = {k: np.zeros((0, w), dtype=dt) for k, (w, dt) in zip(names, shapes)}
    if world == 1:
        return empty, dict(bytes=0, rows=0)
    dev = device if device is not None else "cpu"
    mask = np.asarray(frames[y_name]).reshape(n, -1)[:, y_col] < y_limit
    row_bytes = [w * dt.itemsize for w, dt in shapes]
    n_send = int(mask.sum())
    # one byte buffer, column blocks back to back: [rows of column 0][rows of column 1] ...
    send = torch.empty(max(n_send * sum(row_bytes), 1), dtype=torch.uint8, device=dev)
    o = 0
    for c, rb in zip(cols, row_bytes):
        if n_send:
            blk = torch.from_numpy(np.ascontiguousarray(c[mask]).view(np.uint8).reshape(-1))
            send[o:o + n_send * rb].copy_(blk, non_blocking=True)
        o += n_send * rb
    # row counts travel first (8 bytes to the predecessor)
    cnt_out = torch.tensor([n_send], dtype=torch.int64, device=dev)
    cnt_in = torch.zeros(1, dtype=torch.int64, device=dev)
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, cnt_out, rank - 1, group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.irecv, cnt_in, rank + 1, group))
    for r in (dist.batch_isend_irecv(ops) if ops else []):
        r.wait()
    n_recv = int(cnt_in.item()) if rank < world - 1 else 0
    recv = torch.empty(max(n_recv * sum(row_bytes), 1), dtype=torch.uint8, device=dev)
    ops = []
    if rank > 0 and n_send:
        ops.append(dist.P2POp(dist.isend, send, rank - 1, group))
    if rank < world - 1 and n_recv:
        ops.append(dist.P2POp(dist.irecv, recv, rank + 1, group))
    for r in (dist.batch_isend_irecv(ops) if ops else []):
        r.wait()
    got = recv.cpu().numpy()
    res, o = {}, 0
    for k, (w, dt), rb in zip(names, shapes, row_bytes):
        res[k] = got[o:o + n_recv * rb].view(dt).reshape(n_recv, w).copy() if n_recv else empty[k]
        o += n_recv * rb
    return res, dict(bytes=int(n_recv * sum(row_bytes)), rows=n_recv)


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
