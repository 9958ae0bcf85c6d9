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


def exchange_halo(frames, y_key, y_limit: float, device=None, group=None):
    """Neighbour exchange of border cells: every rank sends the rows with `frames[y_key] < y_limit` (the band of its strip that
    the windows of the PREVIOUS strip reach into) to rank-1 and receives rank+1's band (empty for the last rank).  `y_key` is
    the name of a single-column frame or `(name, column)` to select on one column of a wider frame (e.g. `("xy", 1)`).

    `frames` maps names to arrays sharing their first dimension, in their native dtypes (float64 coordinates and probabilities,
    int32 type codes ...): the selected rows are packed column by column into one byte buffer, so nothing is widened or sent
    twice.  Point-to-point (`batch_isend_irecv`: NCCL over NVLink between GPUs, gloo in the CPU tests): the traffic is the band
    itself, once, whatever the number of ranks.

    numpy arrays in -> numpy arrays out (with `device` set they are staged through that GPU).  torch tensors that already live on
    the GPU in -> tensors on the GPU out: selection, packing, transfer and unpacking all run on the device and nothing touches
    the host except the 8-byte row count.
    -> ({name: received rows}, {"bytes": bytes received, "rows": rows received})"""
    import torch
    dist = _dist()
    rank, world = rank_world(group)
    names = list(frames)
    y_name, y_col = y_key if isinstance(y_key, tuple) else (y_key, 0)
    on_device = all(isinstance(v, torch.Tensor) for v in frames.values())
    if on_device:
        n = len(frames[y_name])
        cols = [frames[k].reshape(n, -1).contiguous() for k in names]
        dev = cols[0].device
        shapes = [(c.shape[1], c.dtype) for c in cols]
        row_bytes = [w * dt.itemsize for w, dt in shapes]
        empty = {k: torch.zeros((0, w), dtype=dt, device=dev) for k, (w, dt) in zip(names, shapes)}
    else:
        n = len(np.asarray(frames[y_name]))
        cols = [np.ascontiguousarray(np.asarray(frames[k]).reshape(n, -1)) for k in names]
        shapes = [(c.shape[1], c.dtype) for c in cols]
        row_bytes = [w * dt.itemsize for w, dt in shapes]
        empty = {k: np.zeros((0, w), dtype=dt) for k, (w, dt) in zip(names, shapes)}
        dev = device if device is not None else "cpu"
    if world == 1:
        return empty, dict(bytes=0, rows=0)
    ycol = cols[names.index(y_name)][:, y_col]
    if on_device:
        idx = torch.nonzero(ycol < y_limit).reshape(-1)
        n_send = int(idx.numel())
        blocks = [c.index_select(0, idx).view(torch.uint8).reshape(-1) for c in cols]
    else:
        mask = ycol < y_limit
        n_send = int(mask.sum())
        blocks = [torch.from_numpy(np.ascontiguousarray(c[mask]).view(np.uint8).reshape(-1)).to(dev, non_blocking=True) for c in cols]
    # one byte buffer, column blocks back to back: [rows of column 0][rows of column 1] ...
    send = torch.cat(blocks) if n_send else torch.empty(1, dtype=torch.uint8, device=dev)
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
    res, o = {}, 0
    if on_device:
        for k, (w, dt), rb in zip(names, shapes, row_bytes):
            res[k] = recv[o:o + n_recv * rb].view(dt).reshape(n_recv, w) if n_recv else empty[k]
            o += n_recv * rb
    else:
        got = recv.cpu().numpy()
        for k, (w, dt), rb in zip(names, shapes, row_bytes):
            res[k] = got[o:o + n_recv * rb].view(dt).reshape(n_recv, w).copy() if n_recv else empty[k]
            o += n_recv * rb
    return res, dict(bytes=int(n_recv * sum(row_bytes)), rows=n_recv)


def allgather_rows(array: np.ndarray, device=None, group=None):
    """Every rank holds the same [N, c] host array.  Rank r moves only rows [r*chunk, (r+1)*chunk), chunk = ceil(N / world), to its
    device — 1/world of the bytes over its own PCIe link — and the slices are all-gathered (NCCL over NVLink; gloo on CPU): every rank
    ends with the whole array on its device having uploaded an eighth of it.  -> torch tensor [N, c] (a view of the gather buffer)."""
    import torch
    dist = _dist()
    rank, world = rank_world(group)
    a = np.ascontiguousarray(array)
    a2 = a.reshape(len(a), -1)
    n, c = a2.shape
    dev = device if device is not None else "cpu"
    t_host = torch.from_numpy(a2)
    if world == 1:
        return t_host.to(dev, non_blocking=True)
    chunk = -(-n // world) if n else 1
    lo, hi = min(rank * chunk, n), min((rank + 1) * chunk, n)
    mine = torch.zeros((chunk, c), dtype=t_host.dtype, device=dev)
    if hi > lo:
        mine[: hi - lo].copy_(t_host[lo:hi], non_blocking=True)
    full = torch.empty((world * chunk, c), dtype=t_host.dtype, device=dev)
    dist.all_gather_into_tensor(full, mine, group=group)
    return full[:n]


def section_from_row_shards(frames, device_index=None, group=None):
    """`Section` of frames that every rank holds on its host (the situation of `distributed_sliding_window_matching`), built
    WITHOUT every rank pushing the whole section through PCIe: each array goes up in row shards (`allgather_rows`) and the section
    is created from the gathered device arrays.  frames = (a_xy, r_xy, a_prob, r_prob, a_type, r_type[, a_size, r_size])."""
    import torch
    from .device import Section, default_device
    di = default_device() if device_index is None else device_index
    dev = torch.device("cuda", di)
    a_xy, r_xy, a_prob, r_prob = (np.ascontiguousarray(f, dtype=np.float64) for f in frames[:4])
    rest = list(frames[4:]) + [None] * (8 - len(frames))
    a_type, r_type = (None if v is None else np.ascontiguousarray(v, dtype=np.int32) for v in rest[:2])
    a_size, r_size = (None if v is None else np.ascontiguousarray(v, dtype=np.float64) for v in rest[2:4])
    host = [a_xy.reshape(-1, 2), r_xy.reshape(-1, 2), a_prob.reshape(len(a_xy.reshape(-1, 2)), -1), r_prob.reshape(len(r_xy.reshape(-1, 2)), -1),
            a_type, r_type, a_size, r_size]
    tens = [None if h is None else allgather_rows(h, dev, group) for h in host]
    ptrs = [None if t is None else int(t.data_ptr()) for t in tens]
    cur = torch.cuda.current_stream(dev)
    if not cur.cuda_stream:
        # the legacy default stream cannot be handed over (handle 0 means "create a private stream" to the library): the gathers
        # must have finished before that private stream reads their output
        cur.synchronize()
    return Section.from_pointers(len(host[0]), len(host[1]), host[2].shape[1], *ptrs, device=di, stream=cur.cuda_stream or None, keep=tens)


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
