"""Thin object wrappers over the C-ABI: `Section` (frames resident in HBM) and `WindowBatch`
(a list of windows processed together).  Arrays come back as numpy; per-window slices are
views into the concatenated result.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


_PINNED_POOL: dict = {}    # size class -> [host pointers ready for reuse]
PINNED_ALLOCS = [0]        # page-locked blocks obtained from the driver so far (diagnostics)


class _PinnedLease:
    """Holds one cudaHostAlloc block; when the last numpy view dies the block goes back to the pool."""

    def __init__(self, nbytes):
        pool = _PINNED_POOL.setdefault(nbytes, [])
        if pool:
            self.ptr = pool.pop()
        else:
            p = C.c_void_p()
            L.check(L.load().same_pinned_alloc(nbytes, C.byref(p)))
            self.ptr = p.value
            PINNED_ALLOCS[0] += 1
        self.nbytes = nbytes

    def __del__(self):
        try:
            pool = _PINNED_POOL.setdefault(self.nbytes, [])
            # generous: cudaFreeHost / cudaHostAlloc synchronise the device and cost milliseconds — a stream of sections keeps
            # three result sets alive (two in flight, one being read), several arrays of which share a size class
            if len(pool) < 32:
                pool.append(self.ptr)
            else:
                L.load().same_pinned_free(C.c_void_p(self.ptr))
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked host memory (1 MiB size classes, recycled through a small pool)."""
    dtype = np.dtype(dtype)
    shape = tuple(int(v) for v in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    count = int(np.prod(shape)) if shape else 1
    cls = max(1 << 20, -(-(count * dtype.itemsize) // (1 << 20)) * (1 << 20))
    lease = _PinnedLease(cls)
    raw = (C.c_ubyte * cls).from_address(lease.ptr)
    raw._lease = lease        # the ctypes array is the base of every view: keeps the lease alive
    return np.frombuffer(raw, dtype=dtype, count=count).reshape(shape)


def default_device() -> int:
    """GPU of this process: SAME_B200_DEVICE if set, else LOCAL_RANK (one process per GPU under torchrun), else 0."""
    import os
    for key in ("SAME_B200_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(key)
        if v is not None and v.strip().lstrip("-").isdigit():
            return int(v)
    return 0


class Section:
    """Both frames of one tissue section pair on the GPU (same_section_create).  `device=None` = `default_device()`."""

    def __init__(self, a_xy, r_xy, a_prob, r_prob, a_type=None, r_type=None, a_size=None, r_size=None,
                 device=None, stream=None):
        lib = L.load()
        if device is None:
            device = default_device()
        self._keep = []
        a_xy, r_xy = _f64(a_xy, (-1, 2)), _f64(r_xy, (-1, 2))
        self.n_aligned, self.n_ref = len(a_xy), len(r_xy)
        def _block(p, n):
            p = _f64(p)
            if p.ndim == 2 and p.shape[0] == n:
                return p
            return p.reshape(n, p.size // n if n else 0)
        a_prob, r_prob = _block(a_prob, self.n_aligned), _block(r_prob, self.n_ref)
        if self.n_aligned == 0:
            a_prob = np.zeros((0, r_prob.shape[1]))
        if self.n_ref == 0:
            r_prob = np.zeros((0, a_prob.shape[1]))
        if a_prob.shape[1] != r_prob.shape[1]:
            raise ValueError("probability blocks of the two frames have different widths")
        self.n_types = a_prob.shape[1]
        conv = lambda v, dt, n: None if v is None else np.ascontiguousarray(v, dtype=dt).reshape(n)
        a_type, r_type = conv(a_type, np.int32, self.n_aligned), conv(r_type, np.int32, self.n_ref)
        a_size, r_size = conv(a_size, np.float64, self.n_aligned), conv(r_size, np.float64, self.n_ref)
        h = C.c_void_p()
        a_prob, r_prob = np.ascontiguousarray(a_prob), np.ascontiguousarray(r_prob)
        L.check(lib.same_section_create(device, C.c_void_p(stream) if stream else None, self.n_aligned, self.n_ref, self.n_types,
                                        L.ptr(a_xy), L.ptr(r_xy), L.ptr(a_prob), L.ptr(r_prob),
                                        L.ptr(a_type), L.ptr(r_type), L.ptr(a_size), L.ptr(r_size), C.byref(h)))
        self._h = h
        self.device = device
        # page-locked inputs are read asynchronously by the auxiliary upload stream (same_section_wait_uploads): the section
        # keeps every array it handed to the library alive
        self._keep = [a_xy, r_xy, a_prob, r_prob, a_type, r_type, a_size, r_size]
        self.h2d_bytes = a_xy.nbytes + r_xy.nbytes + a_prob.nbytes + r_prob.nbytes + sum(
            v.nbytes for v in (a_type, r_type, a_size, r_size) if v is not None)

    @classmethod
    def from_pointers(cls, n_aligned, n_ref, n_types, a_xy, r_xy, a_prob, r_prob, a_type=None, r_type=None, a_size=None, r_size=None,
                      device=None, stream=None, keep=None):
        """Section from raw addresses (host or device memory, `int`; layouts as in `same_section_create`).  `keep` = objects that
        own the memory (kept alive until the uploads are done)."""
        self = cls.__new__(cls)
        if device is None:
            device = default_device()
        self.n_aligned, self.n_ref, self.n_types = int(n_aligned), int(n_ref), int(n_types)
        h = C.c_void_p()
        L.check(L.load().same_section_create(device, C.c_void_p(stream) if stream else None, self.n_aligned, self.n_ref, self.n_types,
                                             L.ptr(a_xy), L.ptr(r_xy), L.ptr(a_prob), L.ptr(r_prob), L.ptr(a_type), L.ptr(r_type),
                                             L.ptr(a_size), L.ptr(r_size), C.byref(h)))
        self._h = h
        self.device = device
        self._keep = [keep]
        self.h2d_bytes = 0
        return self

    @property
    def bbox(self):
        out = np.zeros(4)
        L.check(L.load().same_section_bbox(self._h, L.ptr(out)))
        return out

    def count_rects(self, rects):
        rects = _f64(rects, (-1, 4))
        m = len(rects)
        ca, cr = np.zeros(m, np.int64), np.zeros(m, np.int64)
        L.check(L.load().same_section_count_rects(self._h, m, L.ptr(rects), L.ptr(ca), L.ptr(cr)))
        return ca, cr

    def set_triangles(self, tri_vid, a_vid=None):
        """Precomputed triangulation in vertex-id space (same.py:262-290)."""
        tri = np.ascontiguousarray(tri_vid, dtype=np.int64).reshape(-1, 3)
        vid = None if a_vid is None else np.ascontiguousarray(a_vid, dtype=np.int64).reshape(self.n_aligned)
        L.check(L.load().same_section_set_triangles(self._h, L.ptr(vid), L.ptr(tri), len(tri)))
        self.n_global_triangles = len(tri)

    def batch(self, rects=None):
        return WindowBatch(self, rects)

    def close(self):
        if getattr(self, "_h", None):
            L.check(L.load().same_section_destroy(self._h))
            self._h = None
            self._keep = []

    def wait_uploads(self):
        """Block until every column of both frames is on the device (same_section_wait_uploads)."""
        L.check(L.load().same_section_wait_uploads(self._h))
        self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class CandidateStream:
    """Candidate stage (window subsetting + KNN + pair costs, src/same.py:523-524, src/utils.py:709-742, src/same.py:1182-1189)
    of a SEQUENCE of sections — serial tissue slices, or the strips of one large slice — with the transfers and kernels of
    neighbouring sections overlapped.  Every section runs on one of `depth` CUDA streams (taken in turn) and is driven by its
    own worker thread (the C-ABI calls release the GIL; distinct sections share no mutable state in the library), and its
    results are copied back asynchronously: while section k's pairs and costs cross PCIe towards the host, section k+1's frames
    cross it in the other direction and its kernels run.

        cs = CandidateStream(radius, knn)
        for frames, rects in sections:                # host arrays (ideally page-locked)
            h = cs.submit(frames, rects)              # returns at once
            if prev is not None: out = prev.result()  # {KEEP_A, KEEP_R, ROW_PTR, PAIR_J, COST} + "offsets"
            prev = h

    At most `depth` sections may be outstanding (submitted, `result()` not yet called).  `result()` gives valid_pairs in compact
    form: the aligned index of pair p is the row r with ROW_PTR[r] <= p < ROW_PTR[r+1] (`pairs_from_rows`), the reference index
    is PAIR_J[p].  With `j16=True` PAIR_J comes back as uint16 (two bytes per pair over PCIe instead of four; the index is
    window-local) whenever every window keeps at most 65,536 reference rows, and as int32 otherwise — the key is L.PAIR_J either
    way."""

    ARRAYS = (L.KEEP_A, L.KEEP_R, L.ROW_PTR, L.PAIR_J, L.COST)

    class Handle:
        def __init__(self, owner, future):
            self._owner, self._future = owner, future

        def result(self):
            """Wait for this section's downloads; release its device memory.  -> dict of numpy arrays (page-locked)."""
            try:
                sec, b, arrays = self._future.result()
            finally:
                self._owner._outstanding -= 1
            try:
                b.sync()
                out = dict(arrays)
                out["offsets"] = {w: b.offsets(w).copy() for w in (L.KEEP_A, L.KEEP_R, L.PAIRS)}
            finally:
                b.close()
                sec.close()
            return out

    def __init__(self, radius, knn, priority=False, dist_ct_coeff=1.0, device=None, depth=3, j16=False):
        from concurrent.futures import ThreadPoolExecutor
        self.radius, self.knn, self.priority, self.dist_ct_coeff = float(radius), int(knn), bool(priority), float(dist_ct_coeff)
        self.j16 = bool(j16)
        self.device = default_device() if device is None else device
        self._k = self._outstanding = 0
        # a fixed set of streams, taken in turn: the stream-ordered memory pool then recycles a section's buffers for the
        # section after next without a driver call (a new stream per section cannot reuse memory freed on other streams)
        self._streams = []
        for _ in range(max(1, int(depth))):
            p = C.c_void_p()
            L.check(L.load().same_stream_create(self.device, C.byref(p)))
            self._streams.append(p.value)
        self._pool = ThreadPoolExecutor(max_workers=len(self._streams), thread_name_prefix="same_b200-section")

    def reserve(self, nbytes=None):
        """Give the device memory pool head-room for the overlap of `depth` sections (default: twice what it holds now), so that no
        section of the steady state has to wait for the driver to map new memory — tens of milliseconds when several processes
        share the host.  Call between sections (nothing outstanding), after the first results.  The request is ONE allocation that
        is freed again at once, so the pool ends up holding max(what it held, nbytes)."""
        if nbytes is None:
            nbytes = 2 * L.mempool_stats(self.device)[0]
        L.mempool_reserve(self.device, int(nbytes))

    def close(self):
        if getattr(self, "_pool", None) is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        for p in getattr(self, "_streams", []):
            L.check(L.load().same_stream_destroy(self.device, C.c_void_p(p)))
        self._streams = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _run(self, frames, rects, stream):
        if not self.priority:       # type codes (and sizes) are read by the priority filter / later stages only: they stay on the host
            frames = tuple(frames[:4]) + (None,) * (len(frames) - 4)
        sec = Section(*frames, device=self.device, stream=stream)
        try:
            b = sec.batch(rects)
            try:
                b.candidates(self.radius, self.knn, self.priority, self.dist_ct_coeff)
                if self.j16 and int(np.diff(b.offsets(L.KEEP_R)).max(initial=0)) <= 65536:
                    got = b.get_many(tuple(L.PAIR_J16 if w == L.PAIR_J else w for w in self.ARRAYS), pinned=True, wait=False)
                    got[L.PAIR_J] = got.pop(L.PAIR_J16)
                    return sec, b, got
                return sec, b, b.get_many(self.ARRAYS, pinned=True, wait=False)
            except BaseException:
                b.close()
                raise
        except BaseException:
            sec.close()
            raise

    def submit(self, frames, rects=None):
        """frames = (a_xy, r_xy, a_prob, r_prob, a_type, r_type[, a_size, r_size]); sections take the streams in turn.  The type
        codes are uploaded only when the stream was created with priority=True (nothing else in this stage reads them)."""
        if self._outstanding >= len(self._streams):
            raise RuntimeError(f"CandidateStream: {self._outstanding} sections outstanding; call result() on the oldest first "
                               f"(depth={len(self._streams)})")
        stream = self._streams[self._k % len(self._streams)]
        self._k += 1
        self._outstanding += 1
        return CandidateStream.Handle(self, self._pool.submit(self._run, frames, rects, stream))


def pairs_from_rows(row_ptr, pair_j, row_base=0):
    """valid_pairs [P, 2] (src/utils.py:741) from the compact form ROW_PTR + PAIR_J of one window: `row_ptr` = the window's
    slice of ROW_PTR (nKA + 1 entries, any base), `pair_j` its slice of PAIR_J."""
    rp = np.asarray(row_ptr, dtype=np.int64)
    i = np.repeat(np.arange(len(rp) - 1, dtype=np.int32), np.diff(rp))
    return np.column_stack([i, np.asarray(pair_j, dtype=np.int32)])


def greedy_select(nodes, key, n_nodes, eligible=None, device=None, return_rounds=False):
    """Ordered greedy selection with disjoint endpoints (same_greedy_select): items visited in ascending (key, index) order,
    an item is taken iff it is eligible and none of its endpoints was taken before.  nodes [n, 1..3] int -> bool [n]."""
    nodes = np.ascontiguousarray(nodes, dtype=np.int32)
    if nodes.ndim == 1:
        nodes = nodes.reshape(-1, 1)
    n, degree = nodes.shape
    key = np.ascontiguousarray(key, dtype=np.float64).reshape(n)
    if n and (nodes.min() < 0 or nodes.max() >= n_nodes):
        raise ValueError("endpoint out of range")
    if np.isnan(key).any():
        raise ValueError("keys must not be NaN")
    el = None if eligible is None else np.ascontiguousarray(eligible, dtype=np.uint8).reshape(n)
    sel, used, r = np.zeros(n, np.uint8), np.zeros(int(n_nodes), np.uint8), C.c_int32(0)
    device = default_device() if device is None else device
    L.check(L.load().same_greedy_select(device, n, degree, L.ptr(nodes), L.ptr(key), L.ptr(el), int(n_nodes), L.ptr(sel), L.ptr(used), C.byref(r)))
    return (sel.astype(bool), r.value) if return_rounds else sel.astype(bool)


def collapse_select(xy, type_codes, sizes, tri, max_size, device=None):
    """One collapse iteration of greedy_triangle_collapse on the GPU (same_collapse_select): candidate test, perimeter in the
    reference's arithmetic, ordered disjoint selection.  -> (selected bool [T], perimeter float64 [T])"""
    xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
    n = len(xy)
    tc = np.ascontiguousarray(type_codes, dtype=np.int32).reshape(n)
    sz = np.ascontiguousarray(sizes, dtype=np.float64).reshape(n)
    tri = np.ascontiguousarray(tri, dtype=np.int32).reshape(-1, 3)
    T = len(tri)
    if T and (tri.min() < 0 or tri.max() >= n):
        raise ValueError("triangle vertex out of range")
    sel, per, r = np.zeros(T, np.uint8), np.zeros(T, np.float64), C.c_int32(0)
    device = default_device() if device is None else device
    L.check(L.load().same_collapse_select(device, n, L.ptr(xy), L.ptr(tc), L.ptr(sz), T, L.ptr(tri), float(max_size), L.ptr(sel), L.ptr(per),
                                          C.byref(r)))
    return sel.astype(bool), per


def segment_mean(values, ptr, pos, device=None):
    """out[g] = mean of values[pos[ptr[g]:ptr[g+1]]] per column, in pandas' / numpy's summation order (same_segment_mean)."""
    values = np.ascontiguousarray(values, dtype=np.float64)
    if values.ndim == 1:
        values = values.reshape(-1, 1)
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    G = len(ptr) - 1
    if len(pos) != ptr[-1] or (len(pos) and (pos.min() < 0 or pos.max() >= len(values))):
        raise ValueError("member positions out of range")
    out = np.empty((G, values.shape[1]), dtype=np.float64)
    device = default_device() if device is None else device
    L.check(L.load().same_segment_mean(device, values.shape[0], values.shape[1], L.ptr(values), G, L.ptr(ptr), L.ptr(pos), L.ptr(out)))
    return out


class WindowBatch:
    """A list of windows of one section (same_batch_create).  `rects=None` = the whole section."""

    def __init__(self, section: Section, rects=None):
        lib = L.load()
        self.section = section
        if rects is None:
            self.W, r = 1, None
        else:
            r = _f64(rects, (-1, 4))
            self.W = len(r)
        h = C.c_void_p()
        L.check(lib.same_batch_create(section._h, self.W, L.ptr(r), C.byref(h)))
        self._h = h
        self._off_cache = {}

    # ---- stages ----
    def candidates(self, radius, knn, priority=False, dist_ct_coeff=1.0):
        self._off_cache.clear()
        L.check(L.load().same_batch_candidates(self._h, float(radius), int(knn), int(bool(priority)), float(dist_ct_coeff)))

    def triangles_remap(self):
        self._off_cache.clear()
        L.check(L.load().same_batch_triangles_remap(self._h))

    def triangles_set(self, tri, tri_off):
        self._off_cache.clear()
        tri = np.ascontiguousarray(tri, dtype=np.int32).reshape(-1, 3)
        off = np.ascontiguousarray(tri_off, dtype=np.int64)
        L.check(L.load().same_batch_triangles_set(self._h, L.ptr(tri), L.ptr(off)))

    def tri_classify(self, radius, min_angle_deg, ignore_same_type):
        nb = C.c_int64(0)
        L.check(L.load().same_batch_tri_classify(self._h, float(radius), int(min_angle_deg is not None),
                                                 0.0 if min_angle_deg is None else float(min_angle_deg),
                                                 int(bool(ignore_same_type)), C.byref(nb)))
        return nb.value

    def tri_override(self, idx, cls):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        cls = np.ascontiguousarray(cls, dtype=np.uint8)
        L.check(L.load().same_batch_tri_override(self._h, len(idx), L.ptr(idx), L.ptr(cls)))

    def tri_finalize(self, ignore_same_type, ensure_min=True, remove_unconstrained=False):
        self._off_cache.clear()
        L.check(L.load().same_batch_tri_finalize(self._h, int(bool(ignore_same_type)), int(bool(ensure_min)),
                                                 int(bool(remove_unconstrained))))

    def groups(self, max_matches, multiplier=None):
        self._off_cache.clear()
        L.check(L.load().same_batch_groups(self._h, int(max_matches), -1 if multiplier is None else int(multiplier)))

    def separation(self, x, w_lo=0, w_hi=None, cap=1000):
        """-> (n_viol[nw], n_checked[nw], cuts[nw, cap, 4]); x may be a numpy array or a device address."""
        w_hi = self.W if w_hi is None else w_hi
        nw = w_hi - w_lo
        nv, nc = np.zeros(nw, np.int64), np.zeros(nw, np.int64)
        cuts = pinned_empty((nw, max(cap, 1), 4), np.int32)   # rows beyond min(n_viol, cap) are unspecified
        if not isinstance(x, int):
            x = np.ascontiguousarray(x, dtype=np.float64)
        L.check(L.load().same_batch_separation(self._h, w_lo, w_hi, L.ptr(x), cap, L.ptr(nv), L.ptr(nc), L.ptr(cuts)))
        return nv, nc, cuts

    def postsolve(self, x, w_lo=0, w_hi=None):
        w_hi = self.W if w_hi is None else w_hi
        if not isinstance(x, int):
            x = np.ascontiguousarray(x, dtype=np.float64)
        L.check(L.load().same_batch_postsolve(self._h, w_lo, w_hi, L.ptr(x)))

    def uncertain(self, which, cap=65536):
        """Triangles whose naive orientation sign the static error filter could not certify (same_batch_uncertain):
        which = 0 source signs, 1 = the last separation call.  -> (count, batch-global TRI indices, ascending)."""
        n = C.c_int64(0)
        idx = np.zeros(max(int(cap), 1), np.int32)
        L.check(L.load().same_batch_uncertain(self._h, int(which), int(cap), C.byref(n), L.ptr(idx)))
        m = min(n.value, int(cap), 65536)
        return n.value, np.sort(idx[:m])

    def mip_start(self, no_match_penalty):
        """Greedy MIP start of every window (init_helpers.py:110-132) -> parallel rounds used; results: START_X, START_UNMATCHED."""
        r = C.c_int32(0)
        L.check(L.load().same_batch_mip_start(self._h, float(no_match_penalty), C.byref(r)))
        return r.value

    # ---- results ----
    def offsets(self, what):
        if what not in self._off_cache:
            off = np.zeros(self.W + 1, np.int64)
            L.check(L.load().same_batch_offsets(self._h, what, L.ptr(off)))
            self._off_cache[what] = off
        return self._off_cache[what]

    def length(self, what):
        n = C.c_int64(0)
        L.check(L.load().same_batch_length(self._h, what, C.byref(n)))
        return n.value

    def get(self, what, lo=0, hi=None):
        dt, tail = L.ARRAY_SPEC[what]
        hi = self.length(what) if hi is None else hi
        out = np.empty((hi - lo,) + tail, dtype=dt)
        L.check(L.load().same_batch_get(self._h, what, lo, hi, L.ptr(out)))
        return out

    def get_many(self, whats, pinned=True, wait=True):
        """Fetch several whole arrays with one stream synchronisation -> {what: ndarray}.
        `wait=False` (page-locked destinations only) queues the copies and returns at once: the arrays may be read after
        `sync()`.  That is what lets the next section's upload and kernels overlap this one's download (`CandidateStream`)."""
        if not wait and not pinned:
            raise ValueError("asynchronous downloads need page-locked destinations")
        whats = list(whats)
        n = len(whats)
        outs, lo, hi = [], np.zeros(n, np.int64), np.zeros(n, np.int64)
        for k, what in enumerate(whats):
            dt, tail = L.ARRAY_SPEC[what]
            hi[k] = self.length(what)
            outs.append(pinned_empty((int(hi[k]),) + tail, dt) if pinned else np.empty((int(hi[k]),) + tail, dt))
        dst = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        w = np.asarray(whats, dtype=np.int32)
        fn = L.load().same_batch_get_many if wait else L.load().same_batch_get_many_async
        L.check(fn(self._h, n, L.ptr(w), L.ptr(lo), L.ptr(hi), dst))
        return dict(zip(whats, outs))

    def stat(self, what):
        v = C.c_int64(0)
        L.check(L.load().same_batch_stat(self._h, int(what), C.byref(v)))
        return v.value

    def get_window(self, what, w):
        off = self.offsets(what)
        return self.get(what, int(off[w]), int(off[w + 1]))

    def sync(self):
        L.check(L.load().same_batch_sync(self._h))

    @property
    def stream(self):
        return L.load().same_batch_stream(self._h)

    def close(self):
        if getattr(self, "_h", None):
            L.check(L.load().same_batch_destroy(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- convenience: one window's model arrays in the reference's index spaces ----
    def window_model(self, w):
        """dict of numpy arrays for window w (everything run_same needs to build its model)."""
        p0 = int(self.offsets(L.PAIRS)[w])
        out = dict(keepA=self.get_window(L.KEEP_A, w), keepR=self.get_window(L.KEEP_R, w),
                   pairs=self.get_window(L.PAIRS, w), cost=self.get_window(L.COST, w))
        def stage_missing(e):
            if e.code != L.E_STATE:      # only "stage not run yet" is an expected condition here
                raise e
        try:
            out.update(tri=self.get_window(L.TRI, w), tri_src=self.get_window(L.TRI_SRC, w), weight=self.get_window(L.TRI_WEIGHT, w),
                       sign=self.get_window(L.TRI_SIGN, w), bounds=self.get_window(L.TRI_BOUNDS, w), argv=self.get_window(L.TRI_ARGV, w),
                       unconstrained=self.get_window(L.UNCONSTRAINED, w))
        except L.SameError as e:
            stage_missing(e)
        try:
            goff = self.offsets(L.REF_GROUP_NODE)
            g0, g1 = int(goff[w]), int(goff[w + 1])
            out.update(ref_group_node=self.get(L.REF_GROUP_NODE, g0, g1), ref_group_limit=self.get(L.REF_GROUP_LIMIT, g0, g1),
                       ref_group_ptr=self.get(L.REF_GROUP_PTR, g0, g1 + 1).astype(np.int64) - p0,
                       ref_group_idx=self.get_window(L.REF_GROUP_IDX, w))
        except L.SameError as e:
            stage_missing(e)
        ka = self.offsets(L.KEEP_A)
        out["row_ptr"] = self.get(L.ROW_PTR, int(ka[w]), int(ka[w + 1]) + 1).astype(np.int64) - p0
        return out
