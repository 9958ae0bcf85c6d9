"""ctypes front-end of the CPU oracle (`oracle/same_oracle.c`).  TEST INFRASTRUCTURE ONLY.

Imported by `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs — never by `same_b200/`.  Parity status: pinned against the
reference's own outputs (tests/golden/*.npz, tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsame_oracle.so")
_lib = None

TRI_DROP_RADIUS, TRI_DROP_ANGLE, TRI_SAME_TYPE, TRI_KEEP = 0, 1, 2, 3


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "same_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        for name in ("oracle_knn_compact", "oracle_knn_priority", "oracle_group_pairs", "oracle_remap_triangles",
                     "oracle_tri_select", "oracle_separation", "oracle_subset", "oracle_greedy_select"):
            getattr(_lib, name).restype = C.c_int64
    return _lib


def set_threads(n: int = 0) -> int:
    """OpenMP worker threads of the oracle's loops (n <= 0: just report).  -> threads in effect."""
    return int(lib().oracle_set_threads(C.c_int(int(n))))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def knn_candidates(a_xy, r_xy, radius, knn, brute=False):
    """a1 (utils.py:709-731): per aligned row the <=knn nearest refs within radius.
    Returns (cand_j [Na,knn] int32 padded with -1, cnt [Na] int32)."""
    a_xy, r_xy = _f64(a_xy).reshape(-1, 2), _f64(r_xy).reshape(-1, 2)
    na, nr = len(a_xy), len(r_xy)
    cand = np.full((na, max(knn, 0)), -1, dtype=np.int32)
    cnt = np.zeros(na, dtype=np.int32)
    fn = lib().oracle_knn_brute if brute else lib().oracle_knn_grid
    fn(_p(a_xy), C.c_int64(na), _p(r_xy), C.c_int64(nr), C.c_double(radius), C.c_int(knn), _p(cand), _p(cnt))
    return cand, cnt


def knn_compact(cand, cnt, nr):
    """utils.py:733-742 -> (keepA, keepR, pairs[P,2]) with pairs in the re-indexed space."""
    cand, cnt = _i32(cand), _i32(cnt)
    na, knn = cand.shape
    keepA = np.empty(na, np.int32)
    keepR = np.empty(nr, np.int32)
    pairs = np.empty((max(int(cnt.sum()), 1), 2), np.int32)
    nka, nkr = C.c_int64(0), C.c_int64(0)
    p = lib().oracle_knn_compact(C.c_int64(na), C.c_int64(nr), C.c_int(knn), _p(cand), _p(cnt),
                                 _p(keepA), C.byref(nka), _p(keepR), C.byref(nkr), _p(pairs))
    return keepA[:nka.value].copy(), keepR[:nkr.value].copy(), pairs[:p].copy()


def find_knn_within_radius(a_xy, r_xy, radius, knn, brute=False):
    cand, cnt = knn_candidates(a_xy, r_xy, radius, knn, brute=brute)
    return knn_compact(cand, cnt, len(np.asarray(r_xy).reshape(-1, 2)))


def knn_priority(pairs, typeA, typeR):
    """a2 (knn_utils.py:31-65) on a1's re-indexed pairs; type codes of the re-indexed frames."""
    pairs, typeA, typeR = _i32(pairs).reshape(-1, 2), _i32(typeA), _i32(typeR)
    out = np.empty_like(pairs)
    q = lib().oracle_knn_priority(_p(pairs), C.c_int64(len(pairs)), _p(typeA), _p(typeR), C.c_int64(len(typeR)), _p(out))
    return out[:q].copy()


def pair_cost(pairs, a_xy, r_xy, a_prob, r_prob, dist_ct_coeff):
    """a3 (same.py:1182-1189)."""
    pairs = _i32(pairs).reshape(-1, 2)
    a_xy, r_xy = _f64(a_xy), _f64(r_xy)
    a_prob, r_prob = _f64(a_prob), _f64(r_prob)
    k = a_prob.shape[1] if a_prob.ndim == 2 else 0
    cost = np.empty(len(pairs), np.float64)
    lib().oracle_pair_cost(_p(pairs), C.c_int64(len(pairs)), _p(a_xy), _p(r_xy), _p(a_prob), _p(r_prob), C.c_int(k),
                           C.c_double(dist_ct_coeff), _p(cost))
    return cost


def group_pairs(pairs, column, n_nodes):
    """a4 (helpers.py:105-110): (group_node[G], indptr[G+1], idx[P]); groups in first-appearance order."""
    pairs = _i32(pairs).reshape(-1, 2)
    p = len(pairs)
    node = np.empty(max(n_nodes, 1), np.int32)
    indptr = np.zeros(max(n_nodes, 1) + 1, np.int64)
    idx = np.empty(max(p, 1), np.int32)
    g = lib().oracle_group_pairs(_p(pairs), C.c_int64(p), C.c_int(column), C.c_int64(n_nodes), _p(node), _p(indptr), _p(idx))
    return node[:g].copy(), indptr[:g + 1].copy(), idx[:p].copy()


def remap_triangles(tri, vertex_ids):
    """a5 (same.py:262-290) -> (tri_local[T',3] int32, src[T'] index into the input list)."""
    tri = np.ascontiguousarray(tri, dtype=np.int64).reshape(-1, 3)
    vid = np.ascontiguousarray(vertex_ids, dtype=np.int64)
    out = np.empty((max(len(tri), 1), 3), np.int32)
    src = np.empty(max(len(tri), 1), np.int32)
    t = lib().oracle_remap_triangles(_p(tri), C.c_int64(len(tri)), _p(vid), C.c_int64(len(vid)), _p(out), _p(src))
    return out[:t].copy(), src[:t].copy()


def tri_classify(xy, tri, radius, min_angle_deg, types, ignore_same_type):
    xy, tri = _f64(xy).reshape(-1, 2), _i32(tri).reshape(-1, 3)
    t = len(tri)
    cls = np.empty(max(t, 1), np.uint8)
    score = np.empty(max(t, 1), np.float64)
    band = np.empty(max(t, 1), np.uint8)
    ty = _i32(types) if types is not None else None
    lib().oracle_tri_classify(_p(xy), _p(tri), C.c_int64(t), C.c_double(radius), C.c_int(min_angle_deg is not None),
                              C.c_double(0.0 if min_angle_deg is None else min_angle_deg),
                              _p(ty) if ty is not None else None, C.c_int(bool(ignore_same_type)), _p(cls), _p(score), _p(band))
    return cls[:t], score[:t], band[:t]


def filter_triangles(xy, tri, radius, min_angle_deg, types, ignore_same_type, ensure_min=True):
    """a6 (helpers.py:233-395) -> (kept_src[T'] indices into tri in output order, unconstrained nodes (sorted), band count)."""
    xy, tri = _f64(xy).reshape(-1, 2), _i32(tri).reshape(-1, 3)
    cls, score, band = tri_classify(xy, tri, radius, min_angle_deg, types, ignore_same_type)
    n = len(xy)
    kept = np.empty(max(len(tri), 1), np.int32)
    valid = np.zeros(max(n, 1), np.uint8)
    o = lib().oracle_tri_select(_p(tri), C.c_int64(len(tri)), C.c_int64(n), _p(np.ascontiguousarray(cls)),
                                _p(np.ascontiguousarray(score)), C.c_int(bool(ignore_same_type)), C.c_int(bool(ensure_min)),
                                _p(kept), _p(valid))
    return kept[:o].copy(), np.flatnonzero(valid[:n] == 0).astype(np.int32), int(band.sum())


def tri_tables(xy, size, tri):
    """a8/a9 -> dict(weight, sign, bounds[T,4]=(min_x,max_x,min_y,max_y), argv[T,4]=(max_x,min_x,max_y,min_y vertex))."""
    xy, tri, size = _f64(xy).reshape(-1, 2), _i32(tri).reshape(-1, 3), _f64(size)
    t = len(tri)
    w = np.empty(max(t, 1)); s = np.empty(max(t, 1), np.int8)
    b = np.empty((max(t, 1), 4)); av = np.empty((max(t, 1), 4), np.int32)
    lib().oracle_tri_tables(_p(xy), _p(size), _p(tri), C.c_int64(t), _p(w), _p(s), _p(b), _p(av))
    return dict(weight=w[:t], sign=s[:t], bounds=b[:t], argv=av[:t])


def matching_from_x(x, pairs, na):
    x, pairs = _f64(x), _i32(pairs).reshape(-1, 2)
    mj = np.empty(max(na, 1), np.int32); mp = np.empty(max(na, 1), np.int32)
    lib().oracle_matching_from_x(_p(x), _p(pairs), C.c_int64(len(pairs)), C.c_int64(na), _p(mj), _p(mp))
    return mj[:na], mp[:na]


def separation(tri, source_sign, match_j, r_xy):
    """a10 (same.py:645-669) -> (violated triangle indices ascending, checked)."""
    tri = _i32(tri).reshape(-1, 3)
    s = np.ascontiguousarray(source_sign, dtype=np.int8)
    mj, r_xy = _i32(match_j), _f64(r_xy)
    viol = np.empty(max(len(tri), 1), np.int32)
    ck = C.c_int64(0)
    v = lib().oracle_separation(_p(tri), C.c_int64(len(tri)), _p(s), _p(mj), _p(r_xy), _p(viol), C.byref(ck))
    return viol[:v].copy(), ck.value


def lazy_cuts(x, pairs, tri, source_sign, r_xy, na, allowed_frac, per_inc_limit, max_cuts=None, cuts_added=0):
    """Full callback contract (same.py:621-703) -> list of (pa, pb, pc, t)."""
    if max_cuts is not None and cuts_added >= max_cuts:
        return np.zeros((0, 4), np.int64)
    mj, mp = matching_from_x(x, pairs, na)
    viol, checked = separation(tri, source_sign, mj, r_xy)
    if checked == 0 or len(viol) == 0:
        return np.zeros((0, 4), np.int64)
    if allowed_frac is not None and len(viol) / float(checked) <= allowed_frac:
        return np.zeros((0, 4), np.int64)
    lim = len(viol)
    if per_inc_limit is not None:
        lim = min(lim, per_inc_limit)
    if max_cuts is not None:
        lim = min(lim, max(0, max_cuts - cuts_added))
    tri = np.asarray(tri).reshape(-1, 3)
    v = viol[:lim]
    return np.column_stack([mp[tri[v, 0]], mp[tri[v, 1]], mp[tri[v, 2]], v]).astype(np.int64)


def postsolve(tri, a_xy, r_xy, match_j):
    """a11 + a12 -> dict(mask, area_before, area_after (nan = unmatched), flipped)."""
    tri, a_xy, r_xy, mj = _i32(tri).reshape(-1, 3), _f64(a_xy), _f64(r_xy), _i32(match_j)
    t = len(tri)
    mask = np.empty(max(t, 1), np.int32); ab = np.empty(max(t, 1)); aa = np.empty(max(t, 1)); fl = np.empty(max(t, 1), np.uint8)
    lib().oracle_postsolve(_p(tri), C.c_int64(t), _p(a_xy), _p(r_xy), _p(mj), _p(mask), _p(ab), _p(aa), _p(fl))
    return dict(mask=mask[:t], area_before=ab[:t], area_after=aa[:t], flipped=fl[:t].astype(bool))


def subset(xy, x_min, x_max, y_min, y_max):
    xy = _f64(xy).reshape(-1, 2)
    rows = np.empty(max(len(xy), 1), np.int32)
    o = lib().oracle_subset(_p(xy), C.c_int64(len(xy)), C.c_double(x_min), C.c_double(x_max), C.c_double(y_min),
                            C.c_double(y_max), _p(rows))
    return rows[:o].copy()


def greedy_select(nodes, key, n_nodes, eligible=None):
    """Ordered greedy selection with disjoint endpoints (init_helpers.py:110-132, metacell_utils.py:423-433) -> (selected bool [n], used bool [n_nodes])."""
    nodes = _i32(nodes)
    if nodes.ndim == 1:
        nodes = nodes.reshape(-1, 1)
    n, degree = nodes.shape
    key = _f64(key).reshape(n)
    el = None if eligible is None else np.ascontiguousarray(eligible, dtype=np.uint8).reshape(n)
    sel, used = np.zeros(n, np.uint8), np.zeros(int(n_nodes), np.uint8)
    lib().oracle_greedy_select(C.c_int64(n), C.c_int(degree), _p(nodes), _p(key), _p(el) if el is not None else None, C.c_int64(int(n_nodes)),
                               _p(sel), _p(used))
    return sel.astype(bool), used.astype(bool)


def mip_start_greedy(pairs, cost, n_aligned, n_ref, sizes, no_match_penalty):
    """compute_mip_start_pairs(init_method='greedy') -> (chosen [m,3] (i, j, var_idx) in selection order, unmatched sorted)."""
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    cost = _f64(cost)
    best = np.full(n_aligned, np.inf)
    np.minimum.at(best, pairs[:, 0], cost)
    prefer = best < float(no_match_penalty) * _f64(sizes)
    nodes = np.stack([pairs[:, 0], n_aligned + pairs[:, 1]], axis=1)
    sel, used = greedy_select(nodes, cost, n_aligned + n_ref, prefer[pairs[:, 0]])
    idx = np.flatnonzero(sel)
    idx = idx[np.argsort(cost[idx], kind="stable")]
    return np.stack([pairs[idx, 0], pairs[idx, 1], idx], axis=1), np.flatnonzero(~used[:n_aligned])


def collapse_select(xy, type_codes, sizes, tri, max_size):
    """One collapse iteration of greedy_triangle_collapse (metacell_utils.py:388-433) -> (selected bool [T], perimeter [T])."""
    xy = _f64(xy).reshape(-1, 2)
    tri = _i32(tri).reshape(-1, 3)
    tc, sz = _i32(type_codes), _f64(sizes)
    T = len(tri)
    cand, per = np.zeros(T, np.uint8), np.zeros(T, np.float64)
    lib().oracle_collapse_score(C.c_int64(T), _p(tri), _p(xy), _p(tc), _p(sz), C.c_double(float(max_size)), _p(cand), _p(per))
    sel, _ = greedy_select(tri, per, len(xy), cand)
    return sel, per


def segment_mean(values, ptr, pos):
    """Member means in pandas' / numpy's summation order (metacell_utils.py:446-474) -> [G, C]."""
    values = _f64(values)
    if values.ndim == 1:
        values = values.reshape(-1, 1)
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    pos = _i32(pos)
    G = len(ptr) - 1
    out = np.empty((G, values.shape[1]))
    lib().oracle_segment_mean(C.c_int64(values.shape[0]), C.c_int64(values.shape[1]), _p(values), C.c_int64(G), _p(ptr), _p(pos), _p(out))
    return out
