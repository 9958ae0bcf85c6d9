"""GPU tests at BASELINE.json's full sizes and on the code paths only large batches reach.

Full-size checks use properties that do not need the oracle on the whole input (radius / order / count
of candidates against brute force on sampled rows, exact linearity of the cost in dist_ct_coeff,
idempotence of the separation call, shards == whole) plus full oracle parity on a few sampled windows.
"""
import numpy as np
import pytest
from scipy.spatial import Delaunay

from oracle import oracle as O
from oracle import pipeline as OP

pytestmark = pytest.mark.gpu


def _arrays(ref, qry, ct):
    lut = {c: i for i, c in enumerate(ct)}
    return dict(a_xy=np.ascontiguousarray(qry[["X", "Y"]].to_numpy(np.float64)), r_xy=np.ascontiguousarray(ref[["X", "Y"]].to_numpy(np.float64)),
                a_prob=np.ascontiguousarray(qry[ct].to_numpy(np.float64)), r_prob=np.ascontiguousarray(ref[ct].to_numpy(np.float64)),
                tA=qry["cell_type"].map(lut).to_numpy(np.int32), tR=ref["cell_type"].map(lut).to_numpy(np.int32))


def _incumbent(pairs, seed):
    rng = np.random.default_rng(seed)
    x = np.zeros(len(pairs))
    if len(pairs) == 0:
        return x
    i = pairs[:, 0].astype(np.int64)
    starts = np.flatnonzero(np.r_[True, i[1:] != i[:-1]])
    counts = np.diff(np.r_[starts, len(i)])
    sel = starts + rng.integers(0, 1 << 30, size=len(starts)) % counts
    x[sel[rng.uniform(size=len(starts)) < 0.9]] = 1.0
    return x


def _check_window_vs_oracle(b, w, rect, D, tri_vid, radius, knn, min_angle, x_w):
    """Everything the model builder and the callback consume, for one window, against the CPU oracle."""
    from same_b200 import _lib as L
    ra, rr = O.subset(D["a_xy"], *rect), O.subset(D["r_xy"], *rect)
    res = OP.window_pipeline(D["a_xy"][ra], D["r_xy"][rr], D["a_prob"][ra], D["r_prob"][rr], D["tA"][ra], D["tR"][rr], np.ones(len(ra)), np.ones(len(rr)),
                             radius=radius, knn=knn, min_angle_deg=min_angle, ignore_same_type_triangles=True, tri_global=tri_vid, a_vid=ra.astype(np.int64))
    m = b.window_model(w)
    assert np.array_equal(m["keepA"], ra[res["keepA"]])
    assert np.array_equal(m["keepR"], rr[res["keepR"]])
    for k in ("pairs", "cost", "tri", "sign", "weight", "ref_group_node", "ref_group_idx", "ref_group_limit"):
        assert np.array_equal(m[k], res[k]), k
    mj, mp = O.matching_from_x(x_w, m["pairs"], len(m["keepA"]))
    viol, checked = O.separation(m["tri"], m["sign"], mj, D["r_xy"][m["keepR"]])
    ps = O.postsolve(m["tri"], D["a_xy"][m["keepA"]], D["r_xy"][m["keepR"]], mj)
    return m, mj, mp, viol, checked, ps


def test_full_size_section_properties():
    """BASELINE configs[3]: 2,500 tiles (~1.03 M reference / ~0.93 M query cells), 7x7 windows."""
    from same_b200 import _lib as L
    from same_b200 import datagen
    from same_b200.device import Section
    radius, knn, window, overlap, min_angle = 250.0, 8, 5000, 250, 15.0
    ref, qry, ct = datagen.make_section_pair(n_tiles=2500, n_types=3, seed=2, scale=50.0, tiles_per_row=50)
    D = _arrays(ref, qry, ct)
    a_xy, r_xy = D["a_xy"], D["r_xy"]
    x_min, x_max = min(a_xy[:, 0].min(), r_xy[:, 0].min()), max(a_xy[:, 0].max(), r_xy[:, 0].max())
    y_min, y_max = min(a_xy[:, 1].min(), r_xy[:, 1].min()), max(a_xy[:, 1].max(), r_xy[:, 1].max())
    step = window - overlap
    rects = np.array([[x, x + window, y, y + window] for x in range(int(x_min), int(x_max), step) for y in range(int(y_min), int(y_max), step)], float)
    tri_vid = Delaunay(a_xy).simplices.astype(np.int64)
    rng = np.random.default_rng(0)
    with Section(a_xy, r_xy, D["a_prob"], D["r_prob"], D["tA"], D["tR"]) as sec:
        sec.set_triangles(tri_vid, None)
        with sec.batch(rects) as b:
            b.candidates(radius, knn, False, 1.0)
            off_a, off_r, off_p = b.offsets(L.KEEP_A), b.offsets(L.KEEP_R), b.offsets(L.PAIRS)
            keepA, keepR, pairs, cost, row_ptr = b.get(L.KEEP_A), b.get(L.KEEP_R), b.get(L.PAIRS), b.get(L.COST), b.get(L.ROW_PTR)
            assert len(pairs) == off_p[-1] > 8_000_000 and row_ptr[-1] == len(pairs)
            # window subsetting: instance lists are exactly the rows inside each (half-open) rectangle, ascending
            win_a, woff = b.get(L.WIN_A), b.offsets(L.WIN_A)
            for w in rng.choice(len(rects), 6, replace=False):
                assert np.array_equal(win_a[woff[w]:woff[w + 1]], O.subset(a_xy, *rects[w]))
            # every pair lies within the radius; rows are ordered by (d2, ref index); kept frames are sorted unique rows
            win_of_pair = np.searchsorted(off_p, np.arange(len(pairs)), side="right") - 1
            ga = keepA[off_a[win_of_pair] + pairs[:, 0]]
            gr = keepR[off_r[win_of_pair] + pairs[:, 1]]
            d = a_xy[ga] - r_xy[gr]
            d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
            assert (d2 <= radius * radius).all()
            same_row = (win_of_pair[1:] == win_of_pair[:-1]) & (pairs[1:, 0] == pairs[:-1, 0])
            assert ((d2[1:] > d2[:-1]) | ((d2[1:] == d2[:-1]) & (gr[1:] > gr[:-1])))[same_row].all()
            for w in range(len(rects)):
                assert (np.diff(keepA[off_a[w]:off_a[w + 1]]) > 0).all() and (np.diff(keepR[off_r[w]:off_r[w + 1]]) > 0).all()
            # sampled rows against brute force over the window's reference cells: the k nearest within the radius, in order
            for w in rng.choice(len(rects), 4, replace=False):
                rr = O.subset(r_xy, *rects[w])
                ka = keepA[off_a[w]:off_a[w + 1]]
                for li in rng.choice(len(ka), 150, replace=False):
                    dd = r_xy[rr] - a_xy[ka[li]]
                    dd2 = dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]
                    order = np.lexsort((rr, dd2))
                    order = order[dd2[order] <= radius * radius][:knn]
                    p0, p1 = row_ptr[off_a[w] + li], row_ptr[off_a[w] + li + 1]
                    got = keepR[off_r[w] + pairs[p0:p1, 1]]
                    assert np.array_equal(got, rr[order])
            # the cost is exactly linear in dist_ct_coeff for a power of two
            with sec.batch(rects) as b2:
                b2.candidates(radius, knn, False, 2.0)
                assert np.array_equal(b2.get(L.PAIRS), pairs) and np.array_equal(b2.get(L.COST), 2.0 * cost)
            # triangles, groups, one incumbent
            b.triangles_remap()
            assert b.tri_classify(radius, min_angle, True) == 0
            b.tri_finalize(True, True, True)
            b.groups(1, None)
            pairs, off_p = b.get(L.PAIRS), b.offsets(L.PAIRS)
            x = _incumbent(pairs, 5)
            nv, nc, cuts = b.separation(x, cap=1000)
            nv2, nc2, cuts2 = b.separation(x, cap=1000)
            assert np.array_equal(nv, nv2) and np.array_equal(nc, nc2)
            T = np.diff(b.offsets(L.TRI))
            assert (nv <= nc).all() and (nc <= T).all() and nc.sum() > 500_000
            for w in range(len(rects)):
                k = int(min(nv[w], 1000))
                assert np.array_equal(cuts[w, :k], cuts2[w, :k]) and (np.diff(cuts[w, :k, 3]) > 0).all()
            nv0, nc0, _ = b.separation(np.zeros(len(pairs)), cap=10)
            assert nv0.sum() == 0 and nc0.sum() == 0
            b.postsolve(x)
            flipped = b.get(L.FLIPPED)
            assert flipped.sum() > 0
            # three windows in full against the oracle (corner, edge, interior)
            for w in (0, 3, 24):
                xw = x[off_p[w]:off_p[w + 1]]
                m, mj, mp, viol, checked, ps = _check_window_vs_oracle(b, w, rects[w], D, tri_vid, radius, knn, min_angle, xw)
                assert nv[w] == len(viol) and nc[w] == checked
                k = min(len(viol), 1000)
                want = np.column_stack([mp[m["tri"][viol[:k], 0]], mp[m["tri"][viol[:k], 1]], mp[m["tri"][viol[:k], 2]], viol[:k]])
                assert np.array_equal(cuts[w, :k], want)
                assert np.array_equal(b.get_window(L.TRI_MASK, w), ps["mask"])
                assert np.array_equal(b.get_window(L.FLIPPED, w).astype(bool), ps["flipped"])
                assert np.array_equal(b.get_window(L.AREA_AFTER, w), ps["area_after"], equal_nan=True)
            # shards == whole: the window list split in two batches gives the same per-window arrays
            full_pairs, full_tri = b.get(L.PAIRS), b.get(L.TRI)
            parts_p, parts_t = [], []
            for lo, hi in ((0, 25), (25, len(rects))):
                with sec.batch(rects[lo:hi]) as bs:
                    bs.candidates(radius, knn, False, 1.0)
                    bs.triangles_remap()
                    bs.tri_classify(radius, min_angle, True)
                    bs.tri_finalize(True, True, True)
                    parts_p.append(bs.get(L.PAIRS)); parts_t.append(bs.get(L.TRI))
            assert np.array_equal(np.concatenate(parts_p), full_pairs) and np.array_equal(np.concatenate(parts_t), full_tri)


def test_luad_shape_single_window_vs_oracle():
    """BASELINE configs[4] shape: ~94 K query / 100 K reference cells, K = 5, one 13,000-unit window."""
    from same_b200 import _lib as L
    from same_b200.device import Section
    rng = np.random.default_rng(4)
    nA, nR, K, radius, knn = 94_000, 100_000, 5, 250.0, 8
    a_xy, r_xy = rng.uniform(0, 13_000, (nA, 2)), rng.uniform(0, 13_000, (nR, 2))
    a_prob, r_prob = rng.dirichlet([0.3] * K, nA) * 100, rng.dirichlet([0.3] * K, nR) * 100
    tA, tR = a_prob.argmax(1).astype(np.int32), r_prob.argmax(1).astype(np.int32)
    sA, sR = rng.integers(1, 4, nA).astype(np.float64), rng.integers(1, 4, nR).astype(np.float64)   # metacell sizes 1..3
    with Section(a_xy, r_xy, a_prob, r_prob, tA, tR, sA, sR) as sec, sec.batch() as b:
        b.candidates(radius, knn, False, 1.0)
        keepA = b.get(L.KEEP_A)
        assert b.length(L.KEEP_R) > 65536
        with pytest.raises(L.SameError) as limit:       # a window with more than 65,536 kept reference rows has no 16-bit form
            b.get(L.PAIR_J16)
        assert limit.value.code == L.E_LIMIT
        tri = Delaunay(a_xy[keepA]).simplices.astype(np.int32)
        b.triangles_set(tri, [0, len(tri)])
        nb = b.tri_classify(50.0, 15.0, True)
        assert nb == 0
        b.tri_finalize(True, True, False)
        b.groups(2, None)
        m = b.window_model(0)
        res = OP.window_pipeline(a_xy, r_xy, a_prob, r_prob, tA, tR, sA, sR, radius=radius, knn=knn, max_matches=2)
        # window_pipeline filters with radius as the edge limit; redo the triangle part with the 50-unit limit used above
        kept, unc, band = O.filter_triangles(a_xy[res["keepA"]], tri, 50.0, 15.0, tA[res["keepA"]], True)
        tt = O.tri_tables(a_xy[res["keepA"]], sA[res["keepA"]], tri[kept])
        assert np.array_equal(m["keepA"], res["keepA"]) and np.array_equal(m["keepR"], res["keepR"])
        for k in ("pairs", "cost", "ref_group_node", "ref_group_idx", "ref_group_limit"):
            assert np.array_equal(m[k], res[k]), k
        assert np.array_equal(m["tri"], tri[kept]) and np.array_equal(m["sign"], tt["sign"]) and np.array_equal(m["weight"], tt["weight"])
        x = _incumbent(m["pairs"], 1)
        nv, nc, cuts = b.separation(x, cap=1000)
        mj, mp = O.matching_from_x(x, m["pairs"], len(m["keepA"]))
        viol, checked = O.separation(m["tri"], m["sign"], mj, r_xy[m["keepR"]])
        assert nv[0] == len(viol) and nc[0] == checked
        k = min(len(viol), 1000)
        assert np.array_equal(cuts[0, :k, 3], viol[:k])


@pytest.mark.parametrize("grid,n_cells", [(70, 60_000), (92, 280_000)])
def test_many_small_windows(grid, n_cells):
    """More windows than the shared-memory window histogram holds (4,096); the larger case also needs 64-bit subset keys.
    Most windows hold a handful of cells, many are empty, and a separation tile spans dozens of windows."""
    from same_b200 import _lib as L
    from same_b200.device import Section
    rng = np.random.default_rng(grid)
    ext = 1000.0
    a_xy = rng.uniform(0, ext, (n_cells, 2))
    a_xy = a_xy[(a_xy[:, 0] % 200 > 40) | (a_xy[:, 1] % 300 > 60)]      # holes -> empty windows
    r_xy = a_xy + rng.normal(0, 0.4, a_xy.shape)
    K = 3
    a_prob, r_prob = rng.dirichlet([0.5] * K, len(a_xy)) * 100, rng.dirichlet([0.5] * K, len(a_xy)) * 100
    tA, tR = a_prob.argmax(1).astype(np.int32), r_prob.argmax(1).astype(np.int32)
    pitch = ext / grid
    rects = np.array([[x * pitch, x * pitch + 1.3 * pitch, y * pitch, y * pitch + 1.3 * pitch] for x in range(grid) for y in range(grid)], float)
    radius, knn = 8.0, 5
    tri_vid = Delaunay(a_xy).simplices.astype(np.int64)
    D = dict(a_xy=a_xy, r_xy=r_xy, a_prob=a_prob, r_prob=r_prob, tA=tA, tR=tR)
    with Section(a_xy, r_xy, a_prob, r_prob, tA, tR) as sec:
        sec.set_triangles(tri_vid, None)
        with sec.batch(rects) as b:
            assert b.W > 4096
            b.candidates(radius, knn, False, 1.0)
            win_a, woff = b.get(L.WIN_A), b.offsets(L.WIN_A)
            sizes = np.diff(woff)
            assert (sizes == 0).any()
            b.triangles_remap()
            assert b.tri_classify(radius, 10.0, True) == 0
            b.tri_finalize(True, True, True)
            b.groups(1, None)
            pairs, off_p = b.get(L.PAIRS), b.offsets(L.PAIRS)
            x = _incumbent(pairs, 2)
            nv, nc, cuts = b.separation(x, cap=8)
            b.postsolve(x)
            T = np.diff(b.offsets(L.TRI))
            assert (nv <= nc).all() and (nc <= T).all() and nv.sum() > 0
            assert (nv[T == 0] == 0).all()
            pick = np.r_[rng.choice(len(rects), 25, replace=False), np.argsort(-sizes)[:3], np.flatnonzero(sizes == 0)[:2]]
            n_checked_windows = 0
            for w in pick:
                assert np.array_equal(win_a[woff[w]:woff[w + 1]], O.subset(a_xy, *rects[w]))
                if off_p[w + 1] == off_p[w]:
                    assert nv[w] == 0 and nc[w] == 0
                    continue
                xw = x[off_p[w]:off_p[w + 1]]
                m, mj, mp, viol, checked, ps = _check_window_vs_oracle(b, w, rects[w], D, tri_vid, radius, knn, 10.0, xw)
                assert nv[w] == len(viol) and nc[w] == checked
                k = min(len(viol), 8)
                want = np.column_stack([mp[m["tri"][viol[:k], 0]], mp[m["tri"][viol[:k], 1]], mp[m["tri"][viol[:k], 2]], viol[:k]])
                assert np.array_equal(cuts[w, :k], want)
                assert np.array_equal(b.get_window(L.FLIPPED, w).astype(bool), ps["flipped"])
                assert np.array_equal(b.get_window(L.TRI_MASK, w), ps["mask"])
                n_checked_windows += 1
            assert n_checked_windows >= 20
