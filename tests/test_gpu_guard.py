"""Guard mode (same_debug_guard / SAME_B200_GUARD=1): the library's own memory checker, for GPU boxes where compute-sanitizer cannot
run.  Every device buffer gets canary zones that are verified on release and a poisoned body on every (re)allocation; a pipeline
run in that mode must (a) leave every zone intact and (b) still equal the oracle bit for bit — a kernel reading memory the library
never wrote would see 0xCD bytes instead of whatever a recycled pool block happened to hold."""
import numpy as np
import pytest
from scipy.spatial import Delaunay

from oracle import oracle as O
from oracle import pipeline as OP

pytestmark = pytest.mark.gpu


def test_guard_mode_pipeline_equals_oracle_and_no_zone_is_touched():
    from same_b200 import _lib as L
    from same_b200 import datagen
    from same_b200.device import Section
    ref, qry, ct = datagen.make_section_pair(n_tiles=9, n_types=3, seed=17)
    a_xy, r_xy = qry[["X", "Y"]].to_numpy(), ref[["X", "Y"]].to_numpy()
    a_prob, r_prob = qry[ct].to_numpy(), ref[ct].to_numpy()
    lut = {c: i for i, c in enumerate(ct)}
    tA = qry["cell_type"].map(lut).to_numpy(np.int32)
    tR = ref["cell_type"].map(lut).to_numpy(np.int32)
    sA, sR = np.ones(len(qry)), np.ones(len(ref))
    rng = np.random.default_rng(5)
    vid = (rng.permutation(len(qry)) * 3 + 7).astype(np.int64)
    tri_g = Delaunay(a_xy).simplices
    side = np.linalg.norm(a_xy[tri_g] - a_xy[np.roll(tri_g, 1, axis=1)], axis=2).max(axis=1)
    tri_vid = vid[tri_g[side < 0.9]]
    ext = max(a_xy.max(), r_xy.max()) + 1
    rects = np.array([[x0, x0 + 14.0, y0, y0 + 14.0] for x0 in np.arange(0, ext, 11.0) for y0 in np.arange(0, ext, 11.0)])
    radius, knn = 1.0, 6
    kw = dict(radius=radius, knn=knn, dist_ct_coeff=1.0, min_angle_deg=12, ignore_same_type_triangles=True, max_matches=1)
    bad0, checked0 = L.debug_guard(True)
    try:
        for rep in range(2):                      # the second pass runs on recycled (previously used, now re-poisoned) pool blocks
            with Section(a_xy, r_xy, a_prob, r_prob, tA, tR, sA, sR) as sec, sec.batch(rects) as b:
                sec.set_triangles(tri_vid, vid)
                b.candidates(radius, knn)
                b.triangles_remap()
                assert b.tri_classify(radius, 12.0, True) == 0
                b.tri_finalize(True, True, remove_unconstrained=True)
                b.groups(1, None)
                xs, per_window = [], []
                for w in range(len(rects)):
                    ra, rr = O.subset(a_xy, *rects[w]), O.subset(r_xy, *rects[w])
                    if len(ra) == 0 or len(rr) == 0:
                        continue
                    res = OP.window_pipeline(a_xy[ra], r_xy[rr], a_prob[ra], r_prob[rr], tA[ra], tR[rr], sA[ra], sR[rr],
                                             tri_global=tri_vid, a_vid=vid[ra], **kw)
                    m = b.window_model(w)
                    for k in ("pairs", "cost", "tri", "sign", "weight", "ref_group_node", "ref_group_idx"):
                        assert np.array_equal(m[k], res[k]), (rep, w, k)
                    if len(m["pairs"]) == 0:
                        continue
                    x = np.zeros(len(m["pairs"]))
                    first = np.flatnonzero(np.r_[True, np.diff(m["pairs"][:, 0]) != 0])
                    x[first] = 1.0
                    nv, nc, cuts = b.separation(x, w, w + 1, cap=40)
                    mj, _ = O.matching_from_x(x, m["pairs"], len(m["keepA"]))
                    viol, n_checked = O.separation(m["tri"], m["sign"], mj, r_xy[m["keepR"]])
                    assert nv[0] == len(viol) and nc[0] == n_checked
                    xs.append(x)
                    per_window.append((w, m, mj))
                b.postsolve(np.concatenate(xs))
                for w, m, mj in per_window:
                    ps = O.postsolve(m["tri"], a_xy[m["keepA"]], r_xy[m["keepR"]], mj)
                    assert np.array_equal(b.get_window(L.TRI_MASK, w), ps["mask"])
    finally:
        bad, checked = L.debug_guard(False)
    assert checked - checked0 > 100, "guard mode did not check any buffers"
    assert bad - bad0 == 0, f"{bad - bad0} device buffers were written outside their bounds"


def test_guard_checker_detects_a_deliberate_overrun():
    from same_b200 import _lib as L
    bad0, checked0 = L.debug_guard()
    bad1, checked1 = L.debug_guard("selftest")
    assert bad1 - bad0 == 1 and checked1 - checked0 == 1
    import os
    os.environ["SAME_B200_GUARD_EXPECTED"] = str(int(os.environ.get("SAME_B200_GUARD_EXPECTED", "0")) + 1)   # (conftest's guard report)
