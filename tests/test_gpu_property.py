"""Property tests (hypothesis) of the candidate search on inputs built to provoke ties and threshold cases: integer / half-integer
lattices (many equal distances, duplicates, candidates exactly on the radius), random knn and window splits — CUDA == oracle
brute force, bit for bit.  The bucketed top-k of k_knn takes its exact fallback on most of these rows."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@st.composite
def lattice_case(draw):
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    step = draw(st.sampled_from([1.0, 0.5, 0.25, 3.0]))
    side = draw(st.integers(4, 40))
    na, nr = draw(st.integers(1, 900)), draw(st.integers(1, 900))
    a = rng.integers(0, side, size=(na, 2)).astype(np.float64) * step
    r = rng.integers(0, side, size=(nr, 2)).astype(np.float64) * step
    if draw(st.booleans()):                                    # a sprinkle of off-lattice points between exact ties
        k = rng.choice(nr, max(1, nr // 7), replace=False)
        r[k] += rng.uniform(-0.3, 0.3, size=(len(k), 2)) * step
    knn = draw(st.integers(1, 20))
    radius = step * draw(st.sampled_from([1.0, 2.0, 2.5, 5.0, float(np.sqrt(2.0)), float(np.sqrt(5.0)), 13.0]))
    split = draw(st.booleans())
    return a, r, knn, radius, split, side * step


@settings(max_examples=int(os.environ.get("SAME_B200_HYPOTHESIS_EXAMPLES", "40")), deadline=None, derandomize=os.environ.get("SAME_B200_HYPOTHESIS_RANDOM") is None,
          suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(lattice_case())
def test_candidates_on_lattices(case):
    from same_b200 import _lib as L
    from same_b200.device import Section
    a, r, knn, radius, split, extent = case
    rects = None
    if split:     # two overlapping windows: every window is its own search problem over the cells inside its rectangle
        rects = np.array([[-1.0, 0.6 * extent, -1.0, extent + 1.0], [0.4 * extent, extent + 1.0, -1.0, extent + 1.0]])
    with Section(a, r, np.zeros((len(a), 1)), np.zeros((len(r), 1))) as sec, sec.batch(rects) as b:
        b.candidates(radius, knn)
        pairs, keepA, keepR = b.get(L.PAIRS), b.get(L.KEEP_A), b.get(L.KEEP_R)
        po, ko, ro = b.offsets(L.PAIRS), b.offsets(L.KEEP_A), b.offsets(L.KEEP_R)
        for w in range(1 if rects is None else len(rects)):
            if rects is None:
                ra, rr = np.arange(len(a)), np.arange(len(r))
            else:
                x0, x1, y0, y1 = rects[w]
                ra = np.flatnonzero((a[:, 0] >= x0) & (a[:, 0] < x1) & (a[:, 1] >= y0) & (a[:, 1] < y1))
                rr = np.flatnonzero((r[:, 0] >= x0) & (r[:, 0] < x1) & (r[:, 1] >= y0) & (r[:, 1] < y1))
            kA, kR, pr = O.find_knn_within_radius(a[ra], r[rr], radius, knn, brute=True)
            assert np.array_equal(keepA[ko[w]:ko[w + 1]], ra[kA])
            assert np.array_equal(keepR[ro[w]:ro[w + 1]], rr[kR])
            assert np.array_equal(pairs[po[w]:po[w + 1]], pr)
