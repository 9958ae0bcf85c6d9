"""CPU-side checks of the boundary: the library builds, loads and exports every symbol the header declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    from same_b200 import _lib as L
    from same_b200 import build
    build.build()
    lib = L.load()
    header = open(os.path.join(ROOT, "include", "same_b200.h")).read()
    declared = sorted(set(re.findall(r"^SAME_API [^;(]*?\b(same_[a-z0-9_]+)\(", header, flags=re.M)))
    assert declared, "no declarations found in the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/same_b200.h but not exported"
    assert sorted(L.SYMBOLS) == declared


def test_array_spec_matches_elem_size():
    import numpy as np
    from same_b200 import _lib as L
    lib = L.load()
    for what, (dt, tail) in L.ARRAY_SPEC.items():
        n = int(np.prod(tail)) if tail else 1
        assert lib.same_elem_size(what) == np.dtype(dt).itemsize * n, what


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device the product must fail loudly, not compute on the CPU."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from same_b200 import _lib as L
    from same_b200.device import Section
    with pytest.raises(L.SameError):
        Section(np.zeros((4, 2)), np.zeros((4, 2)), np.zeros((4, 1)), np.zeros((4, 1)))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "same_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "same_oracle" not in src, f


def test_next_rows_have_no_cpu_fallback_either():
    """The MIP start / metacell collapse entry points fail loudly without a device (no numpy fallback behind them)."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import same_b200
    from same_b200 import _lib as L
    from same_b200 import datagen, init_helpers
    from same_b200.device import collapse_select, greedy_select, segment_mean
    with pytest.raises(L.SameError):
        greedy_select(np.array([[0, 1], [1, 2]]), np.array([1.0, 2.0]), 3)
    with pytest.raises(L.SameError):
        collapse_select(np.zeros((3, 2)), np.zeros(3, np.int32), np.ones(3), np.array([[0, 1, 2]]), 3)
    with pytest.raises(L.SameError):
        segment_mean(np.ones((3, 2)), np.array([0, 3]), np.array([0, 1, 2], np.int32))
    with pytest.raises(L.SameError):
        init_helpers.compute_mip_start_pairs(valid_pairs=[(0, 0)], costs=[1.0], n_aligned=1, n_ref=1, aligned_sizes=np.ones(1), no_match_penalty=5.0,
                                             max_matches=1, init_method="greedy", verbose=False)
    ref, qry, ct = datagen.make_section_pair(n_tiles=1, seed=3)
    with pytest.raises(L.SameError):
        same_b200.greedy_triangle_collapse(qry, max_metacell_size=3, r_max=1.5, min_angle_deg=10)
