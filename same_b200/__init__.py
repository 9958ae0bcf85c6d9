"""same_b200 — B200-native hot path of SAME (Spatial Alignment of Multimodal Expression).

Drop-in for the reference's public API (`src/__init__.py:51-65`): the same eight names, signatures and DataFrame
contract; candidate generation, pair costs, constraint grouping, triangle tables, lazy-constraint separation,
post-solve analysis and window subsetting run as sm_100a CUDA kernels in `libsame_b200.so` (C-ABI in
`include/same_b200.h`).  Gurobi stays on the host.  There is no CPU fallback.
"""
from .params import init_gurobi_params, init_optim_params

__version__ = "0.1.0"
__all__ = ["init_gurobi_params", "init_optim_params", "sliding_window_matching", "run_same", "merge_window_matches_unique_ref",
           "MetaCell", "greedy_triangle_collapse", "unpack_metacell_matches"]

_LAZY = {
    "run_same": ("same", "run_same"), "sliding_window_matching": ("same", "sliding_window_matching"),
    "MetaCell": ("metacell_utils", "MetaCell"), "greedy_triangle_collapse": ("metacell_utils", "greedy_triangle_collapse"),
    "unpack_metacell_matches": ("metacell_utils", "unpack_metacell_matches"),
    "merge_window_matches_unique_ref": ("merge", "merge_window_matches_unique_ref"),
}


def __getattr__(name):          # keeps `import same_b200.datagen` light (no pandas/scipy/ctypes import chain)
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    raise AttributeError(name)
