"""Solver back-ends behind `run_same`.

Gurobi stays the reference solver on the host (BASELINE.json north_star).  `GurobiBackend` builds the model
with the reference's variable, constraint and cut order and names (src/same.py:1112-1197,
src/helpers.py:130-158) from the arrays the GPU produced, and routes the MIPSOL callback to the GPU
separation kernel.  `HighsCutLoopBackend` (scipy.optimize.milp) is a stand-in for machines without a Gurobi
licence: solve -> separate -> add cuts -> re-solve; it is not expected to reproduce Gurobi's incumbents.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np


@dataclass
class ModelSpec:
    """Everything the MIP needs, as flat arrays in the reference's index spaces (one window)."""
    n_pairs: int
    n_ref: int
    n_aligned: int
    n_tri: int
    cost: np.ndarray                 # [P]   objective coefficient of x[idx]          (same.py:1182-1189)
    row_ptr: np.ndarray              # [Na+1] pairs of aligned row i = [ptr[i], ptr[i+1])  (helpers.py:107-110)
    ref_group_node: np.ndarray       # [G]   j of each ref group, first-appearance order    (helpers.py:105-106)
    ref_group_ptr: np.ndarray        # [G+1]
    ref_group_idx: np.ndarray        # [P]   pair indices, ascending inside a group
    ref_group_limit: np.ndarray      # [G]   rhs of max_matches_<j>                    (helpers.py:130-138)
    aligned_size: np.ndarray         # [Na]
    tri_weight: np.ndarray           # [T]   (same.py:1128-1135)
    penalty_coeff: float = 100.0
    no_match_penalty: float = 100.0
    delaunay_penalty: float = 5.0


@dataclass
class SolveResult:
    status: str                      # "optimal" | "time_limit" | other
    x: np.ndarray
    no_match: np.ndarray
    penalty: np.ndarray
    q: np.ndarray
    runtime: float
    cuts_added: int
    model: object = None


SeparationFn = Callable[[np.ndarray, int], np.ndarray]   # (x values, cuts added so far) -> [n, 4] (pa, pb, pc, t)


class GurobiBackend:
    """The reference's model, line for line in structure, through gurobipy (>= 13 for METHOD_PDHG)."""

    name = "gurobi"

    def solve(self, spec: ModelSpec, separate: Optional[SeparationFn], gurobi_params: dict, outprefix=None, env_options=None, start=None):
        import gurobipy as gp
        from gurobipy import GRB, Model, quicksum

        log_dir = os.path.join(os.getcwd(), "gurobi_logs")                      # same.py:868-870
        os.makedirs(log_dir, exist_ok=True)
        options = {"OutputFlag": 1, "LogFile": os.path.join(log_dir, f"gurobi_{os.getpid()}.log")}
        options.update(env_options or {})
        try:
            env = gp.Env(params=options)
            model = Model("optimal_matches", env=env)
            P, nr, na, T = spec.n_pairs, spec.n_ref, spec.n_aligned, spec.n_tri
            x = model.addVars(P, vtype=GRB.BINARY, lb=0, ub=1, name="x")                            # same.py:1116
            penalty_vars = model.addVars(nr, vtype=GRB.CONTINUOUS, lb=0, ub=1000, name="penalty")
            no_match_vars = model.addVars(na, vtype=GRB.CONTINUOUS, lb=0, ub=1, name="no_match")
            model.update()
            gp_, gi = spec.ref_group_ptr, spec.ref_group_idx
            groups = [(int(j), [int(k) for k in gi[gp_[g]:gp_[g + 1]]], spec.ref_group_limit[g].item())
                      for g, j in enumerate(spec.ref_group_node)]
            rows = [(i, range(int(spec.row_ptr[i]), int(spec.row_ptr[i + 1]))) for i in range(na) if spec.row_ptr[i + 1] > spec.row_ptr[i]]
            for j, idxs, limit in groups:                                                             # helpers.py:130-138
                model.addConstr(quicksum(x[k] for k in idxs) <= limit, name=f"max_matches_{j}")
            model.update()
            for i, idxs in rows:                                                                      # helpers.py:142-145
                model.addConstr(quicksum(x[k] for k in idxs) <= 1, name=f"one_match_{i}")
            model.update()
            for j, idxs, _ in groups:                                                                 # helpers.py:149-152
                model.addConstr(quicksum(x[k] for k in idxs) - penalty_vars[j] <= 1, name=f"penalty_{j}")
            model.update()
            for i, idxs in rows:                                                                      # helpers.py:156-158
                model.addConstr(quicksum(x[k] for k in idxs) + no_match_vars[i] == 1, name=f"no_match_{i}")
            model.update()
            q_tri = model.addVars(T, vtype=GRB.CONTINUOUS, lb=0, name="q_tri")                        # same.py:1149
            model.update()
            state = {"cuts": 0}
            model.Params.LazyConstraints = 1                                                          # same.py:1165-1170
            model.Params.Method = gp.GRB.METHOD_PDHG
            model.Params.PDHGGPU = 1
            model.update()
            c = spec.cost.tolist()
            sizes = spec.aligned_size.tolist()
            w = spec.tri_weight.tolist()
            model.setObjective(                                                                       # same.py:1191-1197
                quicksum(c[k] * x[k] for k in range(P)) +
                spec.penalty_coeff * quicksum(penalty_vars[j] for j in range(nr)) +
                spec.no_match_penalty * quicksum(sizes[i] * no_match_vars[i] for i in range(na)) +
                spec.delaunay_penalty * quicksum(w[t] * q_tri[t] for t in range(T)),
                GRB.MINIMIZE)
            if start is not None:                                                                     # init_helpers.py:229-236
                x0, nm0 = start
                for k in range(P):
                    x[k].Start = float(x0[k])
                for i in range(na):
                    no_match_vars[i].Start = float(nm0[i])
            if outprefix:                                                                             # same.py:1218-1224
                os.makedirs(outprefix, exist_ok=True)
                model_file = os.path.join(outprefix, "matching_model.lp")
            else:
                model_file = "matching_model.lp"
            model.write(model_file)
            tl = gurobi_params.get("time_limit")
            model.Params.timeLimit = float(tl) if tl is not None else float("inf")
            model.Params.MIPGap = float(gurobi_params.get("mip_gap", 0.05))
            if gurobi_params.get("mip_focus") is not None:
                model.Params.MIPFocus = int(gurobi_params["mip_focus"])
            if gurobi_params.get("cuts") is not None:
                model.Params.Cuts = int(gurobi_params["cuts"])
            if gurobi_params.get("heuristics") is not None:
                model.Params.Heuristics = float(gurobi_params["heuristics"])
            xs = [x[k] for k in range(P)]
            model._x, model._q_tri, model._row_ptr = x, q_tri, spec.row_ptr
            lazy_max = gurobi_params.get("lazy_max_cuts")

            def callback(m, where):                                                                   # same.py:621-703
                if where != GRB.Callback.MIPSOL:
                    return
                if lazy_max is not None and state["cuts"] >= lazy_max:
                    return
                vals = np.asarray(m.cbGetSolution(xs), dtype=np.float64)   # one contiguous fetch, not a P-entry dict
                for pa, pb, pc, t in separate(vals, state["cuts"]):
                    m.cbLazy(x[int(pa)] + x[int(pb)] + x[int(pc)] <= 2 + q_tri[int(t)])
                    state["cuts"] += 1

            if separate is not None:
                model.optimize(callback)
            else:
                model.optimize()
            status = "optimal" if model.status == GRB.OPTIMAL else ("time_limit" if model.status == GRB.TIME_LIMIT else f"status_{model.status}")
            if status in ("optimal", "time_limit"):
                res = SolveResult(status, np.array([x[k].x for k in range(P)], dtype=np.float64),
                                  np.array([no_match_vars[i].x for i in range(na)], dtype=np.float64),
                                  np.array([penalty_vars[j].x for j in range(nr)], dtype=np.float64),
                                  np.array([q_tri[t].x for t in range(T)], dtype=np.float64), float(model.Runtime), state["cuts"], model)
            else:
                res = SolveResult(status, np.zeros(P), np.zeros(na), np.zeros(nr), np.zeros(T), float(model.Runtime), state["cuts"], model)
            return res
        finally:
            try:
                if os.path.exists(log_dir) and not os.listdir(log_dir):
                    os.rmdir(log_dir)
            except Exception:
                pass


def model_matrices(spec: ModelSpec):
    """The base model as ONE sparse matrix in the reference's creation order (src/helpers.py:130-158), assembled with array
    operations straight from the GPU-built CSR (aligned rows) / CSC (reference groups) — no per-row Python expression.

    Variables are numbered as the reference creates them (src/same.py:1116-1149): x[0..P) | penalty[P..P+Nr) | no_match[..+Na) |
    q_tri[..+T).  Rows: max_matches_<j> per reference group (first-appearance order), one_match_<i> per aligned row,
    penalty_<j>, no_match_<i>.  -> dict(ptr, idx, val, sense ('<=' / '=='), rhs, names, n_vars, obj)"""
    P, nr, na, T = spec.n_pairs, spec.n_ref, spec.n_aligned, spec.n_tri
    gp = np.asarray(spec.ref_group_ptr, dtype=np.int64)
    gi = np.asarray(spec.ref_group_idx, dtype=np.int64)
    gn = np.asarray(spec.ref_group_node, dtype=np.int64)
    rp = np.asarray(spec.row_ptr, dtype=np.int64)
    G = len(gn)
    rows = np.flatnonzero(np.diff(rp) > 0)                       # aligned rows that own pairs (all of them after KNN compaction)
    rsz, gsz = np.diff(rp)[rows], np.diff(gp)
    R = len(rows)
    row_members = np.concatenate([np.arange(rp[i], rp[i + 1]) for i in rows]) if R and rsz.sum() != P else np.arange(P, dtype=np.int64)
    # block 3 / 4: members followed by the row's own penalty / no_match variable
    def with_tail(members, sizes, tail):
        n = len(sizes)
        out = np.empty(len(members) + n, dtype=np.int64)
        start = np.r_[0, np.cumsum(sizes)][:-1]
        mpos = np.arange(len(members)) + np.repeat(np.arange(n), sizes)
        out[mpos] = members
        out[start + sizes + np.arange(n)] = tail
        return out
    idx = np.concatenate([gi, row_members, with_tail(gi, gsz, P + gn), with_tail(row_members, rsz, P + nr + rows)])
    one_g, one_r = np.ones(len(gi)), np.ones(len(row_members))
    tail_val = lambda sizes, v: with_tail(np.ones(int(sizes.sum())), sizes, np.full(len(sizes), v)) if len(sizes) else np.zeros(0)
    val = np.concatenate([one_g, one_r, tail_val(gsz, -1.0), tail_val(rsz, 1.0)])
    ptr = np.r_[0, np.cumsum(np.concatenate([gsz, rsz, gsz + 1, rsz + 1]))].astype(np.int64)
    sense = np.asarray(["<="] * (2 * G + R) + ["=="] * R)
    rhs = np.concatenate([np.asarray(spec.ref_group_limit, dtype=np.float64), np.ones(R), np.ones(G), np.ones(R)])
    sg, sr = gn.astype(str), rows.astype(str)
    names = np.concatenate([np.char.add("max_matches_", sg), np.char.add("one_match_", sr), np.char.add("penalty_", sg),
                            np.char.add("no_match_", sr)]) if (G + R) else np.zeros(0, dtype=str)
    obj = np.concatenate([np.asarray(spec.cost, dtype=np.float64), np.full(nr, float(spec.penalty_coeff)),
                          float(spec.no_match_penalty) * np.asarray(spec.aligned_size, dtype=np.float64),
                          float(spec.delaunay_penalty) * np.asarray(spec.tri_weight, dtype=np.float64)])   # same.py:1191-1197
    return dict(ptr=ptr, idx=idx, val=val, sense=sense, rhs=rhs, names=names, n_vars=P + nr + na + T, obj=obj)


class GurobiMatrixBackend:
    """Same model, same variable / constraint / cut order and names as `GurobiBackend`, built through gurobipy's matrix API
    (`addMVar` + one `addMConstr`) from `model_matrices` instead of one `quicksum` per row.  The DEFAULT
    back-end ('gurobi'); the per-row builder stays available as 'gurobi_rows' (SAME_B200_SOLVER / `set_default_backend`) and is
    what this class falls back to when the installed gurobipy has no matrix API.  Needs gurobipy >= 13 like the reference
    (GRB.METHOD_PDHG, src/same.py:1169-1170).  tests/test_gpu_api.py runs BOTH builders against the recording fake gurobipy
    (oracle/ref_loader.py) and compares variables, objective, constraints (names, order, members, sense, rhs) and lazy cuts with the
    reference's recorded models; with a real gurobipy installed, tests/test_gurobi_real.py solves the 144-cell fixtures."""

    name = "gurobi_matrix"

    def solve(self, spec: ModelSpec, separate: Optional[SeparationFn], gurobi_params: dict, outprefix=None, env_options=None, start=None):
        import gurobipy as gp
        from gurobipy import GRB
        from scipy.sparse import csr_matrix

        if not (hasattr(gp.Model, "addMVar") and hasattr(gp.Model, "addMConstr") and hasattr(gp, "MVar") and hasattr(gp.MVar, "fromlist")):
            import warnings
            warnings.warn("this gurobipy has no matrix API (addMVar / addMConstr / MVar.fromlist): building the model row by row")
            return GurobiBackend().solve(spec, separate, gurobi_params, outprefix=outprefix, env_options=env_options, start=start)
        log_dir = os.path.join(os.getcwd(), "gurobi_logs")
        os.makedirs(log_dir, exist_ok=True)
        options = {"OutputFlag": 1, "LogFile": os.path.join(log_dir, f"gurobi_{os.getpid()}.log")}
        options.update(env_options or {})
        env = gp.Env(params=options)
        model = gp.Model("optimal_matches", env=env)
        P, nr, na, T = spec.n_pairs, spec.n_ref, spec.n_aligned, spec.n_tri
        mm = model_matrices(spec)
        x = model.addMVar(P, vtype=GRB.BINARY, lb=0, ub=1, name="x")
        pen = model.addMVar(nr, vtype=GRB.CONTINUOUS, lb=0, ub=1000, name="penalty")
        nom = model.addMVar(na, vtype=GRB.CONTINUOUS, lb=0, ub=1, name="no_match")
        q = model.addMVar(T, vtype=GRB.CONTINUOUS, lb=0, name="q_tri")
        model.update()
        allv = gp.MVar.fromlist(x.tolist() + pen.tolist() + nom.tolist() + q.tolist())
        A = csr_matrix((mm["val"], mm["idx"], mm["ptr"]), shape=(len(mm["rhs"]), mm["n_vars"]))
        sense = np.where(mm["sense"] == "==", GRB.EQUAL, GRB.LESS_EQUAL)
        cons = model.addMConstr(A, allv, sense, mm["rhs"])
        model.update()
        for c, nm in zip(cons.tolist(), mm["names"].tolist()):
            c.ConstrName = nm
        model.setObjective(mm["obj"] @ allv, GRB.MINIMIZE)
        model.Params.LazyConstraints = 1
        model.Params.Method = gp.GRB.METHOD_PDHG
        model.Params.PDHGGPU = 1
        if start is not None:                                                  # init_helpers.apply_mip_start
            x.Start = np.asarray(start[0], dtype=float)
            nom.Start = np.asarray(start[1], dtype=float)
        model.write(os.path.join(outprefix, "matching_model.lp") if outprefix else "matching_model.lp")
        tl = gurobi_params.get("time_limit")
        model.Params.timeLimit = float(tl) if tl is not None else float("inf")
        model.Params.MIPGap = float(gurobi_params.get("mip_gap", 0.05))
        for key, attr, conv in (("mip_focus", "MIPFocus", int), ("cuts", "Cuts", int), ("heuristics", "Heuristics", float)):
            if gurobi_params.get(key) is not None:
                setattr(model.Params, attr, conv(gurobi_params[key]))
        xs, qs = x.tolist(), q.tolist()
        model._x, model._q_tri, model._row_ptr = x, q, spec.row_ptr
        state = {"cuts": 0}
        lazy_max = gurobi_params.get("lazy_max_cuts")

        def callback(m, where):
            if where != GRB.Callback.MIPSOL or (lazy_max is not None and state["cuts"] >= lazy_max):
                return
            vals = np.asarray(m.cbGetSolution(xs), dtype=np.float64)
            for pa, pb, pc, t in separate(vals, state["cuts"]):
                m.cbLazy(xs[int(pa)] + xs[int(pb)] + xs[int(pc)] <= 2 + qs[int(t)])
                state["cuts"] += 1

        model.optimize(callback) if separate is not None else model.optimize()
        status = "optimal" if model.status == GRB.OPTIMAL else ("time_limit" if model.status == GRB.TIME_LIMIT else f"status_{model.status}")
        if status in ("optimal", "time_limit"):
            return SolveResult(status, np.asarray(x.X, dtype=np.float64), np.asarray(nom.X, dtype=np.float64), np.asarray(pen.X, dtype=np.float64),
                               np.asarray(q.X, dtype=np.float64), float(model.Runtime), state["cuts"], model)
        return SolveResult(status, np.zeros(P), np.zeros(na), np.zeros(nr), np.zeros(T), float(model.Runtime), state["cuts"], model)


class HighsCutLoopBackend:
    """scipy.optimize.milp (HiGHS) with an explicit lazy-cut loop.  Variables: x[P] | penalty[Nr] | no_match[Na] | q[T]."""

    name = "highs"

    def __init__(self, max_rounds=50):
        self.max_rounds = max_rounds

    def solve(self, spec: ModelSpec, separate: Optional[SeparationFn], gurobi_params: dict, outprefix=None, env_options=None, start=None):
        # `start` is accepted for interface parity; scipy.optimize.milp has no warm start
        from scipy.optimize import Bounds, LinearConstraint, milp
        from scipy.sparse import coo_matrix

        from scipy.sparse import csr_matrix
        P, nr, na, T = spec.n_pairs, spec.n_ref, spec.n_aligned, spec.n_tri
        mm = model_matrices(spec)
        n, c = mm["n_vars"], mm["obj"]
        A = csr_matrix((mm["val"], mm["idx"], mm["ptr"]), shape=(len(mm["rhs"]), n))
        hi = mm["rhs"]
        lo = np.where(mm["sense"] == "==", mm["rhs"], -np.inf)
        integrality = np.r_[np.ones(P), np.zeros(nr + na + T)]
        bounds = Bounds(np.zeros(n), np.r_[np.ones(P), np.full(nr, 1000.0), np.ones(na), np.full(T, np.inf)])
        cut_rows = []
        cuts_added = 0
        t0 = time.perf_counter()
        tl = gurobi_params.get("time_limit")
        opts = {"mip_rel_gap": float(gurobi_params.get("mip_gap", 0.05))}
        if tl is not None:
            opts["time_limit"] = float(tl)
        status, sol = "status_unknown", None
        lazy_max = gurobi_params.get("lazy_max_cuts")
        for _ in range(self.max_rounds):
            cons = [LinearConstraint(A, lo, hi)]
            if cut_rows:
                r = np.repeat(np.arange(len(cut_rows)), 4)
                cols = np.array([[pa, pb, pc, P + nr + na + t] for pa, pb, pc, t in cut_rows]).ravel()
                vals = np.tile([1.0, 1.0, 1.0, -1.0], len(cut_rows))
                cons.append(LinearConstraint(coo_matrix((vals, (r, cols)), shape=(len(cut_rows), n)).tocsr(), -np.inf, 2.0))
            res = milp(c, constraints=cons, integrality=integrality, bounds=bounds, options=opts)
            if res.x is None:
                status = f"status_{res.status}"
                break
            sol = res.x
            status = "optimal" if res.status == 0 else ("time_limit" if res.status == 1 else f"status_{res.status}")
            if separate is None or (lazy_max is not None and cuts_added >= lazy_max):
                break
            new = separate(np.asarray(sol[:P]), cuts_added)
            if len(new) == 0:
                break
            cut_rows.extend([tuple(int(v) for v in row) for row in new])
            cuts_added += len(new)
        rt = time.perf_counter() - t0
        if sol is None:
            return SolveResult(status, np.zeros(P), np.zeros(na), np.zeros(nr), np.zeros(T), rt, cuts_added)
        return SolveResult(status, sol[:P], sol[P + nr:P + nr + na], sol[P:P + nr], sol[P + nr + na:], rt, cuts_added)


class IncumbentBackend:
    """No solver: `incumbent(spec) -> x[P]` supplies ONE candidate solution, the separation callback fires once on it (as a
    MIPSOL event would) and q_t = 1 on every triangle that received a cut.  For timing and testing everything around the MIP
    (KNN, costs, tables, separation, post-solve, frames) on machines without a Gurobi licence — the same rule the golden
    fixtures were recorded with.  Never selected by default."""

    name = "incumbent"

    def __init__(self, incumbent):
        self.incumbent = incumbent

    def solve(self, spec: ModelSpec, separate: Optional[SeparationFn], gurobi_params: dict, outprefix=None, env_options=None, start=None):
        t0 = time.perf_counter()
        x = np.asarray(self.incumbent(spec), dtype=np.float64)
        q = np.zeros(spec.n_tri)
        cuts = 0
        if separate is not None:
            new = separate(x, 0)
            cuts = len(new)
            if cuts:
                q[np.asarray(new)[:, 3].astype(np.int64)] = 1.0
        rows_matched = np.add.reduceat(x > 0.5, spec.row_ptr[:-1].astype(np.int64)) > 0 if spec.n_pairs else np.zeros(spec.n_aligned, bool)
        return SolveResult("optimal", x, (~rows_matched).astype(np.float64), np.zeros(spec.n_ref), q, time.perf_counter() - t0, cuts)


_BACKENDS = {"gurobi": GurobiMatrixBackend, "gurobi_matrix": GurobiMatrixBackend, "gurobi_rows": GurobiBackend, "highs": HighsCutLoopBackend}
_default_backend = None


def set_default_backend(backend):
    """`'gurobi'` (default: the matrix builder), `'gurobi_rows'` (one quicksum per row, as the reference writes it), `'highs'`, or an object with a `.solve(spec, separate, gurobi_params, outprefix, env_options)` method."""
    global _default_backend
    _default_backend = backend


def get_backend(backend=None):
    b = backend if backend is not None else (_default_backend if _default_backend is not None else os.environ.get("SAME_B200_SOLVER", "gurobi"))
    if isinstance(b, str):
        return _BACKENDS[b]()
    return b
