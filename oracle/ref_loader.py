"""Import the UNMODIFIED reference (`/root/reference/src`) under stub modules.

TEST INFRASTRUCTURE ONLY.  Nothing under `same_b200/` may import this file.  It
exists so that `tests/golden/gen_golden.py` can run the reference's own Python
in the build container and freeze its outputs as golden vectors.  The reference
cannot travel to the GPU box (`/root/reference` does not exist there), so no
`-m gpu` test, `smoke()` or `bench.py` run imports this module.

The reference needs five third-party packages that are absent here
(`gurobipy`, `scanpy`, `matplotlib`, `shapely`, `alphashape`).
None of them is on the hot path except `gurobipy`, which is replaced by a
*recording fake*: it implements just enough of the modelling API
(`Model.addVars/addVar/addConstr/setObjective/optimize/cbGetSolution/cbLazy`)
for `run_same` (`src/same.py:706-1489`) to execute end to end, records every
variable, constraint, objective coefficient and lazy cut in creation order, and
"solves" the model by asking a user-supplied incumbent function.
"""
from __future__ import annotations

import importlib
import sys
import types

REFERENCE_ROOT = "/root/reference"


# --------------------------------------------------------------------------
# recording fake gurobipy
# --------------------------------------------------------------------------
class LinExpr:
    """Sparse linear expression: {var_index: coef} + constant."""

    __slots__ = ("terms", "const")

    def __init__(self, terms=None, const=0.0):
        self.terms = dict(terms) if terms else {}
        self.const = const

    @staticmethod
    def _lift(o):
        if isinstance(o, LinExpr):
            return o
        if isinstance(o, Var):
            return LinExpr({o.index: 1.0})
        return LinExpr(None, o)

    def copy(self):
        return LinExpr(self.terms, self.const)

    def _iadd(self, o, s=1.0):
        o = LinExpr._lift(o)
        for k, v in o.terms.items():
            self.terms[k] = self.terms.get(k, 0.0) + s * v
        self.const = self.const + s * o.const
        return self

    def __add__(self, o):
        return self.copy()._iadd(o)

    __radd__ = __add__

    def __sub__(self, o):
        return self.copy()._iadd(o, -1.0)

    def __rsub__(self, o):
        return LinExpr._lift(o).copy()._iadd(self, -1.0)

    def __mul__(self, c):
        return LinExpr({k: v * c for k, v in self.terms.items()}, self.const * c)

    __rmul__ = __mul__

    def __le__(self, o):
        return TempConstr(self - o, "<=")

    def __ge__(self, o):
        return TempConstr(self - o, ">=")

    def __eq__(self, o):  # noqa: D105
        return TempConstr(self - o, "==")

    __hash__ = None


class TempConstr:
    __slots__ = ("expr", "sense")

    def __init__(self, expr, sense):
        self.expr, self.sense = expr, sense


class Var:
    """Stand-in for gurobipy.Var; `.x` is filled by FakeModel.optimize."""

    def __init__(self, model, index, name, vtype, lb, ub):
        self.model, self.index, self.VarName = model, index, name
        self.vtype, self.lb, self.ub = vtype, lb, ub
        self.x = 0.0
        self.Start = None

    def __hash__(self):
        return self.index

    def __add__(self, o):
        return LinExpr._lift(self) + o

    __radd__ = __add__

    def __sub__(self, o):
        return LinExpr._lift(self) - o

    def __rsub__(self, o):
        return LinExpr._lift(o) - LinExpr._lift(self)

    def __mul__(self, c):
        return LinExpr({self.index: c})

    __rmul__ = __mul__

    def __le__(self, o):
        return LinExpr._lift(self) <= o

    def __ge__(self, o):
        return LinExpr._lift(self) >= o

    def __eq__(self, o):  # noqa: D105
        if isinstance(o, Var):
            return self is o
        return LinExpr._lift(self) == o


class tupledict(dict):
    def values(self):  # gurobipy returns a list
        return list(super().values())


def quicksum(it):
    e = LinExpr()
    for t in it:
        e._iadd(t)
    return e


class _Params:
    def __setattr__(self, k, v):
        object.__setattr__(self, k, v)


class _Callback:
    MIPSOL = 4


class MVar:
    """Stand-in for gurobipy.MVar (1-D): what same_b200.solver.GurobiMatrixBackend uses of the matrix API."""

    def __init__(self, vs):
        self._vars = list(vs)

    @staticmethod
    def fromlist(vs):
        return MVar(vs)

    def tolist(self):
        return list(self._vars)

    def __len__(self):
        return len(self._vars)

    @property
    def X(self):
        import numpy as np
        return np.array([v.x for v in self._vars], dtype=float)

    def _set_start(self, values):
        for v, val in zip(self._vars, values):
            v.Start = float(val)

    Start = property(lambda self: [v.Start for v in self._vars], _set_start)

    def __rmatmul__(self, coeffs):          # ndarray @ MVar -> linear expression
        e = LinExpr()
        for c, v in zip(coeffs, self._vars):
            e.terms[v.index] = e.terms.get(v.index, 0.0) + float(c)
        return e

    __array_ufunc__ = None                  # let `ndarray @ MVar` reach __rmatmul__


class _Constr:
    """Handle of one recorded constraint; setting ConstrName renames the record (gurobipy.Constr.ConstrName)."""

    def __init__(self, model, pos):
        self._model, self._pos = model, pos

    def _get(self):
        return self._model.constrs[self._pos][0]

    def _set(self, name):
        rec = self._model.constrs[self._pos]
        self._model.constrs[self._pos] = (name,) + tuple(rec[1:])

    ConstrName = property(_get, _set)


class MConstr:
    def __init__(self, cs):
        self._cs = cs

    def tolist(self):
        return list(self._cs)


class GRB:
    LESS_EQUAL, EQUAL, GREATER_EQUAL = "<", "=", ">"
    BINARY, CONTINUOUS, INTEGER = "B", "C", "I"
    MINIMIZE, MAXIMIZE = 1, -1
    OPTIMAL, TIME_LIMIT, INFEASIBLE = 2, 9, 3
    METHOD_PDHG = 6
    INFINITY = 1e100
    Callback = _Callback


class Env:
    def __init__(self, *a, **k):
        self.params = k.get("params")


#: set by the test/golden script: f(model) -> sequence of x values (len = n "x" vars)
INCUMBENT_FN = None
#: the last FakeModel built (so callers can inspect the recording)
LAST_MODEL = None
#: all models built since the list was last cleared
MODELS = []


class Model:
    def __init__(self, name="", env=None):
        global LAST_MODEL
        self.name = name
        self.vars = []
        self.constrs = []  # (name, sense, terms dict, rhs)
        self.lazy = []  # (terms dict, sense, rhs)
        self.objective = None
        self.Params = _Params()
        self.status = GRB.OPTIMAL
        self.Runtime = 0.0
        self._sol = None
        self.callback_calls = 0
        LAST_MODEL = self
        MODELS.append(self)

    def addVar(self, lb=0.0, ub=GRB.INFINITY, obj=0.0, vtype=GRB.CONTINUOUS, name=""):
        v = Var(self, len(self.vars), name, vtype, lb, ub)
        self.vars.append(v)
        return v

    def addVars(self, *idx, lb=0.0, ub=GRB.INFINITY, obj=0.0, vtype=GRB.CONTINUOUS, name=""):
        n = idx[0]
        keys = range(n) if isinstance(n, int) else list(n)
        td = tupledict()
        for k in keys:
            td[k] = self.addVar(lb=lb, ub=ub, vtype=vtype, name=f"{name}[{k}]")
        return td

    def addConstr(self, tc, name=""):
        self.constrs.append((name, tc.sense, dict(tc.expr.terms), -tc.expr.const))
        return len(self.constrs) - 1

    def addConstrs(self, gen, name=""):
        return [self.addConstr(tc, name) for tc in gen]

    def addMVar(self, shape, lb=0.0, ub=GRB.INFINITY, obj=0.0, vtype=GRB.CONTINUOUS, name=""):
        n = int(shape if not isinstance(shape, (tuple, list)) else shape[0])
        return MVar([self.addVar(lb=lb, ub=ub, vtype=vtype, name=f"{name}[{k}]") for k in range(n)])

    def addMConstr(self, A, x, sense, b, name=""):
        """Rows of the sparse matrix A (scipy CSR) over the variables of `x` (None = all variables, in order)."""
        vs = self.vars if x is None else x.tolist()
        A = A.tocsr()
        words = {"<": "<=", "=": "==", ">": ">="}
        out = []
        for r in range(A.shape[0]):
            lo, hi = A.indptr[r], A.indptr[r + 1]
            terms = {vs[int(c)].index: float(v) for c, v in zip(A.indices[lo:hi], A.data[lo:hi])}
            sr = sense if isinstance(sense, str) else sense[r]
            self.constrs.append((f"{name}[{r}]" if name else "", words[str(sr)], terms, float(b[r])))
            out.append(_Constr(self, len(self.constrs) - 1))
        return MConstr(out)

    def setObjective(self, expr, sense=GRB.MINIMIZE):
        self.objective = (LinExpr._lift(expr), sense)

    def update(self):
        pass

    def write(self, path):
        pass

    def cbGetSolution(self, vs):
        if isinstance(vs, dict):
            return {k: self._sol[v.index] for k, v in vs.items()}
        if isinstance(vs, Var):
            return self._sol[vs.index]
        return [self._sol[v.index] for v in vs]

    def cbLazy(self, tc):
        self.lazy.append((dict(tc.expr.terms), tc.sense, -tc.expr.const))

    def optimize(self, callback=None):
        """One incumbent from INCUMBENT_FN, one MIPSOL callback, accept."""
        xs = [v for v in self.vars if v.VarName.startswith("x[")]
        vals = INCUMBENT_FN(self) if INCUMBENT_FN is not None else [0.0] * len(xs)
        self._sol = [0.0] * len(self.vars)
        for v, val in zip(xs, vals):
            self._sol[v.index] = float(val)
        if callback is not None:
            self.callback_calls += 1
            callback(self, GRB.Callback.MIPSOL)
        for v in self.vars:
            v.x = self._sol[v.index]
        self.status = GRB.OPTIMAL


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _any_callable(name):
    """Module-level __getattr__ of the plotting stubs: any public name is a no-op callable; dunder lookups (`__file__`, `__path__`,
    ...) fail as on a real module, so that tools which walk sys.modules (hypothesis, importlib) are not handed a function."""
    if name.startswith("__") and name.endswith("__"):
        raise AttributeError(name)
    return lambda *a, **k: None


def install_stubs():
    g = _stub("gurobipy", Model=Model, GRB=GRB, quicksum=quicksum, Env=Env, LinExpr=LinExpr, Var=Var, MVar=MVar)
    g.tupledict = tupledict
    for name in ("scanpy", "alphashape"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name, alphashape=lambda *a, **k: None)
    try:
        importlib.import_module("matplotlib.pyplot")
    except Exception:
        mp = _stub("matplotlib")
        for sub in ("pyplot", "patches", "colors", "gridspec", "lines", "cm", "collections"):
            s = _stub(f"matplotlib.{sub}")
            setattr(mp, sub, s)
            s.__getattr__ = _any_callable  # type: ignore[attr-defined]
        mp.__getattr__ = _any_callable  # type: ignore[attr-defined]
    try:
        importlib.import_module("shapely.geometry")
    except Exception:
        sh = _stub("shapely")
        geo = _stub("shapely.geometry", MultiPolygon=object, Polygon=object, GeometryCollection=object,
                    Point=object, MultiPoint=object)
        sh.geometry = geo
        ops = _stub("shapely.ops", unary_union=lambda *a, **k: None)
        sh.ops = ops


def load_reference(root: str = REFERENCE_ROOT):
    """Return the reference package (`src`) imported from `root` under the stubs."""
    install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    return importlib.import_module("src")


if __name__ == "__main__":
    ref = load_reference()
    print("reference imported:", sorted(ref.__all__))
