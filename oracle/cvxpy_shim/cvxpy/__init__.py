"""
ORACLE — test infrastructure, NOT product code.

A deliberately tiny stand-in for the `cvxpy` package (cvxpy==1.2.1 is the reference's pinned
dependency, environment.yml:31; it is NOT installed in this image and there is no network).
It implements just enough of the modelling API — Variable, Parameter, affine arithmetic,
`<=` / `>=`, `cp.sum`, Minimize, Problem.solve — for the reference's OWN, UNMODIFIED
`DRCVaROptimizer` / `CVaROptimizer` (core/risk_metrics.py:84-265) to build their linear
programs with their own constraint loops.  The LP is then solved by HiGHS
(scipy.optimize.linprog) in place of ECOS (ecos==2.0.14, environment.yml:33).

So: LP construction = the reference's code, verbatim at run time; numerical solver = HiGHS.
Anything non-linear (quad_form, square, ... used by the MPC QP, core/mpc_filter.py) raises
NotImplementedError: the MPC consumer is outside this repo's scope.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse
from scipy.optimize import linprog

OPTIMAL = "optimal"
OPTIMAL_INACCURATE = "optimal_inaccurate"
INFEASIBLE = "infeasible"
UNBOUNDED = "unbounded"

__version__ = "0.0-shim"


class _Row:
    """One scalar affine form: sum(coef*var_elem) + const + sum(coef*param_elem)."""
    __slots__ = ("vars", "const", "params")

    def __init__(self, vars=None, const=0.0, params=None):
        self.vars = vars or {}      # (Variable, idx) -> coef
        self.const = const
        self.params = params or []  # (coef, Parameter, idx)

    def scaled(self, c):
        return _Row({k: v * c for k, v in self.vars.items()}, self.const * c,
                    [(pc * c, p, i) for pc, p, i in self.params])

    def plus(self, other, sign=1.0):
        v = dict(self.vars)
        for k, c in other.vars.items():
            v[k] = v.get(k, 0.0) + sign * c
        return _Row(v, self.const + sign * other.const,
                    self.params + [(sign * pc, p, i) for pc, p, i in other.params])

    def constant_value(self):
        c = self.const
        for pc, p, i in self.params:
            if p.value is None:
                raise ValueError("Parameter value not set")
            c += pc * float(np.asarray(p.value, dtype=np.float64).ravel()[i])
        return c


def _is_number(x):
    return isinstance(x, (int, float, np.integer, np.floating))


class Expression:
    __array_ufunc__ = None  # numpy scalars/arrays defer to our reflected operators

    def __init__(self, rows):
        self.rows = rows

    # ---- shape helpers
    @property
    def size(self):
        return len(self.rows)

    @property
    def shape(self):
        return (len(self.rows),)

    @staticmethod
    def _lift(x, n=None):
        if isinstance(x, Expression):
            return x
        arr = np.asarray(x, dtype=np.float64).ravel()
        return Expression([_Row(const=float(a)) for a in arr])

    def _bcast(self, other):
        a, b = self, Expression._lift(other)
        if a.size == b.size:
            return a.rows, b.rows
        if a.size == 1:
            return a.rows * b.size, b.rows
        if b.size == 1:
            return a.rows, b.rows * a.size
        raise ValueError("shape mismatch in cvxpy shim")

    # ---- arithmetic
    def __add__(self, other):
        ra, rb = self._bcast(other)
        return Expression([x.plus(y) for x, y in zip(ra, rb)])

    __radd__ = __add__

    def __sub__(self, other):
        ra, rb = self._bcast(other)
        return Expression([x.plus(y, -1.0) for x, y in zip(ra, rb)])

    def __rsub__(self, other):
        return Expression._lift(other).__sub__(self)

    def __neg__(self):
        return Expression([r.scaled(-1.0) for r in self.rows])

    def __mul__(self, other):
        if not _is_number(other):
            raise NotImplementedError("cvxpy shim: only scalar * affine is supported")
        return Expression([r.scaled(float(other)) for r in self.rows])

    __rmul__ = __mul__

    def __truediv__(self, other):
        if not _is_number(other):
            raise NotImplementedError("cvxpy shim: only affine / scalar is supported")
        return self * (1.0 / float(other))

    def __getitem__(self, i):
        if isinstance(i, slice):
            return Expression(self.rows[i])
        return Expression([self.rows[i]])

    # ---- constraints:  lhs <= rhs  stored as (lhs - rhs) <= 0
    def __le__(self, other):
        return Constraint(self - other)

    def __ge__(self, other):
        return Constraint(Expression._lift(other) - self)

    def __eq__(self, other):  # pragma: no cover - not used by the LPs
        raise NotImplementedError("cvxpy shim: equality constraints not supported")

    __hash__ = object.__hash__


class Variable(Expression):
    def __init__(self, shape=1, nonneg=False, name=None, **kw):
        if isinstance(shape, tuple):
            if len(shape) != 1:
                raise NotImplementedError("cvxpy shim: only 1-D variables")
            shape = shape[0]
        if kw:
            raise NotImplementedError(f"cvxpy shim: Variable attributes {list(kw)} not supported")
        self.n = int(shape)
        self.nonneg = bool(nonneg)
        self.name = name
        self.value = None
        super().__init__([_Row({(self, i): 1.0}) for i in range(self.n)])


class Parameter(Expression):
    def __init__(self, shape=1, **kw):
        if isinstance(shape, tuple):
            shape = shape[0]
        self.n = int(shape)
        self.value = None
        super().__init__([_Row(params=[(1.0, self, i)]) for i in range(self.n)])


class Constraint:
    def __init__(self, expr_le_zero):
        self.expr = expr_le_zero


class Minimize:
    def __init__(self, expr):
        expr = Expression._lift(expr)
        if expr.size != 1:
            raise ValueError("objective must be scalar")
        self.expr = expr


def sum(expr):  # noqa: A001 - mirrors cvxpy.sum
    acc = _Row()
    for r in expr.rows:
        acc = acc.plus(r)
    return Expression([acc])


def _unsupported(name):
    def f(*a, **k):
        raise NotImplementedError(f"cvxpy shim: {name} is not implemented (LPs only)")
    return f


quad_form = _unsupported("quad_form")
square = _unsupported("square")
multiply = _unsupported("multiply")
norm = _unsupported("norm")
sum_squares = _unsupported("sum_squares")


class Problem:
    def __init__(self, objective, constraints=()):
        self.objective = objective
        self.constraints = list(constraints)
        self.status = None
        self.value = None
        # collect variables, assign columns
        self._vars = []
        seen = set()

        def visit(expr):
            for r in expr.rows:
                for (v, _i) in r.vars:
                    if id(v) not in seen:
                        seen.add(id(v))
                        self._vars.append(v)

        visit(objective.expr)
        for c in self.constraints:
            visit(c.expr)
        self._offset = {}
        off = 0
        for v in self._vars:
            self._offset[id(v)] = off
            off += v.n
        self._ncols = off
        # constant (variable) part of A, built once like a DPP problem
        rows, cols, vals = [], [], []
        self._row_forms = []
        for c in self.constraints:
            for r in c.expr.rows:
                ri = len(self._row_forms)
                for (v, i), coef in r.vars.items():
                    if coef != 0.0:
                        rows.append(ri)
                        cols.append(self._offset[id(v)] + i)
                        vals.append(coef)
                self._row_forms.append(r)
        self._A = sparse.csr_matrix((vals, (rows, cols)), shape=(len(self._row_forms), self._ncols))
        self._c = np.zeros(self._ncols)
        for (v, i), coef in objective.expr.rows[0].vars.items():
            self._c[self._offset[id(v)] + i] += coef
        self._bounds = []
        for v in self._vars:
            self._bounds += [((0.0 if v.nonneg else None), None)] * v.n

    def solve(self, solver=None, **kw):
        b = np.array([-r.constant_value() for r in self._row_forms], dtype=np.float64)
        res = linprog(self._c, A_ub=self._A, b_ub=b, bounds=self._bounds, method="highs")
        if res.status == 0:
            self.status = OPTIMAL
            for v in self._vars:
                o = self._offset[id(v)]
                v.value = np.array(res.x[o:o + v.n], dtype=np.float64)
            self.value = float(res.fun + self.objective.expr.rows[0].constant_value())
        elif res.status == 2:
            self.status = INFEASIBLE
            self.value = np.inf
        elif res.status == 3:
            self.status = UNBOUNDED
            self.value = -np.inf
        else:
            self.status = "solver_error"
            self.value = None
        return self.value
