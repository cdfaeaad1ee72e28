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

The CONSUMER of the halfspaces, the MPC safety filter's QP (core/mpc_filter.py:40-178: 2-D variables, `A @ x[t]`,
equality constraints, `quad_form`, `square`, `multiply`), is supported too, so that the downstream-trajectory parity of
BASELINE's north star can be checked: the QP is assembled from the reference's own constraint loops and solved by a
dense primal-dual interior-point method (`solve_qp` below, tolerance 1e-10) in place of cvxpy's default OSQP
(osqp==1.0.1, environment.yml:42).  `norm` / `sum_squares` remain unimplemented.
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

    def __rmatmul__(self, mat):          # numpy matrix @ affine vector (core/mpc_filter.py:84,96)
        m = np.asarray(mat, dtype=np.float64)
        if m.ndim != 2 or m.shape[1] != self.size:
            raise ValueError("shape mismatch in cvxpy shim matmul")
        out = []
        for k in range(m.shape[0]):
            acc = _Row()
            for j in range(m.shape[1]):
                if m[k, j] != 0.0:
                    acc = acc.plus(self.rows[j].scaled(float(m[k, j])))
            out.append(acc)
        return Expression(out)

    # ---- constraints:  lhs <= rhs  stored as (lhs - rhs) <= 0
    def __le__(self, other):
        return Constraint(self - other)

    def __ge__(self, other):
        return Constraint(Expression._lift(other) - self)

    def __eq__(self, other):             # equality constraints (MPC dynamics, core/mpc_filter.py:80-84)
        return Constraint(self - other, kind="eq")

    __hash__ = object.__hash__


class Variable(Expression):
    def __init__(self, shape=1, nonneg=False, name=None, **kw):
        self.shape2 = None
        if isinstance(shape, tuple):
            if len(shape) == 2:              # x[t] -> row t (core/mpc_filter.py:59-60)
                self.shape2 = (int(shape[0]), int(shape[1]))
                shape = self.shape2[0] * self.shape2[1]
            elif len(shape) == 1:
                shape = shape[0]
            elif len(shape) == 0:
                shape = 1
            else:
                raise NotImplementedError("cvxpy shim: only scalar, 1-D and 2-D variables")
        if kw:
            raise NotImplementedError(f"cvxpy shim: Variable attributes {list(kw)} not supported")
        self.n = int(shape)
        self.nonneg = bool(nonneg)
        self.name = name
        self._value = None
        super().__init__([_Row({(self, i): 1.0}) for i in range(self.n)])

    @property
    def value(self):
        if self._value is None or self.shape2 is None:
            return self._value
        return self._value.reshape(self.shape2)

    @value.setter
    def value(self, v):
        self._value = v

    def __getitem__(self, i):
        if self.shape2 is None:
            return super().__getitem__(i)
        if isinstance(i, (int, np.integer)):
            t = int(i)
            if t < 0:
                t += self.shape2[0]
            return Expression(self.rows[t * self.shape2[1]:(t + 1) * self.shape2[1]])
        raise NotImplementedError("cvxpy shim: 2-D variables support row indexing only")


class Parameter(Expression):
    def __init__(self, shape=1, **kw):
        if isinstance(shape, tuple):
            shape = shape[0]
        self.n = int(shape)
        self.value = None
        super().__init__([_Row(params=[(1.0, self, i)]) for i in range(self.n)])


class Constraint:
    def __init__(self, expr_le_zero, kind="le"):
        self.expr = expr_le_zero     # kind "le": expr <= 0;  kind "eq": expr == 0
        self.kind = kind


class QuadExpr:
    """Scalar convex objective: sum_k coef_k * e_k^T Q_k e_k (e_k affine vectors) + an affine scalar."""
    __array_ufunc__ = None

    def __init__(self, quads=None, affine=None):
        self.quads = quads or []                 # (coef, rows, Q)
        self.affine = affine or _Row()

    def _plus(self, other):
        if isinstance(other, QuadExpr):
            return QuadExpr(self.quads + other.quads, self.affine.plus(other.affine))
        if isinstance(other, Expression):
            if other.size != 1:
                raise ValueError("objective terms must be scalar")
            return QuadExpr(list(self.quads), self.affine.plus(other.rows[0]))
        if _is_number(other):
            return QuadExpr(list(self.quads), self.affine.plus(_Row(const=float(other))))
        return NotImplemented

    __add__ = _plus
    __radd__ = _plus

    def __mul__(self, c):
        if not _is_number(c) or c < 0:
            raise NotImplementedError("cvxpy shim: only nonnegative scalar * quadratic")
        return QuadExpr([(q[0] * float(c), q[1], q[2]) for q in self.quads], self.affine.scaled(float(c)))

    __rmul__ = __mul__


class Minimize:
    def __init__(self, expr):
        self.quad = None
        if isinstance(expr, QuadExpr):
            self.quad = expr
            expr = Expression([expr.affine])
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


def quad_form(expr, Q):            # e^T Q e  (core/mpc_filter.py:70,73)
    expr = Expression._lift(expr)
    Q = np.asarray(Q, dtype=np.float64)
    if Q.shape != (expr.size, expr.size):
        raise ValueError("quad_form shape mismatch")
    return QuadExpr([(1.0, list(expr.rows), 0.5 * (Q + Q.T))])


def square(expr):                  # core/mpc_filter.py:146
    expr = Expression._lift(expr)
    if expr.size != 1:
        raise NotImplementedError("cvxpy shim: square of a scalar only")
    return QuadExpr([(1.0, list(expr.rows), np.ones((1, 1)))])


def multiply(a, expr):             # elementwise constant * affine (core/mpc_filter.py:139)
    expr = Expression._lift(expr)
    arr = np.asarray(a, dtype=np.float64).ravel()
    if arr.size != expr.size:
        raise ValueError("multiply shape mismatch")
    return Expression([r.scaled(float(c)) for r, c in zip(expr.rows, arr)])


def solve_qp(P, q, G, h, A, b, tol=1e-10, max_iter=200):
    """min 0.5 z'Pz + q'z  s.t.  Gz <= h, Az = b  — dense Mehrotra predictor-corrector interior-point method."""
    n, m, p = q.shape[0], h.shape[0], b.shape[0]
    z, y = np.zeros(n), np.zeros(p)
    s, lam = np.ones(m), np.ones(m)
    if m:
        s = np.maximum(h - G @ z, 1.0)
    scale = 1.0 + max(np.abs(q).max(initial=0.0), np.abs(h).max(initial=0.0), np.abs(b).max(initial=0.0))
    for it in range(max_iter):
        rd = P @ z + q + (G.T @ lam if m else 0.0) + (A.T @ y if p else 0.0)
        rp = G @ z + s - h if m else np.zeros(0)
        re = A @ z - b if p else np.zeros(0)
        mu = float(s @ lam) / m if m else 0.0
        if max(np.abs(rd).max(initial=0.0), np.abs(rp).max(initial=0.0), np.abs(re).max(initial=0.0)) <= tol * scale \
                and mu <= tol:
            return z, True, it
        W = lam / s if m else np.zeros(0)
        H = P + (G.T @ (W[:, None] * G) if m else 0.0) + 1e-13 * np.eye(n)
        K = np.block([[H, A.T], [A, -1e-13 * np.eye(p)]]) if p else H

        def direction(rc):
            rhs = -rd - (G.T @ ((lam * rp - rc) / s) if m else 0.0)
            sol = np.linalg.solve(K, np.concatenate([rhs, -re]) if p else rhs)
            dz, dy = sol[:n], sol[n:]
            ds = -rp - G @ dz if m else np.zeros(0)
            dlam = (-rc - lam * ds) / s if m else np.zeros(0)
            return dz, dy, ds, dlam

        def step_len(v, dv):
            neg = dv < 0
            return min(1.0, float(np.min(-v[neg] / dv[neg]))) if neg.any() else 1.0

        dz, dy, ds, dlam = direction(s * lam)
        if m:
            a_aff = min(step_len(s, ds), step_len(lam, dlam))
            mu_aff = float((s + a_aff * ds) @ (lam + a_aff * dlam)) / m
            sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
            dz, dy, ds, dlam = direction(s * lam + ds * dlam - sigma * mu)
            a = 0.995 * min(step_len(s, ds), step_len(lam, dlam))
            a = min(a, 1.0)
            s = s + a * ds
            lam = lam + a * dlam
        else:
            a = 1.0
        z = z + a * dz
        y = y + a * dy
    return z, False, max_iter


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
        if objective.quad is not None:
            for _c, qrows, _Q in objective.quad.quads:
                visit(Expression(qrows))
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
        self._row_kind = []
        for c in self.constraints:
            for r in c.expr.rows:
                self._row_kind.append(c.kind)
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

    def _solve_qp(self):
        n = self._ncols
        P = np.zeros((n, n))
        q = self._c.copy()
        const = self.objective.expr.rows[0].constant_value()
        for coef, qrows, Q in self.objective.quad.quads:      # coef * (M z + c)' Q (M z + c)
            M = np.zeros((len(qrows), n))
            c0 = np.zeros(len(qrows))
            for k, r in enumerate(qrows):
                for (v, i), cf in r.vars.items():
                    M[k, self._offset[id(v)] + i] += cf
                c0[k] = r.constant_value()
            P += 2.0 * coef * (M.T @ Q @ M)
            q += 2.0 * coef * (M.T @ (Q @ c0))
            const += coef * float(c0 @ Q @ c0)
        Aall = self._A.toarray()
        rhs = np.array([-r.constant_value() for r in self._row_forms], dtype=np.float64)
        eq = np.array([k == "eq" for k in self._row_kind], dtype=bool)
        G, h, A, b = Aall[~eq], rhs[~eq], Aall[eq], rhs[eq]
        nn = [self._offset[id(v)] + i for v in self._vars if v.nonneg for i in range(v.n)]
        if nn:
            E = np.zeros((len(nn), n))
            E[np.arange(len(nn)), nn] = -1.0
            G, h = np.vstack([G, E]), np.concatenate([h, np.zeros(len(nn))])
        z, ok, _it = solve_qp(P, q, G, h, A, b)
        if ok:
            self.status = OPTIMAL
            for v in self._vars:
                o = self._offset[id(v)]
                v.value = np.array(z[o:o + v.n], dtype=np.float64)
            self.value = float(0.5 * z @ P @ z + q @ z + const)
        else:
            self.status = "solver_error"
            self.value = None
        return self.value

    def solve(self, solver=None, **kw):
        if self.objective.quad is not None and self.objective.quad.quads:
            return self._solve_qp()
        if any(k == "eq" for k in self._row_kind):
            raise NotImplementedError("cvxpy shim: equality constraints are only supported on the QP path")
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
