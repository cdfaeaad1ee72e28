"""
ORACLE — test infrastructure, NOT product code.

Independent restatement of the reference's two linear programs, solved with HiGHS
(scipy.optimize.linprog) instead of cvxpy->ECOS (which is not installed here).  Used only to
cross-check oracle/closed_form.py; the row layout follows the reference line by line:

  DR-CVaR LP  core/risk_metrics.py:94-125   vars z = [g, tau, lambda, eta_1..eta_N]
      lambda*eps + (1/N) sum(eta) <= delta                                   (:110)
      a_k*hxi_i + b_k*(g - r) + c_k*tau <= eta_i,  k = 0,1                    (:113-119)
          a = [-1/alpha, 0], b = [-1/alpha, 0], c = [1 - 1/alpha, 1]          (:105-107)
      1/alpha <= lambda, lambda >= 0                                          (:97,122)
      minimise g                                                              (:125)
  CVaR LP     core/risk_metrics.py:188-213  vars z = [g, tau, aux_1..aux_N]
      aux >= 0 ; aux_i >= -hxi_i - g + r - tau                                (:198-205)
      tau + (1/(alpha N)) sum(aux) <= delta                                   (:208-210)
      minimise g                                                              (:213)
"""
from __future__ import annotations

import numpy as np
from scipy import sparse
from scipy.optimize import linprog


def _solve(c, A, b, bounds):
    res = linprog(c, A_ub=A, b_ub=b, bounds=bounds, method="highs")
    if res.status != 0:
        return False, 100.0
    return True, float(res.x[0])


def drcvar_lp(h_xi, r, alpha, epsilon, delta):
    """Returns (solved, g_star) for projections h_xi = h @ samples.T and radius term r."""
    h_xi = np.asarray(h_xi, dtype=np.float64)
    n = h_xi.shape[0]
    nv = 3 + n
    ia = 1.0 / alpha
    rows, cols, vals, rhs = [], [], [], []
    # row 0: lambda*eps + (1/N) sum eta <= delta
    rows += [0] * (n + 1)
    cols += [2] + list(range(3, nv))
    vals += [epsilon] + [1.0 / n] * n
    rhs.append(delta)
    # k = 0: -ia*hxi_i - ia*(g - r) + (1 - ia)*tau - eta_i <= 0
    base = 1
    for i in range(n):
        rr = base + i
        rows += [rr, rr, rr]
        cols += [0, 1, 3 + i]
        vals += [-ia, 1.0 - ia, -1.0]
        rhs.append(ia * h_xi[i] - ia * r)
    # k = 1: tau - eta_i <= 0
    base = 1 + n
    for i in range(n):
        rr = base + i
        rows += [rr, rr]
        cols += [1, 3 + i]
        vals += [1.0, -1.0]
        rhs.append(0.0)
    A = sparse.csr_matrix((vals, (rows, cols)), shape=(1 + 2 * n, nv))
    c = np.zeros(nv)
    c[0] = 1.0
    bounds = [(None, None), (None, None), (ia, None)] + [(None, None)] * n
    return _solve(c, A, np.asarray(rhs), bounds)


def cvar_lp(h_xi, r, alpha, delta):
    """Returns (solved, g) for the CVaR LP."""
    h_xi = np.asarray(h_xi, dtype=np.float64)
    n = h_xi.shape[0]
    nv = 2 + n
    rows, cols, vals, rhs = [], [], [], []
    # -g - tau - aux_i <= hxi_i - r
    for i in range(n):
        rows += [i, i, i]
        cols += [0, 1, 2 + i]
        vals += [-1.0, -1.0, -1.0]
        rhs.append(h_xi[i] - r)
    rows += [n] * (n + 1)
    cols += [1] + list(range(2, nv))
    vals += [1.0] + [1.0 / (alpha * n)] * n
    rhs.append(delta)
    A = sparse.csr_matrix((vals, (rows, cols)), shape=(n + 1, nv))
    c = np.zeros(nv)
    c[0] = 1.0
    bounds = [(None, None), (None, None)] + [(0.0, None)] * n
    return _solve(c, A, np.asarray(rhs), bounds)
