"""
Drop-in for the reference's core/mpc_filter.py (MPCSafetyFilter, reference :9-218) — the CONSUMER of the safe halfspaces
(SURVEY.md §8-f4).  Same constructor, attributes, `filter_trajectory` signature, return triple, printed lines and fallback
behaviour; what changes is how the QP is solved: the reference rebuilds a cvxpy problem on every call and hands it to
cvxpy's default solver (OSQP, environment.yml:42); here the same QP is assembled directly as dense matrices and solved by
a primal-dual interior-point method in numpy (residual tolerance 1e-10), so the filter runs where cvxpy is not installed
and returns the optimum to solver-independent accuracy.

The QP (reference :58-149), z = [x (H+1 x n), u (H x m), one slack per (step, halfspace)]:
    min  sum_t (x[t+1]-x_ref[t+1])' Q (x[t+1]-x_ref[t+1]) + u[t]' R u[t]  +  sum_slacks 50 s + 50 s^2
    s.t. x[0] = x0;  x[t+1] = A x[t] + B u[t];  u_min <= u[t] <= u_max;  pos_min <= C x[t] <= pos_max  (t >= 1)
         h . (C x[t]) + g <= s,  s >= 0   for every halfspace of safe_halfspaces[t-1]   (reference :116-146: the halfspace
         of obstacle time t-1 constrains the ego at time t — quirk preserved)
This module runs on the host (as the reference's does): the hot path of this repo is the halfspace computation.
"""
import time

import numpy as np

try:   # the reference's own helpers when this overlay sits on a reference checkout
    from utils.timing import timeit, Timer
except ImportError:   # standalone: dropin/_stopwatch.py
    from _stopwatch import timeit, Timer

OPTIMAL = "optimal"


def _solve_qp(P, q, G, h, A, b, tol=1e-10, max_iter=200):
    """min 0.5 z'Pz + q'z  s.t. Gz <= h, Az = b.  Infeasible-start Mehrotra predictor-corrector; returns (z, converged)."""
    n, m, p = q.size, h.size, b.size
    z = np.zeros(n)
    nu = np.zeros(p)
    slack = np.maximum(h - G @ z, 1.0)
    dual = np.ones(m)
    scale = 1.0 + max(np.abs(q).max(), np.abs(h).max() if m else 0.0, np.abs(b).max() if p else 0.0)
    KKT = np.zeros((n + p, n + p))
    KKT[:n, n:] = A.T
    KKT[n:, :n] = A
    reg = 1e-13
    for _ in range(max_iter):
        r_dual = P @ z + q + G.T @ dual + A.T @ nu
        r_ineq = G @ z + slack - h
        r_eq = A @ z - b
        gap = float(slack @ dual) / max(m, 1)
        if max(np.abs(r_dual).max(), np.abs(r_ineq).max() if m else 0.0, np.abs(r_eq).max() if p else 0.0) <= tol * scale \
                and gap <= tol:
            return z, True
        w = dual / slack
        KKT[:n, :n] = P + G.T @ (w[:, None] * G)
        KKT[np.arange(n), np.arange(n)] += reg
        KKT[np.arange(n, n + p), np.arange(n, n + p)] = -reg

        def newton(comp):   # comp = right-hand side of the linearised complementarity  slack*d_dual + dual*d_slack = -comp
            rhs = np.concatenate([-r_dual - G.T @ ((dual * r_ineq - comp) / slack), -r_eq])
            sol = np.linalg.solve(KKT, rhs)
            dz = sol[:n]
            d_slack = -r_ineq - G @ dz
            return dz, sol[n:], d_slack, (-comp - dual * d_slack) / slack

        def max_step(v, dv):
            neg = dv < 0
            return min(1.0, float((-v[neg] / dv[neg]).min())) if neg.any() else 1.0

        dz, dnu, d_slack, d_dual = newton(slack * dual)                      # predictor
        a_aff = min(max_step(slack, d_slack), max_step(dual, d_dual))
        gap_aff = float((slack + a_aff * d_slack) @ (dual + a_aff * d_dual)) / max(m, 1)
        sigma = (gap_aff / gap) ** 3 if gap > 0 else 0.0
        dz, dnu, d_slack, d_dual = newton(slack * dual + d_slack * d_dual - sigma * gap)   # corrector
        a = min(1.0, 0.995 * min(max_step(slack, d_slack), max_step(dual, d_dual)))
        z = z + a * dz
        nu = nu + a * dnu
        slack = slack + a * d_slack
        dual = dual + a * d_dual
    return z, False


class MPCSafetyFilter:
    """MPC-based safety filter for collision avoidance (reference core/mpc_filter.py:9-37)."""

    def __init__(self, A, B, C, Q, R, horizon, dt):
        self.A = A
        self.B = B
        self.C = C
        self.Q = Q
        self.R = R
        self.horizon = horizon
        self.dt = dt
        self.n_states = A.shape[0]
        self.n_inputs = B.shape[1]
        self.n_outputs = C.shape[0]
        self.last_optimal_u = None

    @timeit
    def filter_trajectory(self, x0, x_ref, u_ref, safe_halfspaces, input_constraints=None, position_constraints=None):
        """Reference core/mpc_filter.py:40-178: returns (x_filtered [H+1,n], u_filtered [H,m], info)."""
        start_time = time.time()
        H, n, m, no = self.horizon, self.n_states, self.n_inputs, self.n_outputs
        A, B, C = np.asarray(self.A, float), np.asarray(self.B, float), np.asarray(self.C, float)
        Q = 0.5 * (np.asarray(self.Q, float) + np.asarray(self.Q, float).T)
        R = 0.5 * (np.asarray(self.R, float) + np.asarray(self.R, float).T)
        x_ref = np.asarray(x_ref, float)
        x0 = np.asarray(x0, float)

        # halfspaces of step t (1..H), with the reference's debug lines (reference :122-128)
        hs_rows = []
        for t in range(1, H + 1):
            if t - 1 < len(safe_halfspaces):
                halfspaces_t = safe_halfspaces[t - 1]
                print(f"Halfspace constraints at step {t}: {len(halfspaces_t)}")
                for i, halfspace in enumerate(halfspaces_t):
                    hh, gg = halfspace.get_constraint_params()
                    print(f"Halfspace {i}: h={hh}, g={gg}")
                for halfspace in halfspaces_t:
                    hh, gg = halfspace.get_constraint_params()
                    hs_rows.append((t, np.asarray(hh, float).ravel(), float(gg)))
        n_s = len(hs_rows)
        ox, ou, os_ = 0, (H + 1) * n, (H + 1) * n + H * m
        nz = os_ + n_s

        def xs(t):
            return slice(ox + t * n, ox + (t + 1) * n)

        def us(t):
            return slice(ou + t * m, ou + (t + 1) * m)

        # objective  0.5 z'Pz + q'z + const
        P = np.zeros((nz, nz))
        q = np.zeros(nz)
        const = 0.0
        for t in range(H):
            P[xs(t + 1), xs(t + 1)] += 2.0 * Q
            q[xs(t + 1)] += -2.0 * (Q @ x_ref[t + 1])
            const += float(x_ref[t + 1] @ Q @ x_ref[t + 1])
            P[us(t), us(t)] += 2.0 * R
        for k in range(n_s):
            P[os_ + k, os_ + k] += 2.0 * 50.0
            q[os_ + k] += 50.0

        # equalities: initial state and dynamics
        Aeq = np.zeros(((H + 1) * n, nz))
        beq = np.zeros((H + 1) * n)
        Aeq[0:n, xs(0)] = np.eye(n)
        beq[0:n] = x0
        for t in range(H):
            r = slice((t + 1) * n, (t + 2) * n)
            Aeq[r, xs(t + 1)] = np.eye(n)
            Aeq[r, xs(t)] = -A
            Aeq[r, us(t)] = -B

        # inequalities  G z <= h
        rows, rhs = [], []

        def add(row, b):
            rows.append(row)
            rhs.append(b)

        if input_constraints is not None:
            u_min, u_max = (np.asarray(v, float) for v in input_constraints)
            for t in range(H):
                for j in range(m):
                    row = np.zeros(nz)
                    row[ou + t * m + j] = -1.0
                    add(row, -u_min[j])
                    row = np.zeros(nz)
                    row[ou + t * m + j] = 1.0
                    add(row, u_max[j])
        if position_constraints is not None:
            pos_min, pos_max = (np.asarray(v, float) for v in position_constraints)
            pos_min, pos_max = pos_min[:no], pos_max[:no]      # reference :101-107: only the position part
            for t in range(1, H + 1):
                for j in range(no):
                    row = np.zeros(nz)
                    row[xs(t)] = -C[j]
                    add(row, -pos_min[j])
                    row = np.zeros(nz)
                    row[xs(t)] = C[j]
                    add(row, pos_max[j])
        for k, (t, hh, gg) in enumerate(hs_rows):
            row = np.zeros(nz)
            row[xs(t)] = hh @ C                                  # h . (C x[t]) + g <= slack
            row[os_ + k] = -1.0
            add(row, -gg)
            row = np.zeros(nz)
            row[os_ + k] = -1.0                                  # slack >= 0
            add(row, 0.0)
        G = np.array(rows) if rows else np.zeros((0, nz))
        h = np.array(rhs) if rhs else np.zeros(0)

        try:
            with Timer("MPC Solve"):
                z, ok = _solve_qp(P, q, G, h, Aeq, beq)
            if ok:
                x_filtered = z[ox:ou].reshape(H + 1, n)
                u_filtered = z[ou:os_].reshape(H, m)
                self.last_optimal_u = u_filtered
                return x_filtered, u_filtered, {
                    'status': OPTIMAL,
                    'solve_time': time.time() - start_time,
                    'objective': float(0.5 * z @ P @ z + q @ z + const),
                }
            return self._fallback(x0, x_ref, u_ref, {
                'status': 'solver_error',
                'error': 'Problem could not be solved optimally'
            })
        except Exception as e:
            return self._fallback(x0, x_ref, u_ref, {
                'status': 'ERROR',
                'error': str(e)
            })

    def _fallback(self, x0, x_ref, u_ref, info):
        """Reference core/mpc_filter.py:180-218: shifted last optimal inputs (or the reference inputs), simulated forward."""
        info['used_fallback'] = True
        if self.last_optimal_u is not None:
            u_filtered = np.zeros((self.horizon, self.n_inputs))
            remaining_steps = min(self.horizon - 1, len(self.last_optimal_u) - 1)
            u_filtered[:remaining_steps] = self.last_optimal_u[1:remaining_steps + 1]
            if remaining_steps < self.horizon:
                u_filtered[remaining_steps:] = u_ref[remaining_steps:]
        else:
            u_filtered = u_ref
        x_filtered = np.zeros((self.horizon + 1, self.n_states))
        x_filtered[0] = x0
        for t in range(self.horizon):
            x_filtered[t + 1] = self.A @ x_filtered[t] + self.B @ u_filtered[t]
        return x_filtered, u_filtered, info
