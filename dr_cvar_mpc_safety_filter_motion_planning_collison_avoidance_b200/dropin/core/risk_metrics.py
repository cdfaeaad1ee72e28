"""
Drop-in for the reference's core/risk_metrics.py: same module globals, class names, call signatures,
return values, printed lines and tmp/timing_info_*.json side channel — but the two cvxpy/ECOS linear
programs (reference core/risk_metrics.py:84-265) are replaced by their exact closed form, evaluated on
the GPU through libdrcvar.so (dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200.engine).

    CVaR LP     min g : tau + (1/(aN)) sum (-(h.xi_i) - g + r - tau)^+ <= delta        =>  g  = CVaR_a(-h.xi) + r - delta
    DR-CVaR LP  (Wasserstein ball eps, lambda* = 1/a)                                  =>  g* = CVaR_a(-h.xi) + r + eps/a - delta

There is no CPU fallback: without the CUDA library the import of the engine fails loudly.
Deviation from the reference (documented in DESIGN.md): the optimizer singletons are rebuilt whenever
alpha / delta / epsilon change, not only when N changes (reference :289, :325 silently ignored them).
"""
import json
import os
import time

import numpy as np

try:   # the reference's own helpers when this overlay sits on a reference checkout
    from utils.timing import timeit
except ImportError:   # standalone: dropin/_stopwatch.py
    from _stopwatch import timeit

from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib as _abi
from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import engine as _engine

# module-global optimizer singletons, as in the reference
drcvar_optimizer = None
cvar_optimizer = None

SENTINEL = 100.0


_timing_files = {}   # (cwd, key) -> open file: the side channel is rewritten in place, not re-created, on every call


def save_timing_info(key, setup_time, solve_time):
    """Write tmp/timing_info_<key>.json (seconds) and print the DEBUG line the reference prints
    (reference core/risk_metrics.py:16-33: same file, same keys, same line)."""
    where = (os.getcwd(), key)
    f = _timing_files.get(where)
    if f is None or f.closed or not os.path.exists(f.name):
        os.makedirs('tmp', exist_ok=True)
        f = _timing_files[where] = open(os.path.abspath(f'tmp/timing_info_{key}.json'), 'w')
    f.seek(0)
    f.write(json.dumps({'setup_time': setup_time, 'solve_time': solve_time}))
    f.truncate()
    f.flush()
    print(f"DEBUG - Saved {key} timing: setup={setup_time*1000:.2f}ms, solve={solve_time*1000:.2f}ms")


def expected_value(samples):
    """Mean of the samples along axis 0."""
    return np.mean(samples, axis=0)


def var_metric(samples, alpha):
    """Value-at-Risk: the ceil(N (1 - alpha))-th smallest sample (1-based)."""
    ordered = np.sort(samples)
    return ordered[int(np.ceil(len(samples) * (1 - alpha))) - 1]


def cvar_metric(samples, alpha):
    """Mean of the samples that are >= VaR (the reference's sample estimator; NOT the LP's CVaR)."""
    var = var_metric(samples, alpha)
    tail = samples[samples >= var]
    return var if len(tail) == 0 else np.mean(tail)


def _gpu_cvar(h, samples, alpha):
    """CVaR_alpha of the loss -(h . xi) over the samples, and whether the input was finite."""
    res = _engine.compute_halfspaces(samples, None, alpha=alpha, delta=0.0, epsilon=0.0, robot_radius=0.0,
                                     obstacle_radius=0.0, h=h)
    ok = not (int(res.status[0]) & _abi.STATUS_NONFINITE)
    return ok, float(res.cvar[0])


class DRCVaROptimizer:
    """Closed-form replacement of the reference's DR-CVaR LP wrapper (same constructor and solve())."""

    def __init__(self, alpha, epsilon, delta, max_samples):
        self.alpha = alpha
        self.epsilon = epsilon
        self.delta = delta
        self.n_samples = max_samples

    def solve(self, h, samples, combined_radius):
        """Returns (solved, g_star, info) with g_star = CVaR_a(-h.xi) + combined_radius + eps/alpha - delta."""
        t0 = time.time()
        h = np.asarray(h, dtype=np.float64)
        t1 = time.time()
        ok, cvar = _gpu_cvar(h, samples, self.alpha)
        t2 = time.time()
        setup_time, solve_time = t1 - t0, t2 - t1
        info = {'setup_time': setup_time, 'solve_time': solve_time, 'solve_call_time': setup_time + solve_time}
        save_timing_info('drcvar', setup_time, solve_time)
        if ok:
            g_star = ((cvar + float(combined_radius)) + self.epsilon / self.alpha) - self.delta
            return True, float(g_star), info
        print("Warning: DR-CVaR optimization failed with status: non-finite input")
        return False, SENTINEL, info


class CVaROptimizer:
    """Closed-form replacement of the reference's CVaR LP wrapper (same constructor and solve())."""

    def __init__(self, alpha, delta, max_samples):
        self.alpha = alpha
        self.delta = delta
        self.n_samples = max_samples

    def solve(self, h, samples, combined_radius):
        """Returns (solved, g, info) with g = CVaR_a(-h.xi) + combined_radius * |h| - delta."""
        t0 = time.time()
        h = np.asarray(h, dtype=np.float64)
        r = float(combined_radius) * float(np.sqrt(h[0] * h[0] + h[1] * h[1]))
        t1 = time.time()
        ok, cvar = _gpu_cvar(h, samples, self.alpha)
        t2 = time.time()
        setup_time, solve_time = t1 - t0, t2 - t1
        info = {'setup_time': setup_time, 'solve_time': solve_time, 'solve_call_time': setup_time + solve_time}
        save_timing_info('cvar', setup_time, solve_time)
        if ok:
            return True, float((cvar + r) - self.delta), info
        print("Warning: CVaR optimization failed with status: non-finite input")
        return False, SENTINEL, info


@timeit
def dr_cvar_halfspace(samples, h, alpha, delta, epsilon, robot_radius, obstacle_radius):
    """(g_star, g_tilde) of the DR-CVaR safe halfspace with normal h; g_tilde = g_star - (r_r + r_o)|h|."""
    global drcvar_optimizer
    o = drcvar_optimizer
    if (o is None or o.n_samples != len(samples) or o.alpha != alpha or o.epsilon != epsilon or o.delta != delta):
        drcvar_optimizer = DRCVaROptimizer(alpha, epsilon, delta, len(samples))
    h = np.asarray(h, dtype=np.float64)
    combined_radius = (robot_radius + obstacle_radius) * float(np.sqrt(h[0] * h[0] + h[1] * h[1]))
    solved, g_star, _ = drcvar_optimizer.solve(h, samples, combined_radius)
    if solved:
        return g_star, g_star - combined_radius
    return SENTINEL, SENTINEL - combined_radius


@timeit
def cvar_halfspace(samples, h, alpha, delta, robot_radius, obstacle_radius):
    """Offset g of the CVaR safe halfspace with normal h (used directly as g_tilde by the caller)."""
    global cvar_optimizer
    o = cvar_optimizer
    if o is None or o.n_samples != len(samples) or o.alpha != alpha or o.delta != delta:
        cvar_optimizer = CVaROptimizer(alpha, delta, len(samples))
    solved, g_value, _ = cvar_optimizer.solve(h, samples, robot_radius + obstacle_radius)
    return g_value if solved else SENTINEL
