"""
Drop-in for the reference's core/halfspaces.py: SafeHalfspace value type and the Mean / CVaR / DR-CVaR
factories with unchanged signatures, printed lines, `.info` dicts and tmp/timing_info_*.json side channel
(reference core/halfspaces.py:11-247).  The offsets come from the GPU engine (libdrcvar.so); the normal h
of each halfspace is the one the kernel derived from the canonical sample mean, so (h, g) is self-consistent.

compute_safe_halfspaces() evaluates all obstacles and all three metrics in ONE launch.
"""
import time

import numpy as np

try:   # the reference's own helpers when this overlay sits on a reference checkout
    from utils.timing import timeit, Timer
except ImportError:   # standalone: dropin/_stopwatch.py
    from _stopwatch import timeit, Timer
from core.geometry import compute_separating_vector  # noqa: F401  (re-exported like the reference module)
from core import risk_metrics as _rm
from core.risk_metrics import dr_cvar_halfspace, cvar_halfspace  # noqa: F401

from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib as _abi
from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import engine as _engine


class SafeHalfspace:
    """Safe halfspace {y | h.y + g <= 0}: h points from the ego to the obstacle, g carries geometry and risk."""

    def __init__(self, h, g_tilde):
        self.h = h
        self.g_tilde = g_tilde
        self.info = None

    def is_point_safe(self, point):
        return np.dot(self.h, point) + self.g_tilde <= 0

    def distance_to_boundary(self, point):
        n = np.linalg.norm(self.h)
        return np.dot(self.h / n, point) + self.g_tilde / n

    def get_constraint_params(self):
        return self.h, self.g_tilde


def _zero_info():
    return {'setup_time': 0, 'solve_time': 0, 'solve_call_time': 0}


def _evaluate(samples_list, ego_ref_pos, alpha, delta, epsilon, robot_radius, obstacle_radius):
    """All obstacles in one launch.  samples_list: list of [N,2] arrays (views allowed, same N)."""
    n = len(samples_list)
    same = all(np.shape(s) == np.shape(samples_list[0]) for s in samples_list)
    if same:
        batch = np.stack([np.asarray(s) for s in samples_list]) if n > 1 else np.asarray(samples_list[0])[None]
        ego = None if ego_ref_pos is None else np.broadcast_to(np.asarray(ego_ref_pos, dtype=np.float64), (n, 2))
        return [_engine.compute_halfspaces(batch, ego, alpha=alpha, delta=delta, epsilon=epsilon,
                                           robot_radius=robot_radius, obstacle_radius=obstacle_radius)], True
    return [_engine.compute_halfspaces(np.asarray(s), ego_ref_pos, alpha=alpha, delta=delta, epsilon=epsilon,
                                       robot_radius=robot_radius, obstacle_radius=obstacle_radius)
            for s in samples_list], False


class MeanSafeHalfspace(SafeHalfspace):
    """Safe halfspace from the mean obstacle position (direction measured from the origin)."""

    @staticmethod
    def create(samples, robot_radius, obstacle_radius):
        res = _engine.compute_halfspaces(np.asarray(samples), None, alpha=1.0, delta=0.0, epsilon=0.0,
                                         robot_radius=robot_radius, obstacle_radius=obstacle_radius)
        halfspace = MeanSafeHalfspace(np.array(res.h_mean[0]), float(res.g[0, 0]))
        halfspace.info = _zero_info()
        return halfspace


class CVaRSafeHalfspace(SafeHalfspace):
    """Safe halfspace from the CVaR of the obstacle positions."""

    @staticmethod
    @timeit
    def create(samples, ego_ref_pos, alpha, delta, robot_radius, obstacle_radius):
        t0 = time.time()
        with Timer("CVaR Optimization"):
            res = _engine.compute_halfspaces(np.asarray(samples), ego_ref_pos, alpha=alpha, delta=delta, epsilon=0.0,
                                             robot_radius=robot_radius, obstacle_radius=obstacle_radius)
            g_value = float(res.g[0, 1])
            if _rm.cvar_optimizer is None or _rm.cvar_optimizer.n_samples != len(samples):
                _rm.cvar_optimizer = _rm.CVaROptimizer(alpha, delta, len(samples))
            solve_time = time.time() - t0
            _rm.save_timing_info('cvar', 0.0, solve_time)
        halfspace = CVaRSafeHalfspace(np.array(res.h[0]), g_value)
        # what the reference reads back from tmp/timing_info_cvar.json (core/halfspaces.py:142-147): the values just written
        halfspace.info = {'setup_time': 0.0, 'solve_time': solve_time}
        return halfspace


class DRCVaRSafeHalfspace(SafeHalfspace):
    """Safe halfspace from the Wasserstein distributionally robust CVaR."""

    @staticmethod
    @timeit
    def create(samples, ego_ref_pos, alpha, delta, epsilon, robot_radius, obstacle_radius):
        t0 = time.time()
        with Timer("DR-CVaR Optimization"):
            res = _engine.compute_halfspaces(np.asarray(samples), ego_ref_pos, alpha=alpha, delta=delta,
                                             epsilon=epsilon, robot_radius=robot_radius, obstacle_radius=obstacle_radius)
            g_tilde = float(res.g[0, 2])
            if _rm.drcvar_optimizer is None or _rm.drcvar_optimizer.n_samples != len(samples):
                _rm.drcvar_optimizer = _rm.DRCVaROptimizer(alpha, epsilon, delta, len(samples))
            solve_time = time.time() - t0
            _rm.save_timing_info('drcvar', 0.0, solve_time)
        halfspace = DRCVaRSafeHalfspace(np.array(res.h[0]), g_tilde)
        # what the reference reads back from tmp/timing_info_drcvar.json (core/halfspaces.py:187-192): the values just written
        halfspace.info = {'setup_time': 0.0, 'solve_time': solve_time}
        return halfspace


def compute_safe_halfspaces(obstacle_samples, ego_ref_pos, robot_radius, obstacle_radius, alpha, delta, epsilon):
    """
    Safe halfspaces of every obstacle for the three risk metrics:
    {'mean': [...], 'cvar': [...], 'dr_cvar': [...]}, list index = obstacle (reference core/halfspaces.py:196-247).
    """
    out = {'mean': [], 'cvar': [], 'dr_cvar': []}
    if len(obstacle_samples) == 0:
        return out
    t0 = time.time()
    results, batched = _evaluate(list(obstacle_samples), ego_ref_pos, alpha, delta, epsilon, robot_radius,
                                 obstacle_radius)
    dt = time.time() - t0
    _rm.save_timing_info('cvar', 0.0, dt)
    _rm.save_timing_info('drcvar', 0.0, dt)
    info = {'setup_time': 0.0, 'solve_time': dt}
    for i in range(len(obstacle_samples)):
        res, j = (results[0], i) if batched else (results[i], 0)
        mean_hs = MeanSafeHalfspace(np.array(res.h_mean[j]), float(res.g[j, 0]))
        mean_hs.info = _zero_info()
        cvar_hs = CVaRSafeHalfspace(np.array(res.h[j]), float(res.g[j, 1]))
        cvar_hs.info = dict(info)
        dr_hs = DRCVaRSafeHalfspace(np.array(res.h[j]), float(res.g[j, 2]))
        dr_hs.info = dict(info)
        out['mean'].append(mean_hs)
        out['cvar'].append(cvar_hs)
        out['dr_cvar'].append(dr_hs)
    return out
