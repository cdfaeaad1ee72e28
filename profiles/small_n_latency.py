"""
Latency of the reference's own small configurations through the drop-in API:
  * per-call latency of DRCVaRSafeHalfspace.create / CVaRSafeHalfspace.create for N = 10..1500 (the reference's
    timing sweep, evaluation/timing_analysis.py; published ECOS numbers: BASELINE.md);
  * per-trajectory latency of SafetyFilteringEnvironment.compute_safe_halfspaces_for_trajectory for head_on
    (30 halfspaces x 3 metrics) and multi_obstacle (90 x 3), one launch each.
Run on the GPU box:  python profiles/small_n_latency.py
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200", "dropin"), ROOT]
os.chdir(os.environ.get("TMPDIR", "/tmp"))

from core.halfspaces import CVaRSafeHalfspace, DRCVaRSafeHalfspace  # noqa: E402
from simulation.environment import SafetyFilteringEnvironment  # noqa: E402

PUBLISHED_MS = {10: 2.205, 50: 4.220, 100: 6.947, 500: 31.546, 1000: 69.011, 1500: 148.878}   # DR-CVaR call, BASELINE.md
rng = np.random.RandomState(0)
sink = io.StringIO()
print("N, dr_cvar create ms (median of 200), cvar create ms, reference published dr_cvar ms, speed-up")
for n in (10, 50, 100, 500, 1000, 1500):
    s = np.array([0.5, 0.0]) + 0.1 * rng.standard_normal((n, 2))
    ego = np.zeros(2)
    td, tc = [], []
    with contextlib.redirect_stdout(sink):
        for _ in range(220):
            t0 = time.perf_counter(); DRCVaRSafeHalfspace.create(s, ego, 0.2, 0.1, 0.15, 0.3, 0.3); td.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); CVaRSafeHalfspace.create(s, ego, 0.2, 0.1, 0.3, 0.3); tc.append(time.perf_counter() - t0)
    d, c = np.median(td[20:]) * 1e3, np.median(tc[20:]) * 1e3
    print(f"{n}, {d:.3f}, {c:.3f}, {PUBLISHED_MS[n]}, {PUBLISHED_MS[n] / d:.0f}x")
for name in ("head_on_seed42.npz", "multi_obstacle_seed42.npz"):
    z = np.load(os.path.join(ROOT, "tests", "golden", name))
    alpha, delta, eps, rr, ro, horizon = (float(v) for v in z["params"])
    env = SafetyFilteringEnvironment(rr, ro, int(horizon), 0.2, alpha, delta, eps)
    traj = [z["sample_trajectories"][i] for i in range(z["sample_trajectories"].shape[0])]
    ts = []
    for _ in range(60):
        t0 = time.perf_counter(); env.compute_safe_halfspaces_for_trajectory(traj, z["x_ref"]); ts.append(time.perf_counter() - t0)
    print(f"{name}: compute_safe_halfspaces_for_trajectory {np.median(ts[10:]) * 1e3:.3f} ms per trajectory "
          f"({z['g_mean'].size} halfspaces x 3 metrics, one launch)")
