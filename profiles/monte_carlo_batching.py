"""Halfspaces of NUM_MC_RUNS = 300 Monte-Carlo runs (config/parameters.py:33): one launch for all runs vs one launch per run.
usage: python profiles/monte_carlo_batching.py [n_runs] [n_obstacles] [n_samples]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200", "dropin"))
os.chdir(__import__("tempfile").mkdtemp())
env_mod = importlib.import_module("simulation.environment")

n_runs = int(sys.argv[1]) if len(sys.argv) > 1 else 300
n_obs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = int(sys.argv[3]) if len(sys.argv) > 3 else 20
H, T1 = 30, 151
rng = np.random.RandomState(0)
env = env_mod.SafetyFilteringEnvironment(0.3, 0.3, H, 0.2, 0.2, 0.1, 0.15)
x_ref = np.zeros((H + 1, 4))
x_ref[:, 0] = np.linspace(-4, 4, H + 1)
runs = [[np.array([3.0, 1.0 * i]) + 0.1 * rng.standard_normal((N, T1, 2)) for i in range(n_obs)] for _ in range(n_runs)]
for _ in range(2):
    env.compute_safe_halfspaces_for_runs(runs, x_ref)
t0 = time.perf_counter()
a = env.compute_safe_halfspaces_for_runs(runs, x_ref)
t1 = time.perf_counter()
b = [env.compute_safe_halfspaces_for_trajectory(r, x_ref) for r in runs]
t2 = time.perf_counter()
same = all(np.array_equal(a[r]["dr_cvar"][t][i].h, b[r]["dr_cvar"][t][i].h) and a[r]["dr_cvar"][t][i].g_tilde == b[r]["dr_cvar"][t][i].g_tilde
           for r in range(n_runs) for t in range(H) for i in range(n_obs))
hs = n_runs * H * n_obs
print(f"{n_runs} runs x {H} steps x {n_obs} obstacles x N={N}: one launch {1e3 * (t1 - t0):.1f} ms ({hs / (t1 - t0) / 1e3:.0f} k halfspaces/s incl. "
      f"Python objects), run by run {1e3 * (t2 - t1):.1f} ms; identical: {same}")
