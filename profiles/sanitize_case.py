"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel path once.
   compute-sanitizer --tool memcheck python profiles/sanitize_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg  # noqa: E402
from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib  # noqa: E402

rng = np.random.RandomState(0)
P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
for dtype in (np.float32, np.float64):
    for n in (10000, 2500, 777, 20):
        B = 40
        s = (rng.uniform(1, 4, size=(B, 1, 2)) + 0.1 * rng.standard_normal((B, n, 2))).astype(dtype)
        s[3] = 2.0 + rng.randint(0, 3, size=(n, 2)) * 0.5          # heavy ties -> window miss -> re-fetch path
        ego = rng.uniform(-1, 1, size=(B, 2))
        a = pkg.compute_halfspaces(s, ego, **P)                                        # window path, bulk loader
        b = pkg.compute_halfspaces(s, ego, want_tail=True, **P)                        # parity mode
        c = pkg.compute_halfspaces(s, ego, flags=_lib.FLAG_NO_BULK | _lib.FLAG_GENERAL_ONLY, **P)
        assert np.array_equal(a.var, b.var) and np.array_equal(a.var, c.var)
print("sanitize_case ok")
