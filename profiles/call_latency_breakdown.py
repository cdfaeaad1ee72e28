"""Where the per-call latency of a single halfspace goes: C ABI (host path) vs engine wrapper vs drop-in create()."""
import contextlib
import ctypes as C
import io
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200", "dropin"), ROOT]
os.chdir(os.environ.get("TMPDIR", "/tmp"))
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg  # noqa: E402
from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib  # noqa: E402
from core.halfspaces import DRCVaRSafeHalfspace  # noqa: E402

lib = _lib.load()
rng = np.random.RandomState(0)
for n in (100, 1500):
    s = np.array([0.5, 0.0]) + 0.1 * rng.standard_normal((n, 2))
    ego = np.zeros((1, 2))
    h, hm, g = np.empty((1, 2)), np.empty((1, 2)), np.empty((1, 3))

    def abi():
        lib.drcvar_halfspaces_f64(s.ctypes.data, 1, n, 2 * n, 2, 1, ego.ctypes.data, None, 0.2, 0.1, 0.15, 0.3, 0.3, 0,
                                  h.ctypes.data, hm.ctypes.data, g.ctypes.data, None, None, None, None, None, -1, None)

    def eng():
        pkg.compute_halfspaces(s, ego, alpha=0.2, delta=0.1, epsilon=0.15, robot_radius=0.3, obstacle_radius=0.3)

    def create():
        with contextlib.redirect_stdout(io.StringIO()):
            DRCVaRSafeHalfspace.create(s, ego[0], 0.2, 0.1, 0.15, 0.3, 0.3)

    for name, fn in (("C ABI (ctypes, host path)", abi), ("engine.compute_halfspaces", eng), ("dropin DRCVaRSafeHalfspace.create", create)):
        for _ in range(50):
            fn()
        ts = []
        for _ in range(300):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        print(f"N={n:5d} {name:36s} median {np.median(ts) * 1e6:7.1f} us  p10 {np.percentile(ts, 10) * 1e6:7.1f} us")
