"""Every kernel path once on small cases — the workload for a checked build (make -C csrc checked; DRCVAR_LIB=.../libdrcvar_checked.so)
or for compute-sanitizer where that tool is available:
   DRCVAR_LIB=$PWD/dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200/libdrcvar_checked.so python profiles/sanitize_case.py
Prints the number of failed device-side assertions (drcvar_debug_check_failures; -1 = not a checked build)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg  # noqa: E402
from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib  # noqa: E402

rng = np.random.RandomState(0)
P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)


def batch(B, n, dtype, kind="gauss"):
    mu = rng.uniform(1, 4, size=(B, 1, 2))
    if kind == "laplace":
        s = mu + rng.laplace(scale=0.07, size=(B, n, 2))
    else:
        s = mu + 0.1 * rng.standard_normal((B, n, 2))
    s = s.astype(dtype)
    s[min(3, B - 1)] = (2.0 + rng.randint(0, 3, size=(n, 2)) * 0.5).astype(dtype)   # heavy ties -> window miss -> redo / re-fetch
    return s, rng.uniform(-1, 1, size=(B, 2))


for dtype in (np.float32, np.float64):
    for n in (10000, 4096, 2500, 1024, 777, 20):
        for B in (40, 700):                                                            # < and > one pass of the persistent grid
            s, ego = batch(B, n, dtype)
            a = pkg.compute_halfspaces(s, ego, **P)                                    # pipelined (fp32, n >= 1024) / resident kernel
            b = pkg.compute_halfspaces(s[:40], ego[:40], want_tail=True, **P)          # parity mode
            c = pkg.compute_halfspaces(s[:40], ego[:40], flags=_lib.FLAG_NO_BULK | _lib.FLAG_GENERAL_ONLY, **P)
            d = pkg.compute_halfspaces(s, ego, flags=_lib.FLAG_NO_PIPELINE, **P)       # resident kernel, inline general path
            e = pkg.compute_halfspaces(s[:40], ego[:40], flags=_lib.FLAG_FORCE_STREAMING, **P)
            assert np.array_equal(a.var[:40], b.var) and np.array_equal(a.var[:40], c.var) and np.array_equal(a.var, d.var)
            assert np.array_equal(a.var[:40], e.var)
    s, ego = batch(600, 4096, dtype, "laplace")                                        # learned windows
    a = pkg.compute_halfspaces(s, ego, **P)
    c = pkg.compute_halfspaces(s, ego, flags=_lib.FLAG_GENERAL_ONLY, **P)
    assert np.array_equal(a.var, c.var)
    nmax = pkg.max_samples(dtype)
    for n in (nmax, nmax + 1, 40000, 100000):                                          # slot limit, cluster kernels (2 / 4 / 8 CTAs), streaming
        s, ego = batch(9, n, dtype)
        a = pkg.compute_halfspaces(s, ego, **P)
        t = pkg.compute_halfspaces(s[:3], ego[:3], want_tail=True, **P)
        f = pkg.compute_halfspaces(s, ego, flags=_lib.FLAG_FORCE_CLUSTER, **P)
        g = pkg.compute_halfspaces(s, ego, flags=_lib.FLAG_NO_CLUSTER, **P)
        assert np.array_equal(a.var[:3], t.var) and np.array_equal(a.var, f.var) and np.array_equal(a.var, g.var)
# generate mode: resident and cluster sizes
for n in (4096, 100000):
    mean = rng.uniform(1, 4, size=(12, 2))
    r = pkg.compute_halfspaces_generated(mean, np.diag([0.01, 0.02]), n, seed=5, ego=np.zeros((12, 2)), **P)
    assert np.isfinite(r.g).all()
# trajectory entry (strided [N, T+1, 2] views packed by the library)
traj = [np.ascontiguousarray(rng.uniform(1, 3, size=(1, 1, 2)) + 0.1 * rng.standard_normal((1500, 31, 2))) for _ in range(3)]
h, hm, g, _ = pkg.compute_trajectory(traj, rng.uniform(-1, 1, size=(20, 2)), **P)
assert np.isfinite(g).all()
site = ctypes.c_int32(0)
n_fail = _lib.load().drcvar_debug_check_failures(ctypes.byref(site))
print(f"sanitize_case ok: device-side assertion failures {n_fail} (first site {site.value}; -1 = not a checked build)")
