"""Generate mode throughput: samples drawn in-kernel.  usage: python profiles/generated_mode_bench.py [B]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg  # noqa: E402

P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
B, N = (int(sys.argv[1]) if len(sys.argv) > 1 else 655360), 10000
rng = np.random.RandomState(0)
ang = rng.uniform(0, 2 * np.pi, B)
mean = np.stack([np.cos(ang), np.sin(ang)], 1) * rng.uniform(1, 5, (B, 1))
ego = np.zeros((B, 2))
cov = np.diag([0.01, 0.01])
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    r = pkg.compute_halfspaces_generated(mean, cov, N, seed=1, ego=ego, device=0, **P)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"device path incl. H2D of the 56 B/halfspace inputs: {ms:.2f} ms -> {B / ms * 1e3 / 1e6:.2f} M hs/s; "
          f"fallback {(r.status.cpu().numpy() & 2 != 0).sum()}")
t0 = time.perf_counter()
r = pkg.compute_halfspaces_generated(mean, cov, N, seed=1, ego=ego, **P)
t1 = time.perf_counter()
print(f"host path e2e: {(t1 - t0) * 1e3:.2f} ms -> {B / (t1 - t0) / 1e6:.2f} M hs/s")
