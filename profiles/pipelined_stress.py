"""Randomised stress of the pipelined resident kernel against halfspace_kernel (FLAG_NO_PIPELINE): bitwise h, h_mean, T; offsets 1e-6; reruns bit-identical.
usage: python profiles/pipelined_stress.py [cases] [seed]     (device-resident fp32 batches of random B, N, alpha)"""
import sys

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg  # noqa: E402
from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
g = torch.Generator(device="cuda").manual_seed(1234)
bad = 0
for c in range(cases):
    N = 2 * int(rng.randint(512, 12000))
    B = int(rng.choice([1, 2, 7, 295, 296, 297, 592, 593, 1000, 3001, 9000]))
    if B * N * 8 > 6e9:
        B = max(1, int(6e9 // (N * 8)))
    alpha = float(rng.choice([0.05, 0.1, 0.2, 0.3]))
    P = dict(alpha=alpha, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
    mu = (torch.rand(B, 1, 2, generator=g, device="cuda") * 8 - 4)
    s = torch.empty(B, N, 2, device="cuda")
    dist = str(rng.choice(["gauss", "gauss", "uniform", "laplace", "mixed"]))
    for b0 in range(0, B, 64):
        nb = min(64, B - b0)
        kind = dist if dist != "mixed" else str(rng.choice(["gauss", "uniform", "laplace"]))
        if kind == "gauss":
            z = 0.1 * torch.randn(nb, N, 2, generator=g, device="cuda")
        elif kind == "uniform":
            z = 0.3464 * (torch.rand(nb, N, 2, generator=g, device="cuda") - 0.5)
        else:
            u = (torch.rand(nb, N, 2, generator=g, device="cuda") - 0.5).clamp(-0.4999999, 0.4999999)
            z = -0.0707 * torch.sign(u) * torch.log1p(-2 * u.abs())
        s[b0:b0 + 64] = mu[b0:b0 + 64] + z
    ego = torch.rand(B, 2, generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    l0 = pkg.launch_count()
    a = pkg.compute_halfspaces(s, ego, **P)
    used_cluster = pkg.launch_count() - l0 == 2   # pipelined kernel + redo pass
    a2 = pkg.compute_halfspaces(s, ego, **P)
    b = pkg.compute_halfspaces(s, ego, flags=_lib.FLAG_NO_PIPELINE, **P)
    torch.cuda.synchronize()
    checks = {"h": bool((a.h == b.h).all()), "h_mean": bool((a.h_mean == b.h_mean).all()), "T": bool((a.var == b.var).all()),
              "g": float((a.g - b.g).abs().max()) <= 1e-6, "rerun_g": bool((a.g == a2.g).all()),
              "rerun_cvar": bool((a.cvar == a2.cvar).all())}
    ok = all(checks.values())
    redo = int((a.status & 2 != 0).sum())
    print(f"case {c:3d}: {dist:7s} B={B:4d} N={N:6d} alpha={alpha:.2f} pipelined={used_cluster} redo={redo} "
          f"max|dg|={float((a.g - b.g).abs().max()):.2e} {'ok' if ok else 'MISMATCH ' + str([k for k, v in checks.items() if not v])}", flush=True)
    bad += not ok
    del s, a, a2, b
print("mismatches:", bad)
sys.exit(1 if bad else 0)
