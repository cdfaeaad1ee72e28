import sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
B, N = (int(sys.argv[1]) if len(sys.argv) > 1 else 4096), 100000
g = torch.Generator(device="cuda").manual_seed(3)
mu = torch.rand(B, 1, 2, generator=g, device="cuda") * 8 - 4
ego = torch.zeros(B, 2, device="cuda", dtype=torch.float64)
for dist in ("gauss", "uniform", "laplace"):
    s = torch.empty(B, N, 2, device="cuda")
    for b0 in range(0, B, 128):
        nb = min(128, B - b0)
        if dist == "gauss":
            z = 0.1 * torch.randn(nb, N, 2, generator=g, device="cuda")
        elif dist == "uniform":
            z = 0.3464 * (torch.rand(nb, N, 2, generator=g, device="cuda") - 0.5)
        else:
            u = (torch.rand(nb, N, 2, generator=g, device="cuda") - 0.5).clamp(-0.4999999, 0.4999999)
            z = -0.0707 * torch.sign(u) * torch.log1p(-2 * u.abs())
        s[b0:b0 + nb] = mu[b0:b0 + nb] + z
    for name, fl in (("cluster+redo", 0), ("streaming", _lib.FLAG_NO_CLUSTER)):
        for it in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            r = pkg.compute_halfspaces(s, ego, flags=fl, **P)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        print(f"{dist:8s} {name:13s}: {ms:8.2f} ms -> {B/ms*1e3/1e6:.3f} M hs/s; general-path {(r.status & 2 != 0).sum().item()}")
