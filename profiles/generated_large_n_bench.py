import sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
B, N = 8192, 100000
rng = np.random.RandomState(0)
ang = rng.uniform(0, 2*np.pi, B)
mean = np.stack([np.cos(ang), np.sin(ang)], 1) * rng.uniform(1, 5, (B, 1))
ego = np.zeros((B, 2)); cov = np.diag([0.01, 0.01])
for name, fl in (("cluster", 0), ("streaming", _lib.FLAG_NO_CLUSTER)):
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        r = pkg.compute_halfspaces_generated(mean, cov, N, seed=1, ego=ego, device=0, flags=fl, **P)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    print(f"{name}: {ms:.2f} ms -> {B/ms*1e3/1e6:.3f} M hs/s, {B*N/ms*1e3/1e9:.1f} G samples/s; general {(r.status.cpu().numpy() & 2 != 0).sum()}")
    if name == "cluster": keep = r
print("T equal", bool((keep.var == r.var).all()), "h equal", bool((keep.h == r.h).all()), float((keep.g - r.g).abs().max()))
