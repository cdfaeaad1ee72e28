import sys
import torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
for dt, N, B in ((torch.float64, 10000, 131072), (torch.float64, 5000, 131072), (torch.float64, 2000, 262144), (torch.float32, 10000, 131072), (torch.float32, 20000, 65536)):
    g = torch.Generator(device="cuda").manual_seed(1)
    mu = torch.rand(B, 1, 2, generator=g, device="cuda") * 8 - 4
    s = torch.empty(B, N, 2, device="cuda", dtype=dt)
    for b0 in range(0, B, 4096):
        b1 = min(B, b0 + 4096)
        s[b0:b1] = (mu[b0:b1] + 0.1 * torch.randn(b1 - b0, N, 2, generator=g, device="cuda")).to(dt)
    ego = torch.zeros(B, 2, device="cuda", dtype=torch.float64)
    eb = 8 if dt == torch.float64 else 4
    for name, fl in (("resident", 0), ("streaming", _lib.FLAG_FORCE_STREAMING)):
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            r = pkg.compute_halfspaces(s, ego, flags=fl, **P)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        print(f"{str(dt):14s} N={N:6d} {name:10s}: {ms:8.2f} ms  {B/ms*1e3/1e6:7.3f} M hs/s  {B*(N*2*eb+56)/ms/1e6:7.1f} GB/s  general {int((r.status & 2 != 0).sum())}")
    del s
