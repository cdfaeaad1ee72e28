"""Coordinates far from the origin (|xi| ~ 500 m): the default fp32 instantiation of the pipelined kernel hands every halfspace to the
redo pass (raw fp32 sums would not be accurate enough), DRCVAR_FLAG_LARGE_COORDS keeps the first-sample-relative kernel.
usage: python profiles/large_coords_bench.py [B]     (device-resident fp32 samples, N = 10 000, CUDA events)"""
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg  # noqa: E402
from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib  # noqa: E402

P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
N = 10000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(5)
for off in (3.0, 500.0):
    s = torch.empty((B, N, 2), dtype=torch.float32, device=dev)
    for lo in range(0, B, 8192):
        hi = min(B, lo + 8192)
        ang = torch.rand((hi - lo, 1, 1), generator=g, device=dev) * 6.2831853
        mu = off * torch.cat([torch.cos(ang), torch.sin(ang)], dim=2)
        s[lo:hi] = mu + 0.1 * torch.randn((hi - lo, N, 2), generator=g, device=dev)
    ego = torch.zeros((B, 2), dtype=torch.float64, device=dev)
    res = {}
    for name, flags in (("default (raw sums, large coordinates -> redo pass)", 0), ("DRCVAR_FLAG_LARGE_COORDS", _lib.FLAG_LARGE_COORDS)):
        out = None
        for _ in range(2):
            out = pkg.compute_halfspaces(s, ego, out=out, flags=flags, **P)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = pkg.compute_halfspaces(s, ego, out=out, flags=flags, **P)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[name] = out
        ms = min(ts)
        print(f"|xi| ~ {off:6.1f} m  {name:52s} {ms:8.3f} ms  {B / ms / 1e3:7.2f} M halfspaces/s")
    a, b = list(res.values())
    same = bool((a.var == b.var).all()) and bool((a.h == b.h).all())
    print(f"               T and h bit-identical between the two: {same};  max |dg| {float((a.g - b.g).abs().max()):.2e}")
    del s
