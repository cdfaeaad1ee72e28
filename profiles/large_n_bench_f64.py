"""fp64 samples at N = 100 000: cluster kernel (8 CTAs per halfspace) vs the two-pass streaming kernel.
usage: python profiles/large_n_bench_f64.py [B] [N]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg  # noqa: E402
from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib  # noqa: E402

P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
PEAK = 6537.3
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
g = torch.Generator(device="cuda").manual_seed(42)
mu = torch.rand(B, 1, 2, generator=g, device="cuda", dtype=torch.float64) * 8 - 4
s = torch.empty(B, N, 2, device="cuda", dtype=torch.float64)
for b0 in range(0, B, 128):
    b1 = min(B, b0 + 128)
    s[b0:b1] = mu[b0:b1] + 0.1 * torch.randn(b1 - b0, N, 2, generator=g, device="cuda", dtype=torch.float64)
ego = torch.zeros(B, 2, device="cuda", dtype=torch.float64)
bytes_per = N * 16 + 56
out = {}
for name, flags in (("cluster", _lib.FLAG_FORCE_CLUSTER), ("streaming", 0)):
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        r = pkg.compute_halfspaces(s, ego, flags=flags, **P)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    print(f"{name:9s}: {ms:8.2f} ms  {B / ms * 1e3 / 1e6:6.3f} M hs/s  {B * bytes_per / ms / 1e6:7.1f} GB/s = "
          f"{B * bytes_per / ms / 1e6 / PEAK * 100:5.1f}% of {PEAK:.0f}  general-path {int((r.status & 2 != 0).sum().item())}")
    out[name] = r
a, b = out["cluster"], out["streaming"]
print("h equal", bool((a.h == b.h).all()), "T equal", bool((a.var == b.var).all()),
      "max rel |dg|", float(((a.g - b.g).abs() / b.g.abs().clamp(min=1.0)).max()))
