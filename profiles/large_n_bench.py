"""Large-N throughput (BASELINE config 5 sample count): cluster / DSMEM kernel vs the two-pass streaming kernel.
usage: python profiles/large_n_bench.py [B] [N]      (device-resident fp32 samples, CUDA events around each launch)"""
import json
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg  # noqa: E402
from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib  # noqa: E402

P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
PEAK = 6537.3
try:
    PEAK = json.load(open(__import__("os").path.join(__import__("os").path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
g = torch.Generator(device="cuda").manual_seed(42)
ang = torch.rand(B, 1, generator=g, device="cuda") * 6.2831853
rad = 1 + 4 * torch.rand(B, 1, generator=g, device="cuda")
mu = torch.cat([rad * torch.cos(ang), rad * torch.sin(ang)], 1).reshape(B, 1, 2)
s = torch.empty(B, N, 2, device="cuda", dtype=torch.float32)
CH = max(1, (1 << 28) // (N * 2))
for b0 in range(0, B, CH):
    b1 = min(B, b0 + CH)
    s[b0:b1] = mu[b0:b1] + 0.1 * torch.randn(b1 - b0, N, 2, generator=g, device="cuda")
ego = torch.zeros(B, 2, device="cuda", dtype=torch.float64)
bytes_per = N * 8 + 56
out = {}
for name, flags in (("cluster", 0), ("streaming", _lib.FLAG_NO_CLUSTER)):
    best = 1e30
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        r = pkg.compute_halfspaces(s, ego, flags=flags, **P)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = min(best, ms)
        gen = int((r.status & 2 != 0).sum().item())
        print(f"{name:9s} it{it}: {ms:8.2f} ms  {B / ms * 1e3 / 1e6:6.3f} M hs/s  {B * bytes_per / ms / 1e6:7.1f} GB/s "
              f"= {B * bytes_per / ms / 1e6 / PEAK * 100:5.1f}% of {PEAK:.0f}  general-path {gen}")
    out[name] = r
    print()
a, b = out["cluster"], out["streaming"]
print("h equal", bool((a.h == b.h).all()), "T equal", bool((a.var == b.var).all()), "max |dg|", float((a.g - b.g).abs().max()))
