"""Sustained run of the config-4 step (device-resident): per-step time, SM clock, board power and throttle reasons over a few
seconds — what the 1 000 W cap does to the kernel once the board is warm.   python profiles/sustained_power.py [seconds] [f32|f64]"""
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg  # noqa: E402
import pynvml  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
dt = torch.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else torch.float32
B, N = 655360, 10000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(42)
s = torch.empty((B, N, 2), dtype=dt, device=dev)
for lo in range(0, B, 8192):
    hi = min(B, lo + 8192)
    mu = torch.rand((hi - lo, 1, 2), generator=g, device=dev) * 4 + 1
    s[lo:hi] = (mu + 0.1 * torch.randn((hi - lo, N, 2), generator=g, device=dev)).to(dt)
ego = torch.zeros((B, 2), dtype=torch.float64, device=dev)
P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False


def poll():
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.02)


torch.cuda.synchronize(); time.sleep(1.0)          # start from an idle board
out = pkg.compute_halfspaces(s, ego, **P)
torch.cuda.synchronize()
th = threading.Thread(target=poll); th.start()
t0 = time.perf_counter(); evs = []
while time.perf_counter() - t0 < secs:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = pkg.compute_halfspaces(s, ego, out=out, **P); e1.record(); torch.cuda.synchronize()
    evs.append((time.perf_counter() - t0, e0.elapsed_time(e1)))
stop = True; th.join()
bytes_per = N * 2 * s.element_size() + 56
peak = 6537.3
print(f"{len(evs)} steps in {secs:.0f} s; B = {B}, N = {N}, {s.dtype}")
print("   t[s]   ms/step   frac   SM MHz   W   throttle")
k = 0
for i in range(0, len(evs), max(1, len(evs) // 24)):
    t, ms = evs[i]
    while k + 1 < len(samples) and samples[k + 1][0] - t0 <= t:
        k += 1
    _, clk, pw, why = samples[k]
    print(f"{t:7.2f} {ms:9.3f} {B * bytes_per / ms / 1e6 / peak:6.3f} {clk:7d} {pw:6.0f}   0x{why:x}")
ms_all = sorted(m for _, m in evs)
late = sorted(m for t, m in evs if t > secs / 2)
print(json.dumps({"steps": len(evs), "ms_min": ms_all[0], "ms_median": ms_all[len(ms_all) // 2], "ms_median_second_half": late[len(late) // 2],
                  "power_w_max": max(p for _, _, p, _ in samples), "sm_mhz_min": min(c for _, c, _, _ in samples)}))
