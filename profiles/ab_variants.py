"""A/B timing of library builds on the config-4 batch (device-resident, CUDA events): python profiles/ab_variants.py lib1.so lib2.so ...
Each build runs in its own process (DRCVAR_LIB), twice, interleaved; prints ms per launch and the fraction of the measured HBM peak."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, torch
sys.path.insert(0, %r)
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
B, N = int(os.environ.get("AB_B", "655360")), int(os.environ.get("AB_N", "10000"))
dt = torch.float64 if os.environ.get("AB_DTYPE") == "f64" else torch.float32
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(42)
s = torch.empty((B, N, 2), dtype=dt, device=dev)
step = 8192
for lo in range(0, B, step):
    hi = min(B, lo + step)
    mu = torch.rand((hi - lo, 1, 2), generator=g, device=dev) * 4 + 1
    s[lo:hi] = (mu + 0.1 * torch.randn((hi - lo, N, 2), generator=g, device=dev)).to(dt)
ego = torch.zeros((B, 2), dtype=torch.float64, device=dev)
P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
out = None
for _ in range(3):
    out = pkg.compute_halfspaces(s, ego, out=out, **P)
torch.cuda.synchronize()
ts = []
clk, pw, why = [], [], 0
try:
    import pynvml
    pynvml.nvmlInit(); hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:
    hnd = None
for _ in range(int(os.environ.get("AB_ITERS", "20"))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = pkg.compute_halfspaces(s, ego, out=out, **P); e1.record()
    if hnd is not None:   # sampled while the launch is in flight
        clk.append(pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM)); pw.append(pynvml.nvmlDeviceGetPowerUsage(hnd) / 1000.0)
        why |= pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort(); clk.sort(); pw.sort()
fallback = int((out.status.cpu().numpy() != 0).sum()) if getattr(out, "status", None) is not None else -1
print(json.dumps({"ms_median": ts[len(ts) // 2], "ms_min": ts[0], "B": B, "N": N, "elem": s.element_size(), "fallback": fallback,
                  "clk_min": clk[0] if clk else None, "clk_med": clk[len(clk) // 2] if clk else None,
                  "pw_med": pw[len(pw) // 2] if pw else None, "why": why}))
''' % ROOT

libs = sys.argv[1:]   # "lib.so" or "lib.so:ENV=value[,ENV=value]" (environment of that variant's child process)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6537.3
res = {l: [] for l in libs}
for rep in range(2):
    for l in libs:
        env = dict(os.environ, DRCVAR_LIB=os.path.abspath(l.split(":")[0]))
        for kv in (l.split(":")[1].split(",") if ":" in l else []):
            env[kv.split("=")[0]] = kv.split("=")[1]
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            print(l, "FAILED", r.stderr[-500:])
            continue
        res[l].append(json.loads(r.stdout.strip().splitlines()[-1]))
for l in libs:
    for d in res[l]:
        gbs = d["B"] * (d["N"] * 2 * d["elem"] + 56) / d["ms_median"] / 1e6
        print(f"{os.path.basename(l):28s} median {d['ms_median']:.3f} ms  min {d['ms_min']:.3f} ms  {d['B'] / d['ms_median'] / 1e3:.2f} M hs/s  {gbs:.0f} GB/s = {gbs / peak:.3f} of measured peak"
              f"  fallbacks {d.get('fallback')}  | SM MHz min/med {d.get('clk_min')}/{d.get('clk_med')}  {d.get('pw_med')} W  throttle reasons 0x{d.get('why') or 0:x}")
