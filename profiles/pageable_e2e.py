"""End-to-end rate of compute_halfspaces() on HOST numpy inputs: pinned vs pageable, for a few staging thread counts.
usage: python profiles/pageable_e2e.py   (DRCVAR_STAGE_THREADS is read once per process: each setting runs in a child)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time, json, numpy as np
sys.path.insert(0, %r)
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
B, N = 8192, 10000
rng = np.random.default_rng(0)
s = (rng.uniform(1, 4, size=(B, 1, 2)) + 0.1 * rng.standard_normal((B, N, 2))).astype(np.float32)
ego = np.zeros((B, 2))
P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
import torch
pinned = torch.empty((B, N, 2), dtype=torch.float32).pin_memory().numpy(); pinned[...] = s
out = {}
for name, arr in (("pinned", pinned), ("pageable", s), ("strided", np.repeat(s[:, :, None, :], 2, axis=2)[:, :, 1, :])):
    r0 = pkg.compute_halfspaces(arr, ego, **P)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); r = pkg.compute_halfspaces(arr, ego, **P); ts.append(time.perf_counter() - t0)
    assert np.array_equal(r.g, r0.g)
    out[name] = {"M_hs_per_s": B / min(ts) / 1e6, "GB_per_s": B * N * 8 / min(ts) / 1e9}
print(json.dumps(out))
''' % ROOT
for thr in ("1", "2", "4", "8", "16"):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, DRCVAR_STAGE_THREADS=thr), capture_output=True, text=True)
    if r.returncode:
        print(thr, "FAILED", r.stderr[-400:])
        continue
    d = json.loads(r.stdout.strip().splitlines()[-1])
    print(f"stage threads {thr:>2}: " + "  ".join(f"{k} {v['M_hs_per_s']:.3f} M hs/s ({v['GB_per_s']:.1f} GB/s)" for k, v in d.items()))
