"""
Per-phase cycle breakdown of cluster_kernel_f32 (profiling build, clock64 instrumentation in sweep warp 2).
  make -C <pkg>/csrc prof && python profiles/cluster_phase_cycles.py [B] [N]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200"
os.environ["DRCVAR_LIB"] = os.path.join(ROOT, PKG, "libdrcvar_prof.so")

import torch  # noqa: E402
import importlib  # noqa: E402

pkg = importlib.import_module(PKG)
lib = importlib.import_module(PKG + "._lib").load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
mu = (torch.rand((B, 1, 2), generator=g, device=dev) * 4 + 1)
s = torch.empty((B, N, 2), device=dev)
for b0 in range(0, B, 256):
    s[b0:b0 + 256] = mu[b0:b0 + 256] + 0.1 * torch.randn((min(256, B - b0), N, 2), generator=g, device=dev)
ego = torch.zeros((B, 2), dtype=torch.float64, device=dev)
buf = torch.zeros((4096, 2, 12), dtype=torch.int64, device=dev)
lib.drcvar_debug_phase_buffer.argtypes = [ctypes.c_void_p]
lib.drcvar_debug_phase_buffer(buf.data_ptr())
P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
for _ in range(3):
    out = pkg.compute_halfspaces(s, ego, **P)
torch.cuda.synchronize()
buf.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = pkg.compute_halfspaces(s, ego, **P); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
c = buf[:, 0, :].double().cpu().numpy()
live = c.sum(axis=1) > 0
c = c[live]
grid = int(live.sum())
names = ["loop top (first / ego loads)", "sweep A (incl. chunk waits)", "octant trees + moments + sync", "exchange 1 send + wait",
         "window + S2", "sweep B", "release tail + hdone wait", "phase 2b + warp sums", "S3 wait", "exchange 2 send",
         "end-of-halfspace team sync", "-"]
print(f"B={B} N={N}: {ms:.3f} ms, {B/ms/1e3:.3f} M halfspaces/s, {B*N*8/ms/1e6:.0f} GB/s; {grid} CTAs instrumented")
tot = c.sum(axis=1).mean()
print(f"sweep warp 2: {tot:.0f} cycles per CTA in total")
for k in range(11):
    print(f"  {names[k]:36s} {c[:, k].mean():12.0f}  ({100 * c[:, k].mean() / tot:5.1f}%)")
