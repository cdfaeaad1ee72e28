"""
Per-phase cycle breakdown of halfspace_kernel (profiling build, clock64 instrumentation).
  make -C <pkg>/csrc prof && python profiles/phase_cycles.py [B] [N]
Loads libdrcvar_prof.so through the same ctypes binding (DRCVAR_LIB) and prints the mean cycles per halfspace
spent in each phase by sweep warp 1 and by the finisher warp.
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else int(os.environ.get('DRCVAR_DEBUG_GRID', '296')) * 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
dt = torch.float32 if (len(sys.argv) <= 3 or sys.argv[3] == "f32") else torch.float64
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
mu = (torch.rand((B, 1, 2), generator=g, device=dev) * 4 + 1)
s = (mu + 0.1 * torch.randn((B, N, 2), generator=g, device=dev)).to(dt)
ego = torch.zeros((B, 2), dtype=torch.float64, device=dev)
buf = torch.zeros((4096, 2, 12), dtype=torch.int64, device=dev)
lib.drcvar_debug_phase_buffer.argtypes = [ctypes.c_void_p]
lib.drcvar_debug_phase_buffer(buf.data_ptr())
P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
for _ in range(3):
    out = pkg.compute_halfspaces(s, ego, **P)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = pkg.compute_halfspaces(s, ego, **P); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
grid = min(B, 296, int(os.environ.get('DRCVAR_DEBUG_GRID', '296')))
c = buf[:grid].double().cpu().numpy()
per = B / grid
names_s = ["wait for the finisher (empty[par])", "sweep A + warp reduce", "S1 wait", "direction (S2 wait)", "sweep B", "phase 2", "partials",
           "S3 wait", "post (handoff/general/tail)", "wait for the first chunk (data0)", "wait for the other chunks (inside sweep A)", "-"]
names_f = ["wait full", "select + epilogue", "-"]
print(f"B={B} N={N} {dt}: {ms:.3f} ms, {B/ms/1e3:.2f} M halfspaces/s, {B*N*2*s.element_size()/ms/1e6:.0f} GB/s; {per:.1f} halfspaces per CTA")
tot = c[:, 0, :].sum(axis=1).mean() / per
print(f"sweep warp 1: {tot:.0f} cycles per halfspace per CTA")
for k in range(12):
    print(f"  {names_s[k]:32s} {c[:, 0, k].mean() / per:8.0f}")
print("finisher warp:")
for k in range(2):
    print(f"  {names_f[k]:32s} {c[:, 1, k].mean() / per:8.0f}")
