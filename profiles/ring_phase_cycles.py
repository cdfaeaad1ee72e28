"""
Per-phase cycle breakdown of halfspace_ring_kernel (profiling build, clock64 instrumentation; sweep warp 1).
  make -C <pkg>/csrc prof && [DRCVAR_DEBUG_FLAGS=k] python profiles/rf_phase_cycles.py [B] [N]
DRCVAR_DEBUG_FLAGS (timing only, results are wrong): 1 = skip sweep B.
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
mu = (torch.rand((B, 1, 2), generator=g, device=dev) * 4 + 1)
s = (mu + 0.1 * torch.randn((B, N, 2), generator=g, device=dev)).float()
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
grid = min(B, int(os.environ.get("DRCVAR_DEBUG_GRID", "148")))
c = buf[:grid].double().cpu().numpy()
per = B / grid
names = ["prologue (first sweep A)", "sweep A of b+1 (incl. wait for its data)", "wait for placer", "sweep B of b", "exact phase of b",
         "-", "-", "-"]
print(f"flags={os.environ.get('DRCVAR_DEBUG_FLAGS', '0')} B={B} N={N}: {ms:.3f} ms, {B/ms/1e3:.2f} M halfspaces/s, "
      f"{B*N*8/ms/1e6:.0f} GB/s; {per:.1f} halfspaces per CTA, grid {grid}")
tot = c[:, 0, :8].sum(axis=1).mean() / per
print(f"sweep warp 1: {tot:.0f} cycles per halfspace")
for k in range(5):
    print(f"  {names[k]:32s} {c[:, 0, k].mean() / per:8.0f}")
