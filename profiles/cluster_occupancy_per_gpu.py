"""Active clusters of the cluster kernel on every GPU of the box (floor-sweeping differs per die) and its throughput there.
usage: make -C <pkg>/csrc prof && python profiles/cluster_occupancy_per_gpu.py   (profiling build prints max_active_clusters)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200"
os.environ["DRCVAR_LIB"] = os.path.join(ROOT, PKG, "libdrcvar_prof.so")
import torch  # noqa: E402
import importlib  # noqa: E402

pkg = importlib.import_module(PKG)
P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
B, N = 4096, 100000
for d in range(torch.cuda.device_count()):
    torch.cuda.set_device(d)
    dev = torch.device("cuda", d)
    g = torch.Generator(device=dev).manual_seed(1)
    mu = torch.rand(B, 1, 2, generator=g, device=dev) * 8 - 4
    s = torch.empty(B, N, 2, device=dev)
    for b0 in range(0, B, 256):
        s[b0:b0 + 256] = mu[b0:b0 + 256] + 0.1 * torch.randn(256, N, 2, generator=g, device=dev)
    ego = torch.zeros(B, 2, device=dev, dtype=torch.float64)
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev); e0.record()
        r = pkg.compute_halfspaces(s, ego, **P)
        e1.record(); torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
    print(f"GPU {d}: {torch.cuda.get_device_properties(d).multi_processor_count} SMs, {ms:.2f} ms -> {B/ms*1e3/1e6:.3f} M hs/s", flush=True)
    del s
