import torch, numpy as np, time, sys
sys.path.insert(0, '/root/repo')
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
p = dict(alpha=0.2, delta=0.1, epsilon=0.15, robot_radius=0.3, obstacle_radius=0.3)
for N, B, dt in ((20, 262144, torch.float64), (100, 262144, torch.float64), (100, 262144, torch.float32), (500, 131072, torch.float64), (1500, 65536, torch.float64), (1500, 65536, torch.float32), (4096, 32768, torch.float32)):
    g = torch.Generator(device='cuda').manual_seed(1)
    s = (torch.randn(B, N, 2, device='cuda', generator=g, dtype=torch.float32) * 0.1 + 3.0).to(dt)
    ego = torch.zeros(B, 2, device='cuda', dtype=torch.float64)
    out = pkg.compute_halfspaces(s, ego, **p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        pkg.compute_halfspaces(s, ego, out=out, **p)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"N={N} B={B} {str(dt)[6:]}: {ms:.3f} ms/launch, {B/ms*1e3/1e6:.2f} M hs/s, {B*N*2*s.element_size()/ms/1e6:.1f} GB/s, general={(out.status.cpu().numpy() & 2 != 0).mean():.2f}")
