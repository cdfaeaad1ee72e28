import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
P = dict(alpha=0.2, delta=0.1, epsilon=0.15, robot_radius=0.3, obstacle_radius=0.3)
rng = np.random.RandomState(0)
for n in (100, 1500):
    s = np.array([0.5, 0.0]) + 0.1 * rng.standard_normal((1, n, 2))
    sd = torch.from_numpy(s).cuda(); ego = torch.zeros((1, 2), dtype=torch.float64, device="cuda")
    out = None
    for _ in range(20): out = pkg.compute_halfspaces(sd, ego, out=out, **P)
    torch.cuda.synchronize()
    # kernel-only device time via events over 100 back-to-back launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): out = pkg.compute_halfspaces(sd, ego, out=out, **P)
    e1.record(); torch.cuda.synchronize()
    print(f"N={n}: device path, 100 launches back to back: {e0.elapsed_time(e1) * 10:.1f} us per launch (device time)")
    ts = []
    for _ in range(200):
        t0 = time.perf_counter(); out = pkg.compute_halfspaces(sd, ego, out=out, **P); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"N={n}: device path call + synchronize: median {np.median(ts) * 1e6:.1f} us")
