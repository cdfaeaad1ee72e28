"""Throughput of the resident kernel when the samples are NOT Gaussian (the statistical window is planned for Gaussian
losses): Gaussian vs Laplace vs uniform vs Student-t(5) noise of the same variance.  usage: python profiles/non_gaussian_throughput.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg  # noqa: E402

P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
B, N = 65536, 10000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
mu = torch.rand(B, 1, 2, device=dev, generator=g) * 4 + 1
ego = torch.zeros(B, 2, device=dev, dtype=torch.float64)


def noise(kind):
    if kind == "gaussian":
        return torch.randn(B, N, 2, device=dev, generator=g)
    if kind == "laplace":
        u = torch.rand(B, N, 2, device=dev, generator=g) - 0.5
        return -torch.sign(u) * torch.log1p(-2 * u.abs()) / 2 ** 0.5
    if kind == "uniform":
        return (torch.rand(B, N, 2, device=dev, generator=g) - 0.5) * 12 ** 0.5
    if kind == "student_t5":
        z = torch.randn(B, N, 2, device=dev, generator=g)
        chi = torch.randn(B, N, 5, device=dev, generator=g).pow(2).sum(-1, keepdim=True)
        return z / (chi / 5).sqrt() * (3 / 5) ** 0.5
    raise ValueError(kind)


for kind in ("gaussian", "laplace", "uniform", "student_t5"):
    s = (mu + 0.1 * noise(kind)).float().contiguous()
    out = pkg.compute_halfspaces(s, ego, **P)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        pkg.compute_halfspaces(s, ego, out=out, **P)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    fb = float((out.status & 2 != 0).float().mean())
    print(f"{kind:11s}: {B / ms * 1e3 / 1e6:6.2f} M hs/s, fallback fraction {fb:.4f}")
    del s
