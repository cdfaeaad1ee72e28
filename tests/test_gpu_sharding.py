"""
GPU test of the scenario-sharded driver: run_sharded() with the CUDA engine over two emulated ranks
on one device must reproduce the single-launch result bit for bit (shards are independent; no collective on the path).
The process-group gather itself is covered on CPU with gloo (tests/test_sharding_gloo.py).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)


def test_shards_union_equals_single_launch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import sharding
    S, OBS, HOR, N = 7, 3, 4, 4000           # 7 scenarios: ragged over 2 and 4 ranks
    per = sharding.halfspaces_per_scenario(OBS, HOR)
    g = torch.Generator(device="cuda").manual_seed(3)
    mu = torch.rand((S * per, 1, 2), generator=g, device="cuda") * 4 + 1
    s = (mu + 0.1 * torch.randn((S * per, N, 2), generator=g, device="cuda")).float()
    ego = torch.rand((S * per, 2), generator=g, device="cuda", dtype=torch.float64) - 0.5
    full = pkg.compute_halfspaces(s, ego, **P)
    torch.cuda.synchronize()
    for world in (2, 4):
        hs, gs = [], []
        for rank in range(world):
            a, b = sharding.shard_of(S, world, rank)
            r = pkg.compute_halfspaces(s[a * per:b * per], ego[a * per:b * per], **P)
            hs.append(r.h)
            gs.append(r.g)
        torch.cuda.synchronize()
        assert torch.equal(torch.cat(hs), full.h) and torch.equal(torch.cat(gs), full.g)
    # without a process group run_sharded degenerates to the local shard
    h, gg = sharding.run_sharded(lambda a, b: (s[a * per:b * per], ego[a * per:b * per]),
                                 lambda x, e: pkg.compute_halfspaces(x, e, **P), S, per, rank=0, world_size=1)
    torch.cuda.synchronize()
    assert torch.equal(gg, full.g)
