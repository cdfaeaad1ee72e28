"""
Scenario sharding across the GPUs of one box.

Every halfspace is independent (core/halfspaces.py:225-246 and simulation/environment.py:82-104 of the
reference have no cross-halfspace term), so the batch is partitioned on the SCENARIO axis: rank r of W owns the
contiguous scenario range plan_shards(S, W)[r] with all of its obstacles and horizon steps.  There is no
collective on the hot path; the only exchange is the final gather of the small result arrays
(h [B,2], g [B,3]) with torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np


def plan_shards(n_scenarios: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [start, stop) scenario ranges, sizes differing by at most one, in rank order."""
    if n_scenarios < 0 or world_size < 1:
        raise ValueError("n_scenarios >= 0 and world_size >= 1 required")
    base, extra = divmod(n_scenarios, world_size)
    out, start = [], 0
    for r in range(world_size):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


def shard_of(n_scenarios: int, world_size: int, rank: int) -> Tuple[int, int]:
    return plan_shards(n_scenarios, world_size)[rank]


def halfspaces_per_scenario(n_obstacles: int, horizon: int) -> int:
    return int(n_obstacles) * int(horizon)


def gather_results(local_h, local_g, n_scenarios: int, per_scenario: int, group=None):
    """
    All-gather the per-rank results into the global [S*per_scenario, ...] arrays (rank order == scenario order).
    Works on torch tensors of any device the process group supports; ragged shards are padded to the largest.
    Returns (h, g) as torch tensors on the local device.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local_h, local_g
    world = dist.get_world_size(group)
    shards = plan_shards(n_scenarios, world)
    max_rows = max(b - a for a, b in shards) * per_scenario
    packed = torch.zeros((max_rows, 5), dtype=torch.float64, device=local_h.device)
    rows = local_h.shape[0]
    packed[:rows, 0:2] = local_h
    packed[:rows, 2:5] = local_g
    out = torch.empty((world * max_rows, 5), dtype=torch.float64, device=local_h.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    parts = []
    for r, (a, b) in enumerate(shards):
        n = (b - a) * per_scenario
        parts.append(out[r * max_rows: r * max_rows + n])
    full = torch.cat(parts, dim=0)
    return full[:, 0:2].contiguous(), full[:, 2:5].contiguous()


def run_sharded(make_shard_inputs: Callable, compute_fn: Callable, n_scenarios: int, per_scenario: int,
                rank: int, world_size: int, group=None):
    """
    make_shard_inputs(start, stop) -> (samples [(stop-start)*per_scenario, N, 2], ego [.., 2]) for a scenario range;
    compute_fn(samples, ego) -> object with .h [B,2] and .g [B,3] (the engine's HalfspaceBatch).
    Returns the gathered global (h, g).
    """
    import torch
    a, b = shard_of(n_scenarios, world_size, rank)
    samples, ego = make_shard_inputs(a, b)
    res = compute_fn(samples, ego)
    h = res.h if isinstance(res.h, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(res.h))
    g = res.g if isinstance(res.g, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(res.g))
    return gather_results(h, g, n_scenarios, per_scenario, group)
