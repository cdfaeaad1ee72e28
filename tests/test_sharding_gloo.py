"""
CPU tests of the multi-GPU host logic: scenario shard planning and the final gather, with world_size 2
over gloo.  The per-shard compute is the oracle here (no GPU in this container); on the GPU box the same
run_sharded() is driven with the CUDA engine (tests/test_gpu_sharding.py).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import sharding

P = dict(alpha=0.2, delta=0.1, epsilon=0.15, robot_radius=0.3, obstacle_radius=0.3)
S, OBS, HOR, N = 5, 2, 3, 40     # 5 scenarios (ragged over 2 ranks) x 2 obstacles x 3 steps


def test_plan_shards_properties():
    for s in (0, 1, 7, 8, 4096, 65536):
        for w in (1, 2, 3, 4, 8):
            sh = sharding.plan_shards(s, w)
            assert len(sh) == w and sh[0][0] == 0 and sh[-1][1] == s
            assert all(sh[i][1] == sh[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in sh]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.plan_shards(4096, 8)[3] == (1536, 2048)
    with pytest.raises(ValueError):
        sharding.plan_shards(4, 0)


def _inputs(a, b):
    rng = np.random.RandomState(1234)
    per = OBS * HOR
    allx = np.array([1.5, -2.0]) + 0.2 * rng.standard_normal((S * per, N, 2))
    ego = rng.uniform(-1, 1, size=(S * per, 2))
    return allx[a * per: b * per], ego[a * per: b * per]


class _Res:
    def __init__(self, h, g):
        self.h, self.g = h, g


def _oracle_compute(samples, ego):
    from oracle import closed_form as cf
    o = cf.halfspaces_batch(samples, ego, P["alpha"], P["delta"], P["epsilon"], P["robot_radius"], P["obstacle_radius"])
    return _Res(o["h"], o["g"])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h, g = sharding.run_sharded(_inputs, _oracle_compute, S, OBS * HOR, rank, world)
        q.put((rank, h.numpy(), g.numpy()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gather_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full_s, full_e = _inputs(0, S)
    ref = _oracle_compute(full_s, full_e)
    for rank, h, g in got:
        assert h.shape == (S * OBS * HOR, 2) and g.shape == (S * OBS * HOR, 3)
        assert np.array_equal(h, ref.h) and np.array_equal(g, ref.g)     # bit-for-bit, both ranks
