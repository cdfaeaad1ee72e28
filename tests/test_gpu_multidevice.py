"""
One process driving several GPUs through the host (DRCVAR_HOST) entry: every device keeps its own streams, events and
staging buffers (csrc/drcvar_abi.cu: HostCtx per device), so alternating cuda:0 / cuda:1 / cuda:0 works and gives the
same bits as a single-device run.  Needs two visible GPUs (skipped otherwise): gpurun --gpus 2.
"""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
    return pkg


def _batch(seed, B=96, N=10000):
    rng = np.random.default_rng(seed)
    mu = rng.uniform(1.0, 5.0, size=(B, 1, 2))
    return (mu + 0.1 * rng.standard_normal((B, N, 2))).astype(np.float32), rng.uniform(-0.5, 0.5, size=(B, 2))


def test_host_path_alternating_devices(eng):
    import torch
    s, ego = _batch(1)
    with torch.cuda.device(0):
        ref = eng.compute_halfspaces(s, ego, **P)
    for dev in (1, 0, 1, 1, 0):
        with torch.cuda.device(dev):
            res = eng.compute_halfspaces(s, ego, **P)
        assert np.array_equal(res.g, ref.g) and np.array_equal(res.h, ref.h) and np.array_equal(res.var, ref.var), dev


def test_host_path_one_thread_per_device(eng):
    import torch
    batches = [_batch(10 + d) for d in range(2)]
    want = []
    with torch.cuda.device(0):
        for s, ego in batches:
            want.append(eng.compute_halfspaces(s, ego, **P).g.copy())
    got, errs = [None, None], []

    def work(d):
        try:
            with torch.cuda.device(d):
                for _ in range(4):
                    got[d] = eng.compute_halfspaces(*batches[d], **P).g.copy()
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(d,)) for d in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for d in range(2):
        assert np.array_equal(got[d], want[d])
