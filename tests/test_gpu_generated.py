"""
GPU parity tests of the generate mode: drcvar_halfspaces_generated_f32 draws the Monte-Carlo samples
inside the kernel.  Bars: the generated samples are BIT-IDENTICAL to oracle/sample_gen.py; the halfspaces computed
from them meet the same bars as the fp32-input path (tests/test_gpu_parity.py) against oracle/closed_form.py
evaluated on the oracle-generated samples; results do not depend on how the batch is sharded.
"""
import numpy as np
import pytest

from oracle import closed_form as cf
from oracle import sample_gen as sg

pytestmark = pytest.mark.gpu

P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
ABS32 = 1e-6


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
    return pkg


def scenario(B, seed):
    rng = np.random.RandomState(seed)
    ang = rng.uniform(0, 2 * np.pi, size=B)
    mean = np.stack([np.cos(ang), np.sin(ang)], axis=1) * rng.uniform(1.0, 5.0, size=(B, 1))
    ego = rng.uniform(-0.5, 0.5, size=(B, 2))
    return mean, ego


@pytest.mark.parametrize("n", [10000, 4097, 1500, 777, 20, 1])
def test_generated_samples_bit_exact_and_halfspaces(eng, n):
    B = 20 if n >= 1000 else 40
    mean, ego = scenario(B, n)
    cov = np.array([[0.01, 0.003], [0.003, 0.02]])
    res = eng.compute_halfspaces_generated(mean, cov, n, seed=1234 + n, ego=ego, want_tail=True, want_samples=True, **P)
    want = sg.generate(mean, sg.cholesky2(cov), n, seed=1234 + n)
    assert res.samples.dtype == np.float32 and res.samples.shape == (B, n, 2)
    assert np.array_equal(res.samples.view(np.uint32), want.view(np.uint32))
    for b in range(B):
        o = cf.halfspace(want[b], ego[b], P["alpha"], P["delta"], P["epsilon"], P["robot_radius"], P["obstacle_radius"])
        assert np.array_equal(res.h[b], o.h) and np.array_equal(res.h_mean[b], o.h_mean), b
        assert res.var[b] == o.var, (b, res.var[b], o.var)
        assert np.array_equal(res.tail_idx[b], o.tail_idx), b
        assert np.abs(res.g[b] - np.array([o.g_mean, o.g_cvar, o.g_dr])).max() <= ABS32
        assert abs(res.cvar[b] - o.cvar) <= ABS32
    # the timed configuration (no dump, no tail) and the plain fp32-input path on the same samples agree
    fast = eng.compute_halfspaces_generated(mean, cov, n, seed=1234 + n, ego=ego, **P)
    assert np.array_equal(fast.var, res.var) and np.array_equal(fast.h, res.h) and np.array_equal(fast.g, res.g)
    plain = eng.compute_halfspaces(want, ego, **P)
    assert np.array_equal(plain.var, res.var) and np.array_equal(plain.h, res.h)
    assert np.abs(plain.g - res.g).max() <= ABS32


def test_generated_sharding_device_path_and_zero_noise(eng):
    import torch
    B, n = 600, 10000          # > 2 halfspaces per CTA of the persistent grid
    mean, ego = scenario(B, 5)
    cov = np.diag([0.01, 0.01])
    whole = eng.compute_halfspaces_generated(mean, cov, n, seed=99, ego=ego, **P)
    assert (whole.status == 0).mean() > 0.99           # the window path, not the fallback
    lo = eng.compute_halfspaces_generated(mean[:250], cov, n, seed=99, ego=ego[:250], **P)
    hi = eng.compute_halfspaces_generated(mean[250:], cov, n, seed=99, ego=ego[250:], index_offset=250, **P)
    assert np.array_equal(np.concatenate([lo.g, hi.g]), whole.g)
    assert np.array_equal(np.concatenate([lo.h, hi.h]), whole.h)
    dev = eng.compute_halfspaces_generated(mean, cov, n, seed=99, ego=ego, device=0, **P)
    torch.cuda.synchronize()
    assert np.array_equal(dev.g.cpu().numpy(), whole.g) and np.array_equal(dev.status.cpu().numpy(), whole.status)
    # a different seed gives different but statistically equivalent offsets
    other = eng.compute_halfspaces_generated(mean, cov, n, seed=100, ego=ego, **P)
    assert not np.array_equal(other.g, whole.g) and np.abs(other.g - whole.g).max() < 0.02
    # zero covariance = the reference's t = 0 row (obstacles.py:63): analytic answer CVaR = -h.xi0
    z = eng.compute_halfspaces_generated(mean[:4], np.zeros((2, 2)), 50, seed=1, ego=ego[:4], want_samples=True, **P)
    assert np.array_equal(z.samples, np.broadcast_to(mean[:4].astype(np.float32)[:, None, :], (4, 50, 2)))
    for b in range(4):
        xi = mean[b].astype(np.float32).astype(np.float64)
        assert abs(z.cvar[b] + z.h[b] @ xi) < 1e-12


def test_generated_bad_arguments(eng):
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    mean = np.zeros((1, 2))
    with pytest.raises(_lib.DrcvarError):
        eng.compute_halfspaces_generated(mean, np.eye(2), 100, seed=1, **dict(P, alpha=0.0))


def test_generated_large_n_streaming(eng):
    """N beyond one CTA's shared memory (BASELINE config 5, N = 100 000): the streaming kernel re-draws the samples in
    every pass; same bit-exact samples, same halfspace bars."""
    n, B = 100000, 3
    assert n > eng.max_samples(np.float32)
    mean, ego = scenario(B, 77)
    cov = np.array([[0.01, -0.002], [-0.002, 0.015]])
    res = eng.compute_halfspaces_generated(mean, cov, n, seed=5, ego=ego, want_tail=True, want_samples=True, **P)
    want = sg.generate(mean, sg.cholesky2(cov), n, seed=5)
    assert np.array_equal(res.samples.view(np.uint32), want.view(np.uint32))
    fast = eng.compute_halfspaces_generated(mean, cov, n, seed=5, ego=ego, **P)
    for b in range(B):
        o = cf.halfspace(want[b], ego[b], P["alpha"], P["delta"], P["epsilon"], P["robot_radius"], P["obstacle_radius"])
        assert np.array_equal(res.h[b], o.h) and res.var[b] == o.var and np.array_equal(res.tail_idx[b], o.tail_idx)
        assert abs(res.cvar[b] - o.cvar) <= ABS32 and abs(fast.cvar[b] - o.cvar) <= ABS32
    assert np.array_equal(fast.var, res.var)


def test_generated_large_n_cluster_kernel(eng):
    """Generate mode on the cluster / DSMEM kernel (N = 100 000, no tail indices): every CTA draws its part of the samples
    into its shared memory once; same Philox stream, so the samples, h and T are bit-identical to the streaming kernel's
    (which re-draws them in each pass) and to the oracle's restatement."""
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    n, B = 100000, 70
    mean, ego = scenario(B, 91)
    cov = np.array([[0.012, 0.003], [0.003, 0.009]])
    before = eng.launch_count()
    res = eng.compute_halfspaces_generated(mean, cov, n, seed=11, ego=ego, want_samples=True, **P)
    assert eng.launch_count() - before == 2          # cluster kernel + its (empty) redo pass
    ref = eng.compute_halfspaces_generated(mean, cov, n, seed=11, ego=ego, flags=_lib.FLAG_NO_CLUSTER, **P)
    assert np.array_equal(res.h, ref.h) and np.array_equal(res.h_mean, ref.h_mean) and np.array_equal(res.var, ref.var)
    assert np.abs(res.g - ref.g).max() <= ABS32
    want = sg.generate(mean, sg.cholesky2(cov), n, seed=11)
    assert np.array_equal(res.samples.view(np.uint32), want.view(np.uint32))
    for b in (0, 33, B - 1):
        o = cf.halfspace(want[b], ego[b], P["alpha"], P["delta"], P["epsilon"], P["robot_radius"], P["obstacle_radius"])
        assert np.array_equal(res.h[b], o.h) and res.var[b] == o.var and abs(res.cvar[b] - o.cvar) <= ABS32
    # shards reproduce their slice of the unsharded stream (index_offset)
    part = eng.compute_halfspaces_generated(mean[40:], cov, n, seed=11, ego=ego[40:], index_offset=40, **P)
    assert np.array_equal(part.var, res.var[40:]) and np.array_equal(part.g, res.g[40:])
