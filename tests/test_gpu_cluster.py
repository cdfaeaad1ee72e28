"""
GPU parity tests of the cluster / DSMEM kernel (csrc/cluster_kernel.cuh): N > 32768 fp32 — BASELINE config 5 is
N = 100 000 — one thread-block cluster per halfspace, every sample read from HBM once.  All calls go through the C ABI.

Bars: h, h_mean and the threshold T (= kc-th largest loss, which fixes the tail-index set) bit-exact against the
oracle; offsets <= 1e-6 m against the oracle on the same fp32 samples; bit-identical T / h against the streaming kernel
(DRCVAR_FLAG_NO_CLUSTER), whose tail-index sets are checked bit for bit in test_gpu_parity.py; run-to-run determinism.
"""
import numpy as np
import pytest

from oracle import closed_form as cf

pytestmark = pytest.mark.gpu

ABS32 = 1e-6
P = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
    return pkg


def gaussian_batch(rng, B, n, spread=4.0):
    mu = rng.uniform(-spread, spread, size=(B, 1, 2))
    s = (mu + 0.1 * rng.standard_normal((B, n, 2))).astype(np.float32)
    ego = rng.uniform(-1, 1, size=(B, 2))
    return s, ego


def check_against_oracle(res, s, ego, p, which, h_in=None):
    for b in which:
        o = cf.halfspace(s[b], ego[b], p["alpha"], p["delta"], p["epsilon"], p["robot_radius"], p["obstacle_radius"],
                         None if h_in is None else h_in[b])
        assert np.array_equal(res.h[b], o.h), (b, res.h[b], o.h)
        assert np.array_equal(res.h_mean[b], o.h_mean), (b, res.h_mean[b], o.h_mean)
        assert res.var[b] == o.var, (b, res.var[b], o.var)                       # T bit-exact -> the tail set is fixed
        ref = np.array([o.g_mean, o.g_cvar, o.g_dr])
        assert np.abs(res.g[b] - ref).max() <= ABS32, (b, res.g[b], ref)
        assert abs(res.cvar[b] - o.cvar) <= ABS32 and abs(res.g_star[b] - o.g_dr_star) <= ABS32
        assert abs(res.g[b, 0] - o.g_mean) <= 1e-9 * max(1.0, abs(o.g_mean))     # the mean halfspace has no fp32 shortcut


def host_chunks(B, n):
    """The host path stages ~64 MB of samples per launch (drcvar_abi.cu: run_host)."""
    per = max(1, (64 << 20) // (n * 8))
    return -(-B // per)


def launches(eng, fn):
    before = eng.launch_count()
    r = fn()
    return r, eng.launch_count() - before


def test_config5_sample_count(eng):
    """N = 100 000 (4 CTAs per halfspace), enough halfspaces for several pipelined iterations per cluster."""
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    rng = np.random.RandomState(101)
    B, n = 240, 100000
    s, ego = gaussian_batch(rng, B, n)
    res, nl = launches(eng, lambda: eng.compute_halfspaces(s, ego, **P))
    assert nl == 2 * host_chunks(B, n)  # per chunk: cluster kernel + the (normally empty) redo pass of the streaming kernel
    ref, nl2 = launches(eng, lambda: eng.compute_halfspaces(s, ego, flags=_lib.FLAG_NO_CLUSTER, **P))
    assert nl2 == host_chunks(B, n)
    assert np.array_equal(res.h, ref.h) and np.array_equal(res.h_mean, ref.h_mean) and np.array_equal(res.var, ref.var)
    assert np.abs(res.g - ref.g).max() <= ABS32 and np.array_equal(res.g[:, 0], ref.g[:, 0])
    assert (res.status & _lib.STATUS_GENERAL).sum() <= 2          # window misses are 3e-5 events
    check_against_oracle(res, s, ego, P, range(0, B, 7))
    again = eng.compute_halfspaces(s, ego, **P)
    for k in ("h", "h_mean", "g", "cvar", "var", "g_star", "status"):
        assert np.array_equal(getattr(res, k), getattr(again, k)), k


@pytest.mark.parametrize("n", [32770, 40000, 52002, 65544, 131072, 150000, 200000])
def test_cluster_sizes_and_ragged_octants(eng, n):
    """2, 4 and 8 CTAs per halfspace; short / empty last octants; parts that end inside a 32 KB chunk."""
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    rng = np.random.RandomState(n % 1000)
    B = 40
    s, ego = gaussian_batch(rng, B, n)
    res, nl = launches(eng, lambda: eng.compute_halfspaces(s, ego, **P))
    assert nl == 2 * host_chunks(B, n)
    ref = eng.compute_halfspaces(s, ego, flags=_lib.FLAG_NO_CLUSTER, **P)
    assert np.array_equal(res.h, ref.h) and np.array_equal(res.var, ref.var) and np.array_equal(res.h_mean, ref.h_mean)
    assert np.abs(res.g - ref.g).max() <= ABS32
    check_against_oracle(res, s, ego, P, (0, 17, B - 1))


def test_other_tail_fractions_and_explicit_normals(eng):
    rng = np.random.RandomState(7)
    B, n = 24, 100000
    s, ego = gaussian_batch(rng, B, n)
    for alpha in (0.2, 0.05, 0.5):
        p = dict(P, alpha=alpha)
        res = eng.compute_halfspaces(s, ego, **p)
        check_against_oracle(res, s, ego, p, (0, 11, B - 1))
    th = rng.uniform(0, 2 * np.pi, size=B)
    h = np.stack([np.cos(th), np.sin(th)], axis=1) * rng.uniform(0.5, 2.0, size=(B, 1))     # non-unit normals too
    res = eng.compute_halfspaces(s, None, h=h, **P)
    check_against_oracle(res, s, np.zeros((B, 2)), P, (0, 5, B - 1), h_in=h)


def test_redo_list_non_gaussian_and_non_finite(eng):
    """Samples the Gaussian window plan does not fit (uniform noise), a NaN sample and a halfspace whose mean sits on
    the ego position: the cluster kernel hands them to the streaming kernel's exact general select."""
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    rng = np.random.RandomState(3)
    B, n = 12, 100000
    mu = rng.uniform(-4, 4, size=(B, 1, 2))
    s = (mu + rng.uniform(-0.3, 0.3, size=(B, n, 2))).astype(np.float32)
    ego = rng.uniform(-1, 1, size=(B, 2))
    s[5, 77777, 1] = np.nan
    s[9] = (mu[9] + 0.1 * rng.standard_normal((n, 2))).astype(np.float32)     # one Gaussian halfspace among them
    res, nl = launches(eng, lambda: eng.compute_halfspaces(s, ego, **P))
    assert nl == 2 * host_chunks(B, n)
    assert res.status[5] & _lib.STATUS_NONFINITE and res.g[5, 1] == 100.0
    assert not (res.status[9] & _lib.STATUS_GENERAL)
    assert (res.status[[0, 1, 2, 3]] & _lib.STATUS_GENERAL).all()
    check_against_oracle(res, s, ego, P, (0, 3, 9, 11))
    # direction undefined: mean == ego -> fallback [1, 0] (core/geometry.py:49-51)
    m = np.array([cf.canonical_mean(s[2])])
    res2 = eng.compute_halfspaces(s[2:3], m, **P)
    assert res2.status[0] & _lib.STATUS_DEGENERATE and np.array_equal(res2.h[0], [1.0, 0.0])
    check_against_oracle(res2, s[2:3], m, P, (0,))


def test_device_pointers_and_single_halfspace(eng):
    import torch
    rng = np.random.RandomState(11)
    s, ego = gaussian_batch(rng, 5, 100000)
    host = eng.compute_halfspaces(s, ego, **P)
    dev = eng.compute_halfspaces(torch.from_numpy(s).cuda(), torch.from_numpy(ego).cuda(), **P)
    torch.cuda.synchronize()
    assert np.array_equal(dev.g.cpu().numpy(), host.g) and np.array_equal(dev.var.cpu().numpy(), host.var)
    one = eng.compute_halfspaces(s[3], ego[3], **P)
    assert np.array_equal(one.g[0], host.g[3])


# ---------------------------------------------------------------------------------------------- fp64 samples
def gaussian_batch64(rng, B, n, spread=4.0):
    mu = rng.uniform(-spread, spread, size=(B, 1, 2))
    return mu + 0.1 * rng.standard_normal((B, n, 2)), rng.uniform(-1, 1, size=(B, 2))


def check_against_oracle64(res, s, ego, p, which, h_in=None):
    for b in which:
        o = cf.halfspace(s[b], ego[b], p["alpha"], p["delta"], p["epsilon"], p["robot_radius"], p["obstacle_radius"],
                         None if h_in is None else h_in[b])
        assert np.array_equal(res.h[b], o.h) and np.array_equal(res.h_mean[b], o.h_mean), b
        assert res.var[b] == o.var, (b, res.var[b], o.var)
        ref = np.array([o.g_mean, o.g_cvar, o.g_dr])
        assert np.all(np.abs(res.g[b] - ref) <= 1e-9 * np.maximum(1.0, np.abs(ref))), (b, res.g[b], ref)
        assert abs(res.cvar[b] - o.cvar) <= 1e-9 * max(1.0, abs(o.cvar))


@pytest.mark.parametrize("n", [32770, 50001, 100000, 100352])
def test_fp64_cluster_kernel(eng, n):
    """fp64 samples (the reference's dtype) on the opt-in fp64 cluster kernel (DRCVAR_FLAG_FORCE_CLUSTER): N = 100 000 is
    1.6 MB per halfspace = 8 CTAs x 200 KB; bit-exact h / T against the oracle and the streaming kernel (the default for
    fp64: it is faster there), offsets within 1e-9 relative."""
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    rng = np.random.RandomState(n % 997)
    B = 45
    s, ego = gaussian_batch64(rng, B, n)
    per = max(1, (64 << 20) // (n * 16))
    res, nl = launches(eng, lambda: eng.compute_halfspaces(s, ego, flags=_lib.FLAG_FORCE_CLUSTER, **P))
    assert nl == 2 * -(-B // per)                 # per host chunk: cluster kernel + its redo pass
    ref, nl2 = launches(eng, lambda: eng.compute_halfspaces(s, ego, **P))
    assert nl2 == -(-B // per)                    # default for fp64: the (faster) two-pass streaming kernel
    assert np.array_equal(res.h, ref.h) and np.array_equal(res.h_mean, ref.h_mean) and np.array_equal(res.var, ref.var)
    assert np.all(np.abs(res.g - ref.g) <= 1e-9 * np.maximum(1.0, np.abs(ref.g)))
    assert (res.status & _lib.STATUS_GENERAL).sum() <= 1
    check_against_oracle64(res, s, ego, P, (0, 21, B - 1))
    again = eng.compute_halfspaces(s, ego, flags=_lib.FLAG_FORCE_CLUSTER, **P)
    assert np.array_equal(again.g, res.g) and np.array_equal(again.cvar, res.cvar)


def test_fp64_cluster_redo_and_explicit_normals(eng):
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    rng = np.random.RandomState(19)
    B, n = 10, 60000
    mu = rng.uniform(-4, 4, size=(B, 1, 2))
    s = mu + rng.uniform(-0.3, 0.3, size=(B, n, 2))                 # uniform noise: the Gaussian window misses
    s[7] = mu[7] + 0.1 * rng.standard_normal((n, 2))
    s[3, 12345, 0] = np.inf
    ego = rng.uniform(-1, 1, size=(B, 2))
    res = eng.compute_halfspaces(s, ego, flags=_lib.FLAG_FORCE_CLUSTER, **P)
    assert res.status[3] & _lib.STATUS_NONFINITE and res.g[3, 1] == 100.0
    assert not (res.status[7] & _lib.STATUS_GENERAL) and (res.status[[0, 1, 2]] & _lib.STATUS_GENERAL).all()
    check_against_oracle64(res, s, ego, P, (0, 7, 9))
    th = rng.uniform(0, 2 * np.pi, size=B)
    h = np.stack([np.cos(th), np.sin(th)], axis=1) * rng.uniform(0.5, 2.0, size=(B, 1))
    s2, _ = gaussian_batch64(rng, B, n)
    res2 = eng.compute_halfspaces(s2, None, h=h, flags=_lib.FLAG_FORCE_CLUSTER, **P)
    check_against_oracle64(res2, s2, np.zeros((B, 2)), P, (0, 4, 9), h_in=h)


@pytest.mark.parametrize("n,ctas", [(40000, 2), (100000, 4), (100002, 4), (180000, 8)])
def test_tail_indices_from_the_cluster_kernel(eng, n, ctas):
    """Parity mode at cluster sizes: every CTA emits the tail indices of its resident part (one read of the samples) at the
    positions the leader's finisher assigns; bit-exact against the oracle and the streaming kernel, ties to the lower index,
    hand-backs (heavy ties, non-finite) through the redo pass."""
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    assert _lib.load().drcvar_cluster_ctas(n, 4, 232448) == ctas
    rng = np.random.RandomState(n % 1000 + ctas)
    B = 7
    p = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
    s = (rng.uniform(-4, 4, size=(B, 1, 2)) + 0.1 * rng.standard_normal((B, n, 2))).astype(np.float32)
    ego = rng.uniform(-1, 1, size=(B, 2))
    # exact ties AT the threshold inside an otherwise continuous cloud: duplicate the sample that sits at the threshold
    o0 = cf.halfspace(s[1], ego[1], p["alpha"], p["delta"], p["epsilon"], p["robot_radius"], p["obstacle_radius"])
    at_T = int(o0.tail_idx[np.argmin(-(s[1][o0.tail_idx].astype(np.float64) @ o0.h))])      # the kc-th largest loss
    for k in (5, n // 3, n // 2 + 1, n - 2):
        s[1, k] = s[1, at_T]
    s[3] = (2.0 + rng.randint(0, 4, size=(n, 2)) * 0.25).astype(np.float32)                 # massive ties: window miss -> redo pass
    s[5, n // 2, 0] = np.nan                                                                 # non-finite: redo pass, indices -1
    l0 = eng.launch_count()
    res = eng.compute_halfspaces(s, ego, want_tail=True, **p)
    assert eng.launch_count() - l0 == 2                                                      # cluster kernel + redo pass
    ref = eng.compute_halfspaces(s, ego, want_tail=True, flags=_lib.FLAG_NO_CLUSTER, **p)    # streaming kernel
    assert np.array_equal(res.tail_idx, ref.tail_idx) and np.array_equal(res.var[[0, 1, 2, 3, 4, 6]], ref.var[[0, 1, 2, 3, 4, 6]])
    assert np.all(res.tail_idx[5] == -1) and (res.status[5] & _lib.STATUS_NONFINITE)
    for b in (0, 1, 3, 6):
        o = cf.halfspace(s[b], ego[b], p["alpha"], p["delta"], p["epsilon"], p["robot_radius"], p["obstacle_radius"])
        assert np.array_equal(res.tail_idx[b], o.tail_idx), b
        assert res.var[b] == o.var and np.array_equal(res.h[b], o.h)
    plain = eng.compute_halfspaces(s, ego, **p)                                              # the timed instantiation: same T
    assert np.array_equal(plain.var[[0, 1, 2, 4, 6]], res.var[[0, 1, 2, 4, 6]])


@pytest.mark.parametrize("n,B", [(40000, 170), (100000, 75)])
def test_tail_indices_with_several_halfspaces_per_cluster(eng, n, B):
    """More halfspaces than resident clusters: the message barrier, the leader rotation and the deferred slot release of the
    parity mode run through several halfspaces per cluster."""
    import torch
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(n + B)
    mu = torch.rand(B, 1, 2, generator=g, device="cuda") * 8 - 4
    s = (mu + 0.1 * torch.randn(B, n, 2, generator=g, device="cuda")).contiguous()
    ego = torch.rand(B, 2, generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    p = dict(alpha=0.05, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
    a = eng.compute_halfspaces(s, ego, want_tail=True, **p)
    a2 = eng.compute_halfspaces(s, ego, want_tail=True, **p)
    b = eng.compute_halfspaces(s, ego, want_tail=True, flags=_lib.FLAG_NO_CLUSTER, **p)
    torch.cuda.synchronize()
    assert bool((a.tail_idx == b.tail_idx).all()) and bool((a.var == b.var).all()) and bool((a.h == b.h).all())
    assert bool((a.tail_idx == a2.tail_idx).all()) and bool((a.g == a2.g).all())
    kc = a.tail_idx.shape[1]
    assert kc == int(np.ceil(0.05 * n - 1e-9))
    idx = a.tail_idx.cpu().numpy()
    assert (np.diff(idx, axis=1) > 0).all() and idx.min() >= 0 and idx.max() < n          # ascending, in range
    sn, en = s[B - 1].cpu().numpy(), ego[B - 1].cpu().numpy()
    o = cf.halfspace(sn, en, p["alpha"], p["delta"], p["epsilon"], p["robot_radius"], p["obstacle_radius"])
    assert np.array_equal(idx[B - 1], o.tail_idx)
