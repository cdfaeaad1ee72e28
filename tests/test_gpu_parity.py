"""
GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI
(libdrcvar.so via ctypes); the checker is oracle/closed_form.py and the committed golden vectors
that the reference's own modules produced (tests/golden/make_golden.py).

Bars (BASELINE.json north_star):
  * tail-index sets bit-exact (ties -> lower index);
  * h, mean: bit-exact (canonical arithmetic contract);
  * CVaR / offsets: <= 1e-9 relative for fp64 inputs; for the fp32-input path <= 1e-5 m vs the fp64 truth
    (north star) and <= 1e-6 m vs the oracle of the same fp32 samples (the kernel sums the losses that are
    surely above the window through linearity from fp32 coordinate partial sums).
"""
import os

import numpy as np
import pytest

from oracle import closed_form as cf

pytestmark = pytest.mark.gpu

REL = 1e-9
ABS32 = 1e-6
PARAMS = dict(alpha=0.2, delta=0.1, epsilon=0.15, robot_radius=0.3, obstacle_radius=0.3)


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
    return pkg


def rel_close(a, b, tol=REL):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b)))


def check_batch(res, samples, ego, p, h_in=None, tail=True, exact_h=True):
    """Compare a HalfspaceBatch (numpy) with the oracle, halfspace by halfspace."""
    B = samples.shape[0]
    ego = np.broadcast_to(np.zeros(2) if ego is None else np.asarray(ego, dtype=np.float64), (B, 2))
    for b in range(B):
        o = cf.halfspace(samples[b], ego[b], p["alpha"], p["delta"], p["epsilon"], p["robot_radius"],
                         p["obstacle_radius"], None if h_in is None else np.broadcast_to(h_in, (B, 2))[b])
        if exact_h:
            assert np.array_equal(res.h[b], o.h), (b, res.h[b], o.h)
            assert np.array_equal(res.h_mean[b], o.h_mean), (b, res.h_mean[b], o.h_mean)
        else:
            assert np.abs(res.h[b] - o.h).max() < 1e-14
        if samples.dtype == np.float32:
            assert np.abs(res.g[b] - np.array([o.g_mean, o.g_cvar, o.g_dr])).max() <= ABS32, (b, res.g[b], (o.g_mean, o.g_cvar, o.g_dr))
            assert abs(res.cvar[b] - o.cvar) <= ABS32 and abs(res.g_star[b] - o.g_dr_star) <= ABS32
            assert rel_close(res.g[b, 0], o.g_mean)                 # the mean halfspace has no fp32 shortcut
        else:
            assert rel_close(res.g[b], [o.g_mean, o.g_cvar, o.g_dr]), (b, res.g[b], (o.g_mean, o.g_cvar, o.g_dr))
            assert rel_close(res.cvar[b], o.cvar), (b, res.cvar[b], o.cvar)
            assert rel_close(res.g_star[b], o.g_dr_star)
        assert res.var[b] == o.var, (b, res.var[b], o.var)          # the threshold is an order statistic: exact
        if tail and res.tail_idx is not None:
            assert np.array_equal(res.tail_idx[b], o.tail_idx), b


def test_golden_scenarios_through_abi(eng, golden_dir):
    for name in ("head_on_seed42.npz", "multi_obstacle_seed42.npz"):
        z = np.load(os.path.join(golden_dir, name))
        alpha, delta, eps, rr, ro, horizon = z["params"]
        traj = z["sample_trajectories"]                      # [n_obs, N, H+1, 2]
        x_ref = z["x_ref"]
        n_steps = z["g_mean"].shape[0]
        n_obs = traj.shape[0]
        # batch (t, i) -> strided views exactly like simulation/environment.py:88
        samples = np.stack([traj[i][:, t, :] for t in range(n_steps) for i in range(n_obs)])
        ego = np.stack([x_ref[t][:2] for t in range(n_steps) for i in range(n_obs)])
        res = eng.compute_halfspaces(samples, ego, alpha=alpha, delta=delta, epsilon=eps, robot_radius=rr,
                                     obstacle_radius=ro, want_tail=True)
        g_ref = np.stack([z["g_mean"].reshape(-1), z["g_cvar"].reshape(-1), z["g_dr_cvar"].reshape(-1)], axis=1)
        assert np.abs(res.g - g_ref).max() < 1e-9
        assert np.abs(res.h - z["h_dr_cvar"].reshape(-1, 2)).max() < 1e-12
        assert np.abs(res.h_mean - z["h_mean"].reshape(-1, 2)).max() < 1e-12
        p = dict(alpha=alpha, delta=delta, epsilon=eps, robot_radius=rr, obstacle_radius=ro)
        check_batch(res, samples, ego, p)
        # one call per halfspace on the strided view itself (no host repacking by the caller)
        t, i = 7, n_obs - 1
        view = traj[i][:, t, :]
        assert not view.flags["C_CONTIGUOUS"]
        one = eng.compute_halfspaces(view, x_ref[t][:2], alpha=alpha, delta=delta, epsilon=eps, robot_radius=rr,
                                     obstacle_radius=ro)
        assert np.array_equal(one.g[0], res.g[t * n_obs + i])


def test_golden_timing_sweep_and_explicit_h(eng, golden_dir):
    z = np.load(os.path.join(golden_dir, "timing_sweep.npz"))
    alpha, delta, eps, rr, ro = z["params"]
    for n in z["sizes"]:
        s = z[f"samples_{n}"]
        ref = z[f"out_{n}"]
        res = eng.compute_halfspaces(s, np.zeros(2), alpha=alpha, delta=delta, epsilon=eps, robot_radius=rr,
                                     obstacle_radius=ro, want_tail=True)
        assert np.abs(res.h[0] - ref[0:2]).max() < 1e-12
        assert abs(res.g[0, 2] - ref[2]) < 1e-9 and abs(res.g[0, 1] - ref[5]) < 1e-9 and abs(res.g[0, 0] - ref[8]) < 1e-9
        check_batch(res, s[None], np.zeros(2), dict(alpha=alpha, delta=delta, epsilon=eps, robot_radius=rr, obstacle_radius=ro))
    z = np.load(os.path.join(golden_dir, "explicit_h.npz"))
    for c in range(int(z["n_cases"])):
        alpha, delta, eps, rr, ro, h0, h1 = z[f"in_{c}"]
        g_star, g_tilde, g_cvar = z[f"out_{c}"]
        s = z[f"samples_{c}"]
        res = eng.compute_halfspaces(s, None, alpha=alpha, delta=delta, epsilon=eps, robot_radius=rr,
                                     obstacle_radius=ro, h=(h0, h1), want_tail=True)
        assert abs(res.g_star[0] - g_star) < 1e-9 and abs(res.g[0, 2] - g_tilde) < 1e-9 and abs(res.g[0, 1] - g_cvar) < 1e-9
        check_batch(res, s[None], None, dict(alpha=alpha, delta=delta, epsilon=eps, robot_radius=rr, obstacle_radius=ro),
                    h_in=np.array([h0, h1]))


def test_golden_n10k(eng, golden_dir):
    z = np.load(os.path.join(golden_dir, "n10k.npz"))
    alpha, delta, eps, rr, ro = z["params"]
    p = dict(alpha=alpha, delta=delta, epsilon=eps, robot_radius=rr, obstacle_radius=ro)
    res = eng.compute_halfspaces(z["samples"], z["ego"], want_tail=True, **p)
    ref = z["out"]
    assert np.abs(res.h[0] - ref[0:2]).max() < 1e-12
    assert abs(res.g[0, 2] - ref[2]) < 1e-9 and abs(res.g[0, 1] - ref[5]) < 1e-9 and abs(res.g[0, 0] - ref[8]) < 1e-9
    check_batch(res, z["samples"][None], z["ego"], p)
    assert res.status[0] == 0                                   # window path, no fallback
    res32 = eng.compute_halfspaces(z["samples32"], z["ego"], want_tail=True, **p)
    assert abs(res32.g[0, 2] - z["out32"][2]) < 1e-7
    assert abs(res32.g[0, 2] - ref[2]) < 1e-5                   # fp32-input path vs fp64 truth: 1e-5 m
    check_batch(res32, z["samples32"][None], z["ego"], p)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,alpha", [(10000, 0.1), (4096, 0.05), (2500, 0.3), (1500, 0.2), (1025, 0.5), (777, 0.2),
                                      (512, 0.1), (33, 0.2), (20, 0.2), (2, 0.5), (1, 1.0), (1000, 1.0), (3000, 0.0123)])
def test_random_batches(eng, dtype, n, alpha):
    rng = np.random.RandomState(n * 7 + int(alpha * 1000))
    B = 24 if n >= 1000 else 64
    mu = rng.uniform(-5, 5, size=(B, 1, 2))
    s = (mu + 0.1 * rng.standard_normal((B, n, 2))).astype(dtype)
    ego = rng.uniform(-1, 1, size=(B, 2))
    p = dict(PARAMS, alpha=alpha, epsilon=0.01)
    res = eng.compute_halfspaces(s, ego, want_tail=True, **p)
    check_batch(res, s, ego, p)
    # The timed configuration (no tail output) is not asked for its tail SET: the set is a function of (h, T) alone — the kc
    # largest canonical losses -(h.xi), ties to the lower index — so bit-equal h and T (`var`) with the parity-mode run above,
    # whose set was just compared with the oracle's, pin it by construction.
    # without the tail output (the timed configuration: pipelined kernel where the window applies) the direction, the mean
    # halfspace and the threshold are the same bits; the CVaR sum is taken in another order (last-bit differences)
    res2 = eng.compute_halfspaces(s, ego, **p)
    assert np.array_equal(res2.h, res.h) and np.array_equal(res2.var, res.var) and np.array_equal(res2.g[:, 0], res.g[:, 0])
    assert rel_close(res2.g, res.g, 1e-12 if dtype == np.float64 else 1e-7) and rel_close(res2.cvar, res.cvar, 1e-12 if dtype == np.float64 else 1e-7)
    # forcing the general select or the generic loader changes nothing but the last-bit summation order
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    res3 = eng.compute_halfspaces(s, ego, want_tail=True, flags=_lib.FLAG_GENERAL_ONLY | _lib.FLAG_NO_BULK, **p)
    assert np.array_equal(res3.tail_idx, res.tail_idx) and np.array_equal(res3.var, res.var)
    assert np.array_equal(res3.h, res.h) and rel_close(res3.g, res.g, 1e-12 if dtype == np.float64 else 1e-7)
    assert np.all(res3.status & _lib.STATUS_GENERAL)


def test_window_miss_falls_back_exactly(eng):
    """Non-Gaussian data (bimodal, heavy tails, constants) defeats the statistical window; result must not change."""
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    rng = np.random.RandomState(3)
    n, B = 6000, 12
    s = np.empty((B, n, 2))
    for b in range(B):
        kind = b % 4
        if kind == 0:      # bimodal
            s[b] = np.where(rng.rand(n, 1) < 0.15, 6.0, 2.0) + 0.05 * rng.standard_normal((n, 2))
        elif kind == 1:    # heavy tails
            s[b] = 3.0 + 0.1 * rng.standard_cauchy((n, 2)).clip(-1e3, 1e3)
        elif kind == 2:    # few distinct values (massive ties)
            s[b] = 2.0 + rng.randint(0, 4, size=(n, 2)) * 0.25
        else:              # all identical (obstacle at t = 0, simulation/obstacles.py:63)
            s[b] = np.array([4.0, -1.0])
    ego = np.zeros((B, 2))
    p = dict(PARAMS, alpha=0.1, epsilon=0.01)
    res = eng.compute_halfspaces(s, ego, want_tail=True, **p)
    check_batch(res, s, ego, p)
    assert (res.status[3::4] & _lib.STATUS_GENERAL).all()
    assert np.array_equal(res.tail_idx[3], np.arange(600))
    # the timed configuration (no tail output) hands the sample slot back early; a window miss must then re-fetch the
    # halfspace (bulk and strided loaders) and still give the same answer, also in fp32 and with many halfspaces per CTA
    for flags in (0, _lib.FLAG_NO_BULK):
        r2 = eng.compute_halfspaces(s, ego, flags=flags, **p)
        assert np.array_equal(r2.var, res.var) and rel_close(r2.g, res.g, 1e-12)
    big = np.concatenate([s] * 60).astype(np.float32)          # 720 halfspaces > 2 per CTA of the persistent grid
    ego_b = np.zeros((big.shape[0], 2))
    r3 = eng.compute_halfspaces(big, ego_b, **p)
    r4 = eng.compute_halfspaces(big, ego_b, flags=_lib.FLAG_GENERAL_ONLY, **p)
    assert np.array_equal(r3.var, r4.var) and np.abs(r3.g - r4.g).max() < 1e-6
    assert np.array_equal(r3.var[:12], r3.var[-12:])
    assert (r3.status & _lib.STATUS_GENERAL).any()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_adversarial_scales_and_offsets(eng, dtype):
    """Stress the rigorous fp32 screening bounds: large coordinate offsets, tiny / large spreads, ego next to the mean
    (cancellation in m - ego), explicit non-unit normals, extreme alpha.  Threshold and tail sets must stay exact."""
    rng = np.random.RandomState(23)
    n = 10000
    cases = []
    for offset in (0.0, 50.0, 1.0e3, 1.0e5):
        for sigma in (1.0e-3, 0.1, 5.0):
            for near in (False, True):
                cases.append((offset, sigma, near))
    B = len(cases)
    s = np.empty((B, n, 2))
    ego = np.empty((B, 2))
    for b, (offset, sigma, near) in enumerate(cases):
        ang = rng.uniform(0, 2 * np.pi)
        mu = offset * np.array([np.cos(ang), np.sin(ang)]) + rng.uniform(-1, 1, size=2)
        s[b] = mu + sigma * rng.standard_normal((n, 2)) * np.array([1.0, rng.uniform(0.2, 3.0)])
        ego[b] = mu + (2.0 * sigma * rng.standard_normal(2) if near else -mu)   # next to the mean / at the origin
    s = s.astype(dtype)
    for alpha in (0.1, 0.002, 0.9):
        p = dict(PARAMS, alpha=alpha, epsilon=0.01)
        res = eng.compute_halfspaces(s, ego, want_tail=True, **p)
        fast = eng.compute_halfspaces(s, ego, **p)                  # the timed configuration (early slot release)
        assert np.array_equal(fast.var, res.var) and np.array_equal(fast.h, res.h)
        if alpha == 0.1:
            assert fast.status[cases.index((0.0, 0.1, False))] == 0     # the window path is the one under test
            print('window-path halfspaces:', int((fast.status == 0).sum()), 'of', B)
        for b, (offset, sigma, near) in enumerate(cases):
            o = cf.halfspace(s[b], ego[b], p["alpha"], p["delta"], p["epsilon"], p["robot_radius"], p["obstacle_radius"])
            assert np.array_equal(res.h[b], o.h) and np.array_equal(res.h_mean[b], o.h_mean), (cases[b], alpha)
            assert res.var[b] == o.var, (cases[b], alpha, res.var[b], o.var)
            assert np.array_equal(res.tail_idx[b], o.tail_idx), (cases[b], alpha)
            if dtype == np.float32:
                tol = 1e-6 * max(1.0, sigma)
                assert abs(res.cvar[b] - o.cvar) <= tol and abs(fast.cvar[b] - o.cvar) <= tol, (cases[b], alpha, res.cvar[b], o.cvar)
            else:
                assert rel_close(res.cvar[b], o.cvar) and rel_close(fast.cvar[b], o.cvar), (cases[b], alpha)
    # explicit normals that are far from unit length, and normals with one tiny component
    for h in ((1.0e-3, 2.0e-3), (3.0e3, -1.0e3), (1.0, 1.0e-12), (0.0, -2.5)):
        p = dict(PARAMS, alpha=0.1, epsilon=0.01)
        res = eng.compute_halfspaces(s, None, want_tail=True, h=h, **p)
        for b in range(0, B, 3):
            o = cf.halfspace(s[b], np.zeros(2), p["alpha"], p["delta"], p["epsilon"], p["robot_radius"],
                             p["obstacle_radius"], np.array(h))
            assert res.var[b] == o.var and np.array_equal(res.tail_idx[b], o.tail_idx), (h, cases[b])
            scale = max(abs(h[0]), abs(h[1])) * max(1.0, cases[b][1])
            tol = (1e-6 * max(1.0, scale)) if dtype == np.float32 else 1e-9 * max(1.0, abs(o.cvar))
            assert abs(res.cvar[b] - o.cvar) <= tol, (h, cases[b], res.cvar[b], o.cvar)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_learned_window_on_non_gaussian_samples(eng, dtype):
    """Laplace / uniform noise: the Gaussian window plan misses; after two consecutive misses a CTA learns where the
    threshold sits and goes back to the single-sweep path.  Thresholds stay exact, values within the usual bars."""
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    rng = np.random.RandomState(31)
    n, B = 4096, 296 * 8                        # 8 halfspaces per CTA of the persistent grid
    mu = rng.uniform(1.0, 4.0, size=(B, 1, 2))
    lap = rng.laplace(scale=0.1 / np.sqrt(2), size=(B // 2, n, 2))
    uni = rng.uniform(-0.1 * np.sqrt(3), 0.1 * np.sqrt(3), size=(B - B // 2, n, 2))
    s = (mu + np.concatenate([lap, uni])).astype(dtype)
    s = s[rng.permutation(B)]                   # mixed distributions inside every CTA's chain
    ego = np.zeros((B, 2))
    p = dict(PARAMS, alpha=0.1, epsilon=0.01)
    res = eng.compute_halfspaces(s, ego, **p)
    exact = eng.compute_halfspaces(s, ego, flags=_lib.FLAG_GENERAL_ONLY, **p)
    assert np.array_equal(res.var, exact.var) and np.array_equal(res.h, exact.h)
    tol = 1e-6 if dtype == np.float32 else 1e-9
    assert np.abs(res.g - exact.g).max() <= tol and np.abs(res.cvar - exact.cvar).max() <= tol
    frac_general = float(((res.status & _lib.STATUS_GENERAL) != 0).mean())
    assert 0.0 < frac_general < 0.75, frac_general          # the first halfspaces of a chain miss, the later ones hit
    for b in list(range(0, 8)) + list(range(B - 8, B)):
        o = cf.halfspace(s[b], ego[b], p["alpha"], p["delta"], p["epsilon"], p["robot_radius"], p["obstacle_radius"])
        assert res.var[b] == o.var and np.array_equal(res.h[b], o.h), b
        assert abs(res.cvar[b] - o.cvar) <= tol * max(1.0, abs(o.cvar)), b
    tail = eng.compute_halfspaces(s[:600], ego[:600], want_tail=True, **p)
    assert np.array_equal(tail.var, exact.var[:600])
    for b in (0, 299, 599):
        o = cf.halfspace(s[b], ego[b], p["alpha"], p["delta"], p["epsilon"], p["robot_radius"], p["obstacle_radius"])
        assert np.array_equal(tail.tail_idx[b], o.tail_idx), b


def test_degenerate_direction_and_nonfinite(eng):
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    s = np.tile(np.array([[4.0, 0.0]]), (3, 20, 1))
    ego = np.array([[-4.0, 0.0], [4.0, 0.0], [0.0, 0.0]])
    s[2, 5, 0] = np.nan
    res = eng.compute_halfspaces(s, ego, want_tail=True, **PARAMS)
    assert res.g[0, 2] == pytest.approx(-3.35, abs=1e-15) and res.g[0, 1] == pytest.approx(-3.5, abs=1e-15)
    assert np.array_equal(res.tail_idx[0], np.arange(4))
    assert np.array_equal(res.h[1], [1.0, 0.0]) and (res.status[1] & _lib.STATUS_DEGENERATE)
    assert res.status[2] & _lib.STATUS_NONFINITE and res.g[2, 1] == 100.0 and np.all(res.tail_idx[2] == -1)


def test_device_path_matches_host_path(eng):
    import torch
    rng = np.random.RandomState(11)
    for dtype in (np.float32, np.float64):
        s = (np.array([2.0, 1.0]) + 0.1 * rng.standard_normal((40, 5000, 2))).astype(dtype)
        ego = rng.uniform(-1, 1, size=(40, 2))
        p = dict(PARAMS, alpha=0.1, epsilon=0.01)
        host = eng.compute_halfspaces(s, ego, want_tail=True, **p)
        dev = eng.compute_halfspaces(torch.from_numpy(s).cuda(), torch.from_numpy(ego).cuda(), want_tail=True, **p)
        torch.cuda.synchronize()
        assert np.array_equal(dev.g.cpu().numpy(), host.g)
        assert np.array_equal(dev.h.cpu().numpy(), host.h)
        assert np.array_equal(dev.tail_idx.cpu().numpy(), host.tail_idx)
        # strided device view [N, T, 2] -> [:, t, :]
        big = torch.from_numpy(np.ascontiguousarray(np.repeat(s[0][:, None, :], 5, axis=1))).cuda()
        big[:, 3, :] = torch.from_numpy(s[1]).cuda()
        one = eng.compute_halfspaces(big[:, 3, :], torch.from_numpy(ego[1]).cuda(), **p)
        torch.cuda.synchronize()
        assert np.array_equal(one.g.cpu().numpy()[0], host.g[1])


def test_bad_arguments_raise(eng):
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    s = np.zeros((1, 8, 2))
    with pytest.raises(_lib.DrcvarError):
        eng.compute_halfspaces(s, None, **dict(PARAMS, alpha=0.0))
    with pytest.raises(_lib.DrcvarError):
        eng.compute_halfspaces(s, None, **dict(PARAMS, alpha=1.5))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_large_n_streaming_kernel(eng, dtype):
    """N beyond one CTA's shared memory (BASELINE config 5: N = 100 000) goes through the multi-pass streaming kernel."""
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    rng = np.random.RandomState(17)
    p = dict(PARAMS, alpha=0.1, epsilon=0.01)
    n = 100000
    assert n > eng.max_samples(dtype)
    s = (rng.uniform(-4, 4, size=(3, 1, 2)) + 0.1 * rng.standard_normal((3, n, 2))).astype(dtype)
    ego = rng.uniform(-1, 1, size=(3, 2))
    res = eng.compute_halfspaces(s, ego, want_tail=True, **p)
    check_batch(res, s, ego, p)
    assert res.tail_idx.shape == (3, 10000)
    assert not (res.status & _lib.STATUS_GENERAL).any()          # two-pass window path, not the multi-pass select
    # forced on small N it must agree bit for bit with the shared-memory kernel's general path
    for nn in (10000, 777, 20, 1):
        ss = (rng.uniform(-4, 4, size=(16, 1, 2)) + 0.1 * rng.standard_normal((16, nn, 2))).astype(dtype)
        ee = rng.uniform(-1, 1, size=(16, 2))
        a = eng.compute_halfspaces(ss, ee, want_tail=True, flags=_lib.FLAG_FORCE_STREAMING, **p)
        b = eng.compute_halfspaces(ss, ee, want_tail=True, flags=_lib.FLAG_GENERAL_ONLY, **p)
        assert np.array_equal(a.tail_idx, b.tail_idx) and np.array_equal(a.var, b.var) and np.array_equal(a.h, b.h)
        assert rel_close(a.g, b.g, 1e-12)
        strided = np.repeat(ss[:, :, None, :], 3, axis=2)[:, :, 1, :]          # non-contiguous view
        c = eng.compute_halfspaces(strided, ee, flags=_lib.FLAG_FORCE_STREAMING, **p)
        assert np.array_equal(c.var, a.var)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_sample_counts_around_the_shared_memory_limit(eng, dtype):
    """N = drcvar_max_samples (the largest resident slot, one CTA per SM), one more (first streaming size), and ragged
    sizes just below: same bars on both sides of the kernel switch."""
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    nmax = eng.max_samples(dtype)
    rng = np.random.RandomState(41)
    p = dict(PARAMS, alpha=0.1, epsilon=0.01)
    for n in (nmax, nmax + 1, nmax - 1, nmax - 511):
        s = (rng.uniform(-3, 3, size=(3, 1, 2)) + 0.1 * rng.standard_normal((3, n, 2))).astype(dtype)
        ego = rng.uniform(-1, 1, size=(3, 2))
        res = eng.compute_halfspaces(s, ego, want_tail=True, **p)
        check_batch(res, s, ego, p)
        fast = eng.compute_halfspaces(s, ego, **p)
        assert np.array_equal(fast.var, res.var) and np.array_equal(fast.h, res.h)
        gen = eng.compute_halfspaces(s, ego, flags=_lib.FLAG_GENERAL_ONLY | _lib.FLAG_NO_BULK, **p)
        assert np.array_equal(gen.var, res.var)


def test_run_to_run_determinism(eng):
    rng = np.random.RandomState(5)
    s = (np.array([1.0, 3.0]) + 0.1 * rng.standard_normal((300, 10000, 2))).astype(np.float32)
    p = dict(PARAMS, alpha=0.1, epsilon=0.01)
    a = eng.compute_halfspaces(s, None, **p)
    b = eng.compute_halfspaces(s, None, **p)
    assert np.array_equal(a.g, b.g) and np.array_equal(a.h, b.h) and np.array_equal(a.cvar, b.cvar)
    assert (a.status == 0).all()
