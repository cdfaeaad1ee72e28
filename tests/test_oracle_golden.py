"""
CPU tests: the closed-form oracle (oracle/closed_form.py) against
  * the committed golden vectors produced by the reference's own modules (tests/golden/make_golden.py),
  * an independent HiGHS restatement of both LPs (oracle/lp_highs.py),
  * the analytic known answers of SURVEY.md §8-c (t = 0 rows).
Tolerance: 1e-9 absolute on offsets (HiGHS optimum vs closed form agree to ~1e-15 in practice),
1e-12 on directions.
"""
import os

import numpy as np
import pytest

from oracle import closed_form as cf
from oracle import lp_highs

TOL_G = 1e-9
TOL_H = 1e-12


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("name", ["head_on_seed42.npz", "multi_obstacle_seed42.npz"])
def test_scenarios_match_reference(golden_dir, name):
    z = _load(golden_dir, name)
    alpha, delta, eps, rr, ro, horizon = z["params"]
    traj = z["sample_trajectories"]          # [n_obs, N, H+1, 2]
    x_ref = z["x_ref"]
    res = cf.trajectory_halfspaces(list(traj), x_ref, int(horizon), alpha, delta, eps, rr, ro)
    assert len(res) == z["g_mean"].shape[0]
    for t, row in enumerate(res):
        for i, o in enumerate(row):
            assert np.abs(o.h - z["h_dr_cvar"][t, i]).max() < TOL_H
            assert np.abs(o.h - z["h_cvar"][t, i]).max() < TOL_H
            assert np.abs(o.h_mean - z["h_mean"][t, i]).max() < TOL_H
            assert abs(o.g_dr - z["g_dr_cvar"][t, i]) < TOL_G
            assert abs(o.g_cvar - z["g_cvar"][t, i]) < TOL_G
            assert abs(o.g_mean - z["g_mean"][t, i]) < TOL_G


def test_survey_known_answers(golden_dir):
    z = _load(golden_dir, "head_on_seed42.npz")
    # t = 0: all samples identical -> analytic
    assert abs(z["g_dr_cvar"][0, 0] - (-3.35)) < 1e-12
    assert abs(z["g_cvar"][0, 0] - (-3.5)) < 1e-12
    assert abs(z["g_mean"][0, 0] - (-3.4)) < 1e-12
    assert abs(z["g_dr_cvar"][1, 0] - (-3.019526038668)) < 1e-9
    assert abs(z["g_dr_cvar"][29, 0] - (-0.987924508451)) < 1e-9
    m = _load(golden_dir, "multi_obstacle_seed42.npz")
    assert abs(m["g_dr_cvar"][15, 2] - 1.537017938097) < 1e-9
    assert abs(m["g_mean"][29, 1] - (-0.066492693416)) < 1e-9
    # invariant: g~_dr - g_cvar = eps/alpha - (r_r + r_o) for unit h
    assert np.allclose(m["g_dr_cvar"] - m["g_cvar"], 0.15 / 0.2 - 0.6, atol=1e-9)


def test_timing_sweep_matches_reference(golden_dir):
    z = _load(golden_dir, "timing_sweep.npz")
    alpha, delta, eps, rr, ro = z["params"]
    for n in z["sizes"]:
        s = z[f"samples_{n}"]
        ref = z[f"out_{n}"]
        o = cf.halfspace(s, (0.0, 0.0), alpha, delta, eps, rr, ro)
        assert np.abs(o.h - ref[0:2]).max() < TOL_H
        assert abs(o.g_dr - ref[2]) < TOL_G
        assert np.abs(o.h - ref[3:5]).max() < TOL_H
        assert abs(o.g_cvar - ref[5]) < TOL_G
        assert np.abs(o.h_mean - ref[6:8]).max() < TOL_H
        assert abs(o.g_mean - ref[8]) < TOL_G


def test_explicit_h_cases_match_reference(golden_dir):
    z = _load(golden_dir, "explicit_h.npz")
    for c in range(int(z["n_cases"])):
        alpha, delta, eps, rr, ro, h0, h1 = z[f"in_{c}"]
        g_star, g_tilde, g_cvar = z[f"out_{c}"]
        o = cf.halfspace(z[f"samples_{c}"], (0.0, 0.0), alpha, delta, eps, rr, ro, h_in=(h0, h1))
        assert abs(o.g_dr_star - g_star) < TOL_G, c
        assert abs(o.g_dr - g_tilde) < TOL_G, c
        assert abs(o.g_cvar - g_cvar) < TOL_G, c


def test_n10k_matches_reference(golden_dir):
    z = _load(golden_dir, "n10k.npz")
    alpha, delta, eps, rr, ro = z["params"]
    ref = z["out"]
    o = cf.halfspace(z["samples"], z["ego"], alpha, delta, eps, rr, ro)
    assert np.abs(o.h - ref[0:2]).max() < TOL_H
    assert abs(o.g_dr - ref[2]) < TOL_G
    assert abs(o.g_cvar - ref[5]) < TOL_G
    assert abs(o.g_mean - ref[8]) < TOL_G
    assert len(o.tail_idx) == 1000
    # fp32-input path: fp32 lane partial sums for the mean (canonical contract), everything else on promoted samples
    o32 = cf.halfspace(z["samples32"], z["ego"], alpha, delta, eps, rr, ro)
    assert np.abs(o32.h - z["out32"][0:2]).max() < 1e-7
    assert abs(o32.g_dr - z["out32"][2]) < 1e-7           # reference run on the same (promoted) fp32 samples
    # and the fp32-input result is within 1e-5 m of the fp64 one (north-star tolerance)
    assert abs(o32.g_dr - o.g_dr) < 1e-5


@pytest.mark.parametrize("n,alpha", [(10, 0.2), (23, 0.2), (3, 0.2), (100, 0.05), (257, 0.37), (64, 1.0), (500, 0.1)])
def test_closed_form_equals_lp(n, alpha):
    rng = np.random.RandomState(n)
    s = rng.normal(size=(n, 2)) * 0.2 + np.array([1.0, 2.0])
    ego = np.array([-0.5, 0.3])
    delta, eps, rr, ro = 0.1, 0.15, 0.3, 0.3
    o = cf.halfspace(s, ego, alpha, delta, eps, rr, ro)
    hxi = cf.projection(o.h, s)
    r = (rr + ro) * cf.norm2(*o.h)
    ok, g_star = lp_highs.drcvar_lp(hxi, r, alpha, eps, delta)
    assert ok and abs(g_star - o.g_dr_star) < TOL_G
    ok, g = lp_highs.cvar_lp(hxi, r, alpha, delta)
    assert ok and abs(g - o.g_cvar) < TOL_G


def test_tail_index_definition():
    rng = np.random.RandomState(0)
    L = rng.normal(size=200)
    L[10] = L[20] = L[30] = np.sort(L)[-5]          # three-way tie at the boundary region
    for alpha in (0.02, 0.025, 0.1, 0.33, 1.0):
        cvar, T, idx, k_f, kc = cf.tail_select(L, alpha)
        expect = np.sort(np.argsort(-L, kind="stable")[:kc])
        assert np.array_equal(idx, expect)
        srt = np.sort(L)[::-1]
        k = int(np.floor(k_f))
        lp_val = (srt[:k].sum() + ((k_f - k) * srt[k] if k < len(L) and k_f > k else 0.0)) / k_f
        assert abs(cvar - lp_val) < 1e-12


def test_all_ties_and_degenerate_direction():
    s = np.tile(np.array([[4.0, 0.0]]), (20, 1))
    o = cf.halfspace(s, (-4.0, 0.0), 0.2, 0.1, 0.15, 0.3, 0.3)
    assert np.array_equal(o.tail_idx, np.arange(4))
    assert o.g_dr == pytest.approx(-3.35, abs=1e-15)
    # obstacle mean on top of the ego -> fallback direction [1, 0] (core/geometry.py:49-51)
    o = cf.halfspace(s, (4.0, 0.0), 0.2, 0.1, 0.15, 0.3, 0.3)
    assert np.array_equal(o.h, [1.0, 0.0])


def test_nonfinite_gives_sentinel():
    s = np.random.RandomState(1).normal(size=(16, 2))
    s[3, 1] = np.nan
    o = cf.halfspace(s, (0.0, 0.0), 0.25, 0.1, 0.15, 0.3, 0.3, h_in=(1.0, 0.0))
    assert o.nonfinite and o.g_cvar == 100.0 and o.g_dr == pytest.approx(100.0 - 0.6)


def test_canonical_sum_properties():
    rng = np.random.RandomState(5)
    for n in (1, 31, 32, 511, 512, 513, 10000):
        v = rng.normal(size=n)
        assert abs(cf.canonical_sum(v) - float(np.sum(v))) <= 1e-12 * max(1.0, np.abs(v).sum())
        v32 = (5.0 + 0.1 * v).astype(np.float32)
        assert abs(cf.canonical_mean_1d(v32) - float(np.mean(v32.astype(np.float64)))) <= 1e-7
        assert abs(cf.canonical_mean_1d(v) - float(np.mean(v))) <= 1e-13
    # exactly representable data: any order gives the same bits
    v = rng.randint(-1000, 1000, size=5000).astype(np.float64)
    assert cf.canonical_sum(v) == float(v.sum())


def test_tail_count_snapping():
    assert cf.tail_count(0.1, 10000) == (1000.0, 1000)
    assert cf.tail_count(0.2, 20) == (4.0, 4)
    assert cf.tail_count(0.1, 100000) == (10000.0, 10000)
    for n in (10, 50, 100, 500, 1000, 1500):
        k_f, kc = cf.tail_count(0.2, n)
        assert k_f == n // 5 and kc == n // 5
    k_f, kc = cf.tail_count(0.2, 23)
    assert kc == 5 and abs(k_f - 4.6) < 1e-12
    assert cf.tail_count(0.2, 3)[1] == 1
    with pytest.raises(ValueError):
        cf.tail_count(0.0, 10)
    with pytest.raises(ValueError):
        cf.tail_count(1.5, 10)


def test_octant_rule_of_the_canonical_mean():
    """N > 32768: 8 octants of whole 4 KB rows, each with its own slot sums + tree (the cluster kernel's split).
    Checked against a direct, loop-level restatement and against the plain mean."""
    rng = np.random.RandomState(12)
    for dtype, n in ((np.float32, 32769), (np.float32, 100000), (np.float64, 50001), (np.float32, 40000)):
        v = (2.5 + 0.1 * rng.standard_normal(n)).astype(dtype)
        item = np.dtype(dtype).itemsize
        ol = cf.octant_len(n, item)
        assert (ol * 2 * item) % 4096 == 0 and 8 * ol >= n and 8 * (ol - 4096 // (2 * item)) < n
        lanes = 1024 if dtype == np.float32 else 512
        tot = []
        for k in range(8):
            part = v[k * ol:(k + 1) * ol]
            acc = np.zeros(lanes, dtype=dtype)
            d = (part - v[0]).astype(np.float32) if dtype == np.float32 else part
            for i in range(part.shape[0]):          # lane index relative to the octant start, increasing i
                acc[i % lanes] = dtype(acc[i % lanes] + d[i])
            s = acc.astype(np.float64)
            if dtype == np.float32:
                s = s[0::2] + s[1::2]
            tot.append(cf._tree512(s))
        S = ((tot[0] + tot[1]) + (tot[2] + tot[3])) + ((tot[4] + tot[5]) + (tot[6] + tot[7]))
        want = (float(v[0]) + S / n) if dtype == np.float32 else S / n
        got = cf.canonical_mean_1d(v)
        assert got == want
        assert abs(got - float(np.mean(v.astype(np.float64)))) < 1e-12
    # N <= 32768 keeps the single-chain rule (the resident kernel's sizes): unchanged goldens
    v = (1.0 + 0.1 * rng.standard_normal(32768)).astype(np.float32)
    assert cf.canonical_mean_1d(v) == float(np.float64(v[0]) + np.float64(cf._slot_total(v, v[0])) / np.float64(32768))
