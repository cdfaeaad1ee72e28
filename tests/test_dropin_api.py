"""
Drop-in boundary tests.

CPU part: the drop-in modules expose exactly the names and call signatures of the reference's
core/risk_metrics.py, core/halfspaces.py, core/geometry.py, simulation/environment.py and the utils/timing.py stand-in
(checked against a frozen table, and against the reference itself when /root/reference is present).
GPU part (-m gpu): the reference's call patterns (timing sweep, main.py single scenario) through the drop-in
modules reproduce the golden vectors the reference's own modules produced.
"""
import importlib
import inspect
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200", "dropin")
_TOP = ("core", "utils", "simulation", "config", "evaluation", "_stopwatch")

EXPECTED = {
    "core.risk_metrics": {
        "save_timing_info": "(key, setup_time, solve_time)",
        "expected_value": "(samples)",
        "var_metric": "(samples, alpha)",
        "cvar_metric": "(samples, alpha)",
        "dr_cvar_halfspace": "(samples, h, alpha, delta, epsilon, robot_radius, obstacle_radius)",
        "cvar_halfspace": "(samples, h, alpha, delta, robot_radius, obstacle_radius)",
        "DRCVaROptimizer.__init__": "(self, alpha, epsilon, delta, max_samples)",
        "DRCVaROptimizer.solve": "(self, h, samples, combined_radius)",
        "CVaROptimizer.__init__": "(self, alpha, delta, max_samples)",
        "CVaROptimizer.solve": "(self, h, samples, combined_radius)",
    },
    "core.halfspaces": {
        "SafeHalfspace.__init__": "(self, h, g_tilde)",
        "SafeHalfspace.is_point_safe": "(self, point)",
        "SafeHalfspace.distance_to_boundary": "(self, point)",
        "SafeHalfspace.get_constraint_params": "(self)",
        "MeanSafeHalfspace.create": "(samples, robot_radius, obstacle_radius)",
        "CVaRSafeHalfspace.create": "(samples, ego_ref_pos, alpha, delta, robot_radius, obstacle_radius)",
        "DRCVaRSafeHalfspace.create": "(samples, ego_ref_pos, alpha, delta, epsilon, robot_radius, obstacle_radius)",
        "compute_safe_halfspaces": "(obstacle_samples, ego_ref_pos, robot_radius, obstacle_radius, alpha, delta, epsilon)",
    },
    "core.geometry": {"compute_separating_vector": "(ego_pos, obstacle_pos)"},
    "simulation.environment": {
        "SafetyFilteringEnvironment.__init__": "(self, ROBOT_RADIUS, OBSTACLE_RADIUS, HORIZON, DT, ALPHA, DELTA, EPSILON)",
        "SafetyFilteringEnvironment.set_bounds": "(self, state_bounds=None, input_bounds=None)",
        "SafetyFilteringEnvironment.compute_safe_halfspaces_for_trajectory": "(self, obstacle_sample_trajectories, ego_ref_trajectory)",
        "SafetyFilteringEnvironment.compute_distance_to_collision": "(self, ego_trajectory, obstacle_trajectories)",
    },
    "simulation.obstacles": {
        "generate_nominal_trajectory": "(start_pos, direction, speed, n_steps, dt)",
        "generate_obstacle_sample_trajectories": "(nominal_trajectory, n_samples, noise_cov, dt)",
        "generate_laplace_realization": "(nominal_trajectory, noise_cov, dt)",
        "generate_obstacle_scenarios": "(scenario_config, horizon, dt, n_samples=100)",
    },
    "core.mpc_filter": {
        "MPCSafetyFilter.__init__": "(self, A, B, C, Q, R, horizon, dt)",
        "MPCSafetyFilter.filter_trajectory":
            "(self, x0, x_ref, u_ref, safe_halfspaces, input_constraints=None, position_constraints=None)",
    },
    "_stopwatch": {   # stands in for utils.timing when no reference checkout is behind the overlay
        "Timer.__init__": "(self, name=None)", "Timer.start": "(self)", "Timer.stop": "(self)", "timeit": "(func)",
        "TimingStats.add": "(self, name, time_value)", "TimingStats.get_stats": "(self, name)",
        "TimingStats.print_stats": "(self)",
    },
}


def _purge():
    for k in list(sys.modules):
        if k.split(".")[0] in _TOP:
            del sys.modules[k]


@pytest.fixture()
def dropin(tmp_path, monkeypatch):
    """The drop-in modules imported fresh, cwd in a scratch dir (they write tmp/timing_info_*.json like the reference)."""
    _purge()
    monkeypatch.syspath_prepend(DROPIN)
    monkeypatch.chdir(tmp_path)
    mods = {name: importlib.import_module(name) for name in EXPECTED}
    yield mods
    _purge()


def _sig(mod, dotted):
    obj = mod
    for part in dotted.split("."):
        obj = inspect.getattr_static(obj, part) if inspect.isclass(obj) else getattr(obj, part)
        if isinstance(obj, staticmethod):
            obj = obj.__func__
    return str(inspect.signature(obj))


def test_signatures_match_frozen_table(dropin):
    for modname, table in EXPECTED.items():
        for dotted, sig in table.items():
            assert _sig(dropin[modname], dotted) == sig, (modname, dotted)
    rm = dropin["core.risk_metrics"]
    assert rm.drcvar_optimizer is None and rm.cvar_optimizer is None          # module-global singletons exist
    o = rm.DRCVaROptimizer(0.2, 0.15, 0.1, 20)
    assert (o.alpha, o.epsilon, o.delta, o.n_samples) == (0.2, 0.15, 0.1, 20)


def test_frozen_table_matches_the_reference_itself():
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("reference tree not present")
    _purge()
    with ref_harness.reference_modules() as ref:
        mods = {"core.risk_metrics": ref.risk_metrics, "core.halfspaces": ref.halfspaces, "core.geometry": ref.geometry,
                "simulation.environment": ref.environment, "simulation.obstacles": ref.obstacles,
                "_stopwatch": importlib.import_module("utils.timing"),
                "core.mpc_filter": ref.mpc_filter}
        for modname, table in EXPECTED.items():
            for dotted, sig in table.items():
                assert _sig(mods[modname], dotted) == sig, (modname, dotted)


def test_obstacle_module_host_parts_equal_the_reference(dropin):
    """simulation/obstacles.py drop-in: nominal trajectories and the Laplace realization are the reference's numbers
    (same recurrence, same numpy draws in the same order); the sample trajectories are a lazy array-like."""
    ob = dropin["simulation.obstacles"]
    nom = ob.generate_nominal_trajectory(np.array([5.0, 0.5]), np.array([-3.0, 0.1]), 0.7, 40, 0.1)
    assert nom.shape == (41, 2) and np.array_equal(nom[0], [5.0, 0.5])
    still = ob.generate_nominal_trajectory(np.array([1.0, 2.0]), np.array([0.0, 0.0]), 1.0, 5, 0.1)
    assert np.array_equal(still, np.tile([1.0, 2.0], (6, 1)))
    np.random.seed(3)
    lazy = ob.generate_obstacle_sample_trajectories(nom, 64, np.diag([0.01, 0.04]), 0.1)
    assert lazy.shape == (64, 41, 2) and len(lazy) == 64 and lazy.ndim == 3 and not lazy.materialised
    mean, chol = lazy.kernel_inputs(30)
    assert np.array_equal(mean, nom[:30]) and np.array_equal(chol[0], [0, 0, 0]) and np.allclose(chol[1], [0.1, 0.0, 0.2])
    np.random.seed(3)
    assert ob.generate_obstacle_sample_trajectories(nom, 64, np.diag([0.01, 0.04]), 0.1).key == lazy.key   # np.random.seed reproduces it
    data = ob.generate_obstacle_scenarios({"obstacles": [{"start": np.array([4.0, 1.0]), "direction": np.array([-1.0, 0.0])},
                                                          {"start": np.array([4.0, -1.0]), "direction": np.array([-1.0, 0.2]),
                                                           "speed": 0.5}]}, 3.0, 0.1, n_samples=32)
    assert [len(v) for v in data.values()] == [2, 2, 2] and data["sample_trajectories"][1].shape == (32, 31, 2)
    assert data["sample_trajectories"][0].key != data["sample_trajectories"][1].key
    from oracle import ref_harness
    if not ref_harness.available():
        return
    _purge()
    with ref_harness.reference_modules() as ref:
        assert np.array_equal(ref.obstacles.generate_nominal_trajectory(np.array([5.0, 0.5]), np.array([-3.0, 0.1]), 0.7, 40, 0.1), nom)
        np.random.seed(11)
        want = ref.obstacles.generate_laplace_realization(nom, np.diag([0.01, 0.04]), 0.1)
    np.random.seed(11)
    assert np.array_equal(ob.generate_laplace_realization(nom, np.diag([0.01, 0.04]), 0.1), want)


def test_dead_helpers_keep_reference_semantics(dropin):
    rm = dropin["core.risk_metrics"]
    x = np.arange(20.0)
    assert rm.var_metric(x, 0.2) == 15.0 and rm.cvar_metric(x, 0.2) == 17.0      # top floor(aN)+1 samples
    assert np.allclose(rm.expected_value(np.ones((4, 2))), [1.0, 1.0])
    g = dropin["core.geometry"]
    assert np.array_equal(g.compute_separating_vector(np.zeros(2), np.zeros(2)), [1.0, 0.0])
    assert np.allclose(g.compute_separating_vector(np.array([1.0, 1.0]), np.array([4.0, 5.0])), [0.6, 0.8])


def test_timer_and_stats(dropin, capsys):
    t = dropin["_stopwatch"]
    with t.Timer("X") as tm:
        pass
    assert "X: " in capsys.readouterr().out and tm.elapsed >= 0
    s = t.TimingStats()
    s.add("a", 1.0); s.add("a", 3.0)
    assert s.get_stats("a")["mean"] == 2.0 and s.get_stats("b") is None


# ------------------------------------------------------------------------------------------------ GPU part
@pytest.mark.gpu
def test_timing_sweep_call_pattern(dropin, golden_dir, capsys):
    z = np.load(os.path.join(golden_dir, "timing_sweep.npz"))
    alpha, delta, eps, rr, ro = (float(v) for v in z["params"])
    H = dropin["core.halfspaces"]
    for n in z["sizes"]:
        s, ref = z[f"samples_{n}"], z[f"out_{n}"]
        ego = np.array([0.0, 0.0])
        dr = H.DRCVaRSafeHalfspace.create(s, ego, alpha, delta, eps, rr, ro)          # evaluation/timing_analysis.py:74
        cv = H.CVaRSafeHalfspace.create(s, ego, alpha, delta, rr, ro)                 # evaluation/timing_analysis.py:101
        mn = H.MeanSafeHalfspace.create(s, rr, ro)
        assert np.abs(dr.h - ref[0:2]).max() < 1e-12 and abs(dr.g_tilde - ref[2]) < 1e-9
        assert np.abs(cv.h - ref[3:5]).max() < 1e-12 and abs(cv.g_tilde - ref[5]) < 1e-9
        assert np.abs(mn.h - ref[6:8]).max() < 1e-12 and abs(mn.g_tilde - ref[8]) < 1e-9
        assert set(dr.info) >= {"setup_time", "solve_time"} and set(mn.info) == {"setup_time", "solve_time", "solve_call_time"}
        for key in ("drcvar", "cvar"):                                                # side channel read by the sweep (:84,:111)
            with open(f"tmp/timing_info_{key}.json") as f:
                assert set(json.load(f)) == {"setup_time", "solve_time"}
        h, g = dr.get_constraint_params()
        assert isinstance(g, float) and h.shape == (2,) and bool(dr.is_point_safe(np.array([-5.0, 0.0])))
    out = capsys.readouterr().out
    assert "DR-CVaR Optimization:" in out and "create:" in out and "DEBUG - Saved drcvar timing" in out


@pytest.mark.gpu
def test_explicit_h_entry_points(dropin, golden_dir):
    z = np.load(os.path.join(golden_dir, "explicit_h.npz"))
    rm = dropin["core.risk_metrics"]
    for c in range(int(z["n_cases"])):
        alpha, delta, eps, rr, ro, h0, h1 = (float(v) for v in z[f"in_{c}"])
        g_star, g_tilde, g_cvar = z[f"out_{c}"]
        a, b = rm.dr_cvar_halfspace(z[f"samples_{c}"], np.array([h0, h1]), alpha, delta, eps, rr, ro)
        assert abs(a - g_star) < 1e-9 and abs(b - g_tilde) < 1e-9
        assert abs(rm.cvar_halfspace(z[f"samples_{c}"], np.array([h0, h1]), alpha, delta, rr, ro) - g_cvar) < 1e-9
    bad = np.full((8, 2), np.nan)
    assert rm.cvar_halfspace(bad, np.array([1.0, 0.0]), 0.25, 0.1, 0.3, 0.3) == 100.0       # failure sentinel


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["head_on_seed42.npz", "multi_obstacle_seed42.npz"])
def test_trajectory_driver_matches_reference(dropin, golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    alpha, delta, eps, rr, ro, horizon = (float(v) for v in z["params"])
    env_mod = dropin["simulation.environment"]
    env = env_mod.SafetyFilteringEnvironment(rr, ro, int(horizon), 0.2, alpha, delta, eps)
    traj = [z["sample_trajectories"][i] for i in range(z["sample_trajectories"].shape[0])]
    hs = env.compute_safe_halfspaces_for_trajectory(traj, z["x_ref"])                     # main.py:95
    n_steps, n_obs = z["g_mean"].shape
    assert len(hs["mean"]) == n_steps and all(len(r) == n_obs for r in hs["dr_cvar"])
    for metric in ("mean", "cvar", "dr_cvar"):
        for t in range(n_steps):
            for i in range(n_obs):
                h, g = hs[metric][t][i].get_constraint_params()
                assert np.abs(h - z[f"h_{metric}"][t, i]).max() < 1e-12
                assert abs(g - z[f"g_{metric}"][t, i]) < 1e-9
    # the per-step API gives the same objects (core/halfspaces.py:196 called from environment.py:95)
    H = dropin["core.halfspaces"]
    t = 5
    row = H.compute_safe_halfspaces([tr[:, t, :] for tr in traj], z["x_ref"][t][:2], rr, ro, alpha, delta, eps)
    for i in range(n_obs):
        assert row["dr_cvar"][i].g_tilde == hs["dr_cvar"][t][i].g_tilde
        assert np.array_equal(row["cvar"][i].h, hs["cvar"][t][i].h)


@pytest.mark.gpu
def test_nominal_trajectory_extension(dropin, golden_dir):
    """compute_safe_halfspaces_for_nominal (SURVEY §8-f2 extension): same structure as the trajectory driver, values equal
    to the halfspaces of the samples the generator is specified to draw (oracle/sample_gen.py)."""
    from oracle import closed_form as cf
    from oracle import sample_gen as sg
    env_mod = dropin["simulation.environment"]
    z = np.load(os.path.join(golden_dir, "multi_obstacle_seed42.npz"))
    alpha, delta, eps, rr, ro, horizon = (float(v) for v in z["params"])
    env = env_mod.SafetyFilteringEnvironment(rr, ro, int(horizon), 0.2, alpha, delta, eps)
    traj = z["sample_trajectories"]                               # [n_obs, N, T+1, 2]: nominal = sample mean at t = 0
    n_obs, T1 = traj.shape[0], traj.shape[2]
    nominal = [np.stack([traj[i][0, 0, :] + 0.05 * t * np.array([1.0, -0.5]) for t in range(T1)]) for i in range(n_obs)]
    cov = np.array([[0.01, 0.002], [0.002, 0.01]])
    n, seed = 2000, 17
    res = env.compute_safe_halfspaces_for_nominal(nominal, z["x_ref"], cov, n, seed=seed)
    n_steps = min(len(z["x_ref"]), int(horizon))
    assert set(res) == {"mean", "cvar", "dr_cvar"} and len(res["dr_cvar"]) == n_steps and len(res["cvar"][0]) == n_obs
    chol = sg.cholesky2(cov)
    for t in (0, 1, n_steps - 1):
        for i in range(n_obs):
            b = t * n_obs + i
            s = sg.generate(nominal[i][t][None], np.zeros((1, 3)) if t == 0 else chol[None], n, seed, index_offset=b)[0]
            o = cf.halfspace(s, z["x_ref"][t][:2], alpha, delta, eps, rr, ro)
            hs = res["dr_cvar"][t][i]
            assert np.array_equal(hs.h, o.h) and abs(hs.g_tilde - o.g_dr) <= 1e-6
            assert abs(res["cvar"][t][i].g_tilde - o.g_cvar) <= 1e-6
            assert abs(res["mean"][t][i].g_tilde - o.g_mean) <= 1e-9


@pytest.mark.gpu
def test_obstacle_dropin_feeds_the_generate_mode(dropin):
    """main.py's flow with the drop-in simulation/obstacles.py: generate_obstacle_scenarios -> lazy sample trajectories ->
    SafetyFilteringEnvironment.compute_safe_halfspaces_for_trajectory draws the samples inside the kernel.  The result
    must equal (a) the oracle on the samples the generator is specified to draw (oracle/sample_gen.py), (b) the ordinary
    stored-sample path fed with the materialised array (what the reference's visualisation code would read)."""
    from oracle import closed_form as cf
    from oracle import sample_gen as sg
    ob, env_mod = dropin["simulation.obstacles"], dropin["simulation.environment"]
    alpha, delta, eps, rr, ro, horizon, dt, n = 0.1, 0.1, 0.01, 0.3, 0.3, 12, 0.2, 4000
    env = env_mod.SafetyFilteringEnvironment(rr, ro, horizon, dt, alpha, delta, eps)
    np.random.seed(42)
    data = ob.generate_obstacle_scenarios({"obstacles": [
        {"start": np.array([4.0, 0.6]), "direction": np.array([-1.0, 0.0]), "speed": 0.8},
        {"start": np.array([3.0, -2.0]), "direction": np.array([-0.3, 1.0])}]}, 4.0, dt, n_samples=n)
    lazy = data["sample_trajectories"]
    x_ref = np.stack([np.array([0.25 * t, 0.0, 1.25, 0.0]) for t in range(21)])
    hs = env.compute_safe_halfspaces_for_trajectory(lazy, x_ref)                     # main.py:95
    assert not any(tr.materialised for tr in lazy)                                   # nothing was stored on the host
    assert len(hs["dr_cvar"]) == horizon and len(hs["cvar"][0]) == 2
    cov = np.diag([0.01, 0.01])
    for i, tr in enumerate(lazy):
        for t in (0, 1, horizon - 1):
            chol = np.zeros((1, 3)) if t == 0 else sg.cholesky2(cov)[None]
            s = sg.generate(data["nominal_trajectories"][i][t][None], chol, n, tr.key, index_offset=t)[0]
            o = cf.halfspace(s, x_ref[t][:2], alpha, delta, eps, rr, ro)
            assert np.array_equal(hs["dr_cvar"][t][i].h, o.h)
            assert abs(hs["dr_cvar"][t][i].g_tilde - o.g_dr) <= 1e-6 and abs(hs["cvar"][t][i].g_tilde - o.g_cvar) <= 1e-6
            assert abs(hs["mean"][t][i].g_tilde - o.g_mean) <= 1e-9
    per_run = env.compute_safe_halfspaces_for_runs([lazy, lazy[:1]], x_ref)         # Monte-Carlo entry: still nothing stored
    assert not any(tr.materialised for tr in lazy) and len(per_run) == 2 and len(per_run[1]["cvar"][0]) == 1
    assert per_run[0]["dr_cvar"][3][1].g_tilde == hs["dr_cvar"][3][1].g_tilde
    dense = [np.asarray(tr) for tr in lazy]                                          # materialised through the kernel's dump
    assert dense[0].shape == (n, 21, 2) and dense[0].dtype == np.float64 and all(tr.materialised for tr in lazy)
    assert np.array_equal(dense[1][:, 0, :], np.tile(data["nominal_trajectories"][1][0], (n, 1)))
    assert np.array_equal(lazy[0][:, 3, :], dense[0][:, 3, :])                       # indexing like environment.py:88
    s3 = sg.generate(data["nominal_trajectories"][0][3][None], sg.cholesky2(cov)[None], n, lazy[0].key, index_offset=3)[0]
    assert np.array_equal(dense[0][:, 3, :].astype(np.float32), s3)                  # the specified stream, bit for bit
    again = env.compute_safe_halfspaces_for_trajectory(dense, x_ref)                 # stored-sample path (fp64 arrays)
    for t in range(1, horizon):
        for i in range(2):
            assert abs(again["dr_cvar"][t][i].g_tilde - hs["dr_cvar"][t][i].g_tilde) <= 1e-6
            assert np.abs(again["dr_cvar"][t][i].h - hs["dr_cvar"][t][i].h).max() <= 1e-9   # fp64 mean vs shifted fp32 lane sums


@pytest.mark.parametrize("name, cfg", [
    ("head_on_seed42.npz", {"obstacle_start": np.array([4.0, 0.0]), "obstacle_direction": np.array([-1.0, 0.0])}),
    ("multi_obstacle_seed42.npz", {"obstacles": [
        {"start": np.array([0.0, 2.0]), "direction": np.array([0.0, -0.5]), "speed": 0.8},
        {"start": np.array([-3.0, 0.5]), "direction": np.array([0.7, 0.0]), "speed": 0.6},
        {"start": np.array([1.5, -2.0]), "direction": np.array([-0.2, 0.5]), "speed": 0.7}]}),
])
def test_reference_stream_mode_reproduces_the_reference_arrays(dropin, golden_dir, name, cfg):
    """SURVEY §8-f2 'a seeded mode that reproduces host arrays for parity': with the reference-stream mode on, np.random.seed(42)
    (main.py:191) + generate_obstacle_scenarios give the SAME float64 sample trajectories the reference's own
    simulation/obstacles.py produced (tests/golden/make_golden.py stored their first HORIZON+1 steps), bit for bit; scenario
    parameters: config/scenarios.py:21-28,50-65, config/parameters.py:17,25,26 (SIM_TIME 30 s, DT 0.2 s, 20 samples)."""
    ob = dropin["simulation.obstacles"]
    want = np.load(os.path.join(golden_dir, name))["sample_trajectories"]          # [n_obs, N, H+1, 2]
    ob.REFERENCE_STREAM = True
    try:
        np.random.seed(42)
        data = ob.generate_obstacle_scenarios(cfg, 30.0, 0.2, 20)
    finally:
        ob.REFERENCE_STREAM = None
    got = data["sample_trajectories"]
    assert len(got) == want.shape[0]
    for k, tr in enumerate(got):
        assert isinstance(tr, np.ndarray) and tr.dtype == np.float64 and tr.shape == (20, 151, 2)
        assert np.array_equal(tr[:, : want.shape[2], :], want[k]), (name, k)
    # and the default mode stays lazy
    np.random.seed(42)
    assert not isinstance(ob.generate_obstacle_scenarios(cfg, 30.0, 0.2, 20)["sample_trajectories"][0], np.ndarray)


@pytest.mark.gpu
def test_reference_stream_mode_end_to_end(dropin, golden_dir):
    """main.py's flow with the seeded reference-stream mode: np.random.seed(42) -> generate_obstacle_scenarios ->
    compute_safe_halfspaces_for_trajectory (main.py:61,95) gives the halfspaces the unmodified reference produced from ITS
    draws of the same seed (tests/golden/multi_obstacle_seed42.npz)."""
    z = np.load(os.path.join(golden_dir, "multi_obstacle_seed42.npz"))
    alpha, delta, eps, rr, ro, horizon = (float(v) for v in z["params"])
    ob, env_mod = dropin["simulation.obstacles"], dropin["simulation.environment"]
    cfg = {"obstacles": [{"start": np.array([0.0, 2.0]), "direction": np.array([0.0, -0.5]), "speed": 0.8},
                         {"start": np.array([-3.0, 0.5]), "direction": np.array([0.7, 0.0]), "speed": 0.6},
                         {"start": np.array([1.5, -2.0]), "direction": np.array([-0.2, 0.5]), "speed": 0.7}]}
    ob.REFERENCE_STREAM = True
    try:
        np.random.seed(42)
        data = ob.generate_obstacle_scenarios(cfg, 30.0, 0.2, 20)
    finally:
        ob.REFERENCE_STREAM = None
    env = env_mod.SafetyFilteringEnvironment(rr, ro, int(horizon), 0.2, alpha, delta, eps)
    hs = env.compute_safe_halfspaces_for_trajectory(data["sample_trajectories"], z["x_ref"])
    n_steps, n_obs = z["g_mean"].shape
    for metric in ("mean", "cvar", "dr_cvar"):
        for t in range(n_steps):
            for i in range(n_obs):
                h, g = hs[metric][t][i].get_constraint_params()
                assert np.abs(h - z[f"h_{metric}"][t, i]).max() < 1e-12 and abs(g - z[f"g_{metric}"][t, i]) < 1e-9
