"""
Monte-Carlo evaluation driver (dropin/evaluation/monte_carlo.py).  The reference's own
evaluation/monte_carlo.py is deleted (only a .pyc survives), so there is no golden output: the tests pin the driver
against a plain run-by-run loop of the single-scenario flow (main.py:37-147).

  * CPU, reference tree present: driver (all runs generated first, halfspaces of all runs at once, drop-in MPC) ==
    run-by-run loop with the REFERENCE's MPC filter (through the cvxpy shim) on the same seed: same minimum distances;
  * GPU: compute_safe_halfspaces_for_runs (one launch for all runs) == the per-run method, bit for bit; the driver runs
    end to end with stand-ins for the reference's obstacle generator / planner (absent on the GPU box), and fanning the
    MPC QPs out over threads changes nothing.
"""
import importlib
import os
import sys
import types

import numpy as np
import pytest

from oracle import closed_form as cf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200", "dropin")
SHIM = os.path.join(ROOT, "oracle", "cvxpy_shim")
REF = "/root/reference"
_TOP = ("core", "utils", "simulation", "config", "evaluation", "cvxpy")


def _purge():
    for k in list(sys.modules):
        if k.split(".")[0] in _TOP:
            del sys.modules[k]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "core")), reason="reference tree not present (GPU box)")
def test_driver_matches_a_run_by_run_loop(tmp_path, monkeypatch):
    _purge()
    for p in (REF, SHIM, DROPIN):                       # drop-in first, then the shim, then the reference
        monkeypatch.syspath_prepend(p)
    monkeypatch.chdir(tmp_path)
    try:
        # the overlay WITHOUT its simulation/obstacles.py (dropin/README.md: optional) — the reference's generator with stored
        # sample arrays: no GPU here to draw lazy samples with
        import importlib.util
        spec = importlib.util.spec_from_file_location("simulation.obstacles", os.path.join(REF, "simulation", "obstacles.py"))
        ref_obstacles = importlib.util.module_from_spec(spec)
        sys.modules["simulation.obstacles"] = ref_obstacles
        spec.loader.exec_module(ref_obstacles)
        env_mod = importlib.import_module("simulation.environment")          # drop-in
        mc = importlib.import_module("evaluation.monte_carlo")               # drop-in
        params = importlib.import_module("config.parameters")                # reference
        scenarios = importlib.import_module("config.scenarios")
        obstacles = importlib.import_module("simulation.obstacles")
        planner_mod = importlib.import_module("simulation.planner")
        assert env_mod.__file__.startswith(DROPIN) and mc.__file__.startswith(DROPIN) and obstacles.__file__.startswith(REF)

        class OracleEnv(env_mod.SafetyFilteringEnvironment):
            """No GPU here: the halfspaces of all runs from the closed-form oracle, wrapped like the kernel's."""

            def compute_safe_halfspaces_for_runs(self, runs_samples, x_ref):
                out = []
                for trajs in runs_samples:
                    rows = cf.trajectory_halfspaces(trajs, x_ref, self.HORIZON, self.ALPHA, self.DELTA, self.EPSILON,
                                                    self.ROBOT_RADIUS, self.OBSTACLE_RADIUS)
                    n_steps, n_obs = len(rows), len(trajs)
                    h = np.array([[o.h for o in row] for row in rows])
                    hm = np.array([[o.h_mean for o in row] for row in rows])
                    g = np.array([[(o.g_mean, o.g_cvar, o.g_dr) for o in row] for row in rows])
                    out.append(self._wrap(h, hm, g, n_steps, n_obs, 0.0))
                return out

        cfg = scenarios.get_scenario_config("multi_obstacle")
        P = params
        n_runs = 3
        env = OracleEnv(P.ROBOT_RADIUS, P.OBSTACLE_RADIUS, P.HORIZON, P.DT, P.ALPHA, P.DELTA, P.EPSILON)
        np.random.seed(7)
        res = mc.run_monte_carlo_simulation(env, cfg, n_runs, P)
        assert set(res) == {"min_distances", "collision_counts", "collision_probs", "timing_stats"}
        assert set(res["min_distances"]) == {"reference", "mean", "cvar", "dr_cvar"}
        assert all(len(v) == n_runs for v in res["min_distances"].values())

        # run-by-run loop of main.py's flow with the REFERENCE's MPC filter (cvxpy shim), same seed
        _purge_mpc = sys.modules.pop("core.mpc_filter")
        sys.path.remove(DROPIN)
        try:
            ref_mpc = importlib.import_module("core.mpc_filter")
            assert ref_mpc.__file__.startswith(REF)
        finally:
            sys.path.insert(0, DROPIN)
        A, B, C = env.A, env.B, env.C
        Q, R = P.Q_WEIGHT * np.eye(4), P.R_WEIGHT * np.eye(2)
        x0 = np.zeros(4)
        x0[:2] = cfg["ego_start"]
        x_ref, u_ref, _ = planner_mod.ReferenceTrajectoryPlanner(A, B, C, Q, R, P.HORIZON, P.DT).straight_line_trajectory(
            cfg["ego_start"], cfg["ego_goal"])
        sb = (np.array([-10, -10, -5, -5]), np.array([10, 10, 5, 5]))
        ib = (np.array([-5, -5]), np.array([5, 5]))
        np.random.seed(7)
        for r in range(n_runs):
            data = obstacles.generate_obstacle_scenarios(cfg, P.SIM_TIME, P.DT, P.NUM_SAMPLES)
            hs = env.compute_safe_halfspaces_for_runs([data["sample_trajectories"]], x_ref)[0]
            mpc = ref_mpc.MPCSafetyFilter(A, B, C, Q, R, P.HORIZON, P.DT)
            d_ref = env.compute_distance_to_collision(x_ref, data["realization_trajectories"])
            assert abs(res["min_distances"]["reference"][r] - d_ref.min()) <= 1e-12
            for metric in ("mean", "cvar", "dr_cvar"):
                x_f, _, info = mpc.filter_trajectory(x0, x_ref, u_ref, hs[metric], ib, sb[:2])
                assert info["status"] == "optimal"
                d = env.compute_distance_to_collision(x_f, data["realization_trajectories"])
                assert abs(res["min_distances"][metric][r] - d.min()) <= 1e-6, (r, metric)
        for m, v in res["min_distances"].items():
            assert res["collision_counts"][m] == int((v < 0).sum())
            assert abs(res["collision_probs"][m] - (v < 0).mean()) <= 1e-15
        table = mc.compare_risk_metrics(res)                 # evaluation/metrics.py of the reference
        assert set(table["dr_cvar"]) >= {"mean", "min", "collision_rate", "expected_shortfall"}
    finally:
        _purge()


def _fake_reference_modules():
    """Stand-ins for simulation/obstacles.py and simulation/planner.py (reference modules, absent on the GPU box)."""
    obstacles = types.ModuleType("simulation.obstacles")

    def generate_obstacle_scenarios(scenario_config, horizon, dt, n_samples=100):
        n_steps = int(horizon / dt)
        nominal, samples, real = [], [], []
        for ob in scenario_config["obstacles"]:
            t = np.arange(n_steps + 1)[:, None] * dt
            nom = np.asarray(ob["start"], float) + t * np.asarray(ob["velocity"], float)
            noise = np.random.multivariate_normal(np.zeros(2), 0.01 * np.eye(2), size=(n_samples, n_steps + 1))
            noise[:, 0, :] = 0.0
            nominal.append(nom)
            samples.append(nom[None, :, :] + noise)
            real.append(nom + np.random.laplace(0.0, 0.05, size=nom.shape))
        return {"nominal_trajectories": nominal, "sample_trajectories": samples, "realization_trajectories": real}

    obstacles.generate_obstacle_scenarios = generate_obstacle_scenarios
    planner = types.ModuleType("simulation.planner")

    class ReferenceTrajectoryPlanner:
        def __init__(self, A, B, C, Q, R, horizon, dt):
            self.horizon, self.dt = horizon, dt

        def straight_line_trajectory(self, start_pos, goal_pos, velocity=1.5):
            start, goal = np.asarray(start_pos, float), np.asarray(goal_pos, float)
            d = goal - start
            v = velocity * d / np.linalg.norm(d)
            x = np.zeros((self.horizon + 1, 4))
            for t in range(self.horizon + 1):
                x[t, :2] = start + min(t * self.dt * velocity, np.linalg.norm(d)) * d / np.linalg.norm(d)
                x[t, 2:] = v
            return x, np.zeros((self.horizon, 2)), {"status": "ok"}

    planner.ReferenceTrajectoryPlanner = ReferenceTrajectoryPlanner
    return obstacles, planner


@pytest.mark.gpu
def test_all_runs_in_one_launch_and_driver_end_to_end(tmp_path, monkeypatch):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _purge()
    monkeypatch.syspath_prepend(DROPIN)
    monkeypatch.chdir(tmp_path)
    try:
        env_mod = importlib.import_module("simulation.environment")
        obstacles, planner = _fake_reference_modules()
        monkeypatch.setitem(sys.modules, "simulation.obstacles", obstacles)
        monkeypatch.setitem(sys.modules, "simulation.planner", planner)
        mc = importlib.import_module("evaluation.monte_carlo")
        params = types.SimpleNamespace(HORIZON=20, DT=0.2, SIM_TIME=6.0, NUM_SAMPLES=64, Q_WEIGHT=1.0, R_WEIGHT=0.1)
        cfg = {"ego_start": [-4.0, 0.0], "ego_goal": [4.0, 0.0],
               "obstacles": [{"start": [4.0, 0.3], "velocity": [-1.0, 0.0]}, {"start": [0.0, 3.0], "velocity": [0.0, -0.8]}]}
        env = env_mod.SafetyFilteringEnvironment(0.3, 0.3, params.HORIZON, params.DT, 0.2, 0.1, 0.15)
        # all runs in one launch == run by run, bit for bit
        np.random.seed(3)
        runs = [obstacles.generate_obstacle_scenarios(cfg, params.SIM_TIME, params.DT, params.NUM_SAMPLES) for _ in range(6)]
        x_ref, _, _ = planner.ReferenceTrajectoryPlanner(env.A, env.B, env.C, None, None, params.HORIZON, params.DT) \
            .straight_line_trajectory(cfg["ego_start"], cfg["ego_goal"])
        import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
        l0 = pkg.launch_count()
        batched = env.compute_safe_halfspaces_for_runs([r["sample_trajectories"] for r in runs], x_ref)
        assert pkg.launch_count() - l0 == 1
        for r, run in enumerate(runs):
            single = env.compute_safe_halfspaces_for_trajectory(run["sample_trajectories"], x_ref)
            for metric in ("mean", "cvar", "dr_cvar"):
                for t in range(len(single[metric])):
                    for i in range(len(single[metric][t])):
                        h1, g1 = single[metric][t][i].get_constraint_params()
                        h2, g2 = batched[r][metric][t][i].get_constraint_params()
                        assert np.array_equal(h1, h2) and g1 == g2
        # the driver end to end; MPC QPs fanned out over two threads give the same numbers
        np.random.seed(11)
        a = mc.run_monte_carlo_simulation(env, cfg, 4, params)
        np.random.seed(11)
        b = mc.run_monte_carlo_simulation(env, cfg, 4, params, n_workers=2)
        for m in ("reference", "mean", "cvar", "dr_cvar"):
            assert np.array_equal(a["min_distances"][m], b["min_distances"][m]) and len(a["min_distances"][m]) == 4
        assert all(np.isfinite(v).all() for v in a["min_distances"].values())
    finally:
        _purge()
