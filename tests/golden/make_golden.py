"""
Generates tests/golden/*.npz by running the reference's OWN, unmodified Python modules
(/root/reference, imported through oracle/ref_harness.py with oracle/cvxpy_shim standing in for the
absent cvxpy; the LPs the reference builds are solved by HiGHS instead of ECOS).

Run here (the GPU box has no /root/reference):   python tests/golden/make_golden.py
The produced fixtures are committed; tests only read them.

Cases
  head_on_seed42.npz / multi_obstacle_seed42.npz   main.py --mode single path (main.py:19-99, seed main.py:191)
  timing_sweep.npz      evaluation/timing_analysis.py:51-104 inputs (N = 10..1500), create() outputs
  explicit_h.npz        core/risk_metrics.py:267,305 entry points with non-unit h, fractional alpha*N, alpha*N<1
  n10k.npz              one N = 10 000, alpha = 0.1, eps = 0.01 halfspace (BASELINE config 4 parameters)
  mpc_head_on_seed42.npz / mpc_multi_obstacle_seed42.npz
                        the CONSUMER: MPCSafetyFilter.filter_trajectory (core/mpc_filter.py:40-178, called as in
                        main.py:101-113) fed with the reference's halfspaces of the same run; the QP is assembled by the
                        reference's own loops and solved by the shim's interior-point method instead of cvxpy/OSQP
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402

warnings.simplefilter("ignore")


def scenario_case(ref, name):
    P = ref.parameters
    np.random.seed(42)  # main.py:191
    cfg = ref.scenarios.get_scenario_config(name)
    data = ref.obstacles.generate_obstacle_scenarios(cfg, P.SIM_TIME, P.DT, P.NUM_SAMPLES)
    traj = data["sample_trajectories"]
    A, B, C = ref.dynamics.create_double_integrator_matrices(P.DT)
    Q = P.Q_WEIGHT * np.eye(4)
    R = P.R_WEIGHT * np.eye(2)
    planner = ref.planner.ReferenceTrajectoryPlanner(A, B, C, Q, R, P.HORIZON, P.DT)
    x_ref, u_ref, _ = planner.straight_line_trajectory(cfg["ego_start"], cfg["ego_goal"])
    env = ref.environment.SafetyFilteringEnvironment(P.ROBOT_RADIUS, P.OBSTACLE_RADIUS, P.HORIZON, P.DT,
                                                     P.ALPHA, P.DELTA, P.EPSILON)
    hs = env.compute_safe_halfspaces_for_trajectory(traj, x_ref)
    n_steps = len(hs["mean"])
    n_obs = len(traj)
    out = {}
    for metric in ("mean", "cvar", "dr_cvar"):
        h = np.zeros((n_steps, n_obs, 2))
        g = np.zeros((n_steps, n_obs))
        for t in range(n_steps):
            for i in range(n_obs):
                hh, gg = hs[metric][t][i].get_constraint_params()
                h[t, i] = hh
                g[t, i] = gg
        out[f"h_{metric}"] = h
        out[f"g_{metric}"] = g
    out["sample_trajectories"] = np.stack([tr[:, : P.HORIZON + 1, :] for tr in traj])  # [n_obs, N, H+1, 2]
    out["x_ref"] = x_ref
    out["params"] = np.array([P.ALPHA, P.DELTA, P.EPSILON, P.ROBOT_RADIUS, P.OBSTACLE_RADIUS, P.HORIZON], dtype=np.float64)
    return out


def mpc_case(ref, name):
    """main.py:37-113 — halfspaces of the reference's path into the reference's MPC filter, all three metrics."""
    P = ref.parameters
    np.random.seed(42)  # main.py:191
    cfg = ref.scenarios.get_scenario_config(name)
    data = ref.obstacles.generate_obstacle_scenarios(cfg, P.SIM_TIME, P.DT, P.NUM_SAMPLES)
    A, B, C = ref.dynamics.create_double_integrator_matrices(P.DT)
    Q = P.Q_WEIGHT * np.eye(4)
    R = P.R_WEIGHT * np.eye(2)
    planner = ref.planner.ReferenceTrajectoryPlanner(A, B, C, Q, R, P.HORIZON, P.DT)
    x_ref, u_ref, _ = planner.straight_line_trajectory(cfg["ego_start"], cfg["ego_goal"])
    env = ref.environment.SafetyFilteringEnvironment(P.ROBOT_RADIUS, P.OBSTACLE_RADIUS, P.HORIZON, P.DT,
                                                     P.ALPHA, P.DELTA, P.EPSILON)
    hs = env.compute_safe_halfspaces_for_trajectory(data["sample_trajectories"], x_ref)
    state_bounds = (np.array([-10, -10, -5, -5]), np.array([10, 10, 5, 5]))   # main.py:54-55
    input_bounds = (np.array([-5, -5]), np.array([5, 5]))
    x0 = np.zeros(4)
    x0[:2] = cfg["ego_start"]
    mpc = ref.mpc_filter.MPCSafetyFilter(A, B, C, Q, R, P.HORIZON, P.DT)
    out = dict(A=A, B=B, C=C, Q=Q, R=R, horizon=P.HORIZON, dt=P.DT, x0=x0, x_ref=x_ref, u_ref=u_ref,
               u_min=input_bounds[0], u_max=input_bounds[1], pos_min=state_bounds[0], pos_max=state_bounds[1])
    for metric in ("mean", "cvar", "dr_cvar"):
        x_f, u_f, info = mpc.filter_trajectory(x0, x_ref, u_ref, hs[metric], input_bounds, state_bounds[:2])
        assert info["status"] == "optimal" and not info.get("used_fallback"), (name, metric, info)
        out[f"x_{metric}"] = x_f
        out[f"u_{metric}"] = u_f
        out[f"objective_{metric}"] = info["objective"]
    return out


def timing_sweep_case(ref):
    P = ref.parameters
    H = ref.halfspaces
    np.random.seed(7)
    sizes = [10, 50, 100, 500, 1000, 1500]
    out = {"sizes": np.array(sizes)}
    for n in sizes:
        # evaluation/timing_analysis.py:63-70
        mean_pos = np.array([0.5, 0.0])
        scale = np.array([0.1, 0.1])
        samples = np.zeros((n, 2))
        for i in range(n):
            samples[i, 0] = np.random.normal(mean_pos[0], scale[0])
            samples[i, 1] = np.random.normal(mean_pos[1], scale[1])
        ego = np.array([0.0, 0.0])
        dr = H.DRCVaRSafeHalfspace.create(samples, ego, P.ALPHA, P.DELTA, P.EPSILON, P.ROBOT_RADIUS, P.OBSTACLE_RADIUS)
        cv = H.CVaRSafeHalfspace.create(samples, ego, P.ALPHA, P.DELTA, P.ROBOT_RADIUS, P.OBSTACLE_RADIUS)
        mn = H.MeanSafeHalfspace.create(samples, P.ROBOT_RADIUS, P.OBSTACLE_RADIUS)
        out[f"samples_{n}"] = samples
        out[f"out_{n}"] = np.array([dr.h[0], dr.h[1], dr.g_tilde, cv.h[0], cv.h[1], cv.g_tilde,
                                    mn.h[0], mn.h[1], mn.g_tilde])
    out["params"] = np.array([P.ALPHA, P.DELTA, P.EPSILON, P.ROBOT_RADIUS, P.OBSTACLE_RADIUS])
    return out


def explicit_h_case(ref):
    RM = ref.risk_metrics
    rng = np.random.RandomState(123)
    cases = [
        # (N, alpha, delta, eps, rr, ro, h)
        (23, 0.2, 0.1, 0.15, 0.3, 0.3, (0.6, 0.8)),          # alpha*N = 4.6 fractional
        (3, 0.2, 0.1, 0.15, 0.3, 0.3, (1.0, 0.0)),           # alpha*N = 0.6 < 1
        (40, 1.0, 0.05, 0.02, 0.2, 0.4, (0.0, -1.0)),        # alpha = 1 (plain mean of the loss)
        (37, 0.3, 0.2, 0.05, 0.25, 0.35, (1.5, -2.0)),       # NON-unit h, alpha*N = 11.1
        (64, 0.125, 0.1, 0.01, 0.3, 0.3, (-0.28, 0.96)),     # alpha*N = 8 exactly
        (1, 0.5, 0.1, 0.1, 0.3, 0.3, (1.0, 0.0)),            # single sample
        (50, 0.1, 0.1, 0.15, 0.3, 0.3, (0.70710678, 0.70710678)),
    ]
    out = {"n_cases": np.array(len(cases))}
    for c, (n, alpha, delta, eps, rr, ro, h) in enumerate(cases):
        samples = rng.normal(size=(n, 2)) * np.array([0.3, 0.2]) + np.array([2.0, -1.0])
        h = np.array(h, dtype=np.float64)
        # the reference caches optimizers on N only (core/risk_metrics.py:289,325): reset so alpha/delta/eps apply
        RM.drcvar_optimizer = None
        RM.cvar_optimizer = None
        g_star, g_tilde = RM.dr_cvar_halfspace(samples, h, alpha, delta, eps, rr, ro)
        g_cvar = RM.cvar_halfspace(samples, h, alpha, delta, rr, ro)
        out[f"samples_{c}"] = samples
        out[f"in_{c}"] = np.array([alpha, delta, eps, rr, ro, h[0], h[1]])
        out[f"out_{c}"] = np.array([g_star, g_tilde, g_cvar])
    return out


def n10k_case(ref):
    H = ref.halfspaces
    RM = ref.risk_metrics
    rng = np.random.RandomState(2024)
    n, alpha, delta, eps, rr, ro = 10000, 0.1, 0.1, 0.01, 0.3, 0.3
    mu = np.array([2.5, -1.75])
    samples = mu + 0.1 * rng.standard_normal((n, 2))
    samples32 = samples.astype(np.float32)
    ego = np.array([0.25, 0.5])
    RM.drcvar_optimizer = None
    RM.cvar_optimizer = None
    dr = H.DRCVaRSafeHalfspace.create(samples, ego, alpha, delta, eps, rr, ro)
    cv = H.CVaRSafeHalfspace.create(samples, ego, alpha, delta, rr, ro)
    mn = H.MeanSafeHalfspace.create(samples, rr, ro)
    RM.drcvar_optimizer = None
    RM.cvar_optimizer = None
    dr32 = H.DRCVaRSafeHalfspace.create(samples32.astype(np.float64), ego, alpha, delta, eps, rr, ro)
    return {
        "samples": samples, "samples32": samples32, "ego": ego,
        "params": np.array([alpha, delta, eps, rr, ro]),
        "out": np.array([dr.h[0], dr.h[1], dr.g_tilde, cv.h[0], cv.h[1], cv.g_tilde, mn.h[0], mn.h[1], mn.g_tilde]),
        "out32": np.array([dr32.h[0], dr32.h[1], dr32.g_tilde]),
    }


def main():
    with ref_harness.reference_modules(quiet=True) as ref:
        np.savez_compressed(os.path.join(HERE, "head_on_seed42.npz"), **scenario_case(ref, "head_on"))
        np.savez_compressed(os.path.join(HERE, "multi_obstacle_seed42.npz"), **scenario_case(ref, "multi_obstacle"))
        np.savez_compressed(os.path.join(HERE, "timing_sweep.npz"), **timing_sweep_case(ref))
        np.savez_compressed(os.path.join(HERE, "explicit_h.npz"), **explicit_h_case(ref))
        np.savez_compressed(os.path.join(HERE, "n10k.npz"), **n10k_case(ref))
        np.savez_compressed(os.path.join(HERE, "mpc_head_on_seed42.npz"), **mpc_case(ref, "head_on"))
        np.savez_compressed(os.path.join(HERE, "mpc_multi_obstacle_seed42.npz"), **mpc_case(ref, "multi_obstacle"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes", file=sys.stderr)


if __name__ == "__main__":
    main()
