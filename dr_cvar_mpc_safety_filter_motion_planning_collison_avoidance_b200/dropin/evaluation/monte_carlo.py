"""
Monte-Carlo evaluation driver — stands in for the reference's evaluation/monte_carlo.py, which is
DELETED from the reference tree: only evaluation/__pycache__/monte_carlo.cpython-310.pyc survives.  Its interface is
rebuilt from the names and strings in that file (run_monte_carlo_simulation(env, scenario_config, n_runs, params),
compare_risk_metrics, result keys min_distances / collision_counts / collision_probs / timing_stats, methods
reference / mean / cvar / dr_cvar, the timer labels and the printed summary) and from main.py:19-147, whose single-run
flow it repeats n_runs times.  PARITY IS UNPINNED: there is no reference source to compare against.

What changes against a run-by-run loop: the obstacle data of all runs is generated first (same order of random draws),
the safe halfspaces of ALL runs are computed in ONE launch (SafetyFilteringEnvironment.compute_safe_halfspaces_for_runs,
falling back to the per-run method of an environment that lacks it), and the MPC QPs — independent per (run, metric) —
can be fanned out over host threads (`n_workers`; the QP time is in LAPACK, which releases the GIL).  This is the real source of BASELINE config 4's scenario axis
(NUM_MC_RUNS = 300, config/parameters.py:33).
"""
import functools
from concurrent.futures import ThreadPoolExecutor

import numpy as np

try:   # the reference's own helpers when this overlay sits on a reference checkout
    from utils.timing import Timer, TimingStats
except ImportError:   # standalone: dropin/_stopwatch.py
    from _stopwatch import Timer, TimingStats

METHODS = ('reference', 'mean', 'cvar', 'dr_cvar')


def _filter_run(w, run):
    """One Monte-Carlo run: ONE filter object serves the three metrics in order, like main.py:19-147 (its
    last_optimal_u fallback then carries over from metric to metric exactly as there).  `w` is this call's work
    description (no module state: concurrent run_monte_carlo_simulation calls do not see each other)."""
    mpc = w['mpc_cls'](w['A'], w['B'], w['C'], w['Q'], w['R'], w['horizon'], w['dt'])
    out = []
    for metric in METHODS[1:]:
        x_f, u_f, info = mpc.filter_trajectory(w['x0'], w['x_ref'], w['u_ref'], w['halfspaces'][run][metric],
                                               w['input_bounds'], w['state_bounds'][:2])
        out.append((run, metric, x_f, info))
    return out


def run_monte_carlo_simulation(env, scenario_config, n_runs, params, n_workers=1):
    """
    Run Monte Carlo simulations to evaluate different risk metrics.

    env: Safety filtering environment; scenario_config: Scenario configuration (config/scenarios.py);
    n_runs: Number of Monte Carlo runs; params: Configuration parameters (config/parameters.py: HORIZON, DT, SIM_TIME,
    NUM_SAMPLES, Q_WEIGHT, R_WEIGHT); n_workers (extension): host threads for the MPC QPs.
    Returns: results: Dictionary of results — 'min_distances' {method: [n_runs]}, 'collision_counts', 'collision_probs'
    {method: value}, 'timing_stats' (TimingStats), for the methods reference (unfiltered), mean, cvar, dr_cvar.
    """
    from simulation.planner import ReferenceTrajectoryPlanner
    from simulation.obstacles import generate_obstacle_scenarios
    from core.mpc_filter import MPCSafetyFilter

    timing_stats = TimingStats()
    A, B, C = env.A, env.B, env.C
    n_states, n_inputs = A.shape[0], B.shape[1]
    Q = params.Q_WEIGHT * np.eye(n_states)
    R = params.R_WEIGHT * np.eye(n_inputs)
    state_bounds = (np.array([-10, -10, -5, -5]), np.array([10, 10, 5, 5]))      # main.py:54-56
    input_bounds = (np.array([-5, -5]), np.array([5, 5]))
    env.set_bounds(state_bounds, input_bounds)

    planner = ReferenceTrajectoryPlanner(A, B, C, Q, R, params.HORIZON, params.DT)
    x0 = np.zeros(n_states)
    x0[:2] = scenario_config['ego_start']
    with Timer("Reference Planning") as timer:
        x_ref, u_ref, plan_info = planner.straight_line_trajectory(scenario_config['ego_start'], scenario_config['ego_goal'])
    timing_stats.add("Reference Planning", timer.elapsed)
    if x_ref is None:
        print("Failed to plan reference trajectory:")
        print(plan_info)
        return None

    print(f"Running {n_runs} Monte Carlo simulations...")
    runs = []
    for _ in range(n_runs):
        with Timer() as timer:
            runs.append(generate_obstacle_scenarios(scenario_config, params.SIM_TIME, params.DT, params.NUM_SAMPLES))
        timing_stats.add("Obstacle Generation", timer.elapsed)

    with Timer("Computing Safe Halfspaces") as timer:
        samples = [r['sample_trajectories'] for r in runs]
        if hasattr(env, 'compute_safe_halfspaces_for_runs'):
            halfspaces = env.compute_safe_halfspaces_for_runs(samples, x_ref)          # all runs, one launch
        else:
            halfspaces = [env.compute_safe_halfspaces_for_trajectory(s, x_ref) for s in samples]
    for _ in range(n_runs):
        timing_stats.add("Computing Safe Halfspaces", timer.elapsed / max(n_runs, 1))

    work = dict(mpc_cls=MPCSafetyFilter, A=A, B=B, C=C, Q=Q, R=R, horizon=params.HORIZON, dt=params.DT, x0=x0, x_ref=x_ref,
                u_ref=u_ref, halfspaces=halfspaces, input_bounds=input_bounds, state_bounds=state_bounds)
    one_run = functools.partial(_filter_run, work)
    filtered = {}
    if n_workers > 1 and n_runs > 1:
        # threads, not processes: the QP time is in LAPACK (GIL released), and forking a process that holds a CUDA context
        # is not safe
        with ThreadPoolExecutor(n_workers) as pool:
            done = [item for per_run in pool.map(one_run, range(n_runs)) for item in per_run]
    else:
        done = [item for r in range(n_runs) for item in one_run(r)]
    for run, metric, x_f, info in done:
        filtered[(run, metric)] = x_f
        timing_stats.add(f"MPC Filtering ({metric})", info.get('solve_time', 0.0))

    min_distances = {m: [] for m in METHODS}
    for r in range(n_runs):
        real = runs[r]['realization_trajectories']
        for method in METHODS:
            traj = x_ref if method == 'reference' else filtered[(r, method)]
            distances = env.compute_distance_to_collision(traj, real)
            min_distances[method].append(float(np.min(distances)))
    results = {
        'min_distances': {m: np.array(v) for m, v in min_distances.items()},
        'collision_counts': {m: int(np.sum(np.array(v) < 0)) for m, v in min_distances.items()},
        'collision_probs': {m: float(np.mean(np.array(v) < 0)) if n_runs else 0.0 for m, v in min_distances.items()},
        'timing_stats': timing_stats,
    }
    print("\nMonte Carlo Simulation Results:")
    print(f"Total runs: {n_runs}")
    print("Collision Counts:")
    for m in METHODS:
        print(f"  {m}: {results['collision_counts'][m]} ({100.0 * results['collision_probs'][m]:.1f}%)")
    print("Minimum Distance Statistics:")
    for m in METHODS:
        d = results['min_distances'][m]
        if len(d):
            print(f"  {m}:")
            print(f"    Mean: {np.mean(d):.4f}")
            print(f"    Min:  {np.min(d):.4f}")
            print(f"    Max:  {np.max(d):.4f}")
            print(f"    Std:  {np.std(d):.4f}")
    return results


def compare_risk_metrics(results):
    """Text form of the reference's comparison plot: safety metrics (evaluation/metrics.py) of the minimum distances per
    method, as {method: dict}."""
    from evaluation.metrics import safety_metrics
    return {m: safety_metrics(np.asarray(results['min_distances'][m])) for m in METHODS}
