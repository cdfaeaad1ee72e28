"""
Drop-in for the reference's simulation/environment.py (SafetyFilteringEnvironment, reference :6-139) with the
(step, obstacle) double loop of compute_safe_halfspaces_for_trajectory replaced by ONE batched launch on the
strided [N, T+1, 2] sample arrays.  Constructor, attributes and return structure are unchanged:
{'mean'|'cvar'|'dr_cvar': [n_steps][n_obstacles] SafeHalfspace}.
"""
import time

import numpy as np

from core.halfspaces import MeanSafeHalfspace, CVaRSafeHalfspace, DRCVaRSafeHalfspace, compute_safe_halfspaces  # noqa: F401

from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import engine as _engine


def _double_integrator(dt, dim=2):
    """A, B, C of the planar double integrator (same matrices as the reference's core/dynamics.py:7-33)."""
    eye, zero = np.eye(dim), np.zeros((dim, dim))
    A = np.block([[eye, dt * eye], [zero, eye]])
    B = np.block([[0.5 * dt ** 2 * eye], [dt * eye]])
    C = np.block([eye, zero])
    return A, B, C


class SafetyFilteringEnvironment:
    def __init__(self, ROBOT_RADIUS, OBSTACLE_RADIUS, HORIZON, DT, ALPHA, DELTA, EPSILON):
        self.ROBOT_RADIUS = ROBOT_RADIUS
        self.OBSTACLE_RADIUS = OBSTACLE_RADIUS
        self.HORIZON = HORIZON
        self.DT = DT
        self.ALPHA = ALPHA
        self.DELTA = DELTA
        self.EPSILON = EPSILON
        self.A, self.B, self.C = _double_integrator(DT)
        self.n_states = self.A.shape[0]
        self.n_inputs = self.B.shape[1]
        self.n_outputs = self.C.shape[0]
        self.state_bounds = None
        self.input_bounds = None

    def set_bounds(self, state_bounds=None, input_bounds=None):
        self.state_bounds = state_bounds
        self.input_bounds = input_bounds

    def compute_safe_halfspaces_for_trajectory(self, obstacle_sample_trajectories, ego_ref_trajectory):
        """Safe halfspaces for each step t < min(len(ego_ref_trajectory), HORIZON) and each obstacle."""
        n_steps = min(len(ego_ref_trajectory), self.HORIZON)
        out = {'mean': [[] for _ in range(n_steps)], 'cvar': [[] for _ in range(n_steps)],
               'dr_cvar': [[] for _ in range(n_steps)]}
        if n_steps == 0 or len(obstacle_sample_trajectories) == 0:
            return out
        ego_steps = np.stack([self.C @ np.asarray(ego_ref_trajectory[t]) for t in range(n_steps)])
        # sample trajectories that exist only as (nominal trajectory, covariance, Philox key) — the drop-in
        # simulation/obstacles.py returns them: the samples are drawn inside the kernel, one launch per obstacle
        lazy = [tr for tr in obstacle_sample_trajectories if hasattr(tr, "kernel_inputs") and not tr.materialised]
        if len(lazy) == len(obstacle_sample_trajectories) and all(
                tr.shape[1] >= n_steps and tr.shape[2] == 2 for tr in lazy):
            t0 = time.time()
            cols = []
            for tr in lazy:
                mean, chol = tr.kernel_inputs(n_steps)
                cols.append(_engine.compute_halfspaces_generated(
                    mean, None, tr.n_samples, tr.key, ego=ego_steps, chol=chol, alpha=self.ALPHA, delta=self.DELTA,
                    epsilon=self.EPSILON, robot_radius=self.ROBOT_RADIUS, obstacle_radius=self.OBSTACLE_RADIUS))
            h = np.stack([c.h for c in cols], axis=1)
            hm = np.stack([c.h_mean for c in cols], axis=1)
            g = np.stack([c.g for c in cols], axis=1)
            return self._wrap(h, hm, g, n_steps, len(cols), time.time() - t0)
        trajs = [np.asarray(tr, dtype=np.float64) for tr in obstacle_sample_trajectories]
        uniform = all(tr.shape == trajs[0].shape for tr in trajs) and trajs[0].shape[1] >= n_steps
        t0 = time.time()
        if uniform:
            h, hm, g, _ = _engine.compute_trajectory(trajs, ego_steps, alpha=self.ALPHA, delta=self.DELTA,
                                                     epsilon=self.EPSILON, robot_radius=self.ROBOT_RADIUS,
                                                     obstacle_radius=self.OBSTACLE_RADIUS)
        else:  # ragged sample counts: one launch per step
            rows = [compute_safe_halfspaces([tr[:, t, :] for tr in trajs], ego_steps[t], self.ROBOT_RADIUS,
                                            self.OBSTACLE_RADIUS, self.ALPHA, self.DELTA, self.EPSILON)
                    for t in range(n_steps)]
            for t in range(n_steps):
                for k in out:
                    out[k][t] = rows[t][k]
            return out
        return self._wrap(h, hm, g, n_steps, len(trajs), time.time() - t0)

    def compute_safe_halfspaces_for_runs(self, runs_sample_trajectories, ego_ref_trajectory):
        """
        EXTENSION (not in the reference): compute_safe_halfspaces_for_trajectory for MANY Monte-Carlo runs
        that share one reference trajectory, in ONE launch — run r contributes its obstacles as extra columns of the
        (step, obstacle) grid.  `runs_sample_trajectories[r]` is what the per-run method takes (a list of [N, T+1, 2]
        arrays); returns one {'mean'|'cvar'|'dr_cvar': [t][obstacle]} dict per run, identical to calling the per-run
        method run by run.  This is where the scenario-batch axis of BASELINE config 4 comes from (NUM_MC_RUNS = 300,
        config/parameters.py:33).
        """
        if all(hasattr(tr, "kernel_inputs") and not tr.materialised for run in runs_sample_trajectories for tr in run):
            # lazy sample trajectories (drop-in simulation/obstacles.py): generate mode, one launch per obstacle, nothing stored
            return [self.compute_safe_halfspaces_for_trajectory(run, ego_ref_trajectory) for run in runs_sample_trajectories]
        runs = [[np.asarray(tr, dtype=np.float64) for tr in run] for run in runs_sample_trajectories]
        n_steps = min(len(ego_ref_trajectory), self.HORIZON)
        flat = [tr for run in runs for tr in run]
        uniform = len(flat) > 0 and all(tr.shape == flat[0].shape for tr in flat) and flat[0].shape[1] >= n_steps
        if n_steps == 0 or not uniform:
            return [self.compute_safe_halfspaces_for_trajectory(run, ego_ref_trajectory) for run in runs]
        ego_steps = np.stack([self.C @ np.asarray(ego_ref_trajectory[t]) for t in range(n_steps)])
        t0 = time.time()
        h, hm, g, _ = _engine.compute_trajectory(flat, ego_steps, alpha=self.ALPHA, delta=self.DELTA, epsilon=self.EPSILON,
                                                 robot_radius=self.ROBOT_RADIUS, obstacle_radius=self.OBSTACLE_RADIUS)
        elapsed = (time.time() - t0) / max(len(runs), 1)
        out, col = [], 0
        for run in runs:
            k = len(run)
            out.append(self._wrap(h[:, col:col + k], hm[:, col:col + k], g[:, col:col + k], n_steps, k, elapsed))
            col += k
        return out

    def _wrap(self, h, hm, g, n_steps, n_obs, elapsed):
        out = {'mean': [[] for _ in range(n_steps)], 'cvar': [[] for _ in range(n_steps)],
               'dr_cvar': [[] for _ in range(n_steps)]}
        info = {'setup_time': 0.0, 'solve_time': elapsed}
        zero = {'setup_time': 0, 'solve_time': 0, 'solve_call_time': 0}
        for t in range(n_steps):
            for i in range(n_obs):
                m = MeanSafeHalfspace(np.array(hm[t, i]), float(g[t, i, 0]))
                m.info = dict(zero)
                c = CVaRSafeHalfspace(np.array(h[t, i]), float(g[t, i, 1]))
                c.info = dict(info)
                d = DRCVaRSafeHalfspace(np.array(h[t, i]), float(g[t, i, 2]))
                d.info = dict(info)
                out['mean'][t].append(m)
                out['cvar'][t].append(c)
                out['dr_cvar'][t].append(d)
        return out

    def compute_safe_halfspaces_for_nominal(self, obstacle_nominal_trajectories, ego_ref_trajectory, noise_cov,
                                            n_samples, seed=0):
        """
        EXTENSION (not in the reference): the same result structure as
        compute_safe_halfspaces_for_trajectory, but from the NOMINAL obstacle trajectories [T+1, 2] and the noise
        covariance — the Monte-Carlo samples that generate_obstacle_sample_trajectories (simulation/obstacles.py:43-77)
        would have produced (nominal[t] + N(0, noise_cov); exactly nominal[0] at t = 0, obstacles.py:63) are drawn
        inside the kernel and never stored.  `seed` selects the Philox stream (oracle/sample_gen.py).
        """
        n_steps = min(len(ego_ref_trajectory), self.HORIZON)
        noms = [np.asarray(tr, dtype=np.float64) for tr in obstacle_nominal_trajectories]
        n_obs = len(noms)
        if n_steps == 0 or n_obs == 0:
            return {'mean': [[] for _ in range(n_steps)], 'cvar': [[] for _ in range(n_steps)],
                    'dr_cvar': [[] for _ in range(n_steps)]}
        ego_steps = np.stack([self.C @ np.asarray(ego_ref_trajectory[t]) for t in range(n_steps)])
        mean = np.stack([noms[i][t, :2] for t in range(n_steps) for i in range(n_obs)])
        ego = np.repeat(ego_steps, n_obs, axis=0)
        chol = np.tile(_engine.cholesky2(np.asarray(noise_cov, dtype=np.float64)), (n_steps * n_obs, 1))
        chol[:n_obs] = 0.0            # t = 0: every sample is the initial position
        t0 = time.time()
        res = _engine.compute_halfspaces_generated(mean, None, int(n_samples), int(seed), ego=ego, chol=chol,
                                                   alpha=self.ALPHA, delta=self.DELTA, epsilon=self.EPSILON,
                                                   robot_radius=self.ROBOT_RADIUS, obstacle_radius=self.OBSTACLE_RADIUS)
        shape = (n_steps, n_obs)
        return self._wrap(res.h.reshape(shape + (2,)), res.h_mean.reshape(shape + (2,)), res.g.reshape(shape + (3,)),
                          n_steps, n_obs, time.time() - t0)

    def compute_distance_to_collision(self, ego_trajectory, obstacle_trajectories):
        """Minimum clearance (centre distance minus both radii) over the obstacles, per time step."""
        n_steps = min(len(ego_trajectory), len(obstacle_trajectories[0]))
        distances = np.inf * np.ones(n_steps)
        for t in range(n_steps):
            ego = self.C @ np.asarray(ego_trajectory[t])
            for traj in obstacle_trajectories:
                d = np.linalg.norm(ego - traj[t]) - self.ROBOT_RADIUS - self.OBSTACLE_RADIUS
                distances[t] = min(distances[t], d)
        return distances
