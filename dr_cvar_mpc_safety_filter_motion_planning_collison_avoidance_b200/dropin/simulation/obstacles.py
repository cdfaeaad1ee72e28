"""
Drop-in for the reference's simulation/obstacles.py (reference :1-196): same four functions, same signatures, same
returned structure — but the Monte-Carlo sample trajectories are not materialised on the host.

`generate_obstacle_sample_trajectories` (reference :43-77: nominal[t] + N(0, noise_cov) for t >= 1, exactly nominal[0] at
t = 0) returns a GeneratedSampleTrajectories: an array-like of shape (n_samples, n_steps+1, dim) that only records WHAT
the samples are drawn from (nominal trajectory, covariance, a 64-bit Philox key taken from numpy's global RNG, so
`np.random.seed(...)` still makes a run reproducible).  The drop-in SafetyFilteringEnvironment
(simulation/environment.py) recognises it and computes the halfspaces with the generate-mode kernel
(`drcvar_halfspaces_generated_f32`: the samples are drawn inside the kernel's staging step and never stored).  Anything
else that treats the object as an array (`np.asarray`, indexing, the reference's visualisation code) materialises it
ONCE through the same kernel's sample dump — the values are the ones the fused path classifies, bit for bit (fp32).

Parity with the reference's random STREAM cannot be had from a counter-based generator (numpy's legacy MT19937 polar method
consumes a data-dependent number of uniforms per normal): in the default (lazy) mode the
distribution is the reference's and the stream is the one `oracle/sample_gen.py` specifies.  The SEEDED REFERENCE-STREAM
mode (`DRCVAR_REFERENCE_STREAM=1` in the environment, or `obstacles.REFERENCE_STREAM = True`) is the parity mode: the sample
trajectories are drawn on the host with the reference's numpy calls in the reference's order and returned as the dense
float64 array the reference returns — bit-identical to `/root/reference/simulation/obstacles.py` under the same
`np.random.seed` (tests/test_dropin_api.py checks the seed-42 `head_on` / `multi_obstacle` arrays of the golden files) —
and take the stored-sample path through the kernel like any other numpy array.  The nominal trajectory and the Laplace realization follow the
reference's arithmetic exactly (same recurrence / same numpy calls in the same order); the realization still differs
from a reference run with the same seed because the sample generation no longer advances numpy's generator.
"""
import os

import numpy as np

from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import engine as _engine

# seeded reference-stream mode (module docstring): None = follow the DRCVAR_REFERENCE_STREAM environment variable
REFERENCE_STREAM = None


def _reference_stream():
    if REFERENCE_STREAM is not None:
        return bool(REFERENCE_STREAM)
    return os.environ.get("DRCVAR_REFERENCE_STREAM", "0") not in ("", "0")


class GeneratedSampleTrajectories:
    """Array-like [n_samples, n_steps+1, 2] of Monte-Carlo obstacle positions that exist only as (nominal, covariance, key)."""

    def __init__(self, nominal_trajectory, n_samples, noise_cov, key):
        self.nominal = np.array(nominal_trajectory, dtype=np.float64)
        self.noise_cov = np.array(noise_cov, dtype=np.float64)
        self.n_samples = int(n_samples)
        self.key = int(key)
        self.dtype = np.dtype(np.float64)
        self._dense = None

    # ---- what the fused path needs
    def kernel_inputs(self, n_steps):
        """(mean [n_steps,2], chol [n_steps,3]) of the first n_steps time steps; halfspace index = time step."""
        mean = np.ascontiguousarray(self.nominal[:n_steps, :2])
        chol = np.tile(_engine.cholesky2(self.noise_cov), (n_steps, 1))
        if n_steps > 0:
            chol[0] = 0.0            # t = 0: every sample is the initial position (reference :63)
        return mean, chol

    @property
    def materialised(self):
        return self._dense is not None

    # ---- ndarray look-alike
    @property
    def shape(self):
        return (self.n_samples, self.nominal.shape[0], self.nominal.shape[1])

    @property
    def ndim(self):
        return 3

    def __len__(self):
        return self.n_samples

    def _materialise(self):
        if self._dense is None:
            if self.nominal.shape[1] != 2:
                raise ValueError("the generate-mode kernel draws planar samples (dim = 2)")
            n_t = self.nominal.shape[0]
            mean, chol = self.kernel_inputs(n_t)
            res = _engine.compute_halfspaces_generated(mean, None, self.n_samples, self.key, chol=chol, alpha=0.5, delta=0.0,
                                                       epsilon=0.0, robot_radius=0.0, obstacle_radius=0.0, want_samples=True)
            dense = np.ascontiguousarray(np.transpose(res.samples.astype(np.float64), (1, 0, 2)))   # [N, T+1, 2]
            dense[:, 0, :] = self.nominal[0]   # the reference keeps the exact float64 start position
            self._dense = dense
        return self._dense

    def __array__(self, dtype=None, copy=None):
        a = self._materialise()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __getitem__(self, idx):
        return self._materialise()[idx]

    def __iter__(self):
        return iter(self._materialise())


def generate_nominal_trajectory(start_pos, direction, speed, n_steps, dt):
    """Positions [n_steps+1, dim] of an obstacle moving at `speed` along `direction` (reference :7-41: single integrator)."""
    start = np.asarray(start_pos, dtype=np.float64)
    length = np.linalg.norm(direction)
    if length < 1e-10:
        return np.tile(start_pos, (n_steps + 1, 1))          # stationary obstacle
    step = dt * (speed * (np.asarray(direction) / length))   # B u with B = dt I (core/dynamics.py:36-57)
    out = np.zeros((n_steps + 1, len(start)))
    out[0] = start
    for t in range(n_steps):
        out[t + 1] = out[t] + step                            # x+ = A x + B u with A = I (core/dynamics.py:59-88)
    return out


def generate_obstacle_sample_trajectories(nominal_trajectory, n_samples, noise_cov, dt):
    """Sample trajectories [n_samples, n_steps+1, dim] around the nominal one (reference :43-77), as a lazy array — or, in
    the seeded reference-stream mode, as the dense array of the reference's own draws."""
    if _reference_stream():
        # one multivariate_normal call per time step t >= 1, n_samples rows each, zero mean: the reference's consumption of
        # numpy's global generator (reference :65-75); step 0 is the nominal start for every sample
        steps, dim = nominal_trajectory.shape
        dense = np.empty((n_samples, steps, dim))
        dense[:, 0, :] = nominal_trajectory[0, :]
        zero = np.zeros(dim)
        for t in range(1, steps):
            dense[:, t, :] = nominal_trajectory[t, :] + np.random.multivariate_normal(mean=zero, cov=noise_cov, size=n_samples)
        return dense
    key = int(np.random.randint(0, 2 ** 31 - 1)) * (2 ** 31) + int(np.random.randint(0, 2 ** 31 - 1))
    return GeneratedSampleTrajectories(nominal_trajectory, n_samples, noise_cov, key)


def generate_laplace_realization(nominal_trajectory, noise_cov, dt):
    """One Laplace-distributed realization [n_steps+1, dim] of the obstacle's path (reference :79-113)."""
    nominal_trajectory = np.asarray(nominal_trajectory)
    dim = nominal_trajectory.shape[1]
    scale = np.sqrt(np.diag(noise_cov) / 2)                   # Laplace: var = 2 b^2
    out = np.zeros_like(nominal_trajectory)
    out[0, :] = nominal_trajectory[0, :]
    for t in range(1, nominal_trajectory.shape[0]):
        e_plus = np.random.exponential(scale=1.0, size=dim)   # difference of two exponentials, drawn in the reference's order
        e_minus = np.random.exponential(scale=1.0, size=dim)
        out[t, :] = nominal_trajectory[t, :] + scale * (e_plus - e_minus)
    return out


def generate_obstacle_scenarios(scenario_config, horizon, dt, n_samples=100):
    """{'nominal_trajectories', 'sample_trajectories', 'realization_trajectories'}: one entry per obstacle (reference :115-196)."""
    n_steps = int(horizon / dt)
    noise_cov = np.diag([0.01, 0.01])
    if 'obstacles' in scenario_config:
        specs = [(o['start'], o['direction'], o.get('speed', 1.0)) for o in scenario_config['obstacles']]
    else:
        specs = [(scenario_config['obstacle_start'], scenario_config['obstacle_direction'],
                  scenario_config.get('obstacle_speed', 1.0))]
    data = {'nominal_trajectories': [], 'sample_trajectories': [], 'realization_trajectories': []}
    for start, direction, speed in specs:
        nominal = generate_nominal_trajectory(start, direction, speed, n_steps, dt)
        data['nominal_trajectories'].append(nominal)
        data['sample_trajectories'].append(generate_obstacle_sample_trajectories(nominal, n_samples, noise_cov, dt))
        data['realization_trajectories'].append(generate_laplace_realization(nominal, noise_cov, dt))
    return data
