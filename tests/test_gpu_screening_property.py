"""
Property test of the fp32 screening bounds (csrc/halfspace_kernel.cuh and csrc/pipelined_kernel.cuh — the timed configuration: window placement with a rigorous direction bound
err_h, classification thresholds thr_above / thr_keep) — run on the B200 box: pytest -m gpu.

The screening only decides HOW the exact threshold is found: every sample it calls "surely above" must really lie above the
window, every sample it ignores below it.  A wrong bound would silently change the tail set.  So, over random scales,
offsets, anisotropies, tail fractions, sample counts, explicit (non-unit) normals and ego positions next to the mean, the
default path must return the SAME threshold T (bit for bit), the same tail-index set and the same direction as
DRCVAR_FLAG_GENERAL_ONLY (no window, no fp32 screening: exact radix select over the canonical fp64 losses of all samples),
for the tail-index instantiation and for the timed one.  45 examples x 64 halfspaces = 2 880 cases.
"""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

pytestmark = pytest.mark.gpu

B = 64


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
    return pkg


def make_case(seed, n, dist):
    rng = np.random.default_rng(seed)
    offset = 10.0 ** rng.uniform(-1.0, 6.0, size=(B, 1, 1)) * rng.choice([-1.0, 1.0], size=(B, 1, 2))
    offset[rng.random(B) < 0.15] = 0.0
    sigma = 10.0 ** rng.uniform(-6.0, 2.0, size=(B, 1, 1))
    aniso = 10.0 ** rng.uniform(-2.0, 0.0, size=(B, 1, 1))
    th = rng.uniform(0, 2 * np.pi, size=(B, 1))
    c, s_ = np.cos(th), np.sin(th)
    if dist == "gauss":
        z = rng.standard_normal((B, n, 2))
    elif dist == "uniform":
        z = rng.uniform(-1.7, 1.7, size=(B, n, 2))
    else:
        z = rng.standard_t(3, size=(B, n, 2))
    z[:, :, 1:2] *= aniso
    zx = c[:, :, None] * z[:, :, 0:1] - s_[:, :, None] * z[:, :, 1:2]
    zy = s_[:, :, None] * z[:, :, 0:1] + c[:, :, None] * z[:, :, 1:2]
    samples = (offset + sigma * np.concatenate([zx, zy], axis=2)).astype(np.float32)
    mean = samples.astype(np.float64).mean(axis=1)
    ego = mean + rng.uniform(-5, 5, size=(B, 2)) * np.maximum(1.0, np.abs(mean)) * 0.5
    near = rng.random(B)
    ego[near < 0.10] = mean[near < 0.10] + 1e-9 * rng.standard_normal((int((near < 0.10).sum()), 2))   # ego within 1e-9 of the mean
    ego[(near >= 0.10) & (near < 0.15)] = mean[(near >= 0.10) & (near < 0.15)]
    # ties and duplicates
    dup = rng.random(B) < 0.1
    samples[dup, n // 2:] = samples[dup, : n - n // 2]
    return samples, ego


@settings(max_examples=45, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 31 - 1),
       n=st.sampled_from([1024, 1500, 2048, 4096, 5000, 8192, 10000, 16384, 24000]),
       alpha=st.floats(0.012, 0.5), dist=st.sampled_from(["gauss", "gauss", "uniform", "student"]),
       explicit_h=st.booleans())
def test_screening_never_changes_the_tail(eng, seed, n, alpha, dist, explicit_h):
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    p = dict(alpha=alpha, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)
    s, ego = make_case(seed, n, dist)
    kw = {}
    if explicit_h:
        rng = np.random.default_rng(seed + 1)
        kw["h"] = rng.standard_normal((B, 2)) * 10.0 ** rng.uniform(-2, 2, size=(B, 1))   # non-unit normals
        ego = None
    exact = eng.compute_halfspaces(s, ego, want_tail=True, flags=_lib.FLAG_GENERAL_ONLY, **p, **kw)
    fast = eng.compute_halfspaces(s, ego, want_tail=True, **p, **kw)
    timed = eng.compute_halfspaces(s, ego, **p, **kw)                  # the instantiation bench.py times (no tail output)
    same = lambda x, y: np.array_equal(np.asarray(x).view(np.uint64), np.asarray(y).view(np.uint64))   # noqa: E731
    assert same(fast.h, exact.h) and same(timed.h, exact.h)
    assert same(fast.var, exact.var), np.flatnonzero(fast.var.view(np.uint64) != exact.var.view(np.uint64))
    assert same(timed.var, exact.var), np.flatnonzero(timed.var.view(np.uint64) != exact.var.view(np.uint64))
    assert np.array_equal(fast.tail_idx, exact.tail_idx)
    assert np.array_equal(fast.status & 1, exact.status & 1) and np.array_equal(timed.status & 1, exact.status & 1)
    fin = np.isfinite(exact.g).all(axis=1)
    scale = np.maximum(1.0, np.abs(exact.g[fin]))
    tol = 1e-6 * np.maximum(1.0, np.abs(s[fin]).max(axis=(1, 2)))[:, None] / scale      # fp32 partial sums: 1e-6 of the data scale
    assert np.all(np.abs(fast.g[fin] - exact.g[fin]) / scale <= tol)
    assert np.all(np.abs(timed.g[fin] - exact.g[fin]) / scale <= tol)
