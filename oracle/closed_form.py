"""
ORACLE — test infrastructure, NOT product code.

CPU (numpy, fp64) restatement of the reference's risk-bounded safe-halfspace path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product path never does.

PARITY STATUS: the reference ships no tests / golden vectors for this path and its solver
(cvxpy 1.2.1 -> ecos 2.0.14, environment.yml:31,33) is not installed in this image, so the
true ECOS output is *unpinned*.  This restatement is pinned instead against
  (1) the reference's own, unmodified core/risk_metrics.py + core/halfspaces.py +
      simulation/environment.py driven through oracle/cvxpy_shim (their LP built by their
      code, solved by HiGHS instead of ECOS)  -> tests/golden/*.json, tests/test_oracle_*.py
  (2) an independent HiGHS restatement of both LPs (oracle/lp_highs.py).

What is restated (reference file:line, relative to /root/reference):
  * h = unit(obstacle_mean - ego), fallback [1,0] if norm < 1e-10     core/geometry.py:35-53
  * mean halfspace (h from the ORIGIN, g~ = -(h.m - R|h|))            core/halfspaces.py:70-106
  * CVaR LP   -> closed form g = CVaR_a(-h.xi) + R|h| - delta         core/risk_metrics.py:179-265, 305-338
  * DR-CVaR LP-> closed form g* = CVaR_a(-h.xi) + R|h| + eps/a - delta,
                 g~ = g* - R|h|                                       core/risk_metrics.py:84-177, 267-303
  * per-obstacle / per-step drivers                                   core/halfspaces.py:196-247,
                                                                      simulation/environment.py:60-106
  * dead numpy helpers expected_value / var_metric / cvar_metric      core/risk_metrics.py:35-82

CANONICAL ARITHMETIC (the reference's BLAS matvec / np.mean are not bit-reproducible, so the
oracle DEFINES the arithmetic the CUDA path must reproduce bit-for-bit where order matters):
  * sample mean (per coordinate), v = the coordinate of the N samples:
      fp64 inputs: 512 slots, slot l accumulates samples i = l (mod 512) in increasing i, in fp64;
      fp32 inputs: d_i = fl32(v_i - v_0) (shift by the FIRST sample), 1024 lanes, lane l accumulates
                   d_i for i = l (mod 1024) in increasing i IN FP32; the 1024 lane sums are widened to
                   fp64 and adjacent lanes (2j, 2j+1) are added, giving 512 slots;
      then u[j] = s[j] + s[j+256] (j < 256); each group of 32 consecutive u is combined by an
      xor-butterfly (1,2,4,8,16); the 8 group totals by an adjacent-pair tree ((w0+w1)+(w2+w3))+((w4+w5)+(w6+w7));
      mean = S / N in fp64 (fp32 inputs: mean = fl64(v_0) + S / N).
      N > 32768 (the sizes served by the cluster / DSMEM kernel, where 4 or 8 CTAs each hold a part of the samples):
      the samples are cut into 8 OCTANTS of ceil(N * 2 * itemsize / 8 / 4096) * 4096 bytes (whole 4 KB rows; the last
      octants may be short or empty); every octant gets the slot sums + tree above on its own sub-array (lane index
      relative to the octant start, fp32 shift still by the GLOBAL first sample v_0), and the 8 octant totals are
      combined by the adjacent-pair tree ((o0+o1)+(o2+o3))+((o4+o5)+(o6+o7)) = S.
  * projection  p_i = rn(rn(h0*x_i) + rn(h1*y_i))  (no FMA);  loss  L_i = -p_i.
  * tail: k_f = alpha*N (snapped to the nearest integer when within 1e-9 relative),
    kc = ceil(k_f); T = kc-th largest loss; index set = {L_i > T} U lowest-index ties, |set| = kc
    (== np.argsort(-L, kind='stable')[:kc]); CVaR = (sum_{L_i>T} L_i + (k_f - #{L_i>T}) T) / k_f,
    which equals the LP optimum (1/(aN)) [sum_{j<=floor(aN)} L_(j) + (aN - floor(aN)) L_(floor(aN)+1)].
  * fp32 inputs: only the lane partial sums of the mean are fp32 (above); the projection, the tail
    selection and every offset use the samples promoted exactly to fp64.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

LANES = 512
OCTANT_MIN_N = 32768   # N above this uses the 8-octant rule
OCTANTS = 8
SENTINEL = 100.0  # core/risk_metrics.py:177,265,303,338

_XOR_IDX = {x: (np.arange(32) ^ x) for x in (1, 2, 4, 8, 16)}


def _tree512(s) -> float:
    """512 fp64 slots -> one fp64 total: pair (j, j+256), xor-butterfly in groups of 32, adjacent tree of 8."""
    u = s[:256] + s[256:]
    w = u.reshape(8, 32)
    for x in (1, 2, 4, 8, 16):
        w = w + w[:, _XOR_IDX[x]]
    w = w[:, 0]
    while w.shape[0] > 1:
        w = w[0::2] + w[1::2]
    return float(w[0])


def _lane_sums(v, lanes, acc_t):
    n = v.shape[0]
    rows = (n + lanes - 1) // lanes
    pad = rows * lanes - n
    if pad:
        v = np.concatenate([v, np.zeros(pad, dtype=v.dtype)])
    a = v.reshape(rows, lanes)
    s = np.zeros(lanes, dtype=acc_t)
    for r in range(rows):  # sequential per lane, increasing sample index, in the accumulation dtype
        s = s + a[r]
    return s


def canonical_sum(v) -> float:
    """Deterministic slot-structured fp64 sum (fp64 contract of the module docstring)."""
    v = np.ascontiguousarray(v, dtype=np.float64).ravel()
    if v.shape[0] == 0:
        return 0.0
    return _tree512(_lane_sums(v, LANES, np.float64))


def octant_len(n: int, itemsize: int) -> int:
    """Samples per octant for N > OCTANT_MIN_N: whole 4 KB rows of (x, y) pairs."""
    row_bytes = n * 2 * itemsize
    ob = -(-row_bytes // (OCTANTS * 4096)) * 4096
    return ob // (2 * itemsize)


def _pair_tree(w) -> float:
    w = np.asarray(w, dtype=np.float64)
    while w.shape[0] > 1:
        w = w[0::2] + w[1::2]
    return float(w[0])


def _slot_total(v, v0) -> float:
    """Canonical total of one (sub-)array: fp32 -> shifted fp32 lanes, fp64 -> fp64 slots; then the 512-tree."""
    if v.shape[0] == 0:
        return 0.0
    if v.dtype == np.float32:
        d = (v - v0).astype(np.float32)                        # fl32(v_i - v_0)
        s = _lane_sums(d, 2 * LANES, np.float32).astype(np.float64)
        s = s[0::2] + s[1::2]                                  # adjacent fp32 lanes, added in fp64
        return _tree512(s)
    return _tree512(_lane_sums(v, LANES, np.float64))


def canonical_mean_1d(v) -> float:
    """Canonical mean of one coordinate; the arithmetic depends on the INPUT dtype and on N (see module docstring)."""
    v = np.ascontiguousarray(v).ravel()
    if v.dtype != np.float32:
        v = v.astype(np.float64, copy=False)
    n = v.shape[0]
    if n > OCTANT_MIN_N:
        ol = octant_len(n, v.dtype.itemsize)
        total = _pair_tree([_slot_total(v[k * ol:(k + 1) * ol], v[0]) for k in range(OCTANTS)])
    else:
        total = _slot_total(v, v[0]) if n else 0.0
    if v.dtype == np.float32:
        return float(np.float64(v[0]) + np.float64(total) / np.float64(n))
    return float(np.float64(total) / np.float64(n))


def canonical_mean(samples) -> np.ndarray:
    """Mean obstacle position; stands in for np.mean(samples, axis=0) at core/halfspaces.py:85,130,174."""
    s = np.asarray(samples)
    if s.dtype != np.float32:
        s = s.astype(np.float64, copy=False)
    return np.array([canonical_mean_1d(s[:, 0]), canonical_mean_1d(s[:, 1])], dtype=np.float64)


def norm2(v0: float, v1: float) -> float:
    """Canonical 2-norm: sqrt(rn(rn(v0*v0) + rn(v1*v1)))."""
    v0 = float(v0)
    v1 = float(v1)
    return math.sqrt(v0 * v0 + v1 * v1)


def separating_vector(ego_pos, obstacle_pos) -> np.ndarray:
    """core/geometry.py:35-53 in canonical arithmetic."""
    d0 = float(obstacle_pos[0]) - float(ego_pos[0])
    d1 = float(obstacle_pos[1]) - float(ego_pos[1])
    nrm = norm2(d0, d1)
    if nrm < 1e-10:
        return np.array([1.0, 0.0])
    return np.array([d0 / nrm, d1 / nrm], dtype=np.float64)


def tail_count(alpha: float, n: int):
    """(k_f, kc): fractional tail mass alpha*N (snapped) and the tail-set size ceil(k_f) >= 1."""
    if not (0.0 < alpha <= 1.0):
        raise ValueError("alpha must be in (0, 1]")
    if n < 1:
        raise ValueError("N must be >= 1")
    k_f = float(alpha) * float(n)
    kr = float(np.rint(k_f))
    if abs(k_f - kr) <= 1e-9 * max(1.0, kr):
        k_f = kr
    if k_f > float(n):
        k_f = float(n)
    kc = int(math.ceil(k_f))
    kc = min(max(kc, 1), n)
    return k_f, kc


def projection(h, samples) -> np.ndarray:
    """p_i = rn(rn(h0*x_i) + rn(h1*y_i)); stands in for `h @ samples.T` (core/risk_metrics.py:145,233)."""
    s = np.asarray(samples)
    x = s[:, 0].astype(np.float64)
    y = s[:, 1].astype(np.float64)
    return (np.float64(h[0]) * x) + (np.float64(h[1]) * y)


def tail_select(losses: np.ndarray, alpha: float):
    """Exact tail of the losses: returns (cvar, T, tail_idx sorted ascending, k_f, kc)."""
    L = np.asarray(losses, dtype=np.float64) + 0.0  # -0.0 -> +0.0
    n = L.shape[0]
    k_f, kc = tail_count(alpha, n)
    T = float(np.partition(L, n - kc)[n - kc])  # kc-th largest
    gt = L > T
    c_gt = int(gt.sum())
    s_gt = math.fsum(L[gt].tolist())
    cvar = (s_gt + (k_f - c_gt) * T) / k_f
    eq_idx = np.nonzero(L == T)[0][: kc - c_gt]
    idx = np.sort(np.concatenate([np.nonzero(gt)[0], eq_idx])).astype(np.int32)
    return cvar, T, idx, k_f, kc


@dataclass
class OracleHalfspace:
    h: np.ndarray
    h_mean: np.ndarray
    mean: np.ndarray
    g_mean: float
    g_cvar: float
    g_dr_star: float
    g_dr: float          # g-tilde of the DR-CVaR halfspace
    cvar: float
    var: float           # threshold T (kc-th largest loss)
    tail_idx: np.ndarray = field(repr=False, default=None)
    nonfinite: bool = False


def halfspace(samples, ego_ref_pos, alpha, delta, epsilon, robot_radius, obstacle_radius, h_in=None) -> OracleHalfspace:
    """
    One (scenario, obstacle, step): everything MeanSafeHalfspace.create, CVaRSafeHalfspace.create and
    DRCVaRSafeHalfspace.create (core/halfspaces.py:70-194) produce, via the closed forms of the two LPs.
    `h_in` overrides the derived direction (the explicit-h entry points core/risk_metrics.py:267,305).
    """
    s = np.asarray(samples)
    if s.ndim != 2 or s.shape[1] != 2:
        raise ValueError("samples must be [N, 2]")
    m = canonical_mean(s)
    R = float(robot_radius) + float(obstacle_radius)

    # mean halfspace: h measured from the ORIGIN (core/halfspaces.py:88), g~ per :94
    hm = separating_vector((0.0, 0.0), m)
    hmn = norm2(hm[0], hm[1])
    g_mean = -(((hm[0] * m[0]) + (hm[1] * m[1])) - R * hmn)

    if h_in is None:
        h = separating_vector(ego_ref_pos, m)  # core/halfspaces.py:130,174
    else:
        h = np.array([float(h_in[0]), float(h_in[1])])
    hn = norm2(h[0], h[1])
    r = R * hn                                   # core/risk_metrics.py:234,293
    eoa = float(epsilon) / float(alpha)          # lambda* = 1/alpha (core/risk_metrics.py:122)

    p = projection(h, s)
    L = -p
    finite = bool(np.isfinite(m).all() and np.isfinite(h).all() and np.isfinite(L).all())
    if not finite:
        # solver failure sentinel (core/risk_metrics.py:173-177, 301-303, 336-338)
        return OracleHalfspace(h=h, h_mean=hm, mean=m, g_mean=float(g_mean), g_cvar=SENTINEL,
                               g_dr_star=SENTINEL, g_dr=SENTINEL - r, cvar=float("nan"), var=float("nan"),
                               tail_idx=np.zeros(0, np.int32), nonfinite=True)
    cvar, T, idx, _, _ = tail_select(L, alpha)
    g_cvar = (cvar + r) - float(delta)                    # CVaR LP optimum, used directly as g~ (core/halfspaces.py:139)
    g_star = ((cvar + r) + eoa) - float(delta)            # DR-CVaR LP optimum
    g_dr = g_star - r                                     # core/risk_metrics.py:299
    return OracleHalfspace(h=h, h_mean=hm, mean=m, g_mean=float(g_mean), g_cvar=float(g_cvar),
                           g_dr_star=float(g_star), g_dr=float(g_dr), cvar=float(cvar), var=float(T), tail_idx=idx)


def halfspaces_batch(samples, ego, alpha, delta, epsilon, robot_radius, obstacle_radius, h_in=None, want_tail=False):
    """Batch driver: samples [B,N,2], ego [B,2] -> dict of arrays (h[B,2], h_mean[B,2], g[B,3]=(mean,cvar,dr~), ...)."""
    s = np.asarray(samples)
    B = s.shape[0]
    ego = np.broadcast_to(np.asarray(ego, dtype=np.float64), (B, 2))
    out = {
        "h": np.empty((B, 2)), "h_mean": np.empty((B, 2)), "g": np.empty((B, 3)),
        "cvar": np.empty(B), "var": np.empty(B), "g_star": np.empty(B),
    }
    tails = []
    for b in range(B):
        o = halfspace(s[b], ego[b], alpha, delta, epsilon, robot_radius, obstacle_radius,
                      None if h_in is None else h_in[b])
        out["h"][b] = o.h
        out["h_mean"][b] = o.h_mean
        out["g"][b] = (o.g_mean, o.g_cvar, o.g_dr)
        out["cvar"][b] = o.cvar
        out["var"][b] = o.var
        out["g_star"][b] = o.g_dr_star
        if want_tail:
            tails.append(o.tail_idx)
    if want_tail:
        out["tail_idx"] = tails
    return out


def trajectory_halfspaces(obstacle_sample_trajectories, ego_ref_trajectory, horizon, alpha, delta, epsilon,
                          robot_radius, obstacle_radius):
    """
    simulation/environment.py:60-106 restated: for t < min(len(x_ref), HORIZON), obstacle i:
    samples = traj[i][:, t, :], ego = C @ x_ref[t] = x_ref[t][:2].  Returns g[t][i] triples and h's.
    """
    n_steps = min(len(ego_ref_trajectory), horizon)
    res = []
    for t in range(n_steps):
        row = []
        ego = np.asarray(ego_ref_trajectory[t])[:2]
        for traj in obstacle_sample_trajectories:
            row.append(halfspace(traj[:, t, :], ego, alpha, delta, epsilon, robot_radius, obstacle_radius))
        res.append(row)
    return res


# ---- dead helpers of the reference, same semantics (core/risk_metrics.py:35-82) -------------------------

def expected_value(samples):
    return np.mean(samples, axis=0)


def var_metric(samples, alpha):
    sorted_samples = np.sort(samples)
    index = int(np.ceil(len(samples) * (1 - alpha)))
    return sorted_samples[index - 1]


def cvar_metric(samples, alpha):
    var = var_metric(samples, alpha)
    tail = samples[samples >= var]
    if len(tail) == 0:
        return var
    return np.mean(tail)
