"""
Batched risk-bounded safe-halfspace engine: thin Python layer over the C ABI (libdrcvar.so).

`compute_halfspaces` is the batch entry the reference lacks; it replaces B calls of
MeanSafeHalfspace.create / CVaRSafeHalfspace.create / DRCVaRSafeHalfspace.create
(core/halfspaces.py:70-194 of the reference) with one kernel launch.  Inputs may be numpy arrays
(host path: staged by the library) or torch CUDA tensors / anything exposing __dlpack__ on a CUDA
device (device path: zero-copy, stream-ordered).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib


@dataclass
class HalfspaceBatch:
    """Results for B halfspaces.  Arrays are numpy (host path) or torch CUDA tensors (device path)."""
    h: object            # [B,2] normal of the CVaR / DR-CVaR halfspaces
    h_mean: object       # [B,2] normal of the mean halfspace (from the origin)
    g: object            # [B,3] (g_mean, g_cvar, g_drcvar) g-tilde per metric
    cvar: object         # [B]
    var: object          # [B]  threshold loss (kc-th largest)
    g_star: object       # [B]  DR-CVaR LP optimum before the radius is subtracted
    status: object       # [B]  int32 STATUS_* bits
    tail_idx: Optional[object] = None   # [B, kc] int32, ascending (parity mode)
    samples: Optional[object] = None    # [B, N, 2] float32: the generated samples (generate mode, want_samples=True)

    @property
    def g_mean(self):
        return self.g[:, 0]

    @property
    def g_cvar(self):
        return self.g[:, 1]

    @property
    def g_drcvar(self):
        return self.g[:, 2]


def tail_count(alpha: float, n: int):
    """(k_f, kc) exactly as the kernels use them."""
    lib = _lib.load()
    k_f = C.c_double()
    kc = _lib.check(lib.drcvar_tail_count(float(alpha), int(n), C.byref(k_f)))
    return k_f.value, int(kc)


def _is_torch_cuda(x):
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def _as_torch_cuda(x):
    """Accept torch CUDA tensors directly, other CUDA DLPack producers via torch.from_dlpack."""
    if _is_torch_cuda(x):
        return x
    if hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray):
        import torch
        t = torch.from_dlpack(x)
        if t.is_cuda:
            return t
    return None


def compute_halfspaces(samples, ego=None, *, alpha, delta, epsilon, robot_radius, obstacle_radius,
                       h=None, want_tail=False, flags=0, stream=None, out: Optional[HalfspaceBatch] = None):
    """
    samples: [B,N,2] float32/float64 (numpy -> host path; torch CUDA tensor -> device path), any strides
             whose elements are multiples of the dtype size.  [N,2] is treated as B = 1.
    ego:     [B,2] (or [2]) float64 ego reference positions; None = origin.
    h:       optional explicit normals [B,2] (the core/risk_metrics.py entry points take h explicitly).
    """
    lib = _lib.load()
    dev_t = _as_torch_cuda(samples)
    if dev_t is not None:
        return _compute_device(lib, dev_t, ego, alpha, delta, epsilon, robot_radius, obstacle_radius, h,
                               want_tail, flags, stream, out)
    s = np.asarray(samples)
    if s.dtype not in (np.float32, np.float64):
        s = s.astype(np.float64)
    if s.ndim == 2:
        s = s[None]
    if s.ndim != 3 or s.shape[2] != 2:
        raise ValueError("samples must be [B,N,2] or [N,2]")
    B, N, _ = s.shape
    it = s.dtype.itemsize
    if any(st % it for st in s.strides) or any(st < 0 for st in s.strides):
        s = np.ascontiguousarray(s)
    sb, sn, sc = (st // it for st in s.strides)
    if B <= 1:
        sb = N * 2 if sb == 0 else sb
    ego_a = None
    if ego is not None:
        ego_a = np.ascontiguousarray(np.broadcast_to(np.asarray(ego, dtype=np.float64), (B, 2)))
    h_a = None
    if h is not None:
        h_a = np.ascontiguousarray(np.broadcast_to(np.asarray(h, dtype=np.float64), (B, 2)))
    _, kc = tail_count(alpha, N)
    res = HalfspaceBatch(
        h=np.empty((B, 2)), h_mean=np.empty((B, 2)), g=np.empty((B, 3)), cvar=np.empty(B), var=np.empty(B),
        g_star=np.empty(B), status=np.zeros(B, dtype=np.int32),
        tail_idx=np.empty((B, kc), dtype=np.int32) if want_tail else None)
    fn = lib.drcvar_halfspaces_f32 if s.dtype == np.float32 else lib.drcvar_halfspaces_f64
    p = lambda a: None if a is None else a.ctypes.data  # noqa: E731
    rc = fn(s.ctypes.data, B, N, sb, sn, sc, p(ego_a), p(h_a), float(alpha), float(delta), float(epsilon),
            float(robot_radius), float(obstacle_radius), int(flags), p(res.h), p(res.h_mean), p(res.g), p(res.cvar),
            p(res.var), p(res.g_star), p(res.status), p(res.tail_idx), _lib.HOST, None)
    _lib.check(rc)
    return res


def _compute_device(lib, s, ego, alpha, delta, epsilon, rr, ro, h, want_tail, flags, stream, out):
    import torch
    if s.dim() == 2:
        s = s.unsqueeze(0)
    if s.dim() != 3 or s.shape[2] != 2:
        raise ValueError("samples must be [B,N,2] or [N,2]")
    if s.dtype not in (torch.float32, torch.float64):
        raise TypeError("device samples must be float32 or float64")
    B, N, _ = s.shape
    dev = s.device
    sb, sn, sc = s.stride()
    if min(sb, sn, sc) < 0:
        s = s.contiguous()
        sb, sn, sc = s.stride()
    f64 = dict(dtype=torch.float64, device=dev)

    def prep(x):
        if x is None:
            return None
        t = x if _is_torch_cuda(x) else torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)
        return t.to(**f64).expand(B, 2).contiguous()

    # Temporaries and outputs are made on the stream the kernel runs on: with an explicit `stream=` the conversions of ego / h
    # are ordered before the launch and the caching allocator cannot hand their memory to another stream while the kernel
    # still reads it.  The batch keeps them alive until the caller drops it.
    if stream is None:
        tstream = torch.cuda.current_stream(dev)
    elif hasattr(stream, "cuda_stream"):
        tstream = stream
    else:
        tstream = torch.cuda.ExternalStream(int(stream), device=dev)
    _, kc = tail_count(alpha, N)
    with torch.cuda.stream(tstream):
        ego_t, h_t = prep(ego), prep(h)
        if out is None:
            out = HalfspaceBatch(
                h=torch.empty((B, 2), **f64), h_mean=torch.empty((B, 2), **f64), g=torch.empty((B, 3), **f64),
                cvar=torch.empty(B, **f64), var=torch.empty(B, **f64), g_star=torch.empty(B, **f64),
                status=torch.zeros(B, dtype=torch.int32, device=dev),
                tail_idx=torch.empty((B, kc), dtype=torch.int32, device=dev) if want_tail else None)
    out._keepalive = (s, ego_t, h_t)
    p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    fn = lib.drcvar_halfspaces_f32 if s.dtype == torch.float32 else lib.drcvar_halfspaces_f64
    rc = fn(s.data_ptr(), B, N, sb, sn, sc, p(ego_t), p(h_t), float(alpha), float(delta), float(epsilon), float(rr),
            float(ro), int(flags), p(out.h), p(out.h_mean), p(out.g), p(out.cvar), p(out.var), p(out.g_star),
            p(out.status), p(out.tail_idx), dev.index if dev.index is not None else torch.cuda.current_device(),
            C.c_void_p(tstream.cuda_stream))
    _lib.check(rc)
    return out


def cholesky2(cov):
    """(l00, l10, l11) of 2x2 covariance matrices [..., 2, 2] (fp64): the `chol` argument of the generate mode."""
    cov = np.asarray(cov, dtype=np.float64)
    l00 = np.sqrt(cov[..., 0, 0])
    l10 = np.where(l00 > 0, cov[..., 1, 0] / np.where(l00 > 0, l00, 1.0), 0.0)
    l11 = np.sqrt(np.maximum(cov[..., 1, 1] - l10 * l10, 0.0))
    return np.stack([l00, l10, l11], axis=-1)


def compute_halfspaces_generated(mean, noise_cov, n_samples, seed, ego=None, *, alpha, delta, epsilon, robot_radius,
                                 obstacle_radius, h=None, chol=None, index_offset=0, want_tail=False,
                                 want_samples=False, flags=0, device=None, stream=None):
    """
    Fused Monte-Carlo sampling + halfspaces: halfspace b draws `n_samples` points  mean[b] + N(0, noise_cov[b])  inside
    the kernel (never stored) — generate_obstacle_sample_trajectories (simulation/obstacles.py:43-77) followed by the
    *SafeHalfspace.create calls (core/halfspaces.py:70-194), for B (obstacle, step) pairs in one launch.

    mean: [B,2] nominal positions; noise_cov: [2,2] or [B,2,2] (or None with chol = [B,3] / [3] Cholesky factors
    (l00, l10, l11) given directly); seed: 64-bit key of the Philox stream; index_offset: global index of halfspace 0
    (so that a shard reproduces its slice of the unsharded batch).
    device=None: host path (numpy outputs); device=int or torch.device: outputs are torch CUDA tensors on it.
    Returns a HalfspaceBatch; with want_samples=True also the generated float32 samples [B,N,2] as `.samples`.
    """
    lib = _lib.load()
    mean_a = np.ascontiguousarray(np.atleast_2d(np.asarray(mean, dtype=np.float64)))
    B = mean_a.shape[0]
    if mean_a.shape != (B, 2):
        raise ValueError("mean must be [B,2]")
    bcast = lambda x, w: np.array(np.broadcast_to(np.asarray(x, dtype=np.float64), (B, w)), order="C")  # noqa: E731
    chol_a = bcast(cholesky2(np.asarray(noise_cov, dtype=np.float64)) if chol is None else chol, 3)
    N = int(n_samples)
    ego_a = None if ego is None else bcast(ego, 2)
    h_a = None if h is None else bcast(h, 2)
    _, kc = tail_count(alpha, N)
    scal = (float(alpha), float(delta), float(epsilon), float(robot_radius), float(obstacle_radius), int(flags))
    if device is None:
        res = HalfspaceBatch(
            h=np.empty((B, 2)), h_mean=np.empty((B, 2)), g=np.empty((B, 3)), cvar=np.empty(B), var=np.empty(B),
            g_star=np.empty(B), status=np.zeros(B, dtype=np.int32),
            tail_idx=np.empty((B, kc), dtype=np.int32) if want_tail else None)
        samples = np.empty((B, N, 2), dtype=np.float32) if want_samples else None
        p = lambda a: None if a is None else a.ctypes.data  # noqa: E731
        rc = lib.drcvar_halfspaces_generated_f32(
            p(mean_a), p(chol_a), int(seed) & (2 ** 64 - 1), int(index_offset), B, N, p(ego_a), p(h_a), *scal,
            p(res.h), p(res.h_mean), p(res.g), p(res.cvar), p(res.var), p(res.g_star), p(res.status), p(res.tail_idx),
            p(samples), _lib.HOST, None)
        _lib.check(rc)
        res.samples = samples
        return res
    import torch
    dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    f64 = dict(dtype=torch.float64, device=dev)
    to = lambda a: None if a is None else torch.as_tensor(a, device=dev)  # noqa: E731
    mean_t, chol_t, ego_t, h_t = to(mean_a), to(chol_a), to(ego_a), to(h_a)
    res = HalfspaceBatch(
        h=torch.empty((B, 2), **f64), h_mean=torch.empty((B, 2), **f64), g=torch.empty((B, 3), **f64),
        cvar=torch.empty(B, **f64), var=torch.empty(B, **f64), g_star=torch.empty(B, **f64),
        status=torch.zeros(B, dtype=torch.int32, device=dev),
        tail_idx=torch.empty((B, kc), dtype=torch.int32, device=dev) if want_tail else None)
    samples = torch.empty((B, N, 2), dtype=torch.float32, device=dev) if want_samples else None
    if stream is None:
        stream = torch.cuda.current_stream(dev).cuda_stream
    elif hasattr(stream, "cuda_stream"):
        stream = stream.cuda_stream
    p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    rc = lib.drcvar_halfspaces_generated_f32(
        p(mean_t), p(chol_t), int(seed) & (2 ** 64 - 1), int(index_offset), B, N, p(ego_t), p(h_t), *scal,
        p(res.h), p(res.h_mean), p(res.g), p(res.cvar), p(res.var), p(res.g_star), p(res.status), p(res.tail_idx),
        p(samples), dev.index if dev.index is not None else torch.cuda.current_device(), C.c_void_p(stream))
    _lib.check(rc)
    res.samples = samples
    res._keepalive = (mean_t, chol_t, ego_t, h_t)   # inputs of the still-running launch
    return res


def launch_count() -> int:
    return int(_lib.load().drcvar_launch_count())


def max_samples(dtype, device: int = -1) -> int:
    return int(_lib.check(_lib.load().drcvar_max_samples(int(np.dtype(dtype).itemsize), int(device))))


def compute_trajectory(obstacle_sample_trajectories, ego_steps, *, alpha, delta, epsilon, robot_radius,
                       obstacle_radius, flags=0):
    """
    One launch for every (step, obstacle) of a reference trajectory — the batched form of
    SafetyFilteringEnvironment.compute_safe_halfspaces_for_trajectory (simulation/environment.py:60-106):
    halfspace (t, i) uses samples traj[i][:, t, :] and ego ego_steps[t].

    obstacle_sample_trajectories: list of n_obs float64 arrays [N, T1, 2] (C-contiguous, same N and T1)
    ego_steps: [n_steps, 2] float64, n_steps <= T1
    Returns (h [n_steps, n_obs, 2], h_mean [n_steps, n_obs, 2], g [n_steps, n_obs, 3], status [n_steps, n_obs]).
    """
    lib = _lib.load()
    trajs = [np.ascontiguousarray(t, dtype=np.float64) for t in obstacle_sample_trajectories]
    n_obs = len(trajs)
    if n_obs == 0:
        raise ValueError("at least one obstacle is required")
    N, T1, two = trajs[0].shape
    if two != 2 or any(t.shape != (N, T1, 2) for t in trajs):
        raise ValueError("every obstacle trajectory must be [N, T1, 2] with the same N and T1")
    ego = np.ascontiguousarray(ego_steps, dtype=np.float64)
    n_steps = ego.shape[0]
    ptrs = (C.c_void_p * n_obs)(*[t.ctypes.data for t in trajs])
    h = np.empty((n_steps, n_obs, 2))
    hm = np.empty((n_steps, n_obs, 2))
    g = np.empty((n_steps, n_obs, 3))
    st = np.zeros((n_steps, n_obs), dtype=np.int32)
    rc = lib.drcvar_trajectory_f64(ptrs, n_obs, N, T1, n_steps, ego.ctypes.data, float(alpha), float(delta),
                                   float(epsilon), float(robot_radius), float(obstacle_radius), int(flags),
                                   h.ctypes.data, hm.ctypes.data, g.ctypes.data, st.ctypes.data)
    _lib.check(rc)
    return h, hm, g, st
