"""
B200-native (sm_100a) implementation of the risk-bounded safe-halfspace hot path of the DR-CVaR MPC
safety filter (mean / CVaR / Wasserstein DR-CVaR offsets), behind the reference's Python interface.

  engine.compute_halfspaces   batched entry over libdrcvar.so (C ABI: include/drcvar.h)
  engine.compute_halfspaces_generated   the same with the Monte-Carlo samples drawn inside the kernel
  dropin/                     drop-in core/risk_metrics.py, core/halfspaces.py, core/geometry.py, ...
  sharding                    scenario sharding across the GPUs of one box (no collective on the hot path)
"""
from .engine import (HalfspaceBatch, cholesky2, compute_halfspaces, compute_halfspaces_generated,  # noqa: F401
                     compute_trajectory, launch_count, max_samples, tail_count)

__all__ = ["HalfspaceBatch", "cholesky2", "compute_halfspaces", "compute_halfspaces_generated", "compute_trajectory",
           "launch_count", "max_samples", "tail_count"]
