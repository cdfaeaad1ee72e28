"""
B200-native (sm_100a) implementation of the risk-bounded safe-halfspace hot path of the DR-CVaR MPC
safety filter (mean / CVaR / Wasserstein DR-CVaR offsets), behind the reference's Python interface.

  engine.compute_halfspaces   batched entry over libdrcvar.so (C ABI: include/drcvar.h)
  dropin/                     drop-in core/risk_metrics.py, core/halfspaces.py, core/geometry.py, ...
  sharding                    scenario sharding across the GPUs of one box (no collective on the hot path)
"""
from .engine import (HalfspaceBatch, compute_halfspaces, compute_trajectory, launch_count, max_samples,  # noqa: F401
                     tail_count)

__all__ = ["HalfspaceBatch", "compute_halfspaces", "compute_trajectory", "launch_count", "max_samples", "tail_count"]
