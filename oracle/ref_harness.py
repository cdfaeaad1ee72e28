"""
ORACLE — test infrastructure, NOT product code.

Imports the UNMODIFIED reference modules from /root/reference (read-only) with
oracle/cvxpy_shim standing in for the missing cvxpy/ECOS, so the reference's own Python code
paths (core/risk_metrics.py, core/halfspaces.py, core/geometry.py, simulation/environment.py,
simulation/obstacles.py, simulation/planner.py, config/*) can be executed here and used to
generate golden vectors (tests/golden/make_golden.py) and differential tests.

/root/reference does not exist on the GPU box: nothing under `-m gpu`, smoke() or bench.py may
import this file.  Everything it produces is committed as fixtures under tests/golden/.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("DRCVAR_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cvxpy_shim")

_REF_TOPLEVEL = ("core", "utils", "config", "simulation", "evaluation")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "core"))


@contextlib.contextmanager
def reference_modules(quiet: bool = True):
    """
    Context manager yielding a namespace with the reference's modules imported fresh.
    cwd is switched to a scratch directory because the reference writes tmp/timing_info_*.json
    relative to cwd on every solve (core/risk_metrics.py:16-33).
    """
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items()
                  if k.split(".")[0] in _REF_TOPLEVEL + ("cvxpy", "matplotlib")}
    for k in list(saved_mods):
        del sys.modules[k]
    saved_cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="drcvar_ref_")
    sink = io.StringIO()
    try:
        sys.path[:0] = [_SHIM, REFERENCE_ROOT]
        # matplotlib is absent too; the hot path never touches it
        if importlib.util.find_spec("matplotlib") is None:
            mpl = types.ModuleType("matplotlib")
            mpl.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = mpl.pyplot
        os.chdir(scratch)
        ctx = contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()
        with ctx:
            ns = types.SimpleNamespace()
            ns.risk_metrics = importlib.import_module("core.risk_metrics")
            ns.halfspaces = importlib.import_module("core.halfspaces")
            ns.geometry = importlib.import_module("core.geometry")
            ns.environment = importlib.import_module("simulation.environment")
            ns.obstacles = importlib.import_module("simulation.obstacles")
            ns.planner = importlib.import_module("simulation.planner")
            ns.dynamics = importlib.import_module("core.dynamics")
            ns.mpc_filter = importlib.import_module("core.mpc_filter")   # QP solved by the shim's interior-point method
            ns.parameters = importlib.import_module("config.parameters")
            ns.scenarios = importlib.import_module("config.scenarios")
            ns.scratch = scratch
            yield ns
    finally:
        os.chdir(saved_cwd)
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k.split(".")[0] in _REF_TOPLEVEL + ("cvxpy", "matplotlib"):
                del sys.modules[k]
        sys.modules.update(saved_mods)
