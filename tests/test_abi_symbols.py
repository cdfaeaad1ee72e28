"""CPU tests: the C-ABI library loads without a GPU and exports every symbol include/drcvar.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200"


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, PKG, "libdrcvar.so")):
        g.build()
    from importlib import import_module
    return import_module(PKG + "._lib")


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "drcvar.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(drcvar_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared_symbols()
    assert len(names) >= 12
    cdll = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in include/drcvar.h but not exported"
        assert n in lib.SYMBOLS, f"{n} has no ctypes prototype in _lib.py"
    assert sorted(lib.SYMBOLS) == names


def test_host_logic_without_gpu(lib):
    L = lib.load()
    assert L.drcvar_version() == 1
    assert L.drcvar_reduction_lanes() == 512
    kf = ctypes.c_double()
    assert L.drcvar_tail_count(0.1, 10000, ctypes.byref(kf)) == 1000 and kf.value == 1000.0
    assert L.drcvar_tail_count(0.2, 23, ctypes.byref(kf)) == 5 and abs(kf.value - 4.6) < 1e-12
    assert L.drcvar_tail_count(0.2, 3, None) == 1
    assert L.drcvar_tail_count(1.0, 7, None) == 7
    assert L.drcvar_tail_count(0.0, 7, None) == lib.ERR_INVALID
    assert b"alpha" in L.drcvar_last_error()
    assert L.drcvar_tail_count(0.5, 0, None) == lib.ERR_INVALID


def test_cluster_sizes_are_pinned(lib):
    """BASELINE config 5 (N = 100 000 fp32) must run on clusters of 4 CTAs (33 clusters = 132 SMs on B200); a grown scratch
    area once pushed it silently onto clusters of 8 (15 clusters, 40 % slower).  232448 = B200's opt-in shared memory per SM."""
    L = lib.load()
    smem = 232448
    assert L.drcvar_cluster_ctas(100000, 4, smem) == 4
    assert L.drcvar_cluster_ctas(100000, 8, smem) == 8
    assert L.drcvar_cluster_ctas(40000, 4, smem) == 2
    assert L.drcvar_cluster_ctas(200000, 4, smem) == 8
    assert L.drcvar_cluster_ctas(32768, 4, smem) == 0          # single-chain contract: resident / streaming kernels
    assert L.drcvar_cluster_ctas(400000, 4, smem) == 0         # beyond 8 CTAs: streaming kernel
    assert L.drcvar_cluster_ctas(100000, 2, smem) == 0


def test_tail_count_matches_oracle(lib):
    from oracle import closed_form as cf
    import numpy as np
    L = lib.load()
    rng = np.random.RandomState(0)
    kf = ctypes.c_double()
    for _ in range(2000):
        n = int(rng.randint(1, 200000))
        alpha = float(rng.choice([0.05, 0.1, 0.2, 0.25, 0.3, 0.5, 1.0, rng.uniform(1e-4, 1.0)]))
        kc = L.drcvar_tail_count(alpha, n, ctypes.byref(kf))
        assert (kf.value, kc) == cf.tail_count(alpha, n)


def test_argument_validation_precedes_any_cuda_call(lib):
    L = lib.load()
    import numpy as np
    s = np.zeros((1, 4, 2))
    h = np.zeros((1, 2)); g = np.zeros((1, 3))
    rc = L.drcvar_halfspaces_f64(s.ctypes.data, 1, 4, 8, 2, 1, None, None, 2.0, 0.1, 0.1, 0.3, 0.3, 0,
                                 h.ctypes.data, None, g.ctypes.data, None, None, None, None, None, lib.HOST, None)
    assert rc == lib.ERR_INVALID
    rc = L.drcvar_halfspaces_f64(None, 1, 4, 8, 2, 1, None, None, 0.2, 0.1, 0.1, 0.3, 0.3, 0,
                                 h.ctypes.data, None, g.ctypes.data, None, None, None, None, None, lib.HOST, None)
    assert rc == lib.ERR_INVALID
    rc = L.drcvar_halfspaces_f32(s.ctypes.data, 1, 0, 8, 2, 1, None, None, 0.2, 0.1, 0.1, 0.3, 0.3, 0,
                                 h.ctypes.data, None, g.ctypes.data, None, None, None, None, None, lib.HOST, None)
    assert rc == lib.ERR_INVALID


def test_no_cpu_fallback_without_a_gpu():
    """The product path must fail loudly (never fall back to the oracle or any CPU code) when there is no CUDA device."""
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the compute path runs")
    import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib
    risk = dict(alpha=0.2, delta=0.1, epsilon=0.1, robot_radius=0.3, obstacle_radius=0.3)
    with pytest.raises(_lib.DrcvarError):
        pkg.compute_halfspaces(np.zeros((1, 8, 2)), None, **risk)
    with pytest.raises(_lib.DrcvarError):
        pkg.compute_halfspaces_generated(np.zeros((1, 2)), np.eye(2), 100, 1, **risk)
    # the shipped package never imports the test-only oracle
    import glob
    import os
    pkg_dir = os.path.dirname(pkg.__file__)
    for path in glob.glob(os.path.join(pkg_dir, "**", "*.py"), recursive=True):
        src = open(path).read()
        assert "import oracle" not in src and "from oracle" not in src, path


def test_flag_and_status_constants_match_the_header(lib):
    """_lib.py mirrors the DRCVAR_FLAG_* / status bits of include/drcvar.h by value (the ctypes stub of INTEGRATION.md)."""
    src = open(os.path.join(ROOT, "include", "drcvar.h")).read()
    flags = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+DRCVAR_FLAG_([A-Z_]+)\s+(\d+)u", src)}
    assert len(flags) >= 8 and "LARGE_COORDS" in flags
    for name, value in flags.items():
        assert getattr(lib, "FLAG_" + name) == value, name
    assert len(set(flags.values())) == len(flags) and all(v & (v - 1) == 0 for v in flags.values())   # distinct single bits
