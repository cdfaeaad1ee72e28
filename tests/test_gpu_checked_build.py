"""
Stand-in for compute-sanitizer (closed on the GPU pool this repo is developed on): the CHECKED build of the library
(make -C csrc checked -> libdrcvar_checked.so; every shared-memory list / histogram / candidate-pool index and every staged byte
range asserted in range on the device) runs every kernel path once — resident, pipelined, streaming, the cluster kernels at 2 / 4 /
8 CTAs, generate mode, trajectory entry; bulk and strided loaders; parity mode; window misses and learned windows — and must
count zero failed assertions.  Results are compared between the paths inside the script (profiles/sanitize_case.py).
"""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, "dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200", "libdrcvar_checked.so")


def test_checked_build_counts_no_out_of_range_access():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(CHECKED):
        pytest.skip("libdrcvar_checked.so not built (make -C csrc checked)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "sanitize_case.py")], env=dict(os.environ, DRCVAR_LIB=CHECKED),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"device-side assertion failures (-?\d+) \(first site (-?\d+)", r.stdout)
    assert m, r.stdout[-500:]
    assert int(m.group(1)) == 0, f"{m.group(1)} failed device assertions, first at site {m.group(2)} (100000 * file id + line)"
