"""The planner's default kernels must stay on the uniform datapath.

ptxas decides per kernel whether constant-bank operands are read with ``LDCU`` (uniform register,
feeds ``DFMA R, R, UR, R`` directly) or per lane with ``LDC``; the decision flips with innocuous
source changes (register budgets, where a row is stored, a guard).  ``tools/check_sass.py`` counts
both in the SASS of the built library; this test fails when a kernel the planner selects by default
falls off the uniform path.  CPU only (cuobjdump)."""

import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_planner_default_kernels_read_the_bank_on_the_uniform_datapath():
    import check_sass

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    if not os.path.exists(check_sass.LIB):
        pytest.skip("library not built")
    rows = check_sass.scan()
    assert len(rows) >= 20
    assert check_sass.regressions(rows) == []
    by = {r[0]: r for r in rows}
    # the tensor-core 2-D kernels really contain FP64 MMAs
    assert by["spline2d_dmma_kernel<0>"][5] >= 8 and by["slider2d_dmma_kernel<0>"][5] >= 8
    # every per-lane case is a known, documented one
    partial = {r[0] for r in rows if r[2] / max(1, r[1] + r[2]) > check_sass.MAX_LDC_SHARE}
    assert partial <= set(check_sass.KNOWN_PARTIAL) | {"ttc_fd_shared_kernel<2, 8, 256, 2>"}, partial
