"""The reference's OWN test files, run unmodified against the swapped-in B200 backend.

``oracle/_ref/ref_tests`` holds byte-identical copies of ``/root/reference/tests`` (placed by
oracle/ref_install.py, git-ignored, shipped to the GPU box).  Each run is a subprocess
``pytest <reference test file> -p ref_suite_plugin``: the plugin imports the unmodified reference
and rebinds its evaluation methods to the CUDA engine, so every ``vectorized_eval*`` / ``eval*``
call those tests make is answered by the kernels.

Default: the hot-path files named in SURVEY.md §4 with the slowest host-side fixtures thinned
(see SLOW_K).  ``PCB_FULL_REF_SUITE=hot`` runs those files in full (357 tests);
``PCB_FULL_REF_SUITE=all`` runs every reference test file (algebra, calculus, extrude/slice, ... also
evaluate through the patched methods).
"""

import os
import re
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "ref_tests")

HOT_PATH_FILES = [
    "test_barycentric.py", "test_tensor_train.py", "test_spline.py", "test_slider.py",
    "test_special_points.py", "test_binary_format.py", "test_v0201_dim_threading.py",
]

#: reference tests that cannot hold under ANY replacement backend, with the reason
DESELECT: dict = {}

#: Default (driver) run only: the reference's `cheb_bs_5d` fixture is function-scoped and rebuilds
#: the 11^5 interpolant with 161,051 Python calls (~22 s of host time) for EACH test that uses it.
#: Four of the fourteen are kept (batch, multi, vectorised-vs-eval, delta); PCB_FULL_REF_SUITE=1
#: runs all of them (log: profiles/r2_reference_suite.md).
SLOW_K = {
    "test_barycentric.py": "not test_price and not test_gamma and not test_vega and not test_rho "
                           "and not test_fast_eval_matches_eval and not test_node_coincidence "
                           "and not test_error_estimate_bs_5d and not test_multi_matches_single",
    "test_binary_format.py": "not 5d_black_scholes",
}


def _mode():
    """'' (default: hot-path files, slow 5-D fixture tests thinned), 'hot' (hot-path files in full),
    anything else (every reference test file in full)."""
    return os.environ.get("PCB_FULL_REF_SUITE", "")


def _files():
    if _mode() not in ("", "hot") and os.path.isdir(REF_TESTS):
        return sorted(f for f in os.listdir(REF_TESTS) if f.startswith("test_") and f.endswith(".py"))
    return HOT_PATH_FILES


def _run(fname, extra=()):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "plugins"), ROOT,
                                         env.get("PYTHONPATH", "")])
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    env.setdefault("NUMBA_CACHE_DIR", "/tmp/pcb_numba_cache")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", "-p",
           "ref_suite_plugin", "--rootdir", REF_TESTS, os.path.join(REF_TESTS, fname)]
    for nodeid in DESELECT.get(fname, ()):
        cmd += ["--deselect", os.path.join(REF_TESTS, fname) + "::" + nodeid]
    if not _mode() and fname in SLOW_K:
        cmd += ["-k", SLOW_K[fname]]
    cmd += list(extra)
    return subprocess.run(cmd, cwd=REF_TESTS, env=env, capture_output=True, text=True,
                          timeout=1500)


@pytest.mark.gpu
@pytest.mark.parametrize("fname", _files())
def test_reference_test_file_passes_on_b200_backend(fname):
    if not os.path.exists(os.path.join(REF_TESTS, fname)):
        pytest.skip("reference tests not installed (run python oracle/ref_install.py)")
    res = _run(fname)
    tail = (res.stdout + res.stderr)[-3000:]
    log = os.environ.get("PCB_REF_SUITE_LOG")
    if log:
        with open(log, "a") as f:
            f.write(f"== {fname}: rc {res.returncode}\n" + "\n".join(res.stdout.splitlines()[-6:]) + "\n")
    assert res.returncode == 0, tail
    m = re.search(r"B200_BACKEND installed=True kernel_launches=(\d+)", res.stdout)
    assert m, tail
    if fname != "test_binary_format.py":  # the format tests never evaluate
        assert int(m.group(1)) > 0, "the reference tests ran without launching a kernel"


def test_reference_install_is_unmodified():
    """oracle/_ref (when present) is byte-identical to what oracle/ref_install.py recorded."""
    sys.path.insert(0, ROOT)
    from oracle import ref_install

    if not ref_install.installed():
        pytest.skip("reference not installed under oracle/_ref")
    assert ref_install.verify(quiet=True) == []
