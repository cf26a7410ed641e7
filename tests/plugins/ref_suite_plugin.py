"""pytest plugin: run the REFERENCE's own test files with the B200 backend swapped in.

Loaded with ``-p ref_suite_plugin`` by tests/test_reference_suite.py (the reference's
tests all go through the evaluation methods, so they are a free regression suite for the drop-in).
It imports the unmodified reference (oracle/_ref/site), rebinds its vectorised evaluation methods
to the CUDA engine (``pychebyshev_b200.dropin.install``) and reports how many kernels ran.
"""

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# at plugin IMPORT time (before pytest imports the reference's conftest.py, which does
# `from pychebyshev import ...`): put the unmodified reference on sys.path and swap the backend in
from oracle import reference as _R  # noqa: E402

_ref = _R.load()
from pychebyshev_b200 import _engine as _eng, dropin as _dropin  # noqa: E402

_dropin.install(_ref)
_before = _eng.launch_count()


def pytest_terminal_summary(terminalreporter):
    from pychebyshev_b200 import _engine, dropin

    n = _engine.launch_count() - _before
    terminalreporter.write_line(
        f"B200_BACKEND installed={dropin.installed()} kernel_launches={n}")
