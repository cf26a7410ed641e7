"""The C-ABI library loads and exports every symbol include/pcb_b200.h declares (no compute
calls: this runs without a GPU)."""

import ctypes
import os
import re

import pytest

from pychebyshev_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcb_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"PCB_API\s+[\w\s\*]+?\b(pcb_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    syms = declared_symbols()
    assert len(syms) >= 17
    assert sorted(_lib.SIGNATURES) == syms


def test_every_declared_symbol_is_exported(lib):
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert getattr(raw, name) is not None, name


def test_version_and_error_plumbing(lib):
    assert lib.pcb_version() == 1
    assert lib.pcb_launch_count() >= 0
    # argument validation happens before any CUDA call
    rc = lib.pcb_plan_destroy(None)
    assert rc == 0
    rc = lib.pcb_tt_eval(None, None, 0, None, None)
    assert rc == _lib.PCB_EINVAL
    assert b"not a TT plan" in lib.pcb_last_error()
    with pytest.raises(ValueError, match="not a TT plan"):
        _lib.check(rc)


def test_no_cpu_fallback_without_gpu():
    import numpy as np
    import torch

    import pychebyshev_b200 as pcb

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    cheb = pcb.ChebyshevApproximation.from_values(np.zeros((3, 3)), 2, [[0, 1], [0, 1]], [3, 3])
    with pytest.raises(pcb.BackendUnavailable):
        cheb.vectorized_eval_batch(np.zeros((4, 2)), [0, 0])
    tt = pcb.ChebyshevTT.from_values(np.ones((3, 3)), 2, [[0, 1], [0, 1]], [3, 3])
    with pytest.raises(pcb.BackendUnavailable):
        tt.eval_batch(np.zeros((4, 2)))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pychebyshev_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), fn
                assert "np_oracle" not in text and "liboracle" not in text, fn
