"""Locate and import the UNMODIFIED reference package (``pychebyshev`` v0.21.1).

TEST INFRASTRUCTURE ONLY: tests/, ``__graft_entry__.smoke()`` and bench.py's CPU legs may use
this; the product package never does.  Search order:

1. ``oracle/_ref/site``   -- installed by oracle/ref_install.py, travels to the GPU box;
2. ``baseline/_ref``      -- a driver-provided ``pip install --target`` (when it exists);
3. ``/root/reference/src`` -- the read-only source tree (dev container only).

``load()`` returns the module (cached) or raises ``ReferenceUnavailable``.
"""

from __future__ import annotations

import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CANDIDATES = [
    os.path.join(HERE, "_ref", "site"),
    os.path.join(ROOT, "baseline", "_ref"),
    "/root/reference/src",
]
REF_TESTS = os.path.join(HERE, "_ref", "ref_tests")

_mod = None
_where = None


class ReferenceUnavailable(RuntimeError):
    pass


def locate():
    for c in CANDIDATES:
        if os.path.exists(os.path.join(c, "pychebyshev", "__init__.py")):
            return c
    return None


def load():
    """Import the reference; nothing is written next to its sources."""
    global _mod, _where
    if _mod is not None:
        return _mod
    where = locate()
    if where is None:
        raise ReferenceUnavailable(
            "reference package not found: run `python oracle/ref_install.py` where "
            "/root/reference exists (oracle/_ref/ then travels with the repo snapshot)")
    sys.dont_write_bytecode = True
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/pcb_numba_cache")
    if where not in sys.path:
        sys.path.insert(0, where)
    _mod = importlib.import_module("pychebyshev")
    _where = where
    return _mod


def where() -> str:
    load()
    return _where


def version() -> str:
    return getattr(load(), "__version__", "unknown")
