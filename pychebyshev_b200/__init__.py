"""pychebyshev_b200 -- B200-native batch-evaluation engine for PyChebyshev interpolants.

Drop-in for the reference's vectorised evaluation path: the classes below keep the reference's
constructors, batch-evaluation entry points (with derivative orders) and the ``.pcb`` file
layout; evaluation runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/pcb_b200.h``.  No CPU fallback: evaluating without the built library or without a
B200 raises.
"""

from dataclasses import dataclass
from typing import Tuple

from ._lib import BackendUnavailable
from .approximation import ChebyshevApproximation
from .slider import ChebyshevSlider
from .spline import ChebyshevSpline
from .tt import ChebyshevTT

__version__ = "0.1.0"


def load_plan(path, device=None, orders=None):
    """Native ``.pcb`` -> device plan (``plan.eval(points)`` -> (N, G)); ``orders``: optional list of
    derivative-order rows (values only otherwise).  See ``pcb_plan_from_file[_orders]`` in
    ``include/pcb_b200.h``."""
    from ._engine import FilePlan

    return FilePlan(path, device, orders)


@dataclass(frozen=True)
class Domain:
    """Typed wrapper for per-dimension ``(lo, hi)`` bounds (reference ``__init__.py:35-45``)."""

    bounds: Tuple[Tuple[float, float], ...]

    def __post_init__(self):
        object.__setattr__(self, "bounds", tuple((float(a), float(b)) for a, b in self.bounds))


@dataclass(frozen=True)
class Ns:
    """Typed wrapper for per-dimension node counts (reference ``__init__.py:48-55``)."""

    counts: Tuple[int, ...]

    def __post_init__(self):
        object.__setattr__(self, "counts", tuple(self.counts))


@dataclass(frozen=True)
class SpecialPoints:
    """Typed wrapper for per-dimension kink locations (reference ``__init__.py:58-66``)."""

    knots_per_dim: Tuple[Tuple[float, ...], ...]

    def __post_init__(self):
        object.__setattr__(self, "knots_per_dim",
                           tuple(tuple(float(v) for v in k) for k in self.knots_per_dim))


__all__ = [
    "BackendUnavailable",
    "ChebyshevApproximation",
    "ChebyshevSlider",
    "ChebyshevSpline",
    "ChebyshevTT",
    "Domain",
    "Ns",
    "SpecialPoints",
    "load_plan",
    "__version__",
]
