"""ctypes binding of ``libpcb_b200.so`` (the C ABI of ``include/pcb_b200.h``).

There is exactly one evaluation backend: the hand-written sm_100a CUDA library.
If it has not been built, or no B200 is visible, evaluation raises -- there is
no CPU fallback on the product path.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpcb_b200.so")
CSRC = os.path.join(_HERE, "csrc")

PCB_OK = 0
PCB_EINVAL = -1
PCB_ECUDA = -2
PCB_ENOMEM = -3
PCB_EUNSUPPORTED = -4

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_vpp = C.POINTER(C.c_void_p)

#: every symbol ``include/pcb_b200.h`` declares: name -> (restype, argtypes)
SIGNATURES = {
    "pcb_version": (C.c_int, []),
    "pcb_last_error": (C.c_char_p, []),
    "pcb_device_count": (C.c_int, []),
    "pcb_device_info": (C.c_int, [C.c_int, _i32p, _i32p, _i32p]),
    "pcb_plan_destroy": (C.c_int, [C.c_void_p]),
    "pcb_tt_plan_create": (C.c_int, [C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, _i32p, _f64p, _vpp]),
    "pcb_tt_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pcb_tt_eval_fd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, _i32p, C.c_void_p,
                                 C.c_int, C.c_void_p]),
    "pcb_tt_plan_info": (C.c_int, [C.c_void_p, _i32p]),
    "pcb_tt_fd_algo": (C.c_int, [C.c_void_p, C.c_int, _i32p]),
    "pcb_tt_fd_path": (C.c_int, [C.c_void_p, C.c_int, _i32p, C.c_int]),
    "pcb_full_plan_create": (C.c_int, [C.c_int, C.c_int, _i32p, _f64p, _f64p, C.c_int,
                                       C.POINTER(_f64p), _vpp]),
    "pcb_full_plan_create_from_values": (C.c_int, [C.c_int, C.c_int, _i32p, _f64p, _f64p, _f64p, _f64p,
                                                   C.c_int, _i32p, _vpp]),
    "pcb_tensor_deriv": (C.c_int, [C.c_int, C.c_int, _i32p, C.c_int, _f64p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "pcb_tensor_contract": (C.c_int, [C.c_int, C.c_int, _i32p, C.c_int, _f64p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "pcb_tensor_extrude": (C.c_int, [C.c_int, C.c_int, _i32p, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "pcb_full_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]),
    "pcb_spline_plan_create": (C.c_int, [C.c_int, C.c_int, _i32p, _f64p, C.c_int, _i32p, _f64p, _f64p,
                                         C.c_int, C.POINTER(_f64p), _vpp]),
    "pcb_spline_lookup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pcb_spline_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "pcb_slider_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, _f64p, _f64p,
                                         C.c_double, C.c_int, _i32p, C.POINTER(_f64p), _vpp]),
    "pcb_slider_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pcb_plan_from_file": (C.c_int, [C.c_int, C.c_char_p, _vpp, _i32p, _i32p]),
    "pcb_plan_from_file_orders": (C.c_int, [C.c_int, C.c_char_p, C.c_int, _i32p, _vpp, _i32p, _i32p]),
    "pcb_file_write_approx": (C.c_int, [C.c_char_p, C.c_int, _f64p, _f64p, _i32p, _f64p]),
    "pcb_file_write_spline": (C.c_int, [C.c_char_p, C.c_int, _f64p, _f64p, _i32p, _i32p, _f64p,
                                        C.c_int, C.POINTER(_f64p)]),
    "pcb_file_rewrite": (C.c_int, [C.c_char_p, C.c_char_p]),
    "pcb_file_grid_arrays": (C.c_int, [C.c_double, C.c_double, C.c_int, _f64p, _f64p, _f64p]),
    "pcb_plan_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pcb_probe_fp64_peak": (C.c_int, [C.c_int, C.c_int, _f64p, _f64p]),
    "pcb_launch_count": (C.c_int64, []),
    "pcb_peer_alloc": (C.c_int, [C.c_int, C.c_uint64, _vpp, C.c_char_p]),
    "pcb_peer_free": (C.c_int, [C.c_int, C.c_void_p]),
    "pcb_peer_open": (C.c_int, [C.c_int, C.c_char_p, _vpp]),
    "pcb_peer_close": (C.c_int, [C.c_int, C.c_void_p]),
    "pcb_peer_push": (C.c_int, [C.c_int, C.c_int, _vpp, C.c_uint64, C.c_void_p, C.c_uint64,
                                C.c_void_p]),
    "pcb_peer_join": (C.c_int, [C.c_int, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


class BackendUnavailable(RuntimeError):
    """The CUDA library is not built / cannot be loaded / no B200 is visible."""


def build(verbose: bool = False) -> str:
    """Compile ``csrc/*.cu`` for sm_100a with nvcc into ``libpcb_b200.so`` (in-tree)."""
    cmd = ["make", "-C", CSRC, f"-j{max(4, min(8, os.cpu_count() or 4))}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc build failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stdout)
    return LIB_PATH


def load():
    """Load the library (once) and declare every signature.  Raises if it is absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise BackendUnavailable(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  pychebyshev_b200 has no CPU evaluation path."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.pcb_version() != 1:
            raise BackendUnavailable(f"ABI version mismatch: {lib.pcb_version()} != 1")
        _lib = lib
        return lib


def last_error() -> str:
    return load().pcb_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map C-ABI return codes to the reference's exception types."""
    if rc == PCB_OK:
        return
    msg = last_error()
    if rc == PCB_EINVAL:
        raise ValueError(msg)
    if rc == PCB_EUNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == PCB_ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def as_i32(seq):
    import numpy as np

    a = np.ascontiguousarray(np.asarray(seq, dtype=np.int32).ravel())
    return a, a.ctypes.data_as(_i32p)


def as_f64(seq):
    import numpy as np

    a = np.ascontiguousarray(np.asarray(seq, dtype=np.float64).ravel())
    return a, a.ctypes.data_as(_f64p)


def ptr_array(arrays):
    """``const double* const*`` from a list of C-contiguous float64 arrays (None -> NULL)."""
    arr = (_f64p * len(arrays))()
    for i, a in enumerate(arrays):
        arr[i] = a.ctypes.data_as(_f64p) if a is not None else None
    return arr
