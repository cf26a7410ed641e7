"""The ``.pcb`` v1 interpolant file layout (byte-compatible with the reference).

Layout (reference ``_binary.py:28-34, 157-185, 208-236, 289-346`` and
``docs/user-guide/binary-format.md:44-93``), everything little-endian, no padding::

    header  12 B : b"PCB\\0" | u8 major=1 | u8 minor=0 | u16 class_tag | u32 reserved=0
    approx  (tag 1): u32 D | f64[D] lo | f64[D] hi | u32[D] n_nodes | f64[prod n] tensor (C-order)
    spline  (tag 2): u32 D | f64[D] lo | f64[D] hi | u32[D] n_nodes | u32[D] num_knots
                     | f64[sum k] knots | u32 P | P x f64[prod n] piece tensors (C-order)

This module only parses/serialises plain arrays; the interpolant classes rebuild their
nodes/weights/differentiation matrices from them, as the reference does on load.
"""

from __future__ import annotations

import os
import struct

import numpy as np

MAGIC = b"PCB\x00"
MAJOR, MINOR = 1, 0
TAG_APPROX, TAG_SPLINE = 1, 2
HEADER = struct.Struct("<4sBBHI")  # 12 bytes


class _Reader:
    def __init__(self, raw: bytes):
        self.raw = raw
        self.pos = 0

    def take(self, dtype: str, count: int, what: str) -> np.ndarray:
        nbytes = count * np.dtype(dtype).itemsize
        if self.pos + nbytes > len(self.raw):
            raise ValueError(
                f"unexpected EOF reading {what} (wanted {nbytes} bytes, got "
                f"{max(0, len(self.raw) - self.pos)})"
            )
        out = np.frombuffer(self.raw, dtype=dtype, count=count, offset=self.pos)
        self.pos += nbytes
        return out

    def u32(self, what: str) -> int:
        return int(self.take("<u4", 1, what)[0])


def is_pcb(path) -> bool:
    """True when the file starts with the ``.pcb`` magic (reference ``detect_format``)."""
    with open(os.fspath(path), "rb") as f:
        return f.read(4) == MAGIC


def peek_format_version(path) -> int:
    with open(os.fspath(path), "rb") as f:
        head = f.read(HEADER.size)
    if len(head) < HEADER.size:
        raise ValueError(f"file {path!r} is shorter than the {HEADER.size}-byte .pcb header")
    if head[:4] != MAGIC:
        raise ValueError(f"file {path!r} is not a .pcb file (magic mismatch: got {head[:4]!r})")
    return head[4]


def _parse_header(rd: _Reader) -> int:
    if len(rd.raw) < HEADER.size:
        raise ValueError(
            f"unexpected EOF reading header (wanted {HEADER.size} bytes, got {len(rd.raw)})")
    magic, major, _minor, tag, reserved = HEADER.unpack_from(rd.raw, 0)
    rd.pos = HEADER.size
    if magic != MAGIC:
        raise ValueError("not a PyChebyshev binary file (bad magic)")
    if major != MAJOR:
        raise ValueError(f"unsupported .pcb major version {major} (this build reads major {MAJOR})")
    if reserved != 0:
        raise ValueError("reserved header bytes nonzero — file may be corrupt")
    return tag


def _parse_grid(rd: _Reader):
    D = rd.u32("num_dimensions")
    if D < 1:
        raise ValueError(f"num_dimensions must be >= 1, got {D}")
    lo = rd.take("<f8", D, "domain_lo")
    hi = rd.take("<f8", D, "domain_hi")
    domain = [[float(a), float(b)] for a, b in zip(lo, hi)]
    for i, (a, b) in enumerate(domain):
        if a >= b:
            raise ValueError(f"domain[{i}]: lo ({a}) must be < hi ({b})")
    n_nodes = [int(v) for v in rd.take("<u4", D, "n_nodes")]
    for i, n in enumerate(n_nodes):
        if n < 1:
            raise ValueError(f"n_nodes[{i}] must be >= 1, got {n}")
    return D, domain, n_nodes


def parse(raw: bytes) -> dict:
    """Parse ``.pcb`` bytes into ``{'kind': 'approx'|'spline', ...arrays}``."""
    rd = _Reader(raw)
    tag = _parse_header(rd)
    if tag == TAG_APPROX:
        D, domain, n_nodes = _parse_grid(rd)
        total = int(np.prod(n_nodes, dtype=np.int64))
        tensor = rd.take("<f8", total, "tensor").astype(np.float64).reshape(n_nodes)
        return dict(kind="approx", num_dimensions=D, domain=domain, n_nodes=n_nodes, tensor=tensor)
    if tag == TAG_SPLINE:
        D, domain, n_nodes = _parse_grid(rd)
        num_knots = [int(v) for v in rd.take("<u4", D, "num_knots")]
        flat = rd.take("<f8", sum(num_knots), "knots")
        knots, off = [], 0
        for i, k in enumerate(num_knots):
            row = [float(v) for v in flat[off:off + k]]
            off += k
            if any(row[j] >= row[j + 1] for j in range(k - 1)):
                raise ValueError(f"knots in dim {i} not strictly ascending")
            knots.append(row)
        P = rd.u32("num_pieces")
        expect = int(np.prod([k + 1 for k in num_knots], dtype=np.int64))
        if P != expect:
            raise ValueError(f"num_pieces={P} does not match prod(num_knots+1)={expect}")
        per = int(np.prod(n_nodes, dtype=np.int64))
        pieces = [rd.take("<f8", per, "piece tensor").astype(np.float64).reshape(n_nodes)
                  for _ in range(P)]
        return dict(kind="spline", num_dimensions=D, domain=domain, n_nodes=n_nodes, knots=knots,
                    pieces=pieces)
    raise ValueError(f"unknown class_tag {tag}")


def read(path) -> dict:
    with open(os.fspath(path), "rb") as f:
        return parse(f.read())


def _f64_bytes(a) -> bytes:
    a = np.asarray(a)
    if a.dtype != np.float64:
        raise TypeError(f"binary format requires float64 arrays, got dtype={a.dtype}")
    return np.ascontiguousarray(a, dtype="<f8").tobytes()


def _grid_bytes(tag, domain, n_nodes) -> bytes:
    D = len(n_nodes)
    out = [HEADER.pack(MAGIC, MAJOR, MINOR, tag, 0), struct.pack("<I", D)]
    out.append(_f64_bytes(np.array([float(d[0]) for d in domain], dtype=np.float64)))
    out.append(_f64_bytes(np.array([float(d[1]) for d in domain], dtype=np.float64)))
    out.append(np.asarray(n_nodes, dtype="<u4").tobytes())
    return b"".join(out)


def approx_bytes(domain, n_nodes, tensor) -> bytes:
    return _grid_bytes(TAG_APPROX, domain, n_nodes) + _f64_bytes(
        np.ascontiguousarray(tensor, dtype=np.float64).ravel())


def spline_bytes(domain, n_nodes, knots, piece_tensors) -> bytes:
    out = [_grid_bytes(TAG_SPLINE, domain, n_nodes)]
    out.append(np.asarray([len(k) for k in knots], dtype="<u4").tobytes())
    flat = [float(v) for k in knots for v in k]
    if flat:
        out.append(_f64_bytes(np.array(flat, dtype=np.float64)))
    out.append(struct.pack("<I", len(piece_tensors)))
    for t in piece_tensors:
        out.append(_f64_bytes(np.ascontiguousarray(t, dtype=np.float64).ravel()))
    return b"".join(out)
