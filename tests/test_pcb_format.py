"""The .pcb v1 layout: byte-identical with files the reference wrote (its own fixtures under
tests/fixtures/ and fresh reference-written bytes, all stored in tests/golden/pcb_files.npz),
field offsets of the spec, and the reference's error behaviour on corrupt input
(reference tests/test_binary_format.py:562-770)."""

import hashlib
import struct

import numpy as np
import pytest

import _golden as G
import pychebyshev_b200 as pcb
from pychebyshev_b200 import pcbfile


@pytest.fixture(scope="module")
def gold():
    return G.load("pcb_files")


def test_parse_reference_fixtures(gold, tmp_path):
    raw = gold["approx_2d_simple_bytes"].tobytes()
    assert hashlib.sha256(raw).hexdigest() == str(gold["approx_2d_simple_sha256"])
    rec = pcbfile.parse(raw)
    assert rec["kind"] == "approx" and rec["num_dimensions"] == 2
    assert np.array_equal(rec["tensor"], gold["approx_2d_simple_tensor"])
    # bytes -> object -> bytes identity
    path = tmp_path / "a.pcb"
    path.write_bytes(raw)
    obj = pcb.ChebyshevApproximation.load(path)
    out = tmp_path / "b.pcb"
    obj.save(out)
    assert out.read_bytes() == raw
    assert pcbfile.peek_format_version(out) == 1

    raw = gold["spline_1d_kink_bytes"].tobytes()
    assert hashlib.sha256(raw).hexdigest() == str(gold["spline_1d_kink_sha256"])
    path = tmp_path / "s.pcb"
    path.write_bytes(raw)
    sp = pcb.ChebyshevSpline.load(path)
    assert isinstance(sp, pcb.ChebyshevSpline) and sp.num_pieces == 2
    out = tmp_path / "s2.pcb"
    sp.save(out)
    assert out.read_bytes() == raw


def test_writer_matches_fresh_reference_bytes(gold):
    tensor = gold["approx_3x4_tensor"]
    raw = pcbfile.approx_bytes([[0.0, 1.0], [-2.0, 2.0]], [3, 4], tensor)
    assert raw == gold["approx_3x4_bytes"].tobytes()


@pytest.mark.parametrize("name", ["spline_bs2d", "spline_bs3d", "spline_multiknot3d", "spline_abs1d"])
def test_spline_writer_matches_reference_bytes(name):
    from oracle import np_oracle as O

    g = G.load(name)
    knots, shape, pieces = G.spline_parts(g, O.diff_matrix)
    n = [int(v) for v in g["piece_n_nodes"][0]]
    dom = [list(map(float, r)) for r in g["domain"]]
    raw = pcbfile.spline_bytes(dom, n, knots, [p[0] for p in pieces])
    assert raw == g["pcb_bytes"].tobytes()
    rec = pcbfile.parse(raw)
    assert rec["knots"] == knots and len(rec["pieces"]) == len(pieces)


def test_field_offsets_of_the_spec():
    raw = pcbfile.approx_bytes([[0.0, 1.0]], [2], np.array([1.5, -2.5]))
    # 12-byte header + u32 D + f64 lo + f64 hi + u32 n + 2 f64
    assert len(raw) == 12 + 4 + 8 + 8 + 4 + 16
    assert raw[:4] == b"PCB\x00" and raw[4] == 1 and raw[5] == 0
    assert struct.unpack_from("<H", raw, 6)[0] == 1 and raw[8:12] == b"\x00" * 4
    assert struct.unpack_from("<I", raw, 12)[0] == 1
    assert struct.unpack_from("<dd", raw, 16) == (0.0, 1.0)
    assert struct.unpack_from("<I", raw, 32)[0] == 2
    assert struct.unpack_from("<dd", raw, 36) == (1.5, -2.5)


def test_corruption_errors():
    good = pcbfile.approx_bytes([[0.0, 1.0]], [2], np.array([1.0, 2.0]))
    with pytest.raises(ValueError, match="bad magic"):
        pcbfile.parse(b"XXXX" + good[4:])
    with pytest.raises(ValueError, match="unsupported .pcb major version"):
        pcbfile.parse(good[:4] + b"\x02" + good[5:])
    with pytest.raises(ValueError, match="reserved"):
        pcbfile.parse(good[:8] + b"\x01\x00\x00\x00" + good[12:])
    with pytest.raises(ValueError, match="unexpected EOF"):
        pcbfile.parse(good[:-3])
    with pytest.raises(ValueError, match="unexpected EOF reading header"):
        pcbfile.parse(good[:7])
    with pytest.raises(ValueError, match="unknown class_tag"):
        pcbfile.parse(good[:6] + struct.pack("<H", 9) + good[8:])
    bad_dom = pcbfile.approx_bytes([[1.0, 1.0]], [2], np.array([1.0, 2.0]))
    with pytest.raises(ValueError, match="must be <"):
        pcbfile.parse(bad_dom)
    with pytest.raises(TypeError, match="float64"):
        pcbfile._f64_bytes(np.array([1.0, 2.0], dtype=np.float32))


def test_class_tag_mismatch_and_save_guards(tmp_path):
    cheb = pcb.ChebyshevApproximation.from_values(np.ones((3, 3)), 2, [[0, 1], [0, 1]], [3, 3])
    p = tmp_path / "a.pcb"
    cheb.save(p)
    with pytest.raises(ValueError, match="class_tag"):
        pcb.ChebyshevSpline.load(p)
    cheb.additional_data = {"k": 1}
    with pytest.raises(NotImplementedError, match="additional_data"):
        cheb.save(tmp_path / "b.pcb")
    unbuilt = pcb.ChebyshevApproximation(lambda x, _: 0.0, 1, [[0, 1]], [4])
    with pytest.raises(RuntimeError, match="unbuilt"):
        unbuilt.save(tmp_path / "c.pcb")
    # pickle round trip keeps the tensor and drops the function
    cheb.additional_data = None
    q = tmp_path / "a.pkl"
    cheb.save(q, format="pickle")
    back = pcb.ChebyshevApproximation.load(q)
    assert np.array_equal(back.tensor_values, cheb.tensor_values) and back.function is None


def test_native_loader_rejects_corrupt_files_before_touching_cuda(tmp_path):
    """pcb_plan_from_file validates like _binary.py (no GPU needed: parsing comes first)."""
    import ctypes as C

    from pychebyshev_b200 import _lib

    lib = _lib.load()
    good = pcbfile.approx_bytes([[0.0, 1.0]], [2], np.array([1.0, 2.0]))
    cases = {
        "bad magic": b"XXXX" + good[4:],
        "unsupported .pcb major version": good[:4] + b"\x02" + good[5:],
        "reserved": good[:8] + b"\x01\x00\x00\x00" + good[12:],
        "unexpected EOF": good[:-3],
        "unknown class_tag": good[:6] + struct.pack("<H", 9) + good[8:],
        "must be <": pcbfile.approx_bytes([[1.0, 1.0]], [2], np.array([1.0, 2.0])),
    }
    for needle, raw in cases.items():
        path = tmp_path / "bad.pcb"
        path.write_bytes(raw)
        plan, kind, ndim = C.c_void_p(), C.c_int32(), C.c_int32()
        rc = lib.pcb_plan_from_file(0, str(path).encode(), C.byref(plan), C.byref(kind), C.byref(ndim))
        assert rc == _lib.PCB_EINVAL, needle
        assert needle in _lib.last_error(), (needle, _lib.last_error())
    rc = lib.pcb_plan_from_file(0, b"/nonexistent/file.pcb", C.byref(plan), None, None)
    assert rc == _lib.PCB_EINVAL and "cannot open" in _lib.last_error()


def test_files_written_here_are_readable_by_the_reference_c_reader(tmp_path):
    """oracle/_ref/pcb_reader is the reference's examples/binary_reader/reader.c compiled as is
    (oracle/ref_reader.mk).  It must parse our files and reproduce the interpolant's values."""
    import os
    import subprocess

    from oracle import np_oracle as O

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "pcb_reader")
    if not os.path.exists(exe):
        if not os.path.exists("/root/reference/examples/binary_reader/reader.c"):
            pytest.skip("reference C reader not built and reference tree absent")
        subprocess.run(["make", "-C", os.path.join(root, "oracle"), "-f", "ref_reader.mk"], check=True,
                       capture_output=True)
    g = G.load("full_3d")
    n, nodes, weights, dms = G.full_parts(g)
    cheb = pcb.ChebyshevApproximation.from_values(g["tensor"], 3, g["domain"].tolist(), n)
    path = tmp_path / "full3d.pcb"
    cheb.save(path)
    pts = g["points"][:12]
    ref = O.full_eval_batch(g["tensor"], nodes, weights, dms, pts, [0, 0, 0])
    for p, want in zip(pts, ref):
        out = subprocess.run([exe, str(path)] + [repr(float(v)) for v in p], capture_output=True,
                             text=True, check=True).stdout
        got = float(out.strip().split()[-1])
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (got, want)


def _build_c_host(tmp_path):
    """gcc examples/pcb_eval.c against include/pcb_b200.h + libpcb_b200.so (+ libcudart)."""
    import os
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "pychebyshev_b200")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA headers are not available")
    exe = str(tmp_path / "pcb_eval")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(root, "include"), "-I", os.path.join(cuda, "include"),
                    os.path.join(root, "examples", "pcb_eval.c"), "-L", libdir, "-lpcb_b200",
                    "-L", os.path.join(cuda, "lib64"), "-lcudart", f"-Wl,-rpath,{libdir}",
                    f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}", "-o", exe],
                   check=True, capture_output=True)
    return exe


def test_c_host_compiles_against_the_header_and_reports_reference_errors(tmp_path):
    """A plain C program is a client of the ABI: it builds from include/pcb_b200.h alone and gets
    the reference's error text for a corrupt file (parsing precedes any CUDA call)."""
    import subprocess

    from pychebyshev_b200 import _lib

    _lib.load()
    exe = _build_c_host(tmp_path)
    good = pcbfile.approx_bytes([[0.0, 1.0]], [2], np.array([1.0, 2.0]))
    bad = tmp_path / "bad.pcb"
    bad.write_bytes(b"XXXX" + good[4:])
    (tmp_path / "pts.f64").write_bytes(np.zeros(4).tobytes())
    res = subprocess.run([exe, str(bad), str(tmp_path / "pts.f64"), str(tmp_path / "out.f64")],
                         capture_output=True, text=True)
    assert res.returncode == 1 and "bad magic" in res.stderr
