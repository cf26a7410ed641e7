"""The .pcb v1 layout: byte-identical with files the reference wrote (its own fixtures under
tests/fixtures/ and fresh reference-written bytes, all stored in tests/golden/pcb_files.npz),
field offsets of the spec, and the reference's error behaviour on corrupt input
(reference tests/test_binary_format.py:562-770)."""

import hashlib
import os
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


# ------------------------------------------------------------------------------------------
# native writer / round trip / grid recipes, no GPU needed
# ------------------------------------------------------------------------------------------

def _fixture_bytes(gold, name):
    """Reference-written fixture bytes; approx_5d_bs (62 KB) is rebuilt from its stored tensor and
    checked against the stored sha256 of the reference's file."""
    if name == "approx_5d_bs":
        raw = pcbfile.approx_bytes([[-1.0, 1.0]] * 5, [6] * 5, gold["approx_5d_bs_tensor"])
    else:
        raw = gold[name + "_bytes"].tobytes()
    assert hashlib.sha256(raw).hexdigest() == str(gold[name + "_sha256"])
    return raw


@pytest.mark.parametrize("name", ["approx_2d_simple", "approx_5d_bs", "spline_1d_kink"])
def test_native_reader_writer_round_trip_is_byte_identical(gold, tmp_path, name):
    """bytes -> native parser -> native writer -> bytes, on the reference's own fixture files."""
    from pychebyshev_b200 import _lib

    raw = _fixture_bytes(gold, name)
    src, dst = tmp_path / "in.pcb", tmp_path / "out.pcb"
    src.write_bytes(raw)
    _lib.check(_lib.load().pcb_file_rewrite(str(src).encode(), str(dst).encode()))
    assert dst.read_bytes() == raw


def test_native_writer_matches_reference_written_bytes(gold, tmp_path):
    from pychebyshev_b200 import _engine

    p = tmp_path / "a.pcb"
    _engine.write_pcb_native(p, [[0.0, 1.0], [-2.0, 2.0]], [3, 4], gold["approx_3x4_tensor"])
    assert p.read_bytes() == gold["approx_3x4_bytes"].tobytes()
    with pytest.raises(ValueError, match="NaN or Inf"):
        _engine.write_pcb_native(p, [[0.0, 1.0]], [2], np.array([1.0, np.nan]))
    with pytest.raises(ValueError, match="must be <"):
        _engine.write_pcb_native(p, [[1.0, 1.0]], [2], np.array([1.0, 2.0]))


@pytest.mark.parametrize("name", ["spline_bs2d", "spline_bs3d", "spline_multiknot3d", "spline_abs1d"])
def test_native_spline_writer_matches_reference_bytes(name, tmp_path):
    from oracle import np_oracle as O
    from pychebyshev_b200 import _engine

    g = G.load(name)
    knots, shape, pieces = G.spline_parts(g, O.diff_matrix)
    n = [int(v) for v in g["piece_n_nodes"][0]]
    dom = [list(map(float, r)) for r in g["domain"]]
    p = tmp_path / "s.pcb"
    _engine.write_pcb_native(p, dom, n, [q[0] for q in pieces], knots=knots)
    assert p.read_bytes() == g["pcb_bytes"].tobytes()


def test_native_grid_recipes_against_numpy():
    """Nodes / weights / differentiation matrix of the native loader vs the NumPy recipes
    (barycentric.py:43-77, _extrude_slice.py:66-70).  libm's sin and NumPy's may differ in the last
    bit, so nodes agree to 1 ulp; given the native nodes, weights and matrix (incl. NumPy's
    pairwise row sums on the diagonal) are bit-identical."""
    import ctypes as C

    from pychebyshev_b200 import _grid, _lib

    lib = _lib.load()
    for lo, hi, n in ((80.0, 120.0, 11), (0.25, 1.0, 16), (-1.0, 1.0, 7), (0.01, 0.08, 33), (0.0, 3.0, 2)):
        x, w, dm = np.empty(n), np.empty(n), np.empty((n, n))
        _lib.check(lib.pcb_file_grid_arrays(lo, hi, n, x.ctypes.data_as(_lib._f64p),
                                            w.ctypes.data_as(_lib._f64p),
                                            dm.ctypes.data_as(_lib._f64p)))
        ref_x = _grid.cheb_nodes(lo, hi, n)
        assert np.all(np.abs(x - ref_x) <= np.spacing(np.abs(ref_x)))
        assert np.array_equal(w, _grid.bary_weights(x))
        assert np.array_equal(dm, _grid.diff_matrix(x, w))
    del C


def test_native_loader_bounds_sizes_by_the_file_length(tmp_path):
    """Crafted headers (huge n_nodes / num_knots / piece counts) are rejected with the reference's
    EOF error before anything is allocated; no exception crosses the C ABI."""
    import ctypes as C

    from pychebyshev_b200 import _lib

    lib = _lib.load()
    head = b"PCB\x00" + bytes([1, 0]) + struct.pack("<H", 1) + b"\x00" * 4

    def approx(n_nodes, payload=b""):
        D = len(n_nodes)
        return (head + struct.pack("<I", D) + struct.pack(f"<{D}d", *([0.0] * D)) +
                struct.pack(f"<{D}d", *([1.0] * D)) + struct.pack(f"<{D}I", *n_nodes) + payload)

    shead = b"PCB\x00" + bytes([1, 0]) + struct.pack("<H", 2) + b"\x00" * 4

    def spline(num_knots, pieces):
        return (shead + struct.pack("<I", 1) + struct.pack("<d", 0.0) + struct.pack("<d", 1.0) +
                struct.pack("<I", 4) + struct.pack("<I", num_knots) + struct.pack("<I", pieces))

    cases = [approx([0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF]), approx([0x7FFFFFFF] * 8),
             approx([1 << 20, 1 << 20], b"\x00" * 64), spline(0xFFFFFFF0, 1), spline(3, 0xFFFFFFFF)]
    for raw in cases:
        path = tmp_path / "evil.pcb"
        path.write_bytes(raw)
        plan = C.c_void_p()
        rc = lib.pcb_plan_from_file(0, str(path).encode(), C.byref(plan), None, None)
        assert rc == _lib.PCB_EINVAL and "unexpected EOF" in _lib.last_error(), _lib.last_error()
        rc = lib.pcb_file_rewrite(str(path).encode(), str(tmp_path / "o.pcb").encode())
        assert rc == _lib.PCB_EINVAL


@pytest.mark.gpu
def test_reference_fixture_approx_5d_through_the_native_loader(gold, tmp_path):
    """tests/fixtures/approx_5d_bs.pcb (6^5): native file -> plan -> values the reference reads out
    of the same file; with derivative rows: the Python host's plan for the same file."""
    raw = _fixture_bytes(gold, "approx_5d_bs")
    path = tmp_path / "approx_5d_bs.pcb"
    path.write_bytes(raw)
    plan = pcb.load_plan(path)
    assert plan.kind == "approx" and plan.ndim == 5 and plan.G == 1
    got = plan.eval(gold["approx_5d_bs_points"])[:, 0]
    G.assert_close_scaled(got, gold["approx_5d_bs_values"], 1.0, "approx_5d_bs native loader")
    orders = [[0] * 5, [1, 0, 0, 0, 0], [0, 0, 2, 0, 0], [0, 1, 0, 1, 0]]
    plan_g = pcb.load_plan(path, orders=orders)
    assert plan_g.G == 4
    host = pcb.ChebyshevApproximation.load(path)
    want = host.eval_batch_multi(gold["approx_5d_bs_points"], orders)
    got = plan_g.eval(gold["approx_5d_bs_points"])
    for j in range(4):
        G.assert_close_scaled(got[:, j], want[:, j], 1.0, f"file plan order {orders[j]}")
    with pytest.raises(ValueError, match="entries"):
        pcb.load_plan(path, orders=[[0, 0]])


@pytest.mark.gpu
def test_spline_file_plan_with_derivative_rows(tmp_path):
    g = G.load("spline_bs2d")
    path = tmp_path / "s.pcb"
    path.write_bytes(g["pcb_bytes"].tobytes())
    orders = [[0, 0], [1, 0], [0, 1]]
    plan = pcb.load_plan(path, orders=orders)
    assert plan.kind == "spline" and plan.G == 3
    host = pcb.ChebyshevSpline.load(path)
    pts = g["points"]
    want = host.eval_batch_multi(pts, orders)
    got = plan.eval(pts)
    inside = (pts[:, 0] >= 80.0) & (pts[:, 0] <= 120.0) & (pts[:, 1] >= 0.25) & (pts[:, 1] <= 1.0)
    for j in range(3):
        G.assert_close_scaled(got[inside, j], want[inside, j], 1.0, f"spline file plan {orders[j]}")


def _reference_round_trip(raw, path, out):
    """bytes -> object -> bytes through the UNMODIFIED reference's reader and writer
    (``_binary.read_approx / read_spline`` construct through ``from_values``, which also rejects
    non-finite tensors), or through the mirror classes when the reference is not installed.
    Returns the re-written bytes, or None when the file is rejected."""
    try:
        from oracle import reference as R

        ref = R.load()
    except Exception:  # noqa: BLE001
        ref = None
    tag = struct.unpack("<H", raw[6:8])[0] if len(raw) >= 8 else 0
    try:
        if ref is not None:
            from pychebyshev import _binary

            with open(path, "rb") as f:
                obj = _binary.read_approx(f) if tag != 2 else _binary.read_spline(f)
            with open(out, "wb") as f:
                (_binary.write_approx if tag != 2 else _binary.write_spline)(f, obj)
        else:
            obj = (pcb.ChebyshevApproximation if tag != 2 else pcb.ChebyshevSpline).load(path)
            obj.save(out)
        return out.read_bytes()
    except (ValueError, MemoryError, OverflowError, EOFError):
        return None


def test_native_parser_differential_fuzz(gold, tmp_path):
    """Seeded mutations of valid files (bit flips, truncations, extreme 32-bit fields, spliced and
    inserted bytes): the native parser must never crash or let an exception cross the C ABI, must
    accept exactly the files the reference's own reader accepts, and must rewrite every accepted
    file to the bytes the reference's writer produces.  Host code only -- no GPU involved."""
    from pychebyshev_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(20261019)
    seeds = [gold[k].tobytes() for k in ("approx_2d_simple_bytes", "spline_1d_kink_bytes")]
    seeds.append(pcbfile.spline_bytes([[0.0, 1.0], [-1.0, 2.0]], [3, 4], [[0.5], [0.0, 1.0]],
                                      [np.arange(12.0).reshape(3, 4) + p for p in range(6)]))
    seeds.append(pcbfile.approx_bytes([[0.0, 1.0]] * 3, [2, 3, 2], np.arange(12.0).reshape(2, 3, 2)))
    extreme = [0, 1, 2, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 0xFFFFFFF0, 1 << 20, 65537]
    inp, outp, refp, mirp = (tmp_path / n for n in ("m.pcb", "o.pcb", "r.pcb", "p.pcb"))
    accepted = rejected = 0
    for it in range(int(os.environ.get("PCB_FUZZ_ITERS", "600"))):
        raw = bytearray(seeds[it % len(seeds)])
        kind = it % 5
        if kind == 0:      # single bit flip anywhere
            pos = int(rng.integers(len(raw)))
            raw[pos] ^= 1 << int(rng.integers(8))
        elif kind == 1:    # truncate
            raw = raw[: int(rng.integers(len(raw)))]
        elif kind == 2:    # an extreme value in a 4-byte-aligned field of the first 96 bytes
            pos = 4 * int(rng.integers(min(len(raw), 96) // 4))
            raw[pos:pos + 4] = struct.pack("<I", extreme[int(rng.integers(len(extreme)))])
        elif kind == 3:    # random bytes spliced over a window
            pos = int(rng.integers(len(raw)))
            n = int(rng.integers(1, 9))
            raw[pos:pos + n] = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        else:              # trailing or inserted bytes
            pos = int(rng.integers(len(raw) + 1))
            raw[pos:pos] = bytes(rng.integers(0, 256, int(rng.integers(1, 17)), dtype=np.uint8))
        raw = bytes(raw)
        inp.write_bytes(raw)
        want = _reference_round_trip(raw, inp, refp)
        rc = lib.pcb_file_rewrite(str(inp).encode(), str(outp).encode())
        assert rc in (_lib.PCB_OK, _lib.PCB_EINVAL, _lib.PCB_ENOMEM), (it, rc)
        tag = struct.unpack("<H", raw[6:8])[0] if len(raw) >= 8 else 0
        try:  # the Python mirror classes hold to the same contract
            obj = (pcb.ChebyshevApproximation if tag != 2 else pcb.ChebyshevSpline).load(inp)
            obj.save(mirp)
            mine = mirp.read_bytes()
        except Exception:  # noqa: BLE001  (a broken magic sends load() to pickle, like the reference's)
            mine = None
        assert mine == want, (it, kind, "mirror classes and the reference disagree on this file")
        if want is None:
            assert rc != _lib.PCB_OK, (it, kind, "native accepted a file the reference's reader rejects")
            rejected += 1
        else:
            assert rc == _lib.PCB_OK, (it, kind, _lib.last_error())
            assert outp.read_bytes() == want, (it, kind)
            accepted += 1
    assert accepted > 50 and rejected > 50, (accepted, rejected)
