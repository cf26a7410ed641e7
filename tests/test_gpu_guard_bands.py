"""Out-of-bounds evidence without a sanitizer (compute-sanitizer is closed on this GPU pool).

Every kernel family is run with its query and result arrays embedded in larger device buffers that
are filled with a canary bit pattern, at ragged batch sizes (no multiple of any tile, warp or
vector width) and at 8-byte-only alignment:

* a write past either end of the result array changes the canaries  -> compared bit for bit;
* a write into the query array changes the queries                   -> compared bit for bit;
* a read past either end of the query array that reaches a result turns it into the canary's NaN
  or at least changes it -> the results must equal, bit for bit, those of the same queries
  evaluated from a plain tensor in a batch of another size.

The workloads are bench.py's (BASELINE.json's configs), so these are the kernels the bench times.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CANARY = -0x000DEAD0BEEF0001          # as float64: a NaN with a recognisable payload
SIZES = [1, 2, 31, 33, 127, 129, 255, 1023, 1025, 4097, 33_333]
KEYS = ["tt_bs5d", "tt_bs5d_value", "tt_basket10d", "tt_rank20", "full_bs5d", "spline2d_lookup",
        "spline2d", "spline2d_greeks", "spline3d", "spline3d_greeks", "slider10d"]


@pytest.mark.parametrize("key", KEYS)
def test_kernels_stay_inside_their_arrays(key):
    import torch

    import bench

    case = bench.make_case(key)
    dev = torch.device("cuda", 0)
    plan = case.build(0)
    D, G = case.D, case.G
    out_dtype = torch.int32 if case.out_dtype == "int32" else torch.float64
    nmax = max(SIZES)
    pts_all = bench.device_queries(case.domain, nmax, dev, 99)
    plain = torch.empty((nmax, G), dtype=out_dtype, device=dev)
    case.launch(plan, pts_all, plain)                      # reference run: one big plain batch
    torch.cuda.synchronize()
    assert not torch.isnan(plain.double()).any()

    for guard in (1024, 1023):                             # 16-byte aligned / 8-byte aligned only
        for n in SIZES:
            if key == "full_bs5d" and n > 4097:
                continue
            pbuf = torch.full((2 * guard + n * D,), CANARY, dtype=torch.int64, device=dev)
            pts = pbuf[guard: guard + n * D].view(torch.float64).view(n, D)
            pts.copy_(pts_all[:n])
            before = pbuf.clone()
            if out_dtype == torch.int32:
                g32 = 2 * guard                            # same byte offsets as the float64 case
                obuf = torch.full((2 * g32 + n * G,), 0x5EEDBEEF, dtype=torch.int32, device=dev)
                out = obuf[g32: g32 + n * G].view(n, G)
                canary = 0x5EEDBEEF
                lo, hi = obuf[:g32], obuf[g32 + n * G:]
            else:
                obuf = torch.full((2 * guard + n * G,), CANARY, dtype=torch.int64, device=dev)
                out = obuf[guard: guard + n * G].view(torch.float64).view(n, G)
                canary = CANARY
                lo, hi = obuf[:guard], obuf[guard + n * G:]
            case.launch(plan, pts, out)
            torch.cuda.synchronize()
            what = f"{key} n={n} guard={guard}"
            assert bool((lo == canary).all()) and bool((hi == canary).all()), f"{what}: wrote outside the result array"
            assert torch.equal(pbuf, before), f"{what}: wrote into the query buffer"
            ivw = torch.int32 if out_dtype == torch.int32 else torch.int64
            ref_n = torch.empty((n, G), dtype=out_dtype, device=dev)
            case.launch(plan, pts_all[:n].clone(), ref_n)     # same batch from a plain tensor
            torch.cuda.synchronize()
            assert torch.equal(out.view(ivw), ref_n.view(ivw)), (
                f"{what}: results differ from the same batch in a plain tensor (a read outside the "
                f"query array reached a result)")
            if n >= 64:                                       # below 64 the full tensor takes its scalar kernel
                same = out.view(ivw) == plain[:n].view(ivw)
                assert bool(same.all()), (
                    f"{what}: {int((~same).sum())} results depend on the batch size or tile position")
