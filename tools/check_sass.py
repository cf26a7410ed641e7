#!/usr/bin/env python
"""Which kernels of libpcb_b200.so really read their broadcast operands on the uniform datapath?

The constant-bank kernels (pcb_tt_const.cu, the *_bank_kernel of pcb_piecewise.cu) are only fast
when ptxas emits `LDCU.64 UR, c[0x3][UR+imm]` feeding `DFMA R, R, UR, R`; whether it does depends
on ptxas' uniformity analysis (DESIGN.md section K-C lists what breaks it).  This script counts, per
kernel, the bank reads on the uniform path (LDCU) and the per-lane ones (LDC) in the SASS of the
built library, and writes a markdown table.  Runs on the CPU container (cuobjdump only).

    python tools/check_sass.py [--out profiles/r1_sass_uniform.md]
"""
import argparse
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pychebyshev_b200", "libpcb_b200.so")

PATTERNS = ("ttc_value_kernel", "ttc_fd_shared_kernel", "ttc_gstep_kernel", "ttc_gcoeff_kernel",
            "spline_bank_kernel", "slider_bank_kernel", "spline2d_dmma_kernel", "slider2d_dmma_kernel",
            "spline3d_dmma_kernel")

#: kernels the planner selects by default (pcb_tt_const.cu kValueVariants / kSharedVariants first
#: entries per rank class, the per-core step kernel, the bank evaluators the benchmark configurations
#: use, the tensor-core 2-D kernels): these MUST read their bank operands on the uniform datapath.
#: tests/test_sass_uniform.py fails when one of them regresses.
MUST_BE_UNIFORM = (
    "ttc_value_kernel<2, 8, 512, 2>", "ttc_value_kernel<2, 12, 512, 2>", "ttc_value_kernel<2, 16, 512, 0>",
    "ttc_fd_shared_kernel<2, 8, 512, 0>", "ttc_fd_shared_kernel<2, 12, 512, 0>",
    "ttc_fd_shared_kernel<2, 16, 384, 0>",
    "spline_bank_kernel<1, 2>", "spline_bank_kernel<2, 2>", "spline_bank_kernel<2, 3>",
    "spline_bank_kernel<4, 2>", "spline_bank_kernel<4, 3>", "spline_bank_kernel<4, 4>",
    "slider_bank_kernel<1, 2>", "slider_bank_kernel<2, 2>", "slider_bank_kernel<4, 2>",
    "spline2d_dmma_kernel<0>", "spline2d_dmma_kernel<11>", "spline2d_dmma_kernel<15>",
    "slider2d_dmma_kernel<0>", "spline3d_dmma_kernel<1, 0, 0>",
)
#: at most this share of a MUST_BE_UNIFORM kernel's bank reads may be per-lane (descriptor fields
#: that feed per-lane predicates are legitimately read with LDC)
MAX_LDC_SHARE = 0.15
#: known partial cases (correct, slower), tracked so that NEW regressions stand out
KNOWN_PARTIAL = ("ttc_gcoeff_kernel<2, 256>", "ttc_gstep_kernel<2, 256>", "spline_bank_kernel<1, 3>",
                 "spline_bank_kernel<1, 4>", "slider_bank_kernel<1, 3>", "slider_bank_kernel<4, 4>",
                 # fixed-node-count variants: per-lane bank reads, but straight-line weight rows --
                 # measured FASTER than the generic kernels on the uniform path (slider 11 x 11:
                 # 5.24e9 against 4.74e9 q/s; 15^3 pieces: 2.81e9 against 2.71e9), and slower when a
                 # register cap forces them back onto it
                 "slider2d_dmma_kernel<8>", "slider2d_dmma_kernel<11>", "slider2d_dmma_kernel<12>",
                 "slider2d_dmma_kernel<15>", "spline3d_dmma_kernel<1, 0, 8>",
                 "spline3d_dmma_kernel<1, 0, 11>", "spline3d_dmma_kernel<1, 0, 15>",
                 "spline3d_dmma_kernel<1, 0, 16>")


def demangle_short(name):
    m = re.search(r"(" + "|".join(PATTERNS) + r")I(.*?)EEv", name)
    if not m:
        m2 = re.search(r"(" + "|".join(PATTERNS) + r")", name)
        return m2.group(1) if m2 else name
    args = re.findall(r"Li(\d+)E", m.group(2))
    return f"{m.group(1)}<{', '.join(args)}>"


def scan(lib=None):
    """[(kernel, ldcu, ldc, fp64_with_ur, fp64)] of the constant-bank kernels in the built library."""
    lib = lib or LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    rows = []
    for blk in sass.split("Function : ")[1:]:
        name = blk.split("\n", 1)[0].strip()
        if not any(p in name for p in PATTERNS):
            continue
        ldcu = len(re.findall(r"LDCU(?:\.64)? UR\d+, c\[0x3\]\[UR", blk))
        ldc = len(re.findall(r"LDC(?:\.64)? R\d+, c\[0x3\]\[R", blk))
        f64_ur = len(re.findall(r"D(?:FMA|MUL|ADD) [^;]*UR\d+", blk))
        f64 = len(re.findall(r"\bD(?:FMA|MUL|ADD) ", blk))
        dmma = len(re.findall(r"\bDMMA", blk))
        rows.append((demangle_short(name), ldcu, ldc, f64_ur, f64, dmma))
    rows.sort()
    return rows


def regressions(rows):
    """MUST_BE_UNIFORM kernels that are missing or read too many bank operands per lane."""
    by = {r[0]: r for r in rows}
    bad = []
    for k in MUST_BE_UNIFORM:
        r = by.get(k)
        if r is None:
            bad.append((k, "not in the library"))
        elif r[1] + r[2] == 0 or r[2] / (r[1] + r[2]) > MAX_LDC_SHARE:
            bad.append((k, f"{r[2]} of {r[1] + r[2]} bank reads are per-lane (LDC)"))
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    if not os.path.exists(LIB):
        sys.exit(f"{LIB} not built")
    rows = scan()
    lines = ["| kernel | bank reads LDCU (uniform) | bank reads LDC (per lane) | fp64 ops with UR operand | fp64 ops | DMMA | planner default |",
             "|---|---:|---:|---:|---:|---:|---|"]
    bad = 0
    for r in rows:
        share = r[2] / max(1, r[1] + r[2])
        flag = "" if share <= MAX_LDC_SHARE else " (!)"
        bad += share > MAX_LDC_SHARE
        lines.append(f"| `{r[0]}`{flag} | {r[1]} | {r[2]} | {r[3]} | {r[4]} | {r[5]} | "
                     f"{'yes' if r[0] in MUST_BE_UNIFORM else ''} |")
    text = "\n".join(lines)
    print(text)
    print(f"\n{len(rows)} constant-bank kernels, {bad} with more than {MAX_LDC_SHARE:.0%} per-lane bank reads")
    reg = regressions(rows)
    for k, why in reg:
        print(f"REGRESSION: {k}: {why}")
    if args.out:
        with open(args.out, "w") as f:
            f.write("# Uniform-datapath check of the constant-bank kernels (static, from cuobjdump -sass)\n\n"
                    "`python tools/check_sass.py` on the built library.  A kernel marked (!) reads more than 15 % of\n"
                    "its bank operands per lane (`LDC`), i.e. ptxas did not keep those addresses in uniform\n"
                    "registers; it is correct but slower.  `planner default` kernels are held on the uniform\n"
                    "path by tests/test_sass_uniform.py.\n\n" + text + "\n")
    sys.exit(1 if reg else 0)


if __name__ == "__main__":
    main()
