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
            "spline_bank_kernel", "slider_bank_kernel")


def demangle_short(name):
    m = re.search(r"(" + "|".join(PATTERNS) + r")I(.*?)EEv", name)
    if not m:
        return name
    args = re.findall(r"Li(\d+)E", m.group(2))
    return f"{m.group(1)}<{', '.join(args)}>"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    if not os.path.exists(LIB):
        sys.exit(f"{LIB} not built")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    rows = []
    for blk in sass.split("Function : ")[1:]:
        name = blk.split("\n", 1)[0].strip()
        if not any(p in name for p in PATTERNS):
            continue
        ldcu = len(re.findall(r"LDCU\.64 UR\d+, c\[0x3\]\[UR", blk))
        ldc = len(re.findall(r"LDC\.64 R\d+, c\[0x3\]\[R", blk))
        f64_ur = len(re.findall(r"D(?:FMA|MUL|ADD) [^;]*UR\d+", blk))
        f64 = len(re.findall(r"\bD(?:FMA|MUL|ADD) ", blk))
        rows.append((demangle_short(name), ldcu, ldc, f64_ur, f64))
    rows.sort()
    lines = ["| kernel | bank reads LDCU (uniform) | bank reads LDC (per lane) | fp64 ops with UR operand | fp64 ops |",
             "|---|---:|---:|---:|---:|"]
    bad = 0
    for r in rows:
        flag = "" if r[2] == 0 else " (!)"
        bad += r[2] != 0
        lines.append(f"| `{r[0]}`{flag} | {r[1]} | {r[2]} | {r[3]} | {r[4]} |")
    text = "\n".join(lines)
    print(text)
    print(f"\n{len(rows)} uniform-path kernels, {bad} with per-lane bank reads left")
    if args.out:
        with open(args.out, "w") as f:
            f.write("# Uniform-datapath check of the constant-bank kernels (static, from cuobjdump -sass)\n\n"
                    "`python tools/check_sass.py` on the built library.  A kernel marked (!) still reads part of\n"
                    "its bank operands per lane (`LDC`), i.e. ptxas did not keep those addresses in uniform\n"
                    "registers; it is correct but slower.\n\n" + text + "\n")


if __name__ == "__main__":
    main()
