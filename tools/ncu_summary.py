#!/usr/bin/env python
"""Turn ncu artefacts in gpurun_out/ into the small text summaries kept under profiles/.

  ncu_summary.py launches <launches.csv> <out.md>        per-kernel time shares of one command
  ncu_summary.py kernel   <file.ncu-rep> <out.md>        key metrics + stall mix + SASS evidence
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def launches(src, out):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ix["Metric Unit"]], 1.0)
        a = agg.setdefault(r[ix["Kernel Name"]], [0, 0.0, r[ix["Grid Size"]], r[ix["Block Size"]]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"source: {src} (ncu --metrics gpu__time_duration.sum --clock-control none; "
                "cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total ms | share | grid | block |\n|---|---|---|---|---|---|\n")
        for k, (n, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:110]}` | {n} | {t / 1e6:.3f} | {100 * t / tot:.1f}% | {g} | {b} |\n")


def kernel(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"source: {rep} (ncu --set full --clock-control none --import-source on)\n\n")
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"## {name}\n\n| metric | value | unit |\n|---|---|---|\n")
            for h, u, v in zip(hdr, units, vals):
                if h in KEYS:
                    f.write(f"| {h} | {v} | {u} |\n")
            f.write("\nwarp-issue stall mix (per issue-active):\n\n| stall | ratio |\n|---|---|\n")
            stalls = [(h, v) for h, v in zip(hdr, vals)
                      if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
                      and "not_issued" not in h]
            for h, v in sorted(stalls, key=lambda kv: -float(kv[1] or 0))[:8]:
                f.write(f"| {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} | {v} |\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(src.splitlines()))
        if len(srows) > 2:
            h2 = srows[1]
            si, ni, xi = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
            ops = collections.Counter()
            samples = collections.Counter()
            for r in srows[2:]:
                if len(r) <= max(si, ni, xi):
                    continue
                toks = r[si].split()
                if not toks or not (r[xi] or '0').isdigit() or not (r[ni] or '0').isdigit():
                    continue  # blank line or a repeated header (several launches in one report)
                op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
                ops[op] += int(r[xi] or 0)
                samples[op] += int(r[ni] or 0)
            tot_i, tot_s = sum(ops.values()) or 1, sum(samples.values()) or 1
            f.write("\nSASS opcode mix (warp instructions executed / stall samples):\n\n"
                    "| opcode | executed | share | samples share |\n|---|---|---|---|\n")
            for op, n in ops.most_common(10):
                f.write(f"| {op} | {n} | {100 * n / tot_i:.1f}% | {100 * samples[op] / tot_s:.1f}% |\n")
            # where in the kernel the stall samples fall: 20 equal bins of the SASS listing (program
            # order), and the 12 hottest instructions
            body = [(r[si], int(r[ni] or 0), int(r[xi] or 0)) for r in srows[2:]
                    if len(r) > max(si, ni, xi) and r[si].split() and (r[xi] or '0').isdigit()
                    and (r[ni] or '0').isdigit()]
            if body:
                nb = 20
                f.write("\nstall samples along the SASS listing (20 bins in program order):\n\n"
                        "| bin | instructions | executed share | samples share | most sampled instruction |\n"
                        "|---|---|---|---|---|\n")
                per = max(1, (len(body) + nb - 1) // nb)
                for b in range(0, len(body), per):
                    chunk = body[b:b + per]
                    hot = max(chunk, key=lambda t: t[1])
                    f.write(f"| {b // per} | {b}-{b + len(chunk) - 1} | "
                            f"{100 * sum(t[2] for t in chunk) / tot_i:.1f}% | "
                            f"{100 * sum(t[1] for t in chunk) / tot_s:.1f}% | `{hot[0][:70]}` |\n")
                f.write("\nhottest instructions:\n\n| # | samples share | executed | SASS |\n|---|---|---|---|\n")
                order = sorted(range(len(body)), key=lambda i: -body[i][1])[:12]
                for i in order:
                    f.write(f"| {i} | {100 * body[i][1] / tot_s:.1f}% | {body[i][2]} | `{body[i][0][:80]}` |\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
