#!/usr/bin/env python
"""Pipe / dispatch / stall summary of one kernel from `ncu --set full --import-source on` reports (read here with ncu -i).

    python tools/ncu_pipes.py LABEL=report.ncu-rep[@N] [LABEL=report.ncu-rep[@N] ...] > profiles/rNN_ncu_..._pipes.txt

For every report: duration, pipe utilisation, instructions per warp, and the stall samples aggregated by opcode and reason
(the source page), which is what tells a pipe-bound kernel from a dispatch-bound one."""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum"]


def page(rep, name):
    sel = []
    if "@" in rep:                         # report@N: the N-th kernel of a report with several
        rep, n = rep.rsplit("@", 1)
        sel = ["--launch-skip", n, "--launch-count", "1"]
    out = subprocess.run(["ncu", "-i", rep, *sel, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    for arg in sys.argv[1:]:
        label, rep = arg.split("=", 1)
        raw = page(rep, "raw")
        hdr, units, val = raw[0], raw[1], raw[-1]
        d = dict(zip(hdr, val))
        u = dict(zip(hdr, units))
        print(f"==== {label}: {d.get('Kernel Name', '?')}  ({rep.split('/')[-1]})")
        for m in METRICS:
            if m in d:
                print(f"   {m:82s} {d[m]:>16s} {u.get(m, '')}")
        stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: float(v) for k, v in d.items()
                  if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:7]
        print("   stalls per issue: " + ", ".join(f"{k} {v:.2f}" for k, v in top))
        src = page(rep, "source")
        # a "Kernel Name" row, a header row, then one row per instruction; a kernel can come with a second, identical table
        # (another view of the same code): the first table of the kernel the raw page names is used.
        norm = lambda n: re.sub(r"\(bool\)|smh::|\s", "", n).split("(")[0]  # noqa: E731
        starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"] + [len(src)]
        mine = [(a, b) for a, b in zip(starts[:-1], starts[1:]) if norm(src[a][1]) == norm(d.get("Kernel Name", ""))]
        mine = mine[:1] if mine else [(starts[0], starts[1])]
        h = src[mine[0][0] + 1]
        ix = {n: i for i, n in enumerate(h)}
        src = [r for a, b in mine for r in src[a:b]]
        reasons = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        samples, by_op, execd = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
        for r in src:
            if len(r) < len(h) or not r[ix["Source"]].strip() or r[ix["# Samples"]] == "# Samples":
                continue
            op = re.sub(r"^@!?U?P\d\s+", "", r[ix["Source"]].strip()).split()[0]
            samples[op] += int(r[ix["# Samples"]] or 0)
            execd[op] += int(r[ix["Instructions Executed"]] or 0)
            for c in reasons:
                by_op[op][c.replace("stall_", "")] += int(r[ix[c]] or 0)
        total, n_exec = sum(samples.values()), sum(execd.values())
        packed = sum(v for k, v in execd.items() if k in ("FADD2", "FMUL2", "FFMA2"))
        print(f"   warp instructions executed {n_exec}, of them packed fp32x2 {packed} ({100.0 * packed / max(n_exec, 1):.1f} %): "
              f"dispatch cycles ~ executed + packed = {n_exec + packed}")
        print("   stall samples by opcode (share of all samples; top reasons):")
        for op, n in samples.most_common(14):
            rs = ", ".join(f"{k} {100.0 * x / total:.1f}" for k, x in by_op[op].most_common(4))
            print(f"      {op:12s} {100.0 * n / total:5.1f} %   executed {execd[op]:>11d}   {rs}")


if __name__ == "__main__":
    main()
