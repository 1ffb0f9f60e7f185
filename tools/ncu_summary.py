"""Summarise an ncu capture for profiles/ (run here, where ncu can read reports without a GPU).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep gpurun_out/launches.csv > profiles/rNN_ncu_summary.txt

Part 1: per-kernel share of one step from the launch list (cold-cache, serialised: compare shares, not absolutes).
Part 2: the full-section metrics bench.py's roofline object and DESIGN.md quote, plus the top warp-stall reasons.
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__inst_executed.sum",
]


def short(name):
    name = name.replace("smh::", "").replace("void ", "")
    return name.split("(")[0]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    per = OrderedDict()
    for r in rows:
        per.setdefault(short(r[4]), []).append(float(r[14]))
    print("== launch list: %s (%d launches; gpu__time_duration.sum, ns)" % (path, len(rows)))
    ours = {k: v for k, v in per.items() if not k.startswith("at::")}
    # bench.py runs the step in both weight modes: kernels of the exact-weights step are the <.., 0, ..> (Q16 = false)
    # instantiations of mpjpe_kernel<Q16, FUSED> / sweep_tc_kernel<BWD, SBF16, Q16, FUSED>; the small kernels are shared
    def is_exact(k):
        if k.startswith("mpjpe_kernel<"):
            return k.split("<")[1].split(",")[0].strip() == "0"
        if k.startswith("sweep_tc_kernel<"):
            return k.split("<")[1].split(",")[2].strip() == "0"
        return False
    shared = {k: v for k, v in ours.items() if not k.startswith(("mpjpe_kernel<", "sweep_tc_kernel<"))}
    for title, pick in (("relaxed-weights step (default)", lambda k: not is_exact(k)), ("exact-weights step", is_exact)):
        mine = {k: v for k, v in ours.items() if k not in shared and pick(k)}
        if not mine:
            continue
        group = dict(shared)
        group.update(mine)
        step = sum(sum(v) / len(v) for v in group.values())
        print("  -- %s" % title)
        for k, v in group.items():
            avg = sum(v) / len(v)
            print("  %-28s n=%-3d avg %10.0f ns   share of step %5.1f %%" % (k, len(v), avg, 100 * avg / step))
        print("  %-28s       sum %10.0f ns" % ("one step (serialised)", step))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("\n== full-section capture: %s" % path)
    for r in rows[2:]:
        print("---- %s  grid %s block %s" % (short(r[hdr.index("Kernel Name")]), r[hdr.index("Grid Size")],
                                              r[hdr.index("Block Size")]))
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                print("   %-64s %18s %s" % (k, r[i], units[i]))
        st = [(float(r[i]), h) for i, h in enumerate(hdr)
              if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i] not in ("", "n/a")]
        for v, h in sorted(st, reverse=True)[:6]:
            print("   stall %-58s %18.3f per issue" % (h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v))


if __name__ == "__main__":
    if len(sys.argv) > 2:
        launches(sys.argv[2])
    full(sys.argv[1])
