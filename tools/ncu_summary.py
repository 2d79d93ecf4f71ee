"""Turn one `ncu --set full [--import-source on]` report into the two small text files kept under profiles/:
    python tools/ncu_summary.py REPORT.ncu-rep OUT_PREFIX [units]
  OUT_PREFIX_ncu_full_selected.csv  - the metrics the design notes quote (duration, DRAM bytes, occupancy limits, issue
                                      and pipe utilisation, stall reasons)
  OUT_PREFIX_opcode_histogram.txt   - executed SASS opcodes with stall samples (needs --import-source on)
`units` (optional) = how many work units the launch processed (lane-steps, heap steps ...) for the per-unit figure."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

SELECT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "lts__t_sectors.sum",
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return out[out.index('"'):]


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    names, unit_row, vals = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(names)}
    kernel = vals[col["Kernel Name"]]
    with open(prefix + "_ncu_full_selected.csv", "w") as f:
        f.write("metric,unit,value\n")
        f.write('Kernel Name,,"%s"\n' % kernel)
        for m in SELECT:
            if m in col:
                f.write("%s,%s,%s\n" % (m, unit_row[col[m]], vals[col[m]].replace(",", "")))
    src = ncu(rep, "source")
    lines = src.splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"Address"'))
    rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
    ex, st = defaultdict(int), defaultdict(int)
    for r in rd:
        if r["Source"] == "Source" or not (r.get("Address") or "").startswith("0x"):
            break                                   # the report holds more launches: the first one is summarised
        ins = r["Source"].split()
        if not ins:
            continue
        op = ins[1] if ins[0].startswith("@") and len(ins) > 1 else ins[0]
        op = ".".join(op.split(".")[:2]).rstrip(";")
        ex[op] += int(r["Instructions Executed"] or 0)
        st[op] += int(r["Warp Stall Sampling (All Samples)"] or 0)
    total = sum(ex.values())
    if total:
        with open(prefix + "_opcode_histogram.txt", "w") as f:
            f.write("# %s (ncu --set full --import-source on, SASS page)\n" % kernel)
            f.write("# warp-instructions executed per launch: %d%s\n" % (total, "  (= %.1f per unit, %g units)" % (total / units, units) if units else ""))
            f.write("# opcode  executed  share  stall-samples\n")
            for op, n in sorted(ex.items(), key=lambda kv: -kv[1])[:40]:
                f.write("%-24s %12d  %4.1f%%  %d\n" % (op, n, 100.0 * n / total, st[op]))
    print("wrote", prefix + "_ncu_full_selected.csv", "and the opcode histogram" if total else "(no source page)")


if __name__ == "__main__":
    main()
