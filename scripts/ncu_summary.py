"""Summarise ncu outputs into small text files for profiles/ (run here, no GPU needed).

  python scripts/ncu_summary.py rep  gpurun_out/x.ncu-rep  profiles/out.txt     # key metrics per captured launch
  python scripts/ncu_summary.py list gpurun_out/launches.csv profiles/out.txt   # per-kernel totals of a launch list
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__sass_inst_executed_op_shared_ld.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as fh:
        fh.write("# ncu --set full --clock-control none summary of %s\n" % path)
        for r in data:
            fh.write("\nkernel: %s\n" % r[hdr.index("Kernel Name")])
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    fh.write("  %-75s %s %s\n" % (k, r[i], units[i]))
    print(open(out).read())


def lst(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    tot = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v_us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        key = (name, r.get("Grid Size", ""))
        t = tot.setdefault(key, [0, 0.0])
        t[0] += 1
        t[1] += v_us
    total = sum(t[1] for t in tot.values())
    with open(out, "w") as fh:
        fh.write("# ncu launch list %s: per (kernel, grid) totals; cold-cache serialised times: compare SHARES\n" % path)
        fh.write("# total %.1f us over %d launches\n" % (total, sum(t[0] for t in tot.values())))
        for (name, grid), (cnt, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            fh.write("%6.2f%%  %10.1f us  %4d x  %s  grid %s\n" % (100 * us / total, us, cnt, name[:110], grid))
    print(open(out).read()[:4000])


if __name__ == "__main__":
    {"rep": rep, "list": lst}[sys.argv[1]](sys.argv[2], sys.argv[3])
