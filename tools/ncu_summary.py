"""Summarise an .ncu-rep (read here on the CPU box) into a small CSV for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx.csv"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_selected",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [m for m in METRICS if m in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + ["%s [%s]" % (m, units[hdr.index(m)]) for m in cols])
        for d in data:
            name = d[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("unnamed>::", "")
            w.writerow([name] + [d[hdr.index(m)] for m in cols])
    print("wrote", out, len(data), "kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
