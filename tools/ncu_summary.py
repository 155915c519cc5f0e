"""Condense an ncu report (.ncu-rep, `--set full`) into one CSV row per kernel with the metrics DESIGN.md quotes.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rN_ncu_full_xxx.csv
"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(head)}
    cols = [m for m in METRICS if m in idx]
    seen = {}
    for r in data:
        name = r[idx["Kernel Name"]]
        if name not in seen:
            seen[name] = r
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + cols)
        w.writerow([""] + [units[idx[c]] for c in cols])
        for name, r in seen.items():
            w.writerow([name] + [r[idx[c]] for c in cols])
    print(open(out).read())


if __name__ == "__main__":
    main()
