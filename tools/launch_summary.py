"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel: share of the step.

    python tools/launch_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/rN_ncu_launch_summary.txt
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    tot = defaultdict(float); cnt = defaultdict(int)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = r["Kernel Name"].split("(")[0]
        tot[name] += v; cnt[name] += 1
    total = sum(tot.values())
    print(f"ncu --metrics gpu__time_duration.sum --clock-control none: {sys.argv[2] if len(sys.argv) > 2 else ''}")
    print("(cold-cache, serialised launches: compare SHARES with bench.py roofline.kernel_share_of_step, not absolutes)")
    for name in sorted(tot, key=tot.get, reverse=True):
        print(f"{100 * tot[name] / total:6.2f}%  n={cnt[name]:4d}  total={tot[name]:9.3f} ms  avg={1e3 * tot[name] / cnt[name]:9.1f} us  {name[:150]}")


if __name__ == "__main__":
    main()
