"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms summed device time")
    print(f"# {'total ms':>10} {'launches':>8} {'avg us':>9} {'share':>7}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {v[1] / 1e6:10.3f} {v[0]:8d} {v[1] / v[0] / 1e3:9.2f} {v[1] / tot * 100:6.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
