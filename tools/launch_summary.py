"""Aggregates an ncu --csv launch list (gpu__time_duration.sum) by kernel name."""
import collections
import csv
import re
import sys


def main(path, top=28):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += v
        n += 1
    tot = sum(v for _, v in agg.values())
    print(f"{n} launches, {tot / 1e6:.1f} ms total")
    print("| ms | share | launches | avg us | kernel |\n|---|---|---|---|---|")
    for name, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| {v / 1e6:.3f} | {100 * v / tot:.1f}% | {c} | {v / c / 1e3:.1f} | `{name[:70]}` |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 28)
