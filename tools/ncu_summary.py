"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count/avg/total."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] in ("ns", "nsecond"):
            v /= 1e3
        elif row["Metric Unit"] in ("ms", "msecond"):
            v *= 1e3
        agg.setdefault(row["Kernel Name"], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'total_us':>10} {'share':>6} {'n':>4} {'avg_us':>9}  kernel")
    for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{sum(v):10.1f} {sum(v) / tot:6.1%} {len(v):4d} {sum(v) / len(v):9.2f}  {name[:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
