"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import re
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    by_grid = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = name.replace("void ", "").replace("bde::<unnamed>::", "").replace("bde::", "")[:60]
        agg[name][0] += 1
        agg[name][1] += v
        if "gemm" in name or "attention" in name or "attn" in name or "mlp" in name:
            by_grid[(name, row.get("Grid Size", ""))][0] += 1
            by_grid[(name, row.get("Grid Size", ""))][1] += v
    tot = sum(v[1] for v in agg.values())
    print("%-62s %6s %11s %9s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-62s %6d %11.1f %9.1f %6.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    print("total_us %.1f launches %d" % (tot, sum(v[0] for v in agg.values())))
    print("\nby grid:")
    for k, v in sorted(by_grid.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-50s %-16s n=%5d total=%9.1f avg=%8.1f" % (k[0], k[1], v[0], v[1], v[1] / v[0]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
