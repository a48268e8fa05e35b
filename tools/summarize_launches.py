"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        short = re.sub(r"\(.*", "", row["Kernel Name"])[:72]
        tot[short][0] += 1
        tot[short][1] += v
    T = sum(v for _, v in tot.values())
    n = sum(c for c, _ in tot.values())
    print("total %.1f us over %d launches (cold-cache, serialised: compare shares)" % (T, n))
    for k, (c, v) in sorted(tot.items(), key=lambda x: -x[1][1])[:top]:
        print("%-74s n=%4d %9.1f us %5.1f%%" % (k, c, v, 100 * v / T))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
