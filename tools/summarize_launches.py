#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (time share, count, mean)."""
import collections
import csv
import re
import sys


def main(path, top=25):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"total {T / 1e3:.2f} ms over {sum(cnt.values())} launches")
    print(f"{'share':>6} {'total us':>10} {'n':>5} {'mean us':>9}  kernel")
    for k, v in tot.most_common(top):
        print(f"{100 * v / T:5.1f}% {v:10.0f} {cnt[k]:5d} {v / cnt[k]:9.1f}  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
