#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list per kernel:
time share, count, mean duration and (when captured) DRAM bytes per launch and the DRAM bandwidth those imply."""
import collections
import csv
import re
import sys


def to_us(v, unit):
    return v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else (v * 1e6 if unit in ("s", "second") else v))


def to_bytes(v, unit):
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def main(path, top=25):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    tot, cnt, rd, wr = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        m = row["Metric Name"]
        if m == "gpu__time_duration.sum":
            tot[name] += to_us(v, row["Metric Unit"])
            cnt[name] += 1
        elif m == "dram__bytes_read.sum":
            rd[name] += to_bytes(v, row["Metric Unit"])
        elif m == "dram__bytes_write.sum":
            wr[name] += to_bytes(v, row["Metric Unit"])
    T = sum(tot.values())
    have_dram = bool(rd)
    print(f"total {T / 1e3:.2f} ms over {sum(cnt.values())} launches" +
          (f"; DRAM read {sum(rd.values()) / 1e9:.2f} GB, write {sum(wr.values()) / 1e9:.2f} GB" if have_dram else ""))
    print(f"{'share':>6} {'total us':>10} {'n':>5} {'mean us':>9}" + (f" {'rd MB/l':>8} {'wr MB/l':>8} {'GB/s':>6}" if have_dram else "") + "  kernel")
    for k, v in tot.most_common(top):
        extra = ""
        if have_dram:
            extra = f" {rd[k] / cnt[k] / 1e6:8.1f} {wr[k] / cnt[k] / 1e6:8.1f} {(rd[k] + wr[k]) / v / 1e3:6.0f}"
        print(f"{100 * v / T:5.1f}% {v:10.0f} {cnt[k]:5d} {v / cnt[k]:9.1f}{extra}  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
