#!/usr/bin/env python
"""Markdown excerpt of `ncu --set full` reports: per captured launch the duration, DRAM bytes (read / write) and bandwidth, tensor-pipe
and issue activity, occupancy limiters and the top warp-stall reasons.  Usage: ncu_excerpt.py report.ncu-rep [...] > excerpt.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "us", lambda v: v / 1e3),
    ("dram__bytes_read.sum", "MB rd", lambda v: v / 1e6),
    ("dram__bytes_write.sum", "MB wr", lambda v: v / 1e6),
    ("sm__inst_executed_pipe_tensor.sum", "tensor inst", lambda v: v),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma %", lambda v: v),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", lambda v: v),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %", lambda v: v),
    ("launch__registers_per_thread", "regs", lambda v: v),
    ("launch__grid_size", "grid", lambda v: v),
    ("launch__block_size", "block", lambda v: v),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts", lambda v: v),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def stall_summary(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    res, hdr, agg, name = {}, None, None, None
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Kernel Name":
            if name is not None and agg:
                res.setdefault(name, agg)
            name, hdr, agg = r[1], None, {}
            continue
        if name is None:
            continue
        if hdr is None:
            hdr = r
            cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        for i in cols:
            try:
                agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
            except (ValueError, IndexError):
                pass
    if name is not None and agg:
        res.setdefault(name, agg)
    return res


def main():
    for rep in sys.argv[1:]:
        hdr, units, rows = raw_rows(rep)
        H = {h: i for i, h in enumerate(hdr)}
        stalls = stall_summary(rep)
        print(f"### `{rep.split('/')[-1]}`\n")
        have = [(m, lab, f) for m, lab, f in METRICS if m in H]
        print("| kernel | " + " | ".join(lab for _, lab, _ in have) + " | DRAM GB/s | top stalls |")
        print("|---|" + "---:|" * (len(have) + 1) + "---|")
        seen = {}
        for r in rows:
            name = r[H["Kernel Name"]]
            short = name.split("(")[0].replace("<unnamed>::", "").replace("void ", "")
            vals = {}
            for m, lab, f in have:
                try:
                    v = float(r[H[m]].replace(",", "")) * UNIT.get(units[H[m]], 1.0)
                except ValueError:
                    v = float("nan")
                vals[m] = v
            key = (short, r[H["launch__grid_size"]] if "launch__grid_size" in H else "")
            seen[key] = seen.get(key, 0) + 1
            if seen[key] > 2:
                continue
            t_ns = vals.get("gpu__time_duration.sum", float("nan"))
            bw = (vals.get("dram__bytes_read.sum", 0) + vals.get("dram__bytes_write.sum", 0)) / t_ns if t_ns == t_ns and t_ns > 0 else float("nan")
            st = stalls.get(name, {})
            tot = sum(st.values()) or 1
            top = ", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:4])
            cells = []
            for m, lab, f in have:
                v = f(vals[m])
                cells.append(f"{v:.1f}" if abs(v) < 1e4 and v != int(v) else f"{v:.0f}")
            print(f"| `{short}` | " + " | ".join(cells) + f" | {bw:.0f} | {top} |")
        print()


if __name__ == "__main__":
    main()
