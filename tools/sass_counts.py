#!/usr/bin/env python
"""Count the SASS mnemonics that prove which hardware path each kernel uses (cuobjdump -sass over the in-tree library):
UTCHMMA = tcgen05.mma (kind::tf32), LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor loads / stores, UBLKCP = cp.async.bulk,
HMMA = legacy mma.sync, LDGSTS = cp.async, SYNCS = mbarrier.  Writes a markdown table (profiles/r2_sass_counts.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
LIB = os.path.join(ROOT, "cmgan_b200", "libcmgan_b200.so")
WANT = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDGSTS", "SYNCS", "FFMA", "MUFU", "LDL", "STL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name)
            kern = name
            counts[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and kern:
            op = m.group(1)
            counts[kern][op] += 1
            total[kern] += 1
    lines = ["# SASS mnemonic counts per kernel (cuobjdump -sass cmgan_b200/libcmgan_b200.so, sm_100a)", "",
             "Static instruction counts, not executed counts.  `UTCHMMA` = tcgen05.mma kind::tf32, `LDTM` = tcgen05.ld, `UTMALDG` / `UTMASTG` = TMA tensor",
             "load / store, `UBLKCP` = cp.async.bulk, `HMMA` = mma.sync (legacy tensor path), `LDGSTS` = cp.async, `SYNCS` = mbarrier ops,",
             "`LDL` / `STL` = local-memory (spill) traffic.", "",
             "| kernel | instr | " + " | ".join(WANT) + " |", "|---|---:|" + "---:|" * len(WANT)]
    for k, c in counts.items():
        if total[k] < 40:
            continue
        lines.append(f"| `{k}` | {total[k]} | " + " | ".join(str(c.get(w, 0)) for w in WANT) + " |")
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_counts.md")
    with open(dst, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print(dst, len(counts), "kernels")


if __name__ == "__main__":
    main()
