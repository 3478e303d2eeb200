#!/usr/bin/env python
"""Summarise a CMGAN_PROBE_DUMP file (per-launch GEMM timings of one instrumented bench step) by shape."""
import collections
import json
import sys

d = json.load(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
agg = collections.defaultdict(lambda: [0, 0.0])
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for rec in d:
    n, M, N, K, us = rec[:5]
    nb = rec[5] if len(rec) > 5 else 4.0 * M * (N + K)
    cnt = rec[6] if len(rec) > 6 else 1          # newer dumps: one record per (shape, bytes) class with the mean replay time
    key = (n.replace("cmgan_gemm_", "").replace("_f32", ""), M, N, K)
    agg[key][0] += cnt
    agg[key][1] += us * cnt
    agg[key][2] += nb * cnt
tot = sum(v[1] for v in agg.values())
print(f"total GEMM time {tot / 1e3:.2f} ms over {len(d)} calls")
for (n, M, N, K), (c, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    fl = 2.0 * M * N * K * c
    print(f"{n:6s} M={M:7d} N={N:4d} K={K:5d} x{c:3d} {us:9.0f} us  {us / c:7.1f} us/call  {fl / us / 1e6:7.1f} TFLOP/s  {by / us / 1e3:7.1f} GB/s (algorithmic)")
