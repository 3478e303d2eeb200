#!/usr/bin/env python
"""Microbenchmark of the relative-position attention kernels (tf32 tensor-core path) at the bench shapes, forward and backward, both axes.
Prints one JSON line: microseconds per launch and (query, key) pairs per second; 96 flop / pair forward, 240 backward."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cmgan_b200  # noqa: E402,F401
from cmgan_b200 import ops  # noqa: E402
from cmgan_b200.ops import call  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    ops.set_precision("tf32")
    dev = "cuda"
    B, T, F2 = a.batch, 321, 101
    M = B * T * F2
    torch.manual_seed(0)
    nb = 4
    qkvs = [torch.randn(M, 192, device=dev) for _ in range(nb)]
    dctxs = [torch.randn(M, 64, device=dev) for _ in range(nb)]
    E = torch.randn(1025, 16, device=dev) * 0.5
    ctx, lse = torch.empty(M, 64, device=dev), torch.empty(M, 4, device=dev)
    dqkv, delta, dE = torch.empty(M, 192, device=dev), torch.empty(M, 4, device=dev), torch.zeros(1025, 16, device=dev)
    res = {"batch": B, "M": M}
    for axis, name in ((0, "time"), (1, "freq")):
        L, N = (T, B * F2) if axis == 0 else (F2, B * T)
        pairs = N * 4 * L * L

        def fwd(i):
            call("cmgan_attention_fwd_tf32", qkvs[i % nb], E, B, T, F2, axis, ctx, lse)

        def fwd1(i):
            call("cmgan_attention_fwd_tf32_nbuf", qkvs[i % nb], E, B, T, F2, axis, ctx, lse, 1)

        def fwd_tc(i):
            call("cmgan_attention_fwd_tc", qkvs[i % nb], E, B, T, F2, axis, ctx, lse)

        def bwd(i):
            call("cmgan_attention_bwd_tf32", qkvs[i % nb], E, ctx, dctxs[i % nb], lse, B, T, F2, axis, delta, dqkv, dE)
        from cmgan_b200._lib import lib as _lib
        nws = _lib().cdll.cmgan_attention_bwd_ws_floats(B, T, F2, axis)
        ws = torch.empty(nws, device=dev)

        def bwd_ws(i):
            call("cmgan_attention_bwd_tf32_ws", qkvs[i % nb], E, ctx, dctxs[i % nb], lse, B, T, F2, axis, delta, dqkv, dE, 7, ws, nws)

        def dq_only(i):
            call("cmgan_attention_bwd_tf32_ws", qkvs[i % nb], E, ctx, dctxs[i % nb], lse, B, T, F2, axis, delta, dqkv, dE, 2, None, 0)

        def dq_only_ws(i):
            call("cmgan_attention_bwd_tf32_ws", qkvs[i % nb], E, ctx, dctxs[i % nb], lse, B, T, F2, axis, delta, dqkv, dE, 2, ws, nws)
        side = torch.cuda.Stream()

        def bwd2(i):
            args = (qkvs[i % nb], E, ctx, dctxs[i % nb], lse, B, T, F2, axis, delta, dqkv, dE)
            call("cmgan_attention_bwd_tf32_parts", *args, 1)
            ops.call_on(side, "cmgan_attention_bwd_tf32_parts", *args, 4)
            call("cmgan_attention_bwd_tf32_parts", *args, 2)
            ops.join(side)
        ctx2, lse2 = torch.empty_like(ctx), torch.empty_like(lse)
        call("cmgan_attention_fwd_tf32_nbuf", qkvs[0], E, B, T, F2, axis, ctx, lse, 2)
        call("cmgan_attention_fwd_tf32_nbuf", qkvs[0], E, B, T, F2, axis, ctx2, lse2, 1)
        torch.cuda.synchronize()
        assert torch.equal(ctx, ctx2) and torch.equal(lse, lse2), "single- and double-buffered forward disagree"
        for fn, key in ((fwd, "fwd"), (fwd1, "fwd_single_buffer"), (fwd_tc, "fwd_tc"), (bwd, "bwd"), (bwd_ws, "bwd_global_dE"), (dq_only, "dq"), (dq_only_ws, "dq_global_dE"),
                        (bwd2, "bwd_two_streams")):
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.reps):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / a.reps * 1e3
            res[f"{name}_{key}"] = {"us": us, "gpairs_per_s": pairs / us / 1e3, "L": L, "pairs": pairs}
    print(json.dumps(res))
    if a.profile:
        torch.cuda.cudart().cudaProfilerStart()
        call("cmgan_attention_fwd_tc", qkvs[0], E, B, T, F2, 0, ctx, lse)
        call("cmgan_attention_fwd_tf32", qkvs[0], E, B, T, F2, 0, ctx, lse)
        call("cmgan_attention_bwd_tf32", qkvs[0], E, ctx, dctxs[0], lse, B, T, F2, 0, delta, dqkv, dE)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
