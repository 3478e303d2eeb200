#!/usr/bin/env python
"""Microbenchmark of the feed-forward module of the conformer block at the bench shapes: the fused tcgen05 kernel (cmgan_ffn_fwd) against
the three-launch sequence it replaces (cmgan_ln_apply + two cmgan_gemm_rows_f32).  Inputs rotate over NBUF buffer sets (> 126 MB L2 in total)
so that every launch reads its rows from HBM.  Prints one JSON line per configuration.

    python tools/bench_ffn.py [--batch 4] [--reps 40] [--profile]      (--profile: cudaProfilerStart/Stop around a few fused launches, for ncu)
"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cmgan_b200  # noqa: E402
from cmgan_b200 import ops  # noqa: E402
from cmgan_b200.ops import EPI_DROP_RES, EPI_SWISH_DUAL, call, gemm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--reps", type=int, default=40)
    ap.add_argument("--drop", type=float, default=0.2)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--timeline-bwd", action="store_true")
    ap.add_argument("--timeline", action="store_true", help="dump CTA 0's per-warp clock64 stamps of one fused forward launch")
    a = ap.parse_args()
    dev = "cuda"
    ops.set_precision("tf32")
    ops.PACK_CACHE = ops.PackCache()
    M = a.batch * 321 * 101
    nbuf = max(2, int(600e6 // (M * 64 * 4 * 2)) + 1)
    torch.manual_seed(0)
    xs = [torch.randn(M, 64, device=dev) for _ in range(nbuf)]
    outs = [torch.empty(M, 64, device=dev) for _ in range(nbuf)]
    g, b = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1
    W1, b1 = torch.randn(256, 64, device=dev) / 8, torch.randn(256, device=dev) * 0.1
    W2, b2 = torch.randn(64, 256, device=dev) / 16, torch.randn(64, device=dev) * 0.1
    thr, inv = ops.drop_params(a.drop)
    W1p, W2p = ops.packed_weight(W1, 0, 1, 64, 64, 1, 256), ops.packed_weight(W2, 0, 1, 256, 256, 1, 64)

    def fused(i):
        call("cmgan_ffn_fwd", xs[i], 64, M, g, b, W1p, b1, W2p, b2, 0.5, 11, 12, thr, inv, None, outs[i], 64)

    xn, st = torch.empty(M, 64, device=dev), torch.empty(M, 2, device=dev)
    h, act = torch.empty(M, 256, device=dev), torch.empty(M, 256, device=dev)

    def unfused(i, keep):
        call("cmgan_ln_apply", xs[i], 64, M, g, b, None, 0, xn, 64, st, 1)
        gemm(A=xn, lda=64, W=W1, sb_k=1, sb_n=64, bias=b1, C=h if keep else None, ldc=256, M=M, N=256, Cin=64, epi=EPI_SWISH_DUAL, C2=act, ldc2=256, seed=11,
             drop_p=a.drop)
        gemm(A=act, lda=256, W=W2, sb_k=1, sb_n=256, bias=b2, C=outs[i], ldc=64, M=M, N=64, Cin=256, epi=EPI_DROP_RES, alpha=0.5, R=xs[i], ldr=64, seed=12,
             drop_p=a.drop)

    # ---- backward: fused kernel vs the three launches it replaces (data gradients only; the weight-gradient GEMMs are common to both)
    douts = [torch.randn(M, 64, device=dev) for _ in range(nbuf)]
    dzs = [torch.randn(M, 64, device=dev) for _ in range(nbuf)]
    dxs = [torch.empty(M, 64, device=dev) for _ in range(nbuf)]
    a_o, dh_o, xn_o = torch.empty(M, 256, device=dev), torch.empty(M, 256, device=dev), torch.empty(M, 64, device=dev)
    dg, db = torch.zeros(64, device=dev), torch.zeros(64, device=dev)
    W2tp, W1tp = ops.packed_weight(W2, 0, 256, 1, 64, 1, 256), ops.packed_weight(W1, 0, 64, 1, 256, 1, 64)

    def fused_bwd(i):
        call("cmgan_ffn_bwd", xs[i], 64, dzs[i], 64, douts[i], 64, None, 0, M, g, b, W1p, b1, W2tp, W1tp, 11, thr, inv, None, dxs[i], 64, a_o, dh_o, xn_o, dg, db)

    dln = torch.empty(M, 64, device=dev)

    def unfused_bwd(i):
        gemm(A=dzs[i], lda=64, W=W2, sb_k=256, sb_n=1, C=dh_o, ldc=256, M=M, N=256, Cin=64, epi=ops.EPI_DSWISH_DROP, aux=h, ldaux=256, seed=11, drop_p=a.drop)
        gemm(A=dh_o, lda=256, W=W1, sb_k=64, sb_n=1, C=dln, ldc=64, M=M, N=64, Cin=256)
        call("cmgan_ln_bwd", dln, 64, xs[i], 64, st, g, M, douts[i], 64, None, 0, dxs[i], 64, dg, db)

    def timeit(fn):
        for i in range(nbuf):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(a.reps):
            fn(r % nbuf)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.reps * 1e3      # us

    fused(0)
    ref = outs[0].clone()
    unfused(0, False)
    torch.cuda.synchronize()
    dev_err = (outs[0] - ref).abs().max().item()
    flop = 2.0 * M * 64 * 256 * 2
    byts = 4.0 * M * 64 * 2
    res = {"M": M, "batch": a.batch, "nbuf": nbuf, "fused_vs_unfused_max_abs": dev_err}
    for name, fn in (("fused", fused), ("unfused_eval", lambda i: unfused(i, False)), ("unfused_train", lambda i: unfused(i, True))):
        us = timeit(fn)
        res[name] = {"us": us, "tflops": flop / us / 1e6, "compulsory_gbs": byts / us / 1e3}
    unfused(0, True)
    bflop = 2.0 * M * 64 * 256 * 3
    bbytes = 4.0 * M * (64 * 5 + 256 * 2)
    for name, fn in (("fused_bwd", fused_bwd), ("unfused_bwd", unfused_bwd)):
        us = timeit(fn)
        res[name] = {"us": us, "tflops": bflop / us / 1e6, "io_gbs": bbytes / us / 1e3}
    print(json.dumps(res))
    if a.timeline_bwd:
        a.timeline = True
    if a.timeline:
        buf = torch.zeros(32 * 64, dtype=torch.int64, device=dev)
        from cmgan_b200._lib import lib
        lib().cdll.cmgan_ffn_debug_timeline(ctypes.c_void_p(buf.data_ptr()))
        (fused_bwd if a.timeline_bwd else fused)(1)
        torch.cuda.synchronize()
        lib().cdll.cmgan_ffn_debug_timeline(None)
        t = buf.view(32, 64).cpu()
        t0 = int(t[t > 0].min())
        for w in range(32):
            row = [int(v) - t0 if v > 0 else -1 for v in t[w].tolist()]
            if any(v >= 0 for v in row):
                print("warp", w, row, file=sys.stderr)
    if a.profile:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        for i in range(2):
            fused(i % nbuf)
            fused_bwd(i % nbuf)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
