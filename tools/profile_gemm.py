#!/usr/bin/env python
"""Isolated launches of the hot GEMM shapes for `ncu --set full` captures (tf32 tcgen05 path)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from cmgan_b200 import ops  # noqa: E402
from cmgan_b200.ops import call, gemm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--which", default="ffn1", choices=["ffn1", "ffn1_dual", "ffn2", "conv4", "wgrad_ffn1", "wgrad_conv4", "plain256", "plain64", "plain64k256"])
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
ops.set_precision("tf32")
dev = "cuda"
M = 129684
torch.manual_seed(0)
if args.which == "ffn1":
    x, W, b = torch.randn(M, 64, device=dev), torch.randn(256, 64, device=dev), torch.randn(256, device=dev)
    g, be, st = torch.randn(64, device=dev), torch.randn(64, device=dev), torch.empty(M, 2, device=dev)
    call("cmgan_ln_stats", x, 64, M, st)
    out = torch.empty(M, 256, device=dev)
    fn = lambda: gemm(A=x, lda=64, W=W, sb_k=1, sb_n=64, bias=b, C=out, ldc=256, M=M, N=256, Cin=64, pro=ops.PRO_LN, p0=st, p1=g, p2=be)
elif args.which in ("plain256", "plain64"):
    N = 256 if args.which == "plain256" else 64
    x, W, b = torch.randn(M, 64, device=dev), torch.randn(N, 64, device=dev), torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev)
    fn = lambda: gemm(A=x, lda=64, W=W, sb_k=1, sb_n=64, bias=b, C=out, ldc=N, M=M, N=N, Cin=64)
elif args.which == "plain64k256":
    x, W, b = torch.randn(M, 256, device=dev), torch.randn(64, 256, device=dev), torch.randn(64, device=dev)
    out = torch.empty(M, 64, device=dev)
    fn = lambda: gemm(A=x, lda=256, W=W, sb_k=1, sb_n=256, bias=b, C=out, ldc=64, M=M, N=64, Cin=256)
elif args.which == "ffn1_dual":
    x, W, b = torch.randn(M, 64, device=dev), torch.randn(256, 64, device=dev), torch.randn(256, device=dev)
    h, a = torch.empty(M, 256, device=dev), torch.empty(M, 256, device=dev)
    fn = lambda: gemm(A=x, lda=64, W=W, sb_k=1, sb_n=64, bias=b, C=h, ldc=256, M=M, N=256, Cin=64, epi=ops.EPI_SWISH_DUAL, C2=a, ldc2=256, seed=5,
                      drop_p=0.2)
elif args.which == "ffn2":
    h, W, b, x = torch.randn(M, 256, device=dev), torch.randn(64, 256, device=dev), torch.randn(64, device=dev), torch.randn(M, 64, device=dev)
    out = torch.empty(M, 64, device=dev)
    fn = lambda: gemm(A=h, lda=256, W=W, sb_k=1, sb_n=256, bias=b, C=out, ldc=64, M=M, N=64, Cin=256, pro=ops.PRO_SWISH_DROP, pro_seed=1,
                      pro_drop_p=0.2, epi=ops.EPI_DROP_RES, alpha=0.5, R=x, ldr=64, seed=2, drop_p=0.2)
elif args.which == "conv4":
    B, T, F = 4, 321, 101
    cat, W, b = torch.randn(M, 320, device=dev), torch.randn(64, 256, 2, 3, device=dev), torch.randn(64, device=dev)
    out = torch.empty(M, 64, device=dev)
    taps = [((kh - 1) * 8, kw - 1) for kh in range(2) for kw in range(3)]
    fn = lambda: gemm(A=(cat, 64), lda=320, W=W, sb_tap=1, sb_k=6, sb_n=1536, bias=b, C=out, ldc=64, M=M, N=64, Cin=256, taps=taps,
                      conv=dict(OH=T, OW=F, IH=T, IW=F))
elif args.which == "wgrad_ffn1":
    x, dh = torch.randn(M, 64, device=dev), torch.randn(M, 256, device=dev)
    g, be, st = torch.randn(64, device=dev), torch.randn(64, device=dev), torch.empty(M, 2, device=dev)
    call("cmgan_ln_stats", x, 64, M, st)
    dw, db = torch.zeros(256, 64, device=dev), torch.zeros(256, device=dev)
    fn = lambda: gemm(wgrad=True, W=None, C=dw, ldc=0, dbias=db, A=x, lda=64, Cin=64, pro=ops.PRO_LN, p0=st, p1=g, p2=be, D=dh, ldd=256, N=256,
                      sb_k=1, sb_n=64, M=M)
else:
    B, T, F = 4, 321, 101
    cat, dy = torch.randn(M, 320, device=dev), torch.randn(M, 64, device=dev)
    dw, db = torch.zeros(64, 256, 2, 3, device=dev), torch.zeros(64, device=dev)
    taps = [((kh - 1) * 8, kw - 1) for kh in range(2) for kw in range(3)]
    fn = lambda: gemm(wgrad=True, W=None, C=dw, ldc=0, dbias=db, A=(cat, 64), lda=320, Cin=256, taps=taps, conv=dict(OH=T, OW=F, IH=T, IW=F),
                      D=dy, ldd=64, N=64, sb_tap=1, sb_k=6, sb_n=1536, M=M)
for _ in range(2):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    fn()
e1.record()
torch.cuda.synchronize()
print(f"{args.which}: {e0.elapsed_time(e1) / args.reps * 1e3:.1f} us per call")
