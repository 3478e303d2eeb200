"""TSCNet generator: the reference's nn.Module interface over the hand-written CUDA path.

``TSCNet(num_channel=64, num_features=201)`` keeps the reference's constructor, ``forward`` signature
(x: (B, 2, T, F), any strides -> (final_real, final_imag), each (B, 1, T, F)) and the 359-key state dict
(ref: generator.py:159-196), so ``load_state_dict(torch.load("best_ckpt/ckpt"))`` and the call sites in
train.py:100 / evaluation.py:40 work unchanged.  Nothing in ``forward``/``backward`` is a PyTorch compute
op: activations live channel-last as rows (b, t, f) x channels, every contraction is a ``cmgan_gemm_*``
call and everything else one of the HBM-bound kernels in csrc/.  PyTorch provides memory, streams and the
autograd graph node (one ``torch.autograd.Function`` for the whole network).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from .ops import (EPI_ACC, EPI_DBNSWISH, EPI_DROP_RES, EPI_DSWISH_DROP, EPI_NONE, PRO_BN_SWISH, PRO_DROP, PRO_IN_PRELU, PRO_LN,
                  PRO_SWISH_DROP, call, gemm)

C = 64          # num_channel (the kernels are specialised for 64 channels = 4 heads x 16)
CAT = 5 * C     # width of a dense-block concat buffer: [out4 | out3 | out2 | out1 | x]
FF_DROP = 0.2   # ref: generator.py:82,89
ATT_DROP = 0.2  # ref: generator.py:81,88


def _empty(*shape, dev, dtype=torch.float32):
    return torch.empty(*shape, dtype=dtype, device=dev)


class _Tabs:
    """scale/shift/mean/rstd per (group, channel) and PReLU slope per channel of a normalisation site
    (or of all five 64-channel slots of a dense-block concat buffer)."""

    def __init__(self, G, width, dev, identity=False):
        self.scale = _empty(G, width, dev=dev)
        self.shift = _empty(G, width, dev=dev)
        self.mean = _empty(G, width, dev=dev)
        self.rstd = _empty(G, width, dev=dev)
        self.slope = _empty(width, dev=dev)
        self.width = width
        if identity:    # slot 4 of a decoder concat buffer holds final activations: act(x) = x
            call("cmgan_fill", self.scale, G * width, 1.0)
            call("cmgan_fill", self.shift, G * width, 0.0)
            call("cmgan_fill", self.slope, width, 1.0)


class _Sums:
    """One zero-initialised double scratch per pass, sliced per statistics site (a single memset)."""

    def __init__(self, n, dev):
        self.buf = torch.zeros(n, dtype=torch.float64, device=dev)
        self.off = 0

    def take(self, n):
        assert self.off + n <= self.buf.numel(), "statistics scratch exhausted"
        o = self.off
        self.off += n
        return (self.buf, o)


def _inst_norm_site(x, ldx, G, rows, Cn, gamma, beta, tabs: _Tabs, c0, slope_w, sums: _Sums):
    """InstanceNorm2d statistics (ref: generator.py:35) -> tables at channel offset c0 of ``tabs``"""
    s = sums.take(G * Cn * 2)
    call("cmgan_norm_stats", x, ldx, G, rows, Cn, s)
    call("cmgan_norm_finalize", s, rows, G, Cn, 0, gamma, beta, None, None, 0.0, (tabs.scale, c0), (tabs.shift, c0), (tabs.mean, c0),
         (tabs.rstd, c0), tabs.width)
    if slope_w is not None:
        call("cmgan_copy_rows", slope_w, Cn, (tabs.slope, c0), Cn, 1, Cn)


def _norm_bwd(x, ldx, dact, ldd, G, rows, Cn, act, batch_stats, tabs: _Tabs, c0, slope, dx, lddx, dgamma, dbeta, dslope, sums: _Sums):
    s = sums.take(G * Cn * 2)
    args = ((tabs.scale, c0), (tabs.shift, c0), (tabs.mean, c0), (tabs.rstd, c0), tabs.width, slope)
    call("cmgan_norm_bwd_reduce", x, ldx, dact, ldd, G, rows, Cn, act, *args, s, dslope)
    call("cmgan_norm_bwd_apply", x, ldx, dact, ldd, G, rows, Cn, act, 1 if batch_stats else 0, *args, s, dx, lddx, dgamma, dbeta)


def _site_seed(seed: int, block_id: int, site: int) -> int:
    return (seed * 1000003 + block_id * 16 + site + 1) & 0xFFFFFFFFFFFFFFFF


# ====================================================================================== conformer block
def conformer_fwd(x, P, p, B, T, F2, axis, training, seed, block_id, sums: _Sums, save: Optional[dict]):
    """ConformerBlock + the outer TSCB residual (ref: conformer.py:216-222, generator.py:95,97).
    x: (M, 64) rows of the (B, T, F2) grid; axis 0 = sequences along T, 1 = along F2.  Returns LN(x4) + x."""
    dev = x.device
    M = x.shape[0]
    dp = FF_DROP if training else 0.0
    da = ATT_DROP if training else 0.0
    sd = [_site_seed(seed, block_id, i) for i in range(5)]

    def ff(xin, name, s1, s2):
        st = _empty(M, 2, dev=dev)
        call("cmgan_ln_stats", xin, C, M, st)
        h = _empty(M, 4 * C, dev=dev)
        gemm(A=xin, lda=C, W=P[f"{p}.{name}.fn.fn.net.0.weight"], sb_k=1, sb_n=C, bias=P[f"{p}.{name}.fn.fn.net.0.bias"], C=h, ldc=4 * C, M=M,
             N=4 * C, Cin=C, pro=PRO_LN, p0=st, p1=P[f"{p}.{name}.fn.norm.weight"], p2=P[f"{p}.{name}.fn.norm.bias"])
        out = _empty(M, C, dev=dev)
        gemm(A=h, lda=4 * C, W=P[f"{p}.{name}.fn.fn.net.3.weight"], sb_k=1, sb_n=4 * C, bias=P[f"{p}.{name}.fn.fn.net.3.bias"], C=out, ldc=C, M=M,
             N=C, Cin=4 * C, pro=PRO_SWISH_DROP, pro_seed=s1, pro_drop_p=dp, epi=EPI_DROP_RES, alpha=0.5, R=xin, ldr=C, seed=s2, drop_p=dp)
        return st, h, out

    st1, h1, x1 = ff(x, "ff1", sd[0], sd[1])
    # ---- attention (ref: conformer.py:90-133)
    st2 = _empty(M, 2, dev=dev)
    call("cmgan_ln_stats", x1, C, M, st2)
    qkv = _empty(M, 3 * C, dev=dev)
    lnw, lnb = P[f"{p}.attn.norm.weight"], P[f"{p}.attn.norm.bias"]
    gemm(A=x1, lda=C, W=P[f"{p}.attn.fn.to_q.weight"], sb_k=1, sb_n=C, C=qkv, ldc=3 * C, M=M, N=C, Cin=C, pro=PRO_LN, p0=st2, p1=lnw, p2=lnb)
    gemm(A=x1, lda=C, W=P[f"{p}.attn.fn.to_kv.weight"], sb_k=1, sb_n=C, C=(qkv, C), ldc=3 * C, M=M, N=2 * C, Cin=C, pro=PRO_LN, p0=st2, p1=lnw,
         p2=lnb)
    ctx = _empty(M, C, dev=dev)
    lse = _empty(M, 4, dev=dev)
    call("cmgan_attention_fwd", qkv, P[f"{p}.attn.fn.rel_pos_emb.weight"], B, T, F2, axis, ctx, lse)
    x2 = _empty(M, C, dev=dev)
    gemm(A=ctx, lda=C, W=P[f"{p}.attn.fn.to_out.weight"], sb_k=1, sb_n=C, bias=P[f"{p}.attn.fn.to_out.bias"], C=x2, ldc=C, M=M, N=C, Cin=C,
         epi=EPI_DROP_RES, alpha=1.0, R=x1, ldr=C, seed=sd[2], drop_p=da)
    # ---- convolution module (ref: conformer.py:160-173)
    st3 = _empty(M, 2, dev=dev)
    call("cmgan_ln_stats", x2, C, M, st3)
    g = _empty(M, 4 * C, dev=dev)
    gemm(A=x2, lda=C, W=P[f"{p}.conv.net.2.weight"], sb_k=1, sb_n=C, bias=P[f"{p}.conv.net.2.bias"], C=g, ldc=4 * C, M=M, N=4 * C, Cin=C,
         pro=PRO_LN, p0=st3, p1=P[f"{p}.conv.net.0.weight"], p2=P[f"{p}.conv.net.0.bias"])
    d = _empty(M, 2 * C, dev=dev)
    call("cmgan_glu_dwconv_fwd", g, P[f"{p}.conv.net.4.conv.weight"], P[f"{p}.conv.net.4.conv.bias"], B, T, F2, axis, d)
    bn = _Tabs(1, 2 * C, dev)
    bnp = (P[f"{p}.conv.net.5.weight"], P[f"{p}.conv.net.5.bias"], P[f"{p}.conv.net.5.running_mean"], P[f"{p}.conv.net.5.running_var"])
    if training:
        s = sums.take(2 * C * 2)
        call("cmgan_norm_stats", d, 2 * C, 1, M, 2 * C, s)
        call("cmgan_norm_finalize", s, M, 1, 2 * C, 0, *bnp, 0.1, bn.scale, bn.shift, bn.mean, bn.rstd, 2 * C)
    else:
        call("cmgan_norm_finalize", None, M, 1, 2 * C, 1, *bnp, 0.1, bn.scale, bn.shift, bn.mean, bn.rstd, 2 * C)
    x3 = _empty(M, C, dev=dev)
    gemm(A=d, lda=2 * C, W=P[f"{p}.conv.net.7.weight"], sb_k=1, sb_n=2 * C, bias=P[f"{p}.conv.net.7.bias"], C=x3, ldc=C, M=M, N=C, Cin=2 * C,
         pro=PRO_BN_SWISH, p0=bn.scale, p1=bn.shift, epi=EPI_DROP_RES, alpha=1.0, R=x2, ldr=C)
    # ---- second feed-forward, post norm, outer residual
    st4, h2, x4 = ff(x3, "ff2", sd[3], sd[4])
    st5 = _empty(M, 2, dev=dev)
    y = _empty(M, C, dev=dev)
    call("cmgan_ln_apply", x4, C, M, P[f"{p}.post_norm.weight"], P[f"{p}.post_norm.bias"], x, C, y, C, st5)
    if save is not None:
        save.update(x=x, st1=st1, h1=h1, x1=x1, st2=st2, qkv=qkv, ctx=ctx, lse=lse, x2=x2, st3=st3, g=g, d=d, bn=bn, x3=x3, st4=st4, h2=h2,
                    x4=x4, st5=st5, sd=sd, dp=dp, da=da, axis=axis, training=training, p=p)
    return y


def conformer_bwd(dy, S: dict, P, G: Dict[str, torch.Tensor], B, T, F2, sums: _Sums):
    """Gradient of conformer_fwd: dy (M, 64) -> dx (M, 64); parameter gradients accumulate into G[name]."""
    dev = dy.device
    M = dy.shape[0]
    p, axis, dp, da, sd = S["p"], S["axis"], S["dp"], S["da"], S["sd"]

    def ff_bwd(dout, xin, st, h, name, s1, s2, res2=None):
        # out = xin + 0.5 * drop2(W2 (swish(h) * drop1) + b2),  h = W1 LN(xin) + b1
        W1, W2 = P[f"{p}.{name}.fn.fn.net.0.weight"], P[f"{p}.{name}.fn.fn.net.3.weight"]
        dh = _empty(M, 4 * C, dev=dev)
        gemm(A=dout, lda=C, W=W2, sb_k=4 * C, sb_n=1, C=dh, ldc=4 * C, M=M, N=4 * C, Cin=C, pro=PRO_DROP, pro_alpha=0.5, pro_seed=s2,
             pro_drop_p=dp, epi=EPI_DSWISH_DROP, aux=h, ldaux=4 * C, seed=s1, drop_p=dp)
        gemm(wgrad=True, A=h, lda=4 * C, Cin=4 * C, pro=PRO_SWISH_DROP, pro_seed=s1, pro_drop_p=dp, D=dout, ldd=C, N=C, prod=1, alpha=0.5, seed=s2,
             drop_p=dp, W=None, C=G[f"{p}.{name}.fn.fn.net.3.weight"], sb_k=1, sb_n=4 * C, ldc=0, M=M, dbias=G[f"{p}.{name}.fn.fn.net.3.bias"])
        dln = _empty(M, C, dev=dev)
        gemm(A=dh, lda=4 * C, W=W1, sb_k=C, sb_n=1, C=dln, ldc=C, M=M, N=C, Cin=4 * C)
        gemm(wgrad=True, A=xin, lda=C, Cin=C, pro=PRO_LN, p0=st, p1=P[f"{p}.{name}.fn.norm.weight"], p2=P[f"{p}.{name}.fn.norm.bias"], D=dh,
             ldd=4 * C, N=4 * C, W=None, C=G[f"{p}.{name}.fn.fn.net.0.weight"], sb_k=1, sb_n=C, ldc=0, M=M, dbias=G[f"{p}.{name}.fn.fn.net.0.bias"])
        dxin = _empty(M, C, dev=dev)
        call("cmgan_ln_bwd", dln, C, xin, C, st, P[f"{p}.{name}.fn.norm.weight"], M, dout, C, res2, C, dxin, C, G[f"{p}.{name}.fn.norm.weight"],
             G[f"{p}.{name}.fn.norm.bias"])
        return dxin

    # y = LN(x4) * g + b + x
    dx4 = _empty(M, C, dev=dev)
    call("cmgan_ln_bwd", dy, C, S["x4"], C, S["st5"], P[f"{p}.post_norm.weight"], M, None, 0, None, 0, dx4, C, G[f"{p}.post_norm.weight"],
         G[f"{p}.post_norm.bias"])
    dx3 = ff_bwd(dx4, S["x3"], S["st4"], S["h2"], "ff2", sd[3], sd[4])
    # ---- convolution module: x3 = x2 + W7 swish(bn(d)) + b7
    bn = S["bn"]
    dbn = _empty(M, 2 * C, dev=dev)
    gemm(A=dx3, lda=C, W=P[f"{p}.conv.net.7.weight"], sb_k=2 * C, sb_n=1, C=dbn, ldc=2 * C, M=M, N=2 * C, Cin=C, epi=EPI_DBNSWISH, aux=S["d"],
         ldaux=2 * C, e0=bn.scale, e1=bn.shift)
    gemm(wgrad=True, A=S["d"], lda=2 * C, Cin=2 * C, pro=PRO_BN_SWISH, p0=bn.scale, p1=bn.shift, D=dx3, ldd=C, N=C, W=None,
         C=G[f"{p}.conv.net.7.weight"], sb_k=1, sb_n=2 * C, ldc=0, M=M, dbias=G[f"{p}.conv.net.7.bias"])
    dd = _empty(M, 2 * C, dev=dev)
    _norm_bwd(S["d"], 2 * C, dbn, 2 * C, 1, M, 2 * C, 0, S["training"], bn, 0, None, dd, 2 * C, G[f"{p}.conv.net.5.weight"],
              G[f"{p}.conv.net.5.bias"], None, sums)
    dg = _empty(M, 4 * C, dev=dev)
    call("cmgan_glu_dwconv_bwd", S["g"], dd, P[f"{p}.conv.net.4.conv.weight"], B, T, F2, axis, dg, G[f"{p}.conv.net.4.conv.weight"],
         G[f"{p}.conv.net.4.conv.bias"])
    dln3 = _empty(M, C, dev=dev)
    gemm(A=dg, lda=4 * C, W=P[f"{p}.conv.net.2.weight"], sb_k=C, sb_n=1, C=dln3, ldc=C, M=M, N=C, Cin=4 * C)
    gemm(wgrad=True, A=S["x2"], lda=C, Cin=C, pro=PRO_LN, p0=S["st3"], p1=P[f"{p}.conv.net.0.weight"], p2=P[f"{p}.conv.net.0.bias"], D=dg,
         ldd=4 * C, N=4 * C, W=None, C=G[f"{p}.conv.net.2.weight"], sb_k=1, sb_n=C, ldc=0, M=M, dbias=G[f"{p}.conv.net.2.bias"])
    dx2 = _empty(M, C, dev=dev)
    call("cmgan_ln_bwd", dln3, C, S["x2"], C, S["st3"], P[f"{p}.conv.net.0.weight"], M, dx3, C, None, 0, dx2, C, G[f"{p}.conv.net.0.weight"],
         G[f"{p}.conv.net.0.bias"])
    # ---- attention: x2 = x1 + drop(ctx Wo^T + bo)
    dctx = _empty(M, C, dev=dev)
    gemm(A=dx2, lda=C, W=P[f"{p}.attn.fn.to_out.weight"], sb_k=C, sb_n=1, C=dctx, ldc=C, M=M, N=C, Cin=C, pro=PRO_DROP, pro_alpha=1.0,
         pro_seed=sd[2], pro_drop_p=da)
    gemm(wgrad=True, A=S["ctx"], lda=C, Cin=C, D=dx2, ldd=C, N=C, prod=1, alpha=1.0, seed=sd[2], drop_p=da, W=None,
         C=G[f"{p}.attn.fn.to_out.weight"], sb_k=1, sb_n=C, ldc=0, M=M, dbias=G[f"{p}.attn.fn.to_out.bias"])
    dqkv = _empty(M, 3 * C, dev=dev)
    delta = _empty(M, 4, dev=dev)
    call("cmgan_attention_bwd", S["qkv"], P[f"{p}.attn.fn.rel_pos_emb.weight"], S["ctx"], dctx, S["lse"], B, T, F2, axis, delta, dqkv,
         G[f"{p}.attn.fn.rel_pos_emb.weight"])
    dln2 = _empty(M, C, dev=dev)
    gemm(A=dqkv, lda=3 * C, W=P[f"{p}.attn.fn.to_q.weight"], sb_k=C, sb_n=1, C=dln2, ldc=C, M=M, N=C, Cin=C)
    gemm(A=(dqkv, C), lda=3 * C, W=P[f"{p}.attn.fn.to_kv.weight"], sb_k=C, sb_n=1, C=dln2, ldc=C, M=M, N=C, Cin=2 * C, epi=EPI_ACC, alpha=1.0)
    lnw, lnb = P[f"{p}.attn.norm.weight"], P[f"{p}.attn.norm.bias"]
    gemm(wgrad=True, A=S["x1"], lda=C, Cin=C, pro=PRO_LN, p0=S["st2"], p1=lnw, p2=lnb, D=dqkv, ldd=3 * C, N=C, W=None,
         C=G[f"{p}.attn.fn.to_q.weight"], sb_k=1, sb_n=C, ldc=0, M=M)
    gemm(wgrad=True, A=S["x1"], lda=C, Cin=C, pro=PRO_LN, p0=S["st2"], p1=lnw, p2=lnb, D=(dqkv, C), ldd=3 * C, N=2 * C, W=None,
         C=G[f"{p}.attn.fn.to_kv.weight"], sb_k=1, sb_n=C, ldc=0, M=M)
    dx1 = _empty(M, C, dev=dev)
    call("cmgan_ln_bwd", dln2, C, S["x1"], C, S["st2"], lnw, M, dx2, C, None, 0, dx1, C, G[f"{p}.attn.norm.weight"], G[f"{p}.attn.norm.bias"])
    # ---- first feed-forward; the outer residual adds dy
    return ff_bwd(dx1, S["x"], S["st1"], S["h1"], "ff1", sd[0], sd[1], res2=dy)


# ====================================================================================== dense blocks
def _dense_taps(dil):
    return [((kh - 1) * dil, kw - 1) for kh in range(2) for kw in range(3)]     # tap = kh*3 + kw  (ref: generator.py:12-13,17,21)


def dense_block_fwd(cat, tabs: _Tabs, P, p, B, T, Fw, sums: _Sums):
    """DilatedDenseNet (ref: generator.py:39-47) on a concat buffer whose slot 4 (channels 256..319) already holds the block
    input (raw values + the tables that turn them into activations).  Layer i writes its raw conv output into slot 4 - i."""
    M, rows = B * T * Fw, T * Fw
    for i in range(1, 5):
        dil, c0, Cin, co = 2 ** (i - 1), (5 - i) * C, C * i, (4 - i) * C
        gemm(A=(cat, c0), lda=CAT, W=P[f"{p}.conv{i}.weight"], sb_tap=1, sb_k=6, sb_n=Cin * 6, bias=P[f"{p}.conv{i}.bias"], C=(cat, co), ldc=CAT,
             M=M, N=C, Cin=Cin, taps=_dense_taps(dil), conv=dict(OH=T, OW=Fw, IH=T, IW=Fw), pro=PRO_IN_PRELU, p0=(tabs.scale, c0),
             p1=(tabs.shift, c0), p2=(tabs.slope, c0), rows_per_batch=rows, pstride=CAT)
        _inst_norm_site((cat, co), CAT, B, rows, C, P[f"{p}.norm{i}.weight"], P[f"{p}.norm{i}.bias"], tabs, co, P[f"{p}.prelu{i}.weight"], sums)


def dense_block_bwd(cat, dcat, tabs: _Tabs, P, G, p, B, T, Fw, sums: _Sums):
    """dcat slot 0 holds the gradient wrt act(out4); on return dcat slot 4 holds the gradient wrt the block input's activations."""
    dev = cat.device
    M, rows = B * T * Fw, T * Fw
    draw = _empty(M, C, dev=dev)
    for i in range(4, 0, -1):
        dil, c0, Cin, co = 2 ** (i - 1), (5 - i) * C, C * i, (4 - i) * C
        _norm_bwd((cat, co), CAT, (dcat, co), CAT, B, rows, C, 1, True, tabs, co, P[f"{p}.prelu{i}.weight"], draw, C, G[f"{p}.norm{i}.weight"],
                  G[f"{p}.norm{i}.bias"], G[f"{p}.prelu{i}.weight"], sums)
        taps = _dense_taps(dil)
        gemm(wgrad=True, A=(cat, c0), lda=CAT, Cin=Cin, taps=taps, conv=dict(OH=T, OW=Fw, IH=T, IW=Fw), pro=PRO_IN_PRELU, p0=(tabs.scale, c0),
             p1=(tabs.shift, c0), p2=(tabs.slope, c0), rows_per_batch=rows, pstride=CAT, D=draw, ldd=C, N=C, W=None, C=G[f"{p}.conv{i}.weight"],
             sb_tap=1, sb_k=6, sb_n=Cin * 6, ldc=0, M=M, dbias=G[f"{p}.conv{i}.bias"])
        gemm(A=draw, lda=C, W=P[f"{p}.conv{i}.weight"], sb_tap=1, sb_k=Cin * 6, sb_n=6, C=(dcat, c0), ldc=CAT, M=M, N=Cin, Cin=C,
             taps=[(-dy, -dx) for dy, dx in taps], conv=dict(OH=T, OW=Fw, IH=T, IW=Fw), epi=EPI_NONE if i == 4 else EPI_ACC, alpha=1.0)


# ====================================================================================== whole network
_W3 = [(0, -1), (0, 0), (0, 1)]


def tscnet_fwd(x, P, training: bool, seed: int, save: Optional[dict]):
    """TSCNet.forward (ref: generator.py:174-196).  x (B, 2, T, F) any strides -> final_real, final_imag (B, 1, T, F)."""
    dev = x.device
    B, two, T, F = x.shape
    assert two == 2 and F % 2 == 1, "expected x of shape (B, 2, T, F) with odd F"
    F2 = (F - 1) // 2 + 1
    M, M2 = B * T * F, B * T * F2
    xs = x.stride()
    sums = _Sums((16 * C + 2) * B * 2 + 8 * 2 * C * 2 + 64, dev)
    # ---- dense encoder (ref: generator.py:50-69)
    catE = _empty(M, CAT, dev=dev)
    tabE = _Tabs(B, CAT, dev)
    pe = "dense_encoder"
    call("cmgan_head_conv", x, xs[0], xs[1], xs[2], xs[3], B, T, F, P[pe + ".conv_1.0.weight"], P[pe + ".conv_1.0.bias"], (catE, 4 * C), CAT)
    _inst_norm_site((catE, 4 * C), CAT, B, T * F, C, P[pe + ".conv_1.1.weight"], P[pe + ".conv_1.1.bias"], tabE, 4 * C, P[pe + ".conv_1.2.weight"], sums)
    dense_block_fwd(catE, tabE, P, pe + ".dilated_dense", B, T, F, sums)
    e2 = _empty(M2, C, dev=dev)
    gemm(A=catE, lda=CAT, W=P[pe + ".conv_2.0.weight"], sb_tap=1, sb_k=3, sb_n=3 * C, bias=P[pe + ".conv_2.0.bias"], C=e2, ldc=C, M=M2, N=C, Cin=C,
         taps=_W3, conv=dict(OH=T, OW=F2, IH=T, IW=F, mul_x=2), pro=PRO_IN_PRELU, p0=tabE.scale, p1=tabE.shift, p2=tabE.slope,
         rows_per_batch=T * F, pstride=CAT)
    tab2 = _Tabs(B, C, dev)
    _inst_norm_site(e2, C, B, T * F2, C, P[pe + ".conv_2.1.weight"], P[pe + ".conv_2.1.bias"], tab2, 0, P[pe + ".conv_2.2.weight"], sums)
    h = _empty(M2, C, dev=dev)
    call("cmgan_norm_apply", e2, C, B, T * F2, C, 1, tab2.scale, tab2.shift, C, tab2.slope, h, C)
    # ---- 4 x TSCB (ref: generator.py:92-99)
    conf_saves = []
    for i in range(1, 5):
        for axis, name in ((0, "time_conformer"), (1, "freq_conformer")):
            sv = {} if save is not None else None
            h = conformer_fwd(h, P, f"TSCB_{i}.{name}", B, T, F2, axis, training, seed, (i - 1) * 2 + axis, sums, sv)
            conf_saves.append(sv)
    # ---- decoders (ref: generator.py:122-156)
    dec = {}
    for pd in ("mask_decoder", "complex_decoder"):
        cat = _empty(M2, CAT, dev=dev)
        tabs = _Tabs(B, CAT, dev, identity=True)
        call("cmgan_copy_rows", h, C, (cat, 4 * C), CAT, M2, C)
        dense_block_fwd(cat, tabs, P, pd + ".dense_block", B, T, F2, sums)
        sp = _empty(M2, 2 * C, dev=dev)      # == (B, T, 2*F2, 64): the sub-pixel shuffle is a free reinterpretation
        gemm(A=cat, lda=CAT, W=P[pd + ".sub_pixel.conv.weight"], sb_tap=1, sb_k=3, sb_n=3 * C, bias=P[pd + ".sub_pixel.conv.bias"], C=sp, ldc=2 * C,
             M=M2, N=2 * C, Cin=C, taps=_W3, conv=dict(OH=T, OW=F2, IH=T, IW=F2), pro=PRO_IN_PRELU, p0=tabs.scale, p1=tabs.shift, p2=tabs.slope,
             rows_per_batch=T * F2, pstride=CAT)
        dec[pd] = dict(cat=cat, tabs=tabs, sp=sp)
    pm, pc = "mask_decoder", "complex_decoder"
    m1 = _empty(M, dev=dev)
    call("cmgan_rowdot_fwd", dec[pm]["sp"], B, T, F, 1, None, None, None, P[pm + ".conv_1.weight"], P[pm + ".conv_1.bias"], m1)
    tabM = _Tabs(B, 1, dev)
    _inst_norm_site(m1, 1, B, T * F, 1, P[pm + ".norm.weight"], P[pm + ".norm.bias"], tabM, 0, None, sums)
    tabC = _Tabs(B, C, dev)
    _inst_norm_site(dec[pc]["sp"], C, B, T * 2 * F2, C, P[pc + ".norm.weight"], P[pc + ".norm.bias"], tabC, 0, P[pc + ".prelu.weight"], sums)
    cplx = _empty(M, 2, dev=dev)
    call("cmgan_rowdot_fwd", dec[pc]["sp"], B, T, F, 2, tabC.scale, tabC.shift, tabC.slope, P[pc + ".conv.weight"], P[pc + ".conv.bias"], cplx)
    fr = _empty(B, 1, T, F, dev=dev)
    fi = _empty(B, 1, T, F, dev=dev)
    call("cmgan_recombine", m1, tabM.scale, tabM.shift, P[pm + ".prelu.weight"], P[pm + ".final_conv.weight"], P[pm + ".final_conv.bias"],
         P[pm + ".prelu_out.weight"], x, xs[0], xs[1], xs[2], xs[3], cplx, B, T, F, fr, fi)
    if save is not None:
        save.update(x=x, B=B, T=T, F=F, F2=F2, catE=catE, tabE=tabE, e2=e2, tab2=tab2, conf=conf_saves, dec=dec, m1=m1, tabM=tabM, tabC=tabC)
    return fr, fi


def tscnet_bwd(S: dict, dfr, dfi, P, G: Dict[str, torch.Tensor]):
    """Backward of tscnet_fwd.  dfr / dfi: gradients wrt final_real / final_imag ((B,1,T,F), any strides, or None).
    Parameter gradients are accumulated (+=) into the tensors of G."""
    x = S["x"]
    dev = x.device
    B, T, F, F2 = S["B"], S["T"], S["F"], S["F2"]
    M, M2 = B * T * F, B * T * F2
    xs = x.stride()
    sums = _Sums((16 * C + 2) * B * 2 + 8 * 2 * C * 2 + 64, dev)
    if dfr is None:
        dfr = torch.zeros(B, 1, T, F, device=dev)
    if dfi is None:
        dfi = torch.zeros(B, 1, T, F, device=dev)
    if dfi.stride() != dfr.stride():
        dfi = dfi.contiguous()
        dfr = dfr.contiguous()
    gs = dfr.stride()
    pm, pc = "mask_decoder", "complex_decoder"
    dec, tabM, tabC = S["dec"], S["tabM"], S["tabC"]
    dcplx = _empty(M, 2, dev=dev)
    dz = _empty(M, dev=dev)
    call("cmgan_recombine_bwd", S["m1"], tabM.scale, tabM.shift, P[pm + ".prelu.weight"], P[pm + ".final_conv.weight"], P[pm + ".final_conv.bias"],
         P[pm + ".prelu_out.weight"], x, xs[0], xs[1], xs[2], xs[3], dfr, dfi, gs[0], gs[2], gs[3], B, T, F, dcplx, dz, G[pm + ".prelu_out.weight"],
         G[pm + ".final_conv.weight"], G[pm + ".final_conv.bias"])
    dm1 = _empty(M, dev=dev)
    _norm_bwd(S["m1"], 1, dz, 1, B, T * F, 1, 1, True, tabM, 0, P[pm + ".prelu.weight"], dm1, 1, G[pm + ".norm.weight"], G[pm + ".norm.bias"],
              G[pm + ".prelu.weight"], sums)
    dsp = {}
    dsp[pm] = _empty(M2, 2 * C, dev=dev)
    call("cmgan_rowdot_bwd", dec[pm]["sp"], B, T, F, 1, None, None, None, P[pm + ".conv_1.weight"], dm1, dsp[pm], G[pm + ".conv_1.weight"],
         G[pm + ".conv_1.bias"])
    dactc = _empty(M2, 2 * C, dev=dev)
    call("cmgan_rowdot_bwd", dec[pc]["sp"], B, T, F, 2, tabC.scale, tabC.shift, tabC.slope, P[pc + ".conv.weight"], dcplx, dactc,
         G[pc + ".conv.weight"], G[pc + ".conv.bias"])
    dsp[pc] = _empty(M2, 2 * C, dev=dev)
    _norm_bwd(dec[pc]["sp"], C, dactc, C, B, T * 2 * F2, C, 1, True, tabC, 0, P[pc + ".prelu.weight"], dsp[pc], C, G[pc + ".norm.weight"],
              G[pc + ".norm.bias"], G[pc + ".prelu.weight"], sums)
    dh = None
    for pd in (pm, pc):
        cat, tabs = dec[pd]["cat"], dec[pd]["tabs"]
        dcat = _empty(M2, CAT, dev=dev)
        gemm(wgrad=True, A=cat, lda=CAT, Cin=C, taps=_W3, conv=dict(OH=T, OW=F2, IH=T, IW=F2), pro=PRO_IN_PRELU, p0=tabs.scale, p1=tabs.shift,
             p2=tabs.slope, rows_per_batch=T * F2, pstride=CAT, D=dsp[pd], ldd=2 * C, N=2 * C, W=None, C=G[pd + ".sub_pixel.conv.weight"], sb_tap=1,
             sb_k=3, sb_n=3 * C, ldc=0, M=M2, dbias=G[pd + ".sub_pixel.conv.bias"])
        gemm(A=dsp[pd], lda=2 * C, W=P[pd + ".sub_pixel.conv.weight"], sb_tap=1, sb_k=3 * C, sb_n=3, C=dcat, ldc=CAT, M=M2, N=C, Cin=2 * C,
             taps=[(0, 1), (0, 0), (0, -1)], conv=dict(OH=T, OW=F2, IH=T, IW=F2))
        dense_block_bwd(cat, dcat, tabs, P, G, pd + ".dense_block", B, T, F2, sums)
        if dh is None:
            dh = _empty(M2, C, dev=dev)
            call("cmgan_copy_rows", (dcat, 4 * C), CAT, dh, C, M2, C)
        else:
            call("cmgan_add_rows", (dcat, 4 * C), CAT, dh, C, M2, C)
    # ---- TSCBs in reverse
    k = 7
    for i in range(4, 0, -1):
        for axis in (1, 0):
            dh = conformer_bwd(dh, S["conf"][k], P, G, B, T, F2, sums)
            k -= 1
    # ---- encoder
    pe = "dense_encoder"
    catE, tabE, tab2 = S["catE"], S["tabE"], S["tab2"]
    de2 = _empty(M2, C, dev=dev)
    _norm_bwd(S["e2"], C, dh, C, B, T * F2, C, 1, True, tab2, 0, P[pe + ".conv_2.2.weight"], de2, C, G[pe + ".conv_2.1.weight"],
              G[pe + ".conv_2.1.bias"], G[pe + ".conv_2.2.weight"], sums)
    gemm(wgrad=True, A=catE, lda=CAT, Cin=C, taps=_W3, conv=dict(OH=T, OW=F2, IH=T, IW=F, mul_x=2), pro=PRO_IN_PRELU, p0=tabE.scale, p1=tabE.shift,
         p2=tabE.slope, rows_per_batch=T * F, pstride=CAT, D=de2, ldd=C, N=C, W=None, C=G[pe + ".conv_2.0.weight"], sb_tap=1, sb_k=3, sb_n=3 * C,
         ldc=0, M=M2, dbias=G[pe + ".conv_2.0.bias"])
    dcatE = _empty(M, CAT, dev=dev)
    gemm(A=de2, lda=C, W=P[pe + ".conv_2.0.weight"], sb_tap=1, sb_k=3 * C, sb_n=3, C=dcatE, ldc=CAT, M=M, N=C, Cin=C,
         taps=[(0, 1), (0, 0), (0, -1)], conv=dict(OH=T, OW=F, IH=T, IW=F2, div_x=2))
    dense_block_bwd(catE, dcatE, tabE, P, G, pe + ".dilated_dense", B, T, F, sums)
    draw1 = _empty(M, C, dev=dev)
    _norm_bwd((catE, 4 * C), CAT, (dcatE, 4 * C), CAT, B, T * F, C, 1, True, tabE, 4 * C, P[pe + ".conv_1.2.weight"], draw1, C,
              G[pe + ".conv_1.1.weight"], G[pe + ".conv_1.1.bias"], G[pe + ".conv_1.2.weight"], sums)
    call("cmgan_head_conv_wgrad", x, xs[0], xs[1], xs[2], xs[3], B, T, F, draw1, C, G[pe + ".conv_1.0.weight"], G[pe + ".conv_1.0.bias"])


# ====================================================================================== parameters / module
def _param_specs(num_channel: int, num_features: int):
    """[(key, shape, kind, fan_in)] in the reference's state-dict order (ref: generator.py:159-172)."""
    c = num_channel
    specs = []

    def conv(key, cout, cin, kh, kw, bias=True):
        specs.append((key + ".weight", (cout, cin, kh, kw), "kaiming", cin * kh * kw))
        if bias:
            specs.append((key + ".bias", (cout,), "bias", cin * kh * kw))

    def affine(key, n):
        specs.append((key + ".weight", (n,), "ones", 0))
        specs.append((key + ".bias", (n,), "zeros", 0))

    def prelu(key, n, init=0.25):
        specs.append((key + ".weight", (n,), ("const", init), 0))

    def dense(p):
        for i in range(1, 5):
            conv(f"{p}.conv{i}", c, c * i, 2, 3)
            affine(f"{p}.norm{i}", c)
            prelu(f"{p}.prelu{i}", c)

    def linear(key, nout, nin, bias=True):
        specs.append((key + ".weight", (nout, nin), "kaiming", nin))
        if bias:
            specs.append((key + ".bias", (nout,), "bias", nin))

    conv("dense_encoder.conv_1.0", c, 3, 1, 1)
    affine("dense_encoder.conv_1.1", c)
    prelu("dense_encoder.conv_1.2", c)
    dense("dense_encoder.dilated_dense")
    conv("dense_encoder.conv_2.0", c, c, 1, 3)
    affine("dense_encoder.conv_2.1", c)
    prelu("dense_encoder.conv_2.2", c)
    for i in range(1, 5):
        for name in ("time_conformer", "freq_conformer"):
            p = f"TSCB_{i}.{name}"
            for ff in ("ff1",):
                linear(f"{p}.{ff}.fn.fn.net.0", 4 * c, c)
                linear(f"{p}.{ff}.fn.fn.net.3", c, 4 * c)
                affine(f"{p}.{ff}.fn.norm", c)
            linear(f"{p}.attn.fn.to_q", c, c, bias=False)
            linear(f"{p}.attn.fn.to_kv", 2 * c, c, bias=False)
            linear(f"{p}.attn.fn.to_out", c, c)
            specs.append((f"{p}.attn.fn.rel_pos_emb.weight", (1025, c // 4), "normal", 0))
            affine(f"{p}.attn.norm", c)
            affine(f"{p}.conv.net.0", c)
            specs.append((f"{p}.conv.net.2.weight", (4 * c, c, 1), "kaiming", c))
            specs.append((f"{p}.conv.net.2.bias", (4 * c,), "bias", c))
            specs.append((f"{p}.conv.net.4.conv.weight", (2 * c, 1, 31), "kaiming", 31))
            specs.append((f"{p}.conv.net.4.conv.bias", (2 * c,), "bias", 31))
            affine(f"{p}.conv.net.5", 2 * c)
            specs.append((f"{p}.conv.net.5.running_mean", (2 * c,), "buf_zeros", 0))
            specs.append((f"{p}.conv.net.5.running_var", (2 * c,), "buf_ones", 0))
            specs.append((f"{p}.conv.net.5.num_batches_tracked", (), "buf_long", 0))
            specs.append((f"{p}.conv.net.7.weight", (c, 2 * c, 1), "kaiming", 2 * c))
            specs.append((f"{p}.conv.net.7.bias", (c,), "bias", 2 * c))
            linear(f"{p}.ff2.fn.fn.net.0", 4 * c, c)
            linear(f"{p}.ff2.fn.fn.net.3", c, 4 * c)
            affine(f"{p}.ff2.fn.norm", c)
            affine(f"{p}.post_norm", c)
    dense("mask_decoder.dense_block")
    conv("mask_decoder.sub_pixel.conv", 2 * c, c, 1, 3)
    conv("mask_decoder.conv_1", 1, c, 1, 2)
    affine("mask_decoder.norm", 1)
    prelu("mask_decoder.prelu", 1)
    conv("mask_decoder.final_conv", 1, 1, 1, 1)
    prelu("mask_decoder.prelu_out", num_features, -0.25)
    dense("complex_decoder.dense_block")
    conv("complex_decoder.sub_pixel.conv", 2 * c, c, 1, 3)
    prelu("complex_decoder.prelu", c)
    affine("complex_decoder.norm", c)
    conv("complex_decoder.conv", 2, c, 1, 2)
    return specs


def _init_tensor(shape, kind, fan_in):
    if kind == "kaiming":       # nn.Conv*/nn.Linear default: kaiming_uniform_(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
        b = 1.0 / math.sqrt(fan_in)
        return torch.empty(shape).uniform_(-b, b)
    if kind == "bias":
        b = 1.0 / math.sqrt(fan_in)
        return torch.empty(shape).uniform_(-b, b)
    if kind == "ones" or kind == "buf_ones":
        return torch.ones(shape)
    if kind == "zeros" or kind == "buf_zeros":
        return torch.zeros(shape)
    if kind == "normal":
        return torch.randn(shape)
    if kind == "buf_long":
        return torch.zeros(shape, dtype=torch.long)
    if isinstance(kind, tuple) and kind[0] == "const":
        return torch.full(shape, float(kind[1]))
    raise ValueError(kind)


class _Holder(nn.Module):
    """Name-space node of the parameter tree (no compute)."""


def _register(root: nn.Module, key: str, tensor: torch.Tensor, is_buffer: bool):
    parts = key.split(".")
    mod = root
    for part in parts[:-1]:
        if not hasattr(mod, part):
            mod.add_module(part, _Holder())
        mod = getattr(mod, part)
    if is_buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor))


class _TSCNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, module, training, seed, *params):
        P = module._tensor_dict()
        save = {} if any(ctx.needs_input_grad) else None
        fr, fi = tscnet_fwd(x, P, training, seed, save)
        ctx.module = module
        ctx.saved = save
        return fr, fi

    @staticmethod
    def backward(ctx, dfr, dfi):
        module = ctx.module
        S = ctx.saved
        if S is None:
            raise RuntimeError("TSCNet backward called but the forward pass did not record state")
        P = module._tensor_dict()
        G, ret = module._grad_targets()
        tscnet_bwd(S, dfr, dfi, P, G)
        ctx.saved = None
        return (None, None, None, None, *ret)


class TSCNet(nn.Module):
    """Drop-in for reference ``models.generator.TSCNet`` (generator.py:159-196)."""

    def __init__(self, num_channel: int = 64, num_features: int = 201):
        super().__init__()
        if num_channel != C:
            raise ValueError("the sm_100a kernels are specialised for num_channel=64 (the reference's only configuration)")
        self.num_channel, self.num_features = num_channel, num_features
        self._keys: List[str] = []
        for key, shape, kind, fan_in in _param_specs(num_channel, num_features):
            is_buf = isinstance(kind, str) and kind.startswith("buf_")
            _register(self, key, _init_tensor(shape, kind, fan_in), is_buf)
            self._keys.append(key)
        self._param_keys = [k for k, _ in self.named_parameters()]
        self.seed = 0
        self._step = 0
        self.flat_grad: Optional[torch.Tensor] = None    # set by enable_flat_grads()
        self._flat_views: Optional[Dict[str, torch.Tensor]] = None

    # -- plumbing ---------------------------------------------------------------------------------
    def _tensor_dict(self) -> Dict[str, torch.Tensor]:
        d = dict(self.named_parameters())
        d.update(dict(self.named_buffers()))
        return d

    def enable_flat_grads(self) -> torch.Tensor:
        """Allocate one flat fp32 gradient buffer, make every ``param.grad`` a view of it, and let the backward kernels accumulate
        straight into it (one NCCL all-reduce per step, no per-parameter copies).  Returns the flat buffer; zero it once per step."""
        params = list(self.named_parameters())
        sizes = [((p.numel() + 3) // 4) * 4 for _, p in params]      # 16-byte aligned segments
        flat = torch.zeros(sum(sizes), device=params[0][1].device)
        views, off = {}, 0
        for (k, p), n in zip(params, sizes):
            v = flat[off:off + p.numel()].view_as(p)
            p.grad = v
            views[k] = v
            off += n
        self.flat_grad, self._flat_views = flat, views
        return flat

    def _grad_targets(self):
        """(name -> tensor the kernels accumulate into, tuple returned to autograd for *params)"""
        if self._flat_views is not None:
            return self._flat_views, tuple(None for _ in self._param_keys)
        named = dict(self.named_parameters())
        G = {k: torch.zeros_like(named[k]) for k in self._param_keys}
        return G, tuple(G[k] if named[k].requires_grad else None for k in self._param_keys)

    def forward(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("cmgan_b200.TSCNet runs on CUDA only (no CPU fallback)")
        if x.dtype != torch.float32:
            raise RuntimeError("cmgan_b200.TSCNet expects float32 input")
        if self.training:
            self._step += 1
            torch._foreach_add_([b for k, b in self.named_buffers() if k.endswith("num_batches_tracked")], 1)   # bookkeeping only
        params = [p for _, p in self.named_parameters()]
        return _TSCNetFn.apply(x, self, self.training, self.seed * 7919 + self._step, *params)
