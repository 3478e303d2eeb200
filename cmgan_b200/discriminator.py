"""Metric discriminator: the reference's nn.Module interface over the CUDA kernels.

``Discriminator(ndf, in_channel=2)`` keeps the reference's constructor, ``forward(x, y)`` (two (B, 1, F, T) magnitude
spectrograms -> (B, 1) in (0, 1)) and state-dict keys including the legacy ``torch.nn.utils.spectral_norm`` triplets
``weight_orig / weight_u / weight_v`` (ref: discriminator.py:29-64, utils.py:42-50).  Per training forward every
spectrally-normalised weight gets one power iteration (u, v updated in place) exactly like the reference's hook.

Layout: channel-last rows (b, h, w) with h = frequency, w = time.  The 4x4 stride-2 convolutions are implicit GEMMs
(16 taps), InstanceNorm + PReLU is materialised between them, AdaptiveMaxPool2d(1) is fused with the last norm/PReLU.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .conformer_block import _Sums, _Tabs, _empty, _inst_norm_site, _norm_bwd
from .generator import _Holder, _register
from .ops import call, gemm

_TAPS = [(kh - 1, kw - 1) for kh in range(4) for kw in range(4)]          # tap = kh*4 + kw, padding 1
_TAPS_T = [(-dy, -dx) for dy, dx in _TAPS]
_CONV_IDX = (0, 3, 6, 9)
DROP_P = 0.3


def _spectral(P, key, training, dev):
    """-> (w_sn, (sigma, uv)): W / sigma with one power iteration in training (the u / v buffers are updated in place, as by the
    reference's hook); ``uv`` = [u | v] as used by THIS forward (torch's spectral_norm clones them per training forward, so the
    backward of a forward is consistent with its own sigma even when a later forward has iterated the buffers again)"""
    w = P[key + ".weight_orig"]
    R = w.shape[0]
    Cc = w.numel() // R
    w_sn = _empty(*w.shape, dev=dev)
    sigma = _empty(1, dev=dev)
    uv = _empty(R + Cc, dev=dev)
    call("cmgan_spectral_norm", w, R, Cc, P[key + ".weight_u"], P[key + ".weight_v"], 1 if training else 0, w_sn, sigma, uv)
    return w_sn, (sigma, uv)


def disc_fwd(x, y, P, training: bool, seed: int, save):
    cache, ops.PACK_CACHE = ops.PACK_CACHE, None      # W / sigma is a fresh tensor every forward: nothing to cache
    try:
        return _disc_fwd(x, y, P, training, seed, save)
    finally:
        ops.PACK_CACHE = cache


def _disc_fwd(x, y, P, training: bool, seed: int, save):
    dev = x.device
    B, one, H, W = x.shape
    assert one == 1 and y.shape == x.shape
    sums = _Sums((16 + 32 + 64 + 128) * B * 2 * 2 + 64, dev)
    xy = _empty(B * H * W, 2, dev=dev)
    xs, ys = x.stride(), y.stride()
    call("cmgan_stack2", x, xs[0], xs[2], xs[3], y, ys[0], ys[2], ys[3], B, H, W, xy)
    act, Cin, ih, iw = xy, 2, H, W
    layers = []
    pooled = arg = None
    for li, idx in enumerate(_CONV_IDX):
        key = f"layers.{idx}"
        w_sn, sigma = _spectral(P, key, training, dev)
        Cout = w_sn.shape[0]
        oh, ow = (ih + 2 - 4) // 2 + 1, (iw + 2 - 4) // 2 + 1
        M = B * oh * ow
        raw = _empty(M, Cout, dev=dev)
        conv = dict(OH=oh, OW=ow, IH=ih, IW=iw, mul_y=2, mul_x=2)
        gemm(A=act, lda=Cin, W=w_sn, sb_tap=1, sb_k=16, sb_n=Cin * 16, C=raw, ldc=Cout, M=M, N=Cout, Cin=Cin, taps=_TAPS, conv=conv)
        tab = _Tabs(B, Cout, dev)
        _inst_norm_site(raw, Cout, B, oh * ow, Cout, P[f"layers.{idx + 1}.weight"], P[f"layers.{idx + 1}.bias"], tab, 0, None, sums)
        slope = P[f"layers.{idx + 2}.weight"]
        layers.append(dict(a_in=act, Cin=Cin, ih=ih, iw=iw, oh=oh, ow=ow, Cout=Cout, raw=raw, tab=tab, w_sn=w_sn, sigma=sigma, key=key, idx=idx))
        if li < 3:
            nxt = _empty(M, Cout, dev=dev)
            call("cmgan_norm_apply", raw, Cout, B, oh * ow, Cout, 1 | (16 if ops.PRECISION == 1 else 0), tab.scale, tab.shift, Cout, slope, nxt, Cout)
            act, Cin, ih, iw = nxt, Cout, oh, ow
        else:
            pooled = _empty(B, Cout, dev=dev)
            arg = torch.empty(B, Cout, dtype=torch.int32, device=dev)
            call("cmgan_norm_maxpool", raw, B, oh * ow, Cout, tab.scale, tab.shift, slope, pooled, arg)
    # ---- SN Linear(128 -> 64) + Dropout(0.3) + PReLU(64) + SN Linear(64 -> 1) + LearnableSigmoid
    w14, s14 = _spectral(P, "layers.14", training, dev)
    n1, n0 = w14.shape
    h1 = _empty(B, n1, dev=dev)
    gemm(A=pooled, lda=n0, W=w14, sb_k=1, sb_n=n0, bias=P["layers.14.bias"], C=h1, ldc=n1, M=B, N=n1, Cin=n0, precision=0)
    a1 = _empty(B, n1, dev=dev)
    thr, inv = ops.drop_params(DROP_P if training else 0.0)
    call("cmgan_drop_prelu", h1, B * n1, n1, P["layers.16.weight"], seed, thr, inv, a1, ops.SEED_DEV)
    w17, s17 = _spectral(P, "layers.17", training, dev)
    h2 = _empty(B, 1, dev=dev)
    gemm(A=a1, lda=n1, W=w17, sb_k=1, sb_n=n1, bias=P["layers.17.bias"], C=h2, ldc=1, M=B, N=1, Cin=n1, precision=0)
    out = _empty(B, 1, dev=dev)
    call("cmgan_lsigmoid", h2, B, P["layers.18.slope"], out)
    if save is not None:
        save.update(x_shape=(B, H, W), layers=layers, pooled=pooled, arg=arg, w14=w14, s14=s14, h1=h1, a1=a1, w17=w17, s17=s17, h2=h2, out=out,
                    seed=seed, thr=thr, inv=inv)
    return out


def _sn_bwd(P, G, key, w_sn, dw_sn, sig_uv):
    w = P[key + ".weight_orig"]
    R = w.shape[0]
    sigma, uv = sig_uv
    call("cmgan_spectral_norm_bwd", w_sn, dw_sn, R, w.numel() // R, uv, (uv, R), sigma, G[key + ".weight_orig"])


class _Sink(dict):
    """gradient target that discards everything (generator step: the discriminator's own gradients are not needed)"""

    def __init__(self, P):
        super().__init__()
        self._scratch = {}
        self._P = P

    def __missing__(self, k):
        t = torch.empty_like(self._P[k])
        self[k] = t
        return t


def disc_bwd(S, dout, P, G, need_dx: bool, need_dy: bool):
    side, ops.WGRAD_STREAM = ops.WGRAD_STREAM, None      # the spectral-norm backward consumes each dW right away: keep these launches in order
    cache, ops.PACK_CACHE = ops.PACK_CACHE, None
    try:
        return _disc_bwd(S, dout, P, G, need_dx, need_dy)
    finally:
        ops.WGRAD_STREAM = side
        ops.PACK_CACHE = cache


def _disc_bwd(S, dout, P, G, need_dx: bool, need_dy: bool):
    dev = dout.device
    need_w = G is not None         # generator step: only the input gradient is wanted -- no weight gradients, no spectral-norm backward
    if G is None:
        G = _Sink(P)
    B, H, W = S["x_shape"]
    sums = _Sums((16 + 32 + 64 + 128) * B * 2 * 2 + 64, dev)
    dout = dout.contiguous()
    n1, n0 = S["w14"].shape
    dh2 = _empty(B, 1, dev=dev)
    call("cmgan_lsigmoid_bwd", S["h2"], S["out"], dout, B, P["layers.18.slope"], dh2, G["layers.18.slope"])
    if need_w:
        dw17 = torch.zeros_like(S["w17"])
        gemm(wgrad=True, A=S["a1"], lda=n1, Cin=n1, D=dh2, ldd=1, N=1, W=None, C=dw17, sb_k=1, sb_n=n1, ldc=0, M=B, dbias=G["layers.17.bias"],
             precision=0)
        _sn_bwd(P, G, "layers.17", S["w17"], dw17, S["s17"])
    da1 = _empty(B, n1, dev=dev)
    gemm(A=dh2, lda=1, W=S["w17"], sb_k=n1, sb_n=1, C=da1, ldc=n1, M=B, N=n1, Cin=1, precision=0)
    dh1 = _empty(B, n1, dev=dev)
    call("cmgan_drop_prelu_bwd", S["h1"], da1, B * n1, n1, P["layers.16.weight"], S["seed"], S["thr"], S["inv"], dh1, G["layers.16.weight"], ops.SEED_DEV)
    if need_w:
        dw14 = torch.zeros_like(S["w14"])
        gemm(wgrad=True, A=S["pooled"], lda=n0, Cin=n0, D=dh1, ldd=n1, N=n1, W=None, C=dw14, sb_k=1, sb_n=n0, ldc=0, M=B, dbias=G["layers.14.bias"],
             precision=0)
        _sn_bwd(P, G, "layers.14", S["w14"], dw14, S["s14"])
    dpool = _empty(B, n0, dev=dev)
    gemm(A=dh1, lda=n1, W=S["w14"], sb_k=n0, sb_n=1, C=dpool, ldc=n0, M=B, N=n0, Cin=n1, precision=0)
    # ---- conv stack in reverse
    dact = None
    for li in range(3, -1, -1):
        L = S["layers"][li]
        idx, Cout, Cin, oh, ow, ih, iw = L["idx"], L["Cout"], L["Cin"], L["oh"], L["ow"], L["ih"], L["iw"]
        M = B * oh * ow
        if li == 3:
            dact = _empty(M, Cout, dev=dev)
            call("cmgan_maxpool_bwd", dpool, S["arg"], B, oh * ow, Cout, dact)
        draw = _empty(M, Cout, dev=dev)
        _norm_bwd(L["raw"], Cout, dact, Cout, B, oh * ow, Cout, 1, True, L["tab"], 0, P[f"layers.{idx + 2}.weight"], draw, Cout,
                  G[f"layers.{idx + 1}.weight"], G[f"layers.{idx + 1}.bias"], G[f"layers.{idx + 2}.weight"], sums, operand=True)
        conv = dict(OH=oh, OW=ow, IH=ih, IW=iw, mul_y=2, mul_x=2)
        if need_w:
            dw_sn = torch.zeros_like(L["w_sn"])
            gemm(wgrad=True, A=L["a_in"], lda=Cin, Cin=Cin, taps=_TAPS, conv=conv, D=draw, ldd=Cout, N=Cout, W=None, C=dw_sn, sb_tap=1, sb_k=16,
                 sb_n=Cin * 16, ldc=0, M=M)
            _sn_bwd(P, G, L["key"], L["w_sn"], dw_sn, L["sigma"])
        if li > 0 or need_dx or need_dy:
            Min = B * ih * iw
            dact = _empty(Min, Cin, dev=dev)
            gemm(A=draw, lda=Cout, W=L["w_sn"], sb_tap=1, sb_k=Cin * 16, sb_n=16, C=dact, ldc=Cin, M=Min, N=Cin, Cin=Cout, taps=_TAPS_T,
                 conv=dict(OH=ih, OW=iw, IH=oh, IW=ow, div_y=2, div_x=2))
    dx = dy = None
    if need_dx or need_dy:
        dx = _empty(B, 1, H, W, dev=dev) if need_dx else None
        dy = _empty(B, 1, H, W, dev=dev) if need_dy else None
        call("cmgan_unstack2", dact, B * H * W, dx, dy)
    return dx, dy


def _d_specs(ndf: int, in_channel: int):
    specs = []
    cin = in_channel
    for i, idx in enumerate(_CONV_IDX):
        cout = ndf * (2 ** i)
        specs.append((f"layers.{idx}.weight_orig", (cout, cin, 4, 4), "kaiming", cin * 16, False))
        specs.append((f"layers.{idx}.weight_u", (cout,), "unit", 0, True))
        specs.append((f"layers.{idx}.weight_v", (cin * 16,), "unit", 0, True))
        specs.append((f"layers.{idx + 1}.weight", (cout,), "ones", 0, False))
        specs.append((f"layers.{idx + 1}.bias", (cout,), "zeros", 0, False))
        specs.append((f"layers.{idx + 2}.weight", (cout,), ("const", 0.25), 0, False))
        cin = cout
    for idx, nout, nin in ((14, ndf * 4, ndf * 8), (17, 1, ndf * 4)):
        specs.append((f"layers.{idx}.bias", (nout,), "bias", nin, False))
        specs.append((f"layers.{idx}.weight_orig", (nout, nin), "kaiming", nin, False))
        specs.append((f"layers.{idx}.weight_u", (nout,), "unit", 0, True))
        specs.append((f"layers.{idx}.weight_v", (nin,), "unit", 0, True))
        if idx == 14:
            specs.append(("layers.16.weight", (ndf * 4,), ("const", 0.25), 0, False))
    specs.append(("layers.18.slope", (1,), "ones", 0, False))
    return specs


def _init(shape, kind, fan_in):
    if kind in ("kaiming", "bias"):
        b = 1.0 / math.sqrt(fan_in)
        return torch.empty(shape).uniform_(-b, b)
    if kind == "unit":
        return F.normalize(torch.randn(shape), dim=0, eps=1e-12)
    if kind == "ones":
        return torch.ones(shape)
    if kind == "zeros":
        return torch.zeros(shape)
    return torch.full(shape, float(kind[1]))


class _DiscFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, module, training, seed, *params):
        P = module._tensor_dict()
        save = {} if any(ctx.needs_input_grad) else None
        out = disc_fwd(x, y, P, training, seed, save)
        ctx.module, ctx.saved = module, save
        ctx.need = (ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return out

    @staticmethod
    def backward(ctx, dout):
        module, S = ctx.module, ctx.saved
        if S is None:
            raise RuntimeError("Discriminator backward called but the forward pass did not record state")
        P = module._tensor_dict()
        G, ret = module._grad_targets()
        dx, dy = disc_bwd(S, dout, P, G, *ctx.need)
        ctx.saved = None
        return (dx, dy, None, None, None, *ret)


class Discriminator(nn.Module):
    """Drop-in for reference ``models.discriminator.Discriminator`` (discriminator.py:29-64)."""

    def __init__(self, ndf: int, in_channel: int = 2):
        super().__init__()
        if in_channel != 2 or ndf not in (16, 32):
            raise ValueError("the CUDA path supports in_channel=2 and ndf in {16, 32} (the reference uses ndf=16)")
        self.ndf = ndf
        for key, shape, kind, fan_in, is_buf in _d_specs(ndf, in_channel):
            _register(self, key, _init(shape, kind, fan_in), is_buf)
        self._param_keys = [k for k, _ in self.named_parameters()]
        self.seed, self._step = 0, 0
        self._flat_views = None
        self.flat_grad = None

    def _tensor_dict(self) -> Dict[str, torch.Tensor]:
        d = dict(self.named_parameters())
        d.update(dict(self.named_buffers()))
        return d

    def enable_flat_grads(self) -> torch.Tensor:
        params = list(self.named_parameters())
        sizes = [((p.numel() + 3) // 4) * 4 for _, p in params]
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
        if self._flat_views is not None:
            # ``optimizer.zero_grad()`` (set_to_none=True by default) detaches p.grad from the flat buffer: re-attach the views and give
            # the call its meaning (gradients start from zero) instead of silently accumulating into a buffer the optimiser no longer sees
            named = dict(self.named_parameters())
            if any(named[k].grad is None for k in self._param_keys):
                from .ops import call
                call("cmgan_fill", self.flat_grad, self.flat_grad.numel(), 0.0)
                for k in self._param_keys:
                    named[k].grad = self._flat_views[k]
            return self._flat_views, tuple(None for _ in self._param_keys)
        named = dict(self.named_parameters())
        G = {k: torch.zeros_like(named[k]) for k in self._param_keys}
        return G, tuple(G[k] if named[k].requires_grad else None for k in self._param_keys)

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if not (x.is_cuda and y.is_cuda):
            raise RuntimeError("cmgan_b200.Discriminator runs on CUDA only (no CPU fallback)")
        self._step += 1
        params = [p for _, p in self.named_parameters()]
        return _DiscFn.apply(x, y, self, self.training, (self.seed * 7919 + self._step) * 31 + 5, *params)
