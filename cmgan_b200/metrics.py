"""PESQ-free scores of the reference's compute_metrics.py on the GPU: segmental SNR (:350-397) and STOI (:400-471), float64 like the
numpy original, through the C ABI (csrc/metrics.cu).  PESQ itself (and CSIG / CBAK / COVL, which are affine in PESQ) needs the
third-party ``pesq`` package and stays on the host when that is installed.

    ssnr, stoi = ssnr_stoi(clean, enhanced)      # (L,) tensors on the GPU, 16 kHz
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

from ._lib import lib
from .ops import stream

_CONST: Dict[Tuple, tuple] = {}


def _resample_filter() -> np.ndarray:
    """the low-pass scipy.signal.resample_poly(x, 10000, 16000) designs: firwin(161, 1/8, window=('kaiser', 5.0)) * 5, restated"""
    half, fc = 80, 1.0 / 8.0
    m = np.arange(2 * half + 1, dtype=np.float64) - half
    h = fc * np.sinc(fc * m) * np.kaiser(2 * half + 1, 5.0)
    return h / h.sum() * 5.0


def _third_octave_bins(fs=10000, nfft=512, nbands=15, first=150.0):
    """[lo, hi) DFT-bin ranges of the 1/3-octave bands (compute_metrics.py:474-519)"""
    f = np.linspace(0, fs, nfft + 1)[: nfft // 2 + 1]
    k = np.arange(nbands)
    cf = first * 2.0 ** (k / 3.0)
    lo = np.sqrt(cf * first * 2.0 ** ((k - 1) / 3.0))
    hi = np.sqrt(cf * first * 2.0 ** ((k + 1) / 3.0))
    a = [int(np.argmin((f - v) ** 2)) for v in lo]
    b = [int(np.argmin((f - v) ** 2)) for v in hi]
    return np.array(a, dtype=np.int32), np.array(b, dtype=np.int32)


def _consts(dev):
    key = (str(dev),)
    if key not in _CONST:
        lo, hi = _third_octave_bins()
        _CONST[key] = (torch.from_numpy(_resample_filter()).to(dev), torch.from_numpy(lo).to(dev), torch.from_numpy(hi).to(dev))
    return _CONST[key]


def ssnr_stoi(clean: torch.Tensor, proc: torch.Tensor, fs: int = 16000) -> Tuple[float, float]:
    """(mean segmental SNR in dB, STOI) of ``proc`` against ``clean``: 1-D tensors of equal length on the GPU"""
    if not (clean.is_cuda and proc.is_cuda):
        raise RuntimeError("cmgan_b200.metrics runs on CUDA only")
    assert clean.dim() == 1 and clean.shape == proc.shape and fs == 16000
    c, p = clean.double().contiguous(), proc.double().contiguous()
    L = c.numel()
    dev = c.device
    h, lo, hi = _consts(dev)
    W = round(30 * fs / 1000)
    skip = W // 4
    nfr = int(L / skip - W / skip)
    out = torch.zeros(3, dtype=torch.float64, device=dev)
    L_ = lib()
    L_.call("cmgan_ssnr_f64", c.data_ptr(), p.data_ptr(), L, W, skip, nfr, out.data_ptr(), stream())
    scratch = torch.empty(int(L_.cdll.cmgan_stoi_scratch_doubles(L)), dtype=torch.float64, device=dev)
    L_.call("cmgan_stoi_f64", c.data_ptr(), p.data_ptr(), L, h.data_ptr(), lo.data_ptr(), hi.data_ptr(), scratch.data_ptr(), out[1:].data_ptr(), stream())
    o = out.cpu()
    return float(o[0]), float(o[1] / o[2]) if o[2] > 0 else math.nan
