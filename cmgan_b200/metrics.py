"""PESQ-free scores of the reference's compute_metrics.py on the GPU: segmental SNR (:350-397), STOI (:400-471), log-likelihood ratio
(:277-347) and weighted spectral slope (:80-274), float64 like the numpy original, through the C ABI (csrc/metrics.cu).  PESQ itself needs
the third-party ``pesq`` package and stays on the host when that is installed; ``composite`` turns a PESQ value plus the three GPU
measures into CSIG / CBAK / COVL (:66-75).

    ssnr, stoi = ssnr_stoi(clean, enhanced)      # (L,) tensors on the GPU, 16 kHz
    llr, wss = llr_wss(clean, enhanced)          # trimmed means as compute_metrics.py:45-55 (16-bit sample scale expected, as there)
    csig, cbak, covl = composite(pesq_mos, llr, wss, ssnr)
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


# ---- LLR / WSS / composite measures
_WSS_CENT = (50.0, 120.0, 190.0, 260.0, 330.0, 400.0, 470.0, 540.0, 617.372, 703.378, 798.717, 904.128, 1020.38, 1148.30, 1288.72, 1442.54,
             1610.70, 1794.16, 1993.93, 2211.08, 2446.71, 2701.97, 2978.04, 3276.17, 3597.63)
_WSS_BW = (70.0, 70.0, 70.0, 70.0, 70.0, 70.0, 70.0, 77.3724, 86.0056, 95.3398, 105.411, 116.256, 127.914, 140.423, 153.823, 168.154, 183.457,
           199.776, 217.153, 235.631, 255.255, 276.072, 298.126, 321.465, 346.136)


def _wss_filters(fs: int, W: int) -> Tuple[np.ndarray, int]:
    """(25, nfft / 2) Gaussian critical-band filters (compute_metrics.py:101-185) and nfft = 2^ceil(log2(2 W))"""
    nfft = 1 << int(math.ceil(math.log2(2 * W)))
    half, fmax = nfft // 2, fs // 2
    cent, bw = np.array(_WSS_CENT), np.array(_WSS_BW)
    j = np.arange(half, dtype=np.float64)[None, :]
    f0 = np.floor(cent / fmax * half)[:, None]
    wd = (bw / fmax * half)[:, None]
    filt = np.exp(-11.0 * ((j - f0) / wd) ** 2 + (math.log(bw[0]) - np.log(bw))[:, None])
    filt[filt <= math.exp(-30.0 / (2.0 * 2.303))] = 0.0
    return np.ascontiguousarray(filt), nfft


def _trimmed_mean(x: torch.Tensor, alpha: float = 0.95) -> float:
    """mean of the smallest round(alpha N) frame values (compute_metrics.py:47-55)"""
    n = round(x.numel() * alpha)
    return float(torch.sort(x).values[:n].mean()) if n > 0 else math.nan


def llr_wss_frames(clean: torch.Tensor, proc: torch.Tensor, fs: int = 16000) -> Tuple[torch.Tensor, torch.Tensor]:
    """per-frame LLR and WSS distortions (float64 tensors on the GPU)"""
    if not (clean.is_cuda and proc.is_cuda):
        raise RuntimeError("cmgan_b200.metrics runs on CUDA only")
    assert clean.dim() == 1 and clean.shape == proc.shape
    c, p = clean.double().contiguous(), proc.double().contiguous()
    L, dev = c.numel(), c.device
    W = round(30 * fs / 1000)
    skip = W // 4
    order = 10 if fs < 10000 else 16
    n_llr = int((L - W) / skip)
    n_wss = int(L / skip - W / skip)
    key = ("wss", str(dev), fs)
    if key not in _CONST:
        filt, nfft = _wss_filters(fs, W)
        _CONST[key] = (torch.from_numpy(filt).to(dev), nfft)
    filt, nfft = _CONST[key]
    llr = torch.empty(max(n_llr, 0), dtype=torch.float64, device=dev)
    wss = torch.empty(max(n_wss, 0), dtype=torch.float64, device=dev)
    L_ = lib()
    L_.call("cmgan_llr_f64", c.data_ptr(), p.data_ptr(), L, W, skip, order, max(n_llr, 0), llr.data_ptr(), stream())
    L_.call("cmgan_wss_f64", c.data_ptr(), p.data_ptr(), L, W, skip, nfft, filt.data_ptr(), max(n_wss, 0), wss.data_ptr(), stream())
    return llr, wss


def llr_wss(clean: torch.Tensor, proc: torch.Tensor, fs: int = 16000) -> Tuple[float, float]:
    """(llr_mean, wss_dist) exactly as compute_metrics.py:45-55 aggregates them: mean over the lowest 95 % of the frame values"""
    llr, wss = llr_wss_frames(clean, proc, fs)
    return _trimmed_mean(llr), _trimmed_mean(wss)


def composite(pesq_mos: float, llr_mean: float, wss_dist: float, seg_snr: float) -> Tuple[float, float, float]:
    """CSIG, CBAK, COVL (compute_metrics.py:66-75), each limited to [1, 5]"""
    csig = 3.093 - 1.029 * llr_mean + 0.603 * pesq_mos - 0.009 * wss_dist
    cbak = 1.634 + 0.478 * pesq_mos - 0.007 * wss_dist + 0.063 * seg_snr
    covl = 1.594 + 0.805 * pesq_mos - 0.512 * llr_mean - 0.007 * wss_dist
    return tuple(min(5.0, max(1.0, v)) for v in (csig, cbak, covl))
