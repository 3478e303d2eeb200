"""Whole-path parity on real speech: the 25 utterances of the reference's AudioSamples (2.1 - 9.8 s; time sequences up to
L = 1564, i.e. the +-512 relative-position clamp inside the whole network) enhanced by the CUDA path (evaluation.py:21-53 glue,
tf32 tensor-core mode = the timed configuration, plus exact-fp32 mode on a subset) against the reference's own output
(fixture: tools/make_golden_audio.py).  Reported per file: waveform max-abs (original scale, north-star bound 1e-3), SNR of
our output vs the reference's, and the deltas of the reference's PESQ-free scores SSNR / STOI (compute_metrics.py:350-471;
PESQ itself is unpinned: the pesq package is absent)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
if torch.cuda.is_available():
    import cmgan_b200
    from cmgan_b200 import ops, signal
from conftest import GOLDEN
from oracle import metrics_oracle as MO


@pytest.fixture(scope="module")
def model(g_weights):
    m = cmgan_b200.TSCNet(64, 201)
    m.load_state_dict(g_weights, strict=True)
    return m.to(DEV).eval()


def _run(model, mode, idx):
    z = np.load(os.path.join(GOLDEN, "audiosamples.npz"))
    off = np.concatenate([[0], np.cumsum(z["lengths"])])
    cols = list(z["metrics_cols"])
    rows = []
    ops.set_precision(mode)
    try:
        for i in idx:
            name = str(z["names"][i])
            noisy = torch.from_numpy(z["noisy"][off[i]:off[i + 1]].astype(np.float32) / 32768.0).unsqueeze(0)
            ref = z["enhanced_ref"][off[i]:off[i + 1]].astype(np.float64)
            clean = z["clean"][off[i]:off[i + 1]].astype(np.float64) / 32768.0
            out = signal.enhance(model, noisy.to(DEV)).double().cpu().numpy()
            err = np.abs(out - ref).max()
            snr = 10 * np.log10((ref ** 2).sum() / max(((out - ref) ** 2).sum(), 1e-300))
            row = dict(zip(cols, z["metrics"][i]))
            d_ssnr = MO.segmental_snr(clean, out) - row["ssnr_ref_enh"]
            d_stoi = MO.stoi(clean, out) - row["stoi_ref_enh"]
            rows.append((name, len(ref) / 16000.0, err, snr, d_ssnr, d_stoi))
            print(f"[audiosamples-{mode}] {name} {len(ref) / 16000.0:5.2f} s: max-abs {err:.3e}  SNR vs reference {snr:5.1f} dB  "
                  f"dSSNR {d_ssnr:+.4f} dB  dSTOI {d_stoi:+.2e}")
    finally:
        ops.set_precision("fp32")
    return rows


def test_audiosamples_tf32(model):
    rows = _run(model, "tf32", range(25))
    worst = max(r[2] for r in rows)
    print(f"[audiosamples-tf32] 25 files: worst max-abs {worst:.3e}, min SNR {min(r[3] for r in rows):.1f} dB, "
          f"mean |dSSNR| {np.mean([abs(r[4]) for r in rows]):.4f} dB (max {max(abs(r[4]) for r in rows):.4f}), "
          f"max |dSTOI| {max(abs(r[5]) for r in rows):.2e}")
    assert worst <= 1e-3, "north-star bound: enhanced waveform max-abs <= 1e-3 vs the reference forward (original scale)"
    assert max(abs(r[4]) for r in rows) <= 0.05 and max(abs(r[5]) for r in rows) <= 1e-3


def test_audiosamples_fp32_subset(model):
    rows = _run(model, "fp32", [0, 8, 10, 12, 13])        # incl. the two longest files (8.5 s, 9.8 s) and the two shortest
    worst = max(r[2] for r in rows)
    print(f"[audiosamples-fp32] worst max-abs {worst:.3e}, min SNR {min(r[3] for r in rows):.1f} dB")
    assert worst <= 1e-4 and max(abs(r[4]) for r in rows) <= 0.01 and max(abs(r[5]) for r in rows) <= 1e-4
