"""GPU scoring kernels (cmgan_b200.metrics: segmental SNR and STOI of compute_metrics.py:350-471) against the values the reference's own
functions produced for the 25 AudioSamples utterances (fixture: tools/make_golden_audio.py), for the reference-enhanced and the noisy signals."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
if torch.cuda.is_available():
    from cmgan_b200 import metrics
from conftest import GOLDEN


def test_gpu_ssnr_stoi_match_reference_values():
    z = np.load(os.path.join(GOLDEN, "audiosamples.npz"))
    off = np.concatenate([[0], np.cumsum(z["lengths"])])
    cols = list(z["metrics_cols"])
    worst = dict(ssnr=0.0, stoi=0.0)
    for i, name in enumerate(z["names"]):
        clean = torch.from_numpy(z["clean"][off[i]:off[i + 1]].astype(np.float64) / 32768.0).cuda()
        noisy = torch.from_numpy(z["noisy"][off[i]:off[i + 1]].astype(np.float64) / 32768.0).cuda()
        enh = torch.from_numpy(z["enhanced_ref"][off[i]:off[i + 1]].astype(np.float64)).cuda()
        row = dict(zip(cols, z["metrics"][i]))
        s_e, t_e = metrics.ssnr_stoi(clean, enh)
        s_n, t_n = metrics.ssnr_stoi(clean, noisy)
        worst["ssnr"] = max(worst["ssnr"], abs(s_e - row["ssnr_ref_enh"]), abs(s_n - row["ssnr_noisy"]))
        worst["stoi"] = max(worst["stoi"], abs(t_e - row["stoi_ref_enh"]), abs(t_n - row["stoi_noisy"]))
    print(f"[gpu-metrics] 25 files x 2 signals: worst |SSNR - reference| {worst['ssnr']:.3e} dB, worst |STOI - reference| {worst['stoi']:.3e}")
    assert worst["ssnr"] < 1e-8 and worst["stoi"] < 1e-9


def test_gpu_llr_wss_match_reference_values():
    """GPU LLR / WSS kernels against the reference's own ``llr`` / ``wss`` values (tools/make_golden_quality.py) on the 25 utterances, at the
    16-bit sample scale the shipped log was produced with and at the unit scale evaluation.py uses; composite measures against the log."""
    z = np.load(os.path.join(GOLDEN, "audiosamples.npz"))
    q = np.load(os.path.join(GOLDEN, "audiosamples_quality.npz"))
    off = np.concatenate([[0], np.cumsum(z["lengths"])])
    cols = list(q["quality_cols"])
    worst = dict(llr=0.0, wss=0.0, frames_llr=0.0, frames_wss=0.0, composite=0.0)
    for i, name in enumerate(z["names"]):
        c16 = torch.from_numpy(z["clean"][off[i]:off[i + 1]].astype(np.float64)).cuda()
        n16 = torch.from_numpy(z["noisy"][off[i]:off[i + 1]].astype(np.float64)).cuda()
        enh = torch.from_numpy(z["enhanced_ref"][off[i]:off[i + 1]].astype(np.float64)).cuda()
        row = dict(zip(cols, q["quality"][i]))
        l_n, w_n = metrics.llr_wss(c16, n16)
        l_e, w_e = metrics.llr_wss(c16 / 32768.0, enh)
        worst["llr"] = max(worst["llr"], abs(l_n - row["llr_noisy_int16"]), abs(l_e - row["llr_ref_enh_unit"]))
        worst["wss"] = max(worst["wss"], abs(w_n - row["wss_noisy_int16"]), abs(w_e - row["wss_ref_enh_unit"]))
        if f"llr_noisy_{i}" in q.files:
            fl, fw = metrics.llr_wss_frames(c16, n16)
            worst["frames_llr"] = max(worst["frames_llr"], float(np.abs(fl.cpu().numpy() - q[f"llr_noisy_{i}"]).max()))
            worst["frames_wss"] = max(worst["frames_wss"], float(np.abs(fw.cpu().numpy() - q[f"wss_noisy_{i}"]).max()))
        if np.isfinite(row["log_pesq"]):
            s_n, _ = metrics.ssnr_stoi(c16, n16)
            got = metrics.composite(row["log_pesq"], l_n, w_n, s_n)
            worst["composite"] = max(worst["composite"], *(abs(g - row[k]) for g, k in zip(got, ("log_csig", "log_cbak", "log_covl"))))
    print(f"[gpu-metrics] LLR / WSS vs the reference functions: aggregated {worst['llr']:.2e} / {worst['wss']:.2e}, per frame "
          f"{worst['frames_llr']:.2e} / {worst['frames_wss']:.2e}; CSIG/CBAK/COVL vs the shipped log {worst['composite']:.2e}")
    assert worst["llr"] < 1e-6 and worst["wss"] < 1e-7 and worst["frames_llr"] < 1e-5 and worst["frames_wss"] < 1e-7
    assert worst["composite"] < 2e-6
