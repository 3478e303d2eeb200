"""The SSNR / STOI restatement (oracle/metrics_oracle.py) against values produced by the reference's own compute_metrics.py
functions on the 25 AudioSamples utterances (fixture written by tools/make_golden_audio.py) and against the per-track lines of
the reference's shipped python_noisy_metrics.log (known-answer test)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import metrics_oracle as MO


@pytest.fixture(scope="module")
def samples():
    z = np.load(os.path.join(GOLDEN, "audiosamples.npz"))
    off = np.concatenate([[0], np.cumsum(z["lengths"])])
    return z, off


def test_metrics_port_matches_reference_functions_and_log(samples):
    z, off = samples
    cols = list(z["metrics_cols"])
    worst = dict(ssnr=0.0, stoi=0.0, log_ssnr=0.0, log_stoi=0.0)
    for i, name in enumerate(z["names"]):
        clean = z["clean"][off[i]:off[i + 1]].astype(np.float64) / 32768.0
        noisy = z["noisy"][off[i]:off[i + 1]].astype(np.float64) / 32768.0
        enh = z["enhanced_ref"][off[i]:off[i + 1]].astype(np.float64)
        row = dict(zip(cols, z["metrics"][i]))
        s_e, s_n = MO.segmental_snr(clean, enh), MO.segmental_snr(clean, noisy)
        worst["ssnr"] = max(worst["ssnr"], abs(s_e - row["ssnr_ref_enh"]), abs(s_n - row["ssnr_noisy"]))
        if i % 3 == 0:          # STOI costs ~0.5 s per call: every third track keeps the CPU suite short
            t_e, t_n = MO.stoi(clean, enh), MO.stoi(clean, noisy)
            worst["stoi"] = max(worst["stoi"], abs(t_e - row["stoi_ref_enh"]), abs(t_n - row["stoi_noisy"]))
            if np.isfinite(row["log_stoi_noisy"]):
                worst["log_stoi"] = max(worst["log_stoi"], abs(t_n - row["log_stoi_noisy"]))
        if np.isfinite(row["log_ssnr_noisy"]):
            worst["log_ssnr"] = max(worst["log_ssnr"], abs(s_n - row["log_ssnr_noisy"]))
    print("[metrics-oracle] worst abs deviation:", worst)
    assert worst["ssnr"] < 1e-8 and worst["stoi"] < 1e-8            # vs the reference's functions run on the same arrays
    assert worst["log_ssnr"] < 2e-6 and worst["log_stoi"] < 2e-6    # vs the shipped log (6 printed decimals)


def test_llr_wss_port_matches_reference_functions_and_log(samples):
    """LLR / WSS restatement against (i) the reference's own ``llr`` / ``wss`` on the 25 utterances (tools/make_golden_quality.py: aggregated
    values for every file, per-frame values for two) and (ii) the CSIG / CBAK / COVL columns of the reference's shipped log: with the log's
    PESQ, the composite measures rebuilt from this port's LLR, WSS and segmental SNR must reproduce the logged values."""
    z, off = samples
    q = np.load(os.path.join(GOLDEN, "audiosamples_quality.npz"))
    cols = list(q["quality_cols"])
    assert list(q["names"]) == list(z["names"])
    worst = dict(llr=0.0, wss=0.0, frames_llr=0.0, frames_wss=0.0, composite=0.0)
    n_log = 0
    for i, name in enumerate(z["names"]):
        c16 = z["clean"][off[i]:off[i + 1]].astype(np.float64)
        n16 = z["noisy"][off[i]:off[i + 1]].astype(np.float64)
        enh = z["enhanced_ref"][off[i]:off[i + 1]].astype(np.float64)
        row = dict(zip(cols, q["quality"][i]))
        l_n, w_n = MO.llr_frames(c16, n16), MO.wss_frames(c16, n16)
        l_e, w_e = MO.llr_frames(c16 / 32768.0, enh), MO.wss_frames(c16 / 32768.0, enh)
        worst["llr"] = max(worst["llr"], abs(MO.trimmed_mean(l_n) - row["llr_noisy_int16"]), abs(MO.trimmed_mean(l_e) - row["llr_ref_enh_unit"]))
        worst["wss"] = max(worst["wss"], abs(MO.trimmed_mean(w_n) - row["wss_noisy_int16"]), abs(MO.trimmed_mean(w_e) - row["wss_ref_enh_unit"]))
        if f"llr_noisy_{i}" in q.files:
            worst["frames_llr"] = max(worst["frames_llr"], np.abs(l_n - q[f"llr_noisy_{i}"]).max(), np.abs(l_e - q[f"llr_enh_{i}"]).max())
            worst["frames_wss"] = max(worst["frames_wss"], np.abs(w_n - q[f"wss_noisy_{i}"]).max(), np.abs(w_e - q[f"wss_enh_{i}"]).max())
        if np.isfinite(row["log_pesq"]):
            n_log += 1
            got = MO.composite(row["log_pesq"], MO.trimmed_mean(l_n), MO.trimmed_mean(w_n), MO.segmental_snr(c16, n16))
            worst["composite"] = max(worst["composite"], *(abs(g - row[k]) for g, k in zip(got, ("log_csig", "log_cbak", "log_covl"))))
    print(f"[metrics-oracle] LLR / WSS vs the reference functions: aggregated {worst['llr']:.2e} / {worst['wss']:.2e}, per frame "
          f"{worst['frames_llr']:.2e} / {worst['frames_wss']:.2e}; CSIG/CBAK/COVL vs the shipped log ({n_log} tracks): {worst['composite']:.2e}")
    assert worst["llr"] < 1e-7 and worst["wss"] < 1e-8 and worst["frames_llr"] < 1e-6 and worst["frames_wss"] < 1e-8
    assert n_log >= 20 and worst["composite"] < 2e-6          # the log prints six decimals
