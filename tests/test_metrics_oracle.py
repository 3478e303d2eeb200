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
