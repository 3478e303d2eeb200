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
