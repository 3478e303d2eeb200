"""wav I/O and file-list helpers of the evaluation front end (CPU part; the GPU part is in tests/test_gpu_audiosamples.py)."""
import os

import numpy as np

from cmgan_b200 import evaluation as ev
from conftest import GOLDEN


def test_read_wav_matches_int16_scaling_and_roundtrips(tmp_path):
    x, sr = ev.read_wav(os.path.join(GOLDEN, "p232_170_noisy.wav"))
    assert sr == 16000 and x.shape[0] == 1 and x.dtype.is_floating_point
    from scipy.io import wavfile
    _, raw = wavfile.read(os.path.join(GOLDEN, "p232_170_noisy.wav"))
    assert np.array_equal(x[0].numpy(), raw.astype(np.float32) / 32768.0)        # torchaudio.load's normalisation (evaluation.py:17)
    p = str(tmp_path / "f.wav")
    ev.write_wav(p, x[0].numpy())
    y, _ = ev.read_wav(p)
    assert np.array_equal(y.numpy(), x.numpy())
    ev.write_wav(p, x[0].numpy(), subtype="PCM_16")
    z, _ = ev.read_wav(p)
    assert np.abs(z.numpy() - x.numpy()).max() <= 1.0 / 32768.0


def test_natural_sort_orders_digit_runs_numerically():
    names = ["p232_10.wav", "p232_2.wav", "p232_1.wav", "p257_1.wav", "p232_100.wav"]
    assert ev.natural_sorted(names) == ["p232_1.wav", "p232_2.wav", "p232_10.wav", "p232_100.wav", "p257_1.wav"]
