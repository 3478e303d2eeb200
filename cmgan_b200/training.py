"""Per-step glue of the reference's Trainer (train.py:72-151) over the CUDA path.

``forward_generator_step`` mirrors ``Trainer.forward_generator_step`` (same dictionary keys, same shapes);
``generator_loss`` mirrors ``calculate_generator_loss`` (including the reference's quirk that the time-domain L1
compares the RMS-normalised estimate with the *un-normalised* clean waveform, train.py:140-142,188).
The loss reductions themselves are a handful of PyTorch element-wise ops on (B, 1, F, T) tensors for now
("negligible FLOPs"); everything upstream of them is libcmgan_b200.so.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import signal


def forward_generator_step(model, clean: torch.Tensor, noisy: torch.Tensor) -> Dict[str, torch.Tensor]:
    c = signal.rms_scale(noisy)
    noisy_spec = signal.stft_compress(noisy, c).permute(0, 1, 3, 2)            # (B, 2, T, F)
    clean_spec = signal.stft_compress(clean, c)                                # (B, 2, F, T) view
    clean_real, clean_imag = clean_spec[:, 0:1], clean_spec[:, 1:2]
    est_real, est_imag = model(noisy_spec)
    est_audio = signal.uncompress_istft(est_real, est_imag)
    est_real, est_imag = est_real.permute(0, 1, 3, 2), est_imag.permute(0, 1, 3, 2)
    est_mag = torch.sqrt(est_real ** 2 + est_imag ** 2)
    clean_mag = torch.sqrt(clean_real ** 2 + clean_imag ** 2)
    return dict(est_real=est_real, est_imag=est_imag, est_mag=est_mag, clean_real=clean_real, clean_imag=clean_imag,
                clean_mag=clean_mag, est_audio=est_audio)


def generator_loss(go: Dict[str, torch.Tensor], clean: torch.Tensor, discriminator: Optional[torch.nn.Module] = None,
                   weights=(0.1, 0.9, 0.2, 0.05)) -> torch.Tensor:
    loss_mag = F.mse_loss(go["est_mag"], go["clean_mag"])
    loss_ri = F.mse_loss(go["est_real"], go["clean_real"]) + F.mse_loss(go["est_imag"], go["clean_imag"])
    time_loss = torch.mean(torch.abs(go["est_audio"] - clean[:, :go["est_audio"].shape[-1]]))
    loss = weights[0] * loss_ri + weights[1] * loss_mag + weights[2] * time_loss
    if discriminator is not None:
        fake = discriminator(go["clean_mag"], go["est_mag"])
        loss = loss + weights[3] * F.mse_loss(fake.flatten(), torch.ones(fake.shape[0], device=fake.device))
    return loss
