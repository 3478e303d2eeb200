"""Graph-free training step: the body of the reference's ``Trainer.train_step`` (train.py:176-205) orchestrated directly over
the forward/backward kernel sequences, without building an autograd graph.

    t = FusedTrainer(TSCNet().cuda(), Discriminator(16).cuda())
    loss = t.generator_step(clean, noisy)                 # forward, losses + gradients (fused), backward, all-reduce, AdamW
    dloss = t.discriminator_step(pesq_target)             # reuses clean_mag / est_mag of the generator step (train.py:153-174)

Parameters and gradients live in flat fp32 buffers (one NCCL all-reduce and one AdamW kernel per network per step).
PESQ itself is host code of the reference (discriminator.py:9-26) and stays outside: the caller passes (pesq-1)/3.5 targets.
"""
from __future__ import annotations

import os

from typing import Optional

import torch

from . import ops, parallel, signal
from .discriminator import Discriminator, disc_bwd, disc_fwd
from .generator import TSCNet
from .network import tscnet_bwd, tscnet_fwd
from .ops import call


def _flatten_params(module: torch.nn.Module):
    """move every parameter into one flat buffer (16-byte aligned segments); returns (flat_param, name -> view)"""
    params = list(module.named_parameters())
    sizes = [((p.numel() + 3) // 4) * 4 for _, p in params]
    flat = torch.zeros(sum(sizes), device=params[0][1].device)
    off = 0
    for (k, p), n in zip(params, sizes):
        v = flat[off:off + p.numel()].view_as(p)
        v.copy_(p.data)
        p.data = v
        off += n
    return flat


class _Adam:
    def __init__(self, flat_p, flat_g, lr, betas=(0.9, 0.999), eps=1e-8, wd=0.01):
        self.p, self.g = flat_p, flat_g
        self.m, self.v = torch.zeros_like(flat_p), torch.zeros_like(flat_p)
        self.lr, self.betas, self.eps, self.wd, self.t = lr, betas, eps, wd, 0
        self.t_dev = torch.zeros(1, dtype=torch.int64, device=flat_p.device)      # device-side step count (CUDA-graph replay safe)

    def step(self):
        self.t += 1
        call("cmgan_counter_add", self.t_dev, 1)
        call("cmgan_adamw", self.p, self.g, self.m, self.v, self.p.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.t, self.t_dev)


class FusedTrainer:
    def __init__(self, model: TSCNet, disc: Optional[Discriminator] = None, lr: float = 5e-4, weights=(0.1, 0.9, 0.2, 0.05), seed: int = 0):
        self.model, self.disc, self.w = model, disc, weights
        if ops.WGRAD_STREAM is None and os.environ.get("CMGAN_WGRAD_STREAM", "1") != "0":
            ops.WGRAD_STREAM = torch.cuda.Stream()        # weight-gradient GEMMs overlap the data-gradient chain (also inside the CUDA graph)
        self.pg = _flatten_params(model)
        self.gg = model.enable_flat_grads()
        self.opt_g = _Adam(self.pg, self.gg, lr)                      # train.py:63
        parallel.broadcast_module(model)
        if disc is not None:
            self.pd = _flatten_params(disc)
            self.gd = disc.enable_flat_grads()
            self.opt_d = _Adam(self.pd, self.gd, 2 * lr)              # train.py:64-66
            parallel.broadcast_module(disc)
        self.seed, self.step_no = seed, 0
        self.last = None
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.pg.device)   # added to every dropout seed on the device
        self._graph = None

    def generator_step(self, clean: torch.Tensor, noisy: torch.Tensor, update: bool = True, allreduce: bool = True) -> torch.Tensor:
        """train.py:179-193.  clean / noisy: (B, L) un-normalised waveforms on the GPU.  Returns the loss (device scalar)."""
        m, dev = self.model, clean.device
        if self._graph is None or torch.cuda.is_current_stream_capturing():
            self.step_no += 1
        seed = self.seed * 7919 + self.step_no
        B, L = noisy.shape
        ops.SEED_DEV = self.step_dev
        call("cmgan_counter_add", self.step_dev, 1)
        call("cmgan_fill", self.gg, self.gg.numel(), 0.0)
        c = signal.rms_scale(noisy)
        noisy_spec = signal.stft_compress(noisy, c).permute(0, 1, 3, 2)           # (B,2,T,F)
        clean_spec = signal.stft_compress(clean, c).permute(0, 1, 3, 2)           # contiguous (B,2,T,F) memory
        T, F = noisy_spec.shape[2], noisy_spec.shape[3]
        P = m._tensor_dict()
        if m.training:
            torch._foreach_add_([b for k, b in m.named_buffers() if k.endswith("num_batches_tracked")], 1)
        S = {}
        fr, fi = tscnet_fwd(noisy_spec, P, m.training, seed, S)
        est_audio = signal.uncompress_istft_fwd(fr, fi)
        n = B * T * F
        acc = torch.zeros(3, dtype=torch.float64, device=dev)
        d_er, d_ei = torch.empty_like(fr), torch.empty_like(fi)
        est_mag, clean_mag = torch.empty_like(fr), torch.empty_like(fr)
        call("cmgan_spec_loss", fr, fi, clean_spec, (clean_spec, T * F), T * F, 2 * T * F, n, self.w[0], self.w[1], acc, d_er, d_ei, est_mag, clean_mag)
        Lo = est_audio.shape[1]
        d_audio = torch.empty_like(est_audio)
        call("cmgan_time_loss", est_audio, est_audio.stride(0), clean, clean.stride(0), B, Lo, self.w[2], acc, d_audio)
        loss = torch.empty(1, device=dev)
        if self.disc is not None:
            Pd = self.disc._tensor_dict()
            Sd = {}
            cm, em = clean_mag.permute(0, 1, 3, 2), est_mag.permute(0, 1, 3, 2)      # (B,1,F,T) views, as the reference passes them
            fake = disc_fwd(cm, em, Pd, self.disc.training, seed * 31 + 5, Sd)
            d_fake = torch.empty_like(fake)
            call("cmgan_gen_loss_finalize", acc, float(n), float(B * Lo), self.w[0], self.w[1], self.w[2], self.w[3], fake, B, loss, d_fake)
            _, d_mag = disc_bwd(Sd, d_fake, Pd, None, False, True)                   # no parameter gradients: optimizer_disc.zero_grad() discards them
            gs = d_mag.stride()
            call("cmgan_mag_bwd_add", fr, fi, d_mag, gs[0], gs[3], gs[2], B, T, F, d_er, d_ei)
        else:
            call("cmgan_gen_loss_finalize", acc, float(n), float(B * Lo), self.w[0], self.w[1], self.w[2], 0.0, None, B, loss, None)
        signal.uncompress_istft_bwd(fr, fi, d_audio, d_er, d_ei, True)
        tscnet_bwd(S, d_er, d_ei, P, m._flat_views)
        ops.SEED_DEV = None
        if allreduce:
            parallel.allreduce_mean_(self.gg)
        if update:
            self.opt_g.step()
        self.last = dict(clean_mag=clean_mag, est_mag=est_mag, est_audio=est_audio, B=B)
        return loss

    def discriminator_step(self, pesq_target: torch.Tensor, update: bool = True) -> torch.Tensor:
        """train.py:161-170,199-201 once the PESQ targets exist: MSE(D(c,c),1) + MSE(D(c, est.detach()), target)."""
        d, L = self.disc, self.last
        dev = pesq_target.device
        ops.SEED_DEV = self.step_dev
        Pd = d._tensor_dict()
        call("cmgan_fill", self.gd, self.gd.numel(), 0.0)
        cm, em = L["clean_mag"].permute(0, 1, 3, 2), L["est_mag"].permute(0, 1, 3, 2)
        s1, s2 = {}, {}
        seed = (self.seed * 7919 + self.step_no) * 131
        d_enh = disc_fwd(cm, em, Pd, d.training, seed + 1, s1)
        d_max = disc_fwd(cm, cm, Pd, d.training, seed + 2, s2)
        loss = torch.empty(1, device=dev)
        g_max, g_enh = torch.empty_like(d_max), torch.empty_like(d_enh)
        call("cmgan_disc_loss", d_max, d_enh, pesq_target, L["B"], loss, g_max, g_enh)
        disc_bwd(s1, g_enh, Pd, d._flat_views, False, False)
        disc_bwd(s2, g_max, Pd, d._flat_views, False, False)
        ops.SEED_DEV = None
        parallel.allreduce_mean_(self.gd)
        if update:
            self.opt_d.step()
        return loss

    # ------------------------------------------------------------------ CUDA graph of the generator step
    def capture_generator_step(self, clean: torch.Tensor, noisy: torch.Tensor, update: bool = True, allreduce: bool = True) -> None:
        """Record ``generator_step`` once into a CUDA graph (all ~500 kernel launches, the NCCL all-reduce and AdamW); afterwards
        ``replay_generator_step`` costs two H2D/D2D copies and one graph launch on the host.  Dropout masks and Adam bias
        corrections stay fresh across replays because seeds / step counts are read from device counters."""
        self.static_clean, self.static_noisy = clean.clone(), noisy.clone()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                self.generator_step(self.static_clean, self.static_noisy, update, allreduce)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self.static_loss = self.generator_step(self.static_clean, self.static_noisy, update, allreduce)

    def replay_generator_step(self, clean: torch.Tensor, noisy: torch.Tensor) -> torch.Tensor:
        self.static_clean.copy_(clean, non_blocking=True)
        self.static_noisy.copy_(noisy, non_blocking=True)
        self._graph.replay()
        return self.static_loss
