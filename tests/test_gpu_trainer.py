"""Graph-free fused training step (cmgan_b200.trainer) against the autograd path built from the same kernels, and the
flat AdamW kernel against torch.optim.AdamW."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
if torch.cuda.is_available():
    import cmgan_b200
    from cmgan_b200 import training
    from cmgan_b200.ops import call
    from cmgan_b200.trainer import FusedTrainer


def _models(g_weights, d_weights):
    m = cmgan_b200.TSCNet(64, 201)
    m.load_state_dict(g_weights, strict=True)
    d = cmgan_b200.Discriminator(16)
    d.load_state_dict(d_weights, strict=True)
    return m.to(DEV).eval(), d.to(DEV).eval()


def test_fused_generator_and_discriminator_step(g_weights, d_weights, golden, monkeypatch):
    import cmgan_b200.discriminator as Dm
    monkeypatch.setattr(Dm, "DROP_P", 0.0)          # deterministic comparison; the discriminator runs in train mode (power iterations)
    clean = torch.from_numpy(golden["grad_clean"]).to(DEV)
    noisy = torch.from_numpy(golden["grad_noisy"]).to(DEV)
    # autograd path
    m, d = _models(g_weights, d_weights)
    d.train()
    go = training.forward_generator_step(m, clean, noisy)
    loss = training.generator_loss(go, clean, d)
    loss.backward()
    ref = {k: p.grad.clone() for k, p in m.named_parameters()}
    pesq_t = torch.tensor([0.35, 0.6], device=DEV)
    for p in d.parameters():
        p.grad = None
    d_enh = d(go["clean_mag"], go["est_mag"].detach())
    d_max = d(go["clean_mag"], go["clean_mag"])
    dl = F.mse_loss(d_max.flatten(), torch.ones(2, device=DEV)) + F.mse_loss(d_enh.flatten(), pesq_t)
    dl.backward()
    dref = {k: p.grad.clone() for k, p in d.named_parameters()}
    # fused path
    m2, d2 = _models(g_weights, d_weights)
    d2.train()
    t = FusedTrainer(m2, d2)
    loss2 = t.generator_step(clean, noisy, update=False)
    print(f"[parity] generator loss fused {loss2.item():.7f} vs autograd {loss.item():.7f}")
    assert abs(loss2.item() - loss.item()) < 2e-6 * max(1.0, abs(loss.item()))
    gmax = max(v.abs().max().item() for v in ref.values())
    for k, p in m2.named_parameters():
        e = (p.grad - ref[k]).abs().max().item() / max(ref[k].abs().max().item(), 1e-3 * gmax)
        assert e < 2e-3, f"{k}: {e}"
    dl2 = t.discriminator_step(pesq_t, update=False)
    print(f"[parity] discriminator loss fused {dl2.item():.7f} vs autograd {dl.item():.7f}")
    assert abs(dl2.item() - dl.item()) < 2e-6
    gmax = max(v.abs().max().item() for v in dref.values())
    for k, p in d2.named_parameters():
        e = (p.grad - dref[k]).abs().max().item() / max(dref[k].abs().max().item(), 1e-3 * gmax, 1e-20)
        assert e < 2e-3, f"D {k}: {e}"


def test_adamw_kernel_matches_torch():
    torch.manual_seed(0)
    p0 = torch.randn(10007, device=DEV)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=5e-4)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for step in range(1, 4):
        g = torch.randn(10007, device=DEV)
        ref.grad = g.clone()
        opt.step()
        call("cmgan_adamw", p, g, m, v, p.numel(), 5e-4, 0.9, 0.999, 1e-8, 0.01, step, None, None)
    err = (p - ref.detach()).abs().max().item()
    print(f"[parity] AdamW 3 steps max-abs {err:.3e}")
    assert err < 1e-6


def test_training_makes_progress(g_weights, d_weights, golden):
    """a few real optimiser steps in train mode (dropout, BatchNorm batch statistics, spectral-norm power iterations): finite, loss decreases"""
    m, d = _models(g_weights, d_weights)
    m.train(); d.train()
    t = FusedTrainer(m, d, lr=2e-4)
    clean = torch.from_numpy(golden["grad_clean"]).to(DEV)
    noisy = torch.from_numpy(golden["grad_noisy"]).to(DEV)
    losses = []
    for _ in range(6):
        losses.append(t.generator_step(clean, noisy).item())
        dl = t.discriminator_step(torch.tensor([0.4, 0.5], device=DEV)).item()
        assert np.isfinite(dl)
    print("[train] generator losses:", [f"{x:.4f}" for x in losses])
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_lr_schedule_under_graph_replay_and_checkpoint_roundtrip(g_weights, d_weights, golden, tmp_path):
    """StepLR (train.py:248-253) must act on a CAPTURED step (lr is a device scalar), capturing must not consume training steps or
    touch BatchNorm / spectral-norm buffers, and save -> load -> step must continue bit-exactly (train.py:273, evaluation.py:64)."""
    clean = torch.from_numpy(golden["grad_clean"]).to(DEV)
    noisy = torch.from_numpy(golden["grad_noisy"]).to(DEV)
    m, d = _models(g_weights, d_weights)
    m.train(); d.train()
    t = FusedTrainer(m, d, lr=1e-3, decay_epoch=1)
    bufs0 = {k: v.clone() for k, v in list(m.named_buffers()) + [("D." + k, v) for k, v in d.named_buffers()]}
    p0 = t.pg.clone()
    t.capture_train_step(clean, noisy)
    assert torch.equal(t.pg, p0), "capture must not update the parameters"
    for k, v in list(m.named_buffers()) + [("D." + k, v) for k, v in d.named_buffers()]:
        assert torch.equal(v, bufs0[k]), f"capture changed buffer {k}"
    assert t.opt_g.t == 0 and int(t.opt_g.t_dev.item()) == 0 and int(t.step_dev.item()) == 0
    pesq_t = torch.tensor([0.4, 0.6], device=DEV)
    lg, ld = t.replay_train_step(clean, noisy, pesq_t)
    assert np.isfinite(lg.item()) and np.isfinite(ld.item())
    step1 = (t.pg - p0).abs().max().item()
    assert step1 > 0 and int(t.opt_g.t_dev.item()) == 1
    # a full checkpoint, then two more replays under a halved learning rate
    ck = str(tmp_path / "ckpt_full")
    t.save_checkpoint(ck, full=True)
    ref_ck = str(tmp_path / "ckpt_ref_format")
    t.save_checkpoint(ref_ck)
    t.scheduler_step()            # decay_epoch = 1: lr 1e-3 -> 5e-4, D 2e-3 -> 1e-3
    assert abs(t.opt_g.lr - 5e-4) < 1e-12 and abs(t.opt_d.lr - 1e-3) < 1e-12
    p1 = t.pg.clone()
    t.replay_train_step(clean, noisy, pesq_t)
    p2 = t.pg.clone()
    # Adam's first steps move every weight by ~lr: the captured graph must follow the new device-side lr
    ratio = (p2 - p1).abs().mean().item() / (p1 - p0).abs().mean().item()
    print(f"[lr] mean |update| after halving lr / before: {ratio:.3f}")
    assert 0.2 < ratio < 0.8
    # resume from the checkpoint in a fresh trainer: same lr schedule position, same next step
    m2, d2 = _models(g_weights, d_weights)
    m2.train(); d2.train()
    t2 = FusedTrainer(m2, d2, lr=1e-3, decay_epoch=1)
    t2.load_checkpoint(ck)
    assert torch.equal(t2.pg, p1)
    t2.scheduler_step()
    t2.generator_step(clean, noisy)
    t2.discriminator_step(pesq_t)
    # same dropout masks (device counter), same statistics; atomics reorder the gradient sums, and AdamW turns a noise-level gradient
    # (conv biases in front of an InstanceNorm: mathematically zero) into a +-lr step, so compare robustly
    diff = (t2.pg - p2).abs()
    print(f"[ckpt] resumed eager step vs original graph replay: parameter difference median {diff.median().item():.3e}, "
          f"99.9th percentile {diff.float().quantile(0.999).item():.3e}, max {diff.max().item():.3e}")
    assert diff.median().item() <= 1e-6 and (diff > 1e-4).float().mean().item() < 0.01
    # the reference-format file is a plain state dict with the reference's 359 keys: strict load into a fresh module
    sd = torch.load(ref_ck, map_location="cpu")
    assert len(sd) == 359
    cmgan_b200.TSCNet(64, 201).load_state_dict(sd, strict=True)


def test_async_pesq_training_steps(g_weights, d_weights, golden):
    """train_step_async: the discriminator update of batch n happens during step n + 1 with the targets the stand-in scorer gave batch n"""
    from cmgan_b200.pesq_pipeline import AsyncPesq
    m, d = _models(g_weights, d_weights)
    m.train(); d.train()
    t = FusedTrainer(m, d, lr=2e-4)
    clean = torch.from_numpy(golden["grad_clean"]).to(DEV)
    noisy = torch.from_numpy(golden["grad_noisy"]).to(DEV)
    seen = []

    def scorer(c, e):
        seen.append((c.shape, e.shape))
        return 1.0 + 3.5 * float(np.clip(1.0 - np.mean((c - e) ** 2) / (np.mean(c ** 2) + 1e-9), 0.0, 1.0))
    p = AsyncPesq(scorer=scorer, workers=2)
    pd0 = t.pd.clone()
    lg, ld = t.train_step_async(clean, noisy, p)
    assert ld is None and torch.equal(t.pd, pd0), "no discriminator update before the first batch has been scored"
    out = [t.train_step_async(clean, noisy, p) for _ in range(3)]
    assert all(np.isfinite(a.item()) and b is not None and np.isfinite(b.item()) for a, b in out)
    assert not torch.equal(t.pd, pd0)
    assert len(seen) >= 6 and seen[0][0] == seen[0][1] == (1600,)
    p.close()
