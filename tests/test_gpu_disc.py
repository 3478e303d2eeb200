"""GPU parity of the metric discriminator (forward, spectral-norm power iteration, backward) vs the oracle / fixtures."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
if torch.cuda.is_available():
    import cmgan_b200
    from cmgan_b200 import ops
    from cmgan_b200.ops import call
from oracle import cmgan_oracle as O


def _chk(got, ref, tol, name=""):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    assert got.shape == ref.shape, f"{name}: {tuple(got.shape)} vs {tuple(ref.shape)}"
    err = (got - ref).abs().max().item()
    den = max(ref.abs().max().item(), 1e-30)
    print(f"[parity] {name}: max-abs {err:.3e} (ref max {den:.3e})")
    assert np.isfinite(err) and err <= tol * max(den, 1e-3), name


def _model(d_weights):
    m = cmgan_b200.Discriminator(ndf=16)
    m.load_state_dict(d_weights, strict=True)
    return m.to(DEV)


def test_disc_eval_and_train_forward(golden, d_weights):
    x, y = torch.from_numpy(golden["d_x"]).to(DEV), torch.from_numpy(golden["d_y"]).to(DEV)
    m = _model(d_weights).eval()
    with torch.no_grad():
        _chk(m(x, y), torch.from_numpy(golden["d_eval_out"]), 1e-5, "D eval vs reference fixture")
    m.train()
    import cmgan_b200.discriminator as D
    old = D.DROP_P
    D.DROP_P = 0.0            # the fixture was generated with Dropout disabled
    try:
        with torch.no_grad():
            out = m(x, y)
    finally:
        D.DROP_P = old
    _chk(out, torch.from_numpy(golden["d_train_out"]), 1e-5, "D train (1 power iteration) vs reference fixture")
    sd = m.state_dict()
    for li in (0, 3, 6, 9, 14, 17):
        _chk(sd[f"layers.{li}.weight_u"], torch.from_numpy(golden[f"d_train_u{li}"]), 1e-5, f"u{li}")
        _chk(sd[f"layers.{li}.weight_v"], torch.from_numpy(golden[f"d_train_v{li}"]), 1e-5, f"v{li}")


def test_disc_backward_with_dropout(golden, d_weights):
    x = torch.from_numpy(golden["d_x"])
    y = torch.from_numpy(golden["d_y"])
    m = _model(d_weights).train()
    m.seed, m._step = 3, 0
    seed = (3 * 7919 + 1) * 31 + 5
    thr, inv = ops.drop_params(0.3)
    mask = torch.empty(2 * 64, device=DEV)
    call("cmgan_dropout_mask", mask, 2 * 64, seed, thr)
    mask = mask.view(2, 64).cpu()
    xd, yd = x.to(DEV).requires_grad_(True), y.to(DEV).requires_grad_(True)
    out = m(xd, yd)
    tgt = torch.tensor([0.3, 0.9], device=DEV)
    ((out.flatten() - tgt) ** 2).mean().backward()
    # oracle (float64, same dropout mask, same power iteration)
    sd = {k: (v.double().requires_grad_(True) if v.is_floating_point() and not k.endswith(("_u", "_v")) else v.double() if v.is_floating_point() else v)
          for k, v in d_weights.items()}
    x64, y64 = x.double().requires_grad_(True), y.double().requires_grad_(True)
    ref = O.discriminator_forward(x64, y64, sd, training=True, drop_mask=mask.double())
    ((ref.flatten() - tgt.cpu().double()) ** 2).mean().backward()
    _chk(out, ref, 1e-5, "D train forward with dropout")
    _chk(xd.grad, x64.grad, 2e-4, "D dx")
    _chk(yd.grad, y64.grad, 2e-4, "D dy")
    gmax = max(v.grad.abs().max().item() for v in sd.values() if getattr(v, "grad", None) is not None)
    for k, p in m.named_parameters():
        ref_g = sd[k].grad
        err = (p.grad.double().cpu() - ref_g).abs().max().item() / max(ref_g.abs().max().item(), 1e-3 * gmax)
        print(f"[parity] D grad {k}: rel {err:.3e}")
        assert err < 2e-3, k


def test_two_train_forwards_then_backward_use_their_own_uv(d_weights, golden):
    """discriminator step order (train.py:162-170): forward(clean, est), forward(clean, clean) -- a second power iteration -- then
    the backward of both.  torch's spectral_norm keeps the (u, v) of each forward for its own backward; so must this path."""
    import cmgan_b200.discriminator as D
    x = torch.from_numpy(golden["d_x"])
    y = torch.from_numpy(golden["d_y"])
    m = _model(d_weights).train()
    old, D.DROP_P = D.DROP_P, 0.0
    try:
        P = m._tensor_dict()
        G = {k: torch.zeros_like(v) for k, v in m.named_parameters()}
        s1, s2 = {}, {}
        o1 = D.disc_fwd(x.to(DEV), y.to(DEV), P, True, 1, s1)
        o2 = D.disc_fwd(x.to(DEV), x.to(DEV), P, True, 2, s2)
        g1, g2 = torch.tensor([[0.7], [-0.3]], device=DEV), torch.tensor([[0.2], [0.5]], device=DEV)
        D.disc_bwd(s1, g1, P, G, False, False)
        D.disc_bwd(s2, g2, P, G, False, False)
    finally:
        D.DROP_P = old
    sd = {k: (v.double().requires_grad_(True) if v.is_floating_point() and not k.endswith(("_u", "_v")) else v.double() if v.is_floating_point() else v)
          for k, v in d_weights.items()}
    uv = {}
    r1 = O.discriminator_forward(x.double(), y.double(), sd, training=True, uv_out=uv)
    sd2 = dict(sd)
    for li, (u, v) in uv.items():
        sd2[f"layers.{li}.weight_u"], sd2[f"layers.{li}.weight_v"] = u, v
    r2 = O.discriminator_forward(x.double(), x.double(), sd2, training=True)
    ((r1 * g1.cpu().double()).sum() + (r2 * g2.cpu().double()).sum()).backward()
    _chk(o1, r1, 1e-5, "first train forward")
    _chk(o2, r2, 1e-5, "second train forward (second power iteration)")
    gmax = max(v.grad.abs().max().item() for v in sd.values() if getattr(v, "grad", None) is not None)
    for k in G:
        err = (G[k].double().cpu() - sd[k].grad).abs().max().item() / max(sd[k].grad.abs().max().item(), 1e-3 * gmax)
        print(f"[parity] D two-forward grad {k}: rel {err:.3e}")
        assert err < 2e-3, k
