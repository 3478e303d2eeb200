"""Data-parallel path on real GPUs (needs >= 2 devices; skipped otherwise): two ranks x B = 2 with one NCCL all-reduce (AVG) of the flat
gradient buffers must give the gradients one GPU computes on the B = 4 batch (BatchNorm in eval mode and no dropout, so that the only
difference is the sharding), for the generator (two overlapped segment all-reduces) and for the discriminator (its own all-reduce).
ref: train.py:68-69,192,200 (DDP gradient averaging)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, gpath, dpath, out):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import cmgan_b200
    import cmgan_b200.discriminator as Dm
    from cmgan_b200 import ops
    from cmgan_b200.trainer import FusedTrainer
    from oracle import cmgan_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ops.set_precision("tf32")
        Dm.DROP_P = 0.0
        m = cmgan_b200.TSCNet(64, 201)
        m.load_state_dict(O.load_weights_npz(gpath), strict=True)
        d = cmgan_b200.Discriminator(16)
        d.load_state_dict(O.load_weights_npz(dpath), strict=True)
        m, d = m.to(dev).eval(), d.to(dev).train()        # D in train mode (power iterations: with the un-iterated u / v of a fresh init the
        # learnable sigmoid saturates and every D gradient is exactly zero); dropout off, so the ranks stay comparable
        t = FusedTrainer(m, d)
        g = torch.Generator().manual_seed(0)
        clean = 0.05 * torch.randn(4, 8000, generator=g)
        noisy = clean + 0.05 * torch.randn(4, 8000, generator=g)
        per = 4 // world
        c, n = clean[rank * per:(rank + 1) * per].to(dev), noisy[rank * per:(rank + 1) * per].to(dev)
        lg = t.generator_step(c, n, update=False)
        ld = t.discriminator_step(torch.full((per,), 0.5, device=dev), update=False)
        torch.cuda.synchronize()
        if rank == 0:
            torch.save(dict(gg=t.gg.cpu(), gd=t.gd.cpu(), lg=lg.cpu(), ld=ld.cpu()), out)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_gradients_equal_single_gpu_batch(tmp_path):
    import torch.multiprocessing as mp
    gpath, dpath = os.path.join(ROOT, "tests", "golden", "weights_g.npz"), os.path.join(ROOT, "tests", "golden", "weights_d.npz")
    out2, out1 = str(tmp_path / "two.pt"), str(tmp_path / "one.pt")
    mp.spawn(_worker, args=(2, _free_port(), gpath, dpath, out2), nprocs=2, join=True)
    mp.spawn(_worker, args=(1, _free_port(), gpath, dpath, out1), nprocs=1, join=True)
    a, b = torch.load(out2), torch.load(out1)
    # the time-domain L1 and the MSE terms are batch means: the average of two half-batch gradients is the full-batch gradient; InstanceNorm is
    # per utterance, BatchNorm runs on running statistics (eval), the spectral-norm power iteration is identical on both ranks
    for key, tol in (("gg", 2e-2), ("gd", 2e-2)):
        ref, got = b[key].double(), a[key].double()
        assert ref.abs().max().item() > 0, f"{key}: reference gradient is identically zero"
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        print(f"[dp2] {key}: 2 x B=2 (all-reduce AVG) vs 1 x B=4: max-abs deviation / max {err:.3e}")
        assert np.isfinite(err) and err < tol, key


def _ddp_worker(rank, world, port, gpath, out):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    import cmgan_b200
    from cmgan_b200 import training
    from oracle import cmgan_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        g = torch.Generator().manual_seed(rank)
        clean = (0.05 * torch.randn(2, 8000, generator=g)).to(dev)
        noisy = clean + (0.05 * torch.randn(2, 8000, generator=g)).to(dev)
        grads = []
        for wrap in (False, True):
            m = cmgan_b200.TSCNet(64, 201)
            m.load_state_dict(O.load_weights_npz(gpath), strict=True)
            m = m.to(dev).eval()                                      # eval: no dropout, so both passes see the same function
            model = DDP(m, device_ids=[rank]) if wrap else m          # train.py:68: DDP(self.model, device_ids=[gpu_id])
            outp = training.forward_generator_step(model, clean, noisy)
            loss = training.generator_loss(outp, clean)
            loss.backward()
            torch.cuda.synchronize()
            grads.append(torch.cat([p.grad.reshape(-1) for p in m.parameters()]).cpu())
            assert set(model.state_dict().keys()) == {("module." + k if wrap else k) for k in m.state_dict().keys()}     # train.py:273 saves model.module
        if rank == 0:
            torch.save(dict(plain=grads[0], ddp=grads[1]), out)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")
def test_module_inside_ddp(tmp_path):
    """the reference wraps TSCNet in DistributedDataParallel (train.py:68-69): parameters must be leaf tensors DDP can hook, the forward must
    run through the wrapper, and with one rank the averaged gradients equal the unwrapped module's"""
    import torch.multiprocessing as mp
    gpath = os.path.join(ROOT, "tests", "golden", "weights_g.npz")
    out = str(tmp_path / "ddp.pt")
    mp.spawn(_ddp_worker, args=(1, _free_port(), gpath, out), nprocs=1, join=True)
    r = torch.load(out)
    ref, got = r["plain"].double(), r["ddp"].double()
    assert ref.abs().max().item() > 0
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    print(f"[ddp] DDP-wrapped vs plain module gradients: {err:.3e}")
    assert err < 1e-5
