"""Host-side logic of the multi-GPU path on CPU: 2 ranks over gloo (rendezvous on 127.0.0.1)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cmgan_b200 import parallel
    import cmgan_b200
    # flat-gradient averaging
    flat = torch.full((1000,), float(rank + 1))
    parallel.allreduce_mean_(flat)
    ok = bool(torch.allclose(flat, torch.full((1000,), 1.5)))
    # rank-0 parameters win
    torch.manual_seed(100 + rank)
    m = cmgan_b200.TSCNet(64, 201)
    parallel.broadcast_module(m, 0)
    ref = [torch.zeros_like(p) for p in m.parameters()]
    for r, p in zip(ref, m.parameters()):
        r.copy_(p.data)
        dist.broadcast(r, 0)
    ok = ok and all(torch.equal(r, p.data) for r, p in zip(ref, m.parameters()))
    # batch sharding
    x = torch.arange(8).view(8, 1)
    (xs,) = parallel.shard_batch([x], rank, world)
    ok = ok and xs.flatten().tolist() == list(range(rank * 4, rank * 4 + 4))
    ret[rank] = ok
    dist.destroy_process_group()


def test_two_rank_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
