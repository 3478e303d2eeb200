"""Data parallelism over the GPUs of one box (ref: train.py:34-42,68-69 -- DDP over NCCL).

The path shards on the batch axis only: every rank runs the whole generator / discriminator on its own utterances
(BatchNorm statistics stay per rank, as in the reference, which does not use SyncBatchNorm) and the single exchange is
one all-reduce (average) of the flat gradient buffer the backward kernels accumulate into
(``TSCNet.enable_flat_grads()``): 7.34 MB for the generator, 0.73 MB for the discriminator, over NVLink/NVSwitch via NCCL.
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.distributed as dist


def broadcast_module(module: torch.nn.Module, src: int = 0) -> None:
    """rank ``src``'s parameters and buffers win (DDP constructor semantics, train.py:68-69)"""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


def allreduce_mean_(flat: torch.Tensor) -> torch.Tensor:
    """in-place average of a flat gradient buffer over all ranks (one collective per optimiser step)"""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return flat
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:                                   # gloo (CPU tests) has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
    return flat


def shard_batch(tensors: Iterable[torch.Tensor], rank: int, world: int):
    """contiguous shard of the leading (utterance) axis for this rank; the batch must divide evenly (drop_last=True in the reference)"""
    out = []
    for t in tensors:
        n = t.shape[0]
        assert n % world == 0, "global batch must be a multiple of the world size"
        per = n // world
        out.append(t[rank * per:(rank + 1) * per])
    return out
