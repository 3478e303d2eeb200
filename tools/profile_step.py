#!/usr/bin/env python
"""One profiled hot-path step for ncu (run under `ncu --profile-from-start off ...`): the same step bench.py times
(B x 2 s, generator forward+backward, train mode).  Warm-up steps run outside the cudaProfilerStart/Stop window."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
import cmgan_b200  # noqa: E402
from cmgan_b200 import training  # noqa: E402
from cmgan_b200.ops import call  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--fwd-only", action="store_true")
ap.add_argument("--gd", action="store_true", help="bench.py's default workload: generator step + discriminator step with both AdamW updates")
ap.add_argument("--precision", default="tf32")
args = ap.parse_args()
dev = torch.device("cuda", 0)
from cmgan_b200 import ops as _ops  # noqa: E402
_ops.set_precision(args.precision)
torch.manual_seed(0)
model = cmgan_b200.TSCNet(64, 201).to(dev).train()
from cmgan_b200.trainer import FusedTrainer  # noqa: E402
disc = cmgan_b200.Discriminator(16).to(dev).train() if args.gd else None
trainer = FusedTrainer(model, disc)
pesq_t = torch.full((args.batch,), 0.5, device=dev)
clean, noisy = bench.synth_batch(args.batch, 1000, device=dev)


def step():
    if args.fwd_only:
        with torch.no_grad():
            training.forward_generator_step(model, clean, noisy)
        return
    if args.gd:                      # the step bench.py captures into its CUDA graph (configs[2])
        trainer.generator_step(clean, noisy)
        trainer.discriminator_step(pesq_t)
        return
    trainer.generator_step(clean, noisy, update=False, allreduce=False)


for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
