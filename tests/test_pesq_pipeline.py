"""Host-side logic of the asynchronous PESQ target pipeline (cmgan_b200.pesq_pipeline) with a stand-in scorer: non-blocking submit,
per-batch results keyed by step, the reference's (score - 1) / 3.5 mapping and whole-batch skip on a failed utterance
(discriminator.py:9-26), staging-slot reuse."""
import time

import numpy as np
import torch

from cmgan_b200.pesq_pipeline import AsyncPesq


def _scorer(delay=0.0, fail_on=None):
    def f(clean, est):
        if delay:
            time.sleep(delay)
        if fail_on is not None and abs(float(clean[0]) - fail_on) < 1e-6:
            raise ValueError("silent period")
        err = float(np.mean((clean - est) ** 2))
        return 1.0 + 3.5 / (1.0 + err)            # in (1, 4.5]
    return f


def test_async_targets_match_synchronous_scoring_and_do_not_block():
    p = AsyncPesq(scorer=_scorer(delay=0.05), workers=4)
    g = torch.Generator().manual_seed(0)
    batches = []
    for step in range(4):
        clean = torch.randn(4, 1000, generator=g)
        est = clean[:, :900] + 0.1 * step * torch.randn(4, 900, generator=g)
        t0 = time.perf_counter()
        p.submit(step, clean, est)
        dt = time.perf_counter() - t0
        if step < 2:
            assert dt < 0.04, f"submit blocked for {dt:.3f} s"       # both staging slots free: no waiting on the scorer
        batches.append((clean, est))
    for step, (clean, est) in enumerate(batches):
        tgt = p.targets(step, wait=True)
        ref = torch.tensor([( _scorer()(clean[i, :900].numpy(), est[i].numpy()) - 1.0) / 3.5 for i in range(4)])
        assert tgt is not None and torch.allclose(tgt, ref.float(), atol=1e-6), step
    p.close()


def test_failed_utterance_skips_the_whole_batch():
    p = AsyncPesq(scorer=_scorer(fail_on=123.0), workers=2)
    clean = torch.randn(3, 500)
    clean[1, 0] = 123.0
    p.submit(7, clean, clean.clone())
    assert p.targets(7, wait=True) is None and p.has_result(7)          # scored, and the verdict is "skip" (train.py:161,171-172)
    p.submit(8, torch.randn(3, 500), torch.randn(3, 500))
    assert p.targets(8, wait=True) is not None
    p.close()


def test_not_ready_is_distinguishable_from_skip():
    p = AsyncPesq(scorer=_scorer(delay=0.2), workers=1)
    p.submit(0, torch.randn(2, 100), torch.randn(2, 100))
    assert p.targets(0) is None and not p.has_result(0)
    assert p.targets(0, wait=True) is not None and p.has_result(0)
    p.close()
