"""Asynchronous PESQ targets for the discriminator step (reference: discriminator.py:9-26 ``batch_pesq`` called synchronously from
train.py:153-174 -- a device-to-host copy, a CPU job per utterance and a host-to-device copy on the critical path of every step).

Here the waveforms of step n leave the GPU on a side stream into one of two pinned staging buffers, a pool of host workers scores them
while the GPU runs step n + 1, and the discriminator update for the batch of step n happens one step late with exactly the targets the
reference would have used for that batch (same clean / enhanced pair, same (pesq - 1) / 3.5 mapping, whole batch skipped when any
utterance fails -- ``-1 in pesq_score`` at discriminator.py:23).  Nothing in the generator step ever waits for the host.

The scorer is pluggable: by default ``pesq.pesq(16000, clean, enhanced, "wb")`` from the ``pesq`` package when it is installed (it is
third-party C code and absent from this image: constructing the pipeline without a scorer then raises).
"""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, Optional

import numpy as np
import torch

Scorer = Callable[[np.ndarray, np.ndarray], float]          # (clean, enhanced) -> PESQ score, raises or returns -1 on failure


def default_scorer() -> Scorer:
    try:
        from pesq import pesq as _pesq
    except ImportError as e:      # the reference's dependency (pesq==0.0.3) is not vendored
        raise RuntimeError("the 'pesq' package is not installed: pass scorer=... to AsyncPesq") from e
    return lambda clean, est: _pesq(16000, clean, est, "wb")


class AsyncPesq:
    def __init__(self, scorer: Optional[Scorer] = None, workers: int = 8, slots: int = 2):
        self.scorer = scorer if scorer is not None else default_scorer()
        self.pool = ThreadPoolExecutor(max_workers=workers)       # the PESQ C code releases the GIL; numpy scorers mostly do too
        self.slots = slots
        self._host: Dict[int, tuple] = {}                         # slot -> (pinned clean, pinned est)
        self._pending: Dict[int, tuple] = {}                      # step -> (event or None, slot, B, L, future or None)
        self._done: Dict[int, Optional[torch.Tensor]] = {}        # step -> host targets (B,) or None (batch skipped)
        self._lock = threading.Lock()
        self._copy_stream = None

    # ------------------------------------------------------------------ producer side (called right after the generator step)
    def submit(self, step: int, clean: torch.Tensor, est_audio: torch.Tensor) -> None:
        """clean: (B, >= L) un-normalised clean waveforms, est_audio: (B, L) enhanced waveforms (train.py:154-157).  Returns immediately."""
        B, L = est_audio.shape
        slot = step % self.slots
        self._drain(block_slot=slot)                              # the staging buffers of this slot must be free again
        if slot not in self._host or self._host[slot][0].shape != (B, L):
            pin = est_audio.is_cuda
            self._host[slot] = (torch.empty(B, L, pin_memory=pin), torch.empty(B, L, pin_memory=pin))
        hc, he = self._host[slot]
        ev = None
        if est_audio.is_cuda:
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream()
            cur = torch.cuda.current_stream()
            self._copy_stream.wait_stream(cur)                    # the copies start when the step's kernels that wrote est_audio are done
            with torch.cuda.stream(self._copy_stream):
                hc.copy_(clean[:, :L], non_blocking=True)
                he.copy_(est_audio.detach(), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            clean.record_stream(self._copy_stream)
            est_audio.record_stream(self._copy_stream)
        else:
            hc.copy_(clean[:, :L])
            he.copy_(est_audio.detach())
        with self._lock:
            self._pending[step] = (ev, slot, B, L, None)
        self._launch_ready()

    def _score_batch(self, slot: int, B: int) -> Optional[torch.Tensor]:
        hc, he = self._host[slot]
        c, e = hc.numpy(), he.numpy()

        def one(i):
            try:
                return float(self.scorer(c[i], e[i]))
            except Exception:          # noqa: BLE001 -- "error can happen due to silent period" (discriminator.py:13)
                return -1.0
        scores = np.array(list(self.pool.map(one, range(B))), dtype=np.float64) if B > 1 else np.array([one(0)])
        if (scores == -1).any():
            return None
        return torch.from_numpy(((scores - 1.0) / 3.5).astype(np.float32))

    def _launch_ready(self) -> None:
        """start scoring every submitted batch whose device-to-host copy has finished (never blocks)"""
        with self._lock:
            items = list(self._pending.items())
        for step, (ev, slot, B, L, fut) in items:
            if fut is None and (ev is None or ev.query()):
                f = _Deferred(self._score_batch, slot, B)
                with self._lock:
                    self._pending[step] = (ev, slot, B, L, f)

    def _drain(self, block_slot: Optional[int] = None) -> None:
        """move finished batches to the results; with ``block_slot`` wait for the batch that still occupies that staging slot"""
        self._launch_ready()
        with self._lock:
            items = list(self._pending.items())
        for step, (ev, slot, B, L, fut) in items:
            must = block_slot is not None and slot == block_slot
            if fut is None:
                if not must:
                    continue
                if ev is not None:
                    ev.synchronize()
                fut = _Deferred(self._score_batch, slot, B)
            if must or fut.done():
                res = fut.result()
                with self._lock:
                    self._done[step] = res
                    self._pending.pop(step, None)

    # ------------------------------------------------------------------ consumer side (called before the discriminator step)
    def ready(self, step: int) -> bool:
        self._drain()
        return step in self._done

    def targets(self, step: int, device=None, wait: bool = False) -> Optional[torch.Tensor]:
        """(B,) tensor of (pesq - 1) / 3.5 for the batch submitted as ``step``, or None when that batch must be skipped (a PESQ call
        failed) or -- with ``wait`` False -- is not scored yet.  ``has_result(step)`` tells the two None cases apart."""
        self._drain()
        if step not in self._done and wait:
            with self._lock:
                ent = self._pending.get(step)
            if ent is not None:
                self._drain(block_slot=ent[1])
        res = self._done.get(step)
        if res is None:
            return None
        return res.to(device, non_blocking=True) if device is not None else res

    def has_result(self, step: int) -> bool:
        return step in self._done

    def forget(self, step: int) -> None:
        self._done.pop(step, None)

    def close(self) -> None:
        self.pool.shutdown(wait=True)


class _Deferred:
    """a batch being scored on a background thread (the per-utterance calls fan out over the shared pool)"""

    def __init__(self, fn, *args):
        self._res, self._exc = None, None
        self._t = threading.Thread(target=self._run, args=(fn, args), daemon=True)
        self._t.start()

    def _run(self, fn, args):
        try:
            self._res = fn(*args)
        except BaseException as e:      # noqa: BLE001
            self._exc = e

    def done(self) -> bool:
        return not self._t.is_alive()

    def result(self):
        self._t.join()
        if self._exc is not None:
            raise self._exc
        return self._res
