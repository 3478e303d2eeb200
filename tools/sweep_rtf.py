#!/usr/bin/env python
"""inference real-time factor over clip length x batch.  Each point = the whole enhancement path
(RMS normalise -> STFT -> compress -> TSCNet eval forward -> uncompress -> iSTFT -> de-normalise; ref: evaluation.py:12-58) captured as one
CUDA graph and replayed; RTF = GPU time / (batch x clip seconds).  Writes a markdown table (stdout) and a JSON list (--json)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import cmgan_b200  # noqa: E402
from cmgan_b200 import ops, signal  # noqa: E402


def time_graph(fn, reps):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    del g
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", default="1,2,4,8")
    ap.add_argument("--batches", default="1,4,16,64")
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    ops.set_precision(args.precision)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = cmgan_b200.TSCNet(64, 201).to(dev).eval()
    rows = []
    for sec in [int(s) for s in args.seconds.split(",")]:
        for B in [int(b) for b in args.batches.split(",")]:
            L = 16000 * sec
            if B * (L // 100 + 1) * 201 * 320 >= 2 ** 31:       # the encoder's (M, 320) concat buffer is indexed with 32-bit element counts
                rows.append(dict(seconds=sec, T=L // 100 + 1, batch=B, ms=None, note="skipped: > 2^31 elements in one buffer"))
                continue
            wav = 0.05 * torch.randn(B, L, device=dev)
            try:
                with torch.no_grad():
                    ms = time_graph(lambda: signal.enhance_batch(model, wav), args.reps)
            except torch.OutOfMemoryError:
                rows.append(dict(seconds=sec, T=L // 100 + 1, batch=B, ms=None))
                torch.cuda.empty_cache()
                continue
            rows.append(dict(seconds=sec, T=L // 100 + 1, batch=B, ms=ms, rtf=ms * 1e-3 / (B * sec), utt_per_s=B / (ms * 1e-3),
                             audio_s_per_s=B * sec / (ms * 1e-3)))
            del wav
            torch.cuda.empty_cache()
    print(f"| clip | T | batch | ms | RTF | x real time |\n|---|---:|---:|---:|---:|---:|")
    for r in rows:
        if r["ms"] is None:
            print(f"| {r['seconds']} s | {r['T']} | {r['batch']} | {r.get('note', 'out of memory')} | | |")
        else:
            print(f"| {r['seconds']} s | {r['T']} | {r['batch']} | {r['ms']:.2f} | {r['rtf']:.5f} | {r['audio_s_per_s']:.0f} |")
    if args.json:
        with open(args.json, "w") as fh:
            json.dump(dict(precision=args.precision, what="enhancement path as one CUDA graph, eval mode, random-init weights, synthetic noise",
                           points=rows), fh)


if __name__ == "__main__":
    main()
