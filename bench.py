#!/usr/bin/env python
"""bench.py -- headline benchmark of the CMGAN hot path on B200 (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json metric: utterances/sec, 2 s @ 16 kHz, generator forward+backward): per rank, one step =
RMS normalise -> STFT -> power compression -> TSCNet forward (train mode: dropout, BatchNorm batch statistics) ->
un-compression -> iSTFT -> generator loss (RI + magnitude + time terms) -> backward through all of it into the flat
gradient buffer (+ one NCCL all-reduce of that buffer when N > 1), on a batch of B = 4 synthetic 2 s clips (train.py's
default batch size and configs[1]'s batch).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line.  Libraries print there too (NCCL's version banner), so file descriptor 1 is pointed at
# stderr for the whole run and the JSON line is written to the saved original descriptor by emit().
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_STDOUT_FD = os.dup(1)
os.dup2(2, 1)


def emit(obj) -> None:
    os.write(_STDOUT_FD, (json.dumps(obj) + "\n").encode())

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "utterances/sec (2 s @16 kHz) generator fwd+bwd"
UNIT = "utt/s"
CLIP = 32000
FWD_GFLOP_PER_UTT = 145.96            # mm + bmm + conv, 2*MAC, 2 s clip
STEP_GFLOP_PER_UTT = 3 * FWD_GFLOP_PER_UTT


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def synth_batch(B, seed, device=None, pin=False):
    import torch
    g = torch.Generator().manual_seed(seed)
    clean = 0.05 * torch.randn(B, CLIP, generator=g)
    noisy = clean + 0.05 * torch.randn(B, CLIP, generator=g)
    if pin:
        return clean.pin_memory(), noisy.pin_memory()
    if device is not None:
        return clean.to(device), noisy.to(device)
    return clean, noisy


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_step(sd, clean, noisy):
    """the oracle's (= the reference algorithm's) generator forward+backward on CPU, same loss as the GPU arm"""
    import torch
    import torch.nn.functional as F
    from oracle import cmgan_oracle as O
    for v in sd.values():
        if v.is_floating_point() and v.grad is not None:
            v.grad = None
    go = O.forward_generator_step(clean, noisy, sd, training=True)    # BatchNorm batch statistics as in train mode
    loss = 0.1 * (F.mse_loss(go["est_real"], go["clean_real"]) + F.mse_loss(go["est_imag"], go["clean_imag"])) \
        + 0.9 * F.mse_loss(go["est_mag"], go["clean_mag"]) + 0.2 * torch.mean(torch.abs(go["est_audio"] - clean))
    loss.backward()
    return loss.item()


def cpu_weights():
    import torch
    from oracle import cmgan_oracle as O
    w = O.load_weights_npz(os.path.join(ROOT, "tests", "golden", "weights_g.npz"))
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v) for k, v in w.items()}


def host_threads():
    """threads for the CPU arm: the cores this process may run on, capped at 32 (the reference's ~700 small ATen ops per
    forward stop scaling well before that; 128 threads measured 17x slower than 8 on the GPU box's host)"""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, 32))


def time_cpu_baseline(n_steps=1, warm=0):
    import torch
    torch.set_num_threads(host_threads())
    sd = cpu_weights()
    clean, noisy = synth_batch(1, 123)
    for _ in range(warm):
        cpu_reference_step(sd, clean, noisy)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        cpu_reference_step(sd, clean, noisy)
    dt = (time.perf_counter() - t0) / n_steps
    return dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = host_threads()
    torch.set_num_threads(cores)
    sd = cpu_weights()
    clean, noisy = synth_batch(1, 123)
    for _ in range(args.warmup):
        cpu_reference_step(sd, clean, noisy)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(sd, clean, noisy)
    dt = (time.perf_counter() - t0) / args.steps
    v = 1.0 / dt
    sample = "1 utterance (B=1 x 2 s) per step of the B=4 workload, oracle CPU port of the reference forward+backward, fp32, torch CPU threads = cores"
    emit(({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch), "reference_sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(B):
    return (f"generator training step fwd+bwd+AdamW (train mode: dropout + BatchNorm batch stats), "
            f"stft->compress->TSCNet->uncompress->istft->loss->backward->update, batch {B} x 2 s @16 kHz per GPU, fp32 storage")


# ------------------------------------------------------------------------------------------------ our arm (GPU)
def run_ours(args):
    import torch
    import torch.distributed as dist
    import cmgan_b200
    from cmgan_b200 import ops as _ops
    _ops.set_precision(args.precision)
    from cmgan_b200 import ops, training
    from cmgan_b200.ops import call

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    B = args.batch
    torch.manual_seed(0)
    model = cmgan_b200.TSCNet(64, 201).to(dev).train()
    from cmgan_b200.trainer import FusedTrainer
    from cmgan_b200 import parallel
    trainer = FusedTrainer(model, None)          # flat parameter/gradient buffers; rank-0 parameters win (train.py:68)
    flat = trainer.gg
    clean, noisy = synth_batch(B, 1000 + rank, device=dev)
    hclean, hnoisy = synth_batch(B, 1000 + rank, pin=True)

    def step(c, n):          # eager: every kernel launched from Python (gradient all-reduce when N > 1, AdamW update)
        return trainer.generator_step(c, n, update=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, K):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(K):
            fn()
        host_ms[0] = (time.perf_counter() - t0) * 1e3 / K       # host time to enqueue one step
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(max(args.warmup, 3)):
        step(clean, noisy)
    ms_eager = timed(lambda: step(clean, noisy), args.steps)
    eager_host_ms = host_ms[0]
    # ---- the whole step as one CUDA graph (the gradient all-reduce stays an eager NCCL call when N > 1)
    l0 = ops.LAUNCHES
    # N = 1: forward, losses, backward and the AdamW update are all inside the graph.  N > 1: the graph ends with the backward pass,
    # then the NCCL all-reduce of the flat gradient buffer and the (single-kernel) AdamW update are issued eagerly.
    trainer.capture_generator_step(clean, noisy, update=(world == 1), allreduce=False)
    launches_per_step = (ops.LAUNCHES - l0) // 3 + (1 if world > 1 else 0)      # 2 warm-up passes + 1 capture pass

    def gstep(c, n):
        loss = trainer.replay_generator_step(c, n)
        if world > 1:
            parallel.allreduce_mean_(flat)
            trainer.opt_g.step()
        return loss

    for _ in range(3):
        gstep(clean, noisy)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: gstep(clean, noisy), args.steps)
    loss_after = float(gstep(clean, noisy).item())
    assert loss_after == loss_after and abs(loss_after) < 1e30, f"training step diverged: loss {loss_after}"
    launches = launches_per_step * args.steps
    host_enqueue_ms = host_ms[0]
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end: pinned host buffers in, loss scalar out, every step
    host_loss = torch.empty(1).pin_memory()

    def e2e_step():
        loss = gstep(hclean, hnoisy)                     # H2D copies of the pinned batch into the graph's input buffers
        host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the caller reads the loss every step (train.py:205)
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)

    # ---- forward only (configs[1]: eval forward, batch 4), also as a CUDA graph
    model.eval()
    with torch.no_grad():
        def fwd_only():
            go = training.forward_generator_step(model, clean, noisy)
            return go["est_audio"]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                fwd_only()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        fgraph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(fgraph):
            fwd_out = fwd_only()
        for _ in range(2):
            fgraph.replay()
        ms_f = timed(fgraph.replay, args.steps)
    fwd_value = world * B * args.steps / (ms_f * 1e-3)
    model.train()

    # ---- dominant kernel family: every GEMM launch of one step is recorded (arguments + operands kept alive) and the whole list is
    # replayed back to back between two CUDA events, so the durations carry no host gaps (an eager step is host-bound)
    ops.PROBE = []
    step(clean, noisy)
    torch.cuda.synchronize()
    probe, ops.PROBE = ops.PROBE, None
    import ctypes as _ct
    from cmgan_b200._lib import lib as _lib

    def replay(entries, reps=3):
        L, st = _lib(), ops.stream()
        def once():
            for p in entries:
                L.call(p[0], _ct.byref(p[5]), st)
        once()
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(reps):
            once()
        r1.record()
        torch.cuda.synchronize()
        return r0.elapsed_time(r1) / reps * 1e-3

    rows = [p for p in probe if p[0] == "cmgan_gemm_rows_f32"]
    wgs = [p for p in probe if p[0] == "cmgan_gemm_wgrad_f32"]
    t_rows, t_wg = replay(rows), replay(wgs)
    if os.environ.get("CMGAN_PROBE_DUMP") and rank == 0:      # per-shape replay timings for analysis (not part of the JSON line)
        shapes = {}
        for p in probe:
            shapes.setdefault((p[0], p[1], p[2], p[3], p[4]), []).append(p)
        dump = [[k[0], k[1], k[2], k[3], replay(v, 2) / len(v) * 1e6, k[4], len(v)] for k, v in shapes.items()]
        with open(os.environ["CMGAN_PROBE_DUMP"], "w") as fh:
            json.dump(dump, fh)
    f_rows = sum(2.0 * p[1] * p[2] * p[3] for p in rows)
    f_wg = sum(2.0 * p[1] * p[2] * p[3] for p in wgs)
    b_rows = float(sum(p[4] for p in rows))
    b_wg = float(sum(p[4] for p in wgs))
    del probe
    peaks, psrc = _peaks()
    hbm_peak = peaks.get("hbm_gbs", 6500.0)
    tf32_peak = peaks.get("bf16_tflops_sustained", 1400.0) / 2.0      # dense tf32 = half the bf16 rate on the same tensor pipe
    step_s = ms / args.steps * 1e-3
    achieved = b_rows / t_rows / 1e9 if t_rows > 0 else 0.0
    # The dominant kernel family is the row-parallel GEMM (tcgen05 tf32): K = 64 .. 256 against N = 64 .. 256 is 13 - 64 flop/byte,
    # far below the ~110 flop/byte balance point of tf32 tensor cores vs HBM, so the family is HBM-bound and is reported as such.
    roofline = {"bound": "hbm", "kernel": "gemm_rows_tc_kernel (every dense contraction of the step: linear, pointwise, dilated/strided conv)",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "peak_source": f"{psrc}: hbm_gbs (copy bandwidth, read + write)",
                "how": "sum of algorithmic bytes (A once, C, epilogue operands, weights) over every launch of one step / CUDA-event time of those "
                       "launches replayed back to back (each includes its ~3 us weight re-tiling kernel)",
                "launches_per_step": len(rows), "share_of_step": t_rows / step_s, "algorithmic_gb_per_step": b_rows / 1e9,
                "tensor": {"achieved_tflops": f_rows / t_rows / 1e12 if t_rows > 0 else 0.0, "peak_tflops": tf32_peak,
                           "algorithmic_gflop_per_step": f_rows / 1e9},
                "wgrad": {"achieved": (b_wg / t_wg / 1e9) if t_wg > 0 else 0.0, "unit": "GB/s", "frac": (b_wg / t_wg / 1e9 / hbm_peak) if t_wg > 0 else 0.0,
                          "share_of_step": t_wg / step_s, "achieved_tflops": (f_wg / t_wg / 1e12) if t_wg > 0 else 0.0},
                "traffic": 239.3e6,
                "traffic_note": "dram read + write of the FFN-1 launch ((129684 x 64) x (64 x 256), Swish dual output; algorithmic 298.8 MB) from "
                                "ncu --set full (profiles/ncu_r1_kernels.md): 33.4 MB + 205.9 MB -- part of the output is still in the 126 MB L2 "
                                "when the kernel ends"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": workload_name(B), "global_batch": B * world, "clip_samples": CLIP, "parallelism": f"dp{world}",
                       "l2": "per-step working set (activations saved for backward, several GB) >> 126 MB L2; no explicit flush",
                       "weights": "torch.manual_seed(0) default init, updated by AdamW every step (lr 5e-4)", "loss_after": loss_after,
                       "launch": "one CUDA graph per step (cmgan_b200.trainer.FusedTrainer): forward, losses, backward, AdamW"
                                 + (" -- all-reduce + AdamW eager after the graph" if world > 1 else "")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * B * CLIP * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms,
            "eager": {"value": world * B * args.steps / (ms_eager * 1e-3), "unit": UNIT, "ms_per_step": ms_eager / args.steps,
                      "host_enqueue_ms_per_step": eager_host_ms, "note": "same step launched kernel by kernel from Python (no CUDA graph)"},
            "clocks": clocks,
            "roofline": roofline,
            "forward_only": {"value": fwd_value, "unit": UNIT, "ms_per_step": ms_f / args.steps, "workload": f"configs[1]: eval forward, batch {B} x 2 s"},
            "model_tflops": value * STEP_GFLOP_PER_UTT / 1e3,
        }
        if world == 1 and not args.no_cpu:
            cores = host_threads()
            dt = time_cpu_baseline(1, 0)
            out["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": "1 step of B=1 x 2 s (same loss) through the oracle CPU port of the reference, fp32, all host threads"}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"], help="dense contractions: tcgen05 tf32 (default) or exact fp32 FFMA")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
