#!/bin/bash
# One GPU-box call that regenerates the round's kernel-level evidence under gpurun_out/ (copied into profiles/ afterwards):
# steady-state microbenchmarks, an ncu --set full capture of the backward kernels, compute-sanitizer memcheck / racecheck logs.
set -u
O=gpurun_out
mkdir -p $O
timeout 200 python tools/bench_ffn.py > $O/ev_ffn.json 2> $O/ev_ffn.err; echo "ffn rc=$?"
timeout 200 python tools/bench_attn.py > $O/ev_attn.json 2> $O/ev_attn.err; echo "attn rc=$?"
timeout 120 python tools/bench_dwconv.py > $O/ev_dwconv.txt 2>&1; echo "dwconv rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"ffn_bwd|ln_bwd|glu_dwconv_bwd|gemm_wgrad_tma|attn_bwd_dq|attn_bwd_dkv" -c 12 -o $O/ev_bwd_kernels \
    python tools/profile_step.py --batch 4 > $O/ev_ncu.log 2>&1; echo "ncu rc=$?"
( timeout 420 compute-sanitizer --tool memcheck python -m pytest -q -x tests/test_gpu_trainmode.py tests/test_gpu_kernels.py tests/test_gpu_disc.py tests/test_module_abi.py \
    -m gpu -k "ffn or attention or dwconv or disc or c_entry or layer_norm or ln" 2>&1 | tail -6; echo "rc=$?" ) > $O/ev_memcheck.log 2>&1
( timeout 420 compute-sanitizer --tool racecheck python -m pytest -q -x tests/test_gpu_trainmode.py tests/test_gpu_kernels.py \
    -m gpu -k "(ffn and 300) or (attention and 130) or dwconv" 2>&1 | tail -12; echo "rc=$?" ) > $O/ev_racecheck.log 2>&1
tail -3 $O/ev_memcheck.log $O/ev_racecheck.log
