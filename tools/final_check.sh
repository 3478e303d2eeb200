#!/bin/bash
# Round-end validation on one GPU box: parity tests, smoke, both bench arms, memcheck over the whole GPU suite, racecheck on the kernels
# whose synchronisation it can model.
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 400 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/final_bench_ref.json 2> $O/final_bench_ref.err; echo "ref rc=$?"; tail -c 400 $O/final_bench_ref.json; echo
( timeout 500 compute-sanitizer --tool memcheck python -m pytest -q tests -m gpu -k "not audiosamples and not tscnet_train_mode and not 129684" 2>&1 | tail -8; echo "rc=$?" ) > $O/final_memcheck.log 2>&1
( timeout 300 compute-sanitizer --tool racecheck python -m pytest -q tests/test_gpu_kernels.py tests/test_gpu_trainmode.py -m gpu -k "dwconv or (attention and 130 and tf32) or layer_norm or narrow or disc" 2>&1 | grep -v "^$" | tail -12; echo "rc=$?" ) > $O/final_racecheck_simt.log 2>&1
tail -4 $O/final_memcheck.log $O/final_racecheck_simt.log
