timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | grep -v Warning | tail -3
timeout 400 python bench.py --no-extras > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r3j_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
