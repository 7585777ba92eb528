#!/bin/bash
# final verification of the round: what the driver runs, on the final code
cd /root/repo; mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== full gpu suite"; timeout -s KILL 1700 python -m pytest tests -m gpu -q 2>&1 | tail -2
echo "== bench"; timeout 1500 python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/f.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02b_bench_n1.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value', 'ms_per_step', 'steps', 'warmup', 'gpu_launches')}, d['e2e'], d['cpu_baseline']['value'], d['clocks'])
PY
echo "== reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/r02b_bench_reference_arm.json 2>> gpurun_out/f.err; tail -c 400 gpurun_out/r02b_bench_reference_arm.json; echo
echo "== v3 / lanes extra stress on other sizes"; SSLPL_WALKER_V3=1 SSLPL_WALKER_WARPS=8 timeout 600 python -m pytest tests/test_line_gpu.py -m gpu -q -k "not every_region" 2>&1 | tail -1
SSLPL_WALKER_LANES=1 SSLPL_WALKER_WARPS=-1 timeout 600 python -m pytest tests/test_line_gpu.py -m gpu -q -k "not every_region" 2>&1 | tail -1
