#!/bin/bash
# full GPU suite with the new automatic walker choice + bench A/B of the one-warp walker forms
cd /root/repo; mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full gpu suite"; timeout -s KILL 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for lean in 1 0; do echo "== bench lean=$lean"; SSLPL_WALKER_LEAN=$lean timeout 900 python bench.py --steps 10 --no-cpu-baseline 2> gpurun_out/w.err > gpurun_out/w_bench_$lean.json; python - <<PY
import json
d = json.loads(open('gpurun_out/w_bench_$lean.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']); print(d['roofline'].get('stage_ms'))
PY
done
