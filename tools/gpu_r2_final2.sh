#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
echo "== line + frame tests"; timeout -s KILL 900 python -m pytest tests/test_line_gpu.py tests/test_frame_gpu.py -m gpu -q 2>&1 | tail -2
echo "== bench"; timeout 1500 python bench.py --steps 10 2> gpurun_out/f.err > gpurun_out/r02b_bench_n1.json; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02b_bench_n1.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e'], d['cpu_baseline']['value'], d.get('cpu_baseline_cv2', {}).get('value')); print(d['roofline'])
PY
timeout 900 python bench.py --workload single1280 --steps 10 2>> gpurun_out/f.err > gpurun_out/r02b_bench_single1280.json; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02b_bench_single1280.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']); print(d['roofline'].get('stage_ms'))
PY
tail -3 gpurun_out/f.err | cut -c1-300
