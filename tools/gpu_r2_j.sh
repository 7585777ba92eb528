#!/bin/bash
# row 3 (line projection matchers, Fuse) on the GPU + full suite + bench exit check
cd /root/repo; mkdir -p gpurun_out
echo "== row3 tests"; timeout -s KILL 900 python -m pytest tests/test_fuse_gpu.py tests/test_ref_golden_gpu.py tests/test_integration_gpu.py -q -x 2>&1 | tail -15
echo "== full gpu suite"; timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "== bench"; timeout 1200 python bench.py --no-cpu-baseline --steps 10 2> gpurun_out/j.err > gpurun_out/j_bench.json; echo "rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/j_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['ms_per_step'], d['roofline']['stage_ms'])
PY
tail -5 gpurun_out/j.err | cut -c1-300
