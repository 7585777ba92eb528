#!/bin/bash
cd /root/repo
echo "== full gpu suite (default)"; timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== line suite, solo kernel"; SSLPL_WALKER_WARPS=-1 timeout -s KILL 900 python -m pytest tests/test_line_gpu.py tests/test_ref_golden_gpu.py -x -q 2>&1 | tail -3
echo "== line suite, 5 warps"; SSLPL_WALKER_WARPS=5 timeout -s KILL 900 python -m pytest tests/test_line_gpu.py tests/test_ref_golden_gpu.py -x -q 2>&1 | tail -3
echo "== defaults"; timeout -s KILL 300 python tools/walker_timing.py 2>&1 | grep -v "turn_regions" | tail -4
echo "== solo"; SSLPL_WALKER_WARPS=-1 timeout -s KILL 300 python tools/walker_timing.py 2>&1 | grep -v "turn_regions" | tail -4
echo "== 4 warps"; SSLPL_WALKER_WARPS=4 timeout -s KILL 300 python tools/walker_timing.py 2>&1 | grep -v "turn_regions" | tail -1
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_walker.json 2> gpurun_out/r02_bench_walker.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_walker.json').read().strip().splitlines()[-1])
print({k:v for k,v in d.items() if k in('value','ms_per_step','e2e')}); print(d['roofline']['stage_ms'])
PY
tail -2 gpurun_out/r02_bench_walker.err | cut -c1-300
