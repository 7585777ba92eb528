#!/bin/bash
# fused FAST parity + lean one-warp walker experiment (register pressure on co-running kernels)
cd /root/repo; mkdir -p gpurun_out
echo "== orb/golden/frame tests"; timeout -s KILL 900 python -m pytest tests/test_orb_gpu.py tests/test_ref_golden_gpu.py tests/test_frame_gpu.py -q -x 2>&1 | tail -4
for lean in 0 1; do
echo "== bench lean=$lean"; SSLPL_SOLO_LEAN=$lean timeout 1200 python bench.py --no-cpu-baseline --steps 10 2> gpurun_out/i.err > gpurun_out/i_bench_$lean.json; python - <<PY
import json
d = json.loads(open('gpurun_out/i_bench_$lean.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['ms_per_step']); print(d['roofline']['stage_ms']['lsd_regions']); print(d['in_pipeline']['kernel_ms_per_step'])
PY
tail -12 gpurun_out/i.err | cut -c1-300
done
