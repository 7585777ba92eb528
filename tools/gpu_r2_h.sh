#!/bin/bash
# fused FAST kernel: parity, stage times, ncu of the ORB kernels
cd /root/repo; mkdir -p gpurun_out
echo "== orb/golden/frame tests"; timeout -s KILL 900 python -m pytest tests/test_orb_gpu.py tests/test_ref_golden_gpu.py tests/test_frame_gpu.py tests/test_abi_gpu.py -q -x 2>&1 | tail -4
echo "== bench"; timeout 1200 python bench.py --no-cpu-baseline --steps 10 2> gpurun_out/h.err > gpurun_out/h_bench.json; python - <<'PY'
import json
d = json.loads(open('gpurun_out/h_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']); print(d.get('stage_ms') or d.get('stages')); print(d['roofline'])
PY
tail -3 gpurun_out/h.err | cut -c1-300
echo "== ncu orb"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fast|k_blur|k_orient_desc|k_octree|k_resize" -s 12 -c 12 -o gpurun_out/prof_r02_orb python tools/profile_run.py orb 513 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
