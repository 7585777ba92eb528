#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
echo "== frame tests"; timeout -s KILL 600 python -m pytest tests/test_frame_gpu.py -q 2>&1 | tail -3
for r in 6 12; do echo "== ring $r"; timeout 1200 python bench.py --no-cpu-baseline --line-ring $r --steps 12 2> gpurun_out/ring.err > gpurun_out/ring_$r.json; python -c "
import sys, json
d = json.loads(open('gpurun_out/ring_$r.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e'])"; tail -3 gpurun_out/ring.err | cut -c1-300; done
