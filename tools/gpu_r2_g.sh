#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for r in 3 9 12; do echo "== ring $r"; timeout 900 python bench.py --no-cpu-baseline --line-ring $r --steps 12 2> gpurun_out/ring.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['ms_per_step'])"; tail -1 gpurun_out/ring.err | cut -c1-200; done
