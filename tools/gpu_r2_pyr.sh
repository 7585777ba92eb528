#!/bin/bash
cd /root/repo
echo "== orb + frame tests"; timeout -s KILL 900 python -m pytest tests/test_orb_gpu.py tests/test_frame_gpu.py tests/test_ref_golden_gpu.py -m gpu -q -x 2>&1 | tail -4
for e in 0 1; do echo "== bench NO_PYR2=$e"; SSLPL_NO_PYR2=$e timeout 600 python bench.py --steps 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k: round(d[k], 2) for k in ('value', 'ms_per_step')}, 'e2e ms', round(d['e2e']['ms_per_step'], 2), {k: round(v, 3) for k, v in d['roofline']['stage_ms'].items() if k in ('pyramid', 'fast', 'blur')}, d['gpu_launches'])"; done
