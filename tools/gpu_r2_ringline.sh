#!/bin/bash
cd /root/repo
for r in 6 8 10 6 8 10; do timeout 600 python bench.py --steps 12 --no-cpu-baseline --line-ring $r 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('line-ring $r', {k: round(d[k], 2) for k in ('value', 'ms_per_step')}, 'e2e ms', round(d['e2e']['ms_per_step'], 2))"; done
