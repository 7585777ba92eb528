#!/bin/bash
# bench: line / frame handles kept in flight (6 = default) against the step time, interleaved repetitions
cd /root/repo
for rep in 1 2; do for r in 8 10 12; do timeout 600 python bench.py --steps 12 --no-cpu-baseline --line-ring $r 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('line-ring $r', 'device ms', round(d['ms_per_step'], 2), 'e2e ms', round(d['e2e']['ms_per_step'], 2))"; done; done
