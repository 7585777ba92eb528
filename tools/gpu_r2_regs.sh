#!/bin/bash
# one-warp walker: registers per thread (minimum resident CTAs per SM in __launch_bounds__) against the pipelined bench step
cd /root/repo
for mb in 28 32 36 42; do
  touch structure-slam-pointline_b200/csrc/line.cu
  make -C structure-slam-pointline_b200/csrc LINE_DEFS="-DSSLPL_SOLO_MINB=$mb" 2>&1 | grep -E " error"
  grep -A2 "k_lsd_regions_soloE" structure-slam-pointline_b200/csrc/line.ptxas.log | grep -E "registers|spill" | tr '\n' ' '; echo
  timeout 600 python bench.py --steps 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('minb $mb', {k: round(d[k], 2) for k in ('value', 'ms_per_step')}, 'e2e ms', round(d['e2e']['ms_per_step'], 2), 'regions alone', round(d['roofline']['stage_ms']['lsd_regions'], 1))"
done
