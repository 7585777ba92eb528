#!/bin/bash
# Round-2 walker bring-up: line parity tests (short timeouts: a protocol bug would hang), then walker timings.
mkdir -p gpurun_out
cd /root/repo
echo "== line tests (icl)"; timeout -s KILL 180 python -m pytest tests/test_line_gpu.py -x -q -k "icl_frame_lines" 2>&1 | tail -5
echo "== line tests (all)"; timeout -s KILL 600 python -m pytest tests/test_line_gpu.py tests/test_ref_golden_gpu.py -x -q 2>&1 | tail -8
for w in 1 4 8 16; do
  echo "== walker warps $w"
  SSLPL_WALKER_WARPS=$w timeout -s KILL 300 python tools/walker_timing.py 2>&1 | tail -4
done
