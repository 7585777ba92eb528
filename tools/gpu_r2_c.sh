#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
echo "== line tests"; timeout -s KILL 600 python -m pytest tests/test_line_gpu.py tests/test_ref_golden_gpu.py -x -q 2>&1 | tail -6
for w in 1 4 8 16; do echo "== walker warps $w"; SSLPL_WALKER_WARPS=$w timeout -s KILL 300 python tools/walker_timing.py 2>&1 | tail -8; done
