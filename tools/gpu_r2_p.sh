#!/bin/bash
# v3 multi-warp walker: parity at several warp counts, then timings
cd /root/repo; mkdir -p gpurun_out
for w in 16 4 2; do
  echo "== line tests, v3, $w warps"; SSLPL_WALKER_V3=1 SSLPL_WALKER_WARPS=$w timeout -s KILL 150 python -m pytest tests/test_line_gpu.py -m gpu -q -x 2>&1 | tail -4
done
echo "== timings v3 16 warps"; SSLPL_WALKER_V3=1 SSLPL_WALKER_WARPS=16 timeout -s KILL 200 python tools/walker_scaling.py 1 64 2>&1 | tail -2
echo "== timings v3 8 warps"; SSLPL_WALKER_V3=1 SSLPL_WALKER_WARPS=8 timeout -s KILL 200 python tools/walker_scaling.py 1 148 513 2>&1 | tail -3
