#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for w in 16 3; do
  echo "== line tests, v3, $w warps"; SSLPL_WALKER_V3=1 SSLPL_WALKER_WARPS=$w timeout -s KILL 150 python -m pytest tests/test_line_gpu.py -m gpu -q -x 2>&1 | tail -3
done
for w in 8 16; do timeout -s KILL 120 python tools/v3_stats.py $w; done
THR=2 timeout -s KILL 120 python tools/v3_stats.py 16
echo "== batch"; for w in 4 8; do SSLPL_WALKER_V3=1 SSLPL_WALKER_WARPS=$w timeout -s KILL 200 python tools/walker_scaling.py 148 513 2>&1 | tail -2; done
