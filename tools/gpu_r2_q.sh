#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for w in 16 3 2; do
  echo "== line tests, v3, $w warps"; SSLPL_WALKER_V3=1 SSLPL_WALKER_WARPS=$w timeout -s KILL 150 python -m pytest tests/test_line_gpu.py -m gpu -q -x 2>&1 | tail -3
done
timeout -s KILL 120 python tools/v3_stats.py 16; timeout -s KILL 120 python tools/v3_stats.py 8; timeout -s KILL 120 python tools/v3_stats.py 16 1280 960
