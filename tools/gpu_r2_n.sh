#!/bin/bash
# lean one-warp walker: parity (line tests with the one-warp walker forced) + timings lean vs round-2a
cd /root/repo; mkdir -p gpurun_out
./tools/lat_probe2.bin
echo "== line tests, one-warp walker, lean"; SSLPL_WALKER_WARPS=-1 timeout 900 python -m pytest tests/test_line_gpu.py tests/test_ref_golden_gpu.py -m gpu -q -x 2>&1 | tail -5
echo "== timings lean"; SSLPL_WALKER_WARPS=-1 timeout 600 python tools/walker_scaling.py 1 148 513 2>&1 | tail -3
echo "== timings round-2a"; SSLPL_WALKER_LEAN=0 SSLPL_WALKER_WARPS=-1 timeout 600 python tools/walker_scaling.py 1 513 2>&1 | tail -2
