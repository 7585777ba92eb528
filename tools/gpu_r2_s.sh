#!/bin/bash
cd /root/repo
run() { echo "== DBG=$1 x3"; for i in 1 2 3; do SSLPL_WALKER_DBG=$1 SSLPL_WALKER_V3=1 SSLPL_WALKER_WARPS=16 timeout -s KILL 150 python -m pytest tests/test_line_gpu.py -m gpu -q 2>&1 | tail -1; done; }
run 0; run 64; run 128; run 192
