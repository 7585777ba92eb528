#!/bin/bash
cd /root/repo
for w in 2 4 8 16; do echo "== W=$w"; SSLPL_WALKER_WARPS=$w timeout -s KILL 400 python -m pytest tests/test_line_gpu.py tests/test_ref_golden_gpu.py -x -q 2>&1 | grep -E "Error|passed|failed" | head -3; done
for w in 4 8 16; do echo "== walker warps $w"; SSLPL_WALKER_WARPS=$w timeout -s KILL 300 python tools/walker_timing.py 2>&1 | tail -8; done
