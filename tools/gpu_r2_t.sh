#!/bin/bash
cd /root/repo
for w in 4 8 16; do echo "== stress W=$w"; timeout -s KILL 250 python tools/v3_stress.py $w 25 0 3 8 2>&1 | grep -E "mismatching|first difference" | tail -3; done
for w in 16 5; do echo "== line tests, v3, $w warps"; SSLPL_WALKER_V3=1 SSLPL_WALKER_WARPS=$w timeout -s KILL 150 python -m pytest tests/test_line_gpu.py -m gpu -q 2>&1 | tail -2; done
for w in 8 16; do timeout -s KILL 120 python tools/v3_stats.py $w; done
