#!/bin/bash
cd /root/repo
for w in 6 8 12 16; do timeout -s KILL 120 python tools/v3_stats.py $w 2>&1 | head -1; done
for w in 8 16; do timeout -s KILL 200 python tools/v3_stats.py $w 1280 960 2>&1 | head -2; done
echo "== old multi-warp 1280"; SSLPL_WALKER_WARPS=16 timeout -s KILL 200 python tools/walker_timing.py 2>&1 | grep -E "^(640|1280)" | head -3
