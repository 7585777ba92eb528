#!/bin/bash
cd /root/repo
for t in 1 2; do for w in 8 16; do echo "THR=$t"; THR=$t timeout -s KILL 120 python tools/v3_stats.py $w 2>&1 | head -1; THR=$t timeout -s KILL 200 python tools/v3_stats.py $w 1280 960 2>&1 | head -2; done; done
timeout -s KILL 250 python tools/v3_stress.py 16 10 0 3 8 2>&1 | tail -1
