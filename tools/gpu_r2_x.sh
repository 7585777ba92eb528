#!/bin/bash
cd /root/repo
echo "== stress lanes"; timeout -s KILL 300 python tools/v3_stress.py lanes 8 0 3 8 2>&1 | tail -6
echo "== timings lanes"; SSLPL_WALKER_LANES=1 SSLPL_WALKER_WARPS=-1 timeout -s KILL 300 python tools/walker_scaling.py 1 148 513 2>&1 | tail -3
