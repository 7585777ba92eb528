#!/bin/bash
# robustness: parity sweeps (defaults, then v3 and lanes forced), extra stress on other sizes
cd /root/repo; mkdir -p gpurun_out
echo "== parity sweep 256 frames (defaults: one-warp walker for the batch)"; timeout 900 python tools/parity_sweep.py 256 7000 2>&1 | tail -2 | tee gpurun_out/r02b_parity_sweep.txt
echo "== parity sweep 96 frames, v3 forced (4 warps per frame)"; SSLPL_WALKER_V3=1 SSLPL_WALKER_WARPS=4 timeout 900 python tools/parity_sweep.py 96 9000 2>&1 | tail -2 | tee -a gpurun_out/r02b_parity_sweep.txt
echo "== parity sweep 96 frames, v3 forced (16 warps per frame)"; SSLPL_WALKER_V3=1 SSLPL_WALKER_WARPS=16 timeout 900 python tools/parity_sweep.py 96 9100 2>&1 | tail -2 | tee -a gpurun_out/r02b_parity_sweep.txt
echo "== parity sweep 48 frames, lanes forced"; SSLPL_WALKER_LANES=1 SSLPL_WALKER_WARPS=-1 timeout 900 python tools/parity_sweep.py 48 9200 2>&1 | tail -2 | tee -a gpurun_out/r02b_parity_sweep.txt
