#!/bin/bash
# source-level ncu capture of the one-warp walker (148 frames, one per SM) + baseline timings
cd /root/repo; mkdir -p gpurun_out
echo "== baseline walker timings (solo forced)"; SSLPL_WALKER_WARPS=-1 timeout 600 python tools/walker_scaling.py 148 513 2>&1 | tail -3
echo "== ncu solo walker"
SSLPL_WALKER_WARPS=-1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_lsd_regions_solo -s 2 -c 1 -o gpurun_out/r02m_solo148 python tools/profile_run.py line 148 2>&1 | tail -3
ls -la gpurun_out/
