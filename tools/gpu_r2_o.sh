#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
SSLPL_WALKER_WARPS=-1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_lsd_regions_lean -s 2 -c 1 -o gpurun_out/r02o_lean148 python tools/profile_run.py line 148 2>&1 | tail -3
