#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
echo "== new gpu tests"; timeout -s KILL 600 python -m pytest tests/test_frame_gpu.py tests/test_projection_gpu.py tests/test_ref_golden_gpu.py tests/test_integration_gpu.py -q 2>&1 | tail -6
echo "== bench n1"; timeout 1200 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 600 gpurun_out/r02_bench_n1.json; tail -2 gpurun_out/r02_bench_n1.err | cut -c1-300
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2>&1; tail -c 700 gpurun_out/r02_bench_ref.json
echo "== bench single1280"; timeout 1200 python bench.py --workload single1280 > gpurun_out/r02_bench_single1280.json 2> gpurun_out/r02_bench_single1280.err; tail -c 400 gpurun_out/r02_bench_single1280.json; tail -2 gpurun_out/r02_bench_single1280.err | cut -c1-300
echo "== ncu walker (64 frames, multi-warp)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lsd_regions -s 1 -c 1 -o gpurun_out/prof_r02_lsd_regions_64 python tools/profile_run.py line 64 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lsd_regions -s 1 -c 1 -o gpurun_out/prof_r02_lsd_regions_513 python tools/profile_run.py line 513 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
