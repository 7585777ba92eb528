#!/bin/bash
# round 2b records: tests, sanitizer on the v3 walker, ncu summaries, walker statistics, bench lines
cd /root/repo; mkdir -p gpurun_out
echo "== full gpu suite"; timeout -s KILL 1700 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== memcheck v3 / lanes / lean (walker test)"
timeout 1200 compute-sanitizer --tool memcheck --log-file gpurun_out/r02b_sanitizer_memcheck_walkers.log python -m pytest tests/test_line_gpu.py -m gpu -q -k "every_region_walker and (v3 or lane or lean)" 2>&1 | tail -2
tail -3 gpurun_out/r02b_sanitizer_memcheck_walkers.log
echo "== synccheck v3"
timeout 900 compute-sanitizer --tool synccheck --log-file gpurun_out/r02b_sanitizer_synccheck_walkers.log python -m pytest tests/test_line_gpu.py -m gpu -q -k "every_region_walker and (v3 or lane)" 2>&1 | tail -2
tail -3 gpurun_out/r02b_sanitizer_synccheck_walkers.log
echo "== ncu v3 (one 640x480 frame, 8 warps)"
timeout 600 ncu --set full --clock-control none -k regex:k_lsd_regions_v3 -s 2 -c 1 -o gpurun_out/r02b_v3_1frame python tools/profile_run.py line 1 2>&1 | tail -1
echo "== ncu lanes (one frame)"
SSLPL_WALKER_LANES=1 SSLPL_WALKER_WARPS=-1 timeout 600 ncu --set full --clock-control none -k regex:k_lsd_regions_lanes -s 2 -c 1 -o gpurun_out/r02b_lanes_1frame python tools/profile_run.py line 1 2>&1 | tail -1
echo "== stats"; (for w in 4 8 16; do python tools/v3_stats.py $w; done; python tools/v3_stats.py 16 1280 960; python tools/lanes_stats.py; ./tools/lat_probe.bin; ./tools/lat_probe2.bin) > gpurun_out/r02b_walker_stats.txt 2>&1; tail -3 gpurun_out/r02b_walker_stats.txt
echo "== walker timing (defaults)"; timeout 600 python tools/walker_timing.py > gpurun_out/r02b_walker_timing.txt 2>&1; grep -E "^(640|1280)" gpurun_out/r02b_walker_timing.txt
echo "== bench"; timeout 1500 python bench.py --steps 10 2> gpurun_out/f.err > gpurun_out/r02b_bench_n1.json; tail -c 600 gpurun_out/r02b_bench_n1.json; echo
timeout 900 python bench.py --workload single1280 --steps 10 2>> gpurun_out/f.err > gpurun_out/r02b_bench_single1280.json; tail -c 400 gpurun_out/r02b_bench_single1280.json; echo
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/f.err > gpurun_out/r02b_bench_reference_arm.json; tail -c 300 gpurun_out/r02b_bench_reference_arm.json; echo
tail -3 gpurun_out/f.err | cut -c1-300
