#!/bin/bash
# One GPU-box session: parity tests, smoke, bench (both arms), ncu launch list and --set full captures of the top kernels.
# Usage: gpurun -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r1}
mkdir -p gpurun_out
cd /root/repo
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 600 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err | cut -c1-300
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>&1; tail -c 900 gpurun_out/bench_ref_${TAG}.json
echo "== ncu launch list (serial: one line handle)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --line-ring 1 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
echo "== ncu --set full: region walker (513 frames), FAST score"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_lsd_regions -s 1 -c 1 -o gpurun_out/prof_${TAG}_lsd_regions_513 python tools/profile_run.py line 513 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_fast_score|k_fast_cells|k_blur" -s 3 -c 3 -o gpurun_out/prof_${TAG}_orb python tools/profile_run.py orb 513 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
nproc; lscpu | grep "Model name"
