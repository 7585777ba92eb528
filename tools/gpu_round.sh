#!/bin/bash
# One GPU-box session: parity tests, bench, ncu launch list.  Usage: gpurun -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r1}
mkdir -p gpurun_out
cd /root/repo
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>&1; tail -c 1500 gpurun_out/bench_ref_${TAG}.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
tail -3 gpurun_out/bench_under_ncu_${TAG}.log | cut -c1-300
nproc; lscpu | grep "Model name"
