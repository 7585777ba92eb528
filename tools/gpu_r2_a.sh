#!/bin/bash
# Round-2 opening GPU session: baseline tests, compute-sanitizer memcheck/racecheck, config-4 (1280x960) bench line + launch list.
mkdir -p gpurun_out
cd /root/repo
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== memcheck"
timeout 700 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck.log \
  python -m pytest tests -m gpu -q -k "icl_frame_1000 or icl_frame_lines or batched_bow or test_line_matchers or edge_cases or medoid" 2>&1 | tail -3
tail -5 gpurun_out/r02_memcheck.log
echo "== racecheck"
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_racecheck.log \
  python -m pytest tests -m gpu -q -k "icl_frame_1000 or icl_frame_lines or batched_bow or test_line_matchers" 2>&1 | tail -3
tail -5 gpurun_out/r02_racecheck.log
echo "== bench single1280"
timeout 900 python bench.py --workload single1280 > gpurun_out/r02_bench_single1280.json 2> gpurun_out/r02_bench_single1280.err; tail -c 1500 gpurun_out/r02_bench_single1280.json; tail -3 gpurun_out/r02_bench_single1280.err | cut -c1-300
echo "== ncu launch list single1280"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_single1280.csv \
    python bench.py --workload single1280 --steps 2 --warmup 3 --no-cpu-baseline --line-ring 1 > gpurun_out/r02_bench_under_ncu_1280.log 2>&1
tail -2 gpurun_out/r02_launches_single1280.csv | cut -c1-200
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; python -c "import os; print(len(os.sched_getaffinity(0)))"
