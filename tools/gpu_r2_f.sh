#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
echo "== bench n1"; timeout 1200 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 300 gpurun_out/r02_bench_n1.json; tail -2 gpurun_out/r02_bench_n1.err | cut -c1-300
echo "== bench single1280"; timeout 1200 python bench.py --workload single1280 > gpurun_out/r02_bench_single1280.json 2> gpurun_out/r02_bench_single1280.err; tail -c 300 gpurun_out/r02_bench_single1280.json; tail -2 gpurun_out/r02_bench_single1280.err | cut -c1-300
