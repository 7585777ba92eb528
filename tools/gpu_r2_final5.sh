#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout -s KILL 1700 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/f.err
timeout 900 python bench.py --workload single1280 > gpurun_out/r02b_bench_single1280.json 2>> gpurun_out/f.err
python - <<'PY'
import json
for f in ("r02b_bench_n1", "r02b_bench_single1280"):
    d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1]); print(f, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 2), d["clocks"]["reasons"], round(d["cpu_baseline"]["value"]), d["roofline"]["frac"])
PY
