#!/bin/bash
# after the k_fast rewrite + matcher create fix: smoke, full suite, sanitizers on the round-2 kernels, bench
cd /root/repo; mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== full gpu suite"; timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
SEL="non_tma or test_batch_equals_single or icl_frame_1000 or icl_frame_lines or test_line_edge_cases or test_frame or test_fuse or test_line_search or map_points or initialization or test_row3 or synthetic_640_lines"
echo "== memcheck (multi-warp walker for single frames, everything else as shipped)"
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02b_memcheck.log python -m pytest tests -m gpu -q -k "$SEL" 2>&1 | tail -3
tail -2 gpurun_out/r02b_memcheck.log
echo "== racecheck with the one-warp walker (the multi-warp walker synchronises warps with locks and flags, which racecheck does not model: see r02b_racecheck.log of the earlier run)"
SSLPL_WALKER_WARPS=-1 timeout 1200 compute-sanitizer --tool racecheck --log-file gpurun_out/r02b_racecheck_rest.log python -m pytest tests -m gpu -q -k "icl_frame_1000 or icl_frame_lines or test_frame_equals or test_fuse or test_line_search or map_points or initialization or non_tma or test_row3" 2>&1 | tail -3
tail -2 gpurun_out/r02b_racecheck_rest.log
echo "== synccheck"
timeout 600 compute-sanitizer --tool synccheck --log-file gpurun_out/r02b_synccheck.log python -m pytest tests -m gpu -q -k "icl_frame_1000 or icl_frame_lines or test_fuse or non_tma" 2>&1 | tail -3
tail -2 gpurun_out/r02b_synccheck.log
echo "== bench"; timeout 1200 python bench.py --steps 10 2> gpurun_out/l.err > gpurun_out/l_bench.json; echo "rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/l_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e'], d['table_check'], d['cpu_baseline'], d.get('cpu_baseline_cv2')); print(d['roofline']['stage_ms'])
PY
tail -3 gpurun_out/l.err | cut -c1-300
