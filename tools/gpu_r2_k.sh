#!/bin/bash
# smoke + compute-sanitizer on the round-2 kernels (walker multi-warp and one-warp, fused FAST, frame.cu, projection / fuse matchers)
cd /root/repo; mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== line tests"; timeout 900 python -m pytest tests/test_line_gpu.py -q 2>&1 | tail -3
SEL="non_tma or test_batch_equals_single or icl_frame_1000 or icl_frame_lines or test_line_edge_cases or test_frame or test_fuse or test_line_search or map_points or initialization or test_row3 or synthetic_640_lines"
echo "== memcheck"
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/r02b_memcheck.log python -m pytest tests -m gpu -q -k "$SEL" 2>&1 | tail -3
tail -4 gpurun_out/r02b_memcheck.log
echo "== memcheck one-warp walker"
SSLPL_WALKER_WARPS=-1 timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02b_memcheck_solo.log python -m pytest tests/test_line_gpu.py -q -k "icl_frame_lines or synthetic_640" 2>&1 | tail -3
tail -4 gpurun_out/r02b_memcheck_solo.log
echo "== racecheck"
timeout 2400 compute-sanitizer --tool racecheck --log-file gpurun_out/r02b_racecheck.log python -m pytest tests -m gpu -q -k "icl_frame_1000 or icl_frame_lines or test_frame_equals or test_fuse or test_line_search or map_points or initialization or non_tma" 2>&1 | tail -3
tail -4 gpurun_out/r02b_racecheck.log
echo "== synccheck"
timeout 1200 compute-sanitizer --tool synccheck --log-file gpurun_out/r02b_synccheck.log python -m pytest tests -m gpu -q -k "icl_frame_1000 or icl_frame_lines or test_fuse" 2>&1 | tail -3
tail -4 gpurun_out/r02b_synccheck.log
