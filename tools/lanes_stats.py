"""Statistics of the lane-parallel walker (SSLPL_WALKER_DBG=32): python tools/lanes_stats.py"""
import os, sys
os.environ["SSLPL_WALKER_LANES"] = "1"; os.environ["SSLPL_WALKER_WARPS"] = "-1"; os.environ["SSLPL_WALKER_DBG"] = "32"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import __graft_entry__ as g, synth
pkg = g.load_package()
frames = synth.batch(640, 480, 1)
ls = pkg.LineSegment(40, max_width=640, max_height=480, max_batch=1)
ls.set_profiling(True)
for _ in range(3): ls.extract_batch(frames)
print({k: round(v, 2) for k, v in ls.stage_ms().items()})
st = list(ls.walker_stats().values())
names = ["iterations", "grow_rounds", "active_lane_rounds", "post_calls", "-", "-", "-", "-", "retire_Mcyc", "claim_assign_Mcyc", "grow_Mcyc", "post_Mcyc", "-", "claims", "-", "-"]
print({n: (round(v / 1e6, 2) if n.endswith("Mcyc") else v) for n, v in zip(names, st) if n != "-"})
