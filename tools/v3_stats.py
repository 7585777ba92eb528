"""Statistics of the v3 multi-warp walker (SSLPL_WALKER_DBG=32): python tools/v3_stats.py [warps] [w h]"""
import os, sys
W = sys.argv[1] if len(sys.argv) > 1 else "16"
os.environ["SSLPL_WALKER_V3"] = "1"; os.environ["SSLPL_WALKER_WARPS"] = W; os.environ["SSLPL_WALKER_DBG"] = str(32 + 256 * int(os.environ.get("THR", "0")))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, ctypes
import __graft_entry__ as g, synth
pkg = g.load_package()
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (640, 480)
frames = synth.batch(w, h, 1)
ls = pkg.LineSegment(40, max_width=w, max_height=h, max_batch=1)
ls.set_profiling(True)
for _ in range(3): ls.extract_batch(frames)
print("warps", W, {k: round(v, 2) for k, v in ls.stage_ms().items()})
st = list(ls.walker_stats().values())
names = ["commits", "redo_poison", "redo_dep", "void", "presumed_swallowed", "presumed_not", "claimer_ringfull_passes(+aborts)", "worker_idle_with_waiting_ready_Mcyc", "ctl_retire_Mcyc", "ctl_claim_Mcyc",
         "worker_busy_Mcyc", "worker_idle_Mcyc", "worker_abort_Mcyc", "claims", "ctl_idle_Mcyc", "head_attempts"]
print({n: (round(v / 1e6, 2) if n.endswith("Mcyc") else v) for n, v in zip(names, st)})
