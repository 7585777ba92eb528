"""Small workload for ncu captures: python tools/profile_run.py {orb|line} [frames]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import __graft_entry__ as g, synth
pkg = g.load_package()
what = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
frames = synth.batch(640, 480, B)
if what == "orb":
    ext = pkg.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=B)
    for _ in range(3): ext.extract_batch(frames)
else:
    ls = pkg.LineSegment(40, max_width=640, max_height=480, max_batch=B)
    for _ in range(3): ls.extract_batch(frames)
print("done")
