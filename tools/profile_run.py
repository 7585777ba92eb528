"""Small workload for ncu captures: python tools/profile_run.py {orb|line} [frames] [width height lines]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import __graft_entry__ as g, synth
pkg = g.load_package()
what = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
Wd, Ht = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (640, 480)
NLN = int(sys.argv[5]) if len(sys.argv) > 5 else 40
frames = synth.batch(Wd, Ht, B)
if what == "orb":
    ext = pkg.ORBextractor(1000 if Wd <= 640 else 4000, 1.2, 8, 20, 7, max_width=Wd, max_height=Ht, max_batch=B)
    for _ in range(3): ext.extract_batch(frames)
else:
    ls = pkg.LineSegment(NLN, max_width=Wd, max_height=Ht, max_batch=B)
    for _ in range(3): ls.extract_batch(frames)
print("done")
