"""Stress of the v3 multi-warp walker against the one-warp walker on the same frames (job traces must be identical):
   python tools/v3_stress.py [warps] [reps] [frames...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import __graft_entry__ as g, synth
pkg = g.load_package()
W = sys.argv[1] if len(sys.argv) > 1 else "16"; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
fr = [int(a) for a in sys.argv[3:]] or [0, 3, 8]
os.environ["SSLPL_LINE_TRACE"] = "1"
os.environ["SSLPL_WALKER_WARPS"] = "-1"; os.environ.pop("SSLPL_WALKER_V3", None)
ref = pkg.LineSegment(40, max_width=640, max_height=480, max_batch=1)
if W == "lanes":
    os.environ["SSLPL_WALKER_WARPS"] = "-1"; os.environ["SSLPL_WALKER_LANES"] = "1"
else:
    os.environ["SSLPL_WALKER_WARPS"] = W; os.environ["SSLPL_WALKER_V3"] = "1"
v3 = pkg.LineSegment(40, max_width=640, max_height=480, max_batch=1)
bad = 0
for f in fr:
    img = synth.frame(640, 480, f)[None]
    ref.extract_batch(img); tr = ref.debug_trace(0)
    for rep in range(reps):
        v3.extract_batch(img); tv = v3.debug_trace(0)
        same = tr.shape == tv.shape and np.array_equal(tr[:, :3], tv[:, :3]) and np.array_equal(tr[:, 4:], tv[:, 4:])
        if not same:
            bad += 1
            n = min(len(tr), len(tv)); d = [i for i in range(n) if not (np.array_equal(tr[i, :3], tv[i, :3]) and np.array_equal(tr[i, 4:], tv[i, 4:]))]
            i = d[0] if d else n
            print(f"frame {f} rep {rep}: rows {len(tr)} vs {len(tv)}, first difference at row {i}")
            for j in range(max(0, i - 1), min(n, i + 3)):
                print("   ref", [int(tr[j, 0]) % 512, int(tr[j, 0]) // 512, int(tr[j, 1])], np.round(tr[j, 4:9], 2).tolist(), " v3", [int(tv[j, 0]) % 512, int(tv[j, 0]) // 512, int(tv[j, 1])], np.round(tv[j, 4:9], 2).tolist())
            if bad >= 4: break
    if bad >= 4: break
print("mismatching runs:", bad)
