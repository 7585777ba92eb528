"""One v3 run of a frame with the per-attempt log of the first ranks (SSLPL_WALKER_DBG bit 16), repeated until the job trace differs
from the one-warp walker's: python tools/v3_debug.py warps frame"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import __graft_entry__ as g, synth
pkg = g.load_package()
W = sys.argv[1]; f = int(sys.argv[2])
os.environ["SSLPL_LINE_TRACE"] = "1"; os.environ["SSLPL_WALKER_SEED"] = str(int(sys.argv[4]) * 512 + int(sys.argv[3]))
os.environ["SSLPL_WALKER_WARPS"] = "-1"
ref = pkg.LineSegment(40, max_width=640, max_height=480, max_batch=1)
os.environ["SSLPL_WALKER_WARPS"] = W; os.environ["SSLPL_WALKER_V3"] = "1"; os.environ["SSLPL_WALKER_DBG"] = str(16 + 64 + 128)
v3 = pkg.LineSegment(40, max_width=640, max_height=480, max_batch=1)
img = synth.frame(640, 480, f)[None]
ref.extract_batch(img); tr = ref.debug_trace(0)
for rep in range(30):
    print(f"##### run {rep}", flush=True)
    v3.extract_batch(img); tv = v3.debug_trace(0)
    same = tr.shape == tv.shape and np.array_equal(tr[:, :3], tv[:, :3])
    print(f"##### run {rep} same={same}", flush=True)
    if not same:
        n = min(len(tr), len(tv)); d = [i for i in range(n) if not np.array_equal(tr[i, :3], tv[i, :3])]
        for j in d[:3]: print("   row", j, "ref", [int(tr[j, 0]) % 512, int(tr[j, 0]) // 512, int(tr[j, 1]), int(tr[j, 2])], "v3", [int(tv[j, 0]) % 512, int(tv[j, 0]) // 512, int(tv[j, 1]), int(tv[j, 2])])
        break
