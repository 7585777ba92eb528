"""Diagnostic: LSD trace GPU vs oracle on frame sizes that are not multiples of 4 (run on the GPU box)."""
import sys, os
os.environ["SSLPL_LINE_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import __graft_entry__ as g, synth
pkg = g.load_package(); O = g.load_oracle()
np.set_printoptions(precision=17, linewidth=250)
for (w, h, s) in [(333, 251, 2), (336, 252, 2), (335, 250, 2), (320, 243, 2), (322, 240, 2), (640, 480, 2)]:
    img = synth.frame(w, h, s)
    ls = pkg.LineSegment(40, max_width=w, max_height=h); ls.ExtractLineSegment(img); tg = ls.debug_trace(); rg = ls.raw_segments()
    lo = O.LineOracle(40); lo.extract(img); to = lo.trace(); ro = lo.raw_segments()
    n = min(len(tg), len(to))
    first = next((i for i in range(n) if not np.array_equal(tg[i][:3], to[i][:3]) or abs(tg[i][3] - to[i][3]) > 1e-9 or np.abs(tg[i][4:] - to[i][4:]).max() > 1e-9), None)
    print((w, h), "sw", round(w * 0.8), "sh", round(h * 0.8), "trace rows gpu", len(tg), "oracle", len(to), "raw", len(rg), len(ro), "first diff row", first)
    if first is not None:
        print("   gpu   ", tg[first]); print("   oracle", to[first])
