"""Diagnostic: first LSD region whose trace differs between GPU and oracle (run on the GPU box with SSLPL_LINE_TRACE=1)."""
import sys, os
os.environ["SSLPL_LINE_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import __graft_entry__ as g, synth
pkg = g.load_package(); O = g.load_oracle()
np.set_printoptions(precision=17, linewidth=250)
for f in (3, 1, 11):
    img = synth.frame(640, 480, f)
    ls = pkg.LineSegment(40, max_width=640, max_height=480); ls.ExtractLineSegment(img); tg = ls.debug_trace()
    lo = O.LineOracle(40); lo.extract(img); to = lo.trace()
    print("frame", f, "trace rows gpu", len(tg), "oracle", len(to))
    n = min(len(tg), len(to))
    for i in range(n):
        if not np.array_equal(tg[i], to[i]):
            d = np.abs(tg[i] - to[i])
            if d[:3].max() > 0 or d[3] > 1e-9 or d[4:].max() > 1e-9:
                print("  first significant diff at row", i); print("   gpu   ", tg[i]); print("   oracle", to[i])
                for j in range(i + 1, min(i + 3, n)): print("   next gpu", tg[j][:4], "oracle", to[j][:4])
                break
    nd = sum(1 for i in range(n) if not np.array_equal(tg[i], to[i]))
    md = max((np.abs(tg[i][3] - to[i][3]) for i in range(n) if abs(to[i][3]) < 1e8 and abs(tg[i][3]) < 1e8), default=0)
    print("  rows not bit-equal:", nd, "max |dlog_nfa| among them:", md)
