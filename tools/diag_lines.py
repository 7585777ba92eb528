"""Diagnostic: where do the GPU and oracle LSD segment lists diverge? (run on the GPU box)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, cv2
import __graft_entry__ as g, synth
pkg = g.load_package(); O = g.load_oracle()
for name, img in [(f"syn{f}", synth.frame(640, 480, f)) for f in range(0, 12)] + [("syn1280", synth.frame(1280, 960, 0))]:
    ls = pkg.LineSegment(40, max_width=img.shape[1], max_height=img.shape[0]); ls.ExtractLineSegment(img); raw = ls.raw_segments()
    lo = O.LineOracle(40); lo.extract(img); oraw = lo.raw_segments()
    res = cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV).detect(img)
    nfa = res[3].ravel() if res[3] is not None else np.zeros(0)
    width = res[1].ravel()
    if raw.shape == oraw.shape and np.array_equal(raw, oraw):
        print(name, "identical", len(raw)); continue
    # align lists
    i = 0
    n = min(len(raw), len(oraw))
    while i < n and np.array_equal(raw[i], oraw[i]): i += 1
    print(name, "GPU", len(raw), "oracle", len(oraw), "first divergence at", i)
    so = {tuple(r) for r in oraw}; sg = {tuple(r) for r in raw}
    for j, r in enumerate(oraw):
        if tuple(r) not in sg: print("   only oracle: idx", j, r, "cv2 nfa", nfa[j] if j < len(nfa) else None, "width", width[j] if j < len(width) else None)
    for j, r in enumerate(raw):
        if tuple(r) not in so: print("   only GPU   : idx", j, r)
