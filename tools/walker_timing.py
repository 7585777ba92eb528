#!/usr/bin/env python3
"""Times the LSD stages of the line path (CUDA events inside the library) for one frame, 16 frames 1280x960 and a 513-frame batch.
Usage: [SSLPL_WALKER_WARPS=n] python tools/walker_timing.py"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as g
import synth

pkg = g.load_package()
for (w, h, B, nl) in [(640, 480, 1, 40), (1280, 960, 1, 500), (1280, 960, 17, 500), (640, 480, 513, 40)]:
    frames = synth.batch(w, h, min(B, 24))
    if B > len(frames):
        frames = np.concatenate([frames] * ((B + len(frames) - 1) // len(frames)))[:B]
    ls = pkg.LineSegment(nl, max_width=w, max_height=h, max_batch=B)
    ls.set_profiling(True)
    best = None
    for it in range(4):
        ls.extract_batch(frames)
        ms = ls.stage_ms()
        if best is None or ms.get("lsd_regions", 1e9) < best.get("lsd_regions", 1e9):
            best = ms
    print(f"{w}x{h} x{B}: " + "  ".join(f"{k}={v:.3f}" for k, v in best.items()), flush=True)
    st = ls.walker_stats()
    print("   ", {k: (round(v / 1e6, 2) if "cycles" in k else v) for k, v in st.items()}, flush=True)
