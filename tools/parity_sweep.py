"""Statistical parity sweep on the GPU box: N synthetic frames through the batched CUDA path vs the CPU oracle.
Prints how many frames differ in ORB keypoints / descriptors, raw LSD segments, KeyLines / LBD descriptors.
Usage: python tools/parity_sweep.py [frames=256] [start=5000]"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
import threading
import __graft_entry__ as g, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
START = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
pkg = g.load_package(); O = g.load_oracle()
frames = synth.batch(640, 480, N, start=START)
ext = pkg.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=N)
k, d, n = ext.extract_batch(frames)
ls = pkg.LineSegment(40, max_width=640, max_height=480, max_batch=N)
kl, ld, eq, nl = ls.extract_batch(frames)
raw = [ls.raw_segments(frame=f) for f in range(N)]
tl = threading.local()

def cpu(f):
    if not hasattr(tl, "o"):
        tl.o = O.OrbOracle(1000, 1.2, 8, 20, 7); tl.l = O.LineOracle(40)
    ok, od = tl.o.extract(frames[f])
    okl, old, oeq = tl.l.extract(frames[f])
    return ok, od, okl, old, oeq, tl.l.raw_segments()

with ThreadPoolExecutor(max_workers=min(64, os.cpu_count() or 8)) as ex:
    ref = list(ex.map(cpu, range(N)))
bad = {"orb_count": 0, "orb_keypoints": 0, "orb_descriptors": 0, "lsd_segments": 0, "lsd_one_borderline": 0, "keylines": 0, "lbd": 0}
for f in range(N):
    ok, od, okl, old, oeq, oraw = ref[f]
    if n[f] != len(ok): bad["orb_count"] += 1; continue
    if k[f, :n[f]].tobytes() != ok.tobytes(): bad["orb_keypoints"] += 1
    if not np.array_equal(d[f, :n[f]], od): bad["orb_descriptors"] += 1
    if raw[f].shape != oraw.shape or not np.array_equal(raw[f], oraw):
        so = {tuple(np.round(r, 3)) for r in oraw}; sg = {tuple(np.round(r, 3)) for r in raw[f]}
        bad["lsd_one_borderline" if len(so ^ sg) <= 1 else "lsd_segments"] += 1
        continue
    if nl[f] != len(okl) or kl[f, :nl[f]].tobytes() != okl.tobytes(): bad["keylines"] += 1
    if nl[f] == len(okl) and not np.array_equal(ld[f, :nl[f]], old): bad["lbd"] += 1
print(json.dumps({"frames": N, "start": START, "keypoints_total": int(n.sum()), "segments_total": int(sum(len(r) for r in raw)),
                  "frames_differing": bad}))
