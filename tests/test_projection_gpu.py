"""GPU parity: ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (ORBmatcher.cc:1331-1473) with the 64 x 48
feature grid (Frame.cc:133-148, 368-421) — SURVEY.md 8(f) row 2 — against the oracle restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def scenario(oracle, synth, seed, n_claimed=0.05, p_valid=0.85, p_obs=0.9, stereo=False, f0=0):
    """Two consecutive synthetic frames; the last frame's features become MapPoints at random depths, the current camera is
    the last one moved by a small rigid motion (so that projections land near, but not on, the current features)."""
    rng = np.random.default_rng(seed)
    orc = oracle.OrbOracle(1000, 1.2, 8, 20, 7)
    k1, d1 = orc.extract(synth.frame(640, 480, f0))
    k2, d2 = orc.extract(synth.frame(640, 480, f0 + 1))
    fx, fy, cx, cy, bf = 481.2, -480.0, 319.5, 239.5, 40.0                      # Examples/ICL.yaml (fy < 0 in ICL-NUIM)
    fy = abs(fy)
    n1, n2 = len(k1), len(k2)
    z = rng.uniform(1.0, 6.0, n1).astype(np.float32)
    Xc = np.stack([(k1["x"] - cx) / fx * z, (k1["y"] - cy) / fy * z, z], 1).astype(np.float32)   # last camera = world
    ang = rng.normal(0, 0.01, 3); t = rng.normal(0, 0.03, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
    Ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
    Rz = np.array([[np.cos(ang[2]), -np.sin(ang[2]), 0], [np.sin(ang[2]), np.cos(ang[2]), 0], [0, 0, 1]])
    Tcw = np.eye(4, dtype=np.float32); Tcw[:3, :3] = (Rz @ Ry @ Rx).astype(np.float32); Tcw[:3, 3] = t.astype(np.float32)
    if stereo:
        Tcw[2, 3] += np.float32(0.3 * (1 if seed % 2 else -1))                  # forward / backward motion beyond the baseline
    Tlw = np.eye(4, dtype=np.float32)
    Xc[rng.random(n1) < 0.03, 2] *= -1                                          # a few points behind the camera
    dmp = d1.copy(); flip = rng.random(dmp.shape) < 0.02; dmp[flip] ^= (1 << rng.integers(0, 8, flip.sum())).astype(np.uint8)
    last = dict(valid=(rng.random(n1) < p_valid).astype(np.uint8), obs=(rng.random(n1) < p_obs).astype(np.uint8), Xw=Xc, dmp=dmp,
                oct=k1["octave"].astype(np.int32), angle=k1["angle"])
    cur = dict(desc=d2, x=k2["x"], y=k2["y"], oct=k2["octave"].astype(np.int32), angle=k2["angle"],
               uright=(np.where(rng.random(n2) < 0.7, k2["x"] - bf / rng.uniform(1, 6, n2), -1).astype(np.float32) if stereo else None),
               claimed=(rng.random(n2) < n_claimed).astype(np.uint8) if n_claimed else None)
    cam = (fx, fy, cx, cy, bf, bf / fx)
    bounds = (0.0, 640.0, 0.0, 480.0)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    return last, cur, Tcw, Tlw, cam, bounds, sf


@pytest.mark.parametrize("seed,th,mono,ori,claimed", [(1, 15.0, True, True, 0.05), (2, 7.0, True, True, 0.0), (3, 30.0, True, False, 0.2),
                                                       (4, 15.0, False, True, 0.05), (5, 15.0, False, True, 0.0), (6, 100.0, True, True, 0.0)])
def test_search_by_projection_frame(pkg, oracle, synth, seed, th, mono, ori, claimed):
    last, cur, Tcw, Tlw, cam, bounds, sf = scenario(oracle, synth, seed, n_claimed=claimed, stereo=not mono, f0=seed)
    n_o, a_o = oracle.search_by_projection_frame(last, cur, Tcw, Tlw, cam, bounds, sf, th, mono, ori)
    mt = pkg.Matcher(max_features=2048, max_lines=64, max_nodes=3072)
    n_g, a_g = mt.search_by_projection_frame(last, cur, Tcw, Tlw, cam, bounds, sf, th, mono, ori)
    assert n_g == n_o and np.array_equal(a_g, a_o), (seed, n_g, n_o, int((a_g != a_o).sum()))
    assert n_o > 50 or th < 10


def test_projection_edge_cases(pkg, oracle, synth):
    last, cur, Tcw, Tlw, cam, bounds, sf = scenario(oracle, synth, 9)
    mt = pkg.Matcher(max_features=2048, max_lines=64, max_nodes=3072)
    # nothing valid / everything claimed / observations all zero (later points may overwrite earlier ones: counted twice, as the reference does)
    for mod in ("novalid", "allclaimed", "noobs"):
        l2, c2 = dict(last), dict(cur)
        if mod == "novalid": l2["valid"] = np.zeros_like(last["valid"])
        if mod == "allclaimed": c2["claimed"] = np.ones(len(cur["x"]), np.uint8)
        if mod == "noobs": l2["obs"] = np.zeros_like(last["obs"]); l2["valid"] = np.ones_like(last["valid"])
        n_o, a_o = oracle.search_by_projection_frame(l2, c2, Tcw, Tlw, cam, bounds, sf, 20.0, True, True)
        n_g, a_g = mt.search_by_projection_frame(l2, c2, Tcw, Tlw, cam, bounds, sf, 20.0, True, True)
        assert n_g == n_o and np.array_equal(a_g, a_o), mod
    with pytest.raises(pkg.SslplError):                                         # needs the 3072-cell grid
        pkg.Matcher(max_features=2048, max_nodes=100).search_by_projection_frame(last, cur, Tcw, Tlw, cam, bounds, sf, 15.0)
