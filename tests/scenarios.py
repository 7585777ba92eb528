"""Shared synthetic matching scenarios (used by the GPU parity tests and by the reference-parity CPU tests)."""
import numpy as np


def projection_scenario(oracle, synth, seed, n_claimed=0.05, p_valid=0.85, p_obs=0.9, stereo=False, f0=0):
    """Two consecutive synthetic frames; the last frame's features become MapPoints at random depths, the current camera is
    the last one moved by a small rigid motion (so that projections land near, but not on, the current features).
    Returns (last, cur, Tcw, Tlw, cam, bounds, scale_factors); last/cur also carry the full keypoint arrays (`kps`)."""
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
                oct=k1["octave"].astype(np.int32), angle=k1["angle"], kps=k1)
    cur = dict(desc=d2, x=k2["x"], y=k2["y"], oct=k2["octave"].astype(np.int32), angle=k2["angle"], kps=k2,
               uright=(np.where(rng.random(n2) < 0.7, k2["x"] - bf / rng.uniform(1, 6, n2), -1).astype(np.float32) if stereo else None),
               claimed=(rng.random(n2) < n_claimed).astype(np.uint8) if n_claimed else None)
    cam = (fx, fy, cx, cy, bf, bf / fx)
    bounds = (0.0, 640.0, 0.0, 480.0)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    return last, cur, Tcw, Tlw, cam, bounds, sf


def local_map_scenario(oracle, synth, seed, f0=0, p_inview=0.9, p_bad=0.03, p_obs=0.8, p_held=0.1):
    """Tracking::SearchLocalPoints: MapPoints of a local map (made from frame f0's features) with the tracking fields
    Frame::isInFrustum leaves on them, projected near the features of frame f0 + 1.  Returns (mp, cur, bounds, scale_factors)."""
    rng = np.random.default_rng(100 + seed)
    orc = oracle.OrbOracle(1000, 1.2, 8, 20, 7)
    k1, d1 = orc.extract(synth.frame(640, 480, f0)); k2, d2 = orc.extract(synth.frame(640, 480, f0 + 1))
    n1, n2 = len(k1), len(k2)
    reps = np.concatenate([np.arange(n1), rng.integers(0, n1, n1 // 2)])          # some features are seen by two MapPoints
    rng.shuffle(reps)
    nmp = len(reps)
    dmp = d1[reps].copy(); flip = rng.random(dmp.shape) < 0.015; dmp[flip] ^= (1 << rng.integers(0, 8, flip.sum())).astype(np.uint8)
    lvl = np.clip(k1["octave"][reps] + rng.integers(-1, 2, nmp), 0, 7).astype(np.int32)
    mp = dict(inview=(rng.random(nmp) < p_inview).astype(np.uint8), bad=(rng.random(nmp) < p_bad).astype(np.uint8),
              obs=(rng.random(nmp) < p_obs).astype(np.uint8),
              projx=(k1["x"][reps] + rng.normal(0, 2.0, nmp)).astype(np.float32), projy=(k1["y"][reps] + rng.normal(0, 2.0, nmp)).astype(np.float32),
              level=lvl, viewcos=np.where(rng.random(nmp) < 0.5, 0.9995, rng.uniform(0.5, 0.998, nmp)).astype(np.float32), desc=dmp)
    held = np.where(rng.random(n2) < p_held, rng.integers(1, 3, n2), 0).astype(np.uint8)
    cur = dict(desc=d2, x=k2["x"], y=k2["y"], oct=k2["octave"].astype(np.int32), kps=k2, held=held, claimed=held)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    return mp, cur, (0.0, 640.0, 0.0, 480.0), sf
