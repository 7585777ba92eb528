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


def _small_motion(rng, stereo_sign=0, k=1.0):
    ang = rng.normal(0, 0.01 * k, 3); t = rng.normal(0, 0.03 * k, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
    Ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
    Rz = np.array([[np.cos(ang[2]), -np.sin(ang[2]), 0], [np.sin(ang[2]), np.cos(ang[2]), 0], [0, 0, 1]])
    T = np.eye(4, dtype=np.float32); T[:3, :3] = (Rz @ Ry @ Rx).astype(np.float32); T[:3, 3] = t.astype(np.float32)
    T[2, 3] += np.float32(0.3 * stereo_sign)
    return T


_LINE_CACHE = {}


def _lines(oracle, synth, f, nlines=200):
    key = (f, nlines)
    if key not in _LINE_CACHE:
        kl, ld, _ = oracle.LineOracle(nlines).extract(synth.frame(640, 480, f))
        _LINE_CACHE[key] = (kl, ld)
    return _LINE_CACHE[key]


ICL = (481.2, 480.0, 319.5, 239.5)        # Examples/ICL.yaml (|fy|)


def line_scenario(oracle, synth, seed, f0=0, nlines=200, p_valid=0.9, p_obs=0.8, p_held=0.1, stereo_sign=0):
    """Lines of two consecutive synthetic frames.  The first frame's lines become MapLines (end points back-projected at random
    depths, last camera = world); the second frame's lines get random octaves 0..2 so that the level tests have something to do.
    Returns dict(last=..., cur=..., Tcw, Tlw, cam5, bounds, sf)."""
    rng = np.random.default_rng(500 + seed)
    kl1, ld1 = _lines(oracle, synth, f0, nlines); kl2, ld2 = _lines(oracle, synth, f0 + 1, nlines)
    n1, n2 = len(kl1), len(kl2)
    fx, fy, cx, cy = ICL
    zs, ze = rng.uniform(1.5, 6.0, n1), rng.uniform(1.5, 6.0, n1)
    SP = np.stack([(kl1["startPointX"] - cx) / fx * zs, (kl1["startPointY"] - cy) / fy * zs, zs], 1)
    EP = np.stack([(kl1["endPointX"] - cx) / fx * ze, (kl1["endPointY"] - cy) / fy * ze, ze], 1)
    Pw = np.concatenate([SP, EP], 1).astype(np.float64)
    Pw[rng.random(n1) < 0.03, 2] *= -1                                         # a few start points behind the camera
    dml = ld1.copy(); flip = rng.random(dml.shape) < 0.02; dml[flip] ^= (1 << rng.integers(0, 8, flip.sum())).astype(np.uint8)
    state = np.where(rng.random(n1) < p_valid, 1, rng.integers(0, 4, n1)).astype(np.uint8)     # 0 none, 1 good, 2 bad, 3 outlier
    last = dict(state=state, obs=(rng.random(n1) < p_obs).astype(np.uint8), Pw=Pw, dml=dml, oct=rng.integers(0, 3, n1).astype(np.int32), kl=kl1)
    kl2a = np.stack([kl2["pt_x"], kl2["pt_y"], kl2["angle"]], 1).astype(np.float32)
    held = np.where(rng.random(n2) < p_held, rng.integers(1, 3, n2), 0).astype(np.uint8)
    cur = dict(ld=ld2, kl=kl2a, oct=rng.integers(0, 3, n2).astype(np.int32), held=held, keylines=kl2)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    return dict(last=last, cur=cur, Tcw=_small_motion(rng, stereo_sign), Tlw=np.eye(4, dtype=np.float32),
                cam5=(fx, fy, cx, cy, 40.0 / fx), bounds=(0.0, 640.0, 0.0, 480.0), sf=sf)


def local_lines_scenario(oracle, synth, seed, f0=0, nlines=200, p_inview=0.9, p_bad=0.03, p_obs=0.8, p_held=0.1):
    """Tracking::SearchLocalLines: MapLines with the tracking fields Frame::isInFrustum leaves on them, projected near the lines of
    frame f0 + 1.  Returns (ml, cur, sf)."""
    rng = np.random.default_rng(700 + seed)
    kl1, ld1 = _lines(oracle, synth, f0, nlines); kl2, ld2 = _lines(oracle, synth, f0 + 1, nlines)
    n1, n2 = len(kl1), len(kl2)
    reps = np.concatenate([np.arange(n1), rng.integers(0, n1, n1 // 2)]); rng.shuffle(reps)
    nml = len(reps)
    dml = ld1[reps].copy(); flip = rng.random(dml.shape) < 0.015; dml[flip] ^= (1 << rng.integers(0, 8, flip.sum())).astype(np.uint8)
    proj = np.stack([kl1["startPointX"][reps], kl1["startPointY"][reps], kl1["endPointX"][reps], kl1["endPointY"][reps]], 1) + rng.normal(0, 1.5, (nml, 4))
    ml = dict(inview=(rng.random(nml) < p_inview).astype(np.uint8), bad=(rng.random(nml) < p_bad).astype(np.uint8),
              obs=(rng.random(nml) < p_obs).astype(np.uint8), proj=proj.astype(np.float32), level=rng.integers(0, 3, nml).astype(np.int32),
              viewcos=np.where(rng.random(nml) < 0.5, 0.9995, rng.uniform(0.5, 0.998, nml)).astype(np.float32), desc=dml)
    kl2a = np.stack([kl2["pt_x"], kl2["pt_y"], kl2["angle"]], 1).astype(np.float32)
    held = np.where(rng.random(n2) < p_held, rng.integers(1, 3, n2), 0).astype(np.uint8)
    cur = dict(ld=ld2, kl=kl2a, oct=rng.integers(0, 3, n2).astype(np.int32), held=held)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    return ml, cur, sf


def fuse_points_scenario(oracle, synth, seed, f0=0, stereo=False):
    """LocalMapping::SearchInNeighbors: MapPoints (frame f0's features back-projected, with normals and distance ranges that mostly
    pass the gates) fused into the KeyFrame made of frame f0 + 1.  Returns dict(mp=..., kf=..., Tcw, cam5, bounds, sf, inv_sigma2)."""
    rng = np.random.default_rng(900 + seed)
    orc = oracle.OrbOracle(1000, 1.2, 8, 20, 7)
    k1, d1 = orc.extract(synth.frame(640, 480, f0)); k2, d2 = k1.copy(), d1.copy()      # the KeyFrame sees the same features, from a pose a hair away
    n1, n2 = len(k1), len(k2)
    fx, fy, cx, cy = ICL; bf = 40.0
    z = rng.uniform(1.0, 6.0, n1)
    Xw = np.stack([(k1["x"] - cx) / fx * z, (k1["y"] - cy) / fy * z, z], 1).astype(np.float32)
    Xw[rng.random(n1) < 0.03, 2] *= -1
    Tcw = _small_motion(rng, k=0.1)
    Ow = -(Tcw[:3, :3].astype(np.float64).T @ Tcw[:3, 3].astype(np.float64))
    PO = Xw.astype(np.float64) - Ow
    dist = np.linalg.norm(PO, axis=1)
    normal = PO / dist[:, None] + rng.normal(0, 0.25, (n1, 3))
    normal /= np.linalg.norm(normal, axis=1)[:, None]
    normal[rng.random(n1) < 0.05] *= -1                                        # seen from behind: fails the viewing-angle gate
    lvl = k1["octave"].astype(np.float64) + np.where(rng.random(n1) < 0.8, rng.uniform(-0.9, -0.05, n1), rng.uniform(-3, 3, n1))
    max_raw = (dist * 1.2 ** lvl).astype(np.float32)
    min_raw = (max_raw / np.float32(1.2 ** 7)).astype(np.float32)
    far = rng.random(n1) < 0.04; max_raw[far] = (dist[far] * 0.5).astype(np.float32)   # out of the scale-invariance range
    dmp = d1.copy(); flip = rng.random(dmp.shape) < 0.02; dmp[flip] ^= (1 << rng.integers(0, 8, flip.sum())).astype(np.uint8)
    state = np.where(rng.random(n1) < 0.9, 1, rng.integers(0, 4, n1)).astype(np.uint8)          # 0 NULL, 1 good, 2 bad, 3 already in the KeyFrame
    mp = dict(state=state, nobs=rng.integers(0, 4, n1).astype(np.int32), Xw=Xw, normal=normal.astype(np.float32), min_raw=min_raw, max_raw=max_raw, desc=dmp)
    kfobs = np.where(rng.random(n2) < 0.3, rng.integers(0, 4, n2), -1).astype(np.int32)
    uright = np.where(rng.random(n2) < 0.7, k2["x"] - bf / rng.uniform(1, 6, n2), -1).astype(np.float32) if stereo else None
    kf = dict(desc=d2, kps=k2, x=k2["x"], y=k2["y"], oct=k2["octave"].astype(np.int32), uright=uright, kfobs=kfobs)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    sigma2 = (sf * sf).astype(np.float32)
    return dict(mp=mp, kf=kf, Tcw=Tcw, cam5=(fx, fy, cx, cy, bf), bounds=(0.0, 640.0, 0.0, 480.0), sf=sf,
                inv_sigma2=(np.float32(1.0) / sigma2).astype(np.float32))


def fuse_lines_scenario(oracle, synth, seed, f0=0, nlines=200):
    """As fuse_points_scenario for MapLines (LSDmatcher::Fuse)."""
    rng = np.random.default_rng(1100 + seed)
    kl1, ld1 = _lines(oracle, synth, f0, nlines); kl2, ld2 = kl1.copy(), ld1.copy()
    n1, n2 = len(kl1), len(kl2)
    fx, fy, cx, cy = ICL
    zs = rng.uniform(1.5, 6.0, n1); ze = zs + rng.normal(0, 0.2, n1)
    SP = np.stack([(kl1["startPointX"] - cx) / fx * zs, (kl1["startPointY"] - cy) / fy * zs, zs], 1)
    EP = np.stack([(kl1["endPointX"] - cx) / fx * ze, (kl1["endPointY"] - cy) / fy * ze, ze], 1)
    Pw = np.concatenate([SP, EP], 1).astype(np.float32).astype(np.float64)      # float-representable, as both sides narrow to float
    Pw[rng.random(n1) < 0.03, 2] *= -1
    Tcw = _small_motion(rng, k=0.2)
    Ow = -(Tcw[:3, :3].astype(np.float64).T @ Tcw[:3, 3].astype(np.float64))
    OM = 0.5 * (Pw[:, :3] + Pw[:, 3:]) - Ow
    dist = np.linalg.norm(OM, axis=1)
    normal = OM / dist[:, None] + rng.normal(0, 0.25, (n1, 3)); normal /= np.linalg.norm(normal, axis=1)[:, None]
    normal[rng.random(n1) < 0.05] *= -1
    normal = normal.astype(np.float32).astype(np.float64)
    lvl = np.where(rng.random(n1) < 0.8, rng.integers(0, 2, n1) + rng.uniform(-0.9, -0.05, n1), rng.uniform(-1, 4, n1))
    max_raw = (dist * 1.2 ** lvl).astype(np.float32); min_raw = (max_raw / np.float32(1.2 ** 7)).astype(np.float32)
    dml = ld1.copy(); flip = rng.random(dml.shape) < 0.02; dml[flip] ^= (1 << rng.integers(0, 8, flip.sum())).astype(np.uint8)
    state = np.where(rng.random(n1) < 0.9, 1, rng.integers(0, 3, n1)).astype(np.uint8)          # 0 NULL, 1 good, 2 bad
    ml = dict(state=state, nobs=rng.integers(0, 4, n1).astype(np.int32), Pw=Pw, normal=normal, min_raw=min_raw, max_raw=max_raw, desc=dml)
    kl2a = np.stack([kl2["pt_x"], kl2["pt_y"], kl2["angle"]], 1).astype(np.float32)
    kf = dict(ld=ld2, kl=kl2a, oct=np.where(rng.random(n2) < 0.8, 0, 1).astype(np.int32), kfobs=np.where(rng.random(n2) < 0.3, rng.integers(0, 4, n2), -1).astype(np.int32))
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    return dict(ml=ml, kf=kf, Tcw=Tcw, cam5=(fx, fy, cx, cy, 0.0), bounds=(0.0, 640.0, 0.0, 480.0), sf=sf)
