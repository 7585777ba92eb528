"""Deterministic synthetic frames: textured scenes with corners and straight edges so that
both FAST and LSD fire.  Identical bytes for the CPU oracle and the GPU path.  Used by tests/ and bench.py."""
import numpy as np
import cv2


def scene(width, height, seed):
    """Base scene: random filled convex quadrilaterals + line strokes over mid-grey, N(0,2^2) noise, 3x3 box blur."""
    rng = np.random.default_rng(1234 + seed)
    big = width * height > 640 * 480
    nquads, nlines = (240, 160) if big else (60, 40)
    img = np.full((height, width), 128, np.uint8)
    for _ in range(nquads):
        c = rng.uniform([0, 0], [width, height])
        r = rng.uniform(15, 90 if not big else 140)
        ang = np.sort(rng.uniform(0, 2 * np.pi, 4))
        pts = np.stack([c[0] + r * np.cos(ang), c[1] + r * np.sin(ang)], 1).astype(np.int32)
        cv2.fillConvexPoly(img, pts, int(rng.integers(0, 256)))
    for _ in range(nlines):
        p0 = rng.uniform([0, 0], [width, height]).astype(int)
        p1 = rng.uniform([0, 0], [width, height]).astype(int)
        cv2.line(img, tuple(int(v) for v in p0), tuple(int(v) for v in p1), int(rng.integers(0, 256)), int(rng.integers(1, 4)))
    f = img.astype(np.float32) + rng.normal(0, 2.0, img.shape).astype(np.float32)
    f = cv2.blur(f, (3, 3))
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def frame(width, height, f, group=8):
    """Frame f: scene (f // group) moved by a small rigid motion (translation <= 4 px, rotation <= 1 deg)."""
    s, k = divmod(f, group)
    base = scene(width, height, s)
    if k == 0:
        return base
    t = k / (group - 1)
    ang = 1.0 * t
    M = cv2.getRotationMatrix2D((width / 2.0, height / 2.0), ang, 1.0)
    M[0, 2] += 4.0 * t
    M[1, 2] += -3.0 * t
    out = cv2.warpAffine(base, M, (width, height), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    rng = np.random.default_rng(99991 + f)
    out = out.astype(np.int16) + rng.integers(-1, 2, out.shape, dtype=np.int16)
    return np.clip(out, 0, 255).astype(np.uint8)


def batch(width, height, nframes, start=0, group=8):
    return np.stack([frame(width, height, start + i, group) for i in range(nframes)])


def vocabulary(nwords=100, seed=4242):
    """Synthetic one-level vocabulary: random 256-bit centroids."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (nwords, 32), dtype=np.uint8)
