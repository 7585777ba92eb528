"""CPU tests (no GPU): pin the oracle's LSD against cv2 4.13's LineSegmentDetector (LSD_REFINE_ADV) — segment
count, order and endpoints must be identical — and check the pre-processing primitives and the KeyLine / LBD
packaging for internal consistency (opencv_contrib line_descriptor is not installable here: parity unpinned)."""
import numpy as np
import cv2
import pytest


def _cv_lsd(img, scale):
    det = cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV, scale) if scale else cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV)
    r = det.detect(img)[0]
    return np.zeros((0, 4), np.float32) if r is None else r.reshape(-1, 4)


def test_lsd_synthetic_shapes_match_cv2(oracle):
    """SURVEY.md A.6 black-box facts (iv)-(vi): step edges of both polarities, diagonals, a bright bar."""
    yy, xx = np.mgrid[0:120, 0:160]
    ims = []
    a = np.full((120, 160), 50, np.uint8); a[:, 80:] = 200; ims.append(a)
    a = np.full((120, 160), 200, np.uint8); a[:, 80:] = 50; ims.append(a)
    a = np.full((120, 160), 50, np.uint8); a[60:, :] = 200; ims.append(a)
    a = np.full((120, 160), 200, np.uint8); a[60:, :] = 50; ims.append(a)
    ims.append(np.where(yy > 0.5 * xx + 10, 200, 50).astype(np.uint8))
    ims.append(np.where(yy > -0.7 * xx + 100, 200, 50).astype(np.uint8))
    a = np.zeros((120, 160), np.uint8); a[40:60, 30:110] = 200; ims.append(a)
    a = np.full((120, 160), 50, np.uint8); a[:, 80:] = 55; ims.append(a)      # step of 5: below rho = 2/sin(22.5 deg)
    a = np.full((120, 160), 50, np.uint8); a[:, 80:] = 56; ims.append(a)      # step of 6: detected
    nseg = []
    for im in ims:
        ref = _cv_lsd(im, 1.0); got = oracle.lsd_detect_scaled(im)
        assert got.shape == ref.shape and np.array_equal(got, ref)
        nseg.append(len(ref))
    assert nseg[0] == 1 and nseg[6] == 4 and nseg[7] == 0 and nseg[8] == 1


def test_lsd_matches_cv2_on_frames(oracle, icl_gray, synth):
    """Whole detector incl. the 0.8x pre-processing: identical segments in identical order (225 on the ICL frame)."""
    for name, im in [("icl", icl_gray), ("syn0", synth.frame(640, 480, 0)), ("syn5", synth.frame(640, 480, 5)),
                     ("syn320", synth.frame(320, 240, 2)),
                     # sizes where 0.8 * size is not an integer: cv::resize maps with 1 / fx, not src / dst
                     ("syn333x251", synth.frame(333, 251, 2)), ("syn322x243", synth.frame(322, 243, 3)), ("syn336x252", synth.frame(336, 252, 2))]:
        lo = oracle.LineOracle(1 << 20)
        lo.extract(im)
        got = lo.raw_segments(); ref = _cv_lsd(im, None)
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        assert np.array_equal(got, ref), name
        S = cv2.resize(cv2.GaussianBlur(im, (7, 7), 0.75), None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR_EXACT)
        assert np.array_equal(lo.scaled(), S), name
        if name == "icl":
            assert len(ref) == 225


def test_lbd_prep_matches_cv2(oracle, icl_gray):
    dx, dy = oracle.lbd_prep(icl_gray)
    b = cv2.GaussianBlur(icl_gray, (5, 5), 1)
    assert np.array_equal(dx, cv2.Sobel(b, cv2.CV_16S, 1, 0, ksize=3)) and np.array_equal(dy, cv2.Sobel(b, cv2.CV_16S, 0, 1, ksize=3))


def test_extract_line_segment_packaging(oracle, icl_gray):
    """ExtractLineSegment.cpp:18-69: top-40 by response, class ids renumbered, line equations normalised."""
    lo = oracle.LineOracle(40)
    kl, ld, eq = lo.extract(icl_gray)
    raw = lo.raw_segments()
    assert len(kl) == 40 and ld.shape == (40, 32) and eq.shape == (40, 3)
    assert np.array_equal(kl["class_id"], np.arange(40)) and np.all(kl["octave"] == 0)
    assert np.all(np.diff(kl["response"]) <= 0)
    L = np.hypot(raw[:, 0] - raw[:, 2], raw[:, 1] - raw[:, 3])
    assert abs(np.sort(L)[::-1][39] - 63.75) < 0.01 and abs(np.sort(L)[::-1][40] - 62.51) < 0.01
    assert np.allclose(kl["lineLength"], np.hypot(kl["startPointX"] - kl["endPointX"], kl["startPointY"] - kl["endPointY"]), rtol=1e-6)
    assert np.allclose(kl["response"], kl["lineLength"] / 640.0, rtol=1e-6)
    assert np.allclose(kl["angle"], np.arctan2(kl["endPointY"] - kl["startPointY"], kl["endPointX"] - kl["startPointX"]), atol=1e-6)
    # endpoints lie on their line: l . (x, y, 1) = 0 and |(l0, l1)| = 1
    sp = np.stack([kl["startPointX"], kl["startPointY"], np.ones(40)], 1).astype(np.float64)
    ep = np.stack([kl["endPointX"], kl["endPointY"], np.ones(40)], 1).astype(np.float64)
    assert np.max(np.abs((eq * sp).sum(1))) < 1e-9 and np.max(np.abs((eq * ep).sum(1))) < 1e-9
    assert np.allclose(np.hypot(eq[:, 0], eq[:, 1]), 1.0)
    # LBD: deterministic, not degenerate, and a line matches itself in a 2-px shifted copy
    kl2, ld2, _ = oracle.LineOracle(40).extract(np.ascontiguousarray(np.roll(icl_gray, 2, axis=1)))
    knn = oracle.knn2(ld, ld2)
    assert (knn[:, 1] < 40).mean() > 0.5 and len(np.unique(ld, axis=0)) == 40
    # fewer than lsdNFeatures lines: no cut, original detection order
    kl3, _, _ = oracle.LineOracle(1000).extract(icl_gray)
    assert len(kl3) == 225 and np.array_equal(kl3["class_id"], np.arange(225))
    # a flat image has no lines
    kl4, ld4, eq4 = oracle.LineOracle(40).extract(np.full((240, 320), 90, np.uint8))
    assert len(kl4) == 0


def test_num_of_pixels_is_the_8_connected_line_iterator_count(oracle, icl_gray):
    """KeyLine.numOfPixels = cv::LineIterator(start, end).count (line_descriptor LSDDetector::detectImpl): for the default
    8-connected iterator that is max(|dx|, |dy|) + 1 of the ROUNDED end points.  cv2 has no LineIterator binding, but cv2.line
    draws exactly the iterator's pixels: check the closed form against it, then the oracle's KeyLines against the closed form."""
    rng = np.random.default_rng(0)
    for _ in range(500):
        a = rng.integers(0, 200, 2); b = rng.integers(0, 200, 2)
        img = np.zeros((200, 200), np.uint8)
        cv2.line(img, (int(a[0]), int(a[1])), (int(b[0]), int(b[1])), 255, 1, cv2.LINE_8)
        assert int((img > 0).sum()) == max(abs(int(b[0] - a[0])), abs(int(b[1] - a[1]))) + 1
    kl, _, _ = oracle.LineOracle(1 << 20).extract(icl_gray)
    r = lambda v: np.rint(v).astype(np.int64)                       # cv::Point(Point2f) rounds half to even, like np.rint
    exp = np.maximum(np.abs(r(kl["endPointX"]) - r(kl["startPointX"])), np.abs(r(kl["endPointY"]) - r(kl["startPointY"]))) + 1
    assert len(kl) > 100 and np.array_equal(kl["numOfPixels"], exp)


def test_clip_line_and_iterator_count_match_cv2(oracle):
    """KeyLine.numOfPixels = cv::LineIterator(img, p1, p2, 8).count: endpoints are cvRound()ed floats in [0, lim), so 639.6
    becomes 640 — outside — and OpenCV clips the segment to the image first (cv::clipLine).  Pinned to cv2.clipLine."""
    import ctypes as C
    import cv2
    L = oracle.lib()
    rng = np.random.default_rng(12)
    W, H = 640, 480
    for _ in range(4000):
        p = rng.integers(-3, [W + 4, H + 4], (2, 2))
        if rng.random() < 0.5:                                      # the case that occurs: one coordinate exactly at the limit
            p = rng.integers(0, [W, H], (2, 2)); p[rng.integers(0, 2), rng.integers(0, 2)] = [W, H][rng.integers(0, 2)]
        ok, q1, q2 = cv2.clipLine((0, 0, W, H), (int(p[0, 0]), int(p[0, 1])), (int(p[1, 0]), int(p[1, 1])))
        x1, y1, x2, y2 = (C.c_longlong(int(v)) for v in (p[0, 0], p[0, 1], p[1, 0], p[1, 1]))
        r = L.orc_clip_line(W, H, C.byref(x1), C.byref(y1), C.byref(x2), C.byref(y2))
        assert bool(r) == bool(ok), p
        if ok:
            assert (x1.value, y1.value, x2.value, y2.value) == (q1[0], q1[1], q2[0], q2[1]), p
            n = L.orc_line_iterator_count(W, H, int(p[0, 0]), int(p[0, 1]), int(p[1, 0]), int(p[1, 1]))
            assert n == max(abs(q2[0] - q1[0]), abs(q2[1] - q1[1])) + 1
