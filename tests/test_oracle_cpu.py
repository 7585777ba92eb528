"""CPU tests (no GPU): pin the oracle's primitives against cv2 4.13, its
constants against the golden values of SURVEY.md 4, and its ORB output on the ICL frame against the
known answers of SURVEY.md 8(c)."""
import hashlib
import numpy as np
import cv2
import pytest


@pytest.fixture(scope="module")
def images(icl_gray, synth):
    rng = np.random.default_rng(0)
    return [icl_gray, synth.frame(640, 480, 1), rng.integers(0, 256, (135, 100), dtype=np.uint8),
            rng.integers(0, 256, (333, 517), dtype=np.uint8)]


def test_resize_linear_matches_cv2(oracle, images):
    for img in images:
        h, w = img.shape
        for s in (1 / 1.2, 0.5, 0.8, 0.37):
            dw, dh = max(int(round(w * s)), 1), max(int(round(h * s)), 1)
            assert np.array_equal(oracle.resize_linear(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)), (img.shape, s)


def test_border_matches_cv2(oracle, images):
    for img in images:
        assert np.array_equal(oracle.border_reflect101(img, 19), cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))


def test_gaussian_blur_matches_cv2(oracle, images):
    for img in images:
        assert np.array_equal(oracle.gauss7_sigma2(img), cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101))
        # the two other fixed-point kernels used on the line path
        assert np.array_equal(oracle.sepfilter_fixed(img, [0, 4, 56, 136, 56, 4, 0]), cv2.GaussianBlur(img, (7, 7), 0.75))
        assert np.array_equal(oracle.sepfilter_fixed(img, [14, 62, 104, 62, 14]), cv2.GaussianBlur(img, (5, 5), 1.0))


def test_fast_matches_cv2(oracle, images):
    for img in images:
        for th in (20, 7, 40):
            det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            kps = det.detect(img)
            xs, ys, ss = oracle.fast9_16(img, th)
            assert len(kps) == len(xs), (img.shape, th)
            assert np.array_equal(np.array([k.pt[0] for k in kps], np.int32), xs)
            assert np.array_equal(np.array([k.pt[1] for k in kps], np.int32), ys)
            assert np.array_equal(np.array([k.response for k in kps], np.int32), ss)
    # small sub-images like the 30-px cells of ORBextractor.cc:809
    sub = images[0][100:137, 200:236]
    det = cv2.FastFeatureDetector_create(threshold=7, nonmaxSuppression=True)
    assert len(det.detect(np.ascontiguousarray(sub))) == len(oracle.fast9_16(np.ascontiguousarray(sub), 7)[0])


def test_fast_atan2_matches_cv2(oracle):
    rng = np.random.default_rng(1)
    ys = np.concatenate([rng.integers(-60000, 60000, 5000), [0, 0, 1, -1, 5, 0]]).astype(np.float32)
    xs = np.concatenate([rng.integers(-60000, 60000, 5000), [0, 1, 0, 0, 5, -3]]).astype(np.float32)
    for y, x in zip(ys, xs):
        assert oracle.fast_atan2(y, x) == np.float32(cv2.fastAtan2(float(y), float(x))), (y, x)


def test_ic_angle_and_rbrief_match_cv2_orb(oracle, icl_gray, synth):
    """IC_Angle (ORBextractor.cc:77-104) and computeOrbDescriptor (:107-147) were copied into ORB-SLAM2 from OpenCV's ORB, and
    cv2.ORB still evaluates the same formulas on its level 0: feed cv2's own keypoints (integer positions at nlevels = 1) through
    the oracle's stage entry points and demand bit-equal angles and descriptors.  cv2.ORB blurs a SUBMATRIX of its pyramid buffer,
    which makes cv::GaussianBlur take its filter-engine path instead of the bit-exact fixed-point one; a float blur rounded to
    8 bits reproduces that path, and the descriptor stage is handed that image."""
    total = 0
    for im in (icl_gray, synth.frame(640, 480, 3), synth.frame(1280, 960, 1)):
        orb = cv2.ORB_create(nfeatures=3000, scaleFactor=1.2, nlevels=1, edgeThreshold=19, firstLevel=0, WTA_K=2,
                             scoreType=cv2.ORB_FAST_SCORE, patchSize=31, fastThreshold=20)
        kps, desc = orb.detectAndCompute(im, None)
        xs = np.array([kp.pt[0] for kp in kps]); ys = np.array([kp.pt[1] for kp in kps])
        ang = np.array([kp.angle for kp in kps], np.float32)
        assert len(kps) > 300 and np.all(xs == np.rint(xs)) and np.all(ys == np.rint(ys))
        xi, yi = xs.astype(np.int32), ys.astype(np.int32)
        o = oracle.OrbOracle(1000, 1.2, 8, 20, 7)
        assert np.array_equal(o.ic_angles(im, xi, yi), ang)
        blur = np.clip(np.rint(cv2.GaussianBlur(im.astype(np.float32), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)), 0, 255).astype(np.uint8)
        assert np.array_equal(oracle.OrbOracle.brief_descriptors(blur, xi, yi, ang), desc)
        total += len(kps)
    assert total > 3000


def test_knn2_matches_bfmatcher(oracle):
    rng = np.random.default_rng(2)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, False)
    for nq, nt in [(40, 40), (100, 7), (5, 2), (300, 300)]:
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8); t = rng.integers(0, 4, (nt, 32), dtype=np.uint8)
        if nt > 4:
            t[nt // 2:] = t[:nt - nt // 2]          # exact ties
        got = oracle.knn2(q, t)
        exp = bf.knnMatch(q, t, 2)
        for i, ms in enumerate(exp):
            assert [ms[0].trainIdx, int(ms[0].distance), ms[1].trainIdx, int(ms[1].distance)] == got[i].tolist()
            assert oracle.descriptor_distance(q[i], t[ms[0].trainIdx]) == int(ms[0].distance)
            assert int(cv2.norm(q[i], t[ms[1].trainIdx], cv2.NORM_HAMMING)) == got[i, 3]


def test_golden_constants(oracle):
    """mnFeaturesPerLevel, umax, pyramid sizes derived from ORBextractor.cc:410-470,1111-1112."""
    o = oracle.OrbOracle(1000, 1.2, 8, 20, 7); t = o.tables()
    assert t["nfeat"].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert t["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert oracle.OrbOracle(4000, 1.2, 8, 20, 7).tables()["nfeat"].tolist() == [869, 724, 603, 503, 419, 349, 291, 242]
    o.extract(np.zeros((480, 640), np.uint8))
    assert [o.level_size(l) for l in range(8)] == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]
    o.extract(np.zeros((960, 1280), np.uint8))
    assert [o.level_size(l) for l in range(1, 8)] == [(1067, 800), (889, 667), (741, 556), (617, 463), (514, 386), (429, 322), (357, 268)]


def test_icl_known_answers(oracle, icl_gray):
    """whole ORB path on images/input.png under the canonical choices."""
    assert hashlib.sha1(icl_gray.tobytes()).hexdigest() == "ab2880f3ee5d99ff25839a348f91514c7e4b8577"
    o = oracle.OrbOracle(1000, 1.2, 8, 20, 7)
    kps, desc = o.extract(icl_gray)
    assert len(kps) == 1002
    assert [len(o.candidates(l)[0]) for l in range(8)] == [784, 450, 284, 231, 158, 104, 82, 60]
    assert [len(o.level_keypoints(l)[0]) for l in range(8)] == [218, 181, 151, 126, 105, 88, 73, 60]
    first = [(616, 422, 0, 265.97216796875, 7), (595, 428, 0, 302.4071044921875, 8), (611, 408, 0, 271.4791259765625, 13)]
    for k, e in zip(kps[:3], first):
        assert (k["x"], k["y"], k["octave"], k["angle"], k["response"]) == e
    assert desc[0].tolist() == [176, 12, 22, 27, 144, 163, 2, 87, 84, 11, 99, 80, 66, 49, 32, 65, 81, 2, 2, 34, 49, 184, 81, 31, 36, 174, 48, 64, 72, 64, 224, 137]
    assert hashlib.sha1(desc.tobytes()).hexdigest() == "e8dce82582b67285476bbe582e3557644739a438"
    tab = np.stack([kps["x"], kps["y"], kps["angle"]], 1).astype(np.float32)
    assert hashlib.sha1(tab.tobytes()).hexdigest() == "1105debd69ac4a65f0375a9da3f130fe9fd64ca5"
    # pyramid levels vs a cv2 chain (resize + copyMakeBorder), ORBextractor.cc:1107-1132
    prev = icl_gray
    for l in range(1, 8):
        w, h = o.level_size(l)
        prev = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(o.level(l), prev)
        assert np.array_equal(o.level(l, bordered=True), cv2.copyMakeBorder(prev, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))


def test_cell_fast_equals_per_cell_cv2(oracle, icl_gray):
    """The per-cell FAST(20) / fallback FAST(7) logic of ORBextractor.cc:789-829 re-composed from cv2 calls."""
    o = oracle.OrbOracle(1000, 1.2, 8, 20, 7); o.extract(icl_gray)
    d20 = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True)
    d7 = cv2.FastFeatureDetector_create(threshold=7, nonmaxSuppression=True)
    for l in (0, 3, 7):
        img = o.level(l); h, w = img.shape
        minB, maxBX, maxBY = 16, w - 16, h - 16
        nC, nR = int((maxBX - minB) / 30.0), int((maxBY - minB) / 30.0)
        wC, hC = int(np.ceil((maxBX - minB) / nC)), int(np.ceil((maxBY - minB) / nR))
        exp = []
        for i in range(nR):
            iniY = minB + i * hC; maxY = min(iniY + hC + 6, maxBY)
            if iniY >= maxBY - 3:
                continue
            for j in range(nC):
                iniX = minB + j * wC; maxX = min(iniX + wC + 6, maxBX)
                if iniX >= maxBX - 6:
                    continue
                sub = np.ascontiguousarray(img[iniY:maxY, iniX:maxX])
                k = d20.detect(sub) or d7.detect(sub)
                exp += [(int(p.pt[0]) + j * wC, int(p.pt[1]) + i * hC, int(p.response)) for p in k]
        xs, ys, rs = o.candidates(l)
        assert exp == list(zip(xs.tolist(), ys.tolist(), rs.tolist())), l


def test_matcher_oracle_selfconsistency(oracle, synth):
    """Structural properties of the restated matchers (no external pin exists)."""
    orc = oracle.OrbOracle(1000, 1.2, 8, 20, 7)
    k1, d1 = orc.extract(synth.frame(640, 480, 0)); k2, d2 = orc.extract(synth.frame(640, 480, 1))
    voc = synth.vocabulary(100)
    n1, n2 = oracle.bow_assign(d1, voc), oracle.bow_assign(d2, voc)
    fv1, fv2 = oracle.feature_vector_csr(n1), oracle.feature_vector_csr(n2)
    n, m = oracle.search_by_bow(d1, d2, fv1, fv2, np.ones(len(d1), np.uint8), k1["angle"], k2["angle"], 0.7, True)
    hit = np.flatnonzero(m >= 0)
    assert n == len(hit) and n > 50
    assert len(set(m[hit].tolist())) == len(hit)                       # a KF feature is used at most once per node walk
    assert all(n1[m[j]] == n2[j] for j in hit)                          # matches stay inside a vocabulary node
    assert all(oracle.descriptor_distance(d1[m[j]], d2[j]) <= 50 for j in hit)
    # identical frames: every feature matches itself when ratio test allows
    n, m = oracle.search_by_bow(d1, d1, fv1, fv1, np.ones(len(d1), np.uint8), k1["angle"], k1["angle"], 0.9, True)
    hit = np.flatnonzero(m >= 0)
    assert np.array_equal(m[hit], hit) and len(hit) > 0.8 * len(d1)
