"""GPU parity: Hamming matchers (through the C-ABI) vs the CPU oracle — match index tables must be identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _two_frames(oracle, synth, w=640, h=480, nf=1000, f0=0):
    orc = oracle.OrbOracle(nf, 1.2, 8, 20, 7)
    k1, d1 = orc.extract(synth.frame(w, h, f0))
    k2, d2 = orc.extract(synth.frame(w, h, f0 + 1))
    return orc, k1, d1, k2, d2


def _fvs(oracle, d1, d2, nwords, synth):
    voc = synth.vocabulary(nwords)
    return oracle.feature_vector_csr(oracle.bow_assign(d1, voc)), oracle.feature_vector_csr(oracle.bow_assign(d2, voc)), voc


def test_descriptor_distance_and_knn2(pkg, oracle):
    rng = np.random.default_rng(3)
    ctx = pkg.Matcher(max_features=4096, max_lines=600)
    a = rng.integers(0, 256, (500, 32), dtype=np.uint8); b = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    b[:50] = a[:50]; b[50:60, 0] ^= 1
    got = ctx.descriptor_distance(a, b)
    exp = np.array([oracle.descriptor_distance(a[i], b[i]) for i in range(500)])
    assert np.array_equal(got, exp)
    # knn2 with many exact ties (few distinct rows): ties -> lower trainIdx
    for nq, nt in [(40, 40), (500, 500), (1, 2), (7, 2), (33, 65), (1000, 1000), (64, 1)]:
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8); t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        if nt > 8:
            t[nt // 2:] = t[:nt - nt // 2]          # duplicates => ties
        assert np.array_equal(ctx.knn2(q, t), oracle.knn2(q, t)), (nq, nt)
    assert ctx.knn2(np.zeros((0, 32), np.uint8), a).shape == (0, 4)


def test_bow_assign(pkg, oracle, synth):
    _, k1, d1, k2, d2 = _two_frames(oracle, synth)
    ctx = pkg.Matcher()
    for nw in (1, 7, 100, 1000):
        voc = synth.vocabulary(nw)
        assert np.array_equal(ctx.bow_assign(d1, voc), oracle.bow_assign(d1, voc)), nw


@pytest.mark.parametrize("nwords,mask,ratio,ori", [(100, False, 0.7, True), (100, True, 0.7, True), (10, True, 0.9, True),
                                                    (100, True, 0.6, False), (1000, True, 0.75, True), (1, False, 0.7, True)])
def test_search_by_bow(pkg, oracle, synth, nwords, mask, ratio, ori):
    """BASELINE.json config 3: two-frame SearchByBoW (ORBmatcher.cc:159) on a synthetic pair + synthetic vocabulary."""
    _, k1, d1, k2, d2 = _two_frames(oracle, synth)
    fv1, fv2, _ = _fvs(oracle, d1, d2, nwords, synth)
    rng = np.random.default_rng(99)
    valid1 = (rng.random(len(d1)) < 0.7).astype(np.uint8) if mask else np.ones(len(d1), np.uint8)
    n_o, m_o = oracle.search_by_bow(d1, d2, fv1, fv2, valid1, k1["angle"], k2["angle"], ratio, ori)
    m = pkg.ORBmatcher(ratio, ori, pkg.Matcher(max_features=2048, max_nodes=1024))
    n_g, m_g = m.SearchByBoW(d1, fv1, valid1, k1["angle"], d2, fv2, k2["angle"])
    assert n_g == n_o and np.array_equal(m_g, m_o)
    assert n_o > 20 or nwords == 1000
    # KeyFrame-KeyFrame variant (ORBmatcher.cc:525)
    valid2 = (rng.random(len(d2)) < 0.8).astype(np.uint8)
    n_o, m_o = oracle.search_by_bow_kf(d1, d2, fv1, fv2, valid1, valid2, k1["angle"], k2["angle"], ratio, ori)
    n_g, m_g = m.SearchByBoW(d1, fv1, valid1, k1["angle"], d2, fv2, k2["angle"], valid2=valid2)
    assert n_g == n_o and np.array_equal(m_g, m_o)


def test_search_by_bow_icl_shifted(pkg, oracle, icl_gray, synth):
    """config 3 on the real frame: input.png vs a 2-px-shifted copy."""
    orc = oracle.OrbOracle(1000, 1.2, 8, 20, 7)
    k1, d1 = orc.extract(icl_gray)
    k2, d2 = orc.extract(np.ascontiguousarray(np.roll(icl_gray, 2, axis=1)))
    fv1, fv2, _ = _fvs(oracle, d1, d2, 100, synth)
    valid1 = np.ones(len(d1), np.uint8)
    n_o, m_o = oracle.search_by_bow(d1, d2, fv1, fv2, valid1, k1["angle"], k2["angle"], 0.7, True)
    n_g, m_g = pkg.ORBmatcher(0.7, True).SearchByBoW(d1, fv1, valid1, k1["angle"], d2, fv2, k2["angle"])
    assert n_g == n_o and np.array_equal(m_g, m_o) and n_o > 100


def test_search_by_bow_sparse_nodes(pkg, oracle):
    """Feature vectors with disjoint / partially overlapping node sets exercise the lower_bound merge walk."""
    rng = np.random.default_rng(5)
    d1 = rng.integers(0, 256, (300, 32), dtype=np.uint8); d2 = d1.copy()
    d2[rng.random((300, 32)) < 0.02] ^= 4
    n1 = rng.choice([1, 4, 9, 16, 25, 400], 300).astype(np.int32); n2 = n1.copy(); n2[::7] = 36; n2[::11] = 2
    fv1, fv2 = oracle.feature_vector_csr(n1), oracle.feature_vector_csr(n2)
    a1 = rng.uniform(0, 360, 300).astype(np.float32); a2 = (a1 + rng.normal(0, 5, 300)).astype(np.float32) % 360
    v = np.ones(300, np.uint8)
    n_o, m_o = oracle.search_by_bow(d1, d2, fv1, fv2, v, a1, a2, 0.8, True)
    n_g, m_g = pkg.ORBmatcher(0.8, True).SearchByBoW(d1, fv1, v, a1, d2, fv2, a2)
    assert n_g == n_o and np.array_equal(m_g, m_o) and n_o > 50
    # empty inputs
    e = np.zeros((0, 32), np.uint8); ef = oracle.feature_vector_csr(np.zeros(0, np.int32))
    n_g, m_g = pkg.ORBmatcher(0.8, True).SearchByBoW(e, ef, np.zeros(0, np.uint8), np.zeros(0, np.float32), d2, fv2, a2)
    assert n_g == 0 and np.all(m_g == -1)


@pytest.mark.parametrize("nwords,ori", [(100, True), (10, True), (100, False)])
def test_search_for_triangulation(pkg, oracle, synth, nwords, ori):
    """ORBmatcher.cc:660-826 with a synthetic fundamental matrix of a small sideways motion."""
    orc, k1, d1, k2, d2 = _two_frames(oracle, synth)
    fv1, fv2, _ = _fvs(oracle, d1, d2, nwords, synth)
    rng = np.random.default_rng(11)
    has1 = (rng.random(len(d1)) < 0.3).astype(np.uint8); has2 = (rng.random(len(d2)) < 0.3).astype(np.uint8)
    t = orc.tables()
    # F12 for a pure x-translation: epipolar lines are the image rows (l = [0, -1, y1] up to scale), epipole far away
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32)
    for ex, ey in [(1e6, 240.0), (320.0, 240.0)]:
        n_o, p_o = oracle.search_for_triangulation(d1, d2, fv1, fv2, has1, has2, k1, k2, F12, ex, ey, t["scale"], t["sigma2"], ori)
        n_g, p_g = pkg.ORBmatcher(0.6, ori).SearchForTriangulation(d1, fv1, has1, k1, d2, fv2, has2, k2, F12, ex, ey, t["scale"], t["sigma2"])
        assert n_g == n_o and np.array_equal(p_g, p_o)
    assert n_o > 10


def test_line_matchers(pkg, oracle):
    """LSDmatcher knn-based entry points (LSDmatcher.cpp:143,257,286,329,382) + lineDescriptorMAD (Frame.cc:190)."""
    rng = np.random.default_rng(21)
    lm = pkg.LSDmatcher(pkg.Matcher(max_lines=600))
    for n1, n2 in [(40, 40), (40, 37), (500, 500), (3, 2), (100, 220)]:
        d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
        d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
        k = min(n1, n2)
        d2[:k] = d1[:k]; flip = rng.random((k, 32)) < 0.08; d2[:k][flip] ^= 16
        h1 = (rng.random(n1) < 0.6).astype(np.uint8); h2 = (rng.random(n2) < 0.4).astype(np.uint8)
        for mode, args in [(0, (h1, None)), (1, (None, None)), (2, (None, h2)), (3, (h1, h2))]:
            n_o, o_o = oracle.line_match(mode, d1, d2, *args)
            n_g, o_g = lm._run(mode, d1, d2, *args)
            assert n_g == n_o and np.array_equal(o_g, o_o), (n1, n2, mode)
        assert lm.last_mad == oracle.line_mad(oracle.knn2(d1, d2))
    with pytest.raises(pkg.SslplError):          # < 2 train rows: reference reads out of bounds -> explicit error
        lm.SerachForInitialize(d1, d2[:1])


def test_descriptor_medoid_batch(pkg, oracle):
    """MapPoint / MapLine ::ComputeDistinctiveDescriptors (MapPoint.cc:247-312): least-median descriptor per group, with
    duplicate rows (ties -> first), singleton and empty groups."""
    rng = np.random.default_rng(17)
    sizes = [1, 2, 3, 0, 7, 20, 64, 150, 5, 0, 33]
    groups = []
    for nrows in sizes:
        base = rng.integers(0, 256, (1, 32), dtype=np.uint8)
        g = np.repeat(base, nrows, 0)
        flip = rng.random(g.shape) < 0.08
        g[flip] ^= (1 << rng.integers(0, 8, int(flip.sum()))).astype(np.uint8)
        if nrows > 4:
            g[3] = g[1]                                  # exact duplicates => equal medians => the first must win
        groups.append(g)
    desc = np.concatenate(groups) if sum(sizes) else np.zeros((0, 32), np.uint8)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    bi_o, bm_o = oracle.descriptor_medoid(desc, off)
    bi_g, bm_g = pkg.Matcher().descriptor_medoid(desc, off)
    assert np.array_equal(bi_g, bi_o) and np.array_equal(bm_g, bm_o)
    assert bi_o[3] == -1 and bi_o[0] == 0
