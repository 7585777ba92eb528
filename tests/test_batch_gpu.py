"""GPU parity of the batched, device-resident pipeline that bench.py measures (BASELINE.json config 5): ORB batch ->
BoW node assignment -> FeatureVector -> SearchByBoW -> rotation filter, and LSD/LBD batch -> knn2 ratio matching, all
through the C-ABI device entry points; every consecutive pair is compared with the oracle's per-pair functions.
Also the DBoW2 vocabulary-tree transform on its own and inside the batched matcher."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _extract_host(pkg, frames, nf=1000):
    ext = pkg.ORBextractor(nf, 1.2, 8, 20, 7, max_width=frames.shape[2], max_height=frames.shape[1], max_batch=len(frames))
    k, d, n = ext.extract_batch(frames)
    return ext, k, d, n


def _device_orb(pkg, ext, frames):
    B, H, W = frames.shape
    dfr = torch.from_numpy(frames).cuda()
    ext.extract_batch_device(dfr.data_ptr(), B, W, H, W, W * H)
    ext.sync()
    return dfr, ext.device_results()


def test_batched_bow_matching_equals_per_pair_oracle(pkg, oracle, synth):
    B, NW = 5, 100
    frames = synth.batch(640, 480, B)
    ext, k, d, n = _extract_host(pkg, frames)
    dfr, (kps, desc, dn, cap) = _device_orb(pkg, ext, frames)
    voc = synth.vocabulary(NW)
    d_voc = torch.from_numpy(voc).cuda()
    mt = pkg.Matcher(max_features=ext.cap, max_lines=64, max_nodes=NW, max_batch=B)
    d_match = torch.empty((B - 1, cap), dtype=torch.int32, device="cuda"); d_nmatch = torch.empty((B - 1,), dtype=torch.int32, device="cuda")
    for ratio, ori in [(0.7, True), (0.9, False)]:
        mt.match_bow_batch_device(desc, kps, dn, B, cap, d_voc.data_ptr(), NW, ratio, ori, d_match.data_ptr(), d_nmatch.data_ptr())
        mt.sync()
        got, gn = d_match.cpu().numpy(), d_nmatch.cpu().numpy()
        for p in range(B - 1):
            d1, d2 = d[p, :n[p]], d[p + 1, :n[p + 1]]
            fv1 = oracle.feature_vector_csr(oracle.bow_assign(d1, voc)); fv2 = oracle.feature_vector_csr(oracle.bow_assign(d2, voc))
            n_o, m_o = oracle.search_by_bow(d1, d2, fv1, fv2, np.ones(len(d1), np.uint8), k[p, :n[p]]["angle"], k[p + 1, :n[p + 1]]["angle"], ratio, ori)
            assert gn[p] == n_o and np.array_equal(got[p, :n[p + 1]], m_o), (p, ratio, ori)
            assert (got[p, n[p + 1]:] == -1).all()


def test_batched_line_matching_equals_per_pair_oracle(pkg, oracle, synth):
    B, NL = 5, 40
    frames = synth.batch(640, 480, B)
    ls = pkg.LineSegment(NL, max_width=640, max_height=480, max_batch=B)
    kl, ld, eq, nl = ls.extract_batch(frames)
    dfr = torch.from_numpy(frames).cuda()
    ls.extract_batch_device(dfr.data_ptr(), B, 640, 480, 640, 640 * 480); ls.sync()
    _, ldesc, _, dnl, capl = ls.device_results()
    lm = pkg.Matcher(max_features=64, max_lines=NL, max_nodes=2, max_batch=B)
    d_lmatch = torch.empty((B - 1, capl), dtype=torch.int32, device="cuda"); d_nl = torch.zeros((B - 1,), dtype=torch.int32, device="cuda")
    lm.match_lines_batch_device(ldesc, dnl, B, capl, d_lmatch.data_ptr(), d_nl.data_ptr()); lm.sync()
    got, gn = d_lmatch.cpu().numpy(), d_nl.cpu().numpy()
    for p in range(B - 1):
        d1, d2 = ld[p, :nl[p]], ld[p + 1, :nl[p + 1]]
        n_o, o_o = oracle.line_match(0, d1, d2, np.ones(len(d1), np.uint8), None)
        assert gn[p] == n_o and np.array_equal(got[p, :nl[p + 1]], o_o), p


@pytest.mark.parametrize("k,L,stop,early", [(10, 3, 0.0, 0.0), (10, 3, 0.1, 0.2), (4, 5, 0.05, 0.3), (2, 1, 0.0, 0.0), (10, 2, 0.0, 0.0)])
def test_vocabulary_transform_matches_dbow2_restatement(pkg, oracle, synth, k, L, stop, early):
    """TemplatedVocabulary::transform (TemplatedVocabulary.h:1218-1259): word, node at every levelsup, weight; with stopped
    words, leaves above the last level and exact distance ties (duplicated node descriptors)."""
    parent, ndesc, weight, is_leaf = pkg.Vocabulary.random_arrays(k, L, seed=k * 10 + L, stop_fraction=stop, early_leaf_fraction=early)
    ndesc[2::7] = ndesc[1::7][:len(ndesc[2::7])]                     # duplicate descriptors => ties: the first child must win
    voc = pkg.Vocabulary(k, L, parent, ndesc, weight, is_leaf)
    assert voc.info()["words"] == int(is_leaf.sum())
    rng = np.random.default_rng(5)
    orc = oracle.OrbOracle(1000, 1.2, 8, 20, 7)
    _, feats = orc.extract(synth.frame(640, 480, 3))
    feats = np.concatenate([feats, ndesc[rng.integers(1, len(ndesc), 64)], rng.integers(0, 256, (64, 32), dtype=np.uint8)])
    mt = pkg.Matcher(max_features=len(feats) + 8)
    for levelsup in range(0, L + 2):
        w_o, n_o, wt_o = oracle.vocab_transform(L, parent, ndesc, weight, is_leaf, feats, levelsup)
        w_g, n_g, wt_g = mt.bow_transform(voc, feats, levelsup)
        assert np.array_equal(w_g, w_o) and np.array_equal(n_g, n_o) and np.array_equal(wt_g, wt_o), levelsup
    assert len(mt.bow_transform(voc, np.zeros((0, 32), np.uint8))[0]) == 0


def test_vocabulary_text_file_round_trip(pkg, oracle, tmp_path):
    """ORBvoc.txt format of loadFromTextFile (TemplatedVocabulary.h:1338-1420)."""
    k, L = 3, 3
    parent, ndesc, weight, is_leaf = pkg.Vocabulary.random_arrays(k, L, seed=11, stop_fraction=0.1)
    path = tmp_path / "voc.txt"
    with open(path, "w") as f:
        f.write(f"{k} {L} 0 0\n")
        for i in range(1, len(parent)):
            f.write(f"{parent[i]} {int(is_leaf[i])} " + " ".join(str(int(b)) for b in ndesc[i]) + f" {float(weight[i])!r}\n")
    voc = pkg.Vocabulary.load_text(path)
    assert (voc.k, voc.L, voc.scoring, voc.weighting) == (k, L, 0, 0) and voc.info()["nodes"] == len(parent)
    feats = np.random.default_rng(2).integers(0, 256, (300, 32), dtype=np.uint8)
    got = pkg.Matcher(max_features=512).bow_transform(voc, feats, 1)
    exp = oracle.vocab_transform(L, parent, ndesc, weight, is_leaf, feats, 1)
    assert all(np.array_equal(a, b) for a, b in zip(got, exp))
    with pytest.raises(pkg.SslplError):
        pkg.Vocabulary.load_text(tmp_path / "missing.txt")


def test_batched_matching_with_vocabulary_tree(pkg, oracle, synth):
    """sslpl_match_bow_batch_device_vocab == Frame::ComputeBoW (levelsup 4 of a k=10, L=6-like tree: here L - levelsup = 2)
    followed by SearchByBoW per pair; stopped words stay out of the FeatureVector."""
    B, k, L, levelsup = 4, 10, 3, 1
    frames = synth.batch(640, 480, B)
    ext, kk, d, n = _extract_host(pkg, frames)
    dfr, (kps, desc, dn, cap) = _device_orb(pkg, ext, frames)
    parent, ndesc, weight, is_leaf = pkg.Vocabulary.random_arrays(k, L, seed=4, stop_fraction=0.05)
    voc = pkg.Vocabulary(k, L, parent, ndesc, weight, is_leaf)
    mt = pkg.Matcher(max_features=ext.cap, max_lines=64, max_nodes=voc.level_nodes(levelsup) + 1, max_batch=B)
    d_match = torch.empty((B - 1, cap), dtype=torch.int32, device="cuda"); d_nmatch = torch.empty((B - 1,), dtype=torch.int32, device="cuda")
    d_word = torch.empty((B, cap), dtype=torch.int32, device="cuda"); d_node = torch.empty((B, cap), dtype=torch.int32, device="cuda")
    d_w = torch.empty((B, cap), dtype=torch.float64, device="cuda")
    mt.match_bow_batch_device_vocab(desc, kps, dn, B, cap, voc, levelsup, 0.7, True, d_match.data_ptr(), d_nmatch.data_ptr(),
                                    d_word.data_ptr(), d_node.data_ptr(), d_w.data_ptr())
    mt.sync()
    got, gn = d_match.cpu().numpy(), d_nmatch.cpu().numpy()
    tr = [oracle.vocab_transform(L, parent, ndesc, weight, is_leaf, d[f, :n[f]], levelsup) for f in range(B)]
    for f in range(B):
        assert np.array_equal(d_word.cpu().numpy()[f, :n[f]], tr[f][0]) and np.array_equal(d_node.cpu().numpy()[f, :n[f]], tr[f][1])
        assert np.array_equal(d_w.cpu().numpy()[f, :n[f]], tr[f][2])
    for p in range(B - 1):
        fv1 = pkg.Vocabulary.feature_vector(tr[p][1], tr[p][2]); fv2 = pkg.Vocabulary.feature_vector(tr[p + 1][1], tr[p + 1][2])
        n_o, m_o = oracle.search_by_bow(d[p, :n[p]], d[p + 1, :n[p + 1]], fv1, fv2, np.ones(n[p], np.uint8),
                                        kk[p, :n[p]]["angle"], kk[p + 1, :n[p + 1]]["angle"], 0.7, True)
        assert gn[p] == n_o and np.array_equal(got[p, :n[p + 1]], m_o), p
    # BowVector assembly (host logic of transform :1145-1195): L1-normalised TF-IDF sums over the non-stopped words
    ids, vals = voc.bow_vector(tr[0][0], tr[0][2])
    assert np.all(np.diff(ids) > 0) and abs(vals.sum() - 1.0) < 1e-12 and len(ids) == len(set(tr[0][0][tr[0][2] > 0].tolist()))
