"""CPU checks of the DBoW2 transform restatement in the oracle (TemplatedVocabulary.h:1127-1259) and of the host-side
BowVector / FeatureVector assembly in the package (pure numpy; no GPU)."""
import numpy as np


def test_tiny_tree_by_hand(oracle):
    # root(0) -> 1, 2 ; 1 -> 3, 4 ; leaves 2, 3, 4 (words 0, 1, 2 in node order)
    parent = np.array([-1, 0, 0, 1, 1], np.int32); is_leaf = np.array([0, 0, 1, 1, 1], np.uint8)
    nd = np.zeros((5, 32), np.uint8); nd[2] = 255; nd[4, 0] = 1
    w = np.array([0, 0, 1.5, 2.5, 0.0])
    f = np.zeros((3, 32), np.uint8); f[1] = 255; f[2, 0] = 1
    word, node, wt = oracle.vocab_transform(2, parent, nd, w, is_leaf, f, 0)
    assert word.tolist() == [1, 0, 2] and node.tolist() == [3, 0, 4] and wt.tolist() == [2.5, 1.5, 0.0]
    assert oracle.vocab_transform(2, parent, nd, w, is_leaf, f, 1)[1].tolist() == [1, 2, 1]
    assert oracle.vocab_transform(2, parent, nd, w, is_leaf, f, 2)[1].tolist() == [0, 0, 0]          # level <= 0: root
    # ties: identical children -> the first (lowest id) wins (strict '<', :1241)
    nd[4] = nd[3]
    assert oracle.vocab_transform(2, parent, nd, w, is_leaf, f[:1], 0)[1].tolist() == [3]


def test_brute_force_descent_on_random_tree(pkg, oracle):
    parent, nd, w, leaf = pkg.Vocabulary.random_arrays(5, 3, seed=1, stop_fraction=0.1, early_leaf_fraction=0.2)
    rng = np.random.default_rng(0)
    feats = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    word, node, wt = oracle.vocab_transform(3, parent, nd, w, leaf, feats, 1)
    children = {}
    for i in range(1, len(parent)):
        children.setdefault(int(parent[i]), []).append(i)
    wid = {n: j for j, n in enumerate(np.nonzero(leaf)[0].tolist())}
    pop = lambda a, b: int(np.unpackbits(a ^ b).sum())
    for f in range(len(feats)):
        cur, lvl, nid = 0, 0, 0
        while cur in children:
            lvl += 1
            ds = [pop(feats[f], nd[c]) for c in children[cur]]
            cur = children[cur][int(np.argmin(ds))]                 # argmin returns the first minimum
            if lvl == 3 - 1:
                nid = cur
        assert (word[f], node[f], wt[f]) == (wid[cur], nid, w[cur])


def test_bow_and_feature_vector_assembly(pkg):
    V = pkg.Vocabulary.__new__(pkg.Vocabulary)
    V.scoring, V.weighting = 0, 0                                    # L1_NORM, TF_IDF (ORBvoc.txt header "10 6 0 0")
    word = np.array([5, 2, 5, 9, 2, 7]); weight = np.array([1.0, 2.0, 3.0, 0.0, 0.5, 4.0]); node = np.array([10, 11, 10, 12, 11, 10])
    ids, vals = V.bow_vector(word, weight)
    assert ids.tolist() == [2, 5, 7] and np.allclose(vals, np.array([2.5, 4.0, 4.0]) / 10.5)
    nodes, off, idx = pkg.Vocabulary.feature_vector(node, weight)
    assert nodes.tolist() == [10, 11] and off.tolist() == [0, 3, 5] and idx.tolist() == [0, 2, 5, 1, 4]   # stopped feature 3 left out
    V.weighting = 3                                                  # BINARY: first weight kept, still normalised
    ids, vals = V.bow_vector(word, np.where(weight > 0, 1.0, 0.0))
    assert ids.tolist() == [2, 5, 7] and np.allclose(vals, 1 / 3)
    V._h = None


def test_features_in_area_equals_brute_force(oracle):
    """Frame::GetFeaturesInArea (Frame.cc:368-421) over AssignFeaturesToGrid (:133-148): the cell walk returns exactly the
    in-grid features inside the window (and level range), ordered by (cell x, cell y, feature index)."""
    rng = np.random.default_rng(4)
    n = 1500
    kx = rng.uniform(-5, 645, n).astype(np.float32); ky = rng.uniform(-5, 485, n).astype(np.float32); oc = rng.integers(0, 8, n).astype(np.int32)
    bounds = (0.0, 640.0, 0.0, 480.0)
    invw = np.float32(64) / np.float32(640); invh = np.float32(48) / np.float32(480)
    px = np.round((kx - np.float32(0)) * invw).astype(int); py = np.round((ky - np.float32(0)) * invh).astype(int)   # PosInGrid (:462-472); ties .5 do not occur
    ingrid = (px >= 0) & (px < 64) & (py >= 0) & (py < 48)
    for (x, y, r, lo, hi) in [(320, 240, 40, -1, -1), (5, 5, 30, -1, -1), (630, 470, 25, 2, 4), (100, 400, 60, 3, -1), (700, 100, 20, -1, -1),
                              (320, 240, 15, 0, 2), (-50, -50, 10, -1, -1)]:
        got = oracle.features_in_area(kx, ky, oc, bounds, x, y, r, lo, hi)
        m = ingrid & (np.abs(kx - np.float32(x)) < r) & (np.abs(ky - np.float32(y)) < r)
        if lo > 0 or hi >= 0:
            m &= oc >= lo
            if hi >= 0:
                m &= oc <= hi
        idx = np.nonzero(m)[0]
        exp = idx[np.lexsort((idx, py[idx], px[idx]))]
        assert np.array_equal(got, exp), (x, y, r, lo, hi)
