"""GPU parity: ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (ORBmatcher.cc:1331-1473) with the 64 x 48
feature grid (Frame.cc:133-148, 368-421) — SURVEY.md 8(f) row 2 — against the oracle restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


from scenarios import projection_scenario as scenario


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
