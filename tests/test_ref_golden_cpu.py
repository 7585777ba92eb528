"""Fixtures frozen from THE REFERENCE ITSELF (tools/make_ref_golden.py ran oracle/_ref/libref.so — the reference's sources
compiled unmodified — in the build container) vs the CPU oracle.  Needs no /root/reference: this is what keeps the oracle
pinned on boxes that only have the repository.  The same fixtures gate the CUDA path in tests/test_ref_golden_gpu.py."""
import os
import numpy as np

from conftest import GOLDEN


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def frames(icl_gray, synth):
    return {"icl1000": (icl_gray, 1000), "icl2000": (icl_gray, 2000), "syn0": (synth.frame(640, 480, 0), 1000), "syn13": (synth.frame(640, 480, 13), 1000)}


def test_oracle_orb_equals_reference_fixtures(oracle, icl_gray, synth):
    g = load("ref_orb.npz")
    for tag, (img, nf) in frames(icl_gray, synth).items():
        k, d = oracle.OrbOracle(nf, 1.2, 8, 20, 7).extract(img)
        assert k.tobytes() == g[f"orb_{tag}_kps"].tobytes() and np.array_equal(d, g[f"orb_{tag}_desc"]), tag
        assert np.array_equal(np.bincount(k["octave"], minlength=8), g[f"orb_{tag}_levels"])


def match_inputs(oracle, synth):
    orc = oracle.OrbOracle(1000, 1.2, 8, 20, 7)
    k1, d1 = orc.extract(synth.frame(640, 480, 0)); k2, d2 = orc.extract(synth.frame(640, 480, 1))
    voc = synth.vocabulary(100)
    fv1 = oracle.feature_vector_csr(oracle.bow_assign(d1, voc)); fv2 = oracle.feature_vector_csr(oracle.bow_assign(d2, voc))
    return orc, k1, d1, k2, d2, fv1, fv2


def test_oracle_matchers_equal_reference_fixtures(oracle, synth):
    g = load("ref_match.npz")
    orc, k1, d1, k2, d2, fv1, fv2 = match_inputs(oracle, synth)
    v1, v2 = g["valid1"], g["valid2"]
    for ratio, ori in [(0.7, True), (0.9, False)]:
        n, m = oracle.search_by_bow(d1, d2, fv1, fv2, v1, k1["angle"], k2["angle"], ratio, ori)
        assert n == int(g[f"bow_{ratio}_{int(ori)}_n"]) and np.array_equal(m, g[f"bow_{ratio}_{int(ori)}"])
        n, m = oracle.search_by_bow_kf(d1, d2, fv1, fv2, v1, v2, k1["angle"], k2["angle"], ratio, ori)
        assert n == int(g[f"bowkf_{ratio}_{int(ori)}_n"]) and np.array_equal(m, g[f"bowkf_{ratio}_{int(ori)}"])
    tb = orc.tables()
    for tag in ("in", "out"):
        ex, ey = g[f"tri_{tag}_epi"]
        n, p = oracle.search_for_triangulation(d1, d2, fv1, fv2, 1 - v1, 1 - v2, k1, k2, g[f"tri_{tag}_F12"], ex, ey, tb["scale"], tb["sigma2"], True)
        assert np.array_equal(p, g[f"tri_{tag}_pairs"]) and n == len(p)
    l1, l2, h1, h2 = g["line_d1"], g["line_d2"], g["line_h1"], g["line_h2"]
    lo = oracle.LineOracle(40)
    assert np.array_equal(lo.extract(synth.frame(640, 480, 0))[1], l1)
    for mode in range(4):
        n, m = oracle.line_match(mode, l1, l2, h1, h2)
        assert n == int(g[f"line_mode{mode}_n"]) and np.array_equal(m, g[f"line_mode{mode}"]), mode
    assert oracle.line_mad(oracle.knn2(l1, l2)) == tuple(g["line_mad"])
    from scenarios import projection_scenario
    for seed, th, mono in [(1, 15.0, True), (4, 15.0, False)]:
        last, cur, Tcw, Tlw, cam, bounds, sf = projection_scenario(oracle, synth, seed, n_claimed=0.05, stereo=not mono, f0=seed)
        n, a = oracle.search_by_projection_frame(last, cur, Tcw, Tlw, cam, bounds, sf, th, mono, True)
        assert n == int(g[f"proj_{seed}_n"]) and np.array_equal(a, g[f"proj_{seed}"]), seed


def test_oracle_frame_lines_equal_reference_fixtures(oracle, icl_gray):
    g = load("ref_frame.npz")
    kl, ld, eq = oracle.LineOracle(40).extract(icl_gray)
    assert kl.tobytes() == g["keylines"].tobytes() and np.array_equal(ld, g["ldesc"]) and np.array_equal(eq, g["lineeq"])
    k, _ = oracle.OrbOracle(1000, 1.2, 8, 20, 7).extract(icl_gray)
    rnd = lambda v: np.floor(v.astype(np.float32) + np.float32(0.5)).astype(int)      # C round() on non-negative floats (PosInGrid, Frame.cc:462-472)
    gx = rnd(k["x"] * np.float32(64 / 640)); gy = rnd(k["y"] * np.float32(48 / 480))
    order = np.lexsort((np.arange(len(k)), gy, gx))
    order = order[(gx[order] >= 0) & (gx[order] < 64) & (gy[order] >= 0) & (gy[order] < 48)]
    assert np.array_equal(order, g["grid_idx"])


def row3_cases(oracle, synth):
    """The scenarios tools/make_ref_golden.py froze the reference on, with the oracle's projection stage
    applied: name -> (search arguments for the search stage, expected result keys in ref_row3.npz)."""
    from scenarios import line_scenario, local_lines_scenario, fuse_points_scenario, fuse_lines_scenario
    g = load("ref_row3.npz")
    c = {}
    sc = line_scenario(oracle, synth, 2, f0=2, stereo_sign=1)
    last, cur = sc["last"], sc["cur"]
    q = oracle.line_project_frame(last["state"] == 1, last["Pw"], last["oct"], sc["Tcw"][:3], sc["Tlw"][:3], sc["cam5"], sc["bounds"], sc["sf"], 20.0, False)
    c["lpf"] = ((q, last["obs"], last["dml"], cur["ld"], cur["kl"], cur["oct"], cur["held"], 0.8), int(g["lpf_n"]), g["lpf_assign"])
    ml, cur, sf = local_lines_scenario(oracle, synth, 2, f0=2)
    q = oracle.line_project_mls(ml["inview"], ml["bad"], ml["level"], ml["viewcos"], sf, 3.0); q["proj"] = ml["proj"]
    c["lpm"] = ((q, ml["obs"], ml["desc"], cur["ld"], cur["kl"], cur["oct"], cur["held"], 0.8), int(g["lpm_n"]), g["lpm_assign"])
    for tag, stereo in (("mono", False), ("stereo", True)):
        sc = fuse_points_scenario(oracle, synth, 2, f0=2, stereo=stereo)
        mp, kf = sc["mp"], sc["kf"]
        q = oracle.fuse_project_points(mp["state"] != 1, mp["Xw"], mp["normal"], g[f"fp_{tag}_min_inv"], g[f"fp_{tag}_max_inv"], mp["max_raw"], sc["Tcw"][:3],
                                       g[f"fp_{tag}_Ow"], sc["cam5"], sc["bounds"], 8, float(g[f"fp_{tag}_log_scale"]))
        c[f"fp_{tag}"] = ((q, mp["desc"], kf["desc"], kf["x"], kf["y"], kf["oct"], kf["uright"], sc["bounds"], sc["sf"], sc["inv_sigma2"], 3.0),
                          int(g[f"fp_{tag}_n"]), g[f"fp_{tag}_idx"])
    sc = fuse_lines_scenario(oracle, synth, 2, f0=2)
    ml, kf = sc["ml"], sc["kf"]
    q = oracle.fuse_project_lines(g["fl_state"] != 1, ml["Pw"], ml["normal"], g["fl_min_inv"], g["fl_max_inv"], ml["max_raw"], sc["Tcw"][:3], g["fl_Ow"],
                                  sc["cam5"], sc["bounds"], 8, float(g["fl_log_scale"]))
    c["fl"] = ((q, ml["desc"], kf["ld"], kf["kl"], kf["oct"], sc["sf"], 10.0), int(g["fl_n"]), g["fl_idx"])
    return c


def fused_from(q, bi, bd):
    return np.where((np.asarray(q["active"]) != 0) & (bd <= 50) & (bi >= 0), bi, -1)


def test_oracle_row3_equals_reference_fixtures(oracle, synth):
    c = row3_cases(oracle, synth)
    for k in ("lpf", "lpm"):
        args, n, a = c[k]
        n_o, a_o = oracle.line_window_search(*args)
        assert n_o == n and np.array_equal(a_o, a), k
    for k in ("fp_mono", "fp_stereo"):
        args, n, f = c[k]
        f_o = fused_from(args[0], *oracle.fuse_points_search(*args))
        assert np.array_equal(f_o, f) and n == (f >= 0).sum(), k
    args, n, f = c["fl"]
    f_o = fused_from(args[0], *oracle.fuse_lines_search(*args))
    assert np.array_equal(f_o, f) and n == (f >= 0).sum()
