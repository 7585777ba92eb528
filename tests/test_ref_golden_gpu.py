"""The CUDA path (through the C ABI) against fixtures frozen from THE REFERENCE ITSELF (tests/golden/ref_*.npz, written by
tools/make_ref_golden.py from oracle/_ref/libref.so = the reference's sources compiled unmodified): no oracle in between."""
import numpy as np
import pytest

from test_ref_golden_cpu import load, frames, match_inputs

pytestmark = pytest.mark.gpu


def test_gpu_orb_equals_reference_fixtures(pkg, icl_gray, synth):
    g = load("ref_orb.npz")
    for tag, (img, nf) in frames(icl_gray, synth).items():
        ext = pkg.ORBextractor(nf, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=1)
        k, d = ext(img)
        assert k.tobytes() == g[f"orb_{tag}_kps"].tobytes() and np.array_equal(d, g[f"orb_{tag}_desc"]), tag


def test_gpu_matchers_equal_reference_fixtures(pkg, oracle, synth):
    g = load("ref_match.npz")
    orc, k1, d1, k2, d2, fv1, fv2 = match_inputs(oracle, synth)       # (inputs only: extraction is covered above)
    v1, v2 = g["valid1"], g["valid2"]
    ctx = pkg.Matcher(max_features=2048, max_lines=64, max_nodes=3072)
    for ratio, ori in [(0.7, True), (0.9, False)]:
        m = pkg.ORBmatcher(ratio, ori, ctx)
        n, mm = m.SearchByBoW(d1, fv1, v1, k1["angle"], d2, fv2, k2["angle"])
        assert n == int(g[f"bow_{ratio}_{int(ori)}_n"]) and np.array_equal(mm, g[f"bow_{ratio}_{int(ori)}"])
        n, mm = m.SearchByBoW(d1, fv1, v1, k1["angle"], d2, fv2, k2["angle"], valid2=v2)
        assert n == int(g[f"bowkf_{ratio}_{int(ori)}_n"]) and np.array_equal(mm, g[f"bowkf_{ratio}_{int(ori)}"])
    tb = orc.tables()
    m = pkg.ORBmatcher(0.6, True, ctx)
    for tag in ("in", "out"):
        ex, ey = g[f"tri_{tag}_epi"]
        n, p = m.SearchForTriangulation(d1, fv1, 1 - v1, k1, d2, fv2, 1 - v2, k2, g[f"tri_{tag}_F12"], ex, ey, tb["scale"], tb["sigma2"])
        assert np.array_equal(p, g[f"tri_{tag}_pairs"]) and n == len(p)
    l1, l2, h1, h2 = g["line_d1"], g["line_d2"], g["line_h1"], g["line_h2"]
    lm = pkg.LSDmatcher(ctx)
    for mode in range(4):
        n, mm = lm._run(mode, l1, l2, h1, h2)
        assert n == int(g[f"line_mode{mode}_n"]) and np.array_equal(mm, g[f"line_mode{mode}"]), mode
        assert lm.last_mad == tuple(g["line_mad"])
    from scenarios import projection_scenario
    for seed, th, mono in [(1, 15.0, True), (4, 15.0, False)]:
        last, cur, Tcw, Tlw, cam, bounds, sf = projection_scenario(oracle, synth, seed, n_claimed=0.05, stereo=not mono, f0=seed)
        n, a = ctx.search_by_projection_frame(last, cur, Tcw, Tlw, cam, bounds, sf, th, mono, True)
        assert n == int(g[f"proj_{seed}_n"]) and np.array_equal(a, g[f"proj_{seed}"]), seed


def test_gpu_frame_lines_equal_reference_fixtures(pkg, icl_gray):
    g = load("ref_frame.npz")
    ls = pkg.LineSegment(40, max_width=640, max_height=480)
    kl, ld, eq = ls.ExtractLineSegment(icl_gray)
    gk = g["keylines"]
    assert len(kl) == len(gk) and np.array_equal(ld, g["ldesc"])
    for fld in ("class_id", "octave", "numOfPixels"):
        assert np.array_equal(kl[fld], gk[fld]), fld
    for fld in ("startPointX", "startPointY", "endPointX", "endPointY", "pt_x", "pt_y"):
        assert np.max(np.abs(kl[fld] - gk[fld])) <= 1e-4, fld
    assert np.max(np.abs(kl["angle"] - gk["angle"])) <= 1e-3 and np.allclose(eq, g["lineeq"], rtol=1e-9, atol=1e-9)


def test_gpu_row3_equals_reference_fixtures(pkg, oracle, synth):
    """Line projection matchers and Fuse: the device search stage on the frozen scenarios (the oracle only
    prepares the projection-stage inputs, as the adapter does with the reference's own accessors) against the reference's results."""
    from test_ref_golden_cpu import row3_cases, fused_from
    c = row3_cases(oracle, synth)
    mt = pkg.Matcher(max_features=2048, max_lines=512, max_nodes=3072)
    for k in ("lpf", "lpm"):
        args, n, a = c[k]
        n_g, a_g = mt.line_search_by_projection(*args)
        assert n_g == n and np.array_equal(a_g, a), k
    for k in ("fp_mono", "fp_stereo"):
        args, n, f = c[k]
        f_g = fused_from(args[0], *mt.fuse_points_search(*args))
        assert np.array_equal(f_g, f) and n == (f >= 0).sum(), k
    args, n, f = c["fl"]
    f_g = fused_from(args[0], *mt.fuse_lines_search(*args))
    assert np.array_equal(f_g, f) and n == (f >= 0).sum()
