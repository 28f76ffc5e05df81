"""CPU: the port restatement of the matchers against the golden outputs of the reference's own bodies
(tests/golden/ref_match.npz) and, where built, against oracle/_ref live.  Match indices bit-exact."""
import os
import numpy as np
import pytest
import match_cases as mc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match.npz"))


@pytest.fixture(scope="module")
def case(oracle):
    P = oracle.Extractor("port", 1000)
    ka, da, kb, db = mc.mono_pair(P.extract)
    assert len(ka) == int(G["n_a"]) and len(kb) == int(G["n_b"])
    pi = mc.projection_inputs(ka, kb)
    sf = P.scale_factors
    return dict(ka=ka, da=da, kb=kb, db=db, pi=pi, sf=sf,
                FA=oracle.FrameData(ka, da, 640, 480, sf), FB=oracle.FrameData(kb, db, 640, 480, sf),
                FBu=oracle.FrameData(kb, db, 640, 480, sf, u_right=pi["u_right"]))


def test_descriptor_distance(oracle, case):
    m = oracle.Matcher("port")
    d = m.descriptor_distance(case["da"][:900], case["db"][:900])
    assert np.array_equal(d, G["dist"])
    assert np.array_equal(d, np.unpackbits(case["da"][:900] ^ case["db"][:900], axis=1).sum(1))
    z = np.zeros((1, 32), np.uint8); o = np.full((1, 32), 255, np.uint8)
    assert m.descriptor_distance(z, z)[0] == 0 and m.descriptor_distance(z, o)[0] == 256


def test_search_for_initialization(oracle, case):
    prev = np.stack([case["ka"]["x"], case["ka"]["y"]], 1)
    nm, m12, prev2 = oracle.Matcher("port", 0.9, True).search_for_initialization(case["FA"], case["FB"], prev, 100)
    assert nm == int(G["init_nm"]) and np.array_equal(m12, G["init_m12"]) and np.array_equal(prev2, G["init_prev"])
    assert nm == (m12 >= 0).sum() and nm > 50
    assert (case["ka"]["octave"][m12 >= 0] == 0).all()                     # only level-0 keypoints of F1 are matched (:537)
    nm, m12, prev2 = oracle.Matcher("port", 0.9, False).search_for_initialization(case["FA"], case["FB"], prev, 30)
    assert nm == int(G["init2_nm"]) and np.array_equal(m12, G["init2_m12"]) and np.array_equal(prev2, G["init2_prev"])


def test_search_by_projection(oracle, case):
    pi = case["pi"]; ka = case["ka"]
    uv, iz = mc.project(pi["xyz"])
    for i, (th, mono) in enumerate(mc.PROJ_FRAME_CASES):
        nm, cm = oracle.Matcher("port", 0.9, True).search_by_projection_frame_port(case["FBu"], uv, iz, ka["octave"], ka["angle"], case["da"], pi["valid"], pi["obs"], pi["occ"], th, 0, 0, 40.0)
        assert nm == int(G["pf%d_nm" % i]) and np.array_equal(cm, G["pf%d_cm" % i])
    for i, th in enumerate(mc.PROJ_POINT_CASES):
        nm, fm = oracle.Matcher("port", 0.8, True).search_by_projection_points(case["FBu"], pi["tuv"], pi["tur"], pi["lvl"], pi["vc"], case["da"], pi["obs"], pi["occ"], th)
        assert nm == int(G["pp%d_nm" % i]) and np.array_equal(fm, G["pp%d_fm" % i])


def test_compute_stereo_matches(oracle):
    L, R = mc.stereo_pair()
    PL, PR = oracle.Extractor("port", 2000), oracle.Extractor("port", 2000)
    kl, dl = PL.extract(L); kr, dr = PR.extract(R)
    assert len(kl) == int(G["stereo_nl"])
    m = oracle.Matcher("port")
    ur, dep = m.compute_stereo_matches(PL, PR, kl, dl, kr, dr, 0.0, mc.BF_KITTI)           # mb == 0 inside the stereo constructor (Frame.cc:131)
    assert np.array_equal(ur, G["stereo_ur"]) and np.array_equal(dep, G["stereo_depth"])
    ok = ur >= 0
    assert ok.sum() > 300
    disp = kl["x"][ok] - ur[ok]
    assert (disp > 0).all() and np.median(np.abs(disp - np.rint(disp))) < 0.35            # piecewise-constant integer disparities recovered
    ur2, dep2 = m.compute_stereo_matches(PL, PR, kl, dl, kr, dr, 0.5372, mc.BF_KITTI)
    assert np.array_equal(ur2, G["stereo2_ur"]) and np.array_equal(dep2, G["stereo2_depth"])


def test_grid_queries_and_live_reference(oracle, case):
    mp = oracle.Matcher("port")
    rng = np.random.default_rng(5)
    qs = [(float(rng.uniform(-50, 700)), float(rng.uniform(-50, 530)), float(rng.uniform(1, 150)), int(rng.integers(-1, 4)), int(rng.integers(-1, 8))) for _ in range(60)]
    res = [mp.features_in_area(case["FA"], *q) for q in qs]
    assert sum(len(r) for r in res) > 100
    # brute-force definition of the same set (order aside)
    ka = case["ka"]
    for q, r in zip(qs[:20], res[:20]):
        x, y, rad, a, b = q
        ok = (np.abs(ka["x"] - np.float32(x)) < rad) & (np.abs(ka["y"] - np.float32(y)) < rad)
        if a > 0 or b >= 0:
            ok &= ka["octave"] >= a
            if b >= 0:
                ok &= ka["octave"] <= b
        assert set(r.tolist()) <= set(np.nonzero(ok)[0].tolist())
    if oracle.have_ref():
        mr = oracle.Matcher("ref")
        for q, r in zip(qs, res):
            assert np.array_equal(mr.features_in_area(case["FA"], *q), r)          # same candidates in the same ORDER


def test_port_keyframe_projection_matches_reference_golden(oracle):
    """SearchByProjection(Frame&, KeyFrame*, set&, th, ORBdist) (ORBmatcher.cc:1731-1863): port vs the reference body's committed outputs."""
    import os
    GK = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match_kf.npz"))
    E = oracle.Extractor("port", 1000, 1.2, 8, 20, 7)
    ka, da, kb, db = mc.mono_pair(lambda img: E.extract(img))
    assert len(ka) == int(GK["n_a"]) and len(kb) == int(GK["n_b"])
    pi = mc.projection_inputs(ka, kb); kf = mc.keyframe_inputs(ka, kb, pi)
    F = oracle.FrameData(kb, db, 640, 480, E.scale_factors)
    uv, _ = mc.project(pi["xyz"])
    assert np.array_equal(uv, GK["uv"])                                   # the caller-side projection equals the body's own
    for i, (th, od, ori) in enumerate(mc.KF_CASES):
        nm, cm = oracle.Matcher("port", 0.9, ori).search_by_projection_keyframe_port(F, uv, kf["lvl"], ka["angle"], da, kf["valid"], kf["occ"], th, od)
        assert nm == int(GK["kf%d_nm" % i]) and nm > 30 and np.array_equal(cm, GK["kf%d_cm" % i])


def test_port_keyframe_points_projection_matches_reference_golden(oracle):
    """SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (ORBmatcher.cc:388-512): port vs the reference body's committed outputs."""
    import os
    GK = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match_kf.npz"))
    E = oracle.Extractor("port", 1000, 1.2, 8, 20, 7)
    ka, da, kb, db = mc.mono_pair(lambda img: E.extract(img))
    pi = mc.projection_inputs(ka, kb); kp = mc.keyframe_points_inputs(ka, kb, pi)
    assert np.array_equal(kp["uv"], GK["uv2"])
    F = oracle.FrameData(kb, db, 640, 480, E.scale_factors)
    for i, th in enumerate(mc.KFP_CASES):
        nm, km = oracle.Matcher("port").search_by_projection_keyframe_points_port(F, kp["uv"], kp["lvl"], da, kp["valid"], kp["kf_matched"], th)
        assert nm == int(GK["kfp%d_nm" % i]) and nm > 30 and np.array_equal(km, GK["kfp%d_km" % i])


def test_port_search_by_sim3_matches_reference_golden(oracle):
    """SearchBySim3 (ORBmatcher.cc:1290-1555): port vs the reference body's committed outputs."""
    import os
    GK = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match_kf.npz"))
    E = oracle.Extractor("port", 1000, 1.2, 8, 20, 7)
    ka, da, kb, db = mc.mono_pair(lambda img: E.extract(img))
    s1, s2, _ = mc.sim3_inputs(ka, kb)
    F1 = oracle.FrameData(ka, da, 640, 480, E.scale_factors); F2 = oracle.FrameData(kb, db, 640, 480, E.scale_factors)
    for i, th in enumerate(mc.SIM3_TH):
        nf, m12 = oracle.Matcher("port").search_by_sim3_port(F1, F2, s1["uv"], s1["lvl"], da, s1["valid"], s2["uv"], s2["lvl"], db, s2["valid"], th)
        assert nf == int(GK["sim3_%d_nf" % i]) and nf > 5 and np.array_equal(m12, GK["sim3_%d_m12" % i])


def test_port_fuse_search_matches_reference_golden(oracle):
    """Both ORBmatcher::Fuse forms (ORBmatcher.cc:1020-1310): the feature each point is fused with, port vs the reference bodies' committed outputs."""
    import os
    GK = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match_kf.npz"))
    E = oracle.Extractor("port", 1000, 1.2, 8, 20, 7)
    ka, da, kb, db = mc.mono_pair(lambda img: E.extract(img))
    pi = mc.projection_inputs(ka, kb); fu = mc.fuse_inputs(ka, kb, pi)
    assert np.array_equal(fu["uv"], GK["fuse_uv"]) and np.array_equal(fu["ur"], GK["fuse_ur"])
    F = oracle.FrameData(kb, db, 640, 480, E.scale_factors, u_right=pi["u_right"])
    for i, th in enumerate(mc.FUSE_TH):
        for sim3 in (0, 1):
            best = oracle.Matcher("port").fuse_search_port(F, fu["uv"], None if sim3 else fu["ur"], fu["lvl"], da, fu["valid"], fu["inv_sigma2"], th)
            assert np.array_equal(best, GK["fuse_%d_%d_best" % (i, sim3)]) and int((best >= 0).sum()) == int(GK["fuse_%d_%d_nf" % (i, sim3)]) > 10
