"""GPU (B200): ORBmatcher / ComputeStereoMatches through the C ABI against the reference's golden outputs and the port
oracle on the same seeded inputs.  Match indices, counts, uRight and depth bit-exact."""
import os
import numpy as np
import pytest
import match_cases as mc

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match.npz"))


@pytest.fixture(scope="module")
def case(orbx):
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    ka, da, kb, db = mc.mono_pair(lambda img: E(img))
    assert len(ka) == int(G["n_a"]) and len(kb) == int(G["n_b"])
    pi = mc.projection_inputs(ka, kb)
    sf = E.GetScaleFactors()
    return dict(ka=ka, da=da, kb=kb, db=db, pi=pi, sf=sf, FA=orbx.FrameView(ka, da, 640, 480, sf), FB=orbx.FrameView(kb, db, 640, 480, sf),
                FBu=orbx.FrameView(kb, db, 640, 480, sf, u_right=pi["u_right"]))


def test_descriptor_distance(orbx, case):
    m = orbx.ORBmatcher(0.9, True)
    assert np.array_equal(m.DescriptorDistance(case["da"][:900], case["db"][:900]), G["dist"])
    z = np.zeros((3, 32), np.uint8); o = np.full((3, 32), 255, np.uint8)
    assert list(m.DescriptorDistance(z, o)) == [256, 256, 256] and list(m.DescriptorDistance(o, o)) == [0, 0, 0]


def test_search_for_initialization(orbx, case):
    prev = np.stack([case["ka"]["x"], case["ka"]["y"]], 1)
    nm, m12, prev2 = orbx.ORBmatcher(0.9, True).SearchForInitialization(case["FA"], case["FB"], prev, 100)
    assert nm == int(G["init_nm"]) and np.array_equal(m12, G["init_m12"]) and np.array_equal(prev2, G["init_prev"])
    nm, m12, prev2 = orbx.ORBmatcher(0.9, False).SearchForInitialization(case["FA"], case["FB"], prev, 30)
    assert nm == int(G["init2_nm"]) and np.array_equal(m12, G["init2_m12"]) and np.array_equal(prev2, G["init2_prev"])


def test_search_by_projection_frame_and_points(orbx, case):
    pi = case["pi"]; ka = case["ka"]
    uv, iz = mc.project(pi["xyz"])
    for i, (th, mono) in enumerate(mc.PROJ_FRAME_CASES):
        nm, cm = orbx.ORBmatcher(0.9, True).SearchByProjectionFrame(case["FBu"], uv, iz, ka["octave"], ka["angle"], case["da"], pi["valid"], pi["obs"], pi["occ"], th, False, False, 40.0)
        assert nm == int(G["pf%d_nm" % i]) and np.array_equal(cm, G["pf%d_cm" % i])
    for i, th in enumerate(mc.PROJ_POINT_CASES):
        nm, fm = orbx.ORBmatcher(0.8, True).SearchByProjectionPoints(case["FBu"], pi["tuv"], pi["tur"], pi["lvl"], pi["vc"], case["da"], pi["obs"], pi["occ"], th)
        assert nm == int(G["pp%d_nm" % i]) and np.array_equal(fm, G["pp%d_fm" % i])


def test_search_by_projection_frame_pose_projects_on_the_device(orbx, case):
    """The pose form (world points + Rcw, tcw in, projection on the device) must give what the (u, v, 1/z) form gives when the caller
    projects with OpenCV's arithmetic: taps equal the float-path restatement bit for bit, matches equal the pre-projected call."""
    pi = case["pi"]; ka = case["ka"]
    Rcw, tcw = mc.small_pose()
    Rf = Rcw.astype(np.float64); xyz_w = ((pi["xyz"].astype(np.float64) - tcw.astype(np.float64)) @ Rf).astype(np.float32)     # world points that land near pi["xyz"]
    xyz_w[5::41, 2] = -xyz_w[5::41, 2]                                      # some behind the camera
    xyz_w[7::53] *= 40                                                      # some far outside the image
    for FBg in (case["FBu"], case["FB"]):
        b = (0.0, 640.0, 0.0, 480.0)                                         # bounds of the undistorted 640 x 480 test frames (match_cases FrameView default)
        uv, iz, va = mc.project_pose(xyz_w, Rcw, tcw, b)
        va = (va & (pi["valid"] != 0)).astype(np.uint8); uv = uv.copy(); uv[va == 0] = 0; iz = iz.copy(); iz[va == 0] = 0
        M = orbx.ORBmatcher(0.9, True)
        for (th, fw, bw) in [(15.0, 0, 0), (7.0, 1, 0), (15.0, 0, 1)]:
            nm, cm, guv, giz, gva = M.SearchByProjectionFramePose(FBg, xyz_w, pi["valid"], Rcw, tcw, mc.FX, mc.FY, mc.CX, mc.CY, ka["octave"], ka["angle"], case["da"],
                                                                  pi["obs"], pi["occ"], th, fw, bw, 40.0, taps=True)
            assert np.array_equal(gva, va) and np.array_equal(guv, uv) and np.array_equal(giz, iz)
            nm2, cm2 = M.SearchByProjectionFrame(FBg, uv, iz, ka["octave"], ka["angle"], case["da"], va, pi["obs"], pi["occ"], th, fw, bw, 40.0)
            assert nm == nm2 and np.array_equal(cm, cm2) and nm > 50
    # the device-Frame form takes the same path
    assert int(va.sum()) > 300 and int((pi["valid"] != 0).sum()) > int(va.sum())


def test_projection_variants_vs_port(orbx, oracle, case):
    """forward / backward level ranges, no uRight, no observations, checkOri off."""
    pi = case["pi"]; ka = case["ka"]
    uv, iz = mc.project(pi["xyz"])
    oFB = oracle.FrameData(case["kb"], case["db"], 640, 480, case["sf"]); oFBu = oracle.FrameData(case["kb"], case["db"], 640, 480, case["sf"], u_right=pi["u_right"])
    for (th, fw, bw, ori, FBg, FBo, obs, occ) in [(15.0, 1, 0, True, case["FBu"], oFBu, pi["obs"], pi["occ"]), (7.0, 0, 1, True, case["FBu"], oFBu, pi["obs"], None),
                                                  (15.0, 0, 0, False, case["FB"], oFB, np.zeros_like(pi["obs"]), pi["occ"]), (30.0, 0, 0, True, case["FB"], oFB, np.ones_like(pi["obs"]), None)]:
        g = orbx.ORBmatcher(0.9, ori).SearchByProjectionFrame(FBg, uv, iz, ka["octave"], ka["angle"], case["da"], pi["valid"], obs, occ, th, fw, bw, 40.0)
        o = oracle.Matcher("port", 0.9, ori).search_by_projection_frame_port(FBo, uv, iz, ka["octave"], ka["angle"], case["da"], pi["valid"], obs, occ, th, fw, bw, 40.0)
        assert g[0] == o[0] and np.array_equal(g[1], o[1])
        g = orbx.ORBmatcher(0.7, ori).SearchByProjectionPoints(FBg, pi["tuv"], pi["tur"], pi["lvl"], pi["vc"], case["da"], obs, occ, th / 5.0)
        o = oracle.Matcher("port", 0.7, ori).search_by_projection_points(FBo, pi["tuv"], pi["tur"], pi["lvl"], pi["vc"], case["da"], obs, occ, th / 5.0)
        assert g[0] == o[0] and np.array_equal(g[1], o[1])


def test_compute_stereo_matches(orbx):
    L, R = mc.stereo_pair()
    EL, ER = orbx.ORBextractor(2000, 1.2, 8, 20, 7), orbx.ORBextractor(2000, 1.2, 8, 20, 7)
    kl, dl = EL(L); kr, dr = ER(R)
    assert len(kl) == int(G["stereo_nl"])
    m = orbx.ORBmatcher()
    ur, dep = m.ComputeStereoMatches(EL, ER, kl, dl, kr, dr, 0.0, mc.BF_KITTI)
    assert np.array_equal(ur, G["stereo_ur"]) and np.array_equal(dep, G["stereo_depth"])
    ur2, dep2 = m.ComputeStereoMatches(EL, ER, kl, dl, kr, dr, 0.5372, mc.BF_KITTI)
    assert np.array_equal(ur2, G["stereo2_ur"]) and np.array_equal(dep2, G["stereo2_depth"])


def test_empty_and_degenerate(orbx, case):
    m = orbx.ORBmatcher(0.9, True)
    E0 = orbx.FrameView(np.zeros(0, orbx.KP_DTYPE), np.zeros((0, 32), np.uint8), 640, 480, case["sf"])
    nm, m12, _ = m.SearchForInitialization(E0, case["FB"], np.zeros((0, 2), np.float32), 100)
    assert nm == 0 and len(m12) == 0
    prev = np.stack([case["ka"]["x"], case["ka"]["y"]], 1)
    nm, m12, _ = m.SearchForInitialization(case["FA"], E0, prev, 100)                      # nothing to match against
    assert nm == 0 and (m12 == -1).all()
    nm, cm = m.SearchByProjectionFrame(case["FB"], np.zeros((0, 2), np.float32), np.zeros(0, np.float32), np.zeros(0, np.int32), np.zeros(0, np.float32),
                                       np.zeros((0, 32), np.uint8), np.zeros(0, np.uint8), None, None, 15.0)
    assert nm == 0 and (cm == -1).all()
    # identical frames: every level-0 keypoint matches itself with distance 0
    nm, m12, _ = m.SearchForInitialization(case["FA"], case["FA"], prev, 10)
    lv0 = case["ka"]["octave"] == 0
    assert (m12[~lv0] == -1).all() and nm > 0.9 * lv0.sum() and (m12[m12 >= 0] == np.nonzero(m12 >= 0)[0]).all()


def test_bruteforce_best2_properties(orbx, case):
    """All-pairs best / second-best (device pointers): against numpy on the full 1000 x 1000 problem."""
    import torch
    m = orbx.ORBmatcher()
    q = torch.from_numpy(case["da"]).cuda(); t = torch.from_numpy(case["db"]).cuda()
    bi = torch.empty(len(q), dtype=torch.int32, device="cuda"); bd = torch.empty_like(bi); sd = torch.empty_like(bi)
    m.match_bruteforce_device(q.data_ptr(), len(q), t.data_ptr(), len(t), bi.data_ptr(), bd.data_ptr(), sd.data_ptr())
    torch.cuda.synchronize()
    D = np.unpackbits(case["da"][:, None, :] ^ case["db"][None, :, :], axis=2).sum(2)
    assert np.array_equal(bi.cpu().numpy(), D.argmin(1)) and np.array_equal(bd.cpu().numpy(), D.min(1))     # first index wins ties
    assert np.array_equal(sd.cpu().numpy(), np.sort(D, axis=1)[:, 1])


def test_bruteforce_batch_pairs(orbx, case):
    """n_pairs frame pairs in one launch == the single-pair call per pair."""
    import torch
    m = orbx.ORBmatcher()
    rng = np.random.default_rng(5)
    P, nq, nt = 3, 257, 300
    Q = rng.integers(0, 256, (P, nq, 32), dtype=np.uint8); T = rng.integers(0, 256, (P, nt, 32), dtype=np.uint8)
    T[1, 7] = Q[1, 100]; T[1, 9] = Q[1, 100]                                     # exact duplicates: first index wins, second distance 0
    q = torch.from_numpy(Q).cuda(); t = torch.from_numpy(T).cuda()
    bi = torch.empty((P, nq), dtype=torch.int32, device="cuda"); bd = torch.empty_like(bi); sd = torch.empty_like(bi)
    m.match_bruteforce_batch_device(P, q.data_ptr(), nq, t.data_ptr(), nt, bi.data_ptr(), bd.data_ptr(), sd.data_ptr())
    torch.cuda.synchronize()
    for p in range(P):
        D = np.unpackbits(Q[p][:, None, :] ^ T[p][None, :, :], axis=2).sum(2)
        assert np.array_equal(bi[p].cpu().numpy(), D.argmin(1)) and np.array_equal(bd[p].cpu().numpy(), D.min(1))
        assert np.array_equal(sd[p].cpu().numpy(), np.sort(D, axis=1)[:, 1])
    assert bi[1, 100].item() == 7 and bd[1, 100].item() == 0 and sd[1, 100].item() == 0


def test_search_by_projection_keyframe(orbx, case):
    """SearchByProjection(Frame&, KeyFrame*, set&, th, ORBdist) (relocalisation) against the reference body's golden, on a host view and on a device frame."""
    GK = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match_kf.npz"))
    ka, kb, pi = case["ka"], case["kb"], case["pi"]
    kf = mc.keyframe_inputs(ka, kb, pi)
    uv, _ = mc.project(pi["xyz"])
    D = orbx.Frame().assign_host(kb, case["db"], case["sf"], orbx.Camera.make(mc.FX, mc.FY, mc.CX, mc.CY), 480, 640)
    D.set_stereo(pi["u_right"], np.where(pi["u_right"] > 0, 1.0, -1.0).astype(np.float32))          # must be ignored by this matcher
    for i, (th, od, ori) in enumerate(mc.KF_CASES):
        for F in (case["FB"], case["FBu"], D):
            nm, cm = orbx.ORBmatcher(0.9, ori).SearchByProjectionKeyFrame(F, uv, kf["lvl"], ka["angle"], case["da"], kf["valid"], kf["occ"], th, od)
            assert nm == int(GK["kf%d_nm" % i]) and np.array_equal(cm, GK["kf%d_cm" % i])


def test_search_by_projection_keyframe_points(orbx, case):
    """SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (loop closing) against the reference body's golden; the rotation check and
    uRight of the frame must play no role."""
    GK = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match_kf.npz"))
    ka, kb, pi = case["ka"], case["kb"], case["pi"]
    kp = mc.keyframe_points_inputs(ka, kb, pi)
    D = orbx.Frame().assign_host(kb, case["db"], case["sf"], orbx.Camera.make(mc.FX, mc.FY, mc.CX, mc.CY), 480, 640)
    for i, th in enumerate(mc.KFP_CASES):
        for F in (case["FB"], case["FBu"], D):
            for ori in (True, False):
                nm, km = orbx.ORBmatcher(0.9, ori).SearchByProjectionKeyFramePoints(F, kp["uv"], kp["lvl"], case["da"], kp["valid"], kp["kf_matched"], th)
                assert nm == int(GK["kfp%d_nm" % i]) and np.array_equal(km, GK["kfp%d_km" % i])


def test_search_by_sim3(orbx, case):
    """SearchBySim3 (loop closing): both passes and the mutual check on the device, against the reference body's golden."""
    GK = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match_kf.npz"))
    ka, kb = case["ka"], case["kb"]
    s1, s2, _ = mc.sim3_inputs(ka, kb)
    for i, th in enumerate(mc.SIM3_TH):
        nf, m12 = orbx.ORBmatcher(0.6, True).SearchBySim3(case["FA"], case["FB"], s1["uv"], s1["lvl"], case["da"], s1["valid"], s2["uv"], s2["lvl"], case["db"], s2["valid"], th)
        assert nf == int(GK["sim3_%d_nf" % i]) and np.array_equal(m12, GK["sim3_%d_m12" % i])
    e = orbx.FrameView(kb[:0], case["db"][:0], 640, 480, case["sf"])
    nf, m12 = orbx.ORBmatcher().SearchBySim3(case["FA"], e, s1["uv"], s1["lvl"], case["da"], s1["valid"], s2["uv"][:0], s2["lvl"][:0], case["db"][:0], s2["valid"][:0], 7.5)
    assert nf == 0 and np.all(m12 == -1)


def test_fuse_search(orbx, case):
    """The search of both ORBmatcher::Fuse forms (chi-square gated / plain) against what the reference bodies fused each point with."""
    GK = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match_kf.npz"))
    ka, kb, pi = case["ka"], case["kb"], case["pi"]
    fu = mc.fuse_inputs(ka, kb, pi)
    for i, th in enumerate(mc.FUSE_TH):
        for sim3 in (0, 1):
            best = orbx.ORBmatcher().FuseSearch(case["FBu"], fu["uv"], None if sim3 else fu["ur"], fu["lvl"], case["da"], fu["valid"], fu["inv_sigma2"], th)
            assert np.array_equal(best, GK["fuse_%d_%d_best" % (i, sim3)])


def test_search_for_initialization_batch(orbx, oracle, case):
    """P pairs per call == P single calls: pair 0 is the reference golden (ref_match.npz), the others are ragged pairs (different
    feature counts, an empty F1, an empty F2, windows that overflow the default candidate slice) checked against the single-pair
    call and the oracle.  Host views and device-resident frames."""
    from tools.synth import synth_frame, warp_affine_nn
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7); E2 = orbx.ORBextractor(300, 1.2, 8, 20, 7)
    sf = case["sf"]
    pairs = [(case["ka"], case["da"], case["kb"], case["db"])]
    for s in range(5):
        ext = E if s % 2 == 0 else E2
        A = synth_frame(800 + s, 640, 480); B = warp_affine_nn(A, 3 + s, -2 - s, 1.0 + 0.5 * s)
        ka, da = ext(A); kb, db = ext(B)
        pairs.append((ka, da, kb, db))
    empty_k, empty_d = case["ka"][:0], case["da"][:0]
    pairs.append((empty_k, empty_d, case["kb"], case["db"]))
    pairs.append((case["ka"], case["da"], empty_k, empty_d))
    F1 = [orbx.FrameView(p[0], p[1], 640, 480, sf) for p in pairs]; F2 = [orbx.FrameView(p[2], p[3], 640, 480, sf) for p in pairs]
    prevs = [np.stack([p[0]["x"], p[0]["y"]], 1).astype(np.float32).reshape(-1, 2) for p in pairs]
    for (nn, ori, win) in ((0.9, True, 100), (0.9, False, 30), (0.7, True, 400)):       # window 400: lists far beyond the default slice -> grow + retry
        M = orbx.ORBmatcher(nn, ori)
        nm, m12, prev2 = M.SearchForInitializationBatch(F1, F2, prevs, win)
        for p in range(len(pairs)):
            n1, a, b = orbx.ORBmatcher(nn, ori).SearchForInitialization(F1[p], F2[p], prevs[p], win)
            assert nm[p] == n1 and np.array_equal(m12[p], a) and np.array_equal(prev2[p], b), (p, win)
            o = oracle.Matcher("port", nn, ori).search_for_initialization(oracle.FrameData(pairs[p][0], pairs[p][1], 640, 480, sf), oracle.FrameData(pairs[p][2], pairs[p][3], 640, 480, sf), prevs[p], win)
            assert nm[p] == o[0] and np.array_equal(m12[p], o[1]) and np.array_equal(prev2[p], o[2]), (p, win)
    nm, m12, prev2 = orbx.ORBmatcher(0.9, True).SearchForInitializationBatch(F1, F2, prevs, 100)
    assert nm[0] == int(G["init_nm"]) and np.array_equal(m12[0], G["init_m12"]) and np.array_equal(prev2[0], G["init_prev"])
    # device-resident frames (grid already built: the batch skips the grid launch work)
    cam = orbx.Camera.make(500.0, 500.0, 320.0, 240.0, 0, 0, 0, 0, 0, 40.0)
    D1 = [orbx.Frame().assign_host(p[0], p[1], sf, cam, 480, 640) for p in pairs[:6]]; D2 = [orbx.Frame().assign_host(p[2], p[3], sf, cam, 480, 640) for p in pairs[:6]]
    nm_d, m12_d, prev_d = orbx.ORBmatcher(0.9, True).SearchForInitializationBatch(D1, D2, prevs[:6], 100)
    for p in range(6):
        assert nm_d[p] == nm[p] and np.array_equal(m12_d[p], m12[p]) and np.array_equal(prev_d[p], prev2[p]), p
    assert orbx.ORBmatcher(0.9, True).SearchForInitializationBatch([], [], [], 100)[0].size == 0


def test_compute_stereo_matches_batch(orbx):
    """B stereo pairs per call (config C4's shard unit) == B single calls; pair 0 is the reference golden.  Host-pointer batch extraction
    on both handles, then ComputeStereoMatches on what they still hold on the device; also through the device-batch extraction."""
    import torch
    B = 4
    pairs = [mc.stereo_pair()] + [mc.stereo_pair(seed=11 + 2 * s) for s in range(B - 1)]
    Ls = np.stack([p[0] for p in pairs]); Rs = np.stack([p[1] for p in pairs])
    EL, ER = orbx.ORBextractor(2000, 1.2, 8, 20, 7), orbx.ORBextractor(2000, 1.2, 8, 20, 7)
    kl, dl, cl = EL.extract_batch(Ls); kr, dr, cr = ER.extract_batch(Rs)
    cap = kl.shape[1]
    m = orbx.ORBmatcher()
    for mb, tag in ((0.0, "stereo"), (0.5372, "stereo2")):
        ur, dep = m.ComputeStereoMatchesBatch(EL, ER, B, cap, mb, mc.BF_KITTI)
        assert np.array_equal(ur[0][:cl[0]], G[tag + "_ur"]) and np.array_equal(dep[0][:cl[0]], G[tag + "_depth"])
        E1, E2 = orbx.ORBextractor(2000, 1.2, 8, 20, 7), orbx.ORBextractor(2000, 1.2, 8, 20, 7)
        for b in range(B):
            k1, d1 = E1(pairs[b][0]); k2, d2 = E2(pairs[b][1])
            u1, p1 = orbx.ORBmatcher().ComputeStereoMatches(E1, E2, k1, d1, k2, d2, mb, mc.BF_KITTI)
            assert cl[b] == len(k1) and np.array_equal(ur[b][:cl[b]], u1) and np.array_equal(dep[b][:cl[b]], p1), b
            assert (ur[b][cl[b]:] == -1).all() and (dep[b][cl[b]:] == -1).all()
        assert ((ur >= 0).sum(1) > 200).all()
    # device-resident form: frames in HBM, extraction and matching queued without a host round trip
    dL = torch.from_numpy(Ls).cuda(); dR = torch.from_numpy(Rs).cuda()
    outs = []
    for E, d in ((EL, dL), (ER, dR)):
        k = torch.empty((B, cap, 28), dtype=torch.uint8, device="cuda"); ds = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda"); c = torch.empty(B, dtype=torch.int32, device="cuda")
        E.extract_batch_raw(d.data_ptr(), B, 376, 1241, 1241, 1241 * 376, k.data_ptr(), ds.data_ptr(), cap, c.data_ptr(), device=True)
        outs.append((k, ds, c))
    dur = torch.empty((B, cap), dtype=torch.float32, device="cuda"); ddep = torch.empty((B, cap), dtype=torch.float32, device="cuda")
    orbx._check(m._lib.orbx_compute_stereo_matches_batch_device(m._h, EL._h, ER._h, B, cap, 0.0, mc.BF_KITTI, dur.data_ptr(), ddep.data_ptr()))
    m.synchronize() if hasattr(m, "synchronize") else torch.cuda.synchronize()
    torch.cuda.synchronize()
    ur0, dep0 = m.ComputeStereoMatchesBatch(EL, ER, B, cap, 0.0, mc.BF_KITTI)
    assert np.array_equal(dur.cpu().numpy(), ur0) and np.array_equal(ddep.cpu().numpy(), dep0)
    assert np.array_equal(ur0[0][:cl[0]], G["stereo_ur"])
    # the asynchronous form followed at once by the extractors' NEXT batch (different images, same buffers): the extractors' streams must wait for the stereo
    # kernels, which are still reading their pyramids and results
    dL2 = torch.flip(dL, dims=[0]).contiguous(); dR2 = torch.flip(dR, dims=[0]).contiguous()
    for rep in range(4):
        for E, d, o in ((EL, dL, outs[0]), (ER, dR, outs[1])):
            E.extract_batch_raw(d.data_ptr(), B, 376, 1241, 1241, 1241 * 376, o[0].data_ptr(), o[1].data_ptr(), cap, o[2].data_ptr(), device=True)
        dur.fill_(7.0); ddep.fill_(7.0)
        orbx._check(m._lib.orbx_compute_stereo_matches_batch_device(m._h, EL._h, ER._h, B, cap, 0.0, mc.BF_KITTI, dur.data_ptr(), ddep.data_ptr()))
        for E, d, o in ((EL, dL2, outs[0]), (ER, dR2, outs[1])):          # queued behind the stereo call without any host synchronisation
            E.extract_batch_raw(d.data_ptr(), B, 376, 1241, 1241, 1241 * 376, o[0].data_ptr(), o[1].data_ptr(), cap, o[2].data_ptr(), device=True)
        torch.cuda.synchronize()
        assert np.array_equal(dur.cpu().numpy(), ur0) and np.array_equal(ddep.cpu().numpy(), dep0), rep
