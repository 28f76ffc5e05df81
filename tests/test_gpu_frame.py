"""GPU (B200): device-resident Frame through the C ABI.  mvKeysUn, mvuRight, mvDepth, image bounds and
mGrid bit-exact against the reference bodies' goldens / the port oracle; the three windowed matchers on device frames must
return exactly what they return on host frame views (which test_gpu_matcher.py pins to the reference)."""
import os
import numpy as np
import pytest
import frame_cases as fc
import match_cases as mc

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "frame_cv2.npz"))
SF = (np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32)


def check(F, name, tag, n):
    ku, ur, dp, b = F.read()
    cs, en = F.grid()
    assert F.N == n
    assert np.array_equal(ku.view(np.uint8), G["%s_%s_keys_un" % (name, tag)])
    assert np.array_equal(ur, G["%s_%s_u_right" % (name, tag)]) and np.array_equal(dp, G["%s_%s_depth" % (name, tag)])
    assert np.array_equal(b, G["%s_%s_bounds" % (name, tag)])
    assert np.array_equal(cs, G["%s_%s_cell_start" % (name, tag)]) and np.array_equal(en, G["%s_%s_entries" % (name, tag)])


@pytest.mark.parametrize("name", list(fc.CAMS))
def test_frame_from_host_arrays_matches_reference(orbx, name):
    k = fc.keys(orbx.KP_DTYPE); desc = np.zeros((len(k), 32), np.uint8); dimg = fc.depth_image()
    F = orbx.Frame()
    check(F.assign_host(k, desc, SF, fc.cam_struct(orbx, name), fc.ROWS, fc.COLS, dimg), name, "rgbd", len(k))
    gathered = dimg[k["y"].astype(np.int32), k["x"].astype(np.int32)]                 # imDepth.at<float>(v, u) done by the caller
    check(F.assign_host(k, desc, SF, fc.cam_struct(orbx, name), fc.ROWS, fc.COLS, gathered), name, "rgbd", len(k))
    padded = np.zeros((fc.ROWS, fc.COLS + 24), np.float32); padded[:, :fc.COLS] = dimg    # row stride != cols * 4
    check(F.assign_host(k, desc, SF, fc.cam_struct(orbx, name), fc.ROWS, fc.COLS, padded[:, :fc.COLS]), name, "rgbd", len(k))
    check(F.assign_host(k, desc, SF, fc.cam_struct(orbx, name), fc.ROWS, fc.COLS, None), name, "mono", len(k))
    ur, dep = G[name + "_rgbd_u_right"], G[name + "_rgbd_depth"]
    F.set_stereo(ur, dep)                                                              # as after ComputeStereoMatches
    _, ur2, dp2, _ = F.read()
    assert np.array_equal(ur2, ur) and np.array_equal(dp2, dep)


@pytest.mark.parametrize("name", ["tum1", "tum3"])
def test_stepwise_calls_match_reference(orbx, name):
    """UndistortKeyPoints / ComputeStereoFromRGBD / AssignFeaturesToGrid as separate calls, in the constructors' orders."""
    k = fc.keys(orbx.KP_DTYPE); n = len(k); desc = np.zeros((n, 32), np.uint8); dimg = fc.depth_image(); cam = fc.cam_struct(orbx, name)
    g = lambda key, tag="rgbd": G["%s_%s_%s" % (name, tag, key)]
    F = orbx.Frame().take_host(k, desc, SF)
    with pytest.raises(orbx.OrbxError):
        F.AssignFeaturesToGrid(g("bounds"), n)                                          # grid before undistortion: call-sequence error
    with pytest.raises(orbx.OrbxError):
        orbx.ORBmatcher(0.9, True).SearchForInitialization(F, F, np.zeros((n, 2), np.float32), 10)   # not gridded yet
    assert np.array_equal(F.UndistortKeyPoints(cam, n).view(np.uint8), g("keys_un"))
    ur, de = F.ComputeStereoFromRGBD(fc.BF, dimg, fc.ROWS, fc.COLS, n)                  # RGB-D order (Frame.cc:640-645)
    assert np.array_equal(ur, g("u_right")) and np.array_equal(de, g("depth"))
    cs, en = F.AssignFeaturesToGrid(g("bounds"), n)
    assert np.array_equal(cs, g("cell_start")) and np.array_equal(en, g("entries")) and F.N == n
    ku, ur2, de2, b = F.read()
    assert np.array_equal(ku.view(np.uint8), g("keys_un")) and np.array_equal(ur2, ur) and np.array_equal(b, g("bounds"))
    # stereo order (Frame.cc:187-240): undistort, stereo results from ComputeStereoMatches, then the grid
    F.take_host(k, desc, SF)
    F.UndistortKeyPoints(cam, n)
    F.set_stereo(g("u_right"), g("depth"))
    cs, en = F.AssignFeaturesToGrid(g("bounds"), n)
    _, ur3, de3, _ = F.read()
    assert np.array_equal(cs, g("cell_start")) and np.array_equal(ur3, g("u_right")) and np.array_equal(de3, g("depth"))
    # gathered depth, mono order
    F.take_host(k, desc, SF); F.UndistortKeyPoints(cam, n)
    ur, de = F.ComputeStereoFromRGBD(fc.BF, dimg[k["y"].astype(np.int32), k["x"].astype(np.int32)], 0, 0, n)
    assert np.array_equal(ur, g("u_right")) and np.array_equal(de, g("depth"))


def test_frame_from_extractor_result_and_amos_sequence(orbx, oracle):
    from tools.synth import synth_frame
    img = synth_frame(4, 640, 480); dimg = fc.depth_image()
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    F = orbx.Frame()
    with pytest.raises(orbx.OrbxError):
        F.assign(E, fc.cam_struct(orbx, "tum1"), 480, 640, dimg)                       # no result on the device yet
    k, d = E(img)
    F.assign(E, fc.cam_struct(orbx, "tum1"), 480, 640, dimg)
    o = oracle.frame_build("port", k, fc.CAMS["tum1"], fc.BF, 480, 640, dimg)
    ku, ur, dp, b = F.read(); cs, en = F.grid()
    assert np.array_equal(ku, o["keys_un"]) and np.array_equal(ur, o["u_right"]) and np.array_equal(dp, o["depth"])
    assert np.array_equal(b, o["bounds"]) and np.array_equal(cs, o["cell_start"]) and np.array_equal(en, o["entries"])
    assert (np.abs(ku["x"] - k["x"]) > 0.5).any() and (ur > 0).sum() > 100
    # Amos order (Frame::CalDyna, Frame.cc:636-645): detect -> MovingKeyPoints -> ProcessDesp -> N -> Undistort -> RGBD -> grid
    kd, counts = E.detect(img)
    yy, xx = np.mgrid[0:480, 0:640]
    mask = (((xx > 200) & (xx < 330) & (yy > 100) & (yy < 250)) * 255).astype(np.uint8)
    label = (1 + (xx // 80) + 8 * (yy // 80)).astype(np.float64)
    rm = np.zeros(64, np.int32); rm[[5, 17]] = 1
    kept, counts2, culled = E.MovingKeyPoints(mask, label, np.arange(64, dtype=np.int32), rm, kd, counts)
    kc, dc = E.ProcessDesp(kept, counts2)
    assert 0 < len(kc) < len(k)
    F.assign(E, fc.cam_struct(orbx, "tum2"), 480, 640, dimg)
    o = oracle.frame_build("port", kc, fc.CAMS["tum2"], fc.BF, 480, 640, dimg)
    ku, ur, dp, b = F.read(); cs, en = F.grid()
    assert F.N == len(kc) and np.array_equal(ku, o["keys_un"]) and np.array_equal(ur, o["u_right"]) and np.array_equal(en, o["entries"]) and np.array_equal(cs, o["cell_start"])
    # a detect without describe leaves no complete result behind
    E.detect(img)
    with pytest.raises(orbx.OrbxError):
        F.assign(E, fc.cam_struct(orbx, "tum1"), 480, 640, None)


@pytest.fixture(scope="module")
def pair(orbx):
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    cam = fc.cam_struct(orbx, "tum3")                                                   # rectified: device frame == host view inputs
    from tools.synth import synth_frame, warp_affine_nn
    A = synth_frame(0, 640, 480); B = warp_affine_nn(A, 7, -4, 2.0)
    ka, da = E(A); FA = orbx.Frame().assign(E, cam, 480, 640)
    kb, db = E(B); FB = orbx.Frame().assign(E, cam, 480, 640)
    pi = mc.projection_inputs(ka, kb)
    FB.set_stereo(pi["u_right"], np.where(pi["u_right"] > 0, 1.0, -1.0).astype(np.float32))
    sf = E.GetScaleFactors()
    return dict(ka=ka, da=da, kb=kb, db=db, pi=pi, FA=FA, FB=FB, VA=orbx.FrameView(ka, da, 640, 480, sf), VB=orbx.FrameView(kb, db, 640, 480, sf, u_right=pi["u_right"]))


def test_matchers_on_device_frames_equal_host_views(orbx, pair):
    p = pair; pi = p["pi"]; ka = p["ka"]
    prev = np.stack([ka["x"], ka["y"]], 1)
    for nn, ori, win in ((0.9, True, 100), (0.9, False, 30), (0.6, True, 10)):
        m = orbx.ORBmatcher(nn, ori)
        a = m.SearchForInitialization(p["FA"], p["FB"], prev, win); b = m.SearchForInitialization(p["VA"], p["VB"], prev, win)
        assert a[0] == b[0] and a[0] > 0 and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    uv, iz = mc.project(pi["xyz"])
    for th, mono in mc.PROJ_FRAME_CASES:
        m = orbx.ORBmatcher(0.9, True)
        a = m.SearchByProjectionFrame(p["FB"], uv, iz, ka["octave"], ka["angle"], p["da"], pi["valid"], pi["obs"], pi["occ"], th, False, False, 40.0)
        b = m.SearchByProjectionFrame(p["VB"], uv, iz, ka["octave"], ka["angle"], p["da"], pi["valid"], pi["obs"], pi["occ"], th, False, False, 40.0)
        assert a[0] == b[0] and a[0] > 20 and np.array_equal(a[1], b[1])
    for th in mc.PROJ_POINT_CASES:
        m = orbx.ORBmatcher(0.8, True)
        a = m.SearchByProjectionPoints(p["FB"], pi["tuv"], pi["tur"], pi["lvl"], pi["vc"], p["da"], pi["obs"], pi["occ"], th)
        b = m.SearchByProjectionPoints(p["VB"], pi["tuv"], pi["tur"], pi["lvl"], pi["vc"], p["da"], pi["obs"], pi["occ"], th)
        assert a[0] == b[0] and a[0] > 20 and np.array_equal(a[1], b[1])


def test_get_features_in_area_matches_oracle(orbx, oracle, pair):
    p = pair
    oF = oracle.FrameData(p["kb"], p["db"], 640, 480, (np.float32(1.2) ** np.arange(8, dtype=np.float32)))
    rng = np.random.default_rng(3)
    nq = 300
    xy = np.stack([rng.uniform(-30, 670, nq), rng.uniform(-30, 510, nq)], 1).astype(np.float32)
    r = rng.choice([3.0, 10.0, 15.0, 40.0, 100.0], nq).astype(np.float32)
    mn = rng.integers(-1, 5, nq).astype(np.int32); mx = rng.integers(-1, 8, nq).astype(np.int32)
    m = orbx.ORBmatcher(0.9, True)
    got = m.GetFeaturesInArea(p["FB"], xy, r, mn, mx)
    om = oracle.Matcher("port")
    total = 0
    for q in range(nq):
        want = om.features_in_area(oF, float(xy[q, 0]), float(xy[q, 1]), float(r[q]), int(mn[q]), int(mx[q]))
        assert np.array_equal(got[q], want), q
        total += len(want)
    assert total > 2000
    assert m.GetFeaturesInArea(p["FB"], np.zeros((0, 2), np.float32), 1.0) == []


def test_empty_and_error_paths(orbx):
    F = orbx.Frame()
    with pytest.raises(orbx.OrbxError):
        F.read()                                                                        # nothing assigned yet
    cam = fc.cam_struct(orbx, "tum1")
    F.assign_host(np.zeros(0, orbx.KP_DTYPE), np.zeros((0, 32), np.uint8), SF, cam, 480, 640, None)
    assert F.N == 0
    cs, en = F.grid()
    assert cs[-1] == 0 and len(en) == 0
    m = orbx.ORBmatcher(0.9, True)
    nm, m12, _ = m.SearchForInitialization(F, F, np.zeros((0, 2), np.float32), 100)
    assert nm == 0 and len(m12) == 0
    with pytest.raises(orbx.OrbxError):
        F.assign_host(fc.keys(orbx.KP_DTYPE), np.zeros((3000, 32), np.uint8), SF, orbx.Camera.make(0.0, 500.0, 320.0, 240.0), 480, 640, None)   # fx == 0
