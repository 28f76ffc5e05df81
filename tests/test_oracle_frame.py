"""CPU: the oracle's frame steps (UndistortKeyPoints / ComputeImageBounds / ComputeStereoFromRGBD / AssignFeaturesToGrid,
/root/reference/src/Frame.cc:1052-1176, 1576-1614, 431-461) against the committed goldens: cv2 4.13's undistortPoints and
the outputs of the reference's own function bodies (tests/golden/make_golden_frame.py)."""
import os
import numpy as np
import pytest
import frame_cases as fc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "frame_cv2.npz"))
KEYS = ("keys_un", "u_right", "depth", "bounds", "cell_start", "entries")


def same(r, name, tag):
    for key in KEYS:
        g = G["%s_%s_%s" % (name, tag, key)]
        v = r[key].view(np.uint8) if key == "keys_un" else r[key]
        assert np.array_equal(v, g), (name, tag, key)


@pytest.mark.parametrize("name", list(fc.CAMS))
def test_undistort_points_matches_cv2(oracle, name):
    pts, _ = fc.points()
    assert np.array_equal(oracle.undistort_points("port", pts, fc.CAMS[name]), G["und_" + name])
    if oracle.have_ref():
        assert np.array_equal(oracle.undistort_points("ref", pts, fc.CAMS[name]), G["und_" + name])


def test_undistort_actually_moves_points():
    pts, _ = fc.points()
    d = np.abs(G["und_tum1"] - pts).max(1)
    assert d.max() > 5.0 and np.array_equal(G["und_tum3"], pts)          # k1 == 0: cv2 returns the input


@pytest.mark.parametrize("name", list(fc.CAMS))
def test_port_frame_build_matches_reference_bodies(oracle, name):
    k = fc.keys(oracle.KP_DTYPE)
    same(oracle.frame_build("port", k, fc.CAMS[name], fc.BF, fc.ROWS, fc.COLS, fc.depth_image()), name, "rgbd")
    same(oracle.frame_build("port", k, fc.CAMS[name], fc.BF, fc.ROWS, fc.COLS, None), name, "mono")


def test_reference_bodies_reproduce_their_golden(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    k = fc.keys(oracle.KP_DTYPE)
    same(oracle.frame_build("ref", k, fc.CAMS["tum1"], fc.BF, fc.ROWS, fc.COLS, fc.depth_image()), "tum1", "rgbd")


def test_grid_and_stereo_properties():
    """Size-independent checks on the golden itself: every in-grid keypoint appears exactly once, cells hold ascending indices,
    uRight = x_un - bf / d exactly where d > 0 and -1 elsewhere."""
    for name in fc.CAMS:
        cs, en = G[name + "_rgbd_cell_start"], G[name + "_rgbd_entries"]
        assert cs[0] == 0 and np.all(np.diff(cs) >= 0) and cs[-1] == len(en) and len(set(en.tolist())) == len(en)
        for c in np.nonzero(np.diff(cs) > 1)[0][:200]:
            assert np.all(np.diff(en[cs[c]:cs[c + 1]]) > 0)
        ku = G[name + "_rgbd_keys_un"].view(np.dtype([("x", "<f4"), ("y", "<f4"), ("r", "V20")]))
        d, ur = G[name + "_rgbd_depth"], G[name + "_rgbd_u_right"]
        pos = d > 0
        assert pos.any() and (~pos).any() and np.all(d[~pos] == -1) and np.all(ur[~pos] == -1)
        assert np.array_equal(ur[pos], ku["x"][pos] - np.float32(fc.BF) / d[pos])
        assert np.all(G[name + "_mono_u_right"] == -1) and np.all(G[name + "_mono_depth"] == -1)
