"""CPU: the OpenCV-free primitive restatements (oracle/cvlite) against (a) golden vectors produced by the real
OpenCV 4.13 (tests/golden/cvlite_cv2.npz, made by tests/golden/make_golden.py) and (b) cv2 itself when importable.
Bit-exact: these are integer / fixed-point / unfused-float algorithms."""
import os
import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "cvlite_cv2.npz"))


def _fast(lib, oracle, img, th, nms):
    kp = np.zeros(200000, oracle.KP_DTYPE)
    n = lib.cvl_c_fast(np.ascontiguousarray(img), img.shape[1], img.shape[0], th, nms, kp, len(kp))
    assert n >= 0
    return np.stack([kp["x"][:n], kp["y"][:n], kp["response"][:n]], 1).astype(np.float32).reshape(-1, 3)


def test_resize_blur_border_golden(oracle):
    lib = oracle.port_lib()
    for name in ("small", "synth"):
        img = np.ascontiguousarray(G["img_" + name]); h, w = img.shape
        ref = G["resize_" + name]
        out = np.zeros_like(ref); lib.cvl_c_resize(img, w, h, out, ref.shape[1], ref.shape[0])
        assert np.array_equal(out, ref)
        b = np.zeros_like(img); lib.cvl_c_blur7(img, w, h, b)
        assert np.array_equal(b, G["blur_" + name])
    img = np.ascontiguousarray(G["img_small"]); h, w = img.shape
    out = np.zeros((h + 38, w + 38), np.uint8); lib.cvl_c_border101(img, w, h, 19, out)
    assert np.array_equal(out, G["border_small"])


def test_fast_golden(oracle):
    lib = oracle.port_lib()
    img = G["img_synth"]
    for th in (20, 7):
        assert np.array_equal(_fast(lib, oracle, img, th, 1), G["fast%d" % th])          # content AND raster order
        assert np.array_equal(_fast(lib, oracle, img, th, 0), G["fast%d_nonms" % th])


def test_fast_score_map_equivalence(oracle):
    """cv::FAST(t, NMS) == strict 3x3 local maxima of the S map with S > t, response S-1, raster order."""
    lib = oracle.port_lib()
    img = np.ascontiguousarray(G["img_synth"]); h, w = img.shape
    S = np.zeros_like(img); lib.cvl_c_fast_smap(img, w, h, S)
    Si = S.astype(np.int32)
    for th in (20, 7):
        P = np.pad(Si, 1)
        c = P[1:-1, 1:-1]
        ismax = np.ones_like(c, bool)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dy or dx:
                    ismax &= c > P[1 + dy:P.shape[0] - 1 + dy, 1 + dx:P.shape[1] - 1 + dx]
        ys, xs = np.nonzero(ismax & (c > th))
        mine = np.stack([xs, ys, c[ys, xs] - 1], 1).astype(np.float32)
        assert np.array_equal(mine, G["fast%d" % th])


def test_atan2_ellipse_close_golden(oracle):
    lib = oracle.port_lib()
    y, x = np.ascontiguousarray(G["atan_y"]), np.ascontiguousarray(G["atan_x"])
    o = np.zeros_like(x); lib.cvl_c_atan2(y, x, o, len(x))
    assert np.array_equal(o, G["atan_out"])                                              # exact float equality
    e = np.zeros((31, 31), np.uint8); lib.cvl_c_ellipse31(e)
    assert np.array_equal(e, G["ellipse31"])
    m = np.ascontiguousarray(G["mask"]); c = np.zeros_like(m); lib.cvl_c_close31(m, m.shape[1], m.shape[0], c)
    assert np.array_equal(c, G["mask_close"])


def test_primitives_against_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    lib = oracle.port_lib()
    rng = np.random.default_rng(7)
    for (h, w) in [(480, 640), (376, 1241), (61, 75)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        dw, dh = int(round(w / 1.2)), int(round(h / 1.2))
        out = np.zeros((dh, dw), np.uint8); lib.cvl_c_resize(img, w, h, out, dw, dh)
        assert np.array_equal(out, cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))
        b = np.zeros_like(img); lib.cvl_c_blur7(img, w, h, b)
        assert np.array_equal(b, cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
    from tools.synth import synth_frame
    img = synth_frame(4, 320, 240)
    for th in (20, 7):
        k = cv2.FastFeatureDetector_create(th, True).detect(img)
        ref = np.array([(p.pt[0], p.pt[1], p.response) for p in k], np.float32).reshape(-1, 3)
        assert np.array_equal(_fast(lib, oracle, img, th, 1), ref)


def test_det_sincos_is_glibc_sincosf(oracle):
    """The rBRIEF rotation (src/ORBextractor.cc:178-181) is glibc's sincosf in the reference build.  The oracle's restatement of
    that function (FMA variant; mirrored by the CUDA kernel) must equal the live libm bit for bit: EVERY float in [0, 7.0) --
    a superset of every angle the path can produce (degrees in [0, 360] times pi/180) -- and 2^27 strided patterns of the rest."""
    import threading
    lib = oracle.port_lib()
    hi = int(np.float32(7.0).view(np.uint32))
    nthr = 4; res = [None] * nthr
    def run(i):
        bad = np.zeros(8, np.uint32)
        lo_i, hi_i = hi * i // nthr, hi * (i + 1) // nthr
        res[i] = (lib.port_sincos_sweep(lo_i, hi_i, bad, 8), bad.copy())
    th = [threading.Thread(target=run, args=(i,)) for i in range(nthr)]
    [t.start() for t in th]; [t.join() for t in th]
    assert sum(r[0] for r in res) == 0, [hex(int(x)) for r in res for x in r[1][:int(min(r[0], 8))]]
    # the rest of the float range (negative, large-argument reduction, inf / nan): blocks of 2^12 patterns every 2^17
    bad = np.zeros(8, np.uint32)
    tot = 0
    for blk in range(0, 1 << 32, 1 << 17):
        if blk + 4096 <= hi: continue
        tot += lib.port_sincos_sweep(blk, min(blk + 4096, (1 << 32) - 1), bad, 8)
    assert tot == 0, [hex(int(x)) for x in bad]
    # and it is NOT the correctly rounded value everywhere (why a structural port is needed): count the angles where they differ
    ang = (np.arange(0, 360000, dtype=np.float32) / np.float32(1000.0)) * np.float32(np.pi / 180.0)
    s1, c1 = np.zeros_like(ang), np.zeros_like(ang)
    lib.port_det_sincos(ang, s1, c1, len(ang))
    exact_s, exact_c = np.sin(ang.astype(np.float64)).astype(np.float32), np.cos(ang.astype(np.float64)).astype(np.float32)
    d = (s1 != exact_s) | (c1 != exact_c)
    assert 0.001 < d.mean() < 0.06
    for a, b in ((s1, exact_s), (c1, exact_c)):
        ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
        assert ulp.max() <= 1


def test_small_matrix_products_are_opencv_gemm(oracle):
    """cv::Mat products of the matcher bodies (Rcw * x3Dw + tcw, src/ORBmatcher.cc:1608): cvlite's operator* / operator+ against
    cv2.gemm of OpenCV 4.13 -- float arithmetic on the hand-unrolled 2..4 path, double accumulation on the general path."""
    lib = oracle.port_lib()
    names = sorted({k[5:-2] for k in G.files if k.startswith("gemm_") and k.endswith("_A")})
    assert len(names) >= 7
    for name in names:
        A, B, D = G["gemm_%s_A" % name], G["gemm_%s_B" % name], G["gemm_%s_D" % name]
        Cm = G["gemm_%s_C" % name] if ("gemm_%s_C" % name) in G.files else None
        for i in range(len(A)):
            out = np.zeros(D[i].shape, np.float32)
            c = np.ascontiguousarray(Cm[i]) if Cm is not None else None
            lib.cvl_c_gemm(np.ascontiguousarray(A[i]), A.shape[1], A.shape[2], np.ascontiguousarray(B[i]), B.shape[2], c.ctypes.data if c is not None else None, out)
            assert np.array_equal(out, D[i]), (name, i)


def test_numpy_projection_helper_is_opencv_gemm():
    """tests/match_cases.project_pose (the expected values of the device projection) evaluates Rcw * x + tcw like cv2.gemm."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import match_cases as mc
    A, B, Cm, D = G["gemm_3x3_3x1_c_A"], G["gemm_3x3_3x1_c_B"], G["gemm_3x3_3x1_c_C"], G["gemm_3x3_3x1_c_D"]
    f = np.float32
    for i in range(len(A)):
        R, X, t = A[i], B[i][:, 0][None, :], Cm[i][:, 0]
        c = [((((R[r, 0] * X[:, 0]).astype(f) + (R[r, 1] * X[:, 1]).astype(f)).astype(f) + (R[r, 2] * X[:, 2]).astype(f)).astype(f) + t[r]).astype(f) for r in range(3)]
        assert np.array_equal(np.array(c, f).reshape(3, 1), D[i]), i
    # and project_pose uses exactly that expression: a point on the optical axis of an identity pose lands on the principal point
    uv, iz, va = mc.project_pose(np.array([[0, 0, 2.0]], f), np.eye(3, dtype=f), np.zeros(3, f), (0, 640, 0, 480))
    assert va[0] == 1 and uv[0, 0] == f(mc.CX) and uv[0, 1] == f(mc.CY) and iz[0] == f(0.5)
