"""GPU (B200) parity tests of the ORBextractor path, all through the C ABI (amos-slam_b200 -> liborbx_b200.so).
Bit-exact against the port oracle on the same seeded inputs, against the committed golden outputs of the
reference's own ORBextractor.cc, and -- where oracle/_ref travelled -- against the live reference build."""
import os
import threading
import numpy as np
import pytest
from tools.synth import synth_frame, synth_batch, synth_mask

pytestmark = pytest.mark.gpu
R = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_extract.npz"))


def kp_equal(a, b):
    return len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in a.dtype.names)


def assert_same(kg, dg, ko, do):
    assert len(kg) == len(ko), (len(kg), len(ko))
    for f in ko.dtype.names:                      # coordinates, octaves, responses, sizes exact; angles exact (<= 1e-3 rad is the contract)
        assert np.array_equal(kg[f], ko[f]), f
    assert np.array_equal(dg, do)


@pytest.mark.parametrize("name", ["c1", "odd", "wide", "lv4"])
def test_extract_matches_reference_golden(orbx, name):
    w, h, nf, nl, it, mt, seed = [int(v) for v in R[name + "_params"]]
    E = orbx.ORBextractor(nf, float(R[name + "_scale"]), nl, it, mt)
    kp, desc = E(synth_frame(seed, w, h))
    assert_same(kp, desc, R[name + "_kp"], R[name + "_desc"])
    assert E.check_overflow() == 0


@pytest.mark.parametrize("cfg", [(640, 480, 1000, 41), (752, 480, 2000, 42), (1241, 376, 2000, 43), (1920, 1080, 1000, 44), (160, 120, 100, 45), (100, 100, 50, 46)])
def test_extract_matches_oracle_configs(orbx, oracle, cfg):
    """C1 / C3 / C4 / C5 geometries (BASELINE.json configs) + a tiny frame."""
    w, h, nf, seed = cfg
    E = orbx.ORBextractor(nf, 1.2, 8, 20, 7); P = oracle.Extractor("port", nf, 1.2, 8, 20, 7)
    img = synth_frame(seed, w, h)
    kg, dg = E(img); ko, do = P.extract(img)
    assert_same(kg, dg, ko, do)
    for l in range(8):
        po = P.pyramid_level(l)
        assert np.array_equal(E.debug_pyramid_level(0, l, po.shape), po)
        co, cg = P.level_candidates(l), E.debug_level_candidates(0, l)
        assert len(co) == len(cg) and all(np.array_equal(co[f], cg[f]) for f in ("x", "y", "response"))
    if oracle.have_ref() and w <= 1241:
        kr, dr = oracle.Extractor("ref", nf, 1.2, 8, 20, 7).extract(img)
        assert_same(kg, dg, kr, dr)


@pytest.mark.parametrize("cfg", [(800, 600, 500, 2.5, 3, 47), (500, 375, 600, 1.07, 12, 48), (640, 480, 500, 2.0, 4, 49)])
def test_unusual_scale_factors(orbx, oracle, cfg):
    """scale > 2 takes the byte-gather resize kernel, scale <= 2 the word-load one; both must equal the oracle."""
    w, h, nf, sf, nl, seed = cfg
    E = orbx.ORBextractor(nf, sf, nl, 20, 7); P = oracle.Extractor("port", nf, sf, nl, 20, 7)
    img = synth_frame(seed, w, h)
    kg, dg = E(img); ko, do = P.extract(img)
    assert_same(kg, dg, ko, do)
    for l in range(nl):
        po = P.pyramid_level(l)
        assert np.array_equal(E.debug_pyramid_level(0, l, po.shape), po)


def test_large_frame_many_features_fallback_paths(orbx, oracle):
    """A big, corner-dense frame with a large quota: more candidates per level than the radix sort / code staging hold in shared
    memory (global-memory bitonic sort, global key searches) and more nodes per level than the round-based quadtree kernel
    handles (serial replay kernel).  Same bit-exact contract."""
    rng = np.random.default_rng(21)
    img = synth_frame(61, 1600, 1200)
    img[200:1000, 300:1300] = rng.integers(0, 256, (800, 1000), dtype=np.uint8)          # dense corners: > 16 k candidates on level 0
    for nf in (6000, 1500):
        E = orbx.ORBextractor(nf, 1.2, 8, 20, 7); P = oracle.Extractor("port", nf, 1.2, 8, 20, 7)
        kg, dg = E(img); ko, do = P.extract(img)
        assert_same(kg, dg, ko, do)
        assert E.check_overflow() == 0
        n0 = len(E.debug_level_candidates(0, 0))
        assert n0 > 16384, n0


def test_textured_and_flat_inputs(orbx, oracle):
    rng = np.random.default_rng(3)
    E = orbx.ORBextractor(500, 1.2, 8, 20, 7); P = oracle.Extractor("port", 500, 1.2, 8, 20, 7)
    noise = rng.integers(0, 256, (240, 320), dtype=np.uint8)                 # corners everywhere: stresses cell capacity + quadtree
    flat = np.full((240, 320), 77, np.uint8)                                  # no corner at all
    grad = np.tile(np.arange(320, dtype=np.uint8), (240, 1))
    checker = ((np.indices((240, 320)).sum(0) // 8) % 2 * 200 + 20).astype(np.uint8)   # many equal responses (tie-breaks)
    for img in (noise, flat, grad, checker):
        kg, dg = E(img); ko, do = P.extract(img)
        assert_same(kg, dg, ko, do)
    assert len(E(flat)[0]) == 0


def test_empty_and_bad_inputs(orbx):
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    kp, desc = E(np.zeros((0, 0), np.uint8))                                   # reference: silent return (ORBextractor.cc:1553)
    assert len(kp) == 0 and desc.shape == (0, 32)
    with pytest.raises(orbx.OrbxError):
        E(np.zeros((480, 640, 3), np.uint8))                                   # assert(image.type()==CV_8UC1)
    kp, desc = E(np.full((60, 60), 9, np.uint8))                               # levels without a 30-px cell yield nothing (as the reference)
    assert len(kp) == 0
    with pytest.raises(orbx.OrbxError) as e:
        E(np.zeros((32, 64), np.uint8))                                        # level 0 is 32 px high: reference's nIni = w/0 is undefined
    assert e.value.code == orbx.E_INVALID
    with pytest.raises(orbx.OrbxError):
        orbx.ORBextractor(0, 1.2, 8, 20, 7)


def test_strided_input_and_geometry_change(orbx, oracle):
    E = orbx.ORBextractor(600, 1.2, 8, 20, 7); P = oracle.Extractor("port", 600, 1.2, 8, 20, 7)
    big = synth_frame(50, 700, 500)
    view = big[10:490, 30:670]                                                 # non-contiguous rows (step 700)
    kg, dg = E(view); ko, do = P.extract(np.ascontiguousarray(view))
    assert_same(kg, dg, ko, do)
    small = synth_frame(51, 400, 300)                                          # same handle, new geometry
    kg, dg = E(small); ko, do = P.extract(small)
    assert_same(kg, dg, ko, do)
    kg, dg = E(view); ko, do = P.extract(np.ascontiguousarray(view))          # and back
    assert_same(kg, dg, ko, do)


def test_batch_equals_single(orbx, oracle):
    imgs = synth_batch(12, 640, 480, seed0=60, distinct=12)
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7); P = oracle.Extractor("port", 1000, 1.2, 8, 20, 7)
    kp, desc, counts = E.extract_batch(imgs)
    for b in (0, 5, 11):
        ko, do = P.extract(imgs[b])
        assert_same(kp[b][:counts[b]], desc[b][:counts[b]], ko, do)
    for b in range(12):                                                        # batch == per-frame call (idempotence)
        k1, d1 = E(imgs[b])
        assert_same(kp[b][:counts[b]], desc[b][:counts[b]], k1, d1)
    assert E.check_overflow() == 0


def test_host_batch_pipeline_many_chunks(orbx):
    """The host-pointer batch call cuts large batches into chunks over several streams (H2D / compute / D2H overlapped):
    a 150-frame batch (ramped chunks 16, 32, 64, 38) must equal the per-frame call frame by frame, also from a strided,
    non-dense host layout (per-frame copy path) and when called twice in a row (buffer reuse across calls)."""
    B = 150
    imgs = synth_batch(B, 640, 480, seed0=400, distinct=10)
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7); E1 = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    probe = sorted(set(list(range(0, B, 7)) + [15, 16, 17, 47, 48, 111, 112, B - 1]))          # incl. both sides of every chunk boundary
    singles = {b: E1(imgs[b]) for b in probe}
    first = None
    for rep in range(2):
        kp, desc, counts = E.extract_batch(imgs)
        for b in probe:
            k1, d1 = singles[b]
            assert counts[b] == len(k1), (rep, b)
            assert kp_equal(kp[b][:counts[b]], k1) and np.array_equal(desc[b][:counts[b]], d1), (rep, b)
        if first is None:
            first = (kp.copy(), desc.copy(), counts.copy())
        else:
            assert np.array_equal(counts, first[2])
            for b in range(B):
                assert np.array_equal(kp[b][:counts[b]], first[0][b][:counts[b]]) and np.array_equal(desc[b][:counts[b]], first[1][b][:counts[b]])
    # odd geometry + row padding: the mirror path does not apply (step % 4 != 0), frames are copied one by one
    big = np.zeros((70, 300, 403), np.uint8)
    src = synth_batch(70, 401, 300, seed0=410, distinct=5)
    big[:, :, :401] = src
    view = big[:, :, :401]                                                       # step 403, width 401
    cap = E.max_keypoints(300, 401)
    kp = np.zeros((70, cap), orbx.KP_DTYPE); desc = np.zeros((70, cap, 32), np.uint8); counts = np.zeros(70, np.int32)
    E.extract_batch_raw(view.ctypes.data, 70, 300, 401, view.strides[1], view.strides[0], kp.ctypes.data, desc.ctypes.data, cap, counts.ctypes.data, device=False)
    for b in range(0, 70, 3):
        k1, d1 = E1(np.ascontiguousarray(src[b]))
        assert counts[b] == len(k1) and kp_equal(kp[b][:counts[b]], k1) and np.array_equal(desc[b][:counts[b]], d1), b
    assert E.check_overflow() == 0


def test_masked_batch_many_chunks(orbx):
    """Same for the masked (C5) host call: 1080p chunks are 10 frames, so 24 frames run as several chunks on several streams."""
    B, w, h = 24, 1920, 1080
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7); E2 = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    imgs = synth_batch(B, w, h, seed0=420, distinct=3)
    masks = np.stack([synth_mask(420 + (b % 4), w, h) for b in range(B)])
    kp, desc, counts, culled = E.extract_masked_batch(imgs, masks)
    lab = np.ones((h, w), np.float64); ids = np.zeros(1, np.int32); rm = np.zeros(1, np.int32)
    for b in (0, 2, 9, 10, 11, 19, 20, B - 1):                                      # both sides of the 1080p chunk boundaries
        kd, cd = E2.detect(imgs[b]); kk, ck, cu = E2.MovingKeyPoints(masks[b], lab, ids, rm, kd, cd); kf, df = E2.ProcessDesp(kk, ck)
        assert counts[b] == len(kf) and culled[b] == len(cu), b
        assert kp_equal(kp[b][:counts[b]], kf) and np.array_equal(desc[b][:counts[b]], df), b


def test_batch_full_size_properties(orbx):
    """C5 shape (1920x1080): batch result is independent of batch position / batch size, and repeatable."""
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    imgs = synth_batch(6, 1920, 1080, seed0=70, distinct=2)
    kp, desc, counts = E.extract_batch(imgs)
    kp2, desc2, counts2 = E.extract_batch(imgs[::-1].copy())
    assert np.array_equal(counts, counts2[::-1])
    for b in range(6):
        assert np.array_equal(kp[b][:counts[b]], kp2[5 - b][:counts[b]]) and np.array_equal(desc[b][:counts[b]], desc2[5 - b][:counts[b]])
    k1, d1 = E(imgs[3])
    assert np.array_equal(k1, kp[3][:counts[3]]) and np.array_equal(d1, desc[3][:counts[3]])
    assert (counts >= 1000).all() and (counts <= 1000 + 2 * 8).all()
    assert E.check_overflow() == 0


def test_quadtree_stage_adversarial(orbx, oracle):
    """DistributeOctTree alone (device kernels) vs the reference's golden outputs and the port on adversarial sets."""
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7); P = oracle.Extractor("port", 1000, 1.2, 8, 20, 7)
    # forms of the stage: the one-launch kernel of the latency form (default for a single frame), the same kernel with a key capacity so small
    # that every set runs on its global-scratch path, the sort + tree pair of the batched path; path codes computed or read from the host's tables
    forms = ({}, {"ORBX_QT_KEYCAP": "64"}, {"ORBX_QT_FUSED": "0"}, {"ORBX_QT_TABLES": "1"}, {"ORBX_QT_FUSED": "0", "ORBX_QT_TABLES": "1"})
    def distribute(c, x0, x1, y0, y1, N):
        outs = []
        for f in forms:
            os.environ.update(f)
            try:
                outs.append(E.debug_distribute(c, x0, x1, y0, y1, N))
            finally:
                for k in f: os.environ.pop(k, None)
        for o in outs[1:]:
            assert len(o) == len(outs[0]) and all(np.array_equal(o[f], outs[0][f]) for f in ("x", "y", "response"))
        return outs[0]
    for k in [k for k in R.files if k.startswith("oct_") and k.endswith("_in")]:
        N = int(k.split("_")[-2])
        out = distribute(R[k], 16, 16 + 608, 16, 16 + 448, N)
        ref = R[k[:-3] + "_out"]
        assert len(out) == len(ref) and all(np.array_equal(out[f], ref[f]) for f in ("x", "y", "response")), k
    rng = np.random.default_rng(9)
    for trial in range(36):
        W, H = int(rng.integers(80, 1900)), int(rng.integers(60, 1060))
        if not (0.5 <= W / H < 15.5):
            continue
        n = int(rng.integers(1, 6000 if trial % 4 else 30000))
        xs = rng.integers(3, W - 3, n); ys = rng.integers(3, H - 3, n)
        if trial % 3 == 0:                                                      # heavy clustering
            xs = np.clip(rng.normal(W * 0.3, 6, n), 3, W - 4).astype(int); ys = np.clip(rng.normal(H * 0.6, 5, n), 3, H - 4).astype(int)
        if trial % 7 == 5:                                                      # a few tight pairs far apart: long chains of one-child divides
            m = int(rng.integers(2, 40)); cx = rng.integers(8, W - 8, m); cy = rng.integers(8, H - 8, m)
            xs = np.concatenate([cx, cx + rng.integers(1, 3, m)]); ys = np.concatenate([cy, cy + rng.integers(0, 2, m)])
        u = np.unique(np.stack([ys, xs], 1), axis=0); u = u[rng.permutation(len(u))]
        c = np.zeros(len(u), oracle.KP_DTYPE); c["x"] = u[:, 1]; c["y"] = u[:, 0]
        c["response"] = rng.integers(7, 40 if trial % 2 else 255, len(u)); c["size"] = 7; c["angle"] = -1; c["class_id"] = -1
        N = int(rng.integers(1, 500)) if trial % 5 else int(rng.integers(500, 1500))
        out = distribute(c, 16, 16 + W, 16, 16 + H, N); ref = P.distribute(c, 16, 16 + W, 16, 16 + H, N)
        assert len(out) == len(ref) and all(np.array_equal(out[f], ref[f]) for f in ("x", "y", "response")), (trial, W, H, n, N)


def test_amos_detect_cull_describe(orbx, oracle):
    """operator()(img, mask, vector<vector<KeyPoint>>) -> MovingKeyPoints -> ProcessDesp vs the reference's golden run."""
    E = orbx.ORBextractor(800, 1.2, 8, 20, 7)
    kp, counts = E.detect(synth_frame(int(R["amos_frame_seed"]), 480, 360))
    assert kp_equal(kp, R["amos_detect_kp"]) and np.array_equal(counts, R["amos_detect_counts"])
    kp2, counts2, culled = E.MovingKeyPoints(R["amos_mask"], R["amos_label"], R["amos_centers_id"], R["amos_rm"], kp, counts)
    assert kp_equal(kp2, R["amos_kept_kp"]) and np.array_equal(counts2, R["amos_kept_counts"]) and kp_equal(culled, R["amos_culled"])
    kp3, desc3 = E.ProcessDesp(kp2, counts2)
    assert kp_equal(kp3, R["amos_final_kp"]) and np.array_equal(desc3, R["amos_final_desc"])
    # all-zero mask / no flagged super-pixel: nothing is culled, describe == 4-arg operator()
    z = np.zeros((360, 480), np.uint8); lab = np.ones((360, 480), np.float64)
    kp4, counts4, culled4 = E.MovingKeyPoints(z, lab, np.zeros(1, np.int32), np.zeros(1, np.int32), kp, counts)
    assert len(culled4) == 0 and kp_equal(kp4, kp)
    k5, d5 = E.ProcessDesp(kp4, counts4)
    k6, d6 = E(synth_frame(int(R["amos_frame_seed"]), 480, 360))
    assert kp_equal(k5, k6) and np.array_equal(d5, d6)


def test_grayscale_and_ragged_masks_vs_oracle(orbx, oracle):
    """The closing is evaluated as binary morphology on (mask != 0) -- exact because the reference only tests closing != 0.
    Grey-scale masks, speckle, masks touching every border and an odd width exercise that equivalence against the port's
    grey-scale erode(dilate()) (cv semantics)."""
    rng = np.random.default_rng(11)
    w, h = 403, 301
    E = orbx.ORBextractor(700, 1.2, 8, 20, 7); P = oracle.Extractor("port", 700, 1.2, 8, 20, 7)
    img = synth_frame(95, w, h)
    kp, counts = E.detect(img); ko, co = P.detect(img)
    assert kp_equal(kp, ko) and np.array_equal(counts, co)
    lab = np.ones((h, w), np.float64); ids = np.zeros(1, np.int32); rm = np.zeros(1, np.int32)
    masks = [synth_mask(3, w, h),
             (synth_mask(4, w, h) // 255 * rng.integers(1, 256, (h, w))).astype(np.uint8),          # grey levels 1..255
             (rng.random((h, w)) < 0.004).astype(np.uint8) * 7,                                      # speckle: closing fills between dots
             np.pad(np.zeros((h - 40, w - 40), np.uint8), 20, constant_values=200),                  # frame touching all four borders
             np.full((h, w), 1, np.uint8), np.zeros((h, w), np.uint8)]
    for m in masks:
        kg, cg, ug = E.MovingKeyPoints(m, lab, ids, rm, kp, counts)
        kr, cr, ur = P.moving_keypoints(m, lab, ids, rm, kp, counts)
        assert kp_equal(kg, kr) and np.array_equal(cg, cr) and kp_equal(ug, ur)
    assert len(E.MovingKeyPoints(masks[4], lab, ids, rm, kp, counts)[0]) == 0


@pytest.mark.parametrize("cfg", [(640, 480, 1000, 5, 300), (1920, 1080, 1000, 3, 310)])
def test_masked_batch_equals_two_stage_path(orbx, cfg):
    """Config C5 (batched extraction with dynamic-mask culling): the batched call must equal, frame by frame,
    detect -> MovingKeyPoints(mask, no flagged super-pixel) -> ProcessDesp, which the tests above pin to the reference."""
    w, h, nf, B, seed = cfg
    E = orbx.ORBextractor(nf, 1.2, 8, 20, 7)
    imgs = synth_batch(B, w, h, seed0=seed, distinct=B)
    masks = np.stack([synth_mask(seed + b, w, h) for b in range(B)])
    masks[B - 1] = 0                                                               # one frame without dynamic objects
    kp, desc, counts, culled = E.extract_masked_batch(imgs, masks)
    lab = np.ones((h, w), np.float64); ids = np.zeros(1, np.int32); rm = np.zeros(1, np.int32)
    E2 = orbx.ORBextractor(nf, 1.2, 8, 20, 7)
    for b in range(B):
        kd, cd = E2.detect(imgs[b])
        kk, ck, cu = E2.MovingKeyPoints(masks[b], lab, ids, rm, kd, cd)
        kf, df = E2.ProcessDesp(kk, ck)
        assert counts[b] == len(kf) and culled[b] == len(cu)
        assert kp_equal(kp[b][:counts[b]], kf) and np.array_equal(desc[b][:counts[b]], df)
    assert culled[:B - 1].sum() > 0 and culled[B - 1] == 0
    k0, d0 = E2(imgs[B - 1])                                                       # empty mask == plain operator()
    assert kp_equal(kp[B - 1][:counts[B - 1]], k0) and np.array_equal(desc[B - 1][:counts[B - 1]], d0)
    assert E.check_overflow() == 0


def test_describe_requires_state(orbx):
    E = orbx.ORBextractor(300, 1.2, 8, 20, 7)
    with pytest.raises(orbx.OrbxError) as e:
        E.ProcessDesp(np.zeros(1, orbx.KP_DTYPE), np.array([1, 0, 0, 0, 0, 0, 0, 0], np.int32))
    assert e.value.code == orbx.E_STATE


def test_pyramid_export_with_reflect101_border(orbx, oracle):
    """mvImagePyramid export: ROI and the reference's padded parent buffer (copyMakeBorder REFLECT_101, 19 px)."""
    E = orbx.ORBextractor(500, 1.2, 8, 20, 7); P = oracle.Extractor("port", 500, 1.2, 8, 20, 7)
    img = synth_frame(80, 400, 300)
    E(img); P.extract(img)
    lib = oracle.port_lib()
    for l in (0, 3, 7):
        roi = P.pyramid_level(l)
        assert np.array_equal(E.pyramid_level(l, 0), roi)
        padded = np.zeros((roi.shape[0] + 38, roi.shape[1] + 38), np.uint8)
        lib.cvl_c_border101(np.ascontiguousarray(roi), roi.shape[1], roi.shape[0], 19, padded)
        assert np.array_equal(E.pyramid_level(l, 19), padded)
    if oracle.have_ref():
        Rf = oracle.Extractor("ref", 500, 1.2, 8, 20, 7); Rf.extract(img)
        roi = Rf.pyramid_level(2)
        padded = np.zeros((roi.shape[0] + 38, roi.shape[1] + 38), np.uint8)
        oracle.ref_lib().ref_pyramid_level_padded(Rf.h, 2, padded)
        assert np.array_equal(E.pyramid_level(2, 19), padded)


def test_two_handles_two_threads(orbx, oracle):
    """Left/right extractors run concurrently in two threads in the reference (Frame.cc:165-173)."""
    imgs = [synth_frame(90, 640, 480), synth_frame(91, 640, 480)]
    exts = [orbx.ORBextractor(1000, 1.2, 8, 20, 7) for _ in range(2)]
    out = [None, None]

    def run(i):
        for _ in range(5):
            out[i] = exts[i](imgs[i])
    th = [threading.Thread(target=run, args=(i,)) for i in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    P = oracle.Extractor("port", 1000, 1.2, 8, 20, 7)
    for i in range(2):
        ko, do = P.extract(imgs[i])
        assert_same(out[i][0], out[i][1], ko, do)


def test_getters(orbx, oracle):
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7); P = oracle.Extractor("port", 1000, 1.2, 8, 20, 7)
    assert E.GetLevels() == 8 and abs(E.GetScaleFactor() - 1.2) < 1e-6
    assert np.array_equal(E.GetScaleFactors(), P.scale_factors)
    assert np.array_equal(E.GetInverseScaleFactors(), np.float32(1.0) / P.scale_factors)
    assert np.array_equal(E.GetScaleSigmaSquares(), P.scale_factors * P.scale_factors)
    assert np.array_equal(E.features_per_level(), P.features_per_level)


def test_device_sincos_is_glibc_sincosf(orbx, oracle):
    """det_sincos on the device == the oracle's restatement of glibc's sincosf (itself pinned against the live libm for every float,
    tests/test_oracle_cvlite.py): 40 blocks of 2^22 consecutive bit patterns spread over [0, 7.0) -- every octave of the angle range,
    both reduction branches, all four quadrants -- plus blocks of negative, large (reduce_large) and non-finite arguments."""
    E = orbx.ORBextractor(500, 1.2, 8, 20, 7)
    lib = oracle.port_lib()
    hi = int(np.float32(7.0).view(np.uint32)); n = 1 << 22
    starts = [int(x) for x in np.linspace(0, hi - n, 32)] + [int(np.float32(v).view(np.uint32)) - n // 2 for v in (np.pi / 4, np.pi / 2, np.pi, 3 * np.pi / 2, 2 * np.pi)]
    starts += [0x80000000 + starts[20], int(np.float32(119.0).view(np.uint32)), int(np.float32(1e6).view(np.uint32)), 0x7F800000 - n // 2]
    for lo in starts:
        sg, cg = E.debug_sincos(lo, n)
        x = (np.arange(n, dtype=np.uint64) + lo).astype(np.uint32).view(np.float32)
        so, co = np.zeros(n, np.float32), np.zeros(n, np.float32)
        lib.port_det_sincos(x, so, co, n)
        for g, o in ((sg, so), (cg, co)):                          # bit for bit; NaN results (inf / nan arguments) only have to be NaN on both sides
            nan = np.isnan(o)
            assert np.array_equal(np.isnan(g), nan) and np.array_equal(g.view(np.uint32)[~nan], o.view(np.uint32)[~nan]), hex(lo)


def test_descriptors_at_sincos_sensitive_angles(orbx, oracle):
    """>= 10^5 ADVERSARIAL orientations -- angles where glibc's sincosf (what the reference calls, src/ORBextractor.cc:181) differs from
    the correctly rounded sin / cos, i.e. exactly the keypoints a 'mathematically right' sin/cos would get wrong -- pushed through
    ProcessDesp on a textured frame: descriptors must equal the reference build's (oracle/_ref where it travelled, else the port)."""
    lib = oracle.port_lib()
    rng = np.random.default_rng(5)
    deg = (rng.random(6_000_000) * 360.0).astype(np.float32)
    rad = deg * np.float32(np.pi / 180.0)                      # angle * factorPI in float, as :178 does
    s, c = np.zeros_like(rad), np.zeros_like(rad)
    lib.port_libm_sincosf(rad, s, c, len(rad))
    adv = (s != np.sin(rad.astype(np.float64)).astype(np.float32)) | (c != np.cos(rad.astype(np.float64)).astype(np.float32))
    deg = deg[adv]
    assert len(deg) >= 100_000, len(deg)
    deg = deg[:120_000]
    img = synth_frame(77, 320, 240)
    E = orbx.ORBextractor(500, 1.2, 8, 20, 7)
    kind = "ref" if oracle.have_ref() else "port"
    P = oracle.Extractor(kind, 500, 1.2, 8, 20, 7)
    kp, counts = E.detect(img); ko, co = P.detect(img)
    assert kp_equal(kp, ko)
    # level-0 keypoints re-used round-robin, each with one adversarial angle
    base = kp[:counts[0]]
    big = np.repeat(base[:1], len(deg)); big[:] = base[np.arange(len(deg)) % len(base)]; big["angle"] = deg
    cnt = np.zeros(8, np.int32); cnt[0] = len(big)
    kg, dg = E.ProcessDesp(big, cnt)
    kr, dr = P.process_desp(big, cnt)
    assert kp_equal(kg, kr)
    nbad = int((dg != dr).any(axis=1).sum())
    assert nbad == 0, "%d of %d sincos-sensitive keypoints differ" % (nbad, len(deg))


def test_masked_batch_with_label_map_vs_reference(orbx, oracle):
    """Batched C5 path WITH the super-pixel term of MovingKeyPoints (src/ORBextractor.cc:1722-1736): frame 0 is the reference build's
    golden run (24-px block label map, tests/golden/ref_extract.npz); the other frames carry the 5x5-block label map of SURVEY.md 8d and
    are checked against the oracle's detect -> MovingKeyPoints -> ProcessDesp (oracle/_ref where it travelled, else the port)."""
    from tools.synth import synth_labels
    w, h, B = 480, 360, 5
    E = orbx.ORBextractor(800, 1.2, 8, 20, 7)
    imgs = np.stack([synth_frame(int(R["amos_frame_seed"]) + b, w, h) for b in range(B)])
    masks = np.stack([R["amos_mask"]] + [synth_mask(600 + b, w, h) for b in range(1, B)])
    labs = [(R["amos_label"], R["amos_centers_id"], R["amos_rm"])] + [synth_labels(b, w, h) for b in range(1, B)]
    masks[B - 1] = 0                                                               # last frame: only the super-pixel term culls
    nl = max(len(l[1]) for l in labs)
    flagged = np.zeros((B, nl), np.uint8)
    for b, (_, cid, rm) in enumerate(labs):
        flagged[b, :len(cid)] = orbx.label_flags(cid, rm)
    labels = np.stack([l[0] for l in labs])
    kp, desc, counts, culled = E.extract_masked_batch(imgs, masks, labels=labels, flagged=flagged)
    assert counts[0] == len(R["amos_final_kp"]) and kp_equal(kp[0][:counts[0]], R["amos_final_kp"]) and np.array_equal(desc[0][:counts[0]], R["amos_final_desc"])
    assert culled[0] == len(R["amos_culled"])
    P = oracle.Extractor("ref" if oracle.have_ref() else "port", 800, 1.2, 8, 20, 7)
    for b in range(1, B):
        kd, cd = P.detect(imgs[b])
        kk, ck, cu = P.moving_keypoints(masks[b], labs[b][0], labs[b][1], labs[b][2], kd, cd)
        kf, df = P.process_desp(kk, ck)
        assert counts[b] == len(kf) and culled[b] == len(cu), b
        assert kp_equal(kp[b][:counts[b]], kf) and np.array_equal(desc[b][:counts[b]], df), b
    assert culled[B - 1] > 0
    # without labels the same call drops fewer keypoints (mask term only)
    _, _, counts_m, culled_m = E.extract_masked_batch(imgs, masks)
    assert (culled_m <= culled).all() and culled_m[B - 1] == 0 and culled_m.sum() < culled.sum()
    assert E.check_overflow() == 0


@pytest.mark.gpu
def test_no_kernel_writes_outside_its_buffers(tmp_path):
    """Sanitizer substitute (compute-sanitizer is closed on this pool): with ORBX_CANARY=1 every device buffer of the extractor handles
    carries 256 guard bytes on both sides; after single-frame, batched, two-stage and masked runs on several geometries -- including
    widths that are not multiples of 4 / 16 and a tiny frame -- none of the guard zones may have changed."""
    import subprocess, sys, textwrap
    code = textwrap.dedent('''
        import importlib, sys, numpy as np
        sys.path.insert(0, %r)
        from tools.synth import synth_frame, synth_batch, synth_mask
        orbx = importlib.import_module("amos-slam_b200")
        for (w, h, nf) in [(640, 480, 1000), (641, 479, 700), (1241, 376, 2000), (203, 147, 300), (120, 100, 100)]:
            E = orbx.ORBextractor(nf, 1.2, 8, 20, 7)
            A = synth_frame(3, w, h)
            k, d = E(A)
            kd, cd = E.detect(A)
            m = synth_mask(1, w, h); lab = np.ones((h, w)); ids = np.zeros(1, np.int32); rm = np.zeros(1, np.int32)
            kk, ck, _ = E.MovingKeyPoints(m, lab, ids, rm, kd, cd)
            E.ProcessDesp(kk, ck)
            fr = synth_batch(5, w, h, seed0=10)
            E.extract_batch(fr)
            E.extract_masked_batch(fr, np.stack([synth_mask(i, w, h) for i in range(5)]))
        bad, n = orbx.debug_canary_check()
        print("CANARY", bad, n)
    ''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ); env["ORBX_CANARY"] = "1"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("CANARY")][-1].split()
    assert int(line[1]) == 0 and int(line[2]) >= 20, line


@pytest.mark.gpu
def test_every_kernel_form_gives_the_same_result(tmp_path):
    """The latency / throughput forms of the stages are chosen by batch size; every one of them must be bit-identical.  Each environment
    below forces one form for BOTH a single frame and a small batch (resize: TMA tiles / per-thread / one-launch tile pyramid / 8-CTA
    cluster chain; blur: long / short strips; FAST: 1 / 8 cells per warp, one CTA per cell; quadtree: sort + tree pair / one-launch kernel; masks packed on the host / on the device) and the digest of all
    outputs must equal the digest of the default run -- which the other tests pin to the reference."""
    import subprocess, sys, textwrap, hashlib
    code = textwrap.dedent('''
        import importlib, sys, hashlib, numpy as np
        sys.path.insert(0, %r)
        from tools.synth import synth_frame, synth_batch, synth_mask
        orbx = importlib.import_module("amos-slam_b200")
        hsh = hashlib.sha256()
        for (w, h, nf) in [(640, 480, 1000), (417, 301, 500)]:
            E = orbx.ORBextractor(nf, 1.2, 8, 20, 7)
            k, d = E(synth_frame(5, w, h)); hsh.update(k.tobytes()); hsh.update(d.tobytes())
            fr = synth_batch(6, w, h, seed0=20)
            kb, db, nb = E.extract_batch(fr)
            for b in range(6): hsh.update(np.ascontiguousarray(kb[b][:nb[b]]).tobytes()); hsh.update(np.ascontiguousarray(db[b][:nb[b]]).tobytes())
            out = E.extract_masked_batch(fr, np.stack([synth_mask(i, w, h) for i in range(6)]))
            for b in range(6): hsh.update(np.ascontiguousarray(out[0][b][:out[2][b]]).tobytes()); hsh.update(np.ascontiguousarray(out[1][b][:out[2][b]]).tobytes()); hsh.update(bytes([int(out[3][b]) & 255]))
        print("DIGEST", hsh.hexdigest())
    ''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

    def run(extra):
        env = dict(os.environ); env.update(extra)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, str(extra) + r.stdout[-1500:] + r.stderr[-3000:]
        return [l for l in r.stdout.splitlines() if l.startswith("DIGEST")][-1]

    base = run({})
    for extra in ({"ORBX_TILEPYR": "0"}, {"ORBX_TILEPYR": "1"}, {"ORBX_TILEPYR": "0", "ORBX_RESIZE_TMA": "1"}, {"ORBX_TILEPYR": "0", "ORBX_RESIZE_TMA": "0"},
                  {"ORBX_CHAIN": "1"}, {"ORBX_BLUR_SMALL": "0"}, {"ORBX_BLUR_SMALL": "1"}, {"ORBX_FAST_CPW": "1"}, {"ORBX_FAST_CPW": "8"}, {"ORBX_FAST_CTA": "0"}, {"ORBX_FAST_CTA": "1"},
                  {"ORBX_HOST_PACK": "0"}, {"ORBX_HOST_PACK": "3"}, {"ORBX_GRAPH": "0"}, {"ORBX_QT_FUSED": "0"}, {"ORBX_QT_FUSED": "1"}, {"ORBX_QT_FUSED": "2"}, {"ORBX_QT_FUSED": "2", "ORBX_QT_BKEYS": "512"}):
        assert run(extra) == base, extra


@pytest.mark.gpu
def test_device_batch_with_unaligned_rows_and_base(orbx):
    """Device-resident frames whose rows (331 px) and base pointer are not 16-byte aligned are re-pitched for TMA by one kernel per batch
    (k_repitch: aligned word reads + funnel shift, byte path at both ends of a row): same result as the host-pointer call, for every base alignment."""
    import torch
    E = orbx.ORBextractor(500, 1.2, 8, 20, 7)
    B, h, w = 5, 203, 331
    fr = synth_batch(B, w, h, seed0=77)
    kb, db, nb = E.extract_batch(fr)
    cap = E.max_keypoints(h, w)
    for off in (0, 1, 2, 3, 5):
        buf = torch.full((B * h * w + 16,), 255, dtype=torch.uint8, device="cuda")
        buf[off:off + B * h * w] = torch.from_numpy(fr.reshape(-1)).cuda()
        kp = torch.zeros((B, cap, 28), dtype=torch.uint8, device="cuda"); ds = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda"); cn = torch.zeros(B, dtype=torch.int32, device="cuda")
        E.extract_batch_raw(buf.data_ptr() + off, B, h, w, w, w * h, kp.data_ptr(), ds.data_ptr(), cap, cn.data_ptr(), device=True)
        torch.cuda.synchronize()
        cnt = cn.cpu().numpy(); kph = kp.cpu().numpy(); dsh = ds.cpu().numpy()
        assert np.array_equal(cnt, nb), off
        for b in range(B):
            assert np.array_equal(kph[b, :cnt[b]].reshape(-1), np.ascontiguousarray(kb[b][:nb[b]]).view(np.uint8).reshape(-1)), (off, b)
            assert np.array_equal(dsh[b, :cnt[b]], db[b][:nb[b]]), (off, b)


@pytest.mark.gpu
def test_large_feature_counts_and_sizes(orbx, oracle):
    """Capacity edges of the quadtree stage: node pools above the default (6000 / 9500 features), the sort + tree fallback (12000 features: more nodes per level than
    the one-launch kernel's pool), levels whose candidates exceed the shared-memory key capacity (1080p, 2200 x 1300), 12 levels, 3 levels -- single frame
    (wide form) and a batch of 6 (lean form), against the oracle."""
    for (w, h, nf, L, sf) in [(1241, 376, 6000, 8, 1.2), (1241, 376, 9500, 8, 1.2), (1241, 376, 12000, 8, 1.2), (1920, 1080, 9000, 8, 1.2), (640, 480, 4000, 12, 1.1),
                              (2200, 1300, 2000, 8, 1.2), (333, 277, 800, 3, 1.5)]:
        img = synth_frame(42, w, h)
        E = orbx.ORBextractor(nf, sf, L, 20, 7); O = oracle.Extractor("port", nf, sf, L, 20, 7)
        kg, dg = E(img); kr, dr = O.extract(img)
        assert kp_equal(kg, kr) and np.array_equal(dg, dr), (w, h, nf, L)
        kb, db, nb = E.extract_batch(np.stack([img] * 6))
        for b in range(6):
            assert nb[b] == len(kr) and kp_equal(kb[b][:nb[b]], kr) and np.array_equal(db[b][:nb[b]], dr), (w, h, nf, L, b)
        assert E.check_overflow() == 0
