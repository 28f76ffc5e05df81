"""Seeded matcher inputs shared by the golden generator and the tests (CPU and GPU): the same arrays are rebuilt from
seeds everywhere, so only the reference's OUTPUTS need to be committed as fixtures."""
import numpy as np
from tools.synth import synth_frame, warp_affine_nn, stereo_right_from_left

FX, FY, CX, CY = 517.3, 516.5, 318.6, 255.3      # TUM1-like intrinsics (Examples/RGB-D/TUM1.yaml)
BF_KITTI = 386.1448                               # Examples/Stereo/KITTI00-02.yaml:25


def mono_pair(extract, seed=0, width=640, height=480):
    """Frame A, frame B = A warped by a small affine; returns keypoints/descriptors of both (extract: img -> (kp, desc))."""
    A = synth_frame(seed, width, height); B = warp_affine_nn(A, 7, -4, 2.0)
    ka, da = extract(A); kb, db = extract(B)
    return ka, da, kb, db


def projection_inputs(ka, kb, seed=1):
    rng = np.random.default_rng(seed)
    n = len(ka)
    z = rng.uniform(0.5, 8, n).astype(np.float32)
    # camera-frame points that project close to A's keypoints shifted by the known motion (+7,-4)
    xyz = np.stack([((ka['x'] + 7 + rng.normal(0, 1.5, n) - CX) / FX * z), ((ka['y'] - 4 + rng.normal(0, 1.5, n) - CY) / FY * z), z], 1).astype(np.float32)
    xyz[::37, 2] *= -1                                      # some points behind the camera (invz < 0)
    valid = (rng.random(n) < 0.8).astype(np.uint8); obs = (rng.random(n) < 0.6).astype(np.uint8)
    occ = (rng.random(len(kb)) < 0.1).astype(np.uint8)
    u_right = np.where(rng.random(len(kb)) < 0.5, kb['x'] - rng.uniform(1, 30, len(kb)).astype(np.float32), -1).astype(np.float32)
    tuv = np.stack([ka['x'] + 7 + rng.normal(0, 2, n), ka['y'] - 4 + rng.normal(0, 2, n)], 1).astype(np.float32)
    tur = (tuv[:, 0] - rng.uniform(1, 30, n)).astype(np.float32)
    lvl = np.clip(ka['octave'] + rng.integers(-1, 2, n), 0, 7).astype(np.int32)
    vc = rng.uniform(0.99, 1.0, n).astype(np.float32)
    return dict(xyz=xyz, valid=valid, obs=obs, occ=occ, u_right=u_right, tuv=tuv, tur=tur, lvl=lvl, vc=vc)


def project(xyz):
    """ORBmatcher.cc:1608-1618 in float32, as the caller of the C ABI computes it (identity pose)."""
    xc, yc = xyz[:, 0].astype(np.float32), xyz[:, 1].astype(np.float32)
    invz = (1.0 / xyz[:, 2].astype(np.float64)).astype(np.float32)
    u = (np.float32(FX) * xc) * invz + np.float32(CX)
    v = (np.float32(FY) * yc) * invz + np.float32(CY)
    return np.stack([u, v], 1).astype(np.float32), invz


def project_pose(xyz_w, Rcw, tcw, bounds):
    """x3Dc = Rcw * x3Dw + tcw as cv::gemm evaluates it for 3 x 3 by 3 x 1 floats (((r0 x + r1 y) + r2 z) + t, every operation rounded
    to float; pinned against cv2.gemm by tests/test_oracle_cvlite.py), then ORBmatcher.cc:1611-1623 in float32.  Returns uv, invz, valid."""
    f = np.float32
    R = np.asarray(Rcw, f).reshape(3, 3); t = np.asarray(tcw, f).reshape(3); X = np.asarray(xyz_w, f)
    c = [(((R[r, 0] * X[:, 0]).astype(f) + (R[r, 1] * X[:, 1]).astype(f)).astype(f) + (R[r, 2] * X[:, 2]).astype(f)).astype(f) + t[r] for r in range(3)]
    c = [v.astype(f) for v in c]
    with np.errstate(divide="ignore", invalid="ignore"):
        invz = (1.0 / c[2].astype(np.float64)).astype(f)
        u = ((f(FX) * c[0]).astype(f) * invz).astype(f) + f(CX)
        v = ((f(FY) * c[1]).astype(f) * invz).astype(f) + f(CY)
    minx, maxx, miny, maxy = bounds
    valid = ~(invz < 0) & ~((u < minx) | (u > maxx) | (v < miny) | (v > maxy))
    uv = np.stack([u, v], 1).astype(f); uv[~valid] = 0; invz = invz.copy(); invz[~valid] = 0
    return uv, invz, valid.astype(np.uint8)


def small_pose(seed=4):
    """A camera motion of a few centimetres / half a degree (TrackWithMotionModel's situation): Rcw, tcw as float32."""
    rng = np.random.default_rng(seed)
    a = rng.normal(0, 0.008, 3)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    th = np.linalg.norm(a)
    R = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)
    return R.astype(np.float32), rng.normal(0, 0.03, 3).astype(np.float32)


PROJ_FRAME_CASES = [(15.0, 1), (7.0, 0), (15.0, 0)]         # (th, bMono); TrackWithMotionModel uses 15 / 7 (Tracking.cc:1929-1944)
PROJ_POINT_CASES = [1.0, 3.0, 5.0]                          # SearchLocalPoints th (Tracking.cc:2378-2389)


def stereo_pair(seed=3, width=1241, height=376):
    L = synth_frame(seed, width, height)
    return L, stereo_right_from_left(L, seed + 1)


KF_CASES = [(10.0, 100, True), (3.0, 64, True), (10.0, 100, False)]     # (th, ORBdist, checkOri); relocalisation uses (10, 100) and (3, 64) (Tracking.cc:2663-2720)


def keyframe_inputs(ka, kb, pi, seed=9):
    """Inputs of SearchByProjection(Frame&, KeyFrame*, set&, th, ORBdist): map-point state (0 none, 1 good, 2 bad, 3 already found),
    predicted level, scale-invariance distance range, occupancy of the current frame; `valid` is what the C-ABI caller derives."""
    rng = np.random.default_rng(seed); n = len(ka)
    state = rng.choice([0, 1, 1, 1, 1, 2, 3], n).astype(np.uint8)
    lvl = np.clip(ka["octave"] + rng.integers(-1, 2, n), 0, 7).astype(np.int32)
    xyz = pi["xyz"].astype(np.float32)
    d3 = np.sqrt((xyz.astype(np.float64) ** 2).sum(1)).astype(np.float32)                # float dist3D = cv::norm(PO)  (double accumulate, :1773-1774)
    mind = (d3 * rng.choice([0.5, 0.5, 1.2], n)).astype(np.float32); maxd = (d3 * rng.choice([2.0, 2.0, 0.8], n)).astype(np.float32)
    occ = (rng.random(len(kb)) < 0.15).astype(np.uint8)
    valid = ((state == 1) & ~((d3 < mind) | (d3 > maxd))).astype(np.uint8)
    return dict(state=state, lvl=lvl, mind=mind, maxd=maxd, occ=occ, valid=valid)


KFP_CASES = [10, 3]                                                     # th of LoopClosing (SearchByProjection(pKF, Scw, vpPoints, vpMatched, 10))


def keyframe_points_inputs(ka, kb, pi, seed=13):
    """Inputs of SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th): per map point its state (1 good, 2 bad, 3 already in vpMatched at
    KeyFrame feature found_at), whether its normal faces the camera, predicted level, distance range; kf_matched = vpMatched[j] != NULL."""
    rng = np.random.default_rng(seed); n = len(ka)
    state = rng.choice([1, 1, 1, 1, 2, 3], n).astype(np.uint8)
    found_at = rng.permutation(len(kb))[:n].astype(np.int32) if len(kb) >= n else rng.integers(0, len(kb), n).astype(np.int32)
    facing = (rng.random(n) < 0.85).astype(np.uint8)
    lvl = np.clip(ka["octave"] + rng.integers(0, 2, n), 0, 7).astype(np.int32)
    xyz = pi["xyz"].astype(np.float32)
    d3 = np.sqrt((xyz.astype(np.float64) ** 2).sum(1)).astype(np.float32)
    mind = (d3 * rng.choice([0.5, 0.5, 1.2], n)).astype(np.float32); maxd = (d3 * rng.choice([2.0, 2.0, 0.8], n)).astype(np.float32)
    kf_matched = (rng.random(len(kb)) < 0.1).astype(np.uint8)
    kf_matched[found_at[state == 3]] = 1
    # projection exactly as the body forms it (:417-421): x = X * invz, u = fx * x + cx
    invz = (np.float32(1.0) / xyz[:, 2]).astype(np.float32)
    u = (np.float32(FX) * (xyz[:, 0] * invz) + np.float32(CX)).astype(np.float32); v = (np.float32(FY) * (xyz[:, 1] * invz) + np.float32(CY)).astype(np.float32)
    inimg = (u >= 0) & (u < 640) & (v >= 0) & (v < 480)                  # KeyFrame::IsInImage
    valid = ((state == 1) & (xyz[:, 2] >= 0) & inimg & ~((d3 < mind) | (d3 > maxd)) & (facing > 0)).astype(np.uint8)
    return dict(state=state, found_at=found_at, facing=facing, lvl=lvl, mind=mind, maxd=maxd, kf_matched=kf_matched, valid=valid, uv=np.stack([u, v], 1).astype(np.float32))


SIM3_TH = [7.5, 3.0]                                                    # LoopClosing::ComputeSim3 uses 7.5


def sim3_inputs(ka, kb, seed=21):
    """Inputs of SearchBySim3 with an identity similarity and both KeyFrames at the origin: a map point of KeyFrame 1 sits where its feature
    back-projects, shifted by the known image motion, so that it lands near the matching feature of KeyFrame 2 (and vice versa)."""
    rng = np.random.default_rng(seed)

    def side(k, dx, dy, other_n):
        n = len(k)
        z = rng.uniform(0.5, 8, n).astype(np.float32)
        xyz = np.stack([(k["x"] + dx + rng.normal(0, 1.0, n) - CX) / FX * z, (k["y"] + dy + rng.normal(0, 1.0, n) - CY) / FY * z, z], 1).astype(np.float32)
        xyz[::41, 2] *= -1
        state = rng.choice([0, 1, 1, 1, 1, 2], n).astype(np.uint8)
        lvl = np.clip(k["octave"] + rng.integers(0, 2, n), 0, 7).astype(np.int32)
        d3 = np.sqrt((xyz.astype(np.float64) ** 2).sum(1)).astype(np.float32)            # cv::norm(p3Dc), double accumulate
        mind = (d3 * rng.choice([0.5, 0.5, 1.2], n)).astype(np.float32); maxd = (d3 * rng.choice([2.0, 2.0, 0.8], n)).astype(np.float32)
        invz = (1.0 / xyz[:, 2].astype(np.float64)).astype(np.float32)                   # const float invz = 1.0/z  (:1379)
        u = (np.float32(FX) * (xyz[:, 0] * invz) + np.float32(CX)).astype(np.float32); v = (np.float32(FY) * (xyz[:, 1] * invz) + np.float32(CY)).astype(np.float32)
        ok = (state == 1) & (xyz[:, 2] >= 0) & (u >= 0) & (u < 640) & (v >= 0) & (v < 480) & ~((d3 < mind) | (d3 > maxd))
        return dict(xyz=xyz, state=state, lvl=lvl, mind=mind, maxd=maxd, uv=np.stack([u, v], 1).astype(np.float32), ok=ok)

    s1, s2 = side(ka, 7, -4, len(kb)), side(kb, -7, 4, len(ka))
    already12 = np.full(len(ka), -1, np.int32)
    pick = rng.choice(len(ka), 40, replace=False); already12[pick] = rng.choice(len(kb), 40, replace=False)
    s1["valid"] = (s1["ok"] & (already12 < 0)).astype(np.uint8)
    am2 = np.zeros(len(kb), bool); am2[already12[already12 >= 0]] = True
    s2["valid"] = (s2["ok"] & ~am2).astype(np.uint8)
    return s1, s2, already12


FUSE_TH = [3.0, 4.0]                                                     # LocalMapping::SearchInNeighbors: Fuse(pKFi, vpMapPointMatches) with th = 3; LoopClosing: 4


def fuse_inputs(ka, kb, pi, seed=29):
    """Inputs of both ORBmatcher::Fuse forms (identity pose): map points near the features of KeyFrame B, state 1 good / 2 bad / 3 already in
    the KeyFrame, normals facing or not, and which KeyFrame features hold a map point."""
    rng = np.random.default_rng(seed); n = len(ka)
    state = rng.choice([1, 1, 1, 1, 2, 3], n).astype(np.uint8)
    facing = (rng.random(n) < 0.85).astype(np.uint8)
    lvl = np.clip(ka["octave"] + rng.integers(0, 2, n), 0, 7).astype(np.int32)
    z = rng.uniform(0.5, 8, n).astype(np.float32)
    xyz = np.stack([(ka["x"] + 7 + rng.normal(0, 0.7, n) - CX) / FX * z, (ka["y"] - 4 + rng.normal(0, 0.7, n) - CY) / FY * z, z], 1).astype(np.float32)
    xyz[::43, 2] *= -1
    d3 = np.sqrt((xyz.astype(np.float64) ** 2).sum(1)).astype(np.float32)
    mind = (d3 * rng.choice([0.5, 0.5, 1.2], n)).astype(np.float32); maxd = (d3 * rng.choice([2.0, 2.0, 0.8], n)).astype(np.float32)
    kf_has_mp = (rng.random(len(kb)) < 0.5).astype(np.uint8)
    invz = (np.float32(1.0) / xyz[:, 2]).astype(np.float32)
    u = (np.float32(FX) * (xyz[:, 0] * invz) + np.float32(CX)).astype(np.float32); v = (np.float32(FY) * (xyz[:, 1] * invz) + np.float32(CY)).astype(np.float32)
    ur = (u - np.float32(40.0) * invz).astype(np.float32)
    inimg = (u >= 0) & (u < 640) & (v >= 0) & (v < 480)
    valid = ((state == 1) & (xyz[:, 2] >= 0) & inimg & ~((d3 < mind) | (d3 > maxd)) & (facing > 0)).astype(np.uint8)
    sf = (np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32)
    inv_sigma2 = (np.float32(1.0) / (sf * sf)).astype(np.float32)                          # mvInvLevelSigma2 (ORBextractor.cc:519-523)
    return dict(state=state, facing=facing, lvl=lvl, xyz=xyz, mind=mind, maxd=maxd, kf_has_mp=kf_has_mp, uv=np.stack([u, v], 1).astype(np.float32), ur=ur, valid=valid,
                inv_sigma2=inv_sigma2)
