"""GPU vs oracle probe for the matchers (python tests/gpu_match_probe.py on the GPU box)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from tools.synth import synth_frame, warp_affine_nn, stereo_right_from_left
orbx = importlib.import_module("amos-slam_b200")

def main():
    rng = np.random.default_rng(0)
    P = oracle.Extractor('port', 1000)
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    A = synth_frame(0); B = warp_affine_nn(A, 7, -4, 2.0)
    ka, da = E(A); kb, db = E(B)
    sf = E.GetScaleFactors()
    M = orbx.ORBmatcher(0.9, True); mp = oracle.Matcher('port', 0.9, True)
    print('dist', np.array_equal(M.DescriptorDistance(da[:900], db[:900]), mp.descriptor_distance(da[:900], db[:900])))
    FA, FB = orbx.FrameView(ka, da, 640, 480, sf), orbx.FrameView(kb, db, 640, 480, sf)
    oA, oB = oracle.FrameData(ka, da, 640, 480, sf), oracle.FrameData(kb, db, 640, 480, sf)
    prev = np.stack([ka['x'], ka['y']], 1)
    g = M.SearchForInitialization(FA, FB, prev, 100); o = mp.search_for_initialization(oA, oB, prev, 100)
    print('init', g[0], o[0], np.array_equal(g[1], o[1]), np.array_equal(g[2], o[2]))
    ur = np.where(rng.random(len(kb)) < 0.5, kb['x'] - rng.uniform(1, 30, len(kb)).astype(np.float32), -1).astype(np.float32)
    FBu, oBu = orbx.FrameView(kb, db, 640, 480, sf, u_right=ur), oracle.FrameData(kb, db, 640, 480, sf, u_right=ur)
    n = len(ka)
    uv = np.stack([ka['x'] + 7 + rng.normal(0, 1.5, n), ka['y'] - 4 + rng.normal(0, 1.5, n)], 1).astype(np.float32)
    iz = (1.0 / rng.uniform(0.5, 8, n)).astype(np.float32)
    valid = (rng.random(n) < 0.8).astype(np.uint8); obs = (rng.random(n) < 0.6).astype(np.uint8); occ = (rng.random(len(kb)) < 0.1).astype(np.uint8)
    for th, fw, bw in [(15, 0, 0), (7, 1, 0), (15, 0, 1)]:
        g = M.SearchByProjectionFrame(FBu, uv, iz, ka['octave'], ka['angle'], da, valid, obs, occ, th, fw, bw, 40.0)
        o = mp.search_by_projection_frame_port(oBu, uv, iz, ka['octave'], ka['angle'], da, valid, obs, occ, th, fw, bw, 40.0)
        print('proj frame', th, fw, bw, g[0], o[0], np.array_equal(g[1], o[1]))
    tur = (uv[:, 0] - rng.uniform(1, 30, n)).astype(np.float32)
    lvl = np.clip(ka['octave'] + rng.integers(-1, 2, n), 0, 7).astype(np.int32); vc = rng.uniform(0.99, 1.0, n).astype(np.float32)
    for th in (1.0, 3.0, 5.0):
        g = orbx.ORBmatcher(0.8, True).SearchByProjectionPoints(FBu, uv, tur, lvl, vc, da, obs, occ, th)
        o = oracle.Matcher('port', 0.8, True).search_by_projection_points(oBu, uv, tur, lvl, vc, da, obs, occ, th)
        print('proj points', th, g[0], o[0], np.array_equal(g[1], o[1]))
    L = synth_frame(3, 1241, 376); R = stereo_right_from_left(L, 1)
    EL, ER = orbx.ORBextractor(2000, 1.2, 8, 20, 7), orbx.ORBextractor(2000, 1.2, 8, 20, 7)
    PL, PR = oracle.Extractor('port', 2000), oracle.Extractor('port', 2000)
    kl, dl = EL(L); kr, dr = ER(R); kl2, dl2 = PL.extract(L); kr2, dr2 = PR.extract(R)
    print('stereo extract equal', np.array_equal(kl, kl2), np.array_equal(dr, dr2))
    u1, d1 = M.ComputeStereoMatches(EL, ER, kl, dl, kr, dr, 0.0, 386.1448)
    u2, d2 = mp.compute_stereo_matches(PL, PR, kl, dl, kr, dr, 0.0, 386.1448)
    print('stereo', np.array_equal(u1, u2), np.array_equal(d1, d2), int((u1 >= 0).sum()), int((u2 >= 0).sum()), len(kl))
    if not np.array_equal(u1, u2):
        bad = np.nonzero(u1 != u2)[0]; print(bad[:10], u1[bad[:10]], u2[bad[:10]])

if __name__ == "__main__":
    main()
