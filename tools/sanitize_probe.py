#!/usr/bin/env python3
"""Small end-to-end pass over every entry point, meant to be run under compute-sanitizer --tool memcheck (one tool per call)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from tools.synth import synth_frame, synth_batch, synth_mask
import match_cases as mc
orbx = importlib.import_module("amos-slam_b200")
E = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
for (w, h) in ((640, 480), (403, 301), (161, 123)):
    k, d = E(synth_frame(1, w, h)); print(w, h, len(k))
imgs = synth_batch(70, 320, 240, seed0=3, distinct=4)
kp, desc, counts = E.extract_batch(imgs); print("batch", counts[:4])
masks = np.stack([synth_mask(b, 320, 240) for b in range(70)])
kp, desc, counts, culled = E.extract_masked_batch(imgs, masks); print("masked", counts[:4], culled[:4])
A = synth_frame(0, 640, 480)
kd, cd = E.detect(A); kk, ck, cu = E.MovingKeyPoints(synth_mask(1, 640, 480), np.ones((480, 640)), np.zeros(1, np.int32), np.zeros(1, np.int32), kd, cd)
k3, d3 = E.ProcessDesp(kk, ck); print("amos", len(k3), len(cu)); print("pyr", E.pyramid_level(3, 19).shape)
ka, da, kb, db = mc.mono_pair(lambda img: E(img))
sf = E.GetScaleFactors()
FA, FB = orbx.FrameView(ka, da, 640, 480, sf), orbx.FrameView(kb, db, 640, 480, sf)
M = orbx.ORBmatcher(0.9, True)
print("init", M.SearchForInitialization(FA, FB, np.stack([ka["x"], ka["y"]], 1), 100)[0])
pi = mc.projection_inputs(ka, kb); uv, iz = mc.project(pi["xyz"])
FBu = orbx.FrameView(kb, db, 640, 480, sf, u_right=pi["u_right"])
print("pf", M.SearchByProjectionFrame(FBu, uv, iz, ka["octave"], ka["angle"], da, pi["valid"], pi["obs"], pi["occ"], 15.0, False, False, 40.0)[0])
print("pp", M.SearchByProjectionPoints(FBu, pi["tuv"], pi["tur"], pi["lvl"], pi["vc"], da, pi["obs"], pi["occ"], 3.0)[0])
print("dd", M.DescriptorDistance(da[:100], db[:100]).sum())
L, R = mc.stereo_pair()
EL, ER = orbx.ORBextractor(2000, 1.2, 8, 20, 7), orbx.ORBextractor(2000, 1.2, 8, 20, 7)
kl, dl = EL(L); kr, dr = ER(R)
ur, dep = M.ComputeStereoMatches(EL, ER, kl, dl, kr, dr, 0.0, mc.BF_KITTI); print("stereo", (ur >= 0).sum())
print("sanitize probe ok")
