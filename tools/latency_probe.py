#!/usr/bin/env python3
"""Per-call latency of the drop-in entry points as Tracking would use them (one frame / one pair per call, host pointers).
usage (GPU box): python tools/latency_probe.py"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from tools.synth import synth_frame, synth_mask
import match_cases as mc

orbx = importlib.import_module("amos-slam_b200")


def timeit(fn, reps=200, warm=10):
    for _ in range(warm):
        fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t) / reps * 1e6


E = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
A = synth_frame(0, 640, 480)
print("orbx_extract 640x480/1000            %8.1f us" % timeit(lambda: E(A)))
print("orbx_detect  640x480/1000            %8.1f us" % timeit(lambda: E.detect(A)))
kd, cd = E.detect(A)
m = synth_mask(1, 640, 480); lab = np.ones((480, 640)); ids = np.zeros(1, np.int32); rm = np.zeros(1, np.int32)
print("orbx_cull (MovingKeyPoints)          %8.1f us" % timeit(lambda: E.MovingKeyPoints(m, lab, ids, rm, kd, cd)))
kk, ck, _ = E.MovingKeyPoints(m, lab, ids, rm, kd, cd)
print("orbx_describe (ProcessDesp)          %8.1f us" % timeit(lambda: E.ProcessDesp(kk, ck)))
E2 = orbx.ORBextractor(2000, 1.2, 8, 20, 7)
L, R = mc.stereo_pair()
print("orbx_extract 1241x376/2000           %8.1f us" % timeit(lambda: E2(L)))
ka, da, kb, db = mc.mono_pair(lambda img: E(img))
sf = E.GetScaleFactors()
FA, FB = orbx.FrameView(ka, da, 640, 480, sf), orbx.FrameView(kb, db, 640, 480, sf)
M = orbx.ORBmatcher(0.9, True)
prev = np.stack([ka["x"], ka["y"]], 1)
print("SearchForInitialization              %8.1f us" % timeit(lambda: M.SearchForInitialization(FA, FB, prev, 100)))
pi = mc.projection_inputs(ka, kb); uv, iz = mc.project(pi["xyz"])
FBu = orbx.FrameView(kb, db, 640, 480, sf, u_right=pi["u_right"])
print("SearchByProjection(Frame,Frame)      %8.1f us" % timeit(lambda: M.SearchByProjectionFrame(FBu, uv, iz, ka["octave"], ka["angle"], da, pi["valid"], pi["obs"], pi["occ"], 15.0, False, False, 40.0)))
print("SearchByProjection(Frame,MapPoints)  %8.1f us" % timeit(lambda: M.SearchByProjectionPoints(FBu, pi["tuv"], pi["tur"], pi["lvl"], pi["vc"], da, pi["obs"], pi["occ"], 3.0)))
EL, ER = orbx.ORBextractor(2000, 1.2, 8, 20, 7), orbx.ORBextractor(2000, 1.2, 8, 20, 7)
kl, dl = EL(L); kr, dr = ER(R)
print("ComputeStereoMatches 2000x2000       %8.1f us" % timeit(lambda: M.ComputeStereoMatches(EL, ER, kl, dl, kr, dr, 0.0, mc.BF_KITTI), reps=100))
