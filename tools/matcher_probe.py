#!/usr/bin/env python3
"""A few matcher calls (one frame pair) for a kernel-level launch list: ncu --metrics gpu__time_duration.sum python tools/matcher_probe.py"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import match_cases as mc
orbx = importlib.import_module("amos-slam_b200")
E = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
ka, da, kb, db = mc.mono_pair(lambda img: E(img))
sf = E.GetScaleFactors()
FA, FB = orbx.FrameView(ka, da, 640, 480, sf), orbx.FrameView(kb, db, 640, 480, sf)
M = orbx.ORBmatcher(0.9, True)
prev = np.stack([ka["x"], ka["y"]], 1)
pi = mc.projection_inputs(ka, kb); uv, iz = mc.project(pi["xyz"])
FBu = orbx.FrameView(kb, db, 640, 480, sf, u_right=pi["u_right"])
for _ in range(4):
    M.SearchForInitialization(FA, FB, prev, 100)
    M.SearchByProjectionFrame(FBu, uv, iz, ka["octave"], ka["angle"], da, pi["valid"], pi["obs"], pi["occ"], 15.0, False, False, 40.0)
    M.SearchByProjectionPoints(FBu, pi["tuv"], pi["tur"], pi["lvl"], pi["vc"], da, pi["obs"], pi["occ"], 3.0)
print("ok")
