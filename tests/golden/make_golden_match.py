#!/usr/bin/env python3
"""Golden OUTPUTS of the reference's own matcher bodies (oracle/_ref) on the seeded inputs of tests/match_cases.py.
Run in the build container (needs /root/reference): python tests/golden/make_golden_match.py"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))
import oracle
import match_cases as mc

assert oracle.have_ref()
g = {}
R = oracle.Extractor("ref", 1000)
ka, da, kb, db = mc.mono_pair(R.extract)
g["n_a"], g["n_b"] = np.int64(len(ka)), np.int64(len(kb))
pi = mc.projection_inputs(ka, kb)
FA = oracle.FrameData(ka, da, 640, 480, R.scale_factors); FB = oracle.FrameData(kb, db, 640, 480, R.scale_factors)
FBu = oracle.FrameData(kb, db, 640, 480, R.scale_factors, u_right=pi["u_right"])
m = oracle.Matcher("ref", 0.9, True)
g["dist"] = m.descriptor_distance(da[:900], db[:900])
prev = np.stack([ka["x"], ka["y"]], 1)
nm, m12, prev2 = m.search_for_initialization(FA, FB, prev, 100)
g["init_nm"], g["init_m12"], g["init_prev"] = np.int64(nm), m12, prev2
nm, m12, prev2 = oracle.Matcher("ref", 0.9, False).search_for_initialization(FA, FB, prev, 30)
g["init2_nm"], g["init2_m12"], g["init2_prev"] = np.int64(nm), m12, prev2
for i, (th, mono) in enumerate(mc.PROJ_FRAME_CASES):
    nm, cm, uv, iz = m.search_by_projection_frame_ref(FBu, pi["xyz"], ka["octave"], ka["angle"], da, pi["valid"], pi["obs"], pi["occ"], th, mono, 0.08, 40.0, mc.FX, mc.FY, mc.CX, mc.CY)
    g["pf%d_nm" % i], g["pf%d_cm" % i] = np.int64(nm), cm
    uv2, iz2 = mc.project(pi["xyz"])
    assert np.array_equal(uv, uv2) and np.array_equal(iz, iz2), "caller-side projection must equal the reference body's"
for i, th in enumerate(mc.PROJ_POINT_CASES):
    nm, fm = oracle.Matcher("ref", 0.8, True).search_by_projection_points(FBu, pi["tuv"], pi["tur"], pi["lvl"], pi["vc"], da, pi["obs"], pi["occ"], th)
    g["pp%d_nm" % i], g["pp%d_fm" % i] = np.int64(nm), fm
L, Rimg = mc.stereo_pair()
RL, RR = oracle.Extractor("ref", 2000), oracle.Extractor("ref", 2000)
kl, dl = RL.extract(L); kr, dr = RR.extract(Rimg)
ur, dep = m.compute_stereo_matches(RL, RR, kl, dl, kr, dr, 0.0, mc.BF_KITTI)
g["stereo_ur"], g["stereo_depth"], g["stereo_nl"] = ur, dep, np.int64(len(kl))
ur, dep = m.compute_stereo_matches(RL, RR, kl, dl, kr, dr, 0.5372, mc.BF_KITTI)      # mb as after the constructor (finite maxD)
g["stereo2_ur"], g["stereo2_depth"] = ur, dep
np.savez_compressed(os.path.join(HERE, "ref_match.npz"), **g)
print("ref_match.npz", os.path.getsize(os.path.join(HERE, "ref_match.npz")), "init", int(g["init_nm"]), "stereo", int((g["stereo_ur"] >= 0).sum()), int((g["stereo2_ur"] >= 0).sum()))
