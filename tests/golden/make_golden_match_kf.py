#!/usr/bin/env python3
"""Generates tests/golden/ref_match_kf.npz: outputs of the reference's own ORBmatcher::SearchByProjection(Frame&, KeyFrame*,
const set<MapPoint*>&, th, ORBdist) (src/ORBmatcher.cc:1731-1863) and SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)
(:388-512) bodies, compiled into oracle/_ref, on the seeded inputs of tests/match_cases.py.  usage: python tests/golden/make_golden_match_kf.py"""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle                      # noqa: E402
import match_cases as mc           # noqa: E402

assert oracle.build_ref(), "needs /root/reference"
E = oracle.Extractor("ref", 1000, 1.2, 8, 20, 7)
ka, da, kb, db = mc.mono_pair(lambda img: E.extract(img))
pi = mc.projection_inputs(ka, kb); kf = mc.keyframe_inputs(ka, kb, pi)
F = oracle.FrameData(kb, db, 640, 480, E.scale_factors)
out = {"n_a": np.array(len(ka)), "n_b": np.array(len(kb))}
for i, (th, od, ori) in enumerate(mc.KF_CASES):
    nm, cm, uv = oracle.Matcher("ref", 0.9, ori).search_by_projection_keyframe_ref(F, pi["xyz"], kf["lvl"], ka["angle"], da, kf["state"], kf["mind"], kf["maxd"], kf["occ"],
                                                                                 th, od, mc.FX, mc.FY, mc.CX, mc.CY)
    out["kf%d_nm" % i] = np.array(nm); out["kf%d_cm" % i] = cm; out["uv"] = uv
# SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)  (src/ORBmatcher.cc:388-512)
kp = mc.keyframe_points_inputs(ka, kb, pi)
for i, th in enumerate(mc.KFP_CASES):
    nm, km, uv2 = oracle.Matcher("ref", 0.9, True).search_by_projection_keyframe_points_ref(F, pi["xyz"], kp["lvl"], da, kp["state"], kp["found_at"], kp["facing"], kp["mind"], kp["maxd"],
                                                                                          kp["kf_matched"], th, mc.FX, mc.FY, mc.CX, mc.CY)
    out["kfp%d_nm" % i] = np.array(nm); out["kfp%d_km" % i] = km; out["uv2"] = uv2
# SearchBySim3  (src/ORBmatcher.cc:1290-1555), identity similarity
F1 = oracle.FrameData(ka, da, 640, 480, E.scale_factors)
s1, s2, already12 = mc.sim3_inputs(ka, kb)
for i, th in enumerate(mc.SIM3_TH):
    nf, m12 = oracle.Matcher("ref").search_by_sim3_ref(F1, F, dict(s1, desc=da), dict(s2, desc=db), already12, th, mc.FX, mc.FY, mc.CX, mc.CY)
    out["sim3_%d_nf" % i] = np.array(nf); out["sim3_%d_m12" % i] = m12
# both ORBmatcher::Fuse forms  (src/ORBmatcher.cc:1020-1175, 1179-1310): which KeyFrame feature every point was fused with
fu = mc.fuse_inputs(ka, kb, pi)
Fu = oracle.FrameData(kb, db, 640, 480, E.scale_factors, u_right=pi["u_right"])
for i, th in enumerate(mc.FUSE_TH):
    for sim3 in (0, 1):
        nf, best, uvf, urf = oracle.Matcher("ref").fuse_ref(sim3, Fu, fu["kf_has_mp"], fu["xyz"], fu["lvl"], da, fu["state"], fu["facing"], fu["mind"], fu["maxd"], fu["inv_sigma2"],
                                                          th, 40.0, mc.FX, mc.FY, mc.CX, mc.CY)
        out["fuse_%d_%d_nf" % (i, sim3)] = np.array(nf); out["fuse_%d_%d_best" % (i, sim3)] = best; out["fuse_uv"] = uvf; out["fuse_ur"] = urf
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_match_kf.npz"), **out)
print("wrote ref_match_kf.npz", [int(out["kf%d_nm" % i]) for i in range(len(mc.KF_CASES))])
