#!/usr/bin/env python3
"""Generates tests/golden/ref_bow.npz (run in the build container, where /root/reference exists): outputs of the reference's own
DBoW2 sources (FORB.cpp, BowVector.cpp, FeatureVector.cpp, the two TemplatedVocabulary::transform bodies) and of the two
ORBmatcher::SearchByBoW bodies, compiled into oracle/_ref, on the seeded inputs of tests/bow_cases.py.
The inputs (descriptor pool, vocabulary, frames) are rebuilt from seeds by the tests; the keypoints / descriptors of the two
frames are stored too so that the CPU tests do not depend on an extractor.
usage: python tests/golden/make_golden_bow.py"""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle                      # noqa: E402
import bow_cases as bc             # noqa: E402

assert oracle.build_ref(), "needs /root/reference"
E = oracle.Extractor("ref", 1000, 1.2, 8, 20, 7)
pool, (ka, da), (kb, db) = bc.pool_and_frames(lambda im: E.extract(im))
voc = bc.build_vocabulary(pool, 10, 3)
out = {"pool_sum": np.array(int(pool.astype(np.int64).sum())), "voc_parent": voc[0], "voc_leaf": voc[1], "voc_desc": voc[2], "voc_weight": voc[3],
       "ka": ka.view(np.uint8), "da": da, "kb": kb.view(np.uint8), "db": db}
for wt, sc in bc.VOC_VARIANTS:
    R = oracle.Vocabulary("ref", 10, 3, *voc, weighting=wt, scoring=sc)
    for lu in bc.LEVELSUP:
        r = R.transform(da, lu)
        for key, v in r.items():
            out["t_%d_%d_%d_%s" % (wt, sc, lu, key)] = v
R = oracle.Vocabulary("ref", 10, 3, *voc)
fa, fb = R.transform(da, 1), R.transform(db, 1)
va, vb = bc.validity(len(ka), 1), bc.validity(len(kb), 2)
for kfkf in (0, 1):
    for i, (nn, ori) in enumerate(bc.MATCH_VARIANTS):
        nm, m12, m21 = oracle.search_by_bow("ref", nn, ori, kfkf, ka, da, va, fa, kb, db, vb, fb)
        out["m_%d_%d_nm" % (kfkf, i)] = np.array(nm); out["m_%d_%d_m12" % (kfkf, i)] = m12; out["m_%d_%d_m21" % (kfkf, i)] = m21
# ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:810-1010)
t = bc.tri_inputs(len(ka), len(kb))
for i, (ori, st) in enumerate(bc.TRI_VARIANTS):
    nm, m12, epi = oracle.search_for_triangulation("ref", ori, ka, da, t["free1"], t["ur1"], fa, kb, db, t["free2"], t["ur2"], fb, bc.TRI_F12, bc.TRI_C2, bc.TRI_CAM, t["sf"], t["sigma2"], st)
    out["tri_%d_nm" % i] = np.array(nm); out["tri_%d_m12" % i] = m12; out["tri_epi"] = epi
# MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:359-439): the descriptor the reference body selects for every seeded map point
di = bc.distinctive_inputs(np.concatenate([da, db]))
out["distinctive_chosen"] = oracle.distinctive_descriptors("ref", di["offsets"], di["desc"], di["kf_of"], di["kf_bad"])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_bow.npz"), **out)
print("wrote ref_bow.npz:", len(out), "arrays")
