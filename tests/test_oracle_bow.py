"""CPU: the oracle's bag-of-words path (DBoW2 transform, ORBmatcher::SearchByBoW x2) against the committed outputs of the
reference's own sources (tests/golden/make_golden_bow.py), and port == ref live where oracle/_ref exists."""
import os
import numpy as np
import pytest
import bow_cases as bc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_bow.npz"))
TKEYS = ("word", "node", "bow_ids", "bow_vals", "fv_nodes", "fv_offsets", "fv_idx")
VOC = (G["voc_parent"], G["voc_leaf"], G["voc_desc"], G["voc_weight"])


def frames(oracle):
    return G["ka"].view(oracle.KP_DTYPE), G["da"], G["kb"].view(oracle.KP_DTYPE), G["db"]


@pytest.mark.parametrize("wt,sc", bc.VOC_VARIANTS)
def test_port_transform_matches_reference_dbow2(oracle, wt, sc):
    P = oracle.Vocabulary("port", 10, 3, *VOC, weighting=wt, scoring=sc)
    for lu in bc.LEVELSUP:
        r = P.transform(G["da"], lu)
        for key in TKEYS:
            assert np.array_equal(r[key], G["t_%d_%d_%d_%s" % (wt, sc, lu, key)]), (wt, sc, lu, key)


def test_port_search_by_bow_matches_reference_bodies(oracle):
    ka, da, kb, db = frames(oracle)
    P = oracle.Vocabulary("port", 10, 3, *VOC)
    fa, fb = P.transform(da, 1), P.transform(db, 1)
    va, vb = bc.validity(len(ka), 1), bc.validity(len(kb), 2)
    for kfkf in (0, 1):
        for i, (nn, ori) in enumerate(bc.MATCH_VARIANTS):
            nm, m12, m21 = oracle.search_by_bow("port", nn, ori, kfkf, ka, da, va, fa, kb, db, vb, fb)
            assert nm == int(G["m_%d_%d_nm" % (kfkf, i)]) and nm > 100
            assert np.array_equal(m12, G["m_%d_%d_m12" % (kfkf, i)]) and np.array_equal(m21, G["m_%d_%d_m21" % (kfkf, i)])


def test_reference_reproduces_golden_and_vocabulary_is_rebuilt_from_seed(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    R = oracle.Vocabulary("ref", 10, 3, *VOC)
    r = R.transform(G["da"], 4)
    for key in TKEYS:
        assert np.array_equal(r[key], G["t_0_0_4_%s" % key])


def test_golden_properties():
    """Size-independent properties of the reference outputs: BowVector is L1-normalised and sorted, every non-stopped feature is in
    exactly one FeatureVector node, the node of a feature is an ancestor of its word, matches are one-to-one."""
    bv = G["t_0_0_4_bow_vals"]; bi = G["t_0_0_4_bow_ids"]
    assert abs(bv.sum() - 1.0) < 1e-12 and np.all(np.diff(bi) > 0) and np.all(bv > 0)
    assert abs(np.sqrt((G["t_1_1_4_bow_vals"] ** 2).sum()) - 1.0) < 1e-12
    parent = np.concatenate([[0], G["voc_parent"]]); leaf_ids = np.nonzero(G["voc_leaf"])[0] + 1
    word, node = G["t_0_0_1_word"], G["t_0_0_1_node"]
    assert np.array_equal(parent[leaf_ids[word]], node)                                  # L = 3, levelsup = 1: the node is the word's parent
    stopped = G["voc_weight"][leaf_ids[word] - 1] <= 0
    assert stopped.any() and len(G["t_0_0_1_fv_idx"]) == int((~stopped).sum()) and len(set(G["t_0_0_1_fv_idx"].tolist())) == int((~stopped).sum())
    for kfkf in (0, 1):
        m12, m21 = G["m_%d_0_m12" % kfkf], G["m_%d_0_m21" % kfkf]
        i = np.nonzero(m12 >= 0)[0]
        assert np.array_equal(m21[m12[i]], i) and int(G["m_%d_0_nm" % kfkf]) == len(i)


def test_port_search_for_triangulation_matches_reference_body(oracle):
    """ORBmatcher::SearchForTriangulation + CheckDistEpipolarLine (ORBmatcher.cc:810-1010, 188-215): port vs the reference bodies' committed outputs."""
    ka, da, kb, db = frames(oracle)
    P = oracle.Vocabulary("port", 10, 3, *VOC)
    fa, fb = P.transform(da, 1), P.transform(db, 1)
    t = bc.tri_inputs(len(ka), len(kb))
    for i, (ori, st) in enumerate(bc.TRI_VARIANTS):
        nm, m12 = oracle.search_for_triangulation("port", ori, ka, da, t["free1"], t["ur1"], fa, kb, db, t["free2"], t["ur2"], fb, bc.TRI_F12, G["tri_epi"], bc.TRI_CAM, t["sf"], t["sigma2"], st)
        assert nm == int(G["tri_%d_nm" % i]) and nm > 10 and np.array_equal(m12, G["tri_%d_m12" % i])


def check_distinctive(best, di):
    ch = G["distinctive_chosen"]
    assert len(best) == len(ch) and (best < 0).sum() >= 1
    for p in range(len(best)):
        if best[p] < 0:
            assert not ch[p].any() and di["f_offsets"][p + 1] == di["f_offsets"][p]
        else:
            assert np.array_equal(ch[p], di["f_desc"][di["f_offsets"][p] + best[p]]), p


def test_port_distinctive_descriptors_match_reference_body(oracle):
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:359-439): port vs the descriptors the reference body selected."""
    di = bc.distinctive_inputs(np.concatenate([G["da"], G["db"]]))
    check_distinctive(oracle.distinctive_descriptors("port", di["f_offsets"], di["f_desc"]), di)
