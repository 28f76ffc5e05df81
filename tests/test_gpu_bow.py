"""GPU (B200): bag of words through the C ABI: DBoW2 transform (word / node per feature, BowVector with its
double-precision weights, FeatureVector) and both ORBmatcher::SearchByBoW forms, bit-exact against the outputs of the reference's own
sources (tests/golden/ref_bow.npz) and the port oracle."""
import os
import numpy as np
import pytest
import bow_cases as bc

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_bow.npz"))
TKEYS = ("word", "node", "bow_ids", "bow_vals", "fv_nodes", "fv_offsets", "fv_idx")
VOC = (G["voc_parent"], G["voc_leaf"], G["voc_desc"], G["voc_weight"])


@pytest.mark.parametrize("wt,sc", bc.VOC_VARIANTS)
def test_transform_matches_reference_dbow2(orbx, wt, sc):
    V = orbx.ORBVocabulary(10, 3, *VOC, weighting=wt, scoring=sc)
    assert V.size() == int(G["voc_leaf"].sum())
    for lu in bc.LEVELSUP:
        r = V.transform(G["da"], lu)
        for key in TKEYS:
            assert np.array_equal(r[key], G["t_%d_%d_%d_%s" % (wt, sc, lu, key)]), (wt, sc, lu, key)


def test_extracted_frame_gives_the_golden_descriptors(orbx):
    """The goldens were produced from the reference extractor's descriptors: the GPU extractor must give the same ones."""
    from tools.synth import synth_frame
    k, d = orbx.ORBextractor(1000, 1.2, 8, 20, 7)(synth_frame(0, 640, 480))
    assert np.array_equal(d, G["da"]) and np.array_equal(k.view(np.uint8), G["ka"])


def test_search_by_bow_matches_reference_bodies(orbx):
    ka, da, kb, db = G["ka"].view(orbx.KP_DTYPE), G["da"], G["kb"].view(orbx.KP_DTYPE), G["db"]
    V = orbx.ORBVocabulary(10, 3, *VOC)
    fa, fb = V.transform(da, 1), V.transform(db, 1)
    va, vb = bc.validity(len(ka), 1), bc.validity(len(kb), 2)
    for kfkf in (0, 1):
        for i, (nn, ori) in enumerate(bc.MATCH_VARIANTS):
            nm, m12, m21 = orbx.ORBmatcher(nn, ori).SearchByBoW(kfkf, ka, da, va, fa, kb, db, vb if kfkf else None, fb)
            assert nm == int(G["m_%d_%d_nm" % (kfkf, i)])
            assert np.array_equal(m12, G["m_%d_%d_m12" % (kfkf, i)]) and np.array_equal(m21, G["m_%d_%d_m21" % (kfkf, i)])


def test_bow_variants_vs_port(orbx, oracle):
    """Other shapes: every feature in one node (levelsup >= L: a single warp replays 1000 x 1000), deep tree, few / no valid features, empty sides."""
    ka, da, kb, db = G["ka"].view(orbx.KP_DTYPE), G["da"], G["kb"].view(orbx.KP_DTYPE), G["db"]
    V = orbx.ORBVocabulary(10, 3, *VOC); P = oracle.Vocabulary("port", 10, 3, *VOC)
    rng = np.random.default_rng(5)
    for lu, pv in ((4, 0.8), (2, 0.3), (0, 1.0), (1, 0.0)):
        fa, fb = V.transform(da, lu), V.transform(db, lu)
        va = (rng.random(len(ka)) < pv).astype(np.uint8); vb = (rng.random(len(kb)) < 0.9).astype(np.uint8)
        for kfkf in (0, 1):
            g = orbx.ORBmatcher(0.75, True).SearchByBoW(kfkf, ka, da, va, fa, kb, db, vb if kfkf else None, fb)
            o = oracle.search_by_bow("port", 0.75, True, kfkf, ka, da, va, fa, kb, db, vb, fb)
            assert g[0] == o[0] and np.array_equal(g[1], o[1]) and np.array_equal(g[2], o[2]), (lu, pv, kfkf)
    # deeper, narrower tree and a different descriptor set
    pool = np.concatenate([da, db])
    voc2 = bc.build_vocabulary(pool, k=4, L=5, seed=3)
    V2 = orbx.ORBVocabulary(4, 5, *voc2); P2 = oracle.Vocabulary("port", 4, 5, *voc2)
    for lu in (4, 3, 1):
        g, o = V2.transform(db, lu), P2.transform(db, lu)
        for key in TKEYS:
            assert np.array_equal(g[key], o[key]), (lu, key)
    # empty inputs
    e = V.transform(np.zeros((0, 32), np.uint8), 4)
    assert len(e["bow_ids"]) == 0 and len(e["fv_nodes"]) == 0
    fa = V.transform(da, 1)
    nm, m12, m21 = orbx.ORBmatcher(0.7, True).SearchByBoW(0, ka, da, np.ones(len(ka), np.uint8), fa, kb[:0], db[:0], None, e)
    assert nm == 0 and np.all(m12 == -1) and len(m21) == 0


def test_vocabulary_text_file_loader(orbx, tmp_path):
    path = str(tmp_path / "voc.txt")
    bc.write_text(path, 10, 3, 1, 1, VOC)                                # scoring L2_NORM, weighting TF
    V = orbx.ORBVocabulary.load_text(path)
    r = V.transform(G["da"], 2)
    for key in TKEYS:
        assert np.array_equal(r[key], G["t_1_1_2_%s" % key]), key
    with pytest.raises(orbx.OrbxError):
        orbx.ORBVocabulary.load_text(str(tmp_path / "missing.txt"))
    with pytest.raises(orbx.OrbxError):
        orbx.ORBVocabulary(10, 3, VOC[0], 1 - VOC[1], VOC[2], VOC[3])    # leaf flags contradict the tree


def test_search_for_triangulation_matches_reference_body(orbx, oracle):
    """SearchForTriangulation (LocalMapping::CreateNewMapPoints): node-by-node replay with the epipole and epipolar-line gates, against the reference's golden."""
    ka, da, kb, db = G["ka"].view(orbx.KP_DTYPE), G["da"], G["kb"].view(orbx.KP_DTYPE), G["db"]
    V = orbx.ORBVocabulary(10, 3, *VOC)
    fa, fb = V.transform(da, 1), V.transform(db, 1)
    t = bc.tri_inputs(len(ka), len(kb))
    for i, (ori, st) in enumerate(bc.TRI_VARIANTS):
        nm, m12 = orbx.ORBmatcher(0.6, ori).SearchForTriangulation(ka, da, t["free1"], t["ur1"], fa, kb, db, t["free2"], t["ur2"], fb, bc.TRI_F12, G["tri_epi"], t["sf"], t["sigma2"], st)
        assert nm == int(G["tri_%d_nm" % i]) and np.array_equal(m12, G["tri_%d_m12" % i])
    # other node granularities (one node holding everything; one node per word) and a different geometry, against the port
    rng = np.random.default_rng(8)
    F2 = rng.normal(0, 0.01, (3, 3)).astype(np.float32)
    for lu in (4, 0):
        fa, fb = V.transform(da, lu), V.transform(db, lu)
        g = orbx.ORBmatcher(0.6, True).SearchForTriangulation(ka, da, t["free1"], t["ur1"], fa, kb, db, t["free2"], t["ur2"], fb, F2, (100.0, 400.0), t["sf"], t["sigma2"], False)
        o = oracle.search_for_triangulation("port", True, ka, da, t["free1"], t["ur1"], fa, kb, db, t["free2"], t["ur2"], fb, F2, (100.0, 400.0), bc.TRI_CAM, t["sf"], t["sigma2"], False)
        assert g[0] == o[0] and np.array_equal(g[1], o[1]), lu


def test_distinctive_descriptors_match_reference_body(orbx, oracle):
    """MapPoint::ComputeDistinctiveDescriptors batched over 300 map points (0..40 observations each) against the reference body's selection."""
    from test_oracle_bow import check_distinctive
    di = bc.distinctive_inputs(np.concatenate([G["da"], G["db"]]))
    best = orbx.ORBmatcher().ComputeDistinctiveDescriptors(di["f_offsets"], di["f_desc"])
    check_distinctive(best, di)
    assert np.array_equal(best, oracle.distinctive_descriptors("port", di["f_offsets"], di["f_desc"]))
    # one point with many observations (more than one histogram pass per lane) and an empty call
    big = np.concatenate([G["da"][:700], G["db"][:300]])
    off = np.array([0, 1000], np.int32)
    assert np.array_equal(orbx.ORBmatcher().ComputeDistinctiveDescriptors(off, big), oracle.distinctive_descriptors("port", off, big))
    assert len(orbx.ORBmatcher().ComputeDistinctiveDescriptors(np.array([0], np.int32), np.zeros((0, 32), np.uint8))) == 0


def test_large_descriptor_sets_take_the_global_memory_paths(orbx, oracle):
    """More than 4096 descriptors: the BowVector / FeatureVector keys no longer sort in shared memory; 2000-feature frames as in C3 / C4."""
    rng = np.random.default_rng(12)
    base = np.concatenate([G["da"], G["db"]])
    big = base[rng.integers(0, len(base), 6000)].copy()
    flip = rng.integers(0, 256, (6000, 3))
    for c in range(3):
        big[np.arange(6000), flip[:, c] >> 3] ^= (1 << (flip[:, c] & 7)).astype(np.uint8)
    V = orbx.ORBVocabulary(10, 3, *VOC); P = oracle.Vocabulary("port", 10, 3, *VOC)
    for n in (4096, 4097, 6000):
        g, o = V.transform(big[:n], 2), P.transform(big[:n], 2)
        for key in TKEYS:
            assert np.array_equal(g[key], o[key]), (n, key)
    # SearchByBoW on 2000 x 2000 features
    k1 = np.zeros(2000, orbx.KP_DTYPE); k2 = np.zeros(2000, orbx.KP_DTYPE)
    k1["angle"] = rng.uniform(0, 360, 2000); k2["angle"] = (k1["angle"] + rng.normal(0, 3, 2000)) % 360
    d1, d2 = big[:2000], big[2000:4000].copy(); d2[:1500] = d1[:1500]; d2[np.arange(1500), 5] ^= 1
    f1, f2 = V.transform(d1, 1), V.transform(d2, 1)
    v1 = (rng.random(2000) < 0.9).astype(np.uint8); v2 = (rng.random(2000) < 0.9).astype(np.uint8)
    for kfkf in (0, 1):
        g = orbx.ORBmatcher(0.7, True).SearchByBoW(kfkf, k1, d1, v1, f1, k2, d2, v2 if kfkf else None, f2)
        o = oracle.search_by_bow("port", 0.7, True, kfkf, k1, d1, v1, f1, k2, d2, v2, f2)
        assert g[0] == o[0] and g[0] > 500 and np.array_equal(g[1], o[1]) and np.array_equal(g[2], o[2])
