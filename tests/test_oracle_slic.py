"""SLIC stage of `cluster` (src/cluster.cc:88-344): the plain port in gather form, the reference's own code in oracle/_ref and cvlite's
Sobel / addWeighted against the committed fixtures (tests/golden/ref_slic.npz: Lab by the real cv2, outputs by the reference build)."""
import os, sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import oracle
import slic_cases as sc

G = np.load(os.path.join(HERE, "golden", "ref_slic.npz"))


@pytest.mark.parametrize("case", [c[0] for c in sc.CASES])
def test_port_matches_reference_golden(case):
    if case == "qvga":
        pytest.skip("the O(pixels x centres) port takes a minute on 320 x 240; the smaller cases cover the same code")
    labels, centers = oracle.slic("port", G[case + "_lab"], G[case + "_depth"])
    assert np.array_equal(labels.astype(np.uint16), G[case + "_labels"])
    assert np.array_equal(centers, G[case + "_centers"])


@pytest.mark.skipif(not (oracle.have_ref() or os.path.isdir(oracle.REFERENCE_ROOT)), reason="oracle/_ref not built and /root/reference absent")
@pytest.mark.parametrize("case", [c[0] for c in sc.CASES])
def test_reference_build_matches_golden(case):
    labels, centers = oracle.slic("ref", G[case + "_lab"], G[case + "_depth"])
    assert np.array_equal(labels.astype(np.uint16), G[case + "_labels"])
    assert np.array_equal(centers, G[case + "_centers"])


@pytest.mark.skipif(not (oracle.have_ref() or os.path.isdir(oracle.REFERENCE_ROOT)), reason="oracle/_ref not built and /root/reference absent")
def test_cvlite_gradient_matches_cv2():
    assert np.array_equal(oracle.slic_gradient_ref(G["odd_lab"]), G["odd_gradient_cv2"])          # golden written by cv2 4.13
    try:
        import cv2
    except ImportError:
        return
    lab = G["tiny_lab"]
    want = cv2.addWeighted(cv2.Sobel(lab, cv2.CV_64F, 0, 1, ksize=3), 0.5, cv2.Sobel(lab, cv2.CV_64F, 1, 0, ksize=3), 0.5, 0)
    assert np.array_equal(oracle.slic_gradient_ref(lab), want)


@pytest.mark.skipif(not (oracle.have_ref() or os.path.isdir(oracle.REFERENCE_ROOT)), reason="oracle/_ref not built and /root/reference absent")
def test_canonical_seed_kmeans_is_deterministic():
    ids = oracle.slic_kmeans_ref(G["qvga_centers"], G["qvga_kmeans_seeds"])
    assert np.array_equal(ids, G["qvga_kmeans_ids"])
    assert ids.min() >= 0 and ids.max() < len(G["qvga_kmeans_seeds"])


def test_lab_fixture_is_what_cv2_computes():
    cv2 = pytest.importorskip("cv2")
    for name, w, h, seed in sc.CASES:
        bgr = sc.bgr_frame(seed, w, h)
        if name == "flat":
            bgr[:] = 90
        assert np.array_equal(cv2.cvtColor(bgr, cv2.COLOR_BGR2Lab), G[name + "_lab"])
        assert np.array_equal(sc.depth_frame(seed, w, h), G[name + "_depth"])
