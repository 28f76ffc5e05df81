"""GPU SLIC (orbx_slic_*, amos-slam_b200/csrc/orbx_slic.cu) against the reference's own cluster::SLIC outputs (tests/golden/ref_slic.npz)
and, where oracle/_ref travelled to the box, against the reference build run live on fresh inputs.  Bit-exact: labels and centres."""
import importlib, os, sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import slic_cases as sc

pytestmark = pytest.mark.gpu
orbx = importlib.import_module("amos-slam_b200")
G = np.load(os.path.join(HERE, "golden", "ref_slic.npz"))
FIELDS = ["x", "y", "L", "A", "B", "D", "label"]


def centers_array(c):
    return np.stack([c[f] for f in FIELDS], 1).astype(np.int32)


@pytest.mark.parametrize("case", [c[0] for c in sc.CASES])
def test_slic_matches_reference_golden(case):
    S = orbx.cluster()
    labels, centers = S.SLIC(G[case + "_lab"], G[case + "_depth"])
    assert labels.dtype == np.float64
    assert np.array_equal(labels, G[case + "_labels"].astype(np.float64))
    assert np.array_equal(centers_array(centers), G[case + "_centers"])
    l16, c2 = S.SLIC(G[case + "_lab"], G[case + "_depth"], labels16=True)                 # the form extract_masked_batch takes; handle reused
    assert l16.dtype == np.uint16 and np.array_equal(l16, G[case + "_labels"]) and np.array_equal(centers_array(c2), G[case + "_centers"])


def test_slic_matches_live_reference_on_fresh_inputs():
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref did not travel to this box")
    rng = np.random.default_rng(99)
    S = orbx.cluster()
    for w, h in [(157, 93), (64, 64), (11, 9)]:
        lab = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)                             # any 8-bit 3-channel image is a valid "Lab" input
        lab[: h // 2] = (lab[: h // 2] // 64) * 64                                        # flat blocks: many exact distance ties
        depth = rng.integers(0, 4000, (h, w)).astype(np.uint16)
        want_l, want_c = oracle.slic("ref", lab, depth)
        got_l, got_c = S.SLIC(lab, depth)
        assert np.array_equal(got_l, want_l) and np.array_equal(centers_array(got_c), want_c), (w, h)


def test_slic_feeds_moving_keypoints_labels():
    """labels16 goes straight into the batched MovingKeyPoints path (ids from 1, flag table indexed by id - 1)."""
    S = orbx.cluster()
    l16, centers = S.SLIC(G["qvga_lab"], G["qvga_depth"], labels16=True)
    assert l16.min() >= 1 and l16.max() == len(centers)


def test_slic_argument_errors():
    S = orbx.cluster()
    with pytest.raises(orbx.OrbxError):
        S.SLIC(np.zeros((4, 4, 2), np.uint8), np.zeros((4, 4), np.uint16))
