"""Seeded bag-of-words inputs shared by the golden generator and the tests: a synthetic ORB vocabulary (the real ORBvoc.txt is a
145 MB download the reference does not vendor) built by hierarchical k-medoid-style clustering of real ORB descriptors, with
idf-like weights and a few stopped (weight 0) words."""
import numpy as np


def hamming(a, b):
    return np.unpackbits(a[:, None, :] ^ b[None, :, :], axis=2).sum(2)


def build_vocabulary(pool, k=10, L=3, seed=11):
    """pool: M x 32 uint8 descriptors.  Returns (parent, is_leaf, desc, weight) for nodes 1..n in the file order of ORBvoc.txt
    (a node's children are listed after it; ids grow in that order)."""
    rng = np.random.default_rng(seed)
    parent, leaf, desc, weight = [], [], [], []

    def grow(pid, members, level):
        # k centres drawn from the members (random descriptors when the cluster ran dry), members assigned to the nearest
        if len(members) >= k:
            centres = pool[rng.choice(members, k, replace=False)]
        else:
            centres = rng.integers(0, 256, (k, 32), dtype=np.uint8)
        assign = hamming(pool[members], centres).argmin(1) if len(members) else np.zeros(0, np.int64)
        ids = []
        for c in range(k):
            parent.append(pid); desc.append(centres[c]); leaf.append(1 if level == L else 0)
            w = 0.0 if (level == L and rng.random() < 0.03) else float(np.log(1.0 + rng.uniform(1.0, 400.0)))
            weight.append(w if level == L else 0.0)
            ids.append(len(parent))
        if level < L:
            for c in range(k):
                grow(ids[c], members[assign == c] if len(members) else members, level + 1)

    # breadth is not required by the loader: depth-first file order keeps every parent before its children
    grow(0, np.arange(len(pool)), 1)
    return np.array(parent, np.int32), np.array(leaf, np.uint8), np.array(desc, np.uint8), np.array(weight, np.float64)


def pool_and_frames(extract, n_pool_frames=6):
    """extract: img -> (keypoints, descriptors).  Returns the descriptor pool and two related frames (A, B = A warped)."""
    from tools.synth import synth_frame, warp_affine_nn
    pool = np.concatenate([extract(synth_frame(500 + s, 640, 480))[1] for s in range(n_pool_frames)])
    A = synth_frame(0, 640, 480); B = warp_affine_nn(A, 7, -4, 2.0)
    ka, da = extract(A); kb, db = extract(B)
    return pool, (ka, da), (kb, db)


def validity(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.random(n) < 0.8).astype(np.uint8)


VOC_VARIANTS = [(0, 0), (1, 1), (2, 5), (3, 0), (0, 5)]      # (weighting, scoring): TF_IDF/L1 (ORBvoc), TF/L2, IDF/DOT, BINARY/L1, TF_IDF/DOT
LEVELSUP = [4, 2, 1, 0]                                      # L = 3: root, level 1, level 2 (what levelsup = 4 gives with L = 6), words
MATCH_VARIANTS = [(0.7, True), (0.75, False), (0.9, True)]   # TrackReferenceKeyFrame 0.7 / relocalisation 0.75 / loop closing 0.75 (Tracking.cc, LoopClosing.cc)


def write_text(path, k, L, scoring, weighting, voc):
    """ORBvoc.txt format (TemplatedVocabulary.h:1336-1424)."""
    parent, leaf, desc, weight = voc
    with open(path, "w") as f:
        f.write("%d %d %d %d\n" % (k, L, scoring, weighting))
        for i in range(len(parent)):
            f.write("%d %d %s %r\n" % (parent[i], leaf[i], " ".join(str(int(b)) for b in desc[i]), float(weight[i])))


# SearchForTriangulation inputs: frame B = frame A moved by (+7, -4) px and rotated by 2 degrees, so a fundamental matrix of a pure image
# translation along (7, -4) puts the true matches within a few pixels of their epipolar lines near the image centre and further away
# towards the borders; the epipole used for the "too close to the epipole" test is an independent input.
TRI_F12 = [[0.0, 0.0, -0.004], [0.0, 0.0, -0.007], [0.004, 0.007, 0.0]]
TRI_C2 = (0.02, 0.01, 0.5)                                   # camera 1's centre in camera 2 -> epipole (339.3, 265.6) with the TUM1-like intrinsics below
TRI_CAM = (517.3, 516.5, 318.6, 255.3)
TRI_VARIANTS = [(True, False), (False, False), (True, True)] # (checkOri, bOnlyStereo)


def tri_inputs(n1, n2, seed=31):
    rng = np.random.default_rng(seed)
    free1 = (rng.random(n1) < 0.7).astype(np.uint8); free2 = (rng.random(n2) < 0.7).astype(np.uint8)
    ur1 = np.where(rng.random(n1) < 0.4, rng.uniform(0, 600, n1), -1).astype(np.float32); ur2 = np.where(rng.random(n2) < 0.4, rng.uniform(0, 600, n2), -1).astype(np.float32)
    sf = (np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32)
    return dict(free1=free1, free2=free2, ur1=ur1, ur2=ur2, sf=sf, sigma2=(sf * sf).astype(np.float32))


def distinctive_inputs(pool, n_points=300, n_kf=60, seed=41):
    """Map points with 0..40 observations: noisy copies of one pool descriptor each (plus a few outliers), spread over n_kf KeyFrames of which some
    are bad.  Returns the observation slots (offsets, descriptors, KeyFrame of every slot, bad flags) and the filtered lists a C-ABI caller passes."""
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, 41, n_points); counts[:6] = [0, 1, 2, 3, 40, 33]
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    desc = np.zeros((offsets[-1], 32), np.uint8); kf_of = np.zeros(offsets[-1], np.int32)
    for p in range(n_points):
        base = pool[rng.integers(0, len(pool))]
        for s in range(offsets[p], offsets[p + 1]):
            d = base.copy() if rng.random() > 0.1 else pool[rng.integers(0, len(pool))].copy()
            flips = rng.integers(0, 256, rng.integers(0, 30))
            for b in flips:
                d[b >> 3] ^= np.uint8(1 << (b & 7))
            desc[s] = d
        kf_of[offsets[p]:offsets[p + 1]] = np.sort(rng.choice(n_kf, counts[p], replace=False))
    kf_bad = (rng.random(n_kf) < 0.1).astype(np.uint8)
    keep = kf_bad[kf_of] == 0
    f_counts = np.array([keep[offsets[p]:offsets[p + 1]].sum() for p in range(n_points)])
    f_offsets = np.concatenate([[0], np.cumsum(f_counts)]).astype(np.int32)
    return dict(offsets=offsets, desc=desc, kf_of=kf_of, kf_bad=kf_bad, f_offsets=f_offsets, f_desc=desc[keep])
