"""GPU (B200): orbx_pool, the multi-sequence / multi-GPU driver of the extractor, through the C ABI.
Sequence affinity (gpu = seq mod G, stream = (seq div G) mod S), per-worker ordering, and results identical to direct calls."""
import numpy as np
import pytest
import torch
from tools.synth import synth_frame, synth_mask, synth_labels

pytestmark = pytest.mark.gpu


def kp_equal(a, b):
    return len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in a.dtype.names)


def test_pool_single_frames_and_affinity(orbx):
    G = min(torch.cuda.device_count(), 2); S = 2
    P = orbx.ExtractorPool(500, 1.2, 8, 20, 7, n_gpus=G, streams_per_gpu=S)
    E = orbx.ORBextractor(500, 1.2, 8, 20, 7)
    cap = E.max_keypoints(240, 320)
    nseq, per = 6, 3
    frames = {(s, t): synth_frame(100 * s + t, 320, 240) for s in range(nseq) for t in range(per)}
    tickets = {k: P.submit(k[0], f, cap) for k, f in frames.items()}
    for k, t in tickets.items():
        kp, desc = P.result(t)
        k0, d0 = E(frames[k])
        assert kp_equal(kp, k0) and np.array_equal(desc, d0), k
    # every sequence was served by exactly the worker the rule names, and only that one
    expect = np.zeros((G, S), np.int64)
    for s in range(nseq):
        g, st = s % G, (s // G) % S
        assert P.device_of(s) == g
        expect[g, st] += per
    got = np.array([[P.frames_done(g, st) for st in range(S)] for g in range(G)])
    assert np.array_equal(got, expect), (got, expect)
    P.close()


def test_pool_masked_batches_with_labels(orbx):
    """C5 as the pool runs it: one batch job per sequence chunk (images + masks + super-pixel labels), all sequences in flight at once."""
    G = min(torch.cuda.device_count(), 2)
    P = orbx.ExtractorPool(800, 1.2, 8, 20, 7, n_gpus=G, streams_per_gpu=2)
    E = orbx.ORBextractor(800, 1.2, 8, 20, 7)
    w, h, B = 480, 360, 3
    jobs = {}
    for s in range(5):
        imgs = np.stack([synth_frame(300 + 10 * s + b, w, h) for b in range(B)]); masks = np.stack([synth_mask(40 + 10 * s + b, w, h) for b in range(B)])
        labs = [synth_labels(10 * s + b, w, h) for b in range(B)]
        labels = np.stack([l[0] for l in labs]); flagged = np.stack([orbx.label_flags(l[1], l[2]) for l in labs])
        use_labels = s % 2 == 0
        jobs[s] = (imgs, masks, labels if use_labels else None, flagged if use_labels else None,
                   P.submit_batch(s, imgs, masks, labels if use_labels else None, flagged if use_labels else None, cap=E.max_keypoints(h, w)))
    for s, (imgs, masks, labels, flagged, t) in jobs.items():
        kp, desc, counts, culled = P.result(t)
        k0, d0, c0, u0 = E.extract_masked_batch(imgs, masks, labels=labels, flagged=flagged)
        assert np.array_equal(counts, c0) and np.array_equal(culled, u0), s
        for b in range(B):
            assert kp_equal(kp[b][:counts[b]], k0[b][:c0[b]]) and np.array_equal(desc[b][:counts[b]], d0[b][:c0[b]]), (s, b)
    # plain batches (no mask) and an error path: a bad job is reported by wait, later jobs still run
    imgs = np.stack([synth_frame(900 + b, w, h) for b in range(B)])
    t_ok = P.submit_batch(1, imgs, cap=E.max_keypoints(h, w))
    t_small = P.submit_batch(1, imgs, cap=8)                                      # capacity far too small -> ORBX_E_CAPACITY from the job
    t_ok2 = P.submit_batch(1, imgs, cap=E.max_keypoints(h, w))
    kp, desc, counts, _ = P.result(t_ok)
    k0, d0, c0 = E.extract_batch(imgs)
    assert np.array_equal(counts, c0) and all(kp_equal(kp[b][:c0[b]], k0[b][:c0[b]]) for b in range(B))
    with pytest.raises(orbx.OrbxError) as e:
        P.result(t_small)
    assert e.value.code == orbx.E_CAPACITY
    assert np.array_equal(P.result(t_ok2)[2], c0)
    P.wait_all()
    P.close()
