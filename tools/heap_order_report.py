#!/usr/bin/env python3
"""SURVEY.md 8c "also report, separately": how far the reference AS SHIPPED (glibc malloc: DistributeOctTree orders equal-size nodes by
heap address, src/ORBextractor.cc:948) is from the canonical parity contract (same sources + monotonic allocator = "newest node first").
Each frame runs in a FRESH PROCESS with ORB_REF_GLIBC_HEAP=1 (oracle/ref/arena.cpp) and is compared with the canonical build.
CPU only; test infrastructure.   python tools/heap_order_report.py [--frames N]"""
import argparse, json, os, subprocess, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r"""
import sys, json, numpy as np
sys.path.insert(0, %r)
import oracle
from tools.synth import synth_frame
seed, w, h, nf = [int(a) for a in sys.argv[1:5]]
k, d = oracle.Extractor("ref", nf, 1.2, 8, 20, 7).extract(synth_frame(seed, w, h))
np.save(sys.stdout.buffer, np.concatenate([k.view(np.uint8).reshape(len(k), 28), d], 1))
"""


def run(seed, w, h, nf, glibc):
    env = dict(os.environ); env["ORB_REF_GLIBC_HEAP"] = "1" if glibc else "0"
    out = subprocess.run([sys.executable, "-c", CHILD % ROOT, str(seed), str(w), str(h), str(nf)], env=env, capture_output=True, check=True).stdout
    import io
    return np.load(io.BytesIO(out))


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=40); a = ap.parse_args()
    rows = []
    for name, w, h, nf in (("C1 640x480/1000", 640, 480, 1000), ("C3 752x480/2000", 752, 480, 2000)):
        tot = dict(frames=0, kp=0, identical_frames=0, same_set_frames=0, kp_not_in_canonical=0, levels=0, levels_identical=0, levels_same_set=0)
        for s in range(a.frames):
            can = run(1000 + s, w, h, nf, False); gl = run(1000 + s, w, h, nf, True)
            tot["frames"] += 1; tot["kp"] += len(gl)
            tot["identical_frames"] += int(can.shape == gl.shape and np.array_equal(can, gl))
            cs = {r.tobytes() for r in can}; gs = {r.tobytes() for r in gl}
            tot["same_set_frames"] += int(cs == gs); tot["kp_not_in_canonical"] += len(gs - cs)
            oc, og = can[:, 20:24].copy().view(np.int32).ravel(), gl[:, 20:24].copy().view(np.int32).ravel()
            for l in range(8):
                a_, b_ = can[oc == l], gl[og == l]
                tot["levels"] += 1; tot["levels_identical"] += int(a_.shape == b_.shape and np.array_equal(a_, b_))
                tot["levels_same_set"] += int({r.tobytes() for r in a_} == {r.tobytes() for r in b_})
        rows.append((name, tot))
    print("| config | frames (one fresh process each) | keypoints (glibc heap) | not in the canonical result | keypoint-set overlap | levels identical (content and order) | levels with the same set, other order | frames identical |")
    print("|---|---|---|---|---|---|---|---|")
    for name, t in rows:
        print("| %s | %d | %d | %d | %.2f %% | %d / %d | %d / %d | %d / %d |" % (name, t["frames"], t["kp"], t["kp_not_in_canonical"], 100.0 * (1 - t["kp_not_in_canonical"] / max(t["kp"], 1)),
              t["levels_identical"], t["levels"], t["levels_same_set"] - t["levels_identical"], t["levels"], t["identical_frames"], t["frames"]))


if __name__ == "__main__":
    main()
