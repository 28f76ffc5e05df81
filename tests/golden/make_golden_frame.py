#!/usr/bin/env python3
"""Generates tests/golden/frame_cv2.npz (run in the build container, where cv2 4.13 and /root/reference exist):
  * cv2.undistortPoints(pts, K, D, None, K) for the cameras of tests/frame_cases.py  -> pins the oracle's undistortPoints
  * the reference's own Frame bodies (oracle/_ref: UndistortKeyPoints, ComputeImageBounds, ComputeStereoFromRGBD,
    AssignFeaturesToGrid compiled from /root/reference/src/Frame.cc) on the seeded keypoints -> pins the port and the GPU path
usage: python tests/golden/make_golden_frame.py"""
import os, sys
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle                      # noqa: E402
import frame_cases as fc           # noqa: E402

out = {"cv2_version": np.array(cv2.__version__)}
pts, _ = fc.points()
for name, c in fc.CAMS.items():
    K = np.array([[c[0], 0, c[2]], [0, c[1], c[3]], [0, 0, 1]], np.float32); D = np.array(c[4:], np.float32).reshape(-1, 1)
    out["und_" + name] = cv2.undistortPoints(pts.reshape(-1, 1, 2).copy(), K, D, None, K).reshape(-1, 2)
assert oracle.build_ref(), "needs /root/reference"
k = fc.keys(oracle.KP_DTYPE); dimg = fc.depth_image()
for name, c in fc.CAMS.items():
    for tag, d in (("rgbd", dimg), ("mono", None)):
        r = oracle.frame_build("ref", k, c, fc.BF, fc.ROWS, fc.COLS, d)
        for key, v in r.items():
            out["%s_%s_%s" % (name, tag, key)] = v.view(np.uint8) if key == "keys_un" else v
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "frame_cv2.npz"), **out)
print("wrote frame_cv2.npz:", len(out), "arrays")
