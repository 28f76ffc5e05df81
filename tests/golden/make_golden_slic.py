#!/usr/bin/env python3
"""Generates tests/golden/ref_slic.npz (run in the build container, where cv2 4.13 and /root/reference exist):
for every case of tests/slic_cases.py the Lab image the REAL OpenCV computes (cv2.cvtColor(bgr, COLOR_BGR2Lab): the input boundary of the
SLIC stage, see oracle/ref/ref_slic_capi.cpp), the depth image, and the outputs of the reference's OWN src/cluster.cc SLIC() compiled into
oracle/_ref: label map (stored as uint16) and centres.  Also cv2's Sobel / addWeighted gradient of one Lab image, which pins cvlite's
restatement of those two primitives, and a canonical-seed k-means result (reference code after explicit seeding)."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))
import cv2
import oracle
import slic_cases as sc

cv2.setNumThreads(1)
g = {}
for name, w, h, seed in sc.CASES:
    bgr = sc.bgr_frame(seed, w, h)
    if name == "flat":
        bgr[:] = 90
    lab = cv2.cvtColor(bgr, cv2.COLOR_BGR2Lab)
    depth = sc.depth_frame(seed, w, h)
    labels, centers = oracle.slic("ref", lab, depth)
    assert labels.max() < 65536 and np.array_equal(labels, np.rint(labels))
    g[name + "_lab"] = lab; g[name + "_depth"] = depth; g[name + "_labels"] = labels.astype(np.uint16); g[name + "_centers"] = centers
    print(name, w, h, "centres", len(centers), "labels used", len(np.unique(labels)), "never covered", int((labels == 0).sum()))
lab = g["odd_lab"]
sx = cv2.Sobel(lab, cv2.CV_64F, 0, 1, ksize=3); sy = cv2.Sobel(lab, cv2.CV_64F, 1, 0, ksize=3)
g["odd_gradient_cv2"] = cv2.addWeighted(sx, 0.5, sy, 0.5, 0)
c = g["qvga_centers"]
valid = np.flatnonzero(c[:, 5] > 0)
seeds = valid[np.random.default_rng(5).choice(len(valid), 4, replace=False)].astype(np.int32)
g["qvga_kmeans_seeds"] = seeds; g["qvga_kmeans_ids"] = oracle.slic_kmeans_ref(c, seeds)
np.savez_compressed(os.path.join(HERE, "ref_slic.npz"), **g)
print("wrote ref_slic.npz", os.path.getsize(os.path.join(HERE, "ref_slic.npz")), "bytes")
