#!/usr/bin/env python3
"""Generates the committed golden fixtures (run in the build container, where cv2 4.13 and /root/reference exist):

  cvlite_cv2.npz   : inputs + outputs of the real OpenCV (cv2 4.13.0) for the primitives the reference calls
                     (resize INTER_LINEAR, GaussianBlur 7x7 s2, FAST 20/7 with NMS, fastAtan2, ellipse 31x31,
                     MORPH close, copyMakeBorder REFLECT_101).  Pins oracle/cvlite against the third-party library.
  ref_extract.npz  : outputs of the reference's OWN ORBextractor.cc (oracle/_ref, monotonic allocator) on seeded
                     synthetic frames: keypoints + descriptors for C1-like and odd geometries; quadtree stage
                     outputs on adversarial candidate sets; MovingKeyPoints / ProcessDesp outputs.
The fixtures travel to the GPU box; nothing under /root/reference is read at test time.
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cv2
import oracle
from tools.synth import synth_frame, synth_mask

cv2.setNumThreads(1)
rng = np.random.default_rng(12345)

# ---------------- cv2 primitives ----------------
g = {}
img = rng.integers(0, 256, (97, 133), dtype=np.uint8)
g["img_small"] = img
g["resize_small"] = cv2.resize(img, (int(round(133 / 1.2)), int(round(97 / 1.2))), interpolation=cv2.INTER_LINEAR)
g["blur_small"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
g["border_small"] = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
f = synth_frame(3, 200, 150)
g["img_synth"] = f
g["resize_synth"] = cv2.resize(f, (167, 125), interpolation=cv2.INTER_LINEAR)
g["blur_synth"] = cv2.GaussianBlur(f, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
for th in (20, 7):
    k = cv2.FastFeatureDetector_create(th, True).detect(f)
    g["fast%d" % th] = np.array([(p.pt[0], p.pt[1], p.response) for p in k], np.float32).reshape(-1, 3)
    k = cv2.FastFeatureDetector_create(th, False).detect(f)
    g["fast%d_nonms" % th] = np.array([(p.pt[0], p.pt[1], p.response) for p in k], np.float32).reshape(-1, 3)
y = np.rint(rng.normal(0, 300000, 4000)).astype(np.float32); x = np.rint(rng.normal(0, 300000, 4000)).astype(np.float32)
y[:10] = 0; x[:5] = 0; x[5:10] = [1, -1, 5, -7, 0]
g["atan_y"], g["atan_x"] = y, x
g["atan_out"] = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
k31 = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (31, 31), (15, 15))
g["ellipse31"] = k31
m = synth_mask(1, 160, 120); m[:6, :9] = 255; m[50:53, 70:72] = 0
g["mask"] = m
g["mask_close"] = cv2.erode(cv2.dilate(m, k31), k31)
# cv::gemm on small 32F matrices (the pose products of the matcher bodies, e.g. Rcw * x3Dw + tcw at src/ORBmatcher.cc:1608):
# 3x3 * 3x1 (+ 3x1), 3x3 * 3x3, 4x4 * 4x4 (hand-unrolled float path), 1x3 * 3x1 and 5x5 * 5x1 (general path, double accumulation)
def gemm_cases(seed=77, n=400):
    r = np.random.default_rng(seed); out = {}
    for name, (ar, ac, bc, withc) in {"3x3_3x1_c": (3, 3, 1, True), "3x3_3x1": (3, 3, 1, False), "3x3_3x3": (3, 3, 3, False), "4x4_4x4": (4, 4, 4, False),
                                       "1x3_3x1": (1, 3, 1, False), "5x5_5x1": (5, 5, 1, False), "2x2_2x1_c": (2, 2, 1, True)}.items():
        A = r.normal(0, 1, (n, ar, ac)).astype(np.float32); B = r.normal(0, 3, (n, ac, bc)).astype(np.float32); Cm = r.normal(0, 1, (n, ar, bc)).astype(np.float32)
        D = np.stack([cv2.gemm(A[i], B[i], 1.0, Cm[i] if withc else None, 1.0 if withc else 0.0) for i in range(n)]).astype(np.float32).reshape(n, ar, bc)
        out["gemm_%s_A" % name] = A; out["gemm_%s_B" % name] = B; out["gemm_%s_D" % name] = D
        if withc: out["gemm_%s_C" % name] = Cm
    return out
g.update(gemm_cases())
np.savez_compressed(os.path.join(HERE, "cvlite_cv2.npz"), **g)

# ---------------- reference extractor ----------------
assert oracle.have_ref(), "build oracle/_ref first (make -C oracle/ref)"
r = {}
cases = {"c1": (640, 480, 1000, 1.2, 8, 20, 7, 0), "odd": (323, 251, 300, 1.2, 8, 20, 7, 5), "wide": (620, 188, 600, 1.2, 8, 20, 7, 9),
         "lv4": (400, 300, 500, 1.5, 4, 25, 9, 11)}
for name, (w, h, nf, sc, nl, it, mt, seed) in cases.items():
    E = oracle.Extractor("ref", nf, sc, nl, it, mt)
    kp, desc = E.extract(synth_frame(seed, w, h))
    r[name + "_params"] = np.array([w, h, nf, nl, it, mt, seed], np.int64); r[name + "_scale"] = np.float32(sc)
    r[name + "_kp"] = kp; r[name + "_desc"] = desc
# quadtree on adversarial candidate sets
E = oracle.Extractor("ref", 1000, 1.2, 8, 20, 7)
def cands(xs, ys, rs):
    c = np.zeros(len(xs), oracle.KP_DTYPE); c["x"] = xs; c["y"] = ys; c["response"] = rs; c["size"] = 7; c["angle"] = -1; c["class_id"] = -1
    return c
W, H = 608, 448
sets = {}
xs = rng.integers(3, W - 3, 3000); ys = rng.integers(3, H - 3, 3000)
u = np.unique(np.stack([ys, xs], 1), axis=0); sets["uniform"] = cands(u[:, 1], u[:, 0], rng.integers(7, 255, len(u)))
xs = np.clip(rng.normal(100, 12, 2500), 3, W - 4).astype(int); ys = np.clip(rng.normal(90, 10, 2500), 3, H - 4).astype(int)
u = np.unique(np.stack([ys, xs], 1), axis=0); u = u[rng.permutation(len(u))]; sets["cluster"] = cands(u[:, 1], u[:, 0], np.full(len(u), 30))
yy, xx = np.mgrid[3:H - 3:7, 3:W - 3:7]; sets["grid_ties"] = cands(xx.ravel(), yy.ravel(), np.full(xx.size, 50))
sets["few"] = cands([10, 300, 301, 600], [10, 200, 200, 440], [20, 30, 30, 9])
sets["one"] = cands([55], [66], [99])
for name, c in sets.items():
    for N in (217, 30, 5):
        out = E.distribute(c, 16, 16 + W, 16, 16 + H, N)
        r["oct_%s_%d_in" % (name, N)] = c; r["oct_%s_%d_out" % (name, N)] = out
# Amos path: detect -> MovingKeyPoints -> ProcessDesp
E = oracle.Extractor("ref", 800, 1.2, 8, 20, 7)
f = synth_frame(21, 480, 360)
kp, counts = E.detect(f)
mask = synth_mask(2, 480, 360)
label = (np.arange(360)[:, None] // 24 * 20 + np.arange(480)[None, :] // 24 + 1).astype(np.float64)
centers_id = (np.arange(int(label.max())) * 7 % 15).astype(np.int32)
rm = np.zeros(15, np.int32); rm[[2, 9]] = 1
kp2, counts2, culled = E.moving_keypoints(mask, label, centers_id, rm, kp, counts)
kp3, desc3 = E.process_desp(kp2, counts2)
r.update(amos_frame_seed=np.int64(21), amos_mask=mask, amos_label=label, amos_centers_id=centers_id, amos_rm=rm, amos_detect_kp=kp, amos_detect_counts=counts,
         amos_kept_kp=kp2, amos_kept_counts=counts2, amos_culled=culled, amos_final_kp=kp3, amos_final_desc=desc3)
np.savez_compressed(os.path.join(HERE, "ref_extract.npz"), **r)
print("golden written:", {k: os.path.getsize(os.path.join(HERE, k)) for k in ("cvlite_cv2.npz", "ref_extract.npz")})
