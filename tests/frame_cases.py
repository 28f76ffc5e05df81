"""Shared inputs of the device-Frame tests (CPU oracle tests and GPU parity tests use the same seeded cases)."""
import numpy as np

# (fx, fy, cx, cy, k1, k2, p1, p2[, k3]) as float32, from the reference's example settings
CAMS = {
    "tum1": (517.306408, 516.469215, 318.643040, 255.313989, 0.262383, -0.953104, -0.005358, 0.002628, 1.163314),   # Examples/RGB-D/TUM1.yaml
    "tum2": (520.908620, 521.007327, 325.141442, 249.701764, 0.231222, -0.784899, -0.003257, -0.000105, 0.917205),  # Examples/RGB-D/TUM2.yaml
    "tum3": (535.4, 539.2, 320.1, 247.6, 0.0, 0.0, 0.0, 0.0, 0.0),                                                  # Examples/RGB-D/TUM3.yaml (rectified)
    "four": (458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05),              # EuRoC-style, 4 coefficients
}
BF = 40.0            # Camera.bf of the TUM settings
ROWS, COLS = 480, 640


def points(seed=5, n=4000):
    """Keypoint-like coordinates: level-l integer positions scaled by 1.2^l as the extractor produces them, plus the image corners."""
    rng = np.random.default_rng(seed)
    lvl = rng.integers(0, 8, n)
    sc = np.float32(1.2) ** lvl.astype(np.float32)
    x = (rng.integers(19, 620, n) / sc).astype(np.int32).astype(np.float32) * sc.astype(np.float32)
    y = (rng.integers(19, 460, n) / sc).astype(np.int32).astype(np.float32) * sc.astype(np.float32)
    pts = np.stack([np.minimum(x, 639), np.minimum(y, 479)], 1).astype(np.float32)
    pts[:4] = [[0, 0], [COLS, 0], [0, ROWS], [COLS, ROWS]]
    return pts, lvl.astype(np.int32)


def keys(kp_dtype, seed=5, n=3000):
    pts, lvl = points(seed, n)
    pts[:4] = [[19, 19], [620, 19], [19, 460], [620, 460]]
    k = np.zeros(n, kp_dtype)
    k["x"], k["y"], k["octave"] = pts[:, 0], pts[:, 1], lvl
    k["size"] = 31.0; k["angle"] = (np.arange(n) * 7 % 360).astype(np.float32); k["response"] = 20 + np.arange(n) % 50; k["class_id"] = -1
    return k


def depth_image():
    """Exact in float32, with non-positive holes like a real depth map."""
    yy, xx = np.mgrid[0:ROWS, 0:COLS]
    return (((xx * 7 + yy * 13) % 97).astype(np.float32) / np.float32(16.0) - np.float32(0.5)).astype(np.float32)


def cam_struct(orbx, name, bf=BF):
    c = CAMS[name]
    return orbx.Camera.make(*c[:4], *c[4:], bf=bf) if len(c) == 9 else orbx.Camera.make(*c[:4], c[4], c[5], c[6], c[7], 0.0, bf)
