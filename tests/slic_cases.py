"""Inputs of the SLIC tests: seeded synthetic colour frames (three synthetic grey frames as B, G, R) and a smooth 16-bit depth image."""
import numpy as np
from tools.synth import synth_frame

# (name, width, height, seed): a VGA-proportioned frame, an odd geometry (ragged last block row / column), a tiny one, and a flat
# image (every distance ties: exercises the lowest-index rule and dead centres)
CASES = [("qvga", 320, 240, 1), ("odd", 203, 147, 2), ("tiny", 23, 17, 3), ("flat", 64, 48, 4)]


def bgr_frame(seed, w, h):
    return np.stack([synth_frame(seed * 3 + k, w, h) for k in range(3)], -1)


def depth_frame(seed, w, h):
    rng = np.random.default_rng(7000 + seed)
    yy, xx = np.mgrid[0:h, 0:w]
    d = 1500 + 800 * np.sin(xx / 37.0 + seed) + 600 * np.cos(yy / 23.0) + rng.normal(0, 20, (h, w))
    d[rng.random((h, w)) < 0.02] = 0                      # holes, as a real depth camera has
    return np.clip(np.rint(d), 0, 65535).astype(np.uint16)
