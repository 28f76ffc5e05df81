"""Seeded synthetic inputs for the ORB front-end.  Pure numpy, so the GPU box and this
container generate identical frames.  Neutral module: used by tests/, bench.py and re-exported by oracle/."""
import numpy as np


def synth_frame(seed, width=640, height=480):
    rng = np.random.default_rng(1000 + int(seed))
    img = np.full((height, width), 128.0, np.float32)
    nshapes = (width * height) // 600
    yy, xx = np.mgrid[0:height, 0:width]
    kinds = rng.integers(0, 2, nshapes)
    cx = rng.integers(0, width, nshapes); cy = rng.integers(0, height, nshapes)
    sw = rng.integers(4, 61, nshapes); sh = rng.integers(4, 61, nshapes)
    grey = rng.integers(0, 256, nshapes)
    for k in range(nshapes):
        x0, x1 = max(cx[k] - sw[k] // 2, 0), min(cx[k] + sw[k] // 2 + 1, width)
        y0, y1 = max(cy[k] - sh[k] // 2, 0), min(cy[k] + sh[k] // 2 + 1, height)
        if kinds[k] == 0:
            img[y0:y1, x0:x1] = grey[k]
        else:
            r = sw[k] / 2.0
            sub = (xx[y0:y1, x0:x1] - cx[k]) ** 2 + (yy[y0:y1, x0:x1] - cy[k]) ** 2 <= r * r
            img[y0:y1, x0:x1][sub] = grey[k]
    # 3x3 Gaussian sigma 0.8 (separable, edge-replicated)
    g = np.exp(-np.array([-1.0, 0.0, 1.0]) ** 2 / (2 * 0.8 * 0.8)); g /= g.sum()
    p = np.pad(img, 1, mode="edge")
    img = g[0] * p[1:-1, :-2] + g[1] * p[1:-1, 1:-1] + g[2] * p[1:-1, 2:]
    p = np.pad(img, 1, mode="edge")
    img = g[0] * p[:-2, 1:-1] + g[1] * p[1:-1, 1:-1] + g[2] * p[2:, 1:-1]
    img = img + rng.normal(0.0, 3.0, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synth_batch(n, width=640, height=480, seed0=0, distinct=None):
    """n frames; `distinct` base frames are generated (slow path) and the rest are derived from them by
    circular shifts + flips so that every frame of the batch is different but generation stays fast."""
    distinct = min(n, distinct or n)
    base = [synth_frame(seed0 + i, width, height) for i in range(distinct)]
    out = np.empty((n, height, width), np.uint8)
    for i in range(n):
        f = base[i % distinct]
        r = i // distinct
        if r:
            f = np.roll(f, (13 * r, 29 * r), axis=(0, 1))
            if r & 1:
                f = f[:, ::-1]
            if r & 2:
                f = f[::-1, :]
        out[i] = f
    return out


def synth_mask(seed, width, height):
    """2-4 random filled ellipses x255 (the YOLACT 'person' mask stand-in)."""
    rng = np.random.default_rng(5000 + int(seed))
    yy, xx = np.mgrid[0:height, 0:width]
    m = np.zeros((height, width), np.uint8)
    for _ in range(int(rng.integers(2, 5))):
        cx, cy = rng.integers(0, width), rng.integers(0, height)
        a, b = rng.integers(width // 20, width // 6), rng.integers(height // 20, height // 4)
        m[((xx - cx) / float(a)) ** 2 + ((yy - cy) / float(b)) ** 2 <= 1.0] = 255
    return m


def warp_affine_nn(img, dx, dy, deg):
    """Frame B = frame A warped by a small known affine (nearest neighbour, edge clamp): matcher inputs."""
    h, w = img.shape
    t = np.deg2rad(deg)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    cx, cy = w / 2.0, h / 2.0
    xs = np.cos(t) * (xx - cx - dx) + np.sin(t) * (yy - cy - dy) + cx
    ys = -np.sin(t) * (xx - cx - dx) + np.cos(t) * (yy - cy - dy) + cy
    xi = np.clip(np.rint(xs).astype(np.int64), 0, w - 1)
    yi = np.clip(np.rint(ys).astype(np.int64), 0, h - 1)
    return img[yi, xi]


def stereo_right_from_left(left, seed=0):
    """Right image = left resampled with piecewise-constant disparity (5..60 px) per horizontal band."""
    h, w = left.shape
    rng = np.random.default_rng(9000 + int(seed))
    right = np.empty_like(left)
    y = 0
    while y < h:
        bh = int(rng.integers(20, 60)); d = int(rng.integers(5, 61))
        band = left[y:y + bh]
        right[y:y + bh] = np.concatenate([band[:, d:], np.repeat(band[:, -1:], d, axis=1)], axis=1)   # xR = xL - d
        y += bh
    return right


def synth_labels(seed, width, height, blocks=5):
    """Super-pixel stand-in of SURVEY.md 8d: a blocks x blocks label map (ids 1..blocks^2, CV_64F like imLS), the cluster id of every
    super-pixel (centers[i].id) and rm_vector with 2-4 flagged clusters."""
    rng = np.random.default_rng(9000 + int(seed))
    by = np.minimum(np.arange(height) * blocks // height, blocks - 1); bx = np.minimum(np.arange(width) * blocks // width, blocks - 1)
    label = (by[:, None] * blocks + bx[None, :] + 1).astype(np.float64)
    nclusters = 8
    centers_id = rng.integers(0, nclusters, blocks * blocks).astype(np.int32)
    rm = np.zeros(nclusters, np.int32); rm[rng.choice(nclusters, int(rng.integers(2, 5)), replace=False)] = 1
    return label, centers_id, rm


def _synth_one(args):
    seed, width, height = args
    return synth_frame(seed, width, height)


def synth_batch_distinct(n, width=640, height=480, seed0=1000, workers=None):
    """n frames from n DISTINCT seeds seed0 .. seed0 + n - 1 ("batch of 1024 distinct seeds"), generated by a process pool
    (call before CUDA is initialised: the pool forks)."""
    import os
    from concurrent.futures import ProcessPoolExecutor
    workers = workers or max(1, min(os.cpu_count() or 1, 32))
    out = np.empty((n, height, width), np.uint8)
    if workers == 1 or n < 8:
        for i in range(n):
            out[i] = synth_frame(seed0 + i, width, height)
        return out
    with ProcessPoolExecutor(workers) as ex:
        for i, f in enumerate(ex.map(_synth_one, [(seed0 + i, width, height) for i in range(n)], chunksize=max(1, n // (4 * workers)))):
            out[i] = f
    return out
