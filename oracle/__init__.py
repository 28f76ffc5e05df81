"""oracle -- TEST INFRASTRUCTURE ONLY.

ctypes loaders for the two CPU oracles of the ORB front-end hot path:

* ``ref``  : oracle/_ref/liborb_ref.so  -- the reference's OWN sources (/root/reference/src/ORBextractor.cc,
             plus bodies of ORBmatcher.cc / Frame.cc) compiled unmodified against the OpenCV-free shim,
             with a monotonic allocator (canonical octree tie-break).  Built by oracle/ref/Makefile.
* ``port`` : oracle/_build/liborb_port.so -- our plain restatement (oracle/port/*.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (amos-slam_b200) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "liborb_ref.so")
PORT_SO = os.path.join(HERE, "_build", "liborb_port.so")
REFERENCE_ROOT = "/root/reference"

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

PORT_SOURCES = ["port/orb_port.cpp", "port/match_port.cpp"]


def build_port(force=False):
    """Compile the port oracle (plain C++, gcc only)."""
    srcs = [os.path.join(HERE, s) for s in PORT_SOURCES if os.path.exists(os.path.join(HERE, s))]
    deps = srcs + [os.path.join(HERE, "cvlite", "cvlite.hpp"), os.path.join(HERE, "port", "brief_pattern.inc")]
    if not force and os.path.exists(PORT_SO) and all(os.path.getmtime(PORT_SO) >= os.path.getmtime(d) for d in deps):
        return PORT_SO
    os.makedirs(os.path.dirname(PORT_SO), exist_ok=True)
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-std=c++14", "-fPIC", "-shared", "-w",
           "-o", PORT_SO] + srcs + ["-lpthread"]
    subprocess.check_call(cmd)
    return PORT_SO


def build_ref(force=False):
    """Compile the reference's own sources into oracle/_ref (only where /root/reference exists)."""
    if not os.path.isdir(REFERENCE_ROOT):
        return REF_SO if os.path.exists(REF_SO) else None
    args = ["make", "-C", os.path.join(HERE, "ref"), "REF=" + REFERENCE_ROOT]
    if force:
        args.append("-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return REF_SO


def have_ref():
    return os.path.exists(REF_SO)


_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_kpp = np.ctypeslib.ndpointer(dtype=KP_DTYPE, flags="C_CONTIGUOUS")

_libs = {}


def _lib(kind):
    if kind in _libs:
        return _libs[kind]
    if kind == "port":
        path = build_port()
    else:
        path = REF_SO
        if not os.path.exists(path):
            build_ref()
    lib = C.CDLL(path)
    p = kind + "_"
    f = getattr(lib, p + "extractor_create"); f.restype = C.c_void_p; f.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    f = getattr(lib, p + "extractor_destroy"); f.restype = None; f.argtypes = [C.c_void_p]
    f = getattr(lib, p + "extractor_info"); f.restype = C.c_int; f.argtypes = [C.c_void_p, C.POINTER(C.c_int), _f32p, _i32p, _i32p]
    f = getattr(lib, p + "extract"); f.restype = C.c_int; f.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, _kpp, _u8p, C.c_int]
    f = getattr(lib, p + "detect"); f.restype = C.c_int; f.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, _kpp, _i32p, C.c_int]
    f = getattr(lib, p + "pyramid_level"); f.restype = C.c_int; f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    f = getattr(lib, p + "distribute_octtree"); f.restype = C.c_int
    f.argtypes = [C.c_void_p, _kpp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _kpp, C.c_int]
    f = getattr(lib, p + "moving_keypoints"); f.restype = C.c_int
    f.argtypes = [C.c_void_p, _u8p, _f64p, C.c_int, C.c_int, _i32p, C.c_int, _i32p, C.c_int, _kpp, _i32p, _kpp]
    f = getattr(lib, p + "process_desp"); f.restype = C.c_int; f.argtypes = [C.c_void_p, _kpp, _i32p, _kpp, _u8p, C.c_int]
    if kind == "port":
        lib.port_level_candidates.restype = C.c_int; lib.port_level_candidates.argtypes = [C.c_void_p, C.c_int, _kpp, C.c_int]
        lib.cvl_c_resize.restype = None; lib.cvl_c_resize.argtypes = [_u8p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int]
        lib.cvl_c_blur7.restype = None; lib.cvl_c_blur7.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
        lib.cvl_c_fast.restype = C.c_int; lib.cvl_c_fast.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _kpp, C.c_int]
        lib.cvl_c_fast_smap.restype = None; lib.cvl_c_fast_smap.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
        lib.cvl_c_atan2.restype = None; lib.cvl_c_atan2.argtypes = [_f32p, _f32p, _f32p, C.c_int]
        lib.cvl_c_close31.restype = None; lib.cvl_c_close31.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
        lib.cvl_c_ellipse31.restype = None; lib.cvl_c_ellipse31.argtypes = [_u8p]
        lib.cvl_c_border101.restype = None; lib.cvl_c_border101.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p]
        lib.port_det_sincos.restype = None; lib.port_det_sincos.argtypes = [_f32p, _f32p, _f32p, C.c_int]
        lib.port_libm_sincosf.restype = None; lib.port_libm_sincosf.argtypes = [_f32p, _f32p, _f32p, C.c_int]
    else:
        lib.ref_pyramid_level_padded.restype = C.c_int; lib.ref_pyramid_level_padded.argtypes = [C.c_void_p, C.c_int, _u8p]
    _libs[kind] = lib
    return lib


class Extractor:
    """Common Python face of both oracles (kind = 'ref' | 'port')."""

    def __init__(self, kind, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.kind = kind
        self.lib = _lib(kind)
        self.p = kind + "_"
        self.h = getattr(self.lib, self.p + "extractor_create")(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.cap = nfeatures * 2 + 64 * nlevels + 4096
        nl = C.c_int()
        self.scale_factors = np.zeros(nlevels, np.float32)
        self.features_per_level = np.zeros(nlevels, np.int32)
        self.umax = np.zeros(16, np.int32)
        getattr(self.lib, self.p + "extractor_info")(self.h, C.byref(nl), self.scale_factors, self.features_per_level, self.umax)

    def __del__(self):
        try:
            getattr(self.lib, self.p + "extractor_destroy")(self.h)
        except Exception:
            pass

    def extract(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        kp = np.zeros(self.cap, KP_DTYPE); desc = np.zeros((self.cap, 32), np.uint8)
        n = getattr(self.lib, self.p + "extract")(self.h, img, img.shape[0], img.shape[1], img.strides[0], kp, desc, self.cap)
        assert n >= 0, n
        return kp[:n].copy(), desc[:n].copy()

    def detect(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        kp = np.zeros(self.cap, KP_DTYPE); counts = np.zeros(self.nlevels, np.int32)
        n = getattr(self.lib, self.p + "detect")(self.h, img, img.shape[0], img.shape[1], img.strides[0], kp, counts, self.cap)
        assert n >= 0, n
        return kp[:n].copy(), counts

    def pyramid_level(self, level):
        r, c = C.c_int(), C.c_int()
        rc = getattr(self.lib, self.p + "pyramid_level")(self.h, level, None, C.byref(r), C.byref(c))
        assert rc == 0
        out = np.zeros((r.value, c.value), np.uint8)
        getattr(self.lib, self.p + "pyramid_level")(self.h, level, out.ctypes.data_as(C.c_void_p), C.byref(r), C.byref(c))
        return out

    def level_candidates(self, level):
        assert self.kind == "port"
        cap = 1 << 20
        out = np.zeros(cap, KP_DTYPE)
        n = self.lib.port_level_candidates(self.h, level, out, cap)
        assert n >= 0
        return out[:n].copy()

    def distribute(self, cand, minX, maxX, minY, maxY, N, level=0):
        cand = np.ascontiguousarray(cand, KP_DTYPE)
        cap = max(4 * N + 64, 256)
        out = np.zeros(cap, KP_DTYPE)
        n = getattr(self.lib, self.p + "distribute_octtree")(self.h, cand, len(cand), minX, maxX, minY, maxY, N, level, out, cap)
        assert n >= 0, n
        return out[:n].copy()

    def moving_keypoints(self, mask, label, centers_id, rm_vector, kp, counts):
        mask = np.ascontiguousarray(mask, np.uint8); label = np.ascontiguousarray(label, np.float64)
        centers_id = np.ascontiguousarray(centers_id, np.int32); rm_vector = np.ascontiguousarray(rm_vector, np.int32)
        kp = np.ascontiguousarray(kp, KP_DTYPE).copy(); counts = np.ascontiguousarray(counts, np.int32).copy()
        culled = np.zeros(max(len(kp), 1), KP_DTYPE)
        nc = getattr(self.lib, self.p + "moving_keypoints")(self.h, mask, label, mask.shape[0], mask.shape[1], centers_id, len(centers_id),
                                                            rm_vector, len(rm_vector), kp, counts, culled)
        return kp[:int(counts.sum())].copy(), counts, culled[:nc].copy()

    def process_desp(self, kp, counts):
        kp = np.ascontiguousarray(kp, KP_DTYPE); counts = np.ascontiguousarray(counts, np.int32)
        out = np.zeros(max(len(kp), 1), KP_DTYPE); desc = np.zeros((max(len(kp), 1), 32), np.uint8)
        n = getattr(self.lib, self.p + "process_desp")(self.h, kp, counts, out, desc, len(out))
        assert n >= 0
        return out[:n].copy(), desc[:n].copy()


def port_lib():
    return _lib("port")


def ref_lib():
    return _lib("ref")


# ------------------------------------------------------------------------------------------------
# Synthetic inputs: seeded, exercises both FAST thresholds.  Pure numpy (no cv2) so
# the same frames are generated on the GPU box and here.
# ------------------------------------------------------------------------------------------------
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
