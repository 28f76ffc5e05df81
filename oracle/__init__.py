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

PORT_SOURCES = ["port/orb_port.cpp", "port/match_port.cpp", "port/bow_port.cpp", "port/slic_port.cpp"]


def build_port(force=False):
    """Compile the port oracle (plain C++, gcc only)."""
    srcs = [os.path.join(HERE, s) for s in PORT_SOURCES if os.path.exists(os.path.join(HERE, s))]
    deps = srcs + [os.path.join(HERE, "cvlite", "cvlite.hpp"), os.path.join(HERE, "port", "brief_pattern.inc")]
    if not force and os.path.exists(PORT_SO) and all(os.path.getmtime(PORT_SO) >= os.path.getmtime(d) for d in deps):
        return PORT_SO
    os.makedirs(os.path.dirname(PORT_SO), exist_ok=True)
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-std=gnu++17", "-fPIC", "-shared", "-w",
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
        lib.cvl_c_gemm.restype = None; lib.cvl_c_gemm.argtypes = [_f32p, C.c_int, C.c_int, _f32p, C.c_int, C.c_void_p, _f32p]
        lib.cvl_c_atan2.restype = None; lib.cvl_c_atan2.argtypes = [_f32p, _f32p, _f32p, C.c_int]
        lib.cvl_c_close31.restype = None; lib.cvl_c_close31.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
        lib.cvl_c_ellipse31.restype = None; lib.cvl_c_ellipse31.argtypes = [_u8p]
        lib.cvl_c_border101.restype = None; lib.cvl_c_border101.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p]
        lib.port_det_sincos.restype = None; lib.port_det_sincos.argtypes = [_f32p, _f32p, _f32p, C.c_int]
        lib.port_libm_sincosf.restype = None; lib.port_libm_sincosf.argtypes = [_f32p, _f32p, _f32p, C.c_int]
        lib.port_sincos_sweep.restype = C.c_longlong; lib.port_sincos_sweep.argtypes = [C.c_uint, C.c_uint, np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS"), C.c_int]
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


# Synthetic inputs live in tools/synth.py (neutral module shared by tests and bench.py); re-exported here.
import sys as _sys
_sys.path.insert(0, os.path.dirname(HERE))
from tools.synth import synth_frame  # noqa: E402,F401


# ------------------------------------------------------------------------------------------------
# Matcher oracles (ORBmatcher / Frame grid / ComputeStereoMatches)
# ------------------------------------------------------------------------------------------------
class FrameViewC(C.Structure):
    """Mirror of orbx_frame_view (include/orbx_b200.h)."""
    _fields_ = [("n", C.c_int), ("keys_un", C.c_void_p), ("descriptors", C.c_void_p), ("u_right", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
                ("gw_inv", C.c_float), ("gh_inv", C.c_float), ("nlevels", C.c_int), ("scale_factors", C.c_void_p)]


class FrameData:
    """What a Frame exposes to the matchers: undistorted keypoints, descriptors, image bounds, 64x48 grid constants
    (mfGridElementWidthInv = 64/(mnMaxX-mnMinX) etc., /root/reference/src/Frame.cc:219-220)."""

    def __init__(self, keys, desc, width, height, scale_factors, u_right=None):
        self.keys = np.ascontiguousarray(keys, KP_DTYPE)
        self.desc = np.ascontiguousarray(desc, np.uint8)
        self.scale_factors = np.ascontiguousarray(scale_factors, np.float32)
        self.u_right = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        self.min_x, self.min_y, self.max_x, self.max_y = np.float32(0), np.float32(0), np.float32(width), np.float32(height)
        self.gw_inv = np.float32(64.0) / (self.max_x - self.min_x)
        self.gh_inv = np.float32(48.0) / (self.max_y - self.min_y)

    def view(self):
        v = FrameViewC()
        v.n = len(self.keys); v.keys_un = self.keys.ctypes.data; v.descriptors = self.desc.ctypes.data
        v.u_right = self.u_right.ctypes.data if self.u_right is not None else None
        v.min_x, v.min_y, v.max_x, v.max_y = float(self.min_x), float(self.min_y), float(self.max_x), float(self.max_y)
        v.gw_inv, v.gh_inv = float(self.gw_inv), float(self.gh_inv)
        v.nlevels = len(self.scale_factors); v.scale_factors = self.scale_factors.ctypes.data
        return v


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, np.uint8)


class Matcher:
    """ORBmatcher(nnratio, checkOri) oracle, kind = 'ref' (reference bodies) | 'port' (restatement)."""

    def __init__(self, kind, nnratio=0.6, check_ori=True):
        self.kind, self.lib, self.p = kind, _lib(kind), kind + "_"
        self.nnratio, self.check_ori = float(nnratio), int(bool(check_ori))
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        fv = C.POINTER(FrameViewC)
        f = getattr(self.lib, self.p + "descriptor_distance"); f.restype = None; f.argtypes = [vp, vp, ci, vp]
        f = getattr(self.lib, self.p + "get_features_in_area"); f.restype = ci; f.argtypes = [fv, cf, cf, cf, ci, ci, vp, ci]
        f = getattr(self.lib, self.p + "search_for_initialization"); f.restype = ci; f.argtypes = [cf, ci, fv, fv, vp, vp, ci]
        f = getattr(self.lib, self.p + "search_by_projection_points"); f.restype = ci
        f.argtypes = [cf, ci, fv, ci, vp, vp, vp, vp, vp, vp, vp, cf, vp]
        if kind == "ref":
            self.lib.ref_search_by_projection_frame.restype = ci
            self.lib.ref_search_by_projection_frame.argtypes = [cf, ci, fv, ci, vp, vp, vp, vp, vp, vp, vp, cf, ci, cf, cf, cf, cf, cf, cf, vp, vp, vp]
            self.lib.ref_compute_stereo_matches.restype = ci
            self.lib.ref_compute_stereo_matches.argtypes = [vp, vp, vp, vp, ci, vp, vp, ci, cf, cf, vp, vp]
        else:
            self.lib.port_search_by_projection_frame.restype = ci
            self.lib.port_search_by_projection_frame.argtypes = [cf, ci, fv, ci, vp, vp, vp, vp, vp, vp, vp, vp, cf, ci, ci, cf, vp]
            self.lib.port_compute_stereo_matches.restype = ci
            self.lib.port_compute_stereo_matches.argtypes = [ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, vp, vp, ci, cf, cf, vp, vp]

    def descriptor_distance(self, a, b):
        a, b = _u8(a), _u8(b)
        out = np.zeros(len(a), np.int32)
        getattr(self.lib, self.p + "descriptor_distance")(a.ctypes.data, b.ctypes.data, len(a), out.ctypes.data)
        return out

    def features_in_area(self, F, x, y, r, min_level=-1, max_level=-1):
        out = np.zeros(max(len(F.keys), 1), np.int32)
        n = getattr(self.lib, self.p + "get_features_in_area")(C.byref(F.view()), x, y, r, min_level, max_level, out.ctypes.data, len(out))
        return out[:n].copy()

    def search_for_initialization(self, F1, F2, prev_matched, window_size=10):
        prev = np.ascontiguousarray(prev_matched, np.float32).copy()
        m12 = np.zeros(len(F1.keys), np.int32)
        nm = getattr(self.lib, self.p + "search_for_initialization")(self.nnratio, self.check_ori, C.byref(F1.view()), C.byref(F2.view()), prev.ctypes.data, m12.ctypes.data, int(window_size))
        return nm, m12, prev

    def search_by_projection_points(self, F, track_uv, track_ur, track_level, track_view_cos, mp_desc, mp_observed, f_occupied, th):
        uv = np.ascontiguousarray(track_uv, np.float32); ur = np.ascontiguousarray(track_ur, np.float32)
        lv = np.ascontiguousarray(track_level, np.int32); vc = np.ascontiguousarray(track_view_cos, np.float32)
        d, ob, oc = _u8(mp_desc), _u8(mp_observed), _u8(f_occupied)
        fm = np.zeros(len(F.keys), np.int32)
        nm = getattr(self.lib, self.p + "search_by_projection_points")(self.nnratio, self.check_ori, C.byref(F.view()), len(lv), uv.ctypes.data, ur.ctypes.data, lv.ctypes.data,
                                                                       vc.ctypes.data, d.ctypes.data, ob.ctypes.data, oc.ctypes.data if oc is not None else None, float(th), fm.ctypes.data)
        return nm, fm

    def search_by_projection_frame_ref(self, cur, cam_xyz, last_octave, last_angle, mp_desc, valid, mp_observed, cur_occupied, th, b_mono, mb, mbf, fx, fy, cx, cy):
        """Reference body with identity poses; returns (nmatches, cur_match, proj_uv, proj_invz)."""
        assert self.kind == "ref"
        xyz = np.ascontiguousarray(cam_xyz, np.float32); n = len(xyz)
        lo = np.ascontiguousarray(last_octave, np.int32); la = np.ascontiguousarray(last_angle, np.float32)
        d, va, ob, oc = _u8(mp_desc), _u8(valid), _u8(mp_observed), _u8(cur_occupied)
        uv = np.zeros((n, 2), np.float32); iz = np.zeros(n, np.float32); cm = np.zeros(len(cur.keys), np.int32)
        nm = self.lib.ref_search_by_projection_frame(self.nnratio, self.check_ori, C.byref(cur.view()), n, xyz.ctypes.data, lo.ctypes.data, la.ctypes.data, d.ctypes.data,
                                                     va.ctypes.data, ob.ctypes.data, oc.ctypes.data if oc is not None else None, float(th), int(b_mono), float(mb), float(mbf),
                                                     float(fx), float(fy), float(cx), float(cy), uv.ctypes.data, iz.ctypes.data, cm.ctypes.data)
        return nm, cm, uv, iz

    def search_by_projection_frame_port(self, cur, proj_uv, proj_invz, last_octave, last_angle, mp_desc, valid, mp_observed, cur_occupied, th, forward, backward, mbf):
        assert self.kind == "port"
        uv = np.ascontiguousarray(proj_uv, np.float32); iz = np.ascontiguousarray(proj_invz, np.float32)
        lo = np.ascontiguousarray(last_octave, np.int32); la = np.ascontiguousarray(last_angle, np.float32)
        d, va, ob, oc = _u8(mp_desc), _u8(valid), _u8(mp_observed), _u8(cur_occupied)
        cm = np.zeros(len(cur.keys), np.int32)
        nm = self.lib.port_search_by_projection_frame(self.nnratio, self.check_ori, C.byref(cur.view()), len(iz), uv.ctypes.data, iz.ctypes.data, lo.ctypes.data, la.ctypes.data,
                                                      d.ctypes.data, va.ctypes.data, ob.ctypes.data, oc.ctypes.data if oc is not None else None, float(th), int(forward), int(backward),
                                                      float(mbf), cm.ctypes.data)
        return nm, cm

    def search_by_projection_keyframe_ref(self, cur, cam_xyz, predicted_level, kf_angle, mp_desc, state, min_dist, max_dist, cur_occupied, th, orb_dist, fx, fy, cx, cy):
        """Reference body (ORBmatcher.cc:1731) with identity pose; state: 0 none, 1 good, 2 bad, 3 already found.  -> (nmatches, cur_match, proj_uv)."""
        assert self.kind == "ref"
        xyz = np.ascontiguousarray(cam_xyz, np.float32); n = len(xyz)
        lv = np.ascontiguousarray(predicted_level, np.int32); an = np.ascontiguousarray(kf_angle, np.float32)
        d, st, oc = _u8(mp_desc), _u8(state), _u8(cur_occupied)
        mn = np.ascontiguousarray(min_dist, np.float32); mx = np.ascontiguousarray(max_dist, np.float32)
        uv = np.zeros((n, 2), np.float32); cm = np.zeros(max(len(cur.keys), 1), np.int32)
        f = self.lib.ref_search_by_projection_keyframe; f.restype = C.c_int
        f.argtypes = [C.c_float, C.c_int, C.POINTER(FrameViewC), C.c_int] + [C.c_void_p] * 8 + [C.c_float, C.c_int] + [C.c_float] * 4 + [C.c_void_p, C.c_void_p]
        nm = f(self.nnratio, self.check_ori, C.byref(cur.view()), n, xyz.ctypes.data, lv.ctypes.data, an.ctypes.data, d.ctypes.data, st.ctypes.data, mn.ctypes.data, mx.ctypes.data,
               oc.ctypes.data if oc is not None else None, float(th), int(orb_dist), float(fx), float(fy), float(cx), float(cy), uv.ctypes.data, cm.ctypes.data)
        return nm, cm[:len(cur.keys)], uv

    def search_by_projection_keyframe_port(self, cur, proj_uv, predicted_level, kf_angle, mp_desc, valid, cur_occupied, th, orb_dist):
        assert self.kind == "port"
        uv = np.ascontiguousarray(proj_uv, np.float32); lv = np.ascontiguousarray(predicted_level, np.int32); an = np.ascontiguousarray(kf_angle, np.float32)
        d, va, oc = _u8(mp_desc), _u8(valid), _u8(cur_occupied)
        cm = np.zeros(max(len(cur.keys), 1), np.int32)
        f = self.lib.port_search_by_projection_keyframe; f.restype = C.c_int
        f.argtypes = [C.c_float, C.c_int, C.POINTER(FrameViewC), C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_int, C.c_void_p]
        nm = f(self.nnratio, self.check_ori, C.byref(cur.view()), len(lv), uv.ctypes.data, lv.ctypes.data, an.ctypes.data, d.ctypes.data, va.ctypes.data,
               oc.ctypes.data if oc is not None else None, float(th), int(orb_dist), cm.ctypes.data)
        return nm, cm[:len(cur.keys)]

    def search_by_projection_keyframe_points_ref(self, kf, cam_xyz, predicted_level, mp_desc, state, found_at, facing, min_dist, max_dist, kf_matched, th, fx, fy, cx, cy):
        """Reference body (ORBmatcher.cc:388) with Scw = identity.  -> (nmatches, kf_match, proj_uv)."""
        assert self.kind == "ref"
        xyz = np.ascontiguousarray(cam_xyz, np.float32); n = len(xyz)
        lv = np.ascontiguousarray(predicted_level, np.int32); fa = np.ascontiguousarray(found_at, np.int32)
        d, st, fc_, km = _u8(mp_desc), _u8(state), _u8(facing), _u8(kf_matched)
        mn = np.ascontiguousarray(min_dist, np.float32); mx = np.ascontiguousarray(max_dist, np.float32)
        uv = np.zeros((n, 2), np.float32); out = np.zeros(max(len(kf.keys), 1), np.int32)
        f = self.lib.ref_search_by_projection_keyframe_points; f.restype = C.c_int
        f.argtypes = [C.c_float, C.c_int, C.POINTER(FrameViewC), C.c_int] + [C.c_void_p] * 9 + [C.c_int] + [C.c_float] * 4 + [C.c_void_p, C.c_void_p]
        nm = f(self.nnratio, self.check_ori, C.byref(kf.view()), n, xyz.ctypes.data, lv.ctypes.data, d.ctypes.data, st.ctypes.data, fa.ctypes.data, fc_.ctypes.data,
               mn.ctypes.data, mx.ctypes.data, km.ctypes.data if km is not None else None, int(th), float(fx), float(fy), float(cx), float(cy), uv.ctypes.data, out.ctypes.data)
        return nm, out[:len(kf.keys)], uv

    def search_by_projection_keyframe_points_port(self, kf, proj_uv, predicted_level, mp_desc, valid, kf_matched, th):
        assert self.kind == "port"
        uv = np.ascontiguousarray(proj_uv, np.float32); lv = np.ascontiguousarray(predicted_level, np.int32)
        d, va, km = _u8(mp_desc), _u8(valid), _u8(kf_matched)
        out = np.zeros(max(len(kf.keys), 1), np.int32)
        f = self.lib.port_search_by_projection_keyframe_points; f.restype = C.c_int
        f.argtypes = [C.POINTER(FrameViewC), C.c_int] + [C.c_void_p] * 5 + [C.c_float, C.c_void_p]
        nm = f(C.byref(kf.view()), len(lv), uv.ctypes.data, lv.ctypes.data, d.ctypes.data, va.ctypes.data, km.ctypes.data if km is not None else None, float(th), out.ctypes.data)
        return nm, out[:len(kf.keys)]

    def search_by_sim3_ref(self, kf1, kf2, side1, side2, already12, th, fx, fy, cx, cy):
        """Reference body (ORBmatcher.cc:1290), identity Sim3.  side = dict(xyz, lvl, desc, state, mind, maxd).  -> (nfound, match12)."""
        assert self.kind == "ref"
        def arrs(s):
            return [np.ascontiguousarray(s["xyz"], np.float32), np.ascontiguousarray(s["lvl"], np.int32), _u8(s["desc"]), _u8(s["state"]),
                    np.ascontiguousarray(s["mind"], np.float32), np.ascontiguousarray(s["maxd"], np.float32)]
        a, b = arrs(side1), arrs(side2); al = np.ascontiguousarray(already12, np.int32)
        m12 = np.zeros(max(len(kf1.keys), 1), np.int32)
        f = self.lib.ref_search_by_sim3; f.restype = C.c_int
        f.argtypes = [C.POINTER(FrameViewC)] * 2 + [C.c_void_p] * 13 + [C.c_float] * 5 + [C.c_void_p]
        nf = f(C.byref(kf1.view()), C.byref(kf2.view()), *[x.ctypes.data for x in a], *[x.ctypes.data for x in b], al.ctypes.data, float(th), float(fx), float(fy), float(cx), float(cy), m12.ctypes.data)
        return nf, m12[:len(kf1.keys)]

    def search_by_sim3_port(self, kf1, kf2, uv1, lvl1, desc1, valid1, uv2, lvl2, desc2, valid2, th):
        assert self.kind == "port"
        a = [np.ascontiguousarray(uv1, np.float32), np.ascontiguousarray(lvl1, np.int32), _u8(desc1), _u8(valid1),
             np.ascontiguousarray(uv2, np.float32), np.ascontiguousarray(lvl2, np.int32), _u8(desc2), _u8(valid2)]
        m12 = np.zeros(max(len(kf1.keys), 1), np.int32)
        f = self.lib.port_search_by_sim3; f.restype = C.c_int
        f.argtypes = [C.POINTER(FrameViewC)] * 2 + [C.c_void_p] * 8 + [C.c_float, C.c_void_p]
        nf = f(C.byref(kf1.view()), C.byref(kf2.view()), *[x.ctypes.data for x in a], float(th), m12.ctypes.data)
        return nf, m12[:len(kf1.keys)]

    def fuse_ref(self, sim3_form, kf, kf_has_mp, cam_xyz, predicted_level, mp_desc, state, facing, min_dist, max_dist, inv_sigma2, th, bf, fx, fy, cx, cy):
        """Both ORBmatcher::Fuse bodies (ORBmatcher.cc:1020, 1179) at an identity pose.  -> (nFused, best_idx, proj_uv, proj_ur)."""
        assert self.kind == "ref"
        xyz = np.ascontiguousarray(cam_xyz, np.float32); n = len(xyz)
        lv = np.ascontiguousarray(predicted_level, np.int32); d, st, fc_, hm = _u8(mp_desc), _u8(state), _u8(facing), _u8(kf_has_mp)
        mn = np.ascontiguousarray(min_dist, np.float32); mx = np.ascontiguousarray(max_dist, np.float32); isg = np.ascontiguousarray(inv_sigma2, np.float32)
        ur_kf = np.ascontiguousarray(kf.u_right if kf.u_right is not None else np.full(len(kf.keys), -1, np.float32), np.float32)
        uv = np.zeros((n, 2), np.float32); ur = np.zeros(n, np.float32); best = np.zeros(max(n, 1), np.int32)
        f = self.lib.ref_fuse; f.restype = C.c_int
        f.argtypes = [C.c_int, C.POINTER(FrameViewC), C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 8 + [C.c_float] * 6 + [C.c_void_p] * 3
        nf = f(int(bool(sim3_form)), C.byref(kf.view()), ur_kf.ctypes.data, hm.ctypes.data, n, xyz.ctypes.data, lv.ctypes.data, d.ctypes.data, st.ctypes.data, fc_.ctypes.data,
               mn.ctypes.data, mx.ctypes.data, isg.ctypes.data, float(th), float(bf), float(fx), float(fy), float(cx), float(cy), uv.ctypes.data, ur.ctypes.data, best.ctypes.data)
        return nf, best[:n], uv, ur

    def fuse_search_port(self, kf, proj_uv, proj_ur, predicted_level, mp_desc, valid, inv_sigma2, th):
        assert self.kind == "port"
        uv = np.ascontiguousarray(proj_uv, np.float32); lv = np.ascontiguousarray(predicted_level, np.int32); d, va = _u8(mp_desc), _u8(valid)
        ur = None if proj_ur is None else np.ascontiguousarray(proj_ur, np.float32); isg = np.ascontiguousarray(inv_sigma2, np.float32)
        best = np.zeros(max(len(lv), 1), np.int32)
        f = self.lib.port_fuse_search; f.restype = None
        f.argtypes = [C.POINTER(FrameViewC), C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_void_p]
        f(C.byref(kf.view()), len(lv), uv.ctypes.data, None if ur is None else ur.ctypes.data, lv.ctypes.data, d.ctypes.data, va.ctypes.data, isg.ctypes.data, float(th), best.ctypes.data)
        return best[:len(lv)]

    def compute_stereo_matches(self, ext_left, ext_right, keys_left, desc_left, keys_right, desc_right, mb, mbf):
        """ext_left / ext_right: oracle.Extractor of the same kind whose last extract() saw the left / right image."""
        kl = np.ascontiguousarray(keys_left, KP_DTYPE); kr = np.ascontiguousarray(keys_right, KP_DTYPE)
        dl, dr = _u8(desc_left), _u8(desc_right)
        ur = np.zeros(len(kl), np.float32); dep = np.zeros(len(kl), np.float32)
        if self.kind == "ref":
            self.lib.ref_compute_stereo_matches(ext_left.h, ext_right.h, kl.ctypes.data, dl.ctypes.data, len(kl), kr.ctypes.data, dr.ctypes.data, len(kr), float(mb), float(mbf),
                                                ur.ctypes.data, dep.ctypes.data)
        else:
            nl = ext_left.nlevels
            pl = [np.ascontiguousarray(ext_left.pyramid_level(l)) for l in range(nl)]; pr = [np.ascontiguousarray(ext_right.pyramid_level(l)) for l in range(nl)]
            wl = np.array([p.shape[1] for p in pl], np.int32); hl = np.array([p.shape[0] for p in pl], np.int32)
            wr = np.array([p.shape[1] for p in pr], np.int32); hr = np.array([p.shape[0] for p in pr], np.int32)
            ptl = (C.c_void_p * nl)(*[p.ctypes.data for p in pl]); ptr_ = (C.c_void_p * nl)(*[p.ctypes.data for p in pr])
            sc = np.ascontiguousarray(ext_left.scale_factors, np.float32); isc = (np.float32(1.0) / sc).astype(np.float32)
            self.lib.port_compute_stereo_matches(nl, wl.ctypes.data, hl.ctypes.data, ptl, wr.ctypes.data, hr.ctypes.data, ptr_, sc.ctypes.data, isc.ctypes.data,
                                                 kl.ctypes.data, dl.ctypes.data, len(kl), kr.ctypes.data, dr.ctypes.data, len(kr), float(mb), float(mbf), ur.ctypes.data, dep.ctypes.data)
        return ur, dep


# ------------------------------------------------------------------------------------------------
# Frame steps between extractor and matchers (UndistortKeyPoints / ComputeImageBounds / ComputeStereoFromRGBD /
# AssignFeaturesToGrid; /root/reference/src/Frame.cc:1052-1176, 1576-1614, 431-461)
# ------------------------------------------------------------------------------------------------
def _cam9(cam):
    """cam = (fx, fy, cx, cy, k1, k2, p1, p2[, k3]) -> (float32[9], ndist)."""
    c = np.zeros(9, np.float32); c[:len(cam)] = np.asarray(cam, np.float32)
    return c, len(cam) - 4


def undistort_points(kind, pts, cam):
    """cv::undistortPoints(pts, pts, K, D, Mat(), K) of the oracle (kind = 'ref': the shim's, 'port': the restatement)."""
    lib = _lib(kind)
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2); out = np.zeros_like(pts)
    c, nd = _cam9(cam)
    f = getattr(lib, kind + "_undistort_points"); f.restype = None; f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    f(pts.ctypes.data, len(pts), c.ctypes.data, nd, out.ctypes.data)
    return out


def frame_build(kind, keys, cam, bf, rows, cols, depth_img=None):
    """Returns dict(keys_un, u_right, depth, bounds[6], cell_start[3073], entries[n]) as the reference's Frame holds them
    after UndistortKeyPoints -> ComputeStereoFromRGBD -> AssignFeaturesToGrid."""
    lib = _lib(kind)
    keys = np.ascontiguousarray(keys, KP_DTYPE); n = len(keys)
    c, nd = _cam9(cam)
    dimg = None if depth_img is None else np.ascontiguousarray(depth_img, np.float32)
    assert dimg is None or dimg.shape == (rows, cols)
    ku = np.zeros(max(n, 1), KP_DTYPE); ur = np.zeros(max(n, 1), np.float32); dep = np.zeros(max(n, 1), np.float32)
    b = np.zeros(6, np.float32); cs = np.zeros(64 * 48 + 1, np.int32); en = np.zeros(max(n, 1), np.int32)
    f = getattr(lib, kind + "_frame_build"); f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p] + [C.c_void_p] * 6
    f(keys.ctypes.data, n, c.ctypes.data, nd, float(bf), rows, cols, None if dimg is None else dimg.ctypes.data,
      ku.ctypes.data, ur.ctypes.data, dep.ctypes.data, b.ctypes.data, cs.ctypes.data, en.ctypes.data)
    return dict(keys_un=ku[:n], u_right=ur[:n], depth=dep[:n], bounds=b, cell_start=cs, entries=en[:cs[-1]])


# ------------------------------------------------------------------------------------------------
# Bag of words: DBoW2 transform (Frame::ComputeBoW) and ORBmatcher::SearchByBoW x2
# (/root/reference/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1124-1259, /root/reference/src/ORBmatcher.cc:230-382, 656-799)
# ------------------------------------------------------------------------------------------------
class Vocabulary:
    """ORBVocabulary from a flat node table (the rows of ORBvoc.txt): parent[i], is_leaf[i], descriptor[i], weight[i] describe node i + 1.
    weighting: 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY; scoring: 0 L1_NORM, 1 L2_NORM, 2 CHI_SQUARE, 3 KL, 4 BHATTACHARYYA, 5 DOT_PRODUCT."""

    def __init__(self, kind, k, L, parent, is_leaf, desc, weight, weighting=0, scoring=0):
        self.kind, self.lib = kind, _lib(kind)
        self.parent = np.ascontiguousarray(parent, np.int32); self.is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); self.weight = np.ascontiguousarray(weight, np.float64)
        f = getattr(self.lib, kind + "_voc_create"); f.restype = C.c_void_p
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.h = f(k, L, weighting, scoring, len(self.parent), self.parent.ctypes.data, self.is_leaf.ctypes.data, self.desc.ctypes.data, self.weight.ctypes.data)
        g = getattr(self.lib, kind + "_voc_destroy"); g.restype = None; g.argtypes = [C.c_void_p]

    def __del__(self):
        try:
            getattr(self.lib, self.kind + "_voc_destroy")(self.h)
        except Exception:
            pass

    def transform(self, desc, levelsup=4):
        """-> dict(word, node: per feature; bow_ids, bow_vals: BowVector in map order; fv_nodes, fv_offsets, fv_idx: FeatureVector)."""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); n = len(d); m = max(n, 1)
        word = np.zeros(m, np.int32); node = np.zeros(m, np.int32); bi = np.zeros(m, np.int32); bv = np.zeros(m, np.float64)
        fn = np.zeros(m, np.int32); fo = np.zeros(m + 1, np.int32); fi = np.zeros(m, np.int32); nb, nf = C.c_int(), C.c_int()
        f = getattr(self.lib, self.kind + "_voc_transform"); f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.POINTER(C.c_int)] + [C.c_void_p] * 3 + [C.POINTER(C.c_int)]
        f(self.h, d.ctypes.data, n, levelsup, word.ctypes.data, node.ctypes.data, bi.ctypes.data, bv.ctypes.data, C.byref(nb), fn.ctypes.data, fo.ctypes.data, fi.ctypes.data, C.byref(nf))
        return dict(word=word[:n], node=node[:n], bow_ids=bi[:nb.value], bow_vals=bv[:nb.value], fv_nodes=fn[:nf.value], fv_offsets=fo[:nf.value + 1], fv_idx=fi[:fo[nf.value]])


def search_by_bow(kind, nnratio, check_ori, kf_kf, keys1, desc1, valid1, fv1, keys2, desc2, valid2, fv2):
    """fv = dict with fv_nodes / fv_offsets / fv_idx (Vocabulary.transform).  -> (nmatches, match12[n1], match21[n2])."""
    lib = _lib(kind)
    k1 = np.ascontiguousarray(keys1, KP_DTYPE); k2 = np.ascontiguousarray(keys2, KP_DTYPE)
    d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
    v1 = np.ascontiguousarray(valid1, np.uint8); v2 = np.ascontiguousarray(np.ones(len(k2), np.uint8) if valid2 is None else valid2, np.uint8)
    a = [np.ascontiguousarray(fv1[x], np.int32) for x in ("fv_nodes", "fv_offsets", "fv_idx")]; b = [np.ascontiguousarray(fv2[x], np.int32) for x in ("fv_nodes", "fv_offsets", "fv_idx")]
    m12 = np.zeros(max(len(k1), 1), np.int32); m21 = np.zeros(max(len(k2), 1), np.int32)
    f = getattr(lib, kind + "_search_by_bow"); f.restype = C.c_int
    f.argtypes = [C.c_float, C.c_int, C.c_int] + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] * 2 + [C.c_void_p, C.c_void_p]
    nm = f(float(nnratio), int(bool(check_ori)), int(bool(kf_kf)), len(k1), k1.ctypes.data, d1.ctypes.data, v1.ctypes.data, len(a[0]), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data,
           len(k2), k2.ctypes.data, d2.ctypes.data, v2.ctypes.data, len(b[0]), b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data, m12.ctypes.data, m21.ctypes.data)
    return nm, m12[:len(k1)], m21[:len(k2)]


def search_for_triangulation(kind, check_ori, keys1, desc1, free1, ur1, fv1, keys2, desc2, free2, ur2, fv2, F12, c2_or_epipole, cam, scale_factors, level_sigma2, only_stereo):
    """ORBmatcher::SearchForTriangulation (ORBmatcher.cc:810).  free = the feature holds no map point.  kind 'ref': c2_or_epipole = camera 1's
    centre in camera 2 (the body derives the epipole, returned as third value); 'port': the epipole itself.  -> (nmatches, match12[, epipole])."""
    lib = _lib(kind)
    k1 = np.ascontiguousarray(keys1, KP_DTYPE); k2 = np.ascontiguousarray(keys2, KP_DTYPE)
    d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
    f1 = np.ascontiguousarray(free1, np.uint8); f2 = np.ascontiguousarray(free2, np.uint8)
    u1 = np.ascontiguousarray(ur1, np.float32); u2 = np.ascontiguousarray(ur2, np.float32)
    a = [np.ascontiguousarray(fv1[x], np.int32) for x in ("fv_nodes", "fv_offsets", "fv_idx")]; b = [np.ascontiguousarray(fv2[x], np.int32) for x in ("fv_nodes", "fv_offsets", "fv_idx")]
    F = np.ascontiguousarray(F12, np.float32).reshape(9); sc = np.ascontiguousarray(scale_factors, np.float32); sg = np.ascontiguousarray(level_sigma2, np.float32)
    m12 = np.zeros(max(len(k1), 1), np.int32)
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    if kind == "ref":
        c2 = np.ascontiguousarray(c2_or_epipole, np.float32); has1 = (1 - f1).astype(np.uint8); has2 = (1 - f2).astype(np.uint8); epi = np.zeros(2, np.float32)
        f = lib.ref_search_for_triangulation; f.restype = ci
        f.argtypes = [cf, ci] + [ci, vp, vp, vp, vp, ci, vp, vp, vp] * 2 + [vp, vp, cf, cf, cf, cf, ci, vp, vp, ci, vp, vp]
        nm = f(0.6, int(bool(check_ori)), len(k1), k1.ctypes.data, d1.ctypes.data, has1.ctypes.data, u1.ctypes.data, len(a[0]), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data,
               len(k2), k2.ctypes.data, d2.ctypes.data, has2.ctypes.data, u2.ctypes.data, len(b[0]), b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data,
               F.ctypes.data, c2.ctypes.data, *[float(x) for x in cam], len(sc), sc.ctypes.data, sg.ctypes.data, int(bool(only_stereo)), m12.ctypes.data, epi.ctypes.data)
        return nm, m12[:len(k1)], epi
    ex, ey = float(c2_or_epipole[0]), float(c2_or_epipole[1])
    f = lib.port_search_for_triangulation; f.restype = ci
    f.argtypes = [ci] + [ci, vp, vp, vp, vp, ci, vp, vp, vp] * 2 + [vp, cf, cf, vp, vp, ci, vp]
    nm = f(int(bool(check_ori)), len(k1), k1.ctypes.data, d1.ctypes.data, f1.ctypes.data, u1.ctypes.data, len(a[0]), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data,
           len(k2), k2.ctypes.data, d2.ctypes.data, f2.ctypes.data, u2.ctypes.data, len(b[0]), b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data,
           F.ctypes.data, ex, ey, sc.ctypes.data, sg.ctypes.data, int(bool(only_stereo)), m12.ctypes.data)
    return nm, m12[:len(k1)]


def distinctive_descriptors(kind, offsets, obs_desc, kf_of=None, kf_bad=None):
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:359).  'ref': the reference body on observation slots (kf_of[s] = KeyFrame of slot s,
    kf_bad per KeyFrame) -> chosen descriptors [n_points, 32]; 'port': best index per point over the given (already filtered) descriptors."""
    lib = _lib(kind)
    off = np.ascontiguousarray(offsets, np.int32); d = np.ascontiguousarray(obs_desc, np.uint8).reshape(-1, 32); n = len(off) - 1
    if kind == "ref":
        ko = np.ascontiguousarray(kf_of, np.int32); kb = np.ascontiguousarray(kf_bad, np.uint8); out = np.zeros((max(n, 1), 32), np.uint8)
        f = lib.ref_distinctive_descriptors; f.restype = C.c_int; f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        f(n, off.ctypes.data, d.ctypes.data, len(kb), ko.ctypes.data, kb.ctypes.data, out.ctypes.data)
        return out[:n]
    best = np.zeros(max(n, 1), np.int32)
    f = lib.port_distinctive_descriptors; f.restype = None; f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    f(n, off.ctypes.data, d.ctypes.data, best.ctypes.data)
    return best[:n]


# ---- SLIC stage of `cluster` (src/cluster.cc): Lab image in (cv2.cvtColor is the input boundary), labels + centres out ----
def slic(kind, lab, depth, length=5, m=10):
    """labels (rows x cols float64, the reference's CV_64F labelMask) and centres (n x 7 int32: x, y, L, A, B, D, label)."""
    lib = _lib(kind)
    f = getattr(lib, kind + "_slic")
    lab = np.ascontiguousarray(lab, np.uint8); depth = np.ascontiguousarray(depth, np.uint16)
    rows, cols = depth.shape
    assert lab.shape == (rows, cols, 3)
    cap = ((rows + length - 1) // length) * ((cols + length - 1) // length)
    labels = np.zeros((rows, cols), np.float64); centers = np.zeros((cap, 7), np.int32); n = C.c_int(0)
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    rc = f(lab.ctypes.data, depth.ctypes.data, rows, cols, length, m, labels.ctypes.data, centers.ctypes.data, cap, C.byref(n))
    assert rc == 0
    return labels, centers[:n.value].copy()


def slic_kmeans_ref(centers, seeds):
    """k-means of the centres with explicit seed indices (canonical-seed contract): cluster id per centre."""
    lib = _lib("ref")
    centers = np.ascontiguousarray(centers, np.int32); seeds = np.ascontiguousarray(seeds, np.int32)
    ids = np.zeros(len(centers), np.int32)
    lib.ref_slic_kmeans.restype = C.c_int
    lib.ref_slic_kmeans.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    assert lib.ref_slic_kmeans(centers.ctypes.data, len(centers), seeds.ctypes.data, len(seeds), ids.ctypes.data) == 0
    return ids


def slic_gradient_ref(lab):
    """cvlite's Sobel(CV_64F, 0/1, 1/0, 3) + addWeighted(0.5, 0.5) as the reference's SLIC() calls them."""
    lib = _lib("ref")
    lab = np.ascontiguousarray(lab, np.uint8)
    out = np.zeros(lab.shape, np.float64)
    lib.ref_slic_gradient.restype = C.c_int
    lib.ref_slic_gradient.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    assert lib.ref_slic_gradient(lab.ctypes.data, lab.shape[0], lab.shape[1], out.ctypes.data) == 0
    return out
