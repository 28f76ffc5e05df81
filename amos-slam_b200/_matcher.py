"""ORBmatcher / Frame-view bindings over the C ABI (include/orbx_b200.h).

:class:`ORBmatcher` mirrors ORB_SLAM2::ORBmatcher (/root/reference/include/ORBmatcher.h:57-215): same method names,
argument meaning and return values (number of matches), with the Frame / MapPoint object graph flattened into arrays
(:class:`FrameView` carries what the matchers read from a Frame: mvKeysUn, mDescriptors, mvuRight, the image bounds and the
64x48 grid constants)."""
import ctypes as C

import numpy as np


class _FrameViewC(C.Structure):
    _fields_ = [("n", C.c_int), ("keys_un", C.c_void_p), ("descriptors", C.c_void_p), ("u_right", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
                ("grid_element_width_inv", C.c_float), ("grid_element_height_inv", C.c_float),
                ("nlevels", C.c_int), ("scale_factors", C.c_void_p)]


class FrameView:
    """mvKeysUn / mDescriptors / mvuRight / mnMinX.. / mfGridElement*Inv / mvScaleFactors of a Frame
    (/root/reference/include/Frame.h; grid constants as computed at /root/reference/src/Frame.cc:219-220)."""

    def __init__(self, keys_un, descriptors, width, height, scale_factors, u_right=None, min_x=0.0, min_y=0.0):
        from . import KP_DTYPE
        self.keys = np.ascontiguousarray(keys_un, KP_DTYPE)
        self.desc = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        self.scale_factors = np.ascontiguousarray(scale_factors, np.float32)
        self.u_right = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        self.min_x, self.min_y = np.float32(min_x), np.float32(min_y)
        self.max_x, self.max_y = np.float32(min_x + width), np.float32(min_y + height)
        self.gw_inv = np.float32(64.0) / (self.max_x - self.min_x)       # static_cast<float>(FRAME_GRID_COLS)/(mnMaxX-mnMinX)
        self.gh_inv = np.float32(48.0) / (self.max_y - self.min_y)

    def c(self):
        v = _FrameViewC()
        v.n = len(self.keys); v.keys_un = self.keys.ctypes.data; v.descriptors = self.desc.ctypes.data
        v.u_right = self.u_right.ctypes.data if self.u_right is not None else None
        v.min_x, v.min_y, v.max_x, v.max_y = float(self.min_x), float(self.min_y), float(self.max_x), float(self.max_y)
        v.grid_element_width_inv, v.grid_element_height_inv = float(self.gw_inv), float(self.gh_inv)
        v.nlevels = len(self.scale_factors); v.scale_factors = self.scale_factors.ctypes.data
        return v


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, np.uint8)


class ORBmatcher:
    TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        from . import lib, _check
        self._lib, self._check = lib(), _check
        if not hasattr(self._lib, "orbx_matcher_create"):
            raise RuntimeError("liborbx_b200.so was built without the matcher")
        h = C.c_void_p()
        _check(self._lib.orbx_matcher_create(float(nnratio), int(bool(checkOri)), int(device), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.orbx_matcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return self._lib.orbx_matcher_stream(self._h)

    @property
    def launch_count(self):
        return self._lib.orbx_matcher_launch_count(self._h)

    # static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b), vectorised over rows
    def DescriptorDistance(self, a, b):
        a, b = _u8(a).reshape(-1, 32), _u8(b).reshape(-1, 32)
        out = np.zeros(len(a), np.int32)
        self._check(self._lib.orbx_descriptor_distance(self._h, _p(a), _p(b), len(a), _p(out)))
        return out

    # int SearchForInitialization(Frame &F1, Frame &F2, vector<Point2f> &vbPrevMatched, vector<int> &vnMatches12, int windowSize=10)
    def SearchForInitialization(self, F1, F2, vbPrevMatched, windowSize=10):
        prev = np.ascontiguousarray(vbPrevMatched, np.float32).copy()
        m12 = np.zeros(len(F1.keys), np.int32); nm = C.c_int()
        v1, v2 = F1.c(), F2.c()
        self._check(self._lib.orbx_search_for_initialization(self._h, C.byref(v1), C.byref(v2), _p(prev), _p(m12), int(windowSize), C.byref(nm)))
        return nm.value, m12, prev

    # int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
    def SearchByProjectionFrame(self, cur, proj_uv, proj_invz, last_octave, last_angle, mp_desc, valid, mp_observed, cur_occupied, th, forward=False, backward=False, mbf=0.0):
        uv = np.ascontiguousarray(proj_uv, np.float32); iz = np.ascontiguousarray(proj_invz, np.float32)
        lo = np.ascontiguousarray(last_octave, np.int32); la = np.ascontiguousarray(last_angle, np.float32)
        d, va, ob, oc = _u8(mp_desc), _u8(valid), _u8(mp_observed), _u8(cur_occupied)
        cm = np.zeros(len(cur.keys), np.int32); nm = C.c_int(); v = cur.c()
        self._check(self._lib.orbx_search_by_projection_frame(self._h, C.byref(v), len(iz), _p(uv), _p(iz), _p(lo), _p(la), _p(d), _p(va), _p(ob), _p(oc),
                                                              float(th), int(forward), int(backward), float(mbf), _p(cm), C.byref(nm)))
        self.last_raw_match = cm.copy()          # -2 marks entries assigned and then reset by the rotation check
        cm[cm == -2] = -1
        return nm.value, cm

    # int SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th=3)
    def SearchByProjectionPoints(self, F, track_uv, track_ur, track_level, track_view_cos, mp_desc, mp_observed, f_occupied, th=3.0):
        uv = np.ascontiguousarray(track_uv, np.float32); ur = np.ascontiguousarray(track_ur, np.float32)
        lv = np.ascontiguousarray(track_level, np.int32); vc = np.ascontiguousarray(track_view_cos, np.float32)
        d, ob, oc = _u8(mp_desc), _u8(mp_observed), _u8(f_occupied)
        fm = np.zeros(len(F.keys), np.int32); nm = C.c_int(); v = F.c()
        self._check(self._lib.orbx_search_by_projection_points(self._h, C.byref(v), len(lv), _p(uv), _p(ur), _p(lv), _p(vc), _p(d), _p(ob), _p(oc), float(th), _p(fm), C.byref(nm)))
        return nm.value, fm

    # void Frame::ComputeStereoMatches()
    def ComputeStereoMatches(self, extractor_left, extractor_right, keys_left, desc_left, keys_right, desc_right, mb, mbf):
        from . import KP_DTYPE
        kl = np.ascontiguousarray(keys_left, KP_DTYPE); kr = np.ascontiguousarray(keys_right, KP_DTYPE)
        dl, dr = _u8(desc_left), _u8(desc_right)
        ur = np.zeros(len(kl), np.float32); dep = np.zeros(len(kl), np.float32)
        self._check(self._lib.orbx_compute_stereo_matches(self._h, extractor_left._h, extractor_right._h, _p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr),
                                                          float(mb), float(mbf), _p(ur), _p(dep)))
        return ur, dep

    def match_bruteforce_device(self, d_query_ptr, n_query, d_train_ptr, n_train, d_best_idx_ptr, d_best_dist_ptr, d_second_dist_ptr):
        self._check(self._lib.orbx_match_bruteforce_device(self._h, C.c_void_p(d_query_ptr), n_query, C.c_void_p(d_train_ptr), n_train,
                                                           C.c_void_p(d_best_idx_ptr), C.c_void_p(d_best_dist_ptr), C.c_void_p(d_second_dist_ptr)))

    def match_bruteforce_batch_device(self, n_pairs, d_query_ptr, n_query, d_train_ptr, n_train, d_best_idx_ptr, d_best_dist_ptr, d_second_dist_ptr):
        """n_pairs independent frame pairs in one launch (device pointers, asynchronous on .stream)."""
        self._check(self._lib.orbx_match_bruteforce_batch_device(self._h, int(n_pairs), C.c_void_p(d_query_ptr), n_query, C.c_void_p(d_train_ptr), n_train,
                                                                 C.c_void_p(d_best_idx_ptr), C.c_void_p(d_best_dist_ptr), C.c_void_p(d_second_dist_ptr)))
