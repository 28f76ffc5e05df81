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


class Camera(C.Structure):
    """orbx_camera: mK (fx, fy, cx, cy), mDistCoef (k1, k2, p1, p2, k3) and mbf, all float like the reference's CV_32F matrices."""
    _fields_ = [(n, C.c_float) for n in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3", "bf")]

    @classmethod
    def make(cls, fx, fy, cx, cy, k1=0.0, k2=0.0, p1=0.0, p2=0.0, k3=0.0, bf=0.0):
        return cls(fx, fy, cx, cy, k1, k2, p1, p2, k3, bf)


class Frame:
    """Device-resident Frame: mvKeysUn, mDescriptors, mvuRight, mvDepth, image bounds and mGrid stay on the GPU between the extractor
    and the matchers (UndistortKeyPoints, ComputeStereoFromRGBD, AssignFeaturesToGrid: /root/reference/src/Frame.cc:1052-1117,
    1576-1614, 431-461, in the order Frame::CalDyna runs them, :636-645)."""

    def __init__(self, device=0):
        from . import lib, _check
        self._lib, self._check = lib(), _check
        h = C.c_void_p()
        _check(self._lib.orbx_frame_create(int(device), C.byref(h)))
        self._h = h
        self.keys = []                     # len(F.keys) == N for the matcher wrappers

    def close(self):
        if getattr(self, "_h", None):
            self._lib.orbx_frame_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _depth_args(self, depth, rows, cols):
        if depth is None:
            return None, None, 0
        d = np.asarray(depth)
        if d.ndim == 2:                    # the whole CV_32F image; a row-strided view (cv::Mat ROI) is passed as it is
            if d.dtype != np.float32 or d.strides[1] != 4 or d.strides[0] < 4 * cols:
                d = np.ascontiguousarray(d, np.float32)
            assert d.shape == (rows, cols)
            return d, _p(d), d.strides[0]
        d = np.ascontiguousarray(d, np.float32)
        return d, _p(d), 0                 # per-keypoint values gathered by the caller

    def assign(self, extractor, cam, rows, cols, depth=None):
        """From the result the extractor handle still holds on the device (its last __call__ / ProcessDesp)."""
        d, dp, stride = self._depth_args(depth, rows, cols)
        self._check(self._lib.orbx_frame_assign(self._h, extractor._h, C.byref(cam), int(rows), int(cols), dp, stride))
        self.keys = range(self.N)
        return self

    def assign_host(self, keys, descriptors, scale_factors, cam, rows, cols, depth=None):
        from . import KP_DTYPE
        k = np.ascontiguousarray(keys, KP_DTYPE); ds = _u8(descriptors).reshape(-1, 32); sf = np.ascontiguousarray(scale_factors, np.float32)
        d, dp, stride = self._depth_args(depth, rows, cols)
        self._check(self._lib.orbx_frame_assign_host(self._h, _p(k), _p(ds), len(k), len(sf), _p(sf), C.byref(cam), int(rows), int(cols), dp, stride))
        self.keys = range(self.N)
        return self

    # ---- the same steps one by one, as the reference's Frame constructors call them ----
    def take(self, extractor):
        self._check(self._lib.orbx_frame_take(self._h, extractor._h)); self.keys = []
        return self

    def take_host(self, keys, descriptors, scale_factors):
        from . import KP_DTYPE
        k = np.ascontiguousarray(keys, KP_DTYPE); ds = _u8(descriptors).reshape(-1, 32); sf = np.ascontiguousarray(scale_factors, np.float32)
        self._check(self._lib.orbx_frame_take_host(self._h, _p(k), _p(ds), len(k), len(sf), _p(sf))); self.keys = []
        self._n_taken = len(k)
        return self

    def UndistortKeyPoints(self, cam, n):
        from . import KP_DTYPE
        ku = np.zeros(max(n, 1), KP_DTYPE)
        self._check(self._lib.orbx_frame_undistort_keypoints(self._h, C.byref(cam), _p(ku)))
        return ku[:n]

    def ComputeStereoFromRGBD(self, bf, depth, rows, cols, n):
        d, dp, stride = self._depth_args(depth, rows, cols)
        ur = np.zeros(max(n, 1), np.float32); de = np.zeros(max(n, 1), np.float32)
        self._check(self._lib.orbx_frame_compute_stereo_from_rgbd(self._h, float(bf), dp, stride, int(rows), int(cols), _p(ur), _p(de)))
        return ur[:n], de[:n]

    def AssignFeaturesToGrid(self, bounds6, n):
        b = np.ascontiguousarray(bounds6, np.float32); cs = np.zeros(64 * 48 + 1, np.int32); en = np.zeros(max(n, 1), np.int32)
        self._check(self._lib.orbx_frame_assign_features_to_grid(self._h, _p(b), _p(cs), _p(en)))
        self.keys = range(self.N)
        return cs, en[:cs[-1]]

    def set_stereo(self, u_right, depth):
        ur = np.ascontiguousarray(u_right, np.float32); dp = np.ascontiguousarray(depth, np.float32)
        n = self._lib.orbx_frame_taken(self._h)            # allowed between UndistortKeyPoints and AssignFeaturesToGrid (stereo constructor order)
        assert len(ur) == n and len(dp) == n
        self._check(self._lib.orbx_frame_set_stereo(self._h, _p(ur), _p(dp)))

    @property
    def N(self):
        return self._lib.orbx_frame_size(self._h)

    def read(self):
        """(mvKeysUn, mvuRight, mvDepth, bounds[6] = mnMinX, mnMaxX, mnMinY, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv)."""
        from . import KP_DTYPE
        n = self.N
        ku = np.zeros(max(n, 1), KP_DTYPE); ur = np.zeros(max(n, 1), np.float32); dp = np.zeros(max(n, 1), np.float32); b = np.zeros(6, np.float32)
        self._check(self._lib.orbx_frame_read(self._h, _p(ku), _p(ur), _p(dp), _p(b)))
        return ku[:n], ur[:n], dp[:n], b

    def grid(self):
        """mGrid flattened: cell (x, y) = entries[cell_start[x * 48 + y] : cell_start[x * 48 + y + 1]]."""
        cs = np.zeros(64 * 48 + 1, np.int32); en = np.zeros(max(self.N, 1), np.int32)
        self._check(self._lib.orbx_frame_grid(self._h, _p(cs), _p(en)))
        return cs, en[:cs[-1]]


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, np.uint8)


class _BowSideC(C.Structure):
    _fields_ = [("n", C.c_int), ("keys", C.c_void_p), ("descriptors", C.c_void_p), ("valid", C.c_void_p),
                ("n_fv", C.c_int), ("fv_nodes", C.c_void_p), ("fv_offsets", C.c_void_p), ("fv_indices", C.c_void_p)]


class ORBVocabulary:
    """ORB_SLAM2::ORBVocabulary (DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>, /root/reference/include/ORBVocabulary.h:40-41) on the
    device, from the node table of ORBvoc.txt: parent[i], is_leaf[i], descriptor[i], weight[i] describe node i + 1."""
    TF_IDF, TF, IDF, BINARY = 0, 1, 2, 3
    L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT = 0, 1, 2, 3, 4, 5

    def __init__(self, k, L, parent, is_leaf, descriptors, weights, weighting=0, scoring=0, device=0):
        from . import lib, _check
        self._lib, self._check = lib(), _check
        pa = np.ascontiguousarray(parent, np.int32); lf = np.ascontiguousarray(is_leaf, np.uint8)
        ds = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32); w = np.ascontiguousarray(weights, np.float64)
        assert len(pa) == len(lf) == len(ds) == len(w)
        h = C.c_void_p()
        _check(self._lib.orbx_vocabulary_create(int(device), int(k), int(L), int(weighting), int(scoring), len(pa), _p(pa), _p(lf), _p(ds), _p(w), C.byref(h)))
        self._h = h

    @classmethod
    def load_text(cls, path, device=0):
        """bool loadFromTextFile(const std::string &filename) (src/System.cc:84)."""
        from . import lib, _check
        self = cls.__new__(cls)
        self._lib, self._check = lib(), _check
        h = C.c_void_p()
        _check(self._lib.orbx_vocabulary_load_text(int(device), str(path).encode(), C.byref(h)))
        self._h = h
        return self

    def close(self):
        if getattr(self, "_h", None):
            self._lib.orbx_vocabulary_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        return self._lib.orbx_vocabulary_words(self._h)

    # void transform(const vector<TDescriptor>& features, BowVector &v, FeatureVector &fv, int levelsup) const
    def transform(self, descriptors, levelsup=4):
        """-> dict(word, node: per feature; bow_ids, bow_vals: mBowVec in map order; fv_nodes, fv_offsets, fv_idx: mFeatVec)."""
        d = _u8(descriptors).reshape(-1, 32); n = len(d); m = max(n, 1)
        word = np.zeros(m, np.int32); node = np.zeros(m, np.int32); bi = np.zeros(m, np.int32); bv = np.zeros(m, np.float64)
        fn = np.zeros(m, np.int32); fo = np.zeros(m + 1, np.int32); fi = np.zeros(m, np.int32); nb, nf = C.c_int(), C.c_int()
        self._check(self._lib.orbx_vocabulary_transform(self._h, _p(d), n, int(levelsup), _p(word), _p(node), _p(bi), _p(bv), C.byref(nb), _p(fn), _p(fo), _p(fi), C.byref(nf)))
        return dict(word=word[:n], node=node[:n], bow_ids=bi[:nb.value], bow_vals=bv[:nb.value], fv_nodes=fn[:nf.value], fv_offsets=fo[:nf.value + 1], fv_idx=fi[:fo[nf.value]])


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
        if isinstance(F1, Frame):          # device-resident frames: no upload, no grid build
            self._check(self._lib.orbx_search_for_initialization_frames(self._h, F1._h, F2._h, _p(prev), _p(m12), int(windowSize), C.byref(nm)))
            return nm.value, m12, prev
        v1, v2 = F1.c(), F2.c()
        self._check(self._lib.orbx_search_for_initialization(self._h, C.byref(v1), C.byref(v2), _p(prev), _p(m12), int(windowSize), C.byref(nm)))
        return nm.value, m12, prev

    def ComputeStereoMatchesBatch(self, ext_left, ext_right, B, cap, mb, mbf):
        """Frame::ComputeStereoMatches for the B stereo pairs of the two extractors' last batched extract call -> (mvuRight[B, cap], mvDepth[B, cap])."""
        ur = np.zeros((B, cap), np.float32); dep = np.zeros((B, cap), np.float32)
        self._check(self._lib.orbx_compute_stereo_matches_batch(self._h, ext_left._h, ext_right._h, int(B), int(cap), float(mb), float(mbf), _p(ur), _p(dep)))
        return ur, dep

    def SearchForInitializationBatch(self, F1s, F2s, prevs, windowSize=10):
        """P independent frame pairs in one call (config C2's shard unit).  F1s / F2s: lists of FrameView (host) or Frame (device);
        prevs: list of (n1, 2) float arrays.  Returns (nmatches[P], [matches12], [prev])."""
        P = len(F1s)
        prev = [np.ascontiguousarray(v, np.float32).copy() for v in prevs]
        m12 = [np.zeros(len(f.keys), np.int32) for f in F1s]
        nm = np.zeros(max(P, 1), np.int32)
        pp = (C.c_void_p * max(P, 1))(*[v.ctypes.data for v in prev]); mp = (C.c_void_p * max(P, 1))(*[v.ctypes.data for v in m12])
        if P and isinstance(F1s[0], Frame):
            a1 = (C.c_void_p * P)(*[f._h.value if hasattr(f._h, "value") else f._h for f in F1s]); a2 = (C.c_void_p * P)(*[f._h.value if hasattr(f._h, "value") else f._h for f in F2s])
            self._check(self._lib.orbx_search_for_initialization_frames_batch(self._h, P, a1, a2, pp, mp, int(windowSize), _p(nm)))
        else:
            v1 = (_FrameViewC * max(P, 1))(*[f.c() for f in F1s]); v2 = (_FrameViewC * max(P, 1))(*[f.c() for f in F2s])
            self._check(self._lib.orbx_search_for_initialization_batch(self._h, P, v1, v2, pp, mp, int(windowSize), _p(nm)))
        return nm[:P], m12, prev

    # int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
    def SearchByProjectionFrame(self, cur, proj_uv, proj_invz, last_octave, last_angle, mp_desc, valid, mp_observed, cur_occupied, th, forward=False, backward=False, mbf=0.0):
        uv = np.ascontiguousarray(proj_uv, np.float32); iz = np.ascontiguousarray(proj_invz, np.float32)
        lo = np.ascontiguousarray(last_octave, np.int32); la = np.ascontiguousarray(last_angle, np.float32)
        d, va, ob, oc = _u8(mp_desc), _u8(valid), _u8(mp_observed), _u8(cur_occupied)
        cm = np.zeros(len(cur.keys), np.int32); nm = C.c_int()
        if isinstance(cur, Frame):
            fn, fa = self._lib.orbx_search_by_projection_frame_dev, cur._h
        else:
            v = cur.c(); fn, fa = self._lib.orbx_search_by_projection_frame, C.byref(v)
        self._check(fn(self._h, fa, len(iz), _p(uv), _p(iz), _p(lo), _p(la), _p(d), _p(va), _p(ob), _p(oc),
                                                              float(th), int(forward), int(backward), float(mbf), _p(cm), C.byref(nm)))
        self.last_raw_match = cm.copy()          # -2 marks entries assigned and then reset by the rotation check
        cm[cm == -2] = -1
        return nm.value, cm

    # the same with the projection on the device: world points + pose of CurrentFrame instead of (u, v, 1/z)  (ORBmatcher.cc:1597-1623)
    def SearchByProjectionFramePose(self, cur, world_xyz, has_point, Rcw, tcw, fx, fy, cx, cy, last_octave, last_angle, mp_desc, mp_observed, cur_occupied, th,
                                    forward=False, backward=False, mbf=0.0, taps=False):
        w = np.ascontiguousarray(world_xyz, np.float32).reshape(-1, 3); hp = _u8(has_point)
        R = np.ascontiguousarray(Rcw, np.float32).reshape(9); t = np.ascontiguousarray(tcw, np.float32).reshape(3)
        lo = np.ascontiguousarray(last_octave, np.int32); la = np.ascontiguousarray(last_angle, np.float32)
        d, ob, oc = _u8(mp_desc), _u8(mp_observed), _u8(cur_occupied)
        cm = np.zeros(len(cur.keys), np.int32); nm = C.c_int()
        uv = np.zeros((len(w), 2), np.float32); iz = np.zeros(len(w), np.float32); va = np.zeros(len(w), np.uint8)
        if isinstance(cur, Frame):
            view, dev = None, cur._h
        else:
            v = cur.c(); view, dev = C.byref(v), None
        self._check(self._lib.orbx_search_by_projection_frame_pose(self._h, view, dev, len(w), _p(w), _p(hp), _p(R), _p(t), float(fx), float(fy), float(cx), float(cy),
                                                                   _p(lo), _p(la), _p(d), _p(ob), _p(oc), float(th), int(forward), int(backward), float(mbf), _p(cm), C.byref(nm),
                                                                   _p(uv) if taps else None, _p(iz) if taps else None, _p(va) if taps else None))
        self.last_raw_match = cm.copy()
        cm[cm == -2] = -1
        return (nm.value, cm, uv, iz, va) if taps else (nm.value, cm)

    # int SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, const float th, const int ORBdist)
    def SearchByProjectionKeyFrame(self, cur, proj_uv, predicted_level, kf_angle, mp_desc, valid, cur_occupied, th, ORBdist):
        uv = np.ascontiguousarray(proj_uv, np.float32); lv = np.ascontiguousarray(predicted_level, np.int32); an = np.ascontiguousarray(kf_angle, np.float32)
        d, va, oc = _u8(mp_desc), _u8(valid), _u8(cur_occupied)
        cm = np.zeros(max(len(cur.keys), 1), np.int32); nm = C.c_int()
        if isinstance(cur, Frame):
            fn, fa = self._lib.orbx_search_by_projection_keyframe_dev, cur._h
        else:
            v = cur.c(); fn, fa = self._lib.orbx_search_by_projection_keyframe, C.byref(v)
        self._check(fn(self._h, fa, len(lv), _p(uv), _p(lv), _p(an), _p(d), _p(va), _p(oc), float(th), int(ORBdist), _p(cm), C.byref(nm)))
        cm = cm[:len(cur.keys)]
        self.last_raw_match = cm.copy()
        cm[cm == -2] = -1
        return nm.value, cm

    # int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, vector<MapPoint*> &vpMatched, int th)
    def SearchByProjectionKeyFramePoints(self, kf, proj_uv, predicted_level, mp_desc, valid, kf_matched, th):
        uv = np.ascontiguousarray(proj_uv, np.float32); lv = np.ascontiguousarray(predicted_level, np.int32)
        d, va, km = _u8(mp_desc), _u8(valid), _u8(kf_matched)
        out = np.zeros(max(len(kf.keys), 1), np.int32); nm = C.c_int()
        if isinstance(kf, Frame):
            fn, fa = self._lib.orbx_search_by_projection_keyframe_points_dev, kf._h
        else:
            v = kf.c(); fn, fa = self._lib.orbx_search_by_projection_keyframe_points, C.byref(v)
        self._check(fn(self._h, fa, len(lv), _p(uv), _p(lv), _p(d), _p(va), _p(km), float(th), _p(out), C.byref(nm)))
        return nm.value, out[:len(kf.keys)]

    # the search of int Fuse(KeyFrame *pKF, const vector<MapPoint*> &vpMapPoints, th) / Fuse(KeyFrame *pKF, cv::Mat Scw, vpPoints, th, vpReplacePoint)
    def FuseSearch(self, kf, proj_uv, proj_ur, predicted_level, mp_desc, valid, inv_level_sigma2, th):
        uv = np.ascontiguousarray(proj_uv, np.float32); lv = np.ascontiguousarray(predicted_level, np.int32); d, va = _u8(mp_desc), _u8(valid)
        ur = None if proj_ur is None else np.ascontiguousarray(proj_ur, np.float32); isg = np.ascontiguousarray(inv_level_sigma2, np.float32)
        best = np.zeros(max(len(lv), 1), np.int32); v = kf.c()
        self._check(self._lib.orbx_fuse_search(self._h, C.byref(v), len(lv), _p(uv), _p(ur), _p(lv), _p(d), _p(va), _p(isg), float(th), _p(best)))
        return best[:len(lv)]

    # void MapPoint::ComputeDistinctiveDescriptors(), batched over map points
    def ComputeDistinctiveDescriptors(self, offsets, descriptors):
        off = np.ascontiguousarray(offsets, np.int32); d = _u8(descriptors).reshape(-1, 32)
        best = np.zeros(max(len(off) - 1, 1), np.int32)
        self._check(self._lib.orbx_distinctive_descriptors(self._h, len(off) - 1, _p(off), _p(d), _p(best)))
        return best[:len(off) - 1]

    # int SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12, const float &s12, const cv::Mat &R12, const cv::Mat &t12, const float th)
    def SearchBySim3(self, kf1, kf2, uv1, lvl1, desc1, valid1, uv2, lvl2, desc2, valid2, th):
        a = [np.ascontiguousarray(uv1, np.float32), np.ascontiguousarray(lvl1, np.int32), _u8(desc1), _u8(valid1),
             np.ascontiguousarray(uv2, np.float32), np.ascontiguousarray(lvl2, np.int32), _u8(desc2), _u8(valid2)]
        m12 = np.zeros(max(len(kf1.keys), 1), np.int32); nf = C.c_int()
        v1, v2 = kf1.c(), kf2.c()
        self._check(self._lib.orbx_search_by_sim3(self._h, C.byref(v1), C.byref(v2), *[_p(x) for x in a], float(th), _p(m12), C.byref(nf)))
        return nf.value, m12[:len(kf1.keys)]

    # int SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th=3)
    def SearchByProjectionPoints(self, F, track_uv, track_ur, track_level, track_view_cos, mp_desc, mp_observed, f_occupied, th=3.0):
        uv = np.ascontiguousarray(track_uv, np.float32); ur = np.ascontiguousarray(track_ur, np.float32)
        lv = np.ascontiguousarray(track_level, np.int32); vc = np.ascontiguousarray(track_view_cos, np.float32)
        d, ob, oc = _u8(mp_desc), _u8(mp_observed), _u8(f_occupied)
        fm = np.zeros(len(F.keys), np.int32); nm = C.c_int()
        if isinstance(F, Frame):
            fn, fa = self._lib.orbx_search_by_projection_points_dev, F._h
        else:
            v = F.c(); fn, fa = self._lib.orbx_search_by_projection_points, C.byref(v)
        self._check(fn(self._h, fa, len(lv), _p(uv), _p(ur), _p(lv), _p(vc), _p(d), _p(ob), _p(oc), float(th), _p(fm), C.byref(nm)))
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

    # int SearchByBoW(KeyFrame* pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches)            (kf_kf = False)
    # int SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12)           (kf_kf = True)
    def SearchByBoW(self, kf_kf, keys1, desc1, valid1, fv1, keys2, desc2, valid2, fv2):
        """fv = dict with fv_nodes / fv_offsets / fv_idx as ORBVocabulary.transform returns.  -> (nmatches, match12, match21)."""
        from . import KP_DTYPE
        hold = []

        def side(keys, desc, valid, fv):
            k = np.ascontiguousarray(keys, KP_DTYPE); d = _u8(desc).reshape(-1, 32); v = _u8(valid)
            a = [np.ascontiguousarray(fv[x], np.int32) for x in ("fv_nodes", "fv_offsets", "fv_idx")]
            hold.extend([k, d, v] + a)
            s = _BowSideC(); s.n = len(k); s.keys = k.ctypes.data; s.descriptors = d.ctypes.data; s.valid = v.ctypes.data if v is not None else None
            s.n_fv = len(a[0]); s.fv_nodes = a[0].ctypes.data; s.fv_offsets = a[1].ctypes.data; s.fv_indices = a[2].ctypes.data
            return s
        s1, s2 = side(keys1, desc1, valid1, fv1), side(keys2, desc2, valid2, fv2)
        m12 = np.zeros(max(s1.n, 1), np.int32); m21 = np.zeros(max(s2.n, 1), np.int32); nm = C.c_int()
        self._check(self._lib.orbx_search_by_bow(self._h, int(bool(kf_kf)), C.byref(s1), C.byref(s2), _p(m12), _p(m21), C.byref(nm)))
        return nm.value, m12[:s1.n], m21[:s2.n]

    # int SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12, vector<pair<size_t,size_t>> &vMatchedPairs, const bool bOnlyStereo)
    def SearchForTriangulation(self, keys1, desc1, free1, ur1, fv1, keys2, desc2, free2, ur2, fv2, F12, epipole, scale_factors2, level_sigma2_2, bOnlyStereo=False):
        """free = the feature holds no map point; epipole = (ex, ey) as the caller computes it (:818-825).  -> (nmatches, match12)."""
        from . import KP_DTYPE
        hold = []

        def side(keys, desc, valid, fv):
            k = np.ascontiguousarray(keys, KP_DTYPE); d = _u8(desc).reshape(-1, 32); v = _u8(valid)
            a = [np.ascontiguousarray(fv[x], np.int32) for x in ("fv_nodes", "fv_offsets", "fv_idx")]
            hold.extend([k, d, v] + a)
            s = _BowSideC(); s.n = len(k); s.keys = k.ctypes.data; s.descriptors = d.ctypes.data; s.valid = v.ctypes.data
            s.n_fv = len(a[0]); s.fv_nodes = a[0].ctypes.data; s.fv_offsets = a[1].ctypes.data; s.fv_indices = a[2].ctypes.data
            return s
        s1, s2 = side(keys1, desc1, free1, fv1), side(keys2, desc2, free2, fv2)
        u1 = np.ascontiguousarray(ur1, np.float32); u2 = np.ascontiguousarray(ur2, np.float32)
        F = np.ascontiguousarray(F12, np.float32).reshape(9); sc = np.ascontiguousarray(scale_factors2, np.float32); sg = np.ascontiguousarray(level_sigma2_2, np.float32)
        m12 = np.zeros(max(s1.n, 1), np.int32); nm = C.c_int()
        self._check(self._lib.orbx_search_for_triangulation(self._h, C.byref(s1), C.byref(s2), _p(u1), _p(u2), _p(F), float(epipole[0]), float(epipole[1]), len(sc), _p(sc), _p(sg),
                                                            int(bool(bOnlyStereo)), _p(m12), C.byref(nm)))
        return nm.value, m12[:s1.n]

    # vector<size_t> Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel) for many windows at once, on a device-resident Frame
    def GetFeaturesInArea(self, F, xy, r, min_level=None, max_level=None):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2); nq = len(xy)
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(r, np.float32), (nq,)), np.float32)
        mn = np.ascontiguousarray(np.broadcast_to(np.asarray(-1 if min_level is None else min_level, np.int32), (nq,)), np.int32)
        mx = np.ascontiguousarray(np.broadcast_to(np.asarray(-1 if max_level is None else max_level, np.int32), (nq,)), np.int32)
        off = np.zeros(nq + 1, np.int32); tot = C.c_int()
        cap = max(64 * nq, 1024)
        while True:
            idx = np.zeros(cap, np.int32)
            rc = self._lib.orbx_frame_features_in_area(self._h, F._h, nq, _p(xy), _p(r), _p(mn), _p(mx), _p(off), _p(idx), cap, C.byref(tot))
            if rc == -3 and tot.value > cap:           # ORBX_E_CAPACITY: grow and repeat
                cap = tot.value; continue
            self._check(rc)
            return [idx[off[q]:off[q + 1]].copy() for q in range(nq)]

    def match_bruteforce_device(self, d_query_ptr, n_query, d_train_ptr, n_train, d_best_idx_ptr, d_best_dist_ptr, d_second_dist_ptr):
        self._check(self._lib.orbx_match_bruteforce_device(self._h, C.c_void_p(d_query_ptr), n_query, C.c_void_p(d_train_ptr), n_train,
                                                           C.c_void_p(d_best_idx_ptr), C.c_void_p(d_best_dist_ptr), C.c_void_p(d_second_dist_ptr)))

    def match_bruteforce_batch_device(self, n_pairs, d_query_ptr, n_query, d_train_ptr, n_train, d_best_idx_ptr, d_best_dist_ptr, d_second_dist_ptr):
        """n_pairs independent frame pairs in one launch (device pointers, asynchronous on .stream)."""
        self._check(self._lib.orbx_match_bruteforce_batch_device(self._h, int(n_pairs), C.c_void_p(d_query_ptr), n_query, C.c_void_p(d_train_ptr), n_train,
                                                                 C.c_void_p(d_best_idx_ptr), C.c_void_p(d_best_dist_ptr), C.c_void_p(d_second_dist_ptr)))
