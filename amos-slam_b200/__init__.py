"""amos-slam_b200 -- B200-native (sm_100a) drop-in for the ORB feature front-end of Amos-SLAM / ORB-SLAM2.

Python face of the C-ABI shared library ``liborbx_b200.so`` (declared in include/orbx_b200.h).  The classes
mirror the reference's C++ interface for this path -- same names, argument meaning and error behaviour:

* :class:`ORBextractor`  <- ORB_SLAM2::ORBextractor   (/root/reference/include/ORBextractor.h:93-168)
* :class:`ORBmatcher`    <- ORB_SLAM2::ORBmatcher     (/root/reference/include/ORBmatcher.h:57-215)

There is NO CPU fallback: importing works anywhere, but creating an extractor / matcher requires the CUDA
library and a B200; a missing library raises immediately.  (The directory name has a hyphen, so import it
with ``importlib.import_module("amos-slam_b200")``.)
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORBX_LIB_PATH") or os.path.join(HERE, "liborbx_b200.so")   # (ORBX_LIB_PATH: A/B builds of the same library while tuning a kernel)
CSRC = os.path.join(HERE, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177"]
SOURCES = ["orbx_extractor.cu", "orbx_matcher.cu", "orbx_pool.cu", "orbx_slic.cu", "host_pack.cpp"]

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])

OK, E_INVALID, E_CUDA, E_CAPACITY, E_STATE, E_OVERFLOW = 0, -1, -2, -3, -4, -5
TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30
FRAME_GRID_COLS, FRAME_GRID_ROWS = 64, 48


class LabelsC(C.Structure):
    """orbx_labels (include/orbx_b200.h): super-pixel ids + per-frame flag tables of the batched MovingKeyPoints."""
    _fields_ = [("labels", C.c_void_p), ("label_step", C.c_size_t), ("label_frame_stride", C.c_size_t), ("flagged", C.c_void_p), ("n_labels", C.c_int)]


def label_flags(centers_id, rm_vector):
    """flagged[id - 1] = (rm_vector[centers[id - 1].id] == 1)  (src/ORBextractor.cc:1727); out-of-range ids count as not flagged."""
    cid = np.asarray(centers_id, np.int64); rm = np.asarray(rm_vector, np.int64)
    ok = (cid >= 0) & (cid < len(rm))
    out = np.zeros(len(cid), np.uint8)
    out[ok] = (rm[cid[ok]] == 1)
    return out


class OrbxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("orbx error %d: %s" % (code, msg))
        self.code = code


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into the in-tree liborbx_b200.so (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "orbx_b200.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + srcs
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    """Load the C-ABI library.  Raises if it has not been built: the product has no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OrbxError(E_CUDA, "liborbx_b200.so is missing (run __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, ci, cf, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    L.orbx_last_error.restype = C.c_char_p
    L.orbx_create.argtypes = [ci, cf, ci, ci, ci, ci, C.POINTER(vp)]
    L.orbx_destroy.argtypes = [vp]; L.orbx_destroy.restype = None
    L.orbx_get_levels.argtypes = [vp]
    L.orbx_get_scale_factor.argtypes = [vp]; L.orbx_get_scale_factor.restype = cf
    for name in ("orbx_get_scale_factors", "orbx_get_inverse_scale_factors", "orbx_get_scale_sigma_squares",
                 "orbx_get_inverse_scale_sigma_squares", "orbx_get_features_per_level"):
        getattr(L, name).argtypes = [vp, vp]
    L.orbx_max_keypoints.argtypes = [vp, ci, ci]
    L.orbx_stream.argtypes = [vp]; L.orbx_stream.restype = vp
    L.orbx_launch_count.argtypes = [vp]; L.orbx_launch_count.restype = C.c_longlong
    L.orbx_check_overflow.argtypes = [vp]
    L.orbx_profile_enable.argtypes = [vp, ci]
    L.orbx_profile_collect.argtypes = [vp, vp, C.POINTER(ci)]
    L.orbx_extract.argtypes = [vp, vp, ci, ci, sz, vp, vp, ci, C.POINTER(ci)]
    L.orbx_detect.argtypes = [vp, vp, ci, ci, sz, vp, vp, ci, C.POINTER(ci)]
    L.orbx_cull.argtypes = [vp, vp, sz, vp, sz, ci, ci, vp, ci, vp, ci, vp, vp, vp, C.POINTER(ci)]
    L.orbx_describe.argtypes = [vp, vp, vp, vp, vp, ci, C.POINTER(ci)]
    L.orbx_pyramid_level.argtypes = [vp, ci, ci, vp, sz, C.POINTER(ci), C.POINTER(ci)]
    L.orbx_extract_batch.argtypes = [vp, vp, ci, ci, ci, sz, sz, vp, vp, ci, vp]
    L.orbx_extract_batch_device.argtypes = [vp, vp, ci, ci, ci, sz, sz, vp, vp, ci, vp]
    L.orbx_extract_masked_batch.argtypes = [vp, vp, vp, ci, ci, ci, sz, sz, sz, sz, vp, vp, ci, vp, vp]
    L.orbx_extract_masked_batch_device.argtypes = [vp, vp, vp, ci, ci, ci, sz, sz, sz, sz, vp, vp, ci, vp, vp]
    L.orbx_extract_masked_batch_labels.argtypes = [vp, vp, vp, vp, ci, ci, ci, sz, sz, sz, sz, vp, vp, ci, vp, vp]
    L.orbx_extract_masked_batch_labels_device.argtypes = [vp, vp, vp, vp, ci, ci, ci, sz, sz, sz, sz, vp, vp, ci, vp, vp]
    L.orbx_slic_create.argtypes = [ci, C.POINTER(vp)]
    L.orbx_slic_destroy.argtypes = [vp]; L.orbx_slic_destroy.restype = None
    L.orbx_slic_run.argtypes = [vp, vp, sz, vp, sz, ci, ci, ci, ci, ci, vp, sz, vp, sz, vp, ci, C.POINTER(ci)]
    L.orbx_pool_shard_of.argtypes = [ci, ci, ci, C.POINTER(ci), C.POINTER(ci)]
    L.orbx_pool_create.argtypes = [ci, cf, ci, ci, ci, ci, vp, ci, C.POINTER(vp)]
    L.orbx_pool_destroy.argtypes = [vp]; L.orbx_pool_destroy.restype = None
    L.orbx_pool_gpus.argtypes = [vp]; L.orbx_pool_streams_per_gpu.argtypes = [vp]; L.orbx_pool_device_of.argtypes = [vp, ci]
    L.orbx_pool_frames_done.argtypes = [vp, ci, ci]; L.orbx_pool_frames_done.restype = C.c_longlong
    L.orbx_pool_submit.argtypes = [vp, ci, vp, ci, ci, sz, vp, vp, ci, vp, C.POINTER(C.c_longlong)]
    L.orbx_pool_submit_batch.argtypes = [vp, ci, vp, vp, vp, ci, ci, ci, sz, sz, sz, sz, vp, vp, ci, vp, vp, C.POINTER(C.c_longlong)]
    L.orbx_pool_wait.argtypes = [vp, C.c_longlong]; L.orbx_pool_wait_all.argtypes = [vp]
    L.orbx_debug_level_candidates.argtypes = [vp, ci, ci, vp, ci, C.POINTER(ci)]
    L.orbx_debug_blurred_level.argtypes = [vp, ci, ci, vp, sz]
    L.orbx_debug_pyramid_level.argtypes = [vp, ci, ci, vp, sz]
    L.orbx_debug_distribute.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, vp, ci, C.POINTER(ci)]
    L.orbx_debug_sincos.argtypes = [vp, C.c_uint, ci, vp, vp]
    L.orbx_debug_canary_check.argtypes = [C.POINTER(ci)]
    if hasattr(L, "orbx_matcher_create"):
        L.orbx_matcher_create.argtypes = [cf, ci, ci, C.POINTER(vp)]
        L.orbx_matcher_destroy.argtypes = [vp]; L.orbx_matcher_destroy.restype = None
        L.orbx_matcher_stream.argtypes = [vp]; L.orbx_matcher_stream.restype = vp
        L.orbx_matcher_launch_count.argtypes = [vp]; L.orbx_matcher_launch_count.restype = C.c_longlong
        L.orbx_matcher_profile_enable.argtypes = [vp, ci]
        L.orbx_matcher_profile_collect.argtypes = [vp, ci, vp, C.POINTER(ci)]
        L.orbx_descriptor_distance.argtypes = [vp, vp, vp, ci, vp]
        L.orbx_search_for_initialization.argtypes = [vp, vp, vp, vp, vp, ci, C.POINTER(ci)]
        L.orbx_search_by_projection_frame.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, cf, ci, ci, cf, vp, C.POINTER(ci)]
        L.orbx_search_by_projection_frame_pose.argtypes = [vp, vp, vp, ci, vp, vp, vp, vp, cf, cf, cf, cf, vp, vp, vp, vp, vp, cf, ci, ci, cf, vp, C.POINTER(ci), vp, vp, vp]
        L.orbx_search_by_projection_points.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, cf, vp, C.POINTER(ci)]
        L.orbx_search_by_projection_keyframe.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, vp, cf, ci, vp, C.POINTER(ci)]
        L.orbx_search_by_projection_keyframe_points.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, cf, vp, C.POINTER(ci)]
        L.orbx_search_by_projection_keyframe_points_dev.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, cf, vp, C.POINTER(ci)]
        L.orbx_fuse_search.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, vp, cf, vp]
        L.orbx_distinctive_descriptors.argtypes = [vp, ci, vp, vp, vp]
        L.orbx_search_by_sim3.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, cf, vp, C.POINTER(ci)]
        L.orbx_search_by_projection_keyframe_dev.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, vp, cf, ci, vp, C.POINTER(ci)]
        L.orbx_compute_stereo_matches.argtypes = [vp, vp, vp, vp, vp, ci, vp, vp, ci, cf, cf, vp, vp]
        L.orbx_compute_stereo_matches_batch.argtypes = [vp, vp, vp, ci, ci, cf, cf, vp, vp]
        L.orbx_compute_stereo_matches_batch_device.argtypes = [vp, vp, vp, ci, ci, cf, cf, vp, vp]
        L.orbx_match_bruteforce_device.argtypes = [vp, vp, ci, vp, ci, vp, vp, vp]
        L.orbx_match_bruteforce_batch_device.argtypes = [vp, ci, vp, ci, vp, ci, vp, vp, vp]
        L.orbx_vocabulary_create.argtypes = [ci, ci, ci, ci, ci, ci, vp, vp, vp, vp, C.POINTER(vp)]
        L.orbx_vocabulary_destroy.argtypes = [vp]; L.orbx_vocabulary_destroy.restype = None
        L.orbx_vocabulary_words.argtypes = [vp]
        L.orbx_vocabulary_load_text.argtypes = [ci, C.c_char_p, C.POINTER(vp)]
        L.orbx_vocabulary_transform.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, C.POINTER(ci), vp, vp, vp, C.POINTER(ci)]
        L.orbx_search_by_bow.argtypes = [vp, ci, vp, vp, vp, vp, C.POINTER(ci)]
        L.orbx_search_for_triangulation.argtypes = [vp, vp, vp, vp, vp, vp, cf, cf, ci, vp, vp, ci, vp, C.POINTER(ci)]
        L.orbx_frame_create.argtypes = [ci, C.POINTER(vp)]
        L.orbx_frame_destroy.argtypes = [vp]; L.orbx_frame_destroy.restype = None
        L.orbx_frame_assign.argtypes = [vp, vp, vp, ci, ci, vp, sz]
        L.orbx_frame_assign_host.argtypes = [vp, vp, vp, ci, ci, vp, vp, ci, ci, vp, sz]
        L.orbx_frame_set_stereo.argtypes = [vp, vp, vp]
        L.orbx_frame_take.argtypes = [vp, vp]
        L.orbx_frame_take_host.argtypes = [vp, vp, vp, ci, ci, vp]
        L.orbx_frame_undistort_keypoints.argtypes = [vp, vp, vp]
        L.orbx_frame_compute_stereo_from_rgbd.argtypes = [vp, cf, vp, sz, ci, ci, vp, vp]
        L.orbx_frame_assign_features_to_grid.argtypes = [vp, vp, vp, vp]
        L.orbx_frame_size.argtypes = [vp]
        L.orbx_frame_taken.argtypes = [vp]
        L.orbx_frame_read.argtypes = [vp, vp, vp, vp, vp]
        L.orbx_frame_grid.argtypes = [vp, vp, vp]
        L.orbx_frame_features_in_area.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, vp, ci, C.POINTER(ci)]
        L.orbx_search_for_initialization_frames.argtypes = [vp, vp, vp, vp, vp, ci, C.POINTER(ci)]
        L.orbx_search_for_initialization_batch.argtypes = [vp, ci, vp, vp, vp, vp, ci, vp]
        L.orbx_search_for_initialization_frames_batch.argtypes = [vp, ci, vp, vp, vp, vp, ci, vp]
        L.orbx_search_by_projection_frame_dev.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, cf, ci, ci, cf, vp, C.POINTER(ci)]
        L.orbx_search_by_projection_points_dev.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, cf, vp, C.POINTER(ci)]
    _lib = L
    return L


def _check(rc):
    if rc != OK:
        raise OrbxError(rc, lib().orbx_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class ORBextractor:
    """ORB_SLAM2::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) on one B200.

    ``extractor(image, mask=None)`` mirrors ``operator()(image, mask, keypoints, descriptors)`` and returns
    ``(keypoints, descriptors)`` (structured array with cv::KeyPoint's layout, N x 32 uint8); the mask is
    ignored exactly as in the reference (/root/reference/src/ORBextractor.cc:1544-1668).
    """

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0):
        self._lib = lib()
        h = C.c_void_p()
        _check(self._lib.orbx_create(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST), int(device), C.byref(h)))
        self._h = h
        self.nfeatures, self.nlevels, self.device = int(nfeatures), int(nlevels), int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.orbx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters (include/ORBextractor.h:117-165) ----
    def GetLevels(self):
        return self._lib.orbx_get_levels(self._h)

    def GetScaleFactor(self):
        return self._lib.orbx_get_scale_factor(self._h)

    def _vec(self, fn, dtype=np.float32):
        out = np.zeros(self.nlevels, dtype)
        _check(getattr(self._lib, fn)(self._h, _ptr(out)))
        return out

    def GetScaleFactors(self):
        return self._vec("orbx_get_scale_factors")

    def GetInverseScaleFactors(self):
        return self._vec("orbx_get_inverse_scale_factors")

    def GetScaleSigmaSquares(self):
        return self._vec("orbx_get_scale_sigma_squares")

    def GetInverseScaleSigmaSquares(self):
        return self._vec("orbx_get_inverse_scale_sigma_squares")

    def features_per_level(self):
        return self._vec("orbx_get_features_per_level", np.int32)

    def max_keypoints(self, rows, cols):
        rc = self._lib.orbx_max_keypoints(self._h, int(rows), int(cols))
        if rc < 0:
            _check(rc)
        return rc

    @property
    def stream(self):
        """cudaStream_t (int) all of this handle's work is ordered on."""
        return self._lib.orbx_stream(self._h)

    @property
    def launch_count(self):
        return self._lib.orbx_launch_count(self._h)

    def check_overflow(self):
        return self._lib.orbx_check_overflow(self._h)

    STAGES = ("pyr_resize", "fast_cells", "octree_sort", "octree_tree", "gauss7", "orient_describe")

    def profile_enable(self, on=True):
        _check(self._lib.orbx_profile_enable(self._h, 1 if on else 0))

    def profile_collect(self):
        """-> (dict stage -> summed ms, number of profiled calls)"""
        ms = np.zeros(6, np.float64); n = C.c_int()
        _check(self._lib.orbx_profile_collect(self._h, _ptr(ms), C.byref(n)))
        return dict(zip(self.STAGES, ms.tolist())), n.value

    # ---- operator()(image, mask, keypoints, descriptors) ----
    def __call__(self, image, mask=None):
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        image = self._img(image)
        cap = self.max_keypoints(image.shape[0], image.shape[1])
        kp = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); n = C.c_int()
        _check(self._lib.orbx_extract(self._h, _ptr(image), image.shape[0], image.shape[1], image.strides[0], _ptr(kp), _ptr(desc), cap, C.byref(n)))
        return kp[:n.value].copy(), desc[:n.value].copy()

    @staticmethod
    def _img(image):
        if image.dtype != np.uint8 or image.ndim != 2:
            raise OrbxError(E_INVALID, "image must be CV_8UC1 (2-D uint8)")     # assert(image.type()==CV_8UC1) :1559
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        return image

    # ---- operator()(image, mask, vector<vector<KeyPoint>>&): keypoints only, level coordinates ----
    def detect(self, image, mask=None):
        image = self._img(image)
        cap = self.max_keypoints(image.shape[0], image.shape[1])
        kp = np.zeros(cap, KP_DTYPE); counts = np.zeros(self.nlevels, np.int32); n = C.c_int()
        _check(self._lib.orbx_detect(self._h, _ptr(image), image.shape[0], image.shape[1], image.strides[0], _ptr(kp), _ptr(counts), cap, C.byref(n)))
        return kp[:n.value].copy(), counts

    # ---- MovingKeyPoints(imGray, imS, imLS, centers, rm_vector, DynaFlag, mvKeysT) ----
    def MovingKeyPoints(self, imS, imLS, centers_id, rm_vector, keypoints, level_counts):
        imS = np.ascontiguousarray(imS, np.uint8); imLS = np.ascontiguousarray(imLS, np.float64)
        centers_id = np.ascontiguousarray(centers_id, np.int32); rm_vector = np.ascontiguousarray(rm_vector, np.int32)
        kp = np.ascontiguousarray(keypoints, KP_DTYPE).copy(); counts = np.ascontiguousarray(level_counts, np.int32).copy()
        culled = np.zeros(max(len(kp), 1), KP_DTYPE); nc = C.c_int()
        _check(self._lib.orbx_cull(self._h, _ptr(imS), imS.strides[0], _ptr(imLS), imLS.strides[0], imS.shape[0], imS.shape[1],
                                   _ptr(centers_id), len(centers_id), _ptr(rm_vector), len(rm_vector), _ptr(kp), _ptr(counts), _ptr(culled), C.byref(nc)))
        return kp[:int(counts.sum())].copy(), counts, culled[:nc.value].copy()

    # ---- ProcessDesp(image, mask, allKeypoints, mKeypoints, descriptors) ----
    def ProcessDesp(self, keypoints, level_counts):
        kp = np.ascontiguousarray(keypoints, KP_DTYPE); counts = np.ascontiguousarray(level_counts, np.int32)
        cap = max(len(kp), 1)
        out = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); n = C.c_int()
        _check(self._lib.orbx_describe(self._h, _ptr(kp), _ptr(counts), _ptr(out), _ptr(desc), cap, C.byref(n)))
        return out[:n.value].copy(), desc[:n.value].copy()

    # ---- mvImagePyramid[level] ----
    def pyramid_level(self, level, border=0):
        r, c = C.c_int(), C.c_int()
        _check(self._lib.orbx_pyramid_level(self._h, level, border, None, 0, C.byref(r), C.byref(c)))
        out = np.zeros((r.value, c.value), np.uint8)
        _check(self._lib.orbx_pyramid_level(self._h, level, border, _ptr(out), out.strides[0], C.byref(r), C.byref(c)))
        return out

    # ---- batched operator() ----
    def extract_batch(self, images, cap=None):
        """images: (B, rows, cols) uint8 host array.  Returns (kp[B,cap], desc[B,cap,32], counts[B])."""
        if images.dtype != np.uint8 or images.ndim != 3:
            raise OrbxError(E_INVALID, "images must be (B, rows, cols) uint8")
        images = np.ascontiguousarray(images)
        B, rows, cols = images.shape
        cap = cap or self.max_keypoints(rows, cols)
        kp = np.zeros((B, cap), KP_DTYPE); desc = np.zeros((B, cap, 32), np.uint8); counts = np.zeros(B, np.int32)
        _check(self._lib.orbx_extract_batch(self._h, _ptr(images), B, rows, cols, images.strides[1], images.strides[0], _ptr(kp), _ptr(desc), cap, _ptr(counts)))
        return kp, desc, counts

    def extract_masked_batch(self, images, masks, labels=None, flagged=None, cap=None):
        """Batched Amos path: images, masks (B, rows, cols) uint8 host arrays -> (kp[B,cap], desc[B,cap,32], counts[B], culled[B]).
        labels (B, rows, cols) super-pixel ids (1-based; any integer-valued dtype, sent as uint16) and flagged (B, n_labels) uint8 =
        label_flags(centers_id, rm_vector) per frame add the super-pixel term of MovingKeyPoints."""
        if images.dtype != np.uint8 or images.ndim != 3 or masks.dtype != np.uint8 or masks.shape != images.shape:
            raise OrbxError(E_INVALID, "images and masks must be (B, rows, cols) uint8 of the same shape")
        images = np.ascontiguousarray(images); masks = np.ascontiguousarray(masks)
        B, rows, cols = images.shape
        cap = cap or self.max_keypoints(rows, cols)
        kp = np.zeros((B, cap), KP_DTYPE); desc = np.zeros((B, cap, 32), np.uint8); counts = np.zeros(B, np.int32); culled = np.zeros(B, np.int32)
        lab = None
        if labels is not None:
            l16 = np.ascontiguousarray(labels, np.uint16); fl = np.ascontiguousarray(flagged, np.uint8)
            if l16.shape != images.shape or fl.ndim != 2 or fl.shape[0] != B:
                raise OrbxError(E_INVALID, "labels must be (B, rows, cols) and flagged (B, n_labels)")
            lab = LabelsC(_ptr(l16), cols, rows * cols, _ptr(fl), fl.shape[1])
        _check(self._lib.orbx_extract_masked_batch_labels(self._h, _ptr(images), _ptr(masks), C.byref(lab) if lab is not None else None, B, rows, cols,
                                                          images.strides[1], images.strides[0], masks.strides[1], masks.strides[0],
                                                          _ptr(kp), _ptr(desc), cap, _ptr(counts), _ptr(culled)))
        return kp, desc, counts, culled

    def extract_masked_batch_raw_device(self, images_ptr, masks_ptr, B, rows, cols, step, frame_stride, mask_step, mask_frame_stride, kp_ptr, desc_ptr, cap, counts_ptr, culled_ptr=0):
        _check(self._lib.orbx_extract_masked_batch_device(self._h, C.c_void_p(images_ptr), C.c_void_p(masks_ptr), B, rows, cols, step, frame_stride, mask_step,
                                                          mask_frame_stride, C.c_void_p(kp_ptr), C.c_void_p(desc_ptr), cap, C.c_void_p(counts_ptr), C.c_void_p(culled_ptr) if culled_ptr else None))

    def extract_batch_raw(self, images_ptr, B, rows, cols, step, frame_stride, kp_ptr, desc_ptr, cap, counts_ptr, device=False):
        """Raw-pointer form (ints): host pointers (synchronous) or device pointers (asynchronous on .stream)."""
        fn = self._lib.orbx_extract_batch_device if device else self._lib.orbx_extract_batch
        _check(fn(self._h, C.c_void_p(images_ptr), B, rows, cols, step, frame_stride, C.c_void_p(kp_ptr), C.c_void_p(desc_ptr), cap, C.c_void_p(counts_ptr)))

    # ---- stage taps (tests) ----
    def debug_level_candidates(self, b, level):
        n = C.c_int()
        cap = 1 << 20
        out = np.zeros(cap, KP_DTYPE)
        _check(self._lib.orbx_debug_level_candidates(self._h, b, level, _ptr(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def debug_pyramid_level(self, b, level, shape):
        out = np.zeros(shape, np.uint8)
        _check(self._lib.orbx_debug_pyramid_level(self._h, b, level, _ptr(out), out.strides[0]))
        return out

    def debug_blurred_level(self, b, level, shape):
        out = np.zeros(shape, np.uint8)
        _check(self._lib.orbx_debug_blurred_level(self._h, b, level, _ptr(out), out.strides[0]))
        return out

    def debug_sincos(self, lo_bits, n):
        """sin / cos of the rBRIEF rotation as the device evaluates them, for the n floats with bit patterns lo_bits, lo_bits + 1, ..."""
        s = np.zeros(n, np.float32); c = np.zeros(n, np.float32)
        _check(self._lib.orbx_debug_sincos(self._h, int(lo_bits), int(n), _ptr(s), _ptr(c)))
        return s, c

    def debug_distribute(self, cand, minX, maxX, minY, maxY, N):
        cand = np.ascontiguousarray(cand, KP_DTYPE)
        cap = max(4 * N + 64, 256)
        out = np.zeros(cap, KP_DTYPE); n = C.c_int()
        _check(self._lib.orbx_debug_distribute(self._h, _ptr(cand), len(cand), minX, maxX, minY, maxY, N, _ptr(out), cap, C.byref(n)))
        return out[:n.value].copy()


class ExtractorPool:
    """orbx_pool: multi-sequence / multi-GPU driver (gpu = seq_id mod G, stream = (seq_id div G) mod S; one worker thread + extractor
    handle per (gpu, stream)).  Buffers handed to submit* must stay alive until wait*()."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, n_gpus=1, streams_per_gpu=2, devices=None):
        self._lib = lib()
        h = C.c_void_p()
        dv = (C.c_int * n_gpus)(*devices) if devices is not None else None
        _check(self._lib.orbx_pool_create(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST), int(n_gpus), dv, int(streams_per_gpu), C.byref(h)))
        self._h = h; self.n_gpus = n_gpus; self.streams_per_gpu = streams_per_gpu; self._keep = {}

    def close(self):
        if getattr(self, "_h", None):
            self._lib.orbx_pool_destroy(self._h); self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_of(self, seq_id):
        return self._lib.orbx_pool_device_of(self._h, int(seq_id))

    def frames_done(self, gpu, stream):
        return self._lib.orbx_pool_frames_done(self._h, int(gpu), int(stream))

    def submit(self, seq_id, image, cap):
        """One frame of sequence seq_id -> ticket; fetch with result(ticket)."""
        img = np.ascontiguousarray(image, np.uint8)
        kp = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); n = np.zeros(1, np.int32); t = C.c_longlong()
        _check(self._lib.orbx_pool_submit(self._h, int(seq_id), _ptr(img), img.shape[0], img.shape[1], img.strides[0], _ptr(kp), _ptr(desc), cap, _ptr(n), C.byref(t)))
        self._keep[t.value] = (img, kp, desc, n, None)
        return t.value

    def submit_batch(self, seq_id, images, masks=None, labels=None, flagged=None, cap=None, out=None):
        """B frames of sequence seq_id (same geometry) -> ticket.  out = (kp, desc, counts, culled) pre-allocated (pinned) arrays, optional."""
        images = np.ascontiguousarray(images, np.uint8); B, rows, cols = images.shape
        if out is None:
            out = (np.zeros((B, cap), KP_DTYPE), np.zeros((B, cap, 32), np.uint8), np.zeros(B, np.int32), np.zeros(B, np.int32))
        kp, desc, counts, culled = out
        cap = kp.shape[1]
        mk = None if masks is None else np.ascontiguousarray(masks, np.uint8)
        lab = None; keep = None
        if labels is not None:
            l16 = np.ascontiguousarray(labels, np.uint16); fl = np.ascontiguousarray(flagged, np.uint8)
            lab = LabelsC(_ptr(l16), cols, rows * cols, _ptr(fl), fl.shape[1]); keep = (l16, fl, lab)
        t = C.c_longlong()
        _check(self._lib.orbx_pool_submit_batch(self._h, int(seq_id), _ptr(images), _ptr(mk) if mk is not None else None, C.byref(lab) if lab is not None else None, B, rows, cols,
                                                images.strides[1], images.strides[0], mk.strides[1] if mk is not None else 0, mk.strides[0] if mk is not None else 0,
                                                _ptr(kp), _ptr(desc), cap, _ptr(counts), _ptr(culled), C.byref(t)))
        self._keep[t.value] = (images, kp, desc, counts, culled, mk, keep)
        return t.value

    def result(self, ticket):
        _check(self._lib.orbx_pool_wait(self._h, C.c_longlong(ticket)))
        r = self._keep.pop(ticket)
        if r[4] is None and len(r) == 5:          # single frame
            n = int(r[3][0]); return r[1][:n].copy(), r[2][:n].copy()
        return r[1], r[2], r[3], r[4]

    def wait_all(self):
        _check(self._lib.orbx_pool_wait_all(self._h)); self._keep.clear()


from ._matcher import ORBmatcher, FrameView, Frame, Camera, ORBVocabulary  # noqa: E402,F401


SLIC_CENTER_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("L", "<i4"), ("A", "<i4"), ("B", "<i4"), ("D", "<i4"), ("label", "<i4")])


class cluster:
    """SLIC stage of ORB_SLAM2::cluster (/root/reference/include/cluster.h:33-83, src/cluster.cc:295-344) on the GPU.

    ``SLIC(lab, depth)`` takes the Lab image the reference obtains from cv::cvtColor(image, COLOR_BGR2Lab) (the conversion stays with
    the caller's OpenCV) and returns (labelMask as float64, centres) exactly as the reference's SLIC() leaves them; ``labels16=True``
    returns the labels as uint16 ids instead, the form ORBextractor.extract_masked_batch takes."""

    def __init__(self, device=0, length=5, m=10, rounds=5):
        self._lib = lib(); self._h = C.c_void_p()
        self.len, self.m, self.rounds = int(length), int(m), int(rounds)
        _check(self._lib.orbx_slic_create(int(device), C.byref(self._h)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.orbx_slic_destroy(self._h); self._h = None
        except Exception:
            pass

    def SLIC(self, lab, depth, labels16=False):
        lab = np.ascontiguousarray(lab, np.uint8); depth = np.ascontiguousarray(depth, np.uint16)
        rows, cols = depth.shape
        if lab.shape != (rows, cols, 3):
            raise OrbxError(E_INVALID, "lab must be rows x cols x 3")
        cap = (rows // self.len + 1) * (cols // self.len + 1)
        centers = np.zeros(cap, SLIC_CENTER_DTYPE); n = C.c_int(0)
        l64 = None if labels16 else np.zeros((rows, cols), np.float64)
        l16 = np.zeros((rows, cols), np.uint16) if labels16 else None
        _check(self._lib.orbx_slic_run(self._h, _ptr(lab), cols * 3, _ptr(depth), cols * 2, rows, cols, self.len, self.m, self.rounds,
                                       _ptr(l64) if l64 is not None else None, cols * 8, _ptr(l16) if l16 is not None else None, cols * 2,
                                       _ptr(centers), cap, C.byref(n)))
        return (l16 if labels16 else l64), centers[:n.value].copy()


def debug_canary_check():
    """(overwritten guard zones, buffers checked) -- only meaningful when the process was started with ORBX_CANARY=1 (test tap)."""
    n = C.c_int(0)
    bad = lib().orbx_debug_canary_check(C.byref(n))
    if bad < 0:
        _check(bad)
    return bad, n.value
