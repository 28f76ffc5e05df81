/* orbx_b200.h -- C ABI of the B200-native ORB feature front-end (liborbx_b200.so).
 *
 * Drop-in boundary for the Amos-SLAM / ORB-SLAM2 hot path: ORBextractor (pyramid, per-cell FAST-9,
 * quadtree distribution, IC orientation, 7x7 blur, rBRIEF, Amos dynamic-mask culling) and the
 * ORBmatcher / Frame::ComputeStereoMatches 256-bit Hamming matching.  Plain pointers and sizes only;
 * every entry point names the reference interface it replaces (paths relative to the reference root).
 * The C++ classes in amos-slam_b200/host/ (ORB_SLAM2::ORBextractor / ORBmatcher with the reference's
 * own signatures) marshal into these calls; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *   - every function returns an int status: ORBX_OK (0) or a negative ORBX_E_* code; no exceptions
 *     or aborts cross this boundary.  orbx_last_error() returns a thread-local message.
 *   - all output buffers are caller-allocated.  Host-pointer calls are synchronous (results are
 *     valid on return); *_device calls are stream-ordered on the handle's stream (orbx_stream()).
 *   - a handle is stateful and non-reentrant exactly like the reference's ORBextractor object
 *     (it keeps mvImagePyramid between detect / cull / describe); distinct handles are independent
 *     and may be driven from different threads (include/ORBextractor.h:93-168, src/Frame.cc:165-173).
 *   - there is NO CPU fallback: if no CUDA device is usable, create fails with ORBX_E_CUDA.
 */
#ifndef ORBX_B200_H
#define ORBX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBX_OK 0
#define ORBX_E_INVALID (-1)   /* bad argument (null pointer, non-positive size, unsupported geometry) */
#define ORBX_E_CUDA (-2)      /* CUDA runtime error; see orbx_last_error() */
#define ORBX_E_CAPACITY (-3)  /* caller buffer too small; *n_out holds the required count */
#define ORBX_E_STATE (-4)     /* call sequence error (e.g. describe before detect) */
#define ORBX_E_OVERFLOW (-5)  /* an internal worst-case bound was exceeded (reported, never silent) */

/* Binary-identical to cv::KeyPoint (28 bytes): the C++ layer memcpy's straight into std::vector<cv::KeyPoint>. */
typedef struct orbx_keypoint {
    float x, y;       /* pt */
    float size;       /* int(31 * scale[octave])            src/ORBextractor.cc:1175,1189 */
    float angle;      /* degrees, IC_Angle / fastAtan2      src/ORBextractor.cc:108-161  */
    float response;   /* FAST score S-1                      cv::FAST                     */
    int32_t octave;
    int32_t class_id; /* always -1 */
} orbx_keypoint;

typedef struct orbx_extractor orbx_extractor;

const char* orbx_last_error(void);
int orbx_version(void);

/* ------------------------------------------------------------------------------------------------
 * ORBextractor::ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
 *   include/ORBextractor.h:93, src/ORBextractor.cc:492-609.   `device` = CUDA ordinal.
 * ---------------------------------------------------------------------------------------------- */
int orbx_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                int device, orbx_extractor** out);
void orbx_destroy(orbx_extractor* h);

/* GetLevels / GetScaleFactor / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares   include/ORBextractor.h:117-165.  Arrays hold nlevels floats/ints. */
int orbx_get_levels(const orbx_extractor* h);
float orbx_get_scale_factor(const orbx_extractor* h);
int orbx_get_scale_factors(const orbx_extractor* h, float* out);
int orbx_get_inverse_scale_factors(const orbx_extractor* h, float* out);
int orbx_get_scale_sigma_squares(const orbx_extractor* h, float* out);
int orbx_get_inverse_scale_sigma_squares(const orbx_extractor* h, float* out);
int orbx_get_features_per_level(const orbx_extractor* h, int* out);   /* mnFeaturesPerLevel (protected member) */
/* upper bound of keypoints one frame can return: sum over levels of max(N_l + 2, 4*nIni_l) for the
 * geometry (rows, cols); use it to size kp/desc buffers. */
int orbx_max_keypoints(orbx_extractor* h, int rows, int cols);
/* the CUDA stream all work of this handle is ordered on (a cudaStream_t) */
void* orbx_stream(orbx_extractor* h);

/* ------------------------------------------------------------------------------------------------
 * void ORBextractor::operator()(InputArray image, InputArray mask, vector<KeyPoint>& keypoints,
 *                               OutputArray descriptors)          src/ORBextractor.cc:1544-1668
 * image: CV_8UC1, rows x cols, `step` bytes per row (host memory).  The mask is ignored by the
 * reference and therefore not part of this call.  Empty image (rows*cols == 0 or NULL) => 0 keypoints.
 * kp_out[cap], desc_out[cap*32] (row-major N x 32, CV_8U); *n_out = N.  Entries past N are left untouched.  If N > cap
 * nothing is written and ORBX_E_CAPACITY is returned (*n_out = N); orbx_max_keypoints() is always enough.
 * ---------------------------------------------------------------------------------------------- */
int orbx_extract(orbx_extractor* h, const uint8_t* image, int rows, int cols, size_t step,
                 orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out);

/* ------------------------------------------------------------------------------------------------
 * void ORBextractor::operator()(InputArray image, InputArray mask, vector<vector<KeyPoint>>& keypoints)
 *   src/ORBextractor.cc:1672-1686  (Amos stage 1: pyramid + keypoints with orientation, no descriptors).
 * Keypoints are returned level-major in LEVEL coordinates; level_counts[nlevels].
 * The pyramid stays resident in the handle for orbx_cull / orbx_describe / orbx_pyramid_level.
 * ---------------------------------------------------------------------------------------------- */
int orbx_detect(orbx_extractor* h, const uint8_t* image, int rows, int cols, size_t step,
                orbx_keypoint* kp_out, int* level_counts, int cap, int* n_out);

/* ------------------------------------------------------------------------------------------------
 * vector<KeyPoint> ORBextractor::MovingKeyPoints(imGray, imS, imLS, centers, rm_vector, DynaFlag, mvKeysT)
 *   src/ORBextractor.cc:1688-1745.  mask = imS (CV_8U rows x cols), label = imLS (CV_64F rows x cols,
 *   super-pixel ids, 1-based), centers_id[i] = centers[i].id, rm_vector[nrm].
 * kp_inout / level_counts: per-level keypoints (level coordinates) filtered in place, order preserved;
 * culled_out (may be NULL, capacity = number of input keypoints) receives the removed keypoints in
 * removal order; *n_culled = return value of the reference call's size().
 * ---------------------------------------------------------------------------------------------- */
int orbx_cull(orbx_extractor* h, const uint8_t* mask, size_t mask_step, const double* label, size_t label_step,
              int rows, int cols, const int* centers_id, int ncenters, const int* rm_vector, int nrm,
              orbx_keypoint* kp_inout, int* level_counts, orbx_keypoint* culled_out, int* n_culled);

/* ------------------------------------------------------------------------------------------------
 * void ORBextractor::ProcessDesp(image, mask, allKeypoints, mKeypoints, descriptors)
 *   src/ORBextractor.cc:1747-1820  (Amos stage 2: blur + rBRIEF on the surviving keypoints, then
 *   pt *= scale).  Uses the pyramid kept by the last orbx_detect on this handle.
 * ---------------------------------------------------------------------------------------------- */
int orbx_describe(orbx_extractor* h, const orbx_keypoint* kp_in, const int* level_counts,
                  orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out);

/* std::vector<cv::Mat> mvImagePyramid   include/ORBextractor.h:168.
 * Copies level `level` of the resident pyramid to host memory.  border = 0 copies the ROI
 * (rows_l x cols_l); border = 19 reproduces the reference's padded parent buffer
 * ((rows_l+38) x (cols_l+38), BORDER_REFLECT_101, src/ORBextractor.cc:1859-1882).
 * dst may be NULL to query the size only. */
int orbx_pyramid_level(orbx_extractor* h, int level, int border, uint8_t* dst, size_t dst_step,
                       int* rows_out, int* cols_out);

/* ------------------------------------------------------------------------------------------------
 * Batched extraction: B independent frames of identical geometry per call (frames / sequences are
 * independent units).  Semantically B calls of operator()(image, mask, kps, desc).
 *   images      : B frames, frame b at images + b*frame_stride, rows x cols, `step` bytes per row
 *   kp_out      : B * cap keypoints   (frame b at kp_out + b*cap)
 *   desc_out    : B * cap * 32 bytes
 *   counts_out  : B ints
 * orbx_extract_batch       : host pointers, synchronous (H2D + kernels + D2H inside the call)
 * orbx_extract_batch_device: device pointers, asynchronous on orbx_stream(h); inputs must stay valid
 *                            until the stream reaches the end of the call's work.
 * Capacity: counts_out[b] is always the TRUE number of keypoints of frame b.  If it exceeds cap only the first cap of
 * them (level-major order) were written: the host-pointer calls then return ORBX_E_CAPACITY, callers of the *_device
 * forms compare counts_out with cap themselves.  cap >= orbx_max_keypoints() never truncates.
 * Overflow: the host-pointer calls return ORBX_E_OVERFLOW (and clear the flag) when an internal worst-case bound of the
 * quadtree stage was exceeded; the *_device forms cannot (they do not synchronise): poll orbx_check_overflow().
 * ---------------------------------------------------------------------------------------------- */
int orbx_extract_batch(orbx_extractor* h, const uint8_t* images, int B, int rows, int cols, size_t step,
                       size_t frame_stride, orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* counts_out);
int orbx_extract_batch_device(orbx_extractor* h, const uint8_t* d_images, int B, int rows, int cols, size_t step,
                              size_t frame_stride, orbx_keypoint* d_kp_out, uint8_t* d_desc_out, int cap,
                              int* d_counts_out);
/* ------------------------------------------------------------------------------------------------
 * Batched Amos path (BASELINE config 5: multi-sequence extraction with dynamic-mask keypoint culling).  Per frame b:
 *   operator()(image_b, mask, vector<vector<KeyPoint>>&)                         src/ORBextractor.cc:1672-1686
 *   MovingKeyPoints(.., imS = mask_b, ..) with no flagged super-pixel (rm_vector all 0)       :1688-1745
 *   ProcessDesp(..)                                                                            :1747-1820
 * i.e. the keypoints of operator()(image, mask, kps, desc) minus those whose position lies where
 * erode(dilate(mask_b, ellipse31), ellipse31) != 0, with descriptors of the survivors only.
 * masks: CV_8UC1, rows x cols, frame b at masks + b*mask_frame_stride, `mask_step` bytes per row.
 * culled_out (may be NULL): number of keypoints removed per frame.
 * ---------------------------------------------------------------------------------------------- */
int orbx_extract_masked_batch(orbx_extractor* h, const uint8_t* images, const uint8_t* masks, int B, int rows, int cols,
                              size_t step, size_t frame_stride, size_t mask_step, size_t mask_frame_stride,
                              orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* counts_out, int* culled_out);
int orbx_extract_masked_batch_device(orbx_extractor* h, const uint8_t* d_images, const uint8_t* d_masks, int B, int rows, int cols,
                                     size_t step, size_t frame_stride, size_t mask_step, size_t mask_frame_stride,
                                     orbx_keypoint* d_kp_out, uint8_t* d_desc_out, int cap, int* d_counts_out, int* d_culled_out);
/* The same with the super-pixel term of MovingKeyPoints (src/ORBextractor.cc:1722-1736): a keypoint is also removed when
 *   rm_vector[centers[imLS(p) - 1].id] == 1,  p = (int)(pt * scale).
 * imLS travels as 16-bit super-pixel ids (the reference keeps them in a CV_64F image; ids are 1-based integers), frame b at
 * labels + b * label_frame_stride, label_step ELEMENTS per row; flagged[b * n_labels + id - 1] = (rm_vector[centers[id - 1].id] == 1),
 * which the caller folds from its two small tables (centers, rm_vector) per frame.  Ids outside [1, n_labels] (undefined behaviour
 * in the reference) count as not flagged.  labels == NULL behaves like the calls above.  Host pointers for the host call,
 * device pointers (the struct itself on the host) for the *_device call. */
typedef struct orbx_labels {
    const uint16_t* labels; size_t label_step, label_frame_stride;   /* in elements */
    const uint8_t* flagged; int n_labels;
} orbx_labels;
int orbx_extract_masked_batch_labels(orbx_extractor* h, const uint8_t* images, const uint8_t* masks, const orbx_labels* labels, int B, int rows, int cols,
                                     size_t step, size_t frame_stride, size_t mask_step, size_t mask_frame_stride,
                                     orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* counts_out, int* culled_out);
int orbx_extract_masked_batch_labels_device(orbx_extractor* h, const uint8_t* d_images, const uint8_t* d_masks, const orbx_labels* d_labels, int B, int rows, int cols,
                                            size_t step, size_t frame_stride, size_t mask_step, size_t mask_frame_stride,
                                            orbx_keypoint* d_kp_out, uint8_t* d_desc_out, int cap, int* d_counts_out, int* d_culled_out);
/* number of kernel launches issued by this handle so far (bench.py's gpu_launches) */
long long orbx_launch_count(const orbx_extractor* h);
/* Per-stage device timing (bench.py): while enabled every batched extract records CUDA events at the stage
 * boundaries on the handle's stream; collect synchronises, writes the summed elapsed ms of the 6 stages
 * (resize, fast_cells, octree_sort, octree_tree, gauss7, orient_describe) and the number of profiled calls. */
int orbx_profile_enable(orbx_extractor* h, int on);
int orbx_profile_collect(orbx_extractor* h, double* stage_ms6, int* ncalls);
/* internal overflow flags raised since the last report (0 = none, else ORBX_OVF bits of the quadtree stage); synchronises the
 * handle's stream and clears the flags, so one bad frame is reported once and later calls on the handle are unaffected */
int orbx_check_overflow(orbx_extractor* h);

/* ------------------------------------------------------------------------------------------------
 * Multi-sequence / multi-GPU driver (BASELINE config 5): camera streams ("sequences") are pinned to workers by
 *     gpu = seq_id mod n_gpus,   stream = (seq_id div n_gpus) mod streams_per_gpu
 * Every worker is one host thread that owns one extractor handle (one CUDA stream, one resident pyramid) on its GPU and runs its
 * jobs in submission order -- the reference's "one ORBextractor object per camera, each in its own thread" (src/Frame.cc:165-173)
 * scaled out over the GPUs of one box.  No collective, no shared device state.  A single process drives all GPUs.
 *   submit*       : asynchronous; returns a ticket.  Input and output buffers (host memory, pinned for full copy / compute overlap) must
 *                   stay valid until the ticket has been waited for.  submit = operator()(image, mask, kps, desc) of one frame;
 *                   submit_batch = orbx_extract_batch (masks == NULL) or orbx_extract_masked_batch_labels (labels may be NULL).
 *   wait / wait_all: block until the job(s) finished; return the job's status (first failing status for wait_all).
 * ---------------------------------------------------------------------------------------------- */
typedef struct orbx_pool orbx_pool;
int  orbx_pool_shard_of(int seq_id, int n_gpus, int streams_per_gpu, int* gpu, int* stream);     /* the mapping above, no device needed */
int  orbx_pool_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int n_gpus, const int* devices /* NULL: 0 .. n_gpus-1 */,
                      int streams_per_gpu, orbx_pool** out);
void orbx_pool_destroy(orbx_pool* p);                                                            /* finishes the queued jobs first */
int  orbx_pool_gpus(const orbx_pool* p);
int  orbx_pool_streams_per_gpu(const orbx_pool* p);
int  orbx_pool_device_of(const orbx_pool* p, int seq_id);                                        /* CUDA ordinal that serves this sequence */
long long orbx_pool_frames_done(const orbx_pool* p, int gpu, int stream);                        /* frames finished by one worker so far */
int  orbx_pool_submit(orbx_pool* p, int seq_id, const uint8_t* image, int rows, int cols, size_t step,
                      orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out, long long* ticket);
int  orbx_pool_submit_batch(orbx_pool* p, int seq_id, const uint8_t* images, const uint8_t* masks, const orbx_labels* labels, int B, int rows, int cols,
                            size_t step, size_t frame_stride, size_t mask_step, size_t mask_frame_stride,
                            orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* counts_out, int* culled_out, long long* ticket);
int  orbx_pool_wait(orbx_pool* p, long long ticket);
int  orbx_pool_wait_all(orbx_pool* p);

/* ------------------------------------------------------------------------------------------------
 * SLIC super-pixels of `cluster` (src/cluster.cc:88-344): the label map MovingKeyPoints reads
 * (src/ORBextractor.cc:1722-1736) and the centres the reference's k-means groups afterwards.
 *   lab       rows x cols x 3, 8-bit: cv::cvtColor(image, COLOR_BGR2Lab) -- the conversion stays with the caller's OpenCV
 *             (src/cluster.cc:305), everything after it (Sobel gradient, initilizeCenters, fituneCenter, `rounds` x clustering +
 *             updateCenter) runs on the device and returns what the reference's SLIC() returns, bit for bit;
 *   depth     rows x cols, 16-bit (imD); only feeds the centres' D field;
 *   len, m    super-pixel size and compactness (the reference passes 5 and 10, :12-14), rounds = 5 (:330);
 *   labels_out   rows x cols doubles = the reference's CV_64F labelMask (label = centre number from 1, 0 = never covered), may be NULL;
 *   labels16_out the same as 16-bit ids -- the form orbx_extract_masked_batch_labels takes -- may be NULL (E_INVALID if > 65535 centres);
 *   centers_out  n records {x, y, L, A, B, D, label} (struct center, include/cluster.h:22-31, without the k-means id), may be NULL.
 * Steps are in BYTES.  *n_out = number of centres (also on E_CAPACITY).
 * ---------------------------------------------------------------------------------------------- */
typedef struct orbx_slic orbx_slic;
typedef struct orbx_slic_center { int x, y, L, A, B, D, label; } orbx_slic_center;
int  orbx_slic_create(int device, orbx_slic** out);
void orbx_slic_destroy(orbx_slic* h);
int  orbx_slic_run(orbx_slic* h, const uint8_t* lab, size_t lab_step, const uint16_t* depth, size_t depth_step, int rows, int cols, int len, int m, int rounds,
                   double* labels_out, size_t labels_step, uint16_t* labels16_out, size_t labels16_step, orbx_slic_center* centers_out, int cap, int* n_out);

/* ================================================================================================
 * ORBmatcher  (include/ORBmatcher.h:57-215, src/ORBmatcher.cc)
 * ================================================================================================ */
#define ORBX_TH_HIGH 100       /* ORBmatcher::TH_HIGH      src/ORBmatcher.cc:49 */
#define ORBX_TH_LOW 50         /* ORBmatcher::TH_LOW       src/ORBmatcher.cc:50 */
#define ORBX_HISTO_LENGTH 30   /* ORBmatcher::HISTO_LENGTH src/ORBmatcher.cc:51 */
#define ORBX_FRAME_GRID_COLS 64   /* include/Frame.h:61 */
#define ORBX_FRAME_GRID_ROWS 48   /* include/Frame.h:56 */

typedef struct orbx_matcher orbx_matcher;

/* ORBmatcher::ORBmatcher(float nnratio = 0.6, bool checkOri = true)   src/ORBmatcher.cc:54 */
int orbx_matcher_create(float nnratio, int check_orientation, int device, orbx_matcher** out);
void orbx_matcher_destroy(orbx_matcher* m);
void* orbx_matcher_stream(orbx_matcher* m);
long long orbx_matcher_launch_count(const orbx_matcher* m);
/* Per-stage device timing of the batched matcher calls (bench.py): see orbx_profile_enable.  Stages of
 * orbx_search_for_initialization[_frames]_batch: grid build, window count, scan, window fill, resolve (nstages = 5);
 * of orbx_compute_stereo_matches_batch[_device]: stereo match, median cut (nstages = 2). */
int orbx_matcher_profile_enable(orbx_matcher* m, int on);
int orbx_matcher_profile_collect(orbx_matcher* m, int nstages, double* stage_ms, int* ncalls);

/* static int ORBmatcher::DescriptorDistance(const Mat& a, const Mat& b)   src/ORBmatcher.cc:1913-1933
 * n pairs: out[i] = popcount(a[i] xor b[i]) over 256 bits (device kernel; host pointers). */
int orbx_descriptor_distance(orbx_matcher* m, const uint8_t* a, const uint8_t* b, int n, int* out);

/* The part of a Frame the matchers read (src/Frame.cc:431-461 AssignFeaturesToGrid, :894-1003
 * GetFeaturesInArea, :1007-1030 PosInGrid; include/Frame.h).  All pointers are host memory. */
typedef struct orbx_frame_view {
    int n;                          /* Frame::N */
    const orbx_keypoint* keys_un;   /* mvKeysUn[n]  (undistorted keypoints; pt, octave, angle are read) */
    const uint8_t* descriptors;     /* mDescriptors, n x 32 */
    const float* u_right;           /* mvuRight[n] or NULL (treated as all -1) */
    float min_x, min_y, max_x, max_y;                 /* mnMinX, mnMinY, mnMaxX, mnMaxY */
    float grid_element_width_inv, grid_element_height_inv;   /* mfGridElementWidthInv / HeightInv */
    int nlevels; const float* scale_factors;          /* mvScaleFactors[nlevels] */
} orbx_frame_view;

/* int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<Point2f>& vbPrevMatched,
 *                                         vector<int>& vnMatches12, int windowSize = 10)
 *   src/ORBmatcher.cc:515-643.  prev_matched: F1.n (x,y) float pairs, updated in place (:638-640);
 *   matches12: F1.n ints (index into F2 or -1); *nmatches = return value. */
int orbx_search_for_initialization(orbx_matcher* m, const orbx_frame_view* F1, const orbx_frame_view* F2,
                                   float* prev_matched_xy, int* matches12, int window_size, int* nmatches);

/* The same for n_pairs independent frame pairs in one call (the frame pair is the shard unit of BASELINE config 2):
 * semantically n_pairs calls of ORBmatcher::SearchForInitialization(F1[p], F2[p], prev_matched_xy[p], matches12[p], window_size) on
 * fresh matcher objects with this handle's (nnratio, checkOri) -- one upload, five kernel launches, one download for the whole batch
 * (grid build, windowed Hamming search and ordered resolve run one CTA / one warp per pair, query).  F1 / F2: arrays of n_pairs views;
 * prev_matched_xy[p]: F1[p].n (x, y) pairs, updated in place; matches12[p]: F1[p].n ints; nmatches[p] = return value of pair p. */
int orbx_search_for_initialization_batch(orbx_matcher* m, int n_pairs, const orbx_frame_view* F1, const orbx_frame_view* F2,
                                         float* const* prev_matched_xy, int* const* matches12, int window_size, int* nmatches);

/* int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, float th, bool bMono)
 *   src/ORBmatcher.cc:1569-1728.  The pose algebra (Rcw, tcw, tlc) and the projection of LastFrame's
 *   map points stay in the caller (they need MapPoint objects); the call receives, per LastFrame
 *   feature i with a valid non-outlier map point that projects inside the image (:1597-1623):
 *     proj_uv[2i], proj_uv[2i+1] = (u, v);  proj_invz[i];  last_octave[i] = LastFrame.mvKeys[i].octave;
 *     last_angle[i] = LastFrame.mvKeysUn[i].angle;  mp_desc + 32 i = pMP->GetDescriptor();
 *     valid[i] != 0;  mp_observed[i] != 0 iff pMP->Observations() > 0.
 *   cur_occupied[j] != 0 marks CurrentFrame features that already hold a map point with Observations() > 0
 *   (:1658-1660).  A feature claimed earlier in the loop blocks later map points iff its claimer is observed;
 *   otherwise it can be re-assigned, exactly as the reference's running mvpMapPoints state behaves.
 *   forward / backward = bForward / bBackward (:1591-1592); mbf = CurrentFrame.mbf.
 *   Output: cur_match[j] = index i of the LastFrame feature whose map point is assigned to current
 *   feature j after the rotation-histogram filter (:1706-1725); -1 = mvpMapPoints[j] was never written by
 *   the call; -2 = it was assigned and then reset to NULL by the histogram filter (:1719).
 *   *nmatches = return value. */
int orbx_search_by_projection_frame(orbx_matcher* m, const orbx_frame_view* cur, int n_last,
                                    const float* proj_uv, const float* proj_invz, const int* last_octave,
                                    const float* last_angle, const uint8_t* mp_desc, const uint8_t* valid,
                                    const uint8_t* mp_observed, const uint8_t* cur_occupied, float th, int forward, int backward, float mbf,
                                    int* cur_match, int* nmatches);

struct orbx_frame;   /* device-resident Frame, declared below */
/* The same call with the projection itself on the device (src/ORBmatcher.cc:1597-1623): instead of (u, v, 1/z) the caller passes, per
 * LastFrame feature i, has_point[i] != 0 iff mvpMapPoints[i] is set and not an outlier, and world_xyz + 3 i = pMP->GetWorldPos(); plus the
 * pose of CurrentFrame (Rcw row-major 3 x 3, tcw) and its intrinsics.  x3Dc = Rcw * x3Dw + tcw is evaluated exactly as cv::gemm
 * evaluates it for these sizes (float arithmetic, ((r0 x + r1 y) + r2 z) + t; pinned against cv2.gemm), invzc = 1.0 / z in double
 * rounded to float, u = fx xc invzc + cx in float; points with invzc < 0 or outside [mnMinX, mnMaxX] x [mnMinY, mnMaxY] drop out.
 * Exactly one of cur (host view) and cur_dev (device frame) is non-NULL.  proj_uv_out (2 n), proj_invz_out (n), valid_out (n) are
 * optional taps of what the device computed (NULL in production). */
int orbx_search_by_projection_frame_pose(orbx_matcher* m, const orbx_frame_view* cur, const struct orbx_frame* cur_dev, int n_last,
                                         const float* world_xyz, const uint8_t* has_point, const float* Rcw, const float* tcw,
                                         float fx, float fy, float cx, float cy, const int* last_octave, const float* last_angle,
                                         const uint8_t* mp_desc, const uint8_t* mp_observed, const uint8_t* cur_occupied, float th,
                                         int forward, int backward, float mbf, int* cur_match, int* nmatches,
                                         float* proj_uv_out, float* proj_invz_out, uint8_t* valid_out);

/* int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, float th)
 *   src/ORBmatcher.cc:70-175.  Per map point p (already filtered by mbTrackInView && !isBad(), :82-86):
 *     track_uv (mTrackProjX, mTrackProjY), track_ur (mTrackProjXR), track_level (mnTrackScaleLevel),
 *     track_view_cos (mTrackViewCos), mp_desc (GetDescriptor()), mp_observed (Observations() > 0).
 *   f_occupied[j] != 0 marks F features that already hold a map point with Observations() > 0 (:124-126).
 *   Output: f_match[j] = index p of the map point assigned to feature j (or -1); *nmatches. */
int orbx_search_by_projection_points(orbx_matcher* m, const orbx_frame_view* F, int n_points,
                                     const float* track_uv, const float* track_ur, const int* track_level,
                                     const float* track_view_cos, const uint8_t* mp_desc, const uint8_t* mp_observed,
                                     const uint8_t* f_occupied, float th, int* f_match, int* nmatches);

/* void Frame::ComputeStereoMatches()   src/Frame.cc:1179-1573.
 *   left / right: the two extractor handles holding the pyramids of the current stereo pair
 *   (mpORBextractorLeft / Right ->mvImagePyramid);  keys / descriptors as returned by orbx_extract
 *   (mvKeys, mvKeysRight, mDescriptors, mDescriptorsRight).  mb and mbf are passed as the reference
 *   has them AT THE TIME OF THE CALL (mb == 0 inside the stereo constructor: src/Frame.cc:131 vs :237).
 *   Output: u_right[nl] (mvuRight), depth[nl] (mvDepth). */
int orbx_compute_stereo_matches(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right,
                                const orbx_keypoint* keys_left, const uint8_t* desc_left, int nl,
                                const orbx_keypoint* keys_right, const uint8_t* desc_right, int nr,
                                float mb, float mbf, float* u_right, float* depth);

/* The same for the B stereo pairs that went through the LAST BATCHED extract call of the two handles (orbx_extract_batch[_device] on
 * `left` with the B left images, on `right` with the B right images, same cap): both pyramids, keypoints, descriptors and counts are
 * still on the device, so nothing is uploaded.  Outputs are [B][cap] floats (pair b at + b * cap; entries past the pair's left keypoint
 * count are -1).  The host form is synchronous; the device form is asynchronous on orbx_matcher_stream(m), which first waits for the
 * work queued on both extractors' streams; in turn, work queued on either extractor AFTER this call waits for the stereo kernels (they read
 * the extractors' pyramids and results).  The stereo pair is the shard unit of BASELINE config 4. */
int orbx_compute_stereo_matches_batch(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int B, int cap, float mb, float mbf,
                                      float* u_right, float* depth);
int orbx_compute_stereo_matches_batch_device(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int B, int cap, float mb, float mbf,
                                             float* d_u_right, float* d_depth);

/* Brute-force all-pairs Hamming with best / second-best (the inner kernel of all matchers, exposed for
 * throughput measurement):  for every query q: best_idx, best_dist, second_dist over all n_train rows. */
int orbx_match_bruteforce_device(orbx_matcher* m, const uint8_t* d_query, int n_query, const uint8_t* d_train,
                                 int n_train, int* d_best_idx, int* d_best_dist, int* d_second_dist);
/* The same for n_pairs independent frame pairs in one launch: pair p matches d_query[p][n_query][32] against
 * d_train[p][n_train][32]; outputs are [n_pairs][n_query].  Asynchronous on orbx_matcher_stream(m). */
int orbx_match_bruteforce_batch_device(orbx_matcher* m, int n_pairs, const uint8_t* d_query, int n_query, const uint8_t* d_train,
                                       int n_train, int* d_best_idx, int* d_best_dist, int* d_second_dist);

/* ------------------------------------------------------------------------------------------------
 * Device-resident Frame: the per-frame steps the reference runs between the
 * extractor and the matchers, kept on the GPU so that keypoints / descriptors are not uploaded again
 * by every matcher call and the 64x48 grid is built once per frame:
 *   Frame::UndistortKeyPoints()          src/Frame.cc:1052-1117   (cv::undistortPoints, 5 iterations, double)
 *   Frame::ComputeImageBounds()          src/Frame.cc:1120-1176   (cached per camera and image size)
 *   Frame::ComputeStereoFromRGBD()       src/Frame.cc:1576-1614
 *   Frame::AssignFeaturesToGrid()        src/Frame.cc:431-461, PosInGrid :1007-1030
 *   Frame::GetFeaturesInArea()           src/Frame.cc:894-1003
 * in the order Frame::CalDyna runs them (src/Frame.cc:636-645) and the stereo / mono constructors do.
 * ------------------------------------------------------------------------------------------------ */
typedef struct orbx_frame orbx_frame;

/* mK (fx, fy, cx, cy), mDistCoef (k1, k2, p1, p2, k3; k1 == 0 means "already rectified", :1058) and mbf,
 * as float like the reference's CV_32F matrices (src/Tracking.cc camera block). */
typedef struct orbx_camera {
    float fx, fy, cx, cy;
    float k1, k2, p1, p2, k3;
    float bf;
} orbx_camera;

int  orbx_frame_create(int device, orbx_frame** out);
void orbx_frame_destroy(orbx_frame* f);

/* Builds the frame from the result the extractor handle still holds on the device (its last orbx_extract or
 * orbx_describe call: mvKeys / mDescriptors), without a host round trip:
 *   N = mvKeys.size(); UndistortKeyPoints(); ComputeStereoFromRGBD(imDepth); AssignFeaturesToGrid().
 * depth: host pointer, CV_32F.  depth_stride_bytes > 0: the whole image (img_rows x img_cols, row stride in
 * bytes), read as imDepth.at<float>(v, u) with the keypoint's float coordinates truncated (:1595);
 * depth_stride_bytes == 0: depth[i] is that value already gathered by the caller for keypoint i;
 * depth == NULL: monocular / stereo frame, mvuRight = mvDepth = -1 until orbx_frame_set_stereo().
 * Image bounds follow ComputeImageBounds for (cam, img_rows, img_cols) and are cached in the handle. */
int orbx_frame_assign(orbx_frame* f, orbx_extractor* h, const orbx_camera* cam, int img_rows, int img_cols,
                      const float* depth, size_t depth_stride_bytes);
/* The same from host arrays (frames whose keypoints did not come from an extractor handle of this process). */
int orbx_frame_assign_host(orbx_frame* f, const orbx_keypoint* keys, const uint8_t* descriptors, int n,
                           int nlevels, const float* scale_factors, const orbx_camera* cam, int img_rows, int img_cols,
                           const float* depth, size_t depth_stride_bytes);
/* The same steps one by one, in whatever order the reference's constructors run them (stereo: :187-240, mono: :378-428,
 * Amos RGB-D: :636-645), for a drop-in that keeps Frame's call sites untouched.  Each call returns after its result is
 * complete; optional *_out pointers receive the host copy the reference's members need.
 *   take                      N = mvKeys.size(): mvKeys / mDescriptors from the extractor's device result or from host arrays
 *   undistort_keypoints       Frame::UndistortKeyPoints()                 -> keys_un_out[N]  (mvKeysUn)
 *   compute_stereo_from_rgbd  Frame::ComputeStereoFromRGBD(imDepth)       -> u_right_out[N], depth_out[N]; depth as in orbx_frame_assign
 *   assign_features_to_grid   Frame::AssignFeaturesToGrid() with bounds6 = Frame::mnMinX, mnMaxX, mnMinY, mnMaxY,
 *                             mfGridElementWidthInv, mfGridElementHeightInv -> cell_start_out / entries_out as orbx_frame_grid;
 *                             only after this call do the matchers accept the frame. */
int orbx_frame_take(orbx_frame* f, orbx_extractor* h);
int orbx_frame_take_host(orbx_frame* f, const orbx_keypoint* keys, const uint8_t* descriptors, int n, int nlevels, const float* scale_factors);
int orbx_frame_undistort_keypoints(orbx_frame* f, const orbx_camera* cam, orbx_keypoint* keys_un_out);
int orbx_frame_compute_stereo_from_rgbd(orbx_frame* f, float bf, const float* depth, size_t depth_stride_bytes, int img_rows, int img_cols,
                                        float* u_right_out, float* depth_out);
int orbx_frame_assign_features_to_grid(orbx_frame* f, const float* bounds6, int* cell_start_out, int* entries_out);
/* mvuRight / mvDepth computed elsewhere (orbx_compute_stereo_matches): host arrays of N floats. */
int orbx_frame_set_stereo(orbx_frame* f, const float* u_right, const float* depth);

int orbx_frame_size(const orbx_frame* f);           /* Frame::N once the grid is built (the matchers accept the frame), else -1 */
int orbx_frame_taken(const orbx_frame* f);          /* keypoints held after orbx_frame_take*, whatever the later steps; -1 = none */
/* Host copies; any pointer may be NULL.  keys_un[N] (mvKeysUn), u_right[N], depth[N],
 * bounds[6] = mnMinX, mnMaxX, mnMinY, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv. */
int orbx_frame_read(orbx_frame* f, orbx_keypoint* keys_un, float* u_right, float* depth, float* bounds);
/* mGrid: cell (x, y) holds entries[cell_start[x * 48 + y] .. cell_start[x * 48 + y + 1]) in push_back order;
 * cell_start has 64 * 48 + 1 ints, entries N ints (keypoints outside the grid are in no cell, :1020-1024). */
int orbx_frame_grid(orbx_frame* f, int* cell_start, int* entries);
/* vector<size_t> Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel) for nq windows at once (src/Frame.cc:894-1003):
 * window q returns counts[q] indices, in the reference's order, at indices[offsets[q]..); offsets has nq + 1 ints.
 * *total = offsets[nq]; if it exceeds cap nothing is written to indices and ORBX_E_CAPACITY is returned. */
int orbx_frame_features_in_area(orbx_matcher* m, const orbx_frame* f, int nq, const float* xy, const float* r,
                                const int* min_level, const int* max_level, int* offsets, int* indices, int cap, int* total);

/* The three windowed matchers on device-resident frames: same semantics as the orbx_frame_view calls above,
 * minus the per-call upload of keypoints / descriptors / mvuRight and the per-call grid build. */
int orbx_search_for_initialization_frames(orbx_matcher* m, const orbx_frame* F1, const orbx_frame* F2,
                                          float* prev_matched_xy, int* matches12, int window_size, int* nmatches);
int orbx_search_for_initialization_frames_batch(orbx_matcher* m, int n_pairs, const orbx_frame* const* F1, const orbx_frame* const* F2,
                                                float* const* prev_matched_xy, int* const* matches12, int window_size, int* nmatches);
int orbx_search_by_projection_frame_dev(orbx_matcher* m, const orbx_frame* cur, int n_last,
                                        const float* proj_uv, const float* proj_invz, const int* last_octave,
                                        const float* last_angle, const uint8_t* mp_desc, const uint8_t* valid,
                                        const uint8_t* mp_observed, const uint8_t* cur_occupied, float th, int forward, int backward, float mbf,
                                        int* cur_match, int* nmatches);
int orbx_search_by_projection_points_dev(orbx_matcher* m, const orbx_frame* F, int n_points,
                                         const float* track_uv, const float* track_ur, const int* track_level,
                                         const float* track_view_cos, const uint8_t* mp_desc, const uint8_t* mp_observed,
                                         const uint8_t* f_occupied, float th, int* f_match, int* nmatches);

/* int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, float th, int ORBdist)
 *   src/ORBmatcher.cc:1731-1863 (relocalisation, src/Tracking.cc:2663; first function of SURVEY.md 8f rank 3).  As in the Frame x Frame
 *   form the pose algebra stays in the caller; per KeyFrame map point i that is neither NULL, bad nor in sAlreadyFound, projects
 *   inside the image bounds and lies within its scale-invariance distances (:1753-1785):  proj_uv (u, v); predicted_level[i] =
 *   pMP->PredictScale(dist3D, &CurrentFrame); kf_angle[i] = pKF->mvKeysUn[i].angle; mp_desc = GetDescriptor(); valid[i] != 0.
 *   cur_occupied[j] != 0 iff CurrentFrame.mvpMapPoints[j] is not NULL (:1808; any map point blocks, so does every claim made earlier in
 *   the call).  A match needs bestDist <= orb_dist (:1820).  cur_match as for orbx_search_by_projection_frame (-2 = reset by :1853). */
int orbx_search_by_projection_keyframe(orbx_matcher* m, const orbx_frame_view* cur, int n_kf, const float* proj_uv,
                                       const int* predicted_level, const float* kf_angle, const uint8_t* mp_desc, const uint8_t* valid,
                                       const uint8_t* cur_occupied, float th, int orb_dist, int* cur_match, int* nmatches);
int orbx_search_by_projection_keyframe_dev(orbx_matcher* m, const orbx_frame* cur, int n_kf, const float* proj_uv,
                                           const int* predicted_level, const float* kf_angle, const uint8_t* mp_desc, const uint8_t* valid,
                                           const uint8_t* cur_occupied, float th, int orb_dist, int* cur_match, int* nmatches);

/* int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, vector<MapPoint*> &vpMatched, int th)
 *   src/ORBmatcher.cc:388-512 (loop closing, src/LoopClosing.cc).  kf = the KeyFrame's mvKeysUn / mDescriptors / grid (KeyFrame copies them
 *   from its Frame; KeyFrame::GetFeaturesInArea, src/KeyFrame.cc:752-797, walks the grid as Frame's does).  Per map point p that is not
 *   bad, not already in vpMatched, in front of the camera, inside the image, within its distance range and seen from less than 60 deg
 *   (:411-446): proj_uv (u, v); predicted_level[p] = PredictScale(dist, pKF); mp_desc; valid[p] != 0.  kf_matched[j] != 0 iff
 *   vpMatched[j] is not NULL.  Candidates must lie on level predicted - 1 or predicted (:462-463); a match needs bestDist <= TH_LOW.
 *   Output: kf_match[j] = index p newly written to vpMatched[j] (or -1); *nmatches = return value. */
int orbx_search_by_projection_keyframe_points(orbx_matcher* m, const orbx_frame_view* kf, int n_points, const float* proj_uv,
                                              const int* predicted_level, const uint8_t* mp_desc, const uint8_t* valid,
                                              const uint8_t* kf_matched, float th, int* kf_match, int* nmatches);
int orbx_search_by_projection_keyframe_points_dev(orbx_matcher* m, const orbx_frame* kf, int n_points, const float* proj_uv,
                                                  const int* predicted_level, const uint8_t* mp_desc, const uint8_t* valid,
                                                  const uint8_t* kf_matched, float th, int* kf_match, int* nmatches);

/* int ORBmatcher::SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12, const float &s12, const cv::Mat &R12,
 *                              const cv::Mat &t12, const float th)        src/ORBmatcher.cc:1290-1555 (loop closing)
 *   kf1 / kf2: the two KeyFrames as frame views.  Side 1, per feature i of pKF1 whose map point exists, is not bad, is not already
 *   matched (vbAlreadyMatched1), lands in front of camera 2, inside its image and within its distance range (:1369-1398):
 *   proj_uv1 = its projection into pKF2, predicted_level1 = PredictScale(dist3D, pKF2), mp_desc1, valid1 != 0.  Side 2 likewise into pKF1.
 *   Both passes take the best feature on level predicted - 1 or predicted with distance <= TH_HIGH, independently of each other;
 *   match12[i] = index in pKF2 for the pairs that agree both ways (-1 otherwise); *nfound = return value. */
int orbx_search_by_sim3(orbx_matcher* m, const orbx_frame_view* kf1, const orbx_frame_view* kf2,
                        const float* proj_uv1, const int* predicted_level1, const uint8_t* mp_desc1, const uint8_t* valid1,
                        const float* proj_uv2, const int* predicted_level2, const uint8_t* mp_desc2, const uint8_t* valid2,
                        float th, int* match12, int* nfound);

/* The search inside both ORBmatcher::Fuse forms:
 *   int Fuse(KeyFrame *pKF, const vector<MapPoint*> &vpMapPoints, const float th)                                   src/ORBmatcher.cc:1020-1175
 *   int Fuse(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, float th, vector<MapPoint*> &vpReplacePoint)   :1179-1310
 * Per map point that passes the caller-side tests (not NULL / bad / already in the KeyFrame, in front of the camera, inside the image, within its
 * distance range, seen from less than 60 deg): proj_uv (u, v), predicted_level = PredictScale(dist3D, pKF), mp_desc, valid != 0.
 * Pose form: proj_ur[p] = u - bf * invz and inv_level_sigma2 = pKF->mvInvLevelSigma2 (kf->u_right = pKF->mvuRight): a candidate must pass the
 * chi-square gate e2 * invSigma2 <= 7.8 (stereo feature) / 5.99 (monocular) (:1097-1137).  Sim3 form: proj_ur = NULL, no gate.
 * best_idx[p] = the KeyFrame feature with the smallest distance on level predicted - 1 or predicted if that distance is <= TH_LOW, else -1.
 * The map surgery that follows a hit (Replace / AddObservation / AddMapPoint / vpReplacePoint, :1146-1170, :1289-1303) needs the MapPoint
 * objects and stays in the caller; it does not feed back into the search of later points. */
int orbx_fuse_search(orbx_matcher* m, const orbx_frame_view* kf, int n_points, const float* proj_uv, const float* proj_ur,
                     const int* predicted_level, const uint8_t* mp_desc, const uint8_t* valid, const float* inv_level_sigma2,
                     float th, int* best_idx);

/* ------------------------------------------------------------------------------------------------
 * Bag of words: DBoW2's vocabulary tree and the two BoW-guided matchers.
 *   ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>   include/ORBVocabulary.h:40-41
 *   transform(features, BowVector&, FeatureVector&, levelsup)   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1124-1197
 *   transform(feature, word_id, weight, nid, levelsup)          Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1217-1259
 *   Frame::ComputeBoW / KeyFrame::ComputeBoW                    src/Frame.cc:1033-1049
 *   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)            src/ORBmatcher.cc:230-382
 *   ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)         src/ORBmatcher.cc:656-799
 * ------------------------------------------------------------------------------------------------ */
typedef struct orbx_vocabulary orbx_vocabulary;

/* The node table of the vocabulary as ORBVocabulary::loadFromTextFile reads it (TemplatedVocabulary.h:1336-1424): row i describes
 * node i + 1 (node 0 is the root): parent id, leaf flag, 32-byte descriptor, weight.  Children keep file order; word ids are given to
 * the leaves in file order.  weighting: 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY; scoring: 0 L1_NORM, 1 L2_NORM, 2 CHI_SQUARE, 3 KL,
 * 4 BHATTACHARYYA, 5 DOT_PRODUCT (BowVector.h:25-49).  The tree is copied to the device (35 MB for ORBvoc: it stays in L2). */
int  orbx_vocabulary_create(int device, int k, int L, int weighting, int scoring, int n_nodes, const int* parent,
                            const uint8_t* is_leaf, const uint8_t* descriptors, const double* weights, orbx_vocabulary** out);
/* The same from the text file ORBVocabulary::loadFromTextFile reads (src/System.cc:84; first line "k L scoring weighting", then one
 * line per node "parent is_leaf d0 ... d31 weight").  A trailing empty line is ignored (the reference's reader turns it into a
 * childless, weightless extra child of the root with an all-zero descriptor, which no ORB descriptor descends into). */
int  orbx_vocabulary_load_text(int device, const char* path, orbx_vocabulary** out);
void orbx_vocabulary_destroy(orbx_vocabulary* v);
int  orbx_vocabulary_words(const orbx_vocabulary* v);      /* size() */

/* mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, levelsup) for n descriptors (host pointers).
 * word_of / node_of [n] (may be NULL): the per-feature word id and the id of its ancestor on level L - levelsup.
 * mBowVec: bow_ids / bow_values [capacity n] in std::map order (ascending word id), *n_bow entries, weights normalised as the
 * scoring type asks.  mFeatVec: fv_nodes [n] ascending, fv_offsets [n + 1], fv_indices [n]: node q holds the feature indices
 * fv_indices[fv_offsets[q] .. fv_offsets[q + 1]) in ascending order; *n_fv nodes.  Stopped words (weight <= 0) are in neither. */
int orbx_vocabulary_transform(orbx_vocabulary* v, const uint8_t* descriptors, int n, int levelsup, int* word_of, int* node_of,
                              int* bow_ids, double* bow_values, int* n_bow, int* fv_nodes, int* fv_offsets, int* fv_indices, int* n_fv);

/* One side of SearchByBoW: keypoints (mvKeysUn of a KeyFrame, mvKeys of a Frame: only .angle is read), descriptors, valid[i] != 0 iff
 * the feature holds a map point that is not bad (side 1 always; side 2 only in the KeyFrame x KeyFrame form, NULL = all valid), and
 * the feature vector as orbx_vocabulary_transform returns it. */
typedef struct orbx_bow_side {
    int n; const orbx_keypoint* keys; const uint8_t* descriptors; const uint8_t* valid;
    int n_fv; const int* fv_nodes; const int* fv_offsets; const int* fv_indices;
} orbx_bow_side;

/* kf_kf == 0: SearchByBoW(pKF = s1, F = s2): match21[j] = index of the KeyFrame feature whose map point goes to vpMapPointMatches[j].
 * kf_kf != 0: SearchByBoW(pKF1 = s1, pKF2 = s2): match12[i] = index of the pKF2 feature whose map point goes to vpMatches12[i].
 * Both arrays are always filled (-1 = no match); *nmatches = return value. */
int orbx_search_by_bow(orbx_matcher* m, int kf_kf, const orbx_bow_side* s1, const orbx_bow_side* s2, int* match12, int* match21, int* nmatches);

/* int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12, vector<pair<size_t,size_t>> &vMatchedPairs, bool bOnlyStereo)
 *   src/ORBmatcher.cc:810-1010 (LocalMapping::CreateNewMapPoints), with CheckDistEpipolarLine :188-215.
 *   s1 / s2: both KeyFrames as bag-of-words sides (mvKeysUn, mDescriptors, mFeatVec); valid[i] != 0 iff the feature holds NO map point
 *   (:843-845, :862).  u_right1 / u_right2: mvuRight (stereo iff >= 0).  F12: the 3 x 3 fundamental matrix, row-major float.
 *   (ex, ey): the epipole of camera 1 in image 2 as the caller computes it (:818-825).  scale_factors2 / level_sigma2_2: pKF2->mvScaleFactors
 *   and mvLevelSigma2.  Inside a common vocabulary node every side-1 feature takes, among the still unmatched side-2 features with distance
 *   <= TH_LOW that are not too close to the epipole (both monocular) and lie on the epipolar line (3.84 sigma^2), the one with the smallest
 *   distance -- the last such one in list order on ties, as the reference's running comparison does; then the rotation histogram.
 *   match12[i] = index in pKF2 (-1 = none): vMatchedPairs is its list of (i, match12[i]) in ascending i.  *nmatches = return value. */
int orbx_search_for_triangulation(orbx_matcher* m, const orbx_bow_side* s1, const orbx_bow_side* s2, const float* u_right1, const float* u_right2,
                                  const float* F12, float ex, float ey, int nlevels2, const float* scale_factors2, const float* level_sigma2_2,
                                  int only_stereo, int* match12, int* nmatches);

/* void MapPoint::ComputeDistinctiveDescriptors()   src/MapPoint.cc:359-439, for n_points map points at once.
 *   Point p owns the descriptors [offsets[p], offsets[p + 1]) of `descriptors` (32 bytes each): those of its observations in non-bad
 *   KeyFrames, in the iteration order of its mObservations map.  best_idx[p] = the index, inside that range, of the descriptor whose sorted
 *   distance row has the smallest vDists[0.5 * (N - 1)] (first one on ties); -1 for a point without descriptors.  The caller then sets
 *   mDescriptor = vDescriptors[best_idx].clone(). */
int orbx_distinctive_descriptors(orbx_matcher* m, int n_points, const int* offsets, const uint8_t* descriptors, int* best_idx);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_B200_H */
