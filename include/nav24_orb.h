/* nav24_orb.h — C ABI of the B200-native ORB front end (drop-in boundary for m-dayani/nav24).
 *
 * One shared library, libnav24orb.so, built from the .cu files under nav24_b200/csrc for sm_100a.  Plain C types
 * only: no torch, no OpenCV, no C++ in the signatures.  Every function returns an int status:
 * >= 0 success (detect: the reference's `monoIndex` return value), < 0 a NAV24_E_* code; nothing
 * throws or aborts across this boundary.  There is NO CPU fallback: without a CUDA device
 * nav24_orb_create fails with NAV24_E_CUDA.
 *
 * Reference interfaces replaced (paths relative to the reference's core/):
 *   OP::FtDt / OP::FtDtOrbSlam      operators/objDetection/OP_FtDt.hpp:14-29,
 *                                   operators/objDetection/OP_FtDtOrbSlam.hpp:27-84
 *   OP::FtAssoc / FtAssocOrbSlam    operators/objAssoc/OP_FtAssoc.hpp:15-22,
 *                                   operators/objAssoc/OP_FtAssocOrbSlam.hpp:13-38
 *   OB::FeatureGrid                 sensorData/observation/FeatureGrid.hpp:20-51
 * INTEGRATION.md shows the C++ subclasses a nav24 maintainer adds on top of these entry points.
 */
#ifndef NAV24_ORB_H
#define NAV24_ORB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NAV24_ABI_VERSION 2

/* error codes */
#define NAV24_OK 0
#define NAV24_E_BADARG (-1)    /* null pointer / empty image; detect() of the reference returns -1 here too
                                  (OP_FtDtOrbSlam.cpp:851-852) */
#define NAV24_E_GEOMETRY (-2)  /* image too small: some level has no 35-px FAST cell, or the quadtree has no
                                  root node (the reference divides by zero there, OP_FtDtOrbSlam.cpp:743-749,505);
                                  also a scale factor above ~1.85 (the source tile of 128 destination pixels must
                                  fit one 256-byte TMA box) or an image side of 8192 px or more */
#define NAV24_E_CAPACITY (-3)  /* caller's output capacity too small; n_out holds the required size */
#define NAV24_E_OVERFLOW (-4)  /* an internal device buffer overflowed (raw FAST corners); raise
                                  nav24_orb_params.raw_keys_per_kpx and retry.  Never silently truncates. */
#define NAV24_E_CUDA (-5)      /* CUDA runtime/driver error; see nav24_last_error_string */
#define NAV24_E_NOMEM (-6)

typedef struct nav24_orb nav24_orb; /* opaque: one CUDA device, its streams and device buffers */

/* Mirrors the five YAML keys of FtDt::create (OP_FtDt.cpp:31-46) + the scale factor (always 1.2f
 * in the reference because of the key-name bug at OP_FtDt.cpp:45-46). */
typedef struct nav24_orb_params {
    int32_t n_features;       /* nFeatures, default 1000 */
    float scale_factor;       /* 1.2f */
    int32_t n_levels;         /* nLevels, default 8 (1..16) */
    int32_t ini_th_fast;      /* iniThFast, default 20 */
    int32_t min_th_fast;      /* minThFast, default 7 */
    int32_t raw_keys_per_kpx; /* device capacity for raw FAST corners per 1000 level pixels; 0 = default (125,
                                 i.e. one corner per 8 px, the 3x3-NMS packing bound is 250) */
} nav24_orb_params;

/* = cv::KeyPoint field order (28 bytes).  pt in level-0 pixel coordinates. */
typedef struct nav24_kp {
    float x, y;
    float size;
    float angle;    /* degrees, [0,360) */
    float response; /* FAST score */
    int32_t octave;
    int32_t class_id; /* always -1 */
} nav24_kp;

/* FeatureGrid::setImageBounds (FeatureGrid.cpp:100-113): cols = W/10, rows = H/10, bounds from
 * Calibration::computeImageBounds.  Passed explicitly: no process-global state. */
typedef struct nav24_grid_cfg {
    int32_t cols, rows;
    float min_x, max_x, min_y, max_y;
} nav24_grid_cfg;

/* Camera model of Calibration::undistort (core/sensor/camera/Calibration.cpp:135-149), the step between detect and
 * matchV (core/frontEnd/FE_SlamMonoV.cpp:104-122).  K and D are the float matrices the reference hands to OpenCV
 * (models/GeometricCamera.h:62-66: K = (fx 0 cx; 0 fy cy; 0 0 1), D = 4 coefficients, R = I, P = K):
 *   NAV24_CAM_PINHOLE  identity                                   (models/Pinhole.hpp:75-78)
 *   NAV24_CAM_RADTAN   cv::undistortPoints, d = k1 k2 p1 p2         (models/PinholeRadTan.cpp:11-25)
 *   NAV24_CAM_KB8      cv::fisheye::undistortPoints, d = k1..k4     (models/KannalaBrandt8.cpp:231-243) */
#define NAV24_CAM_PINHOLE 0
#define NAV24_CAM_RADTAN 1
#define NAV24_CAM_KB8 2
typedef struct nav24_camera {
    int32_t model;
    float fx, fy, cx, cy;
    float d[4];
} nav24_camera;

/* ---- life cycle ------------------------------------------------------------------------- */
int nav24_abi_version(void);
/* Replaces FtDtOrbSlam::FtDtOrbSlam (OP_FtDtOrbSlam.cpp:441-500). */
int nav24_orb_create(const nav24_orb_params* params, int device, nav24_orb** out);
void nav24_orb_destroy(nav24_orb* ctx);
/* Replaces FtDtOrbSlam::setNumFeatures (OP_FtDtOrbSlam.cpp:962-976): recomputes the per-level quotas. */
int nav24_orb_set_num_features(nav24_orb* ctx, int n_features);
int nav24_orb_get_num_features(const nav24_orb* ctx);
/* GetScaleFactors / GetInverseScaleFactors / per-level quotas (OP_FtDtOrbSlam.hpp:35-55). Arrays of n_levels. */
int nav24_orb_get_tables(const nav24_orb* ctx, float* scale, float* inv_scale, int32_t* features_per_level);
const char* nav24_last_error_string(const nav24_orb* ctx);

/* ---- detector --------------------------------------------------------------------------- */
/* Replaces FtDtOrbSlam::detect (OP_FtDtOrbSlam.cpp:844-934) for one host image (8-bit grey).
 * kps/desc: host buffers for `cap` keypoints / cap*32 bytes, filled in the reference's two-ended
 * output order.  Returns monoIndex (>= 0) or a NAV24_E_* code. */
int nav24_orb_detect(nav24_orb* ctx, const uint8_t* gray, int width, int height, size_t stride_bytes,
                     nav24_kp* kps, uint8_t* desc, int cap, int* n_out);

/* Batched form: n_frames images of identical shape, frame f at gray + f*frame_stride_bytes (host memory;
 * pinned memory makes the copies asynchronous).  Outputs are [n_frames][cap]; n_out / mono_out are
 * [n_frames].  Frames are independent (same results as n_frames calls of nav24_orb_detect). */
int nav24_orb_detect_batch(nav24_orb* ctx, const uint8_t* gray, int n_frames, int width, int height,
                           size_t stride_bytes, size_t frame_stride_bytes, nav24_kp* kps, uint8_t* desc,
                           int cap, int* n_out, int* mono_out);

/* Device-resident form: frames already in HBM (device pointer, 16-byte aligned base and strides are
 * used in place, anything else is copied once).  Results stay on the device inside the context until
 * fetched; the call returns after enqueueing (no host synchronisation). */
int nav24_orb_detect_device(nav24_orb* ctx, const uint8_t* d_gray, int n_frames, int width, int height,
                            size_t stride_bytes, size_t frame_stride_bytes);
/* Fused detect + window matching, the stereo / frame-to-frame form of the front end (FE_SlamMonoV.cpp:104-122 calls
 * detect and then match on every frame).  pairs_ab = n_pairs x (a, b) frame indices of this batch; matching uses the
 * keypoints as detected (identity undistortion, Pinhole.hpp:75-78).  The batch is processed in chunks on two compute
 * streams: the latency-bound tail of one chunk (quadtree, the sequential part of the matcher) overlaps the head of
 * the next, and in the host form the copies overlap the kernels.  Host form: synchronous, outputs as in
 * nav24_orb_detect_batch plus matches12 [n_pairs][mcap] (mcap >= nav24_orb_max_keypoints) and n_matches [n_pairs]. */
int nav24_orb_detect_match_batch(nav24_orb* ctx, const uint8_t* gray, int n_frames, int width, int height,
                                 size_t stride_bytes, size_t frame_stride_bytes, nav24_kp* kps, uint8_t* desc, int cap,
                                 int* n_out, int* mono_out, int n_pairs, const int* pairs_ab, const nav24_grid_cfg* grid,
                                 float window, float nnratio, int th_low, int check_ori, int32_t* matches12, int mcap,
                                 int* n_matches);
/* Device-resident form: frames in HBM (16-byte aligned base / strides required), returns after enqueueing; results
 * stay on the device until nav24_orb_fetch / nav24_match_fetch. */
int nav24_orb_detect_match_device(nav24_orb* ctx, const uint8_t* d_gray, int n_frames, int width, int height,
                                  size_t stride_bytes, size_t frame_stride_bytes, int n_pairs, const int* pairs_ab,
                                  const nav24_grid_cfg* grid, float window, float nnratio, int th_low, int check_ori);
/* Waits for the last fused call and copies its matches out; returns the total number of matches. */
int nav24_match_fetch(nav24_orb* ctx, int32_t* matches12, int mcap, int* n_matches);
/* Waits for the last nav24_orb_detect_device and copies its results out. kps/desc may be NULL (counts only). */
int nav24_orb_fetch(nav24_orb* ctx, nav24_kp* kps, uint8_t* desc, int cap, int* n_out, int* mono_out);
/* The same for frames [first_frame, first_frame + n_frames) / pairs [first_pair, first_pair + n_pairs) of the last call:
 * a caller that keeps many sequences resident (BASELINE.json configs[4]: 64 sequences x 100 frames) reads the results
 * back one sequence at a time instead of holding the whole batch in host memory. */
int nav24_orb_fetch_range(nav24_orb* ctx, int first_frame, int n_frames, nav24_kp* kps, uint8_t* desc, int cap,
                          int* n_out, int* mono_out);
int nav24_match_fetch_range(nav24_orb* ctx, int first_pair, int n_pairs, int32_t* matches12, int mcap, int* n_matches);
int nav24_orb_sync(nav24_orb* ctx);
/* Upper bound of keypoints per frame for the current n_features (quota + 3 per level). */
int nav24_orb_max_keypoints(const nav24_orb* ctx);

/* Parity / debug accessors on the results of the last detect call (frame index within the batch).
 * Level pixels of ComputePyramid (OP_FtDtOrbSlam.cpp:936-960) without the unread 19-px border;
 * which=0 pyramid level, which=1 blurred level (GaussianBlur at :890-891). dst may be NULL to query w/h. */
int nav24_orb_get_level(nav24_orb* ctx, int frame, int level, int which, uint8_t* dst, size_t dst_stride,
                        int* w, int* h);
/* vToDistributeKeys of a level (OP_FtDtOrbSlam.cpp:807-815) in the reference's order: (x, y, response)
 * triplets relative to (minBorderX, minBorderY). Returns the count; xyr may be NULL. */
int nav24_orb_get_raw_keys(nav24_orb* ctx, int frame, int level, float* xyr, int cap);
/* Keypoints of one level after DistributeOctTree + orientation, level coordinates, quadtree order. */
int nav24_orb_get_level_keypoints(nav24_orb* ctx, int frame, int level, nav24_kp* kps, int cap);
/* Device time (ms, CUDA events on the context's stream) of the stages of the last nav24_orb_detect_device call (the
 * only entry point that records them: one chunk on one stream, nothing overlapped); NAV24_E_BADARG when the last
 * detect call was another one:  [0] pyramid, [1] FAST, [2] quadtree, [3] order+orientation+blur+descriptors, [4] total. */
int nav24_orb_stage_ms(nav24_orb* ctx, float* ms5);
/* Same stages summed over the detect calls since the last reset (at most the last 64); *calls = how many. */
int nav24_orb_stage_ms_sum(nav24_orb* ctx, float* ms5, int* calls, int reset);
/* Number of kernels launched by this context so far. */
long long nav24_orb_launch_count(const nav24_orb* ctx);
/* CUDA-event stopwatch on the context's own stream (the stream every kernel of this context runs on):
 * start records an event; stop records a second one, waits for it and returns the elapsed device ms. */
int nav24_orb_timer_start(nav24_orb* ctx);
int nav24_orb_timer_stop(nav24_orb* ctx, float* ms);

/* ---- matchers --------------------------------------------------------------------------- */
/* Replaces FtAssocOrbSlam::matchV (OP_FtAssocOrbSlam.cpp:91-223) + FeatureGrid (FeatureGrid.cpp:20-152).
 * ud*_xy: undistorted (x,y) pairs as produced by Calibration::undistort (identity for pinhole).
 * matches12: n1 ints, -1 or an index into frame 2.  Reference defaults: window 100, nnratio 0.6,
 * th_low 50, check_ori 1, levels 0..0.  Returns the number of matches. */
int nav24_match_window(nav24_orb* ctx, const nav24_kp* k1, const float* ud1_xy, const uint8_t* d1, int n1,
                       const nav24_kp* k2, const float* ud2_xy, const uint8_t* d2, int n2,
                       const nav24_grid_cfg* grid, float window, float nnratio, int th_low, int check_ori,
                       int32_t* matches12);

/* Batched pairs, all host memory: pair p has n1[p]/n2[p] keypoints stored at offset p*cap in each array. */
int nav24_match_window_batch(nav24_orb* ctx, int n_pairs, int cap, const nav24_kp* k1, const float* ud1_xy,
                             const uint8_t* d1, const int* n1, const nav24_kp* k2, const float* ud2_xy,
                             const uint8_t* d2, const int* n2, const nav24_grid_cfg* grid, float window,
                             float nnratio, int th_low, int check_ori, int32_t* matches12, int* n_matches);

/* Device-resident: matches frame a against frame b of the LAST detect batch (identity undistortion), no
 * host round trip of keypoints/descriptors.  pairs = n_pairs x (a,b) frame indices. matches12 is
 * [n_pairs][cap] on the host, cap >= nav24_orb_max_keypoints.  With matches12 == NULL and n_matches == NULL the call only
 * enqueues the kernel (no host synchronisation; results stay on the device). */
int nav24_match_window_frames(nav24_orb* ctx, int n_pairs, const int* pairs_ab, const nav24_grid_cfg* grid,
                              float window, float nnratio, int th_low, int check_ori, int32_t* matches12,
                              int cap, int* n_matches);

/* ---- undistortion (keeps the frame on the device between detect and matchV) ------------- */
/* Replaces Calibration::undistort / GeometricCamera::UndistortKeyPoints for n points in host memory: xy and ud_xy are
 * n x (x, y) floats (may alias).  Double-precision restatement of the OpenCV routines, results rounded to float like
 * cv::Point2f.  Also what Calibration::computeImageBounds (Calibration.cpp:196-228) feeds the four image corners to. */
int nav24_undistort_points(nav24_orb* ctx, const nav24_camera* cam, const float* xy, int n, float* ud_xy);
/* Camera of the fused detect + match entry points and of nav24_match_window_frames: with a model other than pinhole
 * every detect call also undistorts its keypoints on the device and the matchers consume those coordinates (grid
 * cells, window test), exactly like matchV reading getPointUd().  cam == NULL resets to pinhole. */
int nav24_orb_set_camera(nav24_orb* ctx, const nav24_camera* cam);
/* Undistorted coordinates of the last detect batch: ud_xy is [n_frames][cap][2] floats (cap >= keypoints per frame;
 * rows beyond a frame's count are left untouched).  Synchronises. */
int nav24_orb_fetch_undistorted(nav24_orb* ctx, float* ud_xy, int cap);

/* Intended semantics of FtAssocOCV::match (OP_FtAssoc.cpp:63-99): brute-force kNN-2 + ratio test.
 * norm 0 = Hamming (popcount), 1 = L2 on the u8 bytes (cv::DescriptorMatcher::BRUTEFORCE default, :20).
 * Lowest train index wins ties.  pass[i] = dist0 < ratio*dist1 (reference ratio 0.7). */
#define NAV24_NORM_HAMMING 0
#define NAV24_NORM_L2_U8 1
int nav24_match_bf_knn2(nav24_orb* ctx, const uint8_t* d1, int n1, const uint8_t* d2, int n2, int norm,
                        float ratio, int32_t* idx0, int32_t* idx1, float* dist0, float* dist1, uint8_t* pass);

/* Device time (CUDA events) of the kernels of the last nav24_match_bf_knn2 call, for tools/bench_bf.py. */
int nav24_debug_last_kernel_ms(nav24_orb* ctx, float* ms);
/* Test hook for the quadtree's "largest first" ordering (std::sort at OP_FtDtOrbSlam.cpp:646, comparator :358-373):
 * sorts n records by key with the device restatement of libstdc++'s introsort and returns the permutation
 * (perm[i] = original index of the record at sorted position i).  Equal keys are "equivalent": their order is
 * whatever std::sort leaves, which the device code must reproduce exactly. */
int nav24_debug_sort_u32(nav24_orb* ctx, const uint32_t* keys, int n, int32_t* perm);

/* ---- image ingest ------------------------------------------------------------------------ */
/* A ring of pinned host slots the camera decodes into directly, replacing the four full-image host copies the
 * reference makes per frame (cv::imread + image.clone() at core/sensor/camera/Camera.cpp:454-455, the clone in
 * core/sensorData/Image.hpp:22, two more clones + cv::cvtColor at core/frontEnd/FE_SlamMonoV.cpp:90-94): the frame
 * goes from the decoder's output buffer to the device in one asynchronous copy, and colour frames (channels = 3,
 * interleaved BGR as cv::imread / cv::imdecode produce) are converted to grey ON THE DEVICE with cv::cvtColor's
 * COLOR_BGR2GRAY fixed point, so detect sees CV_8UC1 as OP_FtDtOrbSlam.cpp:853 asserts.  channels = 1: grey. */
typedef struct nav24_ingest nav24_ingest;
int nav24_ingest_create(nav24_orb* ctx, int width, int height, int channels, int n_slots, nav24_ingest** out);
void nav24_ingest_destroy(nav24_ingest* ring);
/* Host address of slot `slot` (height rows of width*channels bytes, tightly packed); NULL when out of range. */
uint8_t* nav24_ingest_slot(nav24_ingest* ring, int slot);
size_t nav24_ingest_slot_bytes(const nav24_ingest* ring);
/* detect (and, with pairs, windowed matching: pairs_ab index the frames of THIS call) on slots [first_slot,
 * first_slot + n_frames); arguments and results as nav24_orb_detect_batch / nav24_orb_detect_match_batch.  The grey
 * level 0 the device computed is readable with nav24_orb_get_level(ctx, frame, 0, 0, ...). */
int nav24_ingest_detect(nav24_ingest* ring, int first_slot, int n_frames, nav24_kp* kps, uint8_t* desc, int cap,
                        int* n_out, int* mono_out);
int nav24_ingest_detect_match(nav24_ingest* ring, int first_slot, int n_frames, nav24_kp* kps, uint8_t* desc, int cap,
                              int* n_out, int* mono_out, int n_pairs, const int* pairs_ab, const nav24_grid_cfg* grid,
                              float window, float nnratio, int th_low, int check_ori, int32_t* matches12, int mcap,
                              int* n_matches);

/* ---- two-view RANSAC scoring -------------------------------------------------------------- */
/* Replaces the scoring half of TwoViewReconstruction::FindHomography / FindFundamental
 * (core/operators/mapInit/OP_2ViewReconstruction.cpp:266-365), which the reference runs in two std::threads (:133-134):
 * CheckHomography (:447-530) and CheckFundamental (:532-610) of ALL n_hyp RANSAC hypotheses in one launch.
 *   xy1, xy2      n_matches x (x, y): the matched keypoints in match order (mvKeys1[mvMatches12[i].first].pt, ...second)
 *   H21, H12      n_hyp x 9 row-major: T2inv*Hn*T1 and its inverse per iteration (:302-303); NULL skips the homography
 *   F21           n_hyp x 9 row-major: T2t*Fn*T1 (:354); NULL skips the fundamental matrix
 *   sigma, th_h, th_f, th_score   mSigma, mThChiSqScore (5.991), mThChiSqF (3.841), mThChiSqScore
 *   score_h/f     n_hyp floats, bit-equal to the reference's sequential float sums (IEEE single, no contraction)
 *   inliers_h/f   optional n_hyp x n_matches bytes (vbCurrentInliers of every iteration)
 *   best_h/f      optional: the iteration the reference's `if (currentScore > score)` loop keeps (-1: no score above 0)
 * The minimal-set solvers (ComputeH21 / ComputeF21: 8-point SVDs) stay on the host. */
int nav24_two_view_score(nav24_orb* ctx, const float* xy1, const float* xy2, int n_matches, const float* H21,
                         const float* H12, const float* F21, int n_hyp, float sigma, float th_h, float th_f,
                         float th_score, float* score_h, float* score_f, uint8_t* inliers_h, uint8_t* inliers_f,
                         int* best_h, int* best_f);
/* The same scoring, returning what FindHomography / FindFundamental hand back (:266-365): the scores of all iterations, the
 * kept iteration, and ONLY ITS inlier mask (kept_inliers_h/f: n_matches bytes each, all 0 when no iteration is kept) —
 * n_matches bytes per model over PCIe instead of n_hyp x n_matches (2000 matches, host buffers in and out: 0.12 ms per call instead of 0.29). */
int nav24_two_view_score_kept(nav24_orb* ctx, const float* xy1, const float* xy2, int n_matches, const float* H21,
                              const float* H12, const float* F21, int n_hyp, float sigma, float th_h, float th_f,
                              float th_score, float* score_h, float* score_f, uint8_t* kept_inliers_h,
                              uint8_t* kept_inliers_f, int* best_h, int* best_f);

/* ---- memory helpers (so that a C/C++ host needs no CUDA headers) --------------------------- */
int nav24_host_alloc(size_t bytes, void** out);   /* pinned host memory: makes detect_batch copies asynchronous */
int nav24_host_free(void* p);
int nav24_device_alloc(size_t bytes, void** out);
int nav24_device_free(void* p);
int nav24_memcpy_h2d(void* dst, const void* src, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* NAV24_ORB_H */
