/*
 * vstab.h -- C ABI of libvstab.so, the B200 (sm_100a) hot path of ComfyUI-Video-Stabilizer.
 *
 * Every entry point replaces one OpenCV call site of the reference (file:line relative to the
 * reference repository root) and is what a ctypes / cffi / pybind stub on the reference side
 * would bind.
 *
 * Conventions
 *   - plain C types only; no CUDA or torch types in any signature (`stream` is a cudaStream_t
 *     passed as void*, NULL = legacy default stream);
 *   - every pointer named *_dev is a DEVICE pointer owned by the caller; the library owns only
 *     the workspace inside its handle.  *_host pointers are small host arrays read before the
 *     call returns;
 *   - all work is enqueued on the caller's stream; no entry point synchronises unless its
 *     comment says so;
 *   - return value 0 = VSTAB_OK, negative = error; vstab_last_error() gives the text.  Nothing
 *     throws across the boundary;
 *   - images are row-major, channel-interleaved: IMAGE = float32 [N][H][W][3] in 0..1
 *     (ComfyUI IMAGE layout, nodes/stabilizer_utils.py:200-221), gray = uint8 [N][h][w].
 */
#ifndef VSTAB_H_
#define VSTAB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSTAB_ABI_VERSION 1

#if defined(__GNUC__)
#define VSTAB_API __attribute__((visibility("default")))
#else
#define VSTAB_API
#endif

#define VSTAB_OK 0
#define VSTAB_ERR_INVALID (-1)     /* bad argument */
#define VSTAB_ERR_CUDA (-2)        /* CUDA runtime error, text in vstab_last_error */
#define VSTAB_ERR_NOMEM (-3)       /* workspace allocation failed */
#define VSTAB_ERR_UNSUPPORTED (-4) /* shape outside what the kernels were built for */

/* interpolation (nodes/motion_apply.py:24-29 _interpolation_flag) */
#define VSTAB_INTERP_BILINEAR 0 /* cv2.INTER_LINEAR */
#define VSTAB_INTERP_BICUBIC 1  /* cv2.INTER_CUBIC  */

/* padding-mask footprint rule of warpPerspective(ones, INTER_NEAREST) */
#define VSTAB_MASK_RULE_P 0 /* closed rectangle on continuous coordinates (cv2 4.13 IPP HAL) */
#define VSTAB_MASK_RULE_C 1 /* classic: round-half-even, then range check */
/* What cv2 itself does per call: Rule P, unless one of the wheel's min(threads, ceil(W'H'/2^14)) destination
 * stripes misses the source frame entirely -- then Rule C for that whole (frame, sample).  The thread count of the
 * reference machine (cv2.getNumThreads(), by default its logical CPU count) rides in the upper bits of mask_rule:
 * VSTAB_MASK_RULE_AUTO_THREADS(8).  With 1 thread AUTO == Rule P. */
#define VSTAB_MASK_RULE_AUTO 2
#define VSTAB_MASK_RULE_AUTO_THREADS(t) (VSTAB_MASK_RULE_AUTO | ((t) << 8))

/* source-tile staging of the fused resampler (debug / A-B switch) */
#define VSTAB_STAGE_AUTO 0   /* shared-memory tile when the footprint fits, else global gather */
#define VSTAB_STAGE_GLOBAL 1 /* always gather through L1 */

/* transform models (nodes/stabilizer_utils.py:15 TransformMode) */
#define VSTAB_MODE_TRANSLATION 0
#define VSTAB_MODE_SIMILARITY 1
#define VSTAB_MODE_PERSPECTIVE 2

typedef struct vstab_handle vstab_handle;

/* ---- lifecycle ------------------------------------------------------------------------- */

VSTAB_API int vstab_abi_version(void);

/* One handle per (device, host thread).  Allocates nothing on the device until first use. */
VSTAB_API int vstab_create(int device, vstab_handle** out);
VSTAB_API void vstab_destroy(vstab_handle* h);

/* Text of the last error recorded on this handle (or process-wide when h == NULL). */
VSTAB_API const char* vstab_last_error(const vstab_handle* h);

/* Number of kernels this handle has launched since creation (bench.py "gpu_launches"). */
VSTAB_API uint64_t vstab_launch_count(const vstab_handle* h);

/* ---- K1 + K2 : gray + working-size downscale ------------------------------------------- */

/*
 * Working (estimation) size rule, nodes/stabilizer_utils.py:248-268 _working_estimation_size:
 * longest side capped at 960.  Writes the size actually used (== input size when no resize
 * happens).  Host-only helper.
 */
VSTAB_API int vstab_working_size(int width, int height, int* work_w, int* work_h);

/*
 * Replaces nodes/stabilizer_utils.py:236-242 (_make_gray: cv2.cvtColor(RGB2GRAY), *255, clip,
 * truncating uint8 cast) and :271-276 (_make_gray_for_estimation: cv2.resize INTER_AREA).
 * rgb_dev  [n][h][w][3] float32, gray_dev [n][work_h][work_w] uint8.
 * Bit-exact against cv2 4.13: Y = fma(B,.114f, fma(R,.299f, G*.587f)); INTER_AREA x2
 * ((a+b+c+d+2)>>2), integer xK (rint(sum/K^2)), and the general fractional-coverage path.
 */
VSTAB_API int vstab_gray_working(vstab_handle* h, const float* rgb_dev, int n, int height, int width,
                       uint8_t* gray_dev, int work_h, int work_w, void* stream);

/* ---- K10 + K11 + K12 : fused inverse-map resampler -------------------------------------- */

/*
 * Replaces, in ONE launch for the whole batch,
 *   nodes/video_stabilizer_flow.py:560-588, nodes/video_stabilizer_classic.py:491-519,
 *   nodes/motion_apply.py:75-122 (_warp_with_matrices) and :137-202 (_warp_with_motion_blur):
 *     cv2.warpPerspective(frame, M, out, INTER_LINEAR|INTER_CUBIC, BORDER_CONSTANT, border)
 *     cv2.warpPerspective(ones,  M, out, INTER_NEAREST, BORDER_CONSTANT, 0) -> >0.5 -> 1-x
 *     and the S-sample f32 accumulate / S.
 *
 * src_dev      [n][src_h][src_w][3] float32
 * fwd_dev      [n][samples][9] float32 FORWARD matrices exactly as the reference hands them to
 *              cv2 (row-major 3x3).  The kernel inverts them in double by cofactors like
 *              cv::invert, evaluates source coordinates in double, quantises to 1/32 px with
 *              round-half-even and uses cv2's f32 weight tables and per-tap constant border.
 * samples      1 (no blur) or 3..33 (motion_apply.py:166)
 * border_host  3 floats = padding_rgb / 255 as float32 (motion_apply.py:70-72)
 * dst_dev      [n][out_h][out_w][3] float32
 * mask_dev     [n][out_h][out_w] float32, 1 = padding (0/1 when samples == 1, 1 - count/S
 *              otherwise); may be NULL (masks_zero path of motion_apply.py:102-105)
 * pad_count_dev [n] uint32, number of mask pixels > 1e-3 per frame (for padding_fraction_*,
 *              video_stabilizer_flow.py:585-587); may be NULL; zeroed by the call.
 */
VSTAB_API int vstab_warp_fused(vstab_handle* h, const float* src_dev, int n, int src_h, int src_w,
                     const float* fwd_dev, int samples, int interp, int out_h, int out_w,
                     const float* border_host, int mask_rule, int stage_mode, float* dst_dev,
                     float* mask_dev, uint32_t* pad_count_dev, void* stream);

/*
 * Coverage only (crop solvers, nodes/stabilizer_utils.py:604-656, :759-780 and
 * nodes/motion_apply.py:205-228 _common_valid_mask): AND-reduces the INTER_NEAREST coverage of
 * n matrices into common_dev [out_h][out_w] uint8 (1 = valid in every frame).
 */
VSTAB_API int vstab_common_coverage(vstab_handle* h, const float* fwd_dev, int n, int src_h, int src_w,
                          int out_h, int out_w, int mask_rule, uint8_t* common_dev, void* stream);

/*
 * Crop solver helper (nodes/stabilizer_utils.py:604-656 finalize_with_masks): for every matrix, the
 * bounding box of erode3x3(dilate3x3(INTER_NEAREST coverage > 0.5)).
 * bbox_dev [n][4] int32 = {xmin, ymin, xmax, ymax}; xmax < 0 when the closed coverage is empty.
 */
VSTAB_API int vstab_coverage_bbox(vstab_handle* h, const float* fwd_dev, int n, int src_h, int src_w,
                                  int out_h, int out_w, int mask_rule, int32_t* bbox_dev, void* stream);

/*
 * Padding mask of a single-sample resampling (values exactly 0.0f / 1.0f) packed to one byte per pixel for the
 * device->host copy: the MASK output of nodes/video_stabilizer_flow.py:560-588 / classic.py:491-519 goes back as a
 * quarter of its float32 bytes and is widened on the host (pipeline.fused_warp, output="host").
 * mask_dev [n] float32, out_dev [n] uint8 (both 16-byte aligned), odd_dev: one uint32 the caller zeroed; bit 0 is set
 * when a value other than 0 / 1 was seen (a soft motion-blur mask must not take this route).
 */
VSTAB_API int vstab_mask_pack_u8(vstab_handle* h, const float* mask_dev, size_t n, uint8_t* out_dev, uint32_t* odd_dev,
                                 void* stream);

/*
 * K1 + K2 with the input adapter's range rule fused into the same read of the source.
 * Replaces, for float32 frames, the per-frame `arr.max() > 1.5  =>  arr /= 255.0` of
 * nodes/stabilizer_utils.py:96-147 (_to_numpy_frame) together with :236-276 (_make_gray_for_estimation):
 * the luma kernel notes per frame whether it saw an element > 1.5 (bit 0) or a NaN (bit 1; numpy's max() is then
 * NaN and the test false).  Frames with flags == 1 are 0..255 content: their working image is recomputed from
 * value / 255 and the frame is divided by 255 IN PLACE, all on `stream`, without a host round trip.
 * rgb_dev is therefore read AND written; flags_dev [n] uint32 is zeroed by the call.
 */
VSTAB_API int vstab_gray_working_adapt(vstab_handle* h, float* rgb_dev, int n, int height, int width, uint8_t* gray_dev,
                                       int work_h, int work_w, uint32_t* flags_dev, void* stream);

/* The range rule alone (Motion Apply and the legacy inverse have no estimation pass): one read for the flags, then
 * the in-place division of the frames with flags == 1.  rgb_dev [n][height][width][channels] float32. */
VSTAB_API int vstab_range_normalize(vstab_handle* h, float* rgb_dev, int n, int height, int width, int channels,
                                    uint32_t* flags_dev, void* stream);

/*
 * Host-side helper (no device work): out[i] = f(a[i] [, b[i]]) with glibc's libm, for the float64 trajectory maths that
 * stays on the host like in the reference (nodes/stabilizer_utils.py:300-358 _matrix_to_params / _params_to_matrix:
 * math.atan2, math.log, math.exp, math.cos, math.sin per frame).  Same libm as Python's math module => same bits.
 */
#define VSTAB_LIBM_ATAN2 0 /* atan2(a, b) */
#define VSTAB_LIBM_LOG 1
#define VSTAB_LIBM_EXP 2
#define VSTAB_LIBM_COS 3
#define VSTAB_LIBM_SIN 4
VSTAB_API int vstab_host_libm(int op, const double* a, const double* b, double* out, int n);

/*
 * Host-side trajectory solve (no device work), the common case of the path between the fit kernels and the resampler
 * in one pass each instead of some 80 numpy operations -- the GPU idles while it runs.  Same IEEE operations in the same
 * order and widths as the reference (and as hostmath.py, which stays the general path and the test oracle for these).
 *
 * vstab_host_trajectory: nodes/video_stabilizer_flow.py:156-210 / classic.py:104-158 (acceptance test of the requested
 * model), stabilizer_utils.py:279-297 (_rescale_transform_to_full), :300-326 (_matrix_to_params), flow.py:341-349 (cumsum).
 *   raw            [n_pairs][3] vstab_fit_result as copied from the device (12 eight-byte words each)
 *   detected       [n_pairs] corners found per pair (Classic: < 12 => identity) or NULL
 *   mode           VSTAB_MODE_* requested;  work_w/work_h = 0: estimation ran at full size (no rescale)
 *   matrices       out [n_pairs][9] float32 full-resolution per-pair transforms
 *   path           out [n_pairs + 1][K] float64 cumulative parameters, K = 2 / 4 / 8
 *   first_fallback out: n_pairs when every pair accepts the requested model; otherwise the first pair that does not --
 *                  the sticky fallback ladder starts there, the outputs are NOT written and the caller replays it.
 * vstab_host_framing: stabilizer_utils.py:329-358 (_params_to_matrix) and :1010-1032 (_compute_bounding_boxes).
 *   diffs [n_frames][K] float64 -> apply [n_frames][9] float32, mins / maxs [n_frames][2] float64 (corner bounding boxes),
 *   box [9]: inner rectangle x0 y0 x1 y1 (max of minima, min of maxima), union x0 y0 x1 y1, 1.0 if all matrices are
 *   finite with a (0, 0, 1) last row.
 * vstab_host_shift: [[1,0,ox],[0,1,oy],[0,0,1]] @ apply for such matrices (flow.py:471-489; the float32 product has one
 *   inexact operation per element there, so its bits do not depend on the BLAS); VSTAB_ERR_UNSUPPORTED otherwise.
 */
VSTAB_API int vstab_host_trajectory(const double* raw, const int32_t* detected, int n_pairs, int min_points, int mode,
                                    int src_w, int src_h, int work_w, int work_h, float* matrices, double* path,
                                    int* first_fallback);
VSTAB_API int vstab_host_framing(const double* diffs, int n_frames, int mode, int width, int height, float* apply,
                                 double* mins, double* maxs, double* box);
VSTAB_API int vstab_host_shift(const float* apply, int n_frames, float off_x, float off_y, float* out);
/* vstab_host_target: flow.py:351-374 with the box filter of stabilizer_utils.py:361-383: path [n_frames][n_params] ->
 * target = path + strength * (filter(path) - path) (zeros under camera lock) and diffs = target - path.  window = 0: no
 * filter; windows of more than 11 taps return VSTAB_ERR_UNSUPPORTED (numpy sums those through the BLAS of the machine). */
VSTAB_API int vstab_host_target(const double* path, int n_frames, int n_params, int window, double strength, int camera_lock,
                                double* target, double* diffs);

/* ---- K3 + K4 : DIS dense optical flow, batched over frame pairs ------------------------- */

/*
 * Replaces nodes/video_stabilizer_flow.py:76-87 (_create_flow_backend: PRESET_MEDIUM with
 * finest_scale 2, patch 8, stride 4, 25 GD iterations, 5 variational-refinement iterations,
 * mean normalisation, spatial propagation) and :140 backend.calc(prev, curr) for all pairs
 * (i, i+1), i in [0, n_frames-1), plus the 8-px grid sampling of :141-147.
 *
 * gray_dev   [n_frames][h][w] uint8 working-size frames (h, w <= 960 longest side)
 * flow_dev   [n_frames-1][h][w][2] float32 full working-size flow (cv2 layout) or NULL
 * grid_dev   [n_frames-1][gh][gw][2] float32 flow sampled at (x, y) = (gx*step, gy*step),
 *            gw = ceil(w/step), gh = ceil(h/step); or NULL
 */
VSTAB_API int vstab_dis_flow(vstab_handle* h, const uint8_t* gray_dev, int n_frames, int height, int width,
                   float* flow_dev, float* grid_dev, int grid_step, void* stream);

/*
 * The same for a run of pairs that does not start at the head of the clip (frame-range shards, streamed
 * chunks): first_pair_index is the clip-wide index of the pair (gray_dev[0], gray_dev[1]).
 * It matters on small frames only.  The reference keeps ONE backend object per clip
 * (nodes/video_stabilizer_flow.py:312) and cv2's calc() rewrites that object's finest scale when the frame
 * is too small for the configured one (longest side < ~91 px or shortest < 32 px: automatic scale selection,
 * the flow is then computed down to full resolution).  Pair 0 therefore runs with the automatically selected
 * levels and every later pair with the levels calc() derives from the rewritten state, which can be fewer
 * (90x50: levels 2..0, then 1..0).  vstab_dis_flow(...) == vstab_dis_flow_at(..., first_pair_index = 0, ...).
 * Sizes on which cv2 itself raises (< 12 px) or reads outside its coarsest level (that level smaller than one
 * 8x8 patch, e.g. 100x30) return VSTAB_ERR_UNSUPPORTED.
 */
VSTAB_API int vstab_dis_flow_at(vstab_handle* h, const uint8_t* gray_dev, int n_frames, int height, int width,
                                int first_pair_index, float* flow_dev, float* grid_dev, int grid_step, void* stream);

/* ---- K5 + K6 : Shi-Tomasi corners + pyramidal Lucas-Kanade, batched over frame pairs ---- */

/*
 * Replaces nodes/video_stabilizer_classic.py:76-83 (cv2.goodFeaturesToTrack(prev, maxCorners=400,
 * qualityLevel=0.01, minDistance=7, blockSize=21)) and :88-100 (cv2.calcOpticalFlowPyrLK(prev, curr,
 * features, winSize=(31,31), maxLevel=3, criteria=(EPS|COUNT, 50, 0.01)) + the status == 1 filter) for
 * all pairs (i, i+1), i in [0, n_frames-1).
 *
 * gray_dev      [n_frames][h][w] uint8 working-size frames
 * prev_dev      [n_frames-1][max_corners][2] float32 corner positions in frame i (NaN beyond the
 *               number of corners found)
 * curr_dev      [n_frames-1][max_corners][2] float32 tracked positions in frame i+1 (NaN when the
 *               track was lost: cv2 status 0); feed both to vstab_fit_batch
 * detected_dev  [n_frames-1] int32 number of corners found in frame i (classic.py:84 `< 12` test and
 *               the translation confidence tracked / detected of :157)
 */
VSTAB_API int vstab_gftt_lk(vstab_handle* h, const uint8_t* gray_dev, int n_frames, int height, int width,
                            int max_corners, float* prev_dev, float* curr_dev, int32_t* detected_dev,
                            void* stream);

/* ---- K4 + K7 + K8 + K9 : robust model fit, batched over frame pairs --------------------- */

typedef struct vstab_fit_result {
  double matrix[9];  /* 3x3 prev->curr at working resolution, before the reference's float32 cast */
  double residual;   /* mean |affine(prev) - curr| over both axes, flow.py:174,189,207 */
  int32_t n_inliers; /* RANSAC consensus size (similarity, perspective) or n_valid (translation) */
  int32_t n_valid;   /* finite correspondences */
  int32_t n_total;   /* correspondences offered */
  int32_t ok;        /* 1 = the estimator returned a model (cv2 result not None) */
} vstab_fit_result;

/*
 * Replaces nodes/video_stabilizer_flow.py:148-210 for every pair: finite filter, then ALL
 * three candidate models (translation = per-axis median, similarity =
 * estimateAffinePartial2D RANSAC 2.0 px / 2000 / 0.992, perspective = findHomography RANSAC
 * 2.5 px / 2000 / 0.992).  The sticky mode ladder (:161, :338-339) is replayed by the caller
 * over the table so that frame-range shards agree.
 *
 * prev_dev / curr_dev  [n_pairs][n_pts][2] float32 correspondences; prev_dev may be NULL with
 *                      grid_w/grid_h/grid_step > 0, meaning the regular sampling grid and
 *                      curr = prev + flow where curr_dev then holds the sampled FLOW.
 * mode_mask            bit (1 << VSTAB_MODE_*) set = compute that candidate
 * out_dev              [n_pairs][3] results indexed by VSTAB_MODE_*.  confidence is left to the
 *                      caller: n_inliers / n_valid (similarity, perspective) or n_valid / n_total
 */
VSTAB_API int vstab_fit_batch(vstab_handle* h, const float* prev_dev, const float* curr_dev, int n_pairs,
                    int n_pts, int grid_w, int grid_h, int grid_step, int mode_mask,
                    vstab_fit_result* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VSTAB_H_ */
