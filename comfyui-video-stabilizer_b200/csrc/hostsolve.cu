// hostsolve.cu -- the O(frames) trajectory maths between the estimation kernels and the resampler, in C on the host.
//
// While these few hundred scalar operations run, the GPU idles: the resampler needs the solved matrices.  In numpy
// they cost ~0.45 ms per 121-frame clip (some 80 small-array operations at 3-10 us of interpreter / dispatch
// overhead each); here the same IEEE operations in the same order take a few microseconds.  No GPU involved, no CUDA
// call: this file is host code that happens to live in libvstab.so.
//
// Reference (file:line in the reference repository), restated operation by operation, float widths included:
//   ladder acceptance          nodes/video_stabilizer_flow.py:156-210, nodes/video_stabilizer_classic.py:104-158
//   rescale to full size       nodes/stabilizer_utils.py:279-297   (float64 products, float32 result)
//   matrix -> params           nodes/stabilizer_utils.py:300-326   (float32 a*a + c*c, libm atan2 / log in float64)
//   path cumsum                nodes/video_stabilizer_flow.py:341-349
//   params -> matrix           nodes/stabilizer_utils.py:329-358   (libm exp / cos / sin, float32 result)
//   bounding boxes             nodes/stabilizer_utils.py:1010-1032 (float32 matrix promoted to float64, no FMA)
//   crop_and_pad recentring    nodes/video_stabilizer_flow.py:471-489
// Box smoothing (np.convolve) stays in numpy between the two calls: its summation order belongs to the BLAS build.
// Compiled with -ffp-contract=off: every product and sum below rounds once, like numpy's element-wise loops.
#include "../../include/vstab.h"

#include <math.h>
#include <stdint.h>
#include <string.h>

namespace {

inline int params_per_mode(int mode) { return mode == VSTAB_MODE_TRANSLATION ? 2 : (mode == VSTAB_MODE_SIMILARITY ? 4 : 8); }

struct FitWords {  // one vstab_fit_result as 12 eight-byte words
  double m[9];
  double residual;
  int32_t n_inliers, n_valid, n_total, ok;
};

// np.minimum / np.maximum propagate NaNs
inline double np_min(double a, double b) { return (a <= b || a != a) ? a : b; }
inline double np_max(double a, double b) { return (a >= b || a != a) ? a : b; }

}  // namespace

extern "C" int vstab_host_trajectory(const double* raw, const int32_t* detected, int n_pairs, int min_points, int mode,
                                     int src_w, int src_h, int work_w, int work_h, float* matrices, double* path,
                                     int* first_fallback) {
  if (!raw || !matrices || !path || !first_fallback || n_pairs < 0 || mode < 0 || mode > 2 || src_w <= 0 || src_h <= 0)
    return VSTAB_ERR_INVALID;
  static_assert(sizeof(FitWords) == 96, "vstab_fit_result is 12 words");
  const FitWords* fit = reinterpret_cast<const FitWords*>(raw);
  const int need = mode == VSTAB_MODE_PERSPECTIVE ? 4 : 3;
  const double thr = mode == VSTAB_MODE_PERSPECTIVE ? 0.15 : 0.1;
  // every pair must accept the requested model itself; the first one that does not starts the sticky ladder,
  // which the caller replays the long way
  *first_fallback = n_pairs;
  for (int p = 0; p < n_pairs; ++p) {
    const FitWords* f = fit + 3 * p;
    int n_valid = f[0].n_valid;
    if (f[1].n_valid > n_valid) n_valid = f[1].n_valid;
    if (f[2].n_valid > n_valid) n_valid = f[2].n_valid;
    bool ok = n_valid >= min_points;
    if (detected && detected[p] < 12) ok = false;
    if (ok && mode != VSTAB_MODE_TRANSLATION) {
      const double conf = (double)f[mode].n_inliers / (double)(n_valid > 1 ? n_valid : 1);
      ok = f[mode].ok != 0 && n_valid >= need && conf >= thr;
    }
    if (!ok) {
      *first_fallback = p;
      return VSTAB_OK;
    }
  }
  const bool rescale = work_w > 0 && work_h > 0;
  double up[3] = {1.0, 1.0, 1.0}, down[3] = {1.0, 1.0, 1.0};
  if (rescale) {
    const double kx = work_w / (double)src_w, ky = work_h / (double)src_h;
    down[0] = kx;
    down[1] = ky;
    up[0] = 1.0 / kx;
    up[1] = 1.0 / ky;
  }
  const int K = params_per_mode(mode);
  for (int k = 0; k < K; ++k) path[k] = 0.0;
  for (int p = 0; p < n_pairs; ++p) {
    float* m = matrices + 9 * p;
    const double* src = fit[3 * p + mode].m;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const float m32 = (float)src[3 * i + j];  // the per-pair transform is float32 before anything else happens to it
        m[3 * i + j] = rescale ? (float)(((double)m32 * up[i]) * down[j]) : m32;
      }
    double d[8];
    if (mode == VSTAB_MODE_TRANSLATION) {
      d[0] = m[2];
      d[1] = m[5];
    } else if (mode == VSTAB_MODE_SIMILARITY) {
      const float a = m[0], c = m[3];
      const float aa = a * a, cc = c * c;
      const float mag2 = aa + cc;  // float32 arithmetic, like numpy scalars of the float32 matrix
      d[0] = m[2];
      d[1] = m[5];
      d[2] = atan2((double)c, (double)a);
      const double clamped = mag2 > 1e-10f ? (double)mag2 : 1e-10;
      d[3] = log(sqrt(clamped));
    } else {
      d[0] = (double)(m[0] - 1.0f);
      d[1] = m[1];
      d[2] = m[2];
      d[3] = m[3];
      d[4] = (double)(m[4] - 1.0f);
      d[5] = m[5];
      d[6] = m[6];
      d[7] = m[7];
    }
    const double* prev = path + (size_t)K * p;
    double* next = path + (size_t)K * (p + 1);
    for (int k = 0; k < K; ++k) next[k] = prev[k] + d[k];
  }
  return VSTAB_OK;
}

extern "C" int vstab_host_framing(const double* diffs, int n_frames, int mode, int width, int height, float* apply,
                                  double* mins, double* maxs, double* box) {
  if (!diffs || !apply || !mins || !maxs || !box || n_frames <= 0 || mode < 0 || mode > 2) return VSTAB_ERR_INVALID;
  const int K = params_per_mode(mode);
  const double cx[4] = {0.0, (double)width, 0.0, (double)width};
  const double cy[4] = {0.0, 0.0, (double)height, (double)height};
  double in_x0 = 0, in_y0 = 0, in_x1 = 0, in_y1 = 0, out_x0 = 0, out_y0 = 0, out_x1 = 0, out_y1 = 0;
  int affine = 1;
  for (int n = 0; n < n_frames; ++n) {
    const double* p = diffs + (size_t)K * n;
    double m64[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};
    if (mode == VSTAB_MODE_TRANSLATION) {
      m64[2] = p[0];
      m64[5] = p[1];
    } else if (mode == VSTAB_MODE_SIMILARITY) {
      const double k = exp(p[3]), cs = cos(p[2]), sn = sin(p[2]);
      m64[0] = k * cs;
      m64[1] = -k * sn;
      m64[2] = p[0];
      m64[3] = k * sn;
      m64[4] = k * cs;
      m64[5] = p[1];
    } else {
      m64[0] = p[0] + 1.0;
      m64[1] = p[1];
      m64[2] = p[2];
      m64[3] = p[3];
      m64[4] = p[4] + 1.0;
      m64[5] = p[5];
      m64[6] = p[6];
      m64[7] = p[7];
    }
    float* m = apply + 9 * n;
    for (int i = 0; i < 9; ++i) {
      m[i] = (float)m64[i];
      if (!isfinite(m[i])) affine = 0;
    }
    if (m[6] != 0.0f || m[7] != 0.0f || m[8] != 1.0f) affine = 0;
    // corners of the frame through the float32 matrix, in float64 term by term: (m0*cx + m1*cy) + m2*1
    double x[4], y[4];
    for (int c = 0; c < 4; ++c) {
      double q[3];
      for (int r = 0; r < 3; ++r) {
        const double t0 = (double)m[3 * r] * cx[c], t1 = (double)m[3 * r + 1] * cy[c], t2 = (double)m[3 * r + 2] * 1.0;
        const double s = t0 + t1;
        q[r] = s + t2;
      }
      x[c] = q[0] / q[2];
      y[c] = q[1] / q[2];
    }
    const double lo_x = np_min(np_min(x[0], x[1]), np_min(x[2], x[3])), hi_x = np_max(np_max(x[0], x[1]), np_max(x[2], x[3]));
    const double lo_y = np_min(np_min(y[0], y[1]), np_min(y[2], y[3])), hi_y = np_max(np_max(y[0], y[1]), np_max(y[2], y[3]));
    mins[2 * n] = lo_x, mins[2 * n + 1] = lo_y, maxs[2 * n] = hi_x, maxs[2 * n + 1] = hi_y;
    if (n == 0) {
      in_x0 = out_x0 = lo_x, in_y0 = out_y0 = lo_y, in_x1 = out_x1 = hi_x, in_y1 = out_y1 = hi_y;
    } else {
      in_x0 = np_max(in_x0, lo_x), in_y0 = np_max(in_y0, lo_y), in_x1 = np_min(in_x1, hi_x), in_y1 = np_min(in_y1, hi_y);
      out_x0 = np_min(out_x0, lo_x), out_y0 = np_min(out_y0, lo_y), out_x1 = np_max(out_x1, hi_x), out_y1 = np_max(out_y1, hi_y);
    }
  }
  box[0] = in_x0, box[1] = in_y0, box[2] = in_x1, box[3] = in_y1;      // common inner rectangle
  box[4] = out_x0, box[5] = out_y0, box[6] = out_x1, box[7] = out_y1;  // union (expand framing)
  box[8] = (double)affine;  // 1: every matrix is finite with a (0, 0, 1) last row -> vstab_host_shift applies
  return VSTAB_OK;
}

// [[1, 0, ox], [0, 1, oy], [0, 0, 1]] @ m for affine float32 matrices: with a (0, 0, 1) last row every element of the
// product has at most one inexact operation (m02 + ox, m12 + oy), so the float32 BLAS product the reference runs per
// frame has these bits whatever its accumulation order or fusing.  Perspective matrices go through numpy's matmul.
extern "C" int vstab_host_shift(const float* apply, int n_frames, float off_x, float off_y, float* out) {
  if (!apply || !out || n_frames < 0) return VSTAB_ERR_INVALID;
  for (int n = 0; n < n_frames; ++n) {
    const float* m = apply + 9 * n;
    float* o = out + 9 * n;
    if (m[6] != 0.0f || m[7] != 0.0f || m[8] != 1.0f) return VSTAB_ERR_UNSUPPORTED;
    // the product accumulates from +0, so a -0 entry comes out as +0: keep that too (x + 0.0f is not a no-op for -0)
    for (int i = 0; i < 9; ++i) o[i] = m[i] + 0.0f;
    o[2] = o[2] + off_x;
    o[5] = o[5] + off_y;
  }
  return VSTAB_OK;
}

// path -> target path and per-frame deltas (nodes/video_stabilizer_flow.py:351-374): camera lock = zero target, otherwise
// path + strength * (box_filter(path) - path) with the edge-padded odd box filter of stabilizer_utils.py:361-383.
// numpy evaluates that filter with np.convolve: for kernels of up to 11 taps its own scalar loop, which adds the products
// in tap order starting from +0 (restated here, and checked against np.convolve once per process and window by
// hostmath.native_target); longer kernels go through the BLAS dot product whose summation order belongs to the CPU it
// runs on -- VSTAB_ERR_UNSUPPORTED, and the caller uses numpy.
extern "C" int vstab_host_target(const double* path, int n_frames, int n_params, int window, double strength, int camera_lock,
                                 double* target, double* diffs) {
  if (!path || !target || !diffs || n_frames <= 0 || n_params <= 0 || n_params > 8 || window < 0) return VSTAB_ERR_INVALID;
  const int K = n_params;
  if (camera_lock) {
    for (int i = 0; i < n_frames * K; ++i) {
      target[i] = 0.0;
      diffs[i] = target[i] - path[i];
    }
    return VSTAB_OK;
  }
  const bool filtered = window > 0 && n_frames > 2;  // smooth <= 0 or a clip of one or two frames: the filter is the identity
  if (filtered && (window > 11 || (window & 1) == 0)) return VSTAB_ERR_UNSUPPORTED;
  const int half = window / 2;
  const double tap = 1.0 / (double)window;
  for (int i = 0; i < n_frames; ++i) {
    // the parameters of a frame are independent sums: their (up to 8) chains of dependent additions run side by side,
    // each one in tap order like numpy's loop
    double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (filtered) {
      for (int t = 0; t < window; ++t) {
        int src = i + t - half;
        src = src < 0 ? 0 : (src >= n_frames ? n_frames - 1 : src);
        const double* row = path + (size_t)src * K;
        for (int k = 0; k < K; ++k) acc[k] += row[k] * tap;
      }
    }
    for (int k = 0; k < K; ++k) {
      const double p = path[(size_t)i * K + k];
      const double sm = filtered ? acc[k] : p;
      const double d = sm - p;
      const double scaled = strength * d;
      target[(size_t)i * K + k] = p + scaled;
      diffs[(size_t)i * K + k] = target[(size_t)i * K + k] - p;
    }
  }
  return VSTAB_OK;
}
