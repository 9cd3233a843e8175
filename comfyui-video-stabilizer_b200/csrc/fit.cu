// fit.cu -- K4 + K7 + K8 + K9: robust per-pair model fit, batched over all frame pairs.
//
// Replaces nodes/video_stabilizer_flow.py:148-210 and nodes/video_stabilizer_classic.py:112-158:
// the finite filter plus cv2.findHomography (RANSAC 2.5 px / 2000 / 0.992),
// cv2.estimateAffinePartial2D (RANSAC 2.0 px / 2000 / 0.992) and the per-axis median.  All three
// candidates are produced for every pair so the caller can replay the reference's sticky mode
// ladder over the table (frame-range shards then agree without talking to each other).
//
// The RANSAC drivers replay OpenCV's sampler exactly -- RNG(2^64-1) multiply-with-carry stream,
// distinct-index redraw, subset checks, strict-> acceptance, RANSACUpdateNumIters -- so the
// winning hypothesis, its consensus set and therefore `confidence` are the reference's, not merely
// "a" robust fit.  One CTA per pair: hypotheses are generated 8 at a time by one thread (the RNG
// stream is sequential), scored by all 256 threads over the correspondences with warp-shuffle
// reductions, and the sequential accept / early-terminate rule is replayed over the 8 counts.
// Final models: closed-form least squares (similarity; what cv2's LM converges to on a linear
// problem) and normalised DLT + Levenberg-Marquardt in cv::LMSolver's schedule (perspective), all
// in double with warp-reduced normal equations.
#include "common.cuh"

#include <float.h>
#include <math.h>

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kBatch = 8;
constexpr int kMaxIters = 2000;
constexpr double kConfidence = 0.992;

struct FitOut {  // mirrors vstab_fit_result
  double matrix[9];
  double residual;
  int n_inliers, n_valid, n_total, ok;
};

// ---- OpenCV RNG (multiply with carry) ----
struct CvRng {
  unsigned long long state;
  __device__ unsigned next() {
    state = (unsigned long long)(unsigned)state * 4164903690ULL + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  __device__ int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

__device__ int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = fmin(fmax(p, 0.), 1.);
  ep = fmin(fmax(ep, 0.), 1.);
  double num = fmax(1. - p, DBL_MIN);
  double denom = 1. - pow(1. - ep, (double)model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

// ---- block reductions ----
template <int K>
__device__ void block_sum(double (&v)[K], double* smem /* [kWarps][K] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; k++) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) smem[warp * K + k] = x;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0;
    for (int w = 0; w < kWarps; w++) s += smem[w * K + threadIdx.x];
    smem[threadIdx.x] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; k++) v[k] = smem[k];
  __syncthreads();
}

template <int K>
__device__ void block_sum_int(int (&v)[K], int* smem /* [kWarps][K] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; k++) {
    int x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) smem[warp * K + k] = x;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    int s = 0;
    for (int w = 0; w < kWarps; w++) s += smem[w * K + threadIdx.x];
    smem[threadIdx.x] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; k++) v[k] = smem[k];
  __syncthreads();
}

// ---- stage 0: correspondences + finite filter (order-preserving compaction) ----
__global__ void __launch_bounds__(kThreads) prepare_kernel(const float* __restrict__ prev_in,
                                                           const float* __restrict__ curr_in, int n_pts, int grid_w,
                                                           int grid_step, float2* __restrict__ P, float2* __restrict__ C,
                                                           int* __restrict__ n_valid) {
  __shared__ int s_warp[kWarps];
  __shared__ int s_base;
  const int pair = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int start = 0; start < n_pts; start += kThreads) {
    const int i = start + threadIdx.x;
    float2 p = make_float2(0, 0), c = make_float2(0, 0);
    bool ok = false;
    if (i < n_pts) {
      const size_t k = ((size_t)pair * n_pts + i) * 2;
      if (prev_in) {
        p = make_float2(prev_in[k], prev_in[k + 1]);
        c = make_float2(curr_in[k], curr_in[k + 1]);
        ok = isfinite(p.x) && isfinite(p.y);
      } else {  // regular sampling grid; curr_in holds the sampled flow
        p = make_float2((float)((i % grid_w) * grid_step), (float)((i / grid_w) * grid_step));
        c = make_float2(p.x + curr_in[k], p.y + curr_in[k + 1]);
        ok = true;
      }
      ok = ok && isfinite(c.x) && isfinite(c.y);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    int offset = s_base;
    for (int w = 0; w < warp; w++) offset += s_warp[w];
    if (ok) {
      const int dst = offset + __popc(ballot & ((1u << lane) - 1));
      P[(size_t)pair * n_pts + dst] = p;
      C[(size_t)pair * n_pts + dst] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < kWarps; w++) t += s_warp[w];
      s_base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) n_valid[pair] = s_base;
}

// ---- translation: per-axis median of float32 shifts ----
// Exact selection instead of a sort: the shifts become order-preserving 32-bit keys and four 8-bit
// histogram passes (both axes at once, warp 0 / warp 1 resolve the bins) pin down the element of rank
// n/2; for even n the element of rank n/2 - 1 is that same value when it has duplicates below the
// rank, else the largest smaller key.  Same result as np.median on the float32 shifts.
__device__ __forceinline__ unsigned float_key(float v) {
  const unsigned u = __float_as_uint(v + 0.0f);  // -0 -> +0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(1024) translation_kernel(const float2* __restrict__ P, const float2* __restrict__ C,
                                                           const int* __restrict__ n_valid, int n_pts, int cap,
                                                           FitOut* __restrict__ out) {
  extern __shared__ unsigned s_key[];  // [2][cap]
  __shared__ unsigned s_hist[2][256];
  __shared__ unsigned s_prefix[2], s_rank[2], s_less[2], s_maxless[2];
  __shared__ float s_t[2];
  const int pair = blockIdx.x;
  const int n = n_valid[pair];
  const float2* p = P + (size_t)pair * n_pts;
  const float2* c = C + (size_t)pair * n_pts;
  FitOut* o = out + (size_t)pair * 3 + VSTAB_MODE_TRANSLATION;
  if (n <= 0) {
    if (threadIdx.x == 0) {
      for (int k = 0; k < 9; k++) o->matrix[k] = (k % 4 == 0) ? 1.0 : 0.0;
      o->residual = 0;
      o->n_inliers = 0;
      o->n_valid = 0;
      o->n_total = n_pts;
      o->ok = 0;
    }
    return;
  }
  unsigned* kx = s_key;
  unsigned* ky = s_key + cap;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    kx[i] = float_key(c[i].x - p[i].x);
    ky[i] = float_key(c[i].y - p[i].y);
  }
  if (threadIdx.x < 2) {
    s_prefix[threadIdx.x] = 0;
    s_rank[threadIdx.x] = (unsigned)(n / 2);
    s_less[threadIdx.x] = 0;
    s_maxless[threadIdx.x] = 0;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int pass = 3; pass >= 0; pass--) {
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const unsigned mask = pass == 3 ? 0u : (0xffffffffu << (8 * (pass + 1)));
    const unsigned px = s_prefix[0], py = s_prefix[1];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned a = kx[i], b2 = ky[i];
      if ((a & mask) == px) atomicAdd(&s_hist[0][(a >> (8 * pass)) & 255u], 1u);
      if ((b2 & mask) == py) atomicAdd(&s_hist[1][(b2 >> (8 * pass)) & 255u], 1u);
    }
    __syncthreads();
    if (warp < 2) {  // warp `axis`: which bin holds the element of the wanted rank
      const unsigned* hst = s_hist[warp];
      unsigned loc = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) loc += hst[lane * 8 + q];
      unsigned inc = loc;
      for (int ofs = 1; ofs < 32; ofs <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, ofs);
        if (lane >= ofs) inc += t;
      }
      const unsigned rank = s_rank[warp];
      const unsigned before = inc - loc;
      if (rank >= before && rank < inc) {  // exactly one lane
        unsigned cum = before;
        int bin = lane * 8;
        for (int q = 0; q < 8; q++) {
          const unsigned hq = hst[lane * 8 + q];
          if (rank < cum + hq) { bin = lane * 8 + q; break; }
          cum += hq;
        }
        s_rank[warp] = rank - cum;
        s_prefix[warp] |= (unsigned)bin << (8 * pass);
      }
    }
    __syncthreads();
  }
  {
    const unsigned hx = s_prefix[0], hy = s_prefix[1];
    unsigned lx = 0, ly = 0, mx = 0, my = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned a = kx[i], b2 = ky[i];
      if (a < hx) { lx++; mx = max(mx, a); }
      if (b2 < hy) { ly++; my = max(my, b2); }
    }
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
      lx += __shfl_down_sync(0xffffffffu, lx, ofs);
      ly += __shfl_down_sync(0xffffffffu, ly, ofs);
      mx = max(mx, __shfl_down_sync(0xffffffffu, mx, ofs));
      my = max(my, __shfl_down_sync(0xffffffffu, my, ofs));
    }
    if (lane == 0) {
      atomicAdd(&s_less[0], lx);
      atomicAdd(&s_less[1], ly);
      atomicMax(&s_maxless[0], mx);
      atomicMax(&s_maxless[1], my);
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      const int axis = threadIdx.x;
      const float hi = key_float(s_prefix[axis]);
      float med = hi;
      if ((n & 1) == 0) {
        // numpy: the float32 mean of the two middle elements
        const float lo = (s_less[axis] < (unsigned)(n / 2)) ? hi : key_float(s_maxless[axis]);
        med = (lo + hi) * 0.5f;
      }
      s_t[axis] = med;
    }
    __syncthreads();
  }
  const float tx = s_t[0], ty = s_t[1];
  double acc = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float ex = fabsf((p[i].x + tx) - c[i].x), ey = fabsf((p[i].y + ty) - c[i].y);
    acc += (double)ex + (double)ey;
  }
  for (int ofs = 16; ofs > 0; ofs >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, ofs);
  __shared__ double s_w[32];
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += s_w[w];
    for (int k = 0; k < 9; k++) o->matrix[k] = (k % 4 == 0) ? 1.0 : 0.0;
    o->matrix[2] = tx;
    o->matrix[5] = ty;
    o->residual = tot / (2.0 * n);
    o->n_inliers = n;
    o->n_valid = n;
    o->n_total = n_pts;
    o->ok = 1;
  }
}

// ---- similarity -------------------------------------------------------------------------------

__device__ void similarity_from_2(const float2 f0, const float2 f1, const float2 t0, const float2 t1, double* M) {
  const double x1 = f0.x, y1 = f0.y, x2 = f1.x, y2 = f1.y;
  const double X1 = t0.x, Y1 = t0.y, X2 = t1.x, Y2 = t1.y;
  const double d = 1. / ((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
  const double S0 = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2));
  const double S1 = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2));
  const double S2 = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2) - (X1 * x2 - X2 * x1) * (x1 - x2));
  const double S3 = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2) - (Y1 * y2 - Y2 * y1) * (y1 - y2));
  M[0] = S0; M[1] = -S1; M[2] = S2; M[3] = S1; M[4] = S0; M[5] = S3;
}

__device__ __forceinline__ bool affine_inlier(const double* M, const float2 p, const float2 c, float t) {
  const double a = M[0] * p.x + M[1] * p.y + M[2] - c.x;
  const double b = M[3] * p.x + M[4] * p.y + M[5] - c.y;
  return (float)(a * a + b * b) <= t;
}

// The same decision from float arithmetic with a rigorous error band: cv2 evaluates the residual in double, but a point
// is only in doubt when its float residual lands within the rounding error of the threshold.  Everything else is
// decided at FP32 rate; the doubtful points (a handful per hypothesis) pay for the double evaluation.  When the motion
// is not a similarity (perspective clips) the RANSAC runs all 2000 hypotheses over 8160 points and was bound by the
// FP64 pipe: 1.44 ms per 48 pairs.
__device__ __forceinline__ bool affine_inlier_fast(const double* M, const float* Mf, const float2 p, const float2 c, float t) {
  const float t0 = Mf[0] * p.x, t1 = Mf[1] * p.y, t3 = Mf[3] * p.x, t4 = Mf[4] * p.y;
  const float a = t0 + t1 + Mf[2] - c.x;
  const float b = t3 + t4 + Mf[5] - c.y;
  const float e = a * a + b * b;
  // |a - a_exact| <= da: four roundings of the sum + the float rounding of the three coefficients, each relative 2^-24
  const float da = 4e-7f * (fabsf(t0) + fabsf(t1) + fabsf(Mf[2]) + fabsf(c.x));
  const float db = 4e-7f * (fabsf(t3) + fabsf(t4) + fabsf(Mf[5]) + fabsf(c.y));
  const float band = 2.f * (fabsf(a) * da + fabsf(b) * db) + da * da + db * db + 1e-6f * e;
  if (e < t - band) return true;
  if (e > t + band) return false;
  return affine_inlier(M, p, c, t);
}

__global__ void __launch_bounds__(kThreads) similarity_kernel(const float2* __restrict__ P, const float2* __restrict__ C,
                                                              const int* __restrict__ n_valid, int n_pts,
                                                              FitOut* __restrict__ out) {
  __shared__ double s_model[kBatch][6];
  __shared__ float s_modelf[kBatch][6];
  __shared__ double s_best[6];
  __shared__ double s_red[kWarps * 8];
  __shared__ int s_ired[kWarps * kBatch];
  __shared__ int s_ctl[4];  // done, best_count, it, niters
  __shared__ CvRng s_rng;
  const int pair = blockIdx.x;
  const int n = n_valid[pair];
  const float2* p = P + (size_t)pair * n_pts;
  const float2* c = C + (size_t)pair * n_pts;
  FitOut* o = out + (size_t)pair * 3 + VSTAB_MODE_SIMILARITY;
  const float thr = 2.0f * 2.0f;
  if (threadIdx.x == 0) {
    s_rng.state = 0xffffffffffffffffULL;
    s_ctl[0] = 0; s_ctl[1] = 0; s_ctl[2] = 0; s_ctl[3] = kMaxIters;
    o->n_valid = n; o->n_total = n_pts; o->ok = 0; o->n_inliers = 0; o->residual = 0;
    for (int k = 0; k < 9; k++) o->matrix[k] = (k % 4 == 0) ? 1.0 : 0.0;
  }
  __syncthreads();
  if (n < 2) return;
  if (n == 2) {
    if (threadIdx.x == 0) { similarity_from_2(p[0], p[1], c[0], c[1], s_best); s_ctl[1] = 2; }
    __syncthreads();
  } else {
    while (true) {
      if (threadIdx.x == 0) {
        for (int b = 0; b < kBatch; b++) {
          int i0 = s_rng.uniform(0, n);
          int i1 = s_rng.uniform(0, n);
          while (i1 == i0) i1 = s_rng.uniform(0, n);
          similarity_from_2(p[i0], p[i1], c[i0], c[i1], s_model[b]);
          for (int k = 0; k < 6; k++) s_modelf[b][k] = (float)s_model[b][k];
        }
      }
      __syncthreads();
      int cnt[kBatch];
#pragma unroll
      for (int b = 0; b < kBatch; b++) cnt[b] = 0;
      for (int i = threadIdx.x; i < n; i += kThreads) {
        const float2 pp = p[i], cc = c[i];
#pragma unroll
        for (int b = 0; b < kBatch; b++) cnt[b] += affine_inlier_fast(s_model[b], s_modelf[b], pp, cc, thr) ? 1 : 0;
      }
      block_sum_int<kBatch>(cnt, s_ired);
      if (threadIdx.x == 0) {
        int best = s_ctl[1], it = s_ctl[2], niters = s_ctl[3];
        for (int b = 0; b < kBatch && it < niters; b++, it++) {
          const int good = cnt[b];
          if (good > max(best, 1)) {
            best = good;
            for (int k = 0; k < 6; k++) s_best[k] = s_model[b][k];
            niters = ransac_update_num_iters(kConfidence, (double)(n - good) / n, 2, niters);
          }
        }
        s_ctl[1] = best; s_ctl[2] = it; s_ctl[3] = niters;
        s_ctl[0] = it >= niters;
      }
      __syncthreads();
      if (s_ctl[0]) break;
    }
  }
  const int best = s_ctl[1];
  if (best <= 0) return;
  // least squares over the consensus set of the winning hypothesis (two passes: means, centred sums)
  double m[8];
  for (int k = 0; k < 8; k++) m[k] = 0;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    const float2 pp = p[i], cc = c[i];
    if (affine_inlier(s_best, pp, cc, thr)) { m[0] += pp.x; m[1] += pp.y; m[2] += cc.x; m[3] += cc.y; m[4] += 1.0; }
  }
  block_sum<8>(m, s_red);
  const double cntd = m[4];
  const double xc = m[0] / cntd, yc = m[1] / cntd, Xc = m[2] / cntd, Yc = m[3] / cntd;
  double q[8];
  for (int k = 0; k < 8; k++) q[k] = 0;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    const float2 pp = p[i], cc = c[i];
    if (affine_inlier(s_best, pp, cc, thr)) {
      const double xd = pp.x - xc, yd = pp.y - yc, Xd = cc.x - Xc, Yd = cc.y - Yc;
      q[0] += xd * xd + yd * yd;
      q[1] += xd * Xd + yd * Yd;
      q[2] += xd * Yd - yd * Xd;
    }
  }
  block_sum<8>(q, s_red);
  const double a = q[1] / q[0], b = q[2] / q[0];
  const double tx = Xc - a * xc + b * yc, ty = Yc - b * xc - a * yc;
  double r[8];
  for (int k = 0; k < 8; k++) r[k] = 0;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    const float2 pp = p[i], cc = c[i];
    r[0] += fabs((a * pp.x - b * pp.y + tx) - cc.x) + fabs((b * pp.x + a * pp.y + ty) - cc.y);
  }
  block_sum<8>(r, s_red);
  if (threadIdx.x == 0) {
    o->matrix[0] = a; o->matrix[1] = -b; o->matrix[2] = tx;
    o->matrix[3] = b; o->matrix[4] = a; o->matrix[5] = ty;
    o->matrix[6] = 0; o->matrix[7] = 0; o->matrix[8] = 1;
    o->residual = r[0] / (2.0 * n);
    o->n_inliers = (int)(cntd + 0.5);
    o->ok = 1;
  }
}

// ---- perspective -------------------------------------------------------------------------------

__device__ bool have_collinear(const float2* pts, int count) {
  const int i = count - 1;
  for (int j = 0; j < i; j++) {
    const double dx1 = (double)pts[j].x - pts[i].x, dy1 = (double)pts[j].y - pts[i].y;
    for (int k = 0; k < j; k++) {
      const double dx2 = (double)pts[k].x - pts[i].x, dy2 = (double)pts[k].y - pts[i].y;
      if (fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return true;
    }
  }
  return false;
}

__device__ double det3(const float2 a, const float2 b, const float2 c) {
  // | a.x a.y 1 ; b.x b.y 1 ; c.x c.y 1 |
  const double m00 = a.x, m01 = a.y, m10 = b.x, m11 = b.y, m20 = c.x, m21 = c.y;
  return m00 * (m11 - m21) - m01 * (m10 - m20) + (m10 * m21 - m11 * m20);
}

__device__ bool homography_subset_ok(const float2* src, const float2* dst) {
  if (have_collinear(src, 4) || have_collinear(dst, 4)) return false;
  const int tt[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
  int negative = 0;
  for (int i = 0; i < 4; i++) {
    const double da = det3(src[tt[i][0]], src[tt[i][1]], src[tt[i][2]]);
    const double db = det3(dst[tt[i][0]], dst[tt[i][1]], dst[tt[i][2]]);
    negative += (da * db < 0) ? 1 : 0;
  }
  return negative == 0 || negative == 4;
}

// Cyclic Jacobi on a symmetric 9x9 (double); returns the eigenvector of the smallest eigenvalue.
__device__ void smallest_eigvec9(double A[9][9], double* vec) {
  double V[9][9];
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < 9; i++) {
      diag += A[i][i] * A[i][i];
      for (int j = i + 1; j < 9; j++) off += A[i][j] * A[i][j];
    }
    if (off <= 1e-40 * diag || off == 0) break;
    for (int pi = 0; pi < 8; pi++)
      for (int qi = pi + 1; qi < 9; qi++) {
        const double apq = A[pi][qi];
        if (apq == 0) continue;
        const double theta = (A[qi][qi] - A[pi][pi]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < 9; k++) {
          const double akp = A[k][pi], akq = A[k][qi];
          A[k][pi] = cs * akp - sn * akq;
          A[k][qi] = sn * akp + cs * akq;
        }
        for (int k = 0; k < 9; k++) {
          const double apk = A[pi][k], aqk = A[qi][k];
          A[pi][k] = cs * apk - sn * aqk;
          A[qi][k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < 9; k++) {
          const double vkp = V[k][pi], vkq = V[k][qi];
          V[k][pi] = cs * vkp - sn * vkq;
          V[k][qi] = sn * vkp + cs * vkq;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 9; i++)
    if (A[i][i] < A[best][best]) best = i;
  for (int k = 0; k < 9; k++) vec[k] = V[k][best];
}

// H = invHnorm * H0 * Hnorm2, scaled so that H[8] = 1
__device__ void denormalise_h(const double* h0, double cMx, double cMy, double sMx, double sMy, double cmx, double cmy,
                              double smx, double smy, double* H) {
  const double inv[9] = {1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1};
  const double n2[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += inv[i * 3 + k] * h0[k * 3 + j];
      t[i * 3 + j] = s;
    }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += t[i * 3 + k] * n2[k * 3 + j];
      H[i * 3 + j] = s;
    }
  const double sc = 1. / H[8];
  for (int k = 0; k < 9; k++) H[k] *= sc;
}

// HomographyEstimatorCallback::runKernel for a minimal 4-point sample (one thread).
__device__ bool homography_from_4(const float2* M, const float2* m, double* H) {
  double cMx = 0, cMy = 0, cmx = 0, cmy = 0;
  for (int i = 0; i < 4; i++) { cmx += m[i].x; cmy += m[i].y; cMx += M[i].x; cMy += M[i].y; }
  cmx /= 4; cmy /= 4; cMx /= 4; cMy /= 4;
  double smx = 0, smy = 0, sMx = 0, sMy = 0;
  for (int i = 0; i < 4; i++) {
    smx += fabs(m[i].x - cmx); smy += fabs(m[i].y - cmy); sMx += fabs(M[i].x - cMx); sMy += fabs(M[i].y - cMy);
  }
  if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON) return false;
  smx = 4 / smx; smy = 4 / smy; sMx = 4 / sMx; sMy = 4 / sMy;
  double L[9][9];
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++) L[i][j] = 0;
  for (int i = 0; i < 4; i++) {
    const double x = (m[i].x - cmx) * smx, y = (m[i].y - cmy) * smy;
    const double X = (M[i].x - cMx) * sMx, Y = (M[i].y - cMy) * sMy;
    const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
    const double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
    for (int j = 0; j < 9; j++)
      for (int k = j; k < 9; k++) L[j][k] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
  }
  for (int j = 0; j < 9; j++)
    for (int k = 0; k < j; k++) L[j][k] = L[k][j];
  double h0[9];
  smallest_eigvec9(L, h0);
  denormalise_h(h0, cMx, cMy, sMx, sMy, cmx, cmy, smx, smy, H);
  for (int k = 0; k < 9; k++)
    if (!isfinite(H[k])) return false;
  return true;
}

__device__ __forceinline__ bool homography_inlier(const float* Hf, const float2 M, const float2 m, float t) {
  const float ww = 1.f / (Hf[6] * M.x + Hf[7] * M.y + 1.f);
  const float dx = (Hf[0] * M.x + Hf[1] * M.y + Hf[2]) * ww - m.x;
  const float dy = (Hf[3] * M.x + Hf[4] * M.y + Hf[5]) * ww - m.y;
  return dx * dx + dy * dy <= t;
}

// Solves the n x n system A d = v in place (Gaussian elimination, partial pivoting); n <= 8.
__device__ bool solve_small(double A[8][8], double* v, double* d, int n) {
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++)
      if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
    if (A[piv][c] == 0) return false;
    if (piv != c) {
      for (int k = 0; k < n; k++) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
      const double t = v[c]; v[c] = v[piv]; v[piv] = t;
    }
    for (int r = c + 1; r < n; r++) {
      const double f = A[r][c] / A[c][c];
      for (int k = c; k < n; k++) A[r][k] -= f * A[c][k];
      v[r] -= f * v[c];
    }
  }
  for (int r = n - 1; r >= 0; r--) {
    double s = v[r];
    for (int k = r + 1; k < n; k++) s -= A[r][k] * d[k];
    d[r] = s / A[r][r];
  }
  return true;
}

// Residual / Jacobian sums of HomographyRefineCallback at parameters h (8 doubles) over the
// flagged points: out[0..35] = upper triangle of JtJ, out[36..43] = Jt r, out[44] = |r|^2,
// out[45] = max |r| (summed as a max through a separate path).
constexpr int kLmSums = 45;

__device__ void lm_accumulate(const double* h, const float2 M, const float2 m, double* acc, double& rmax) {
  const double Mx = M.x, My = M.y;
  double ww = h[6] * Mx + h[7] * My + 1.;
  ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
  const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww;
  const double yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
  const double ex = xi - m.x, ey = yi - m.y;
  const double Jx[8] = {Mx * ww, My * ww, ww, 0, 0, 0, -Mx * ww * xi, -My * ww * xi};
  const double Jy[8] = {0, 0, 0, Mx * ww, My * ww, ww, -Mx * ww * yi, -My * ww * yi};
  int k = 0;
  for (int i = 0; i < 8; i++)
    for (int j = i; j < 8; j++) acc[k++] += Jx[i] * Jx[j] + Jy[i] * Jy[j];
  for (int i = 0; i < 8; i++) acc[36 + i] += Jx[i] * ex + Jy[i] * ey;
  acc[44] += ex * ex + ey * ey;
  rmax = fmax(rmax, fmax(fabs(ex), fabs(ey)));
}

__global__ void __launch_bounds__(kThreads) perspective_kernel(const float2* __restrict__ P, const float2* __restrict__ C,
                                                               const int* __restrict__ n_valid, int n_pts,
                                                               unsigned char* __restrict__ flags /* [pairs][n_pts] */,
                                                               FitOut* __restrict__ out) {
  __shared__ double s_H[kBatch][9];
  __shared__ float s_Hf[kBatch][8];
  __shared__ int s_valid[kBatch];
  __shared__ int s_idx[kBatch][4];
  __shared__ double s_best[9];
  __shared__ double s_red[kWarps * kLmSums];
  __shared__ int s_ired[kWarps * kBatch];
  __shared__ int s_ctl[5];  // done, best_count, it, niters, failed
  __shared__ CvRng s_rng;
  __shared__ double s_x[8], s_xd[8], s_scal[8];
  __shared__ float s_rmax[kWarps];
  const int pair = blockIdx.x;
  const int n = n_valid[pair];
  const float2* p = P + (size_t)pair * n_pts;
  const float2* c = C + (size_t)pair * n_pts;
  unsigned char* flag = flags + (size_t)pair * n_pts;
  FitOut* o = out + (size_t)pair * 3 + VSTAB_MODE_PERSPECTIVE;
  const float thr = 2.5f * 2.5f;
  if (threadIdx.x == 0) {
    s_rng.state = 0xffffffffffffffffULL;
    s_ctl[0] = 0; s_ctl[1] = 0; s_ctl[2] = 0; s_ctl[3] = kMaxIters; s_ctl[4] = 0;
    o->n_valid = n; o->n_total = n_pts; o->ok = 0; o->n_inliers = 0; o->residual = 0;
    for (int k = 0; k < 9; k++) o->matrix[k] = (k % 4 == 0) ? 1.0 : 0.0;
  }
  __syncthreads();
  if (n < 4) return;

  if (n == 4) {
    if (threadIdx.x == 0) {
      float2 M[4] = {p[0], p[1], p[2], p[3]}, m[4] = {c[0], c[1], c[2], c[3]};
      if (homography_from_4(M, m, s_best)) s_ctl[1] = 4;
    }
    __syncthreads();
  } else {
    while (true) {
      // 1. sequential sampler (RNG stream + subset checks), one thread
      if (threadIdx.x == 0) {
        for (int b = 0; b < kBatch; b++) {
          bool found = false;
          for (int attempt = 0; attempt < 10000 && !found; attempt++) {
            int idx[4];
            float2 M[4], m[4];
            for (int i = 0; i < 4; i++) {
              int v;
              bool dup;
              do {
                v = s_rng.uniform(0, n);
                dup = false;
                for (int j = 0; j < i; j++) dup = dup || idx[j] == v;
              } while (dup);
              idx[i] = v;
              M[i] = p[v];
              m[i] = c[v];
            }
            if (homography_subset_ok(M, m)) {
              found = true;
              for (int i = 0; i < 4; i++) s_idx[b][i] = idx[i];
            }
          }
          s_valid[b] = found ? 1 : -1;  // -1: sampler exhausted => RANSAC stops here
        }
      }
      __syncthreads();
      // 2. minimal-sample models, one thread per hypothesis
      if (threadIdx.x < kBatch && s_valid[threadIdx.x] == 1) {
        const int b = threadIdx.x;
        float2 M[4], m[4];
        for (int i = 0; i < 4; i++) { M[i] = p[s_idx[b][i]]; m[i] = c[s_idx[b][i]]; }
        double H[9];
        if (homography_from_4(M, m, H)) {
          for (int k = 0; k < 9; k++) s_H[b][k] = H[k];
          for (int k = 0; k < 8; k++) s_Hf[b][k] = (float)H[k];
        } else {
          s_valid[b] = 0;  // runKernel returned no model: iteration counted, nothing scored
        }
      }
      __syncthreads();
      // 3. score
      int cnt[kBatch];
#pragma unroll
      for (int b = 0; b < kBatch; b++) cnt[b] = 0;
      for (int i = threadIdx.x; i < n; i += kThreads) {
        const float2 pp = p[i], cc = c[i];
#pragma unroll
        for (int b = 0; b < kBatch; b++)
          if (s_valid[b] == 1) cnt[b] += homography_inlier(s_Hf[b], pp, cc, thr) ? 1 : 0;
      }
      block_sum_int<kBatch>(cnt, s_ired);
      // 4. replay the sequential accept / terminate rule
      if (threadIdx.x == 0) {
        int best = s_ctl[1], it = s_ctl[2], niters = s_ctl[3];
        bool stop = false;
        for (int b = 0; b < kBatch && it < niters && !stop; b++) {
          if (s_valid[b] == -1) {  // getSubset failed
            if (it == 0) s_ctl[4] = 1;
            stop = true;
            break;
          }
          it++;
          if (s_valid[b] == 0) continue;
          const int good = cnt[b];
          if (good > max(best, 3)) {
            best = good;
            for (int k = 0; k < 9; k++) s_best[k] = s_H[b][k];
            niters = ransac_update_num_iters(kConfidence, (double)(n - good) / n, 4, niters);
          }
        }
        s_ctl[1] = best; s_ctl[2] = it; s_ctl[3] = niters;
        s_ctl[0] = stop || it >= niters;
      }
      __syncthreads();
      if (s_ctl[0]) break;
    }
  }
  if (s_ctl[4] || s_ctl[1] <= 0) return;

  // consensus set of the winning hypothesis
  {
    float Hf[8];
    for (int k = 0; k < 8; k++) Hf[k] = (float)s_best[k];
    for (int i = threadIdx.x; i < n; i += kThreads) flag[i] = homography_inlier(Hf, p[i], c[i], thr) ? 1 : 0;
  }
  __syncthreads();

  if (n > 4) {
    // normalised DLT over the consensus set
    double a[kLmSums];
    for (int k = 0; k < kLmSums; k++) a[k] = 0;
    for (int i = threadIdx.x; i < n; i += kThreads)
      if (flag[i]) { a[0] += c[i].x; a[1] += c[i].y; a[2] += p[i].x; a[3] += p[i].y; a[4] += 1.0; }
    block_sum<kLmSums>(a, s_red);
    const double cnt = a[4];
    const double cmx = a[0] / cnt, cmy = a[1] / cnt, cMx = a[2] / cnt, cMy = a[3] / cnt;
    for (int k = 0; k < kLmSums; k++) a[k] = 0;
    for (int i = threadIdx.x; i < n; i += kThreads)
      if (flag[i]) {
        a[0] += fabs(c[i].x - cmx); a[1] += fabs(c[i].y - cmy); a[2] += fabs(p[i].x - cMx); a[3] += fabs(p[i].y - cMy);
      }
    block_sum<kLmSums>(a, s_red);
    bool degenerate = fabs(a[0]) < DBL_EPSILON || fabs(a[1]) < DBL_EPSILON || fabs(a[2]) < DBL_EPSILON || fabs(a[3]) < DBL_EPSILON;
    if (degenerate) return;
    const double smx = cnt / a[0], smy = cnt / a[1], sMx = cnt / a[2], sMy = cnt / a[3];
    for (int k = 0; k < kLmSums; k++) a[k] = 0;
    for (int i = threadIdx.x; i < n; i += kThreads)
      if (flag[i]) {
        const double x = (c[i].x - cmx) * smx, y = (c[i].y - cmy) * smy;
        const double X = (p[i].x - cMx) * sMx, Y = (p[i].y - cMy) * sMy;
        const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
        const double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
        int k = 0;
        for (int j = 0; j < 9; j++)
          for (int l = j; l < 9; l++) a[k++] += Lx[j] * Lx[l] + Ly[j] * Ly[l];
      }
    block_sum<kLmSums>(a, s_red);
    if (threadIdx.x == 0) {
      double L[9][9];
      int k = 0;
      for (int j = 0; j < 9; j++)
        for (int l = j; l < 9; l++) { L[j][l] = a[k]; L[l][j] = a[k]; k++; }
      double h0[9], H[9];
      smallest_eigvec9(L, h0);
      denormalise_h(h0, cMx, cMy, sMx, sMy, cmx, cmy, smx, smy, H);
      for (int q = 0; q < 8; q++) s_x[q] = H[q];
    }
    __syncthreads();

    // Levenberg-Marquardt, cv::LMSolverImpl::run schedule, <= 10 iterations
    double S = 0, lambda = 1, lc = 0.75;
    double A[8][8], v[8], D[8];
    {
      double acc[kLmSums];
      for (int k = 0; k < kLmSums; k++) acc[k] = 0;
      double rmax = 0;
      for (int i = threadIdx.x; i < n; i += kThreads)
        if (flag[i]) lm_accumulate(s_x, p[i], c[i], acc, rmax);
      block_sum<kLmSums>(acc, s_red);
      int k = 0;
      for (int i = 0; i < 8; i++)
        for (int j = i; j < 8; j++) { A[i][j] = acc[k]; A[j][i] = acc[k]; k++; }
      for (int i = 0; i < 8; i++) { v[i] = acc[36 + i]; D[i] = A[i][i]; }
      S = acc[44];
      float rm = (float)rmax;
      for (int ofs = 16; ofs > 0; ofs >>= 1) rm = fmaxf(rm, __shfl_down_sync(0xffffffffu, rm, ofs));
      if ((threadIdx.x & 31) == 0) s_rmax[threadIdx.x >> 5] = rm;
      __syncthreads();
      if (threadIdx.x == 0) {
        double r0 = 0;
        for (int w = 0; w < kWarps; w++) r0 = fmax(r0, (double)s_rmax[w]);
        s_scal[0] = r0;
      }
      __syncthreads();
    }
    for (int iter = 0; iter < 10;) {
      // every thread holds identical A, v, D, S, lambda: solve redundantly (no divergence in data)
      double Ap[8][8], vv[8], d[8];
      for (int i = 0; i < 8; i++) {
        for (int j = 0; j < 8; j++) Ap[i][j] = A[i][j];
        Ap[i][i] += lambda * D[i];
        vv[i] = v[i];
      }
      if (!solve_small(Ap, vv, d, 8)) break;
      if (threadIdx.x == 0)
        for (int q = 0; q < 8; q++) s_xd[q] = s_x[q] - d[q];
      __syncthreads();
      double acc[kLmSums];
      for (int k = 0; k < kLmSums; k++) acc[k] = 0;
      double rmax = 0;
      for (int i = threadIdx.x; i < n; i += kThreads)
        if (flag[i]) lm_accumulate(s_xd, p[i], c[i], acc, rmax);
      block_sum<kLmSums>(acc, s_red);
      // max |r| at xd across the block
      float rm = (float)rmax;
      for (int ofs = 16; ofs > 0; ofs >>= 1) rm = fmaxf(rm, __shfl_down_sync(0xffffffffu, rm, ofs));
      if ((threadIdx.x & 31) == 0) s_rmax[threadIdx.x >> 5] = rm;
      __syncthreads();
      double rmax_xd = 0;
      for (int w = 0; w < kWarps; w++) rmax_xd = fmax(rmax_xd, (double)s_rmax[w]);
      __syncthreads();
      const double Sd = acc[44];
      double dS = 0, tdv = 0, dmax = 0;
      for (int i = 0; i < 8; i++) {
        double t = 0;
        for (int j = 0; j < 8; j++) t += A[i][j] * d[j];
        dS += d[i] * (-t + 2 * v[i]);
        tdv += d[i] * v[i];
        dmax = fmax(dmax, fabs(d[i]));
      }
      const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
      if (R > 0.75) {
        lambda *= 0.5;
        if (lambda < lc) lambda = 0;
      } else if (R < 0.25) {
        double nu = (Sd - S) / (fabs(tdv) > DBL_EPSILON ? tdv : 1) + 2;
        nu = fmin(fmax(nu, 2.), 10.);
        if (lambda == 0) {
          // lambda = lc = 1 / max |diag(A^-1)|
          double maxval = DBL_EPSILON;
          for (int col = 0; col < 8; col++) {
            double Ai[8][8], e[8], xcol[8];
            for (int i = 0; i < 8; i++) { for (int j = 0; j < 8; j++) Ai[i][j] = A[i][j]; e[i] = i == col ? 1.0 : 0.0; }
            if (solve_small(Ai, e, xcol, 8)) maxval = fmax(maxval, fabs(xcol[col]));
          }
          lambda = lc = 1. / maxval;
          nu *= 0.5;
        }
        lambda *= nu;
      }
      double rcur_max;
      if (Sd < S) {
        S = Sd;
        if (threadIdx.x == 0)
          for (int q = 0; q < 8; q++) s_x[q] = s_xd[q];
        int k = 0;
        for (int i = 0; i < 8; i++)
          for (int j = i; j < 8; j++) { A[i][j] = acc[k]; A[j][i] = acc[k]; k++; }
        for (int i = 0; i < 8; i++) v[i] = acc[36 + i];
        rcur_max = rmax_xd;
      } else {
        rcur_max = s_scal[0];
      }
      __syncthreads();
      if (threadIdx.x == 0) s_scal[0] = rcur_max;
      __syncthreads();
      iter++;
      const bool proceed = iter < 10 && dmax >= (double)FLT_EPSILON && rcur_max >= (double)FLT_EPSILON;
      if (!proceed) break;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int q = 0; q < 8; q++) s_best[q] = s_x[q];
      s_best[8] = 1.0;
    }
    __syncthreads();
  }

  // final mask with the refined model (cv2 >= 4.5), residual of the affine part (flow.py:174)
  float Hf[8];
  for (int k = 0; k < 8; k++) Hf[k] = (float)s_best[k];
  double r[8];
  for (int k = 0; k < 8; k++) r[k] = 0;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    const float2 pp = p[i], cc = c[i];
    r[0] += homography_inlier(Hf, pp, cc, thr) ? 1.0 : 0.0;
    r[1] += fabs((s_best[0] * pp.x + s_best[1] * pp.y + s_best[2]) - cc.x) +
            fabs((s_best[3] * pp.x + s_best[4] * pp.y + s_best[5]) - cc.y);
  }
  block_sum<8>(r, s_red);
  if (threadIdx.x == 0) {
    for (int k = 0; k < 9; k++) o->matrix[k] = s_best[k];
    o->n_inliers = (int)(r[0] + 0.5);
    o->residual = r[1] / (2.0 * n);
    o->ok = 1;
  }
}

}  // namespace

extern "C" int vstab_fit_batch(vstab_handle* h, const float* prev_dev, const float* curr_dev, int n_pairs, int n_pts,
                               int grid_w, int grid_h, int grid_step, int mode_mask, vstab_fit_result* out_dev,
                               void* stream) {
  if (!h) return vstab_fail(nullptr, VSTAB_ERR_INVALID, "vstab_fit_batch: null handle");
  if (!curr_dev || !out_dev || n_pairs < 0 || n_pts <= 0)
    return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_fit_batch: bad argument");
  if (!prev_dev && (grid_w <= 0 || grid_h <= 0 || grid_step <= 0 || grid_w * grid_h != n_pts))
    return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_fit_batch: grid description does not match n_pts");
  if (n_pts > 16384) return vstab_fail(h, VSTAB_ERR_UNSUPPORTED, "vstab_fit_batch: more than 16384 correspondences per pair");
  if (n_pairs == 0) return VSTAB_OK;
  static_assert(sizeof(FitOut) == sizeof(vstab_fit_result), "FitOut must mirror vstab_fit_result");
  cudaStream_t st = (cudaStream_t)stream;
  VSTAB_ENTER(h);
  // workspace: compacted correspondences, valid counts, consensus flags
  const size_t bytes_pts = sizeof(float2) * (size_t)n_pairs * n_pts;
  const size_t off_c = (bytes_pts + 255) & ~(size_t)255;
  const size_t off_n = off_c + ((bytes_pts + 255) & ~(size_t)255);
  const size_t off_f = off_n + (((size_t)n_pairs * 4 + 255) & ~(size_t)255);
  const size_t total = off_f + (size_t)n_pairs * n_pts;
  void* ws = nullptr;
  int rc = vstab_workspace(h, total, &ws);
  if (rc != VSTAB_OK) return rc;
  float2* P = (float2*)ws;
  float2* C = (float2*)((unsigned char*)ws + off_c);
  int* nvalid = (int*)((unsigned char*)ws + off_n);
  unsigned char* flags = (unsigned char*)ws + off_f;
  FitOut* out = (FitOut*)out_dev;

  prepare_kernel<<<n_pairs, kThreads, 0, st>>>(prev_dev, curr_dev, n_pts, grid_w, grid_step, P, C, nvalid);
  VSTAB_LAUNCH_CHECK(h, "fit prepare_kernel");
  // The three models of a pair are independent (disjoint outputs) and each kernel is a chain of short serial phases
  // per pair (RNG replay, 8x8 / 9x9 solves by one thread, LM steps), so they run side by side on forked streams:
  // 250 pairs of a perspective clip, all three models: 3.05 ms back to back.
  const bool want_t = mode_mask & (1 << VSTAB_MODE_TRANSLATION), want_s = mode_mask & (1 << VSTAB_MODE_SIMILARITY);
  const bool want_p = mode_mask & (1 << VSTAB_MODE_PERSPECTIVE);
  cudaStream_t st_s = st, st_p = st;
  const int forks = (want_s && want_t ? 1 : 0) + (want_p && (want_t || want_s) ? 1 : 0);
  if (forks > 0) {
    const int rcs = vstab_aux_streams(h, forks);
    if (rcs != VSTAB_OK) return rcs;
    VSTAB_CUDA(h, cudaEventRecord(h->fork_event, st));
    int k = 0;
    if (want_s && want_t) { st_s = h->aux_stream[k++]; VSTAB_CUDA(h, cudaStreamWaitEvent(st_s, h->fork_event, 0)); }
    if (want_p && (want_t || want_s)) { st_p = h->aux_stream[k++]; VSTAB_CUDA(h, cudaStreamWaitEvent(st_p, h->fork_event, 0)); }
  }
  if (want_t) {
    int cap = 1;
    while (cap < n_pts) cap <<= 1;
    const size_t smem = sizeof(unsigned) * 2 * cap;
    VSTAB_CUDA(h, cudaFuncSetAttribute(translation_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    translation_kernel<<<n_pairs, 1024, smem, st>>>(P, C, nvalid, n_pts, cap, out);
    VSTAB_LAUNCH_CHECK(h, "fit translation_kernel");
  }
  if (want_s) {
    similarity_kernel<<<n_pairs, kThreads, 0, st_s>>>(P, C, nvalid, n_pts, out);
    VSTAB_LAUNCH_CHECK(h, "fit similarity_kernel");
  }
  if (want_p) {
    perspective_kernel<<<n_pairs, kThreads, 0, st_p>>>(P, C, nvalid, n_pts, flags, out);
    VSTAB_LAUNCH_CHECK(h, "fit perspective_kernel");
  }
  {
    int k = 0;
    if (st_s != st) { VSTAB_CUDA(h, cudaEventRecord(h->join_event[k], st_s)); VSTAB_CUDA(h, cudaStreamWaitEvent(st, h->join_event[k], 0)); k++; }
    if (st_p != st) { VSTAB_CUDA(h, cudaEventRecord(h->join_event[k], st_p)); VSTAB_CUDA(h, cudaStreamWaitEvent(st, h->join_event[k], 0)); k++; }
  }
  return VSTAB_OK;
}
