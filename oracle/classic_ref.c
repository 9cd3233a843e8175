/*
 * classic_ref.c -- plain-C restatement of the two cv2 calls of the Classic estimator.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Reference call sites: nodes/video_stabilizer_classic.py:76-83
 *   cv2.goodFeaturesToTrack(prev_gray, maxCorners=400, qualityLevel=0.01, minDistance=7, blockSize=21)
 * and :88-96
 *   cv2.calcOpticalFlowPyrLK(prev_gray, curr_gray, features, None, winSize=(31,31), maxLevel=3,
 *                            criteria=(EPS|COUNT, 50, 0.01))
 * The arithmetic lives in opencv-python-headless (4.13.0.92 here); this restates the published
 * algorithms (Shi-Tomasi minimum-eigenvalue corners; Bouguet's pyramidal Lucas-Kanade as in
 * modules/video/src/lkpyramid.cpp) with the operation order identified black-box against the wheel:
 *   Sobel 3x3 scaled by 1/(4*21*255):  Dx = fma(r0 + r2, s, r1 * 2s),  rows of Dy = fma(s,c,fma(2s,b,s*a)) in the
 *   32-pixel blocks of a row and (s*a + 2s*b) + s*c, unfused, in the last w % 32 pixels
 *   unnormalised 21x21 box filter with DOUBLE running sums (row sums, then column sums)
 *   minEig = (a/2 + c/2) - sqrt((a/2 - c/2)^2 + b^2), threshold-to-zero at 0.01*max, 3x3 local maxima,
 *   descending sort (ties: higher address first), greedy 7-px minimum distance on a 7-px grid
 *   LK: pyrDown (1 4 6 4 1), Scharr derivatives, 14-bit fixed-point bilinear windows (x32 intensities)
 * Pinned in tests/test_oracle_classic.py against live cv2.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
  return p;
}

/* ------------------------------------------------------------------ min-eigenvalue map ---- */

void classicref_min_eigen(const uint8_t* img, int h, int w, int block, float* eig) {
  const double scale_d = 1.0 / ((double)(1 << 2) * block * 255.0);
  const float s = (float)scale_d, s2 = s * 2.0f;
  const size_t n = (size_t)h * w;
  float* dx = (float*)malloc(n * sizeof(float));
  float* dy = (float*)malloc(n * sizeof(float));
  float* rowsm = (float*)malloc(n * sizeof(float));
  /* Dy: row filter [s 2s s], then rows y+1 minus y-1.  The wheel's row filter works on blocks of 32 pixels with a
   * fused chain; the last w % 32 pixels of a row go through its scalar loop, which does not fuse. */
  const int vec_end = (w / 32) * 32;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      float a = img[y * w + reflect101(x - 1, w)], b = img[y * w + x], c = img[y * w + reflect101(x + 1, w)];
      rowsm[y * w + x] = x < vec_end ? fmaf(s, c, fmaf(s2, b, s * a)) : (s * a + s2 * b) + s * c;
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      dy[y * w + x] = rowsm[reflect101(y + 1, h) * w + x] - rowsm[reflect101(y - 1, h) * w + x];
  /* Dx: row filter [-1 0 1] (exact), column filter [s 2s s]: fma(r0 + r2, s, r1 * 2s) */
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) rowsm[y * w + x] = (float)((int)img[y * w + reflect101(x + 1, w)] - (int)img[y * w + reflect101(x - 1, w)]);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      float r0 = rowsm[reflect101(y - 1, h) * w + x], r1 = rowsm[y * w + x], r2 = rowsm[reflect101(y + 1, h) * w + x];
      dx[y * w + x] = fmaf(r0 + r2, s, r1 * s2);
    }
  /* covariance + box filter with double running sums (reflect-101 border) */
  const int r = block / 2;
  double* rows = (double*)malloc(sizeof(double) * (size_t)(h + 2 * r) * w * 3);
  for (int yy = 0; yy < h + 2 * r; yy++) {
    const int y = reflect101(yy - r, h);
    double acc[3] = {0, 0, 0};
    for (int i = 0; i < block; i++) {
      const int x = reflect101(i - r, w);
      const float gx = dx[y * w + x], gy = dy[y * w + x];
      acc[0] += (double)(gx * gx); acc[1] += (double)(gx * gy); acc[2] += (double)(gy * gy);
    }
    double* o = rows + (size_t)yy * w * 3;
    o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
    for (int x = 1; x < w; x++) {
      const int xn = reflect101(x + block - 1 - r, w), xo = reflect101(x - 1 - r, w);
      const float gxn = dx[y * w + xn], gyn = dy[y * w + xn], gxo = dx[y * w + xo], gyo = dy[y * w + xo];
      acc[0] += (double)(gxn * gxn) - (double)(gxo * gxo);
      acc[1] += (double)(gxn * gyn) - (double)(gxo * gyo);
      acc[2] += (double)(gyn * gyn) - (double)(gyo * gyo);
      o[x * 3] = acc[0]; o[x * 3 + 1] = acc[1]; o[x * 3 + 2] = acc[2];
    }
  }
  double* SUM = (double*)calloc((size_t)w * 3, sizeof(double));
  for (int i = 0; i < block - 1; i++)
    for (int k = 0; k < w * 3; k++) SUM[k] += rows[(size_t)i * w * 3 + k];
  for (int y = 0; y < h; y++) {
    const double* Sp = rows + (size_t)(y + block - 1) * w * 3;
    const double* Sm = rows + (size_t)y * w * 3;
    for (int x = 0; x < w; x++) {
      float c3[3];
      for (int k = 0; k < 3; k++) {
        double s0 = SUM[x * 3 + k] + Sp[x * 3 + k];
        c3[k] = (float)s0;
        SUM[x * 3 + k] = s0 - Sm[x * 3 + k];
      }
      float a = c3[0] * 0.5f, b = c3[1], c = c3[2] * 0.5f;
      eig[y * w + x] = (a + c) - sqrtf((a - c) * (a - c) + b * b);
    }
  }
  free(dx); free(dy); free(rowsm); free(rows); free(SUM);
}

/* --------------------------------------------------------------- goodFeaturesToTrack ---- */

typedef struct { float v; int idx; } cand;
static int cand_cmp(const void* pa, const void* pb) {
  const cand* a = (const cand*)pa; const cand* b = (const cand*)pb;
  if (a->v > b->v) return -1;
  if (a->v < b->v) return 1;
  return a->idx > b->idx ? -1 : (a->idx < b->idx ? 1 : 0);
}

/* returns the number of corners written to xy (x0,y0,x1,y1,...) */
int classicref_good_features(const uint8_t* img, int h, int w, int max_corners, double quality, double min_distance,
                             int block, float* xy) {
  const size_t n = (size_t)h * w;
  float* eig = (float*)malloc(n * sizeof(float));
  classicref_min_eigen(img, h, w, block, eig);
  double maxv = -DBL_MAX;
  for (size_t k = 0; k < n; k++) if (eig[k] > maxv) maxv = eig[k];
  const float thr = (float)(maxv * quality);
  for (size_t k = 0; k < n; k++) if (!(eig[k] > thr)) eig[k] = 0.f;
  cand* cs = (cand*)malloc(sizeof(cand) * n);
  int nc = 0;
  for (int y = 1; y < h - 1; y++)
    for (int x = 1; x < w - 1; x++) {
      const float v = eig[y * w + x];
      if (v == 0.f) continue;
      float m = v;
      for (int j = -1; j <= 1; j++)
        for (int i = -1; i <= 1; i++) { const float t = eig[(y + j) * w + x + i]; if (t > m) m = t; }
      if (v == m) { cs[nc].v = v; cs[nc].idx = y * w + x; nc++; }
    }
  qsort(cs, nc, sizeof(cand), cand_cmp);
  int ncorners = 0;
  if (min_distance >= 1) {
    const int cell = (int)lrint(min_distance);
    const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
    const int cap = 8;
    int* cnt = (int*)calloc((size_t)gw * gh, sizeof(int));
    float* pts = (float*)malloc(sizeof(float) * (size_t)gw * gh * cap * 2);
    const double md2 = min_distance * min_distance;
    for (int i = 0; i < nc; i++) {
      const int y = cs[i].idx / w, x = cs[i].idx - y * w;
      const int xc = x / cell, yc = y / cell;
      int x1 = xc - 1, y1 = yc - 1, x2 = xc + 1, y2 = yc + 1;
      if (x1 < 0) x1 = 0; if (y1 < 0) y1 = 0; if (x2 > gw - 1) x2 = gw - 1; if (y2 > gh - 1) y2 = gh - 1;
      int good = 1;
      for (int yy = y1; yy <= y2 && good; yy++)
        for (int xx = x1; xx <= x2 && good; xx++) {
          const int c = yy * gw + xx;
          for (int j = 0; j < cnt[c]; j++) {
            const float ddx = x - pts[(c * cap + j) * 2], ddy = y - pts[(c * cap + j) * 2 + 1];
            if (ddx * ddx + ddy * ddy < md2) { good = 0; break; }
          }
        }
      if (good) {
        const int c = yc * gw + xc;
        if (cnt[c] < cap) { pts[(c * cap + cnt[c]) * 2] = (float)x; pts[(c * cap + cnt[c]) * 2 + 1] = (float)y; cnt[c]++; }
        xy[ncorners * 2] = (float)x; xy[ncorners * 2 + 1] = (float)y;
        ncorners++;
        if (max_corners > 0 && ncorners == max_corners) break;
      }
    }
    free(cnt); free(pts);
  } else {
    for (int i = 0; i < nc; i++) {
      const int y = cs[i].idx / w, x = cs[i].idx - y * w;
      xy[ncorners * 2] = (float)x; xy[ncorners * 2 + 1] = (float)y;
      ncorners++;
      if (max_corners > 0 && ncorners == max_corners) break;
    }
  }
  free(eig); free(cs);
  return ncorners;
}

/* ----------------------------------------------------------------- pyramidal Lucas-Kanade ---- */

static void pyr_down(const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw) {
  /* cv::pyrDown for 8-bit: separable [1 4 6 4 1], integer sums, (sum + 128) >> 8, reflect-101 */
  int* tmp = (int*)malloc(sizeof(int) * (size_t)sh * dw);
  for (int y = 0; y < sh; y++)
    for (int x = 0; x < dw; x++) {
      const int c = 2 * x;
      const uint8_t* r = src + (size_t)y * sw;
      tmp[y * dw + x] = r[reflect101(c - 2, sw)] + 4 * r[reflect101(c - 1, sw)] + 6 * r[reflect101(c, sw)] +
                        4 * r[reflect101(c + 1, sw)] + r[reflect101(c + 2, sw)];
    }
  for (int y = 0; y < dh; y++)
    for (int x = 0; x < dw; x++) {
      const int c = 2 * y;
      const int s = tmp[reflect101(c - 2, sh) * dw + x] + 4 * tmp[reflect101(c - 1, sh) * dw + x] + 6 * tmp[reflect101(c, sh) * dw + x] +
                    4 * tmp[reflect101(c + 1, sh) * dw + x] + tmp[reflect101(c + 2, sh) * dw + x];
      dst[(size_t)y * dw + x] = (uint8_t)((s + 128) >> 8);
    }
  free(tmp);
}

/* Scharr derivatives, int16 interleaved (dx, dy), reflect-101 inside the image */
static void scharr(const uint8_t* I, int h, int w, int16_t* d) {
  for (int y = 0; y < h; y++) {
    const uint8_t* r0 = I + (size_t)reflect101(y - 1, h) * w;
    const uint8_t* r1 = I + (size_t)y * w;
    const uint8_t* r2 = I + (size_t)reflect101(y + 1, h) * w;
    for (int x = 0; x < w; x++) {
      const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
      const int t0p = (r0[xp] + r2[xp]) * 3 + r1[xp] * 10, t0m = (r0[xm] + r2[xm]) * 3 + r1[xm] * 10;
      const int t1m = r2[xm] - r0[xm], t1c = r2[x] - r0[x], t1p = r2[xp] - r0[xp];
      d[((size_t)y * w + x) * 2] = (int16_t)(t0p - t0m);
      d[((size_t)y * w + x) * 2 + 1] = (int16_t)((t1p + t1m) * 3 + t1c * 10);
    }
  }
}

static inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

/* image sample with reflect-101 border (the pyramid levels carry a winSize border made that way) */
static inline int pix(const uint8_t* I, int h, int w, int y, int x) { return I[(size_t)reflect101(y, h) * w + reflect101(x, w)]; }
/* derivative sample: zero outside the image (BORDER_CONSTANT) */
static inline int dpix(const int16_t* D, int h, int w, int y, int x, int c) {
  if (x < 0 || y < 0 || x >= w || y >= h) return 0;
  return D[((size_t)y * w + x) * 2 + c];
}

/* prev/next: float [n][2]; status: uint8 [n].  win = 31, max_level = 3, 50 iterations, eps 0.01.
 * cv_order = 0: the window sums (structure tensor A11/A12/A22, mismatch vector b1/b2) run serially over the window --
 *   what the CUDA tracker is checked against; positions within float rounding of cv2 (<= 2e-3 px, same status flags).
 * cv_order = 1: the sums in the order of cv2's 128-bit SIMD code, found black-box (minEig through
 *   OPTFLOW_LK_GET_MIN_EIGENVALS, positions after one iteration): per row the first 24 of the 31 columns go to four
 *   lane accumulators (A: lane = x mod 4, reduced (l0+l2)+(l1+l3); b: int32 pair sums d[x]g[x]+d[x+4]g[x+4] converted
 *   to float, eight lanes folded pairwise), columns 24..30 to one scalar accumulator carried across the rows.
 *   Bit-exact against the wheel: 9288 of 9288 tracks over 30 random frame sizes (tests/test_oracle_classic.py). */
static void pyr_lk_body(const uint8_t* I0, const uint8_t* J0, int h, int w, const float* prev_pts, int n, int win, int max_level,
                        int max_iter, double eps, float* next_pts, uint8_t* status, int cv_order) {
  enum { MAXL = 8 };
  uint8_t* Ip[MAXL]; uint8_t* Jp[MAXL]; int16_t* Dp[MAXL]; int hh[MAXL], ww[MAXL];
  int levels = 0;
  Ip[0] = (uint8_t*)I0; Jp[0] = (uint8_t*)J0; hh[0] = h; ww[0] = w;
  for (int l = 1; l <= max_level; l++) {
    const int dh = (hh[l - 1] + 1) / 2, dw = (ww[l - 1] + 1) / 2;
    if (dw <= win || dh <= win) break; /* buildOpticalFlowPyramid stops when the level is not larger than the window */
    Ip[l] = (uint8_t*)malloc((size_t)dh * dw); Jp[l] = (uint8_t*)malloc((size_t)dh * dw);
    pyr_down(Ip[l - 1], hh[l - 1], ww[l - 1], Ip[l], dh, dw);
    pyr_down(Jp[l - 1], hh[l - 1], ww[l - 1], Jp[l], dh, dw);
    hh[l] = dh; ww[l] = dw; levels = l;
  }
  for (int l = 0; l <= levels; l++) { Dp[l] = (int16_t*)malloc(sizeof(int16_t) * 2 * (size_t)hh[l] * ww[l]); scharr(Ip[l], hh[l], ww[l], Dp[l]); }
  const float half = (win - 1) * 0.5f;
  const double eps2 = eps * eps;
  const int W_BITS = 14;
  const float FLT_SCALE = 1.f / (1 << 20);
  int16_t* IWin = (int16_t*)malloc(sizeof(int16_t) * win * win);
  int16_t* DWin = (int16_t*)malloc(sizeof(int16_t) * win * win * 2);
  for (int p = 0; p < n; p++) status[p] = 1;
  for (int level = levels; level >= 0; level--) {
    const uint8_t* I = Ip[level]; const uint8_t* J = Jp[level]; const int16_t* D = Dp[level];
    const int lh = hh[level], lw = ww[level];
    for (int p = 0; p < n; p++) {
      float px = prev_pts[p * 2] * (float)(1. / (1 << level)), py = prev_pts[p * 2 + 1] * (float)(1. / (1 << level));
      float nx, ny;
      if (level == levels) { nx = px; ny = py; }
      else { nx = next_pts[p * 2] * 2.f; ny = next_pts[p * 2 + 1] * 2.f; }
      next_pts[p * 2] = nx; next_pts[p * 2 + 1] = ny;
      px -= half; py -= half;
      int ipx = (int)floorf(px), ipy = (int)floorf(py);
      if (ipx < -win || ipx >= lw || ipy < -win || ipy >= lh) { if (level == 0) status[p] = 0; continue; }
      float a = px - ipx, b = py - ipy;
      int iw00 = (int)lrintf((1.f - a) * (1.f - b) * (1 << W_BITS)), iw01 = (int)lrintf(a * (1.f - b) * (1 << W_BITS));
      int iw10 = (int)lrintf((1.f - a) * b * (1 << W_BITS)), iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
      float A11 = 0, A12 = 0, A22 = 0, q11[4] = {0, 0, 0, 0}, q12[4] = {0, 0, 0, 0}, q22[4] = {0, 0, 0, 0};
      for (int y = 0; y < win; y++)
        for (int x = 0; x < win; x++) {
          const int sy = ipy + y, sx = ipx + x;
          const int ival = descale(pix(I, lh, lw, sy, sx) * iw00 + pix(I, lh, lw, sy, sx + 1) * iw01 + pix(I, lh, lw, sy + 1, sx) * iw10 +
                                   pix(I, lh, lw, sy + 1, sx + 1) * iw11, W_BITS - 5);
          const int ixval = descale(dpix(D, lh, lw, sy, sx, 0) * iw00 + dpix(D, lh, lw, sy, sx + 1, 0) * iw01 +
                                    dpix(D, lh, lw, sy + 1, sx, 0) * iw10 + dpix(D, lh, lw, sy + 1, sx + 1, 0) * iw11, W_BITS);
          const int iyval = descale(dpix(D, lh, lw, sy, sx, 1) * iw00 + dpix(D, lh, lw, sy, sx + 1, 1) * iw01 +
                                    dpix(D, lh, lw, sy + 1, sx, 1) * iw10 + dpix(D, lh, lw, sy + 1, sx + 1, 1) * iw11, W_BITS);
          IWin[y * win + x] = (int16_t)ival; DWin[(y * win + x) * 2] = (int16_t)ixval; DWin[(y * win + x) * 2 + 1] = (int16_t)iyval;
          if (cv_order && x < (win / 8) * 8) { const int l = x & 3; q11[l] += (float)(ixval * ixval); q12[l] += (float)(ixval * iyval); q22[l] += (float)(iyval * iyval); }
          else { A11 += (float)(ixval * ixval); A12 += (float)(ixval * iyval); A22 += (float)(iyval * iyval); }
        }
      A11 += (q11[0] + q11[2]) + (q11[1] + q11[3]); A12 += (q12[0] + q12[2]) + (q12[1] + q12[3]); A22 += (q22[0] + q22[2]) + (q22[1] + q22[3]);
      A11 *= FLT_SCALE; A12 *= FLT_SCALE; A22 *= FLT_SCALE;
      float Dd = A11 * A22 - A12 * A12;
      const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
      if (minEig < 1e-4f || Dd < FLT_EPSILON) { if (level == 0) status[p] = 0; continue; }
      Dd = 1.f / Dd;
      nx -= half; ny -= half;
      float pdx = 0, pdy = 0;
      for (int j = 0; j < max_iter; j++) {
        const int inx = (int)floorf(nx), iny = (int)floorf(ny);
        if (inx < -win || inx >= lw || iny < -win || iny >= lh) { if (level == 0) status[p] = 0; break; }
        a = nx - inx; b = ny - iny;
        iw00 = (int)lrintf((1.f - a) * (1.f - b) * (1 << W_BITS)); iw01 = (int)lrintf(a * (1.f - b) * (1 << W_BITS));
        iw10 = (int)lrintf((1.f - a) * b * (1 << W_BITS)); iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
        float b1 = 0, b2 = 0, qb0[4] = {0, 0, 0, 0}, qb1[4] = {0, 0, 0, 0};
        for (int y = 0; y < win; y++) {
          int dd[64];
          for (int x = 0; x < win; x++) {
            const int sy = iny + y, sx = inx + x;
            dd[x] = descale(pix(J, lh, lw, sy, sx) * iw00 + pix(J, lh, lw, sy, sx + 1) * iw01 + pix(J, lh, lw, sy + 1, sx) * iw10 +
                            pix(J, lh, lw, sy + 1, sx + 1) * iw11, W_BITS - 5) - IWin[y * win + x];
          }
          const int16_t* g0 = DWin + y * win * 2;
          int x = 0;
          for (; cv_order && x <= win - 8; x += 8) {
            const int* d = dd + x; const int16_t* g = g0 + x * 2;
            qb0[0] += (float)(d[0] * g[0] + d[4] * g[8]);  qb0[1] += (float)(d[0] * g[1] + d[4] * g[9]);
            qb0[2] += (float)(d[1] * g[2] + d[5] * g[10]); qb0[3] += (float)(d[1] * g[3] + d[5] * g[11]);
            qb1[0] += (float)(d[2] * g[4] + d[6] * g[12]); qb1[1] += (float)(d[2] * g[5] + d[6] * g[13]);
            qb1[2] += (float)(d[3] * g[6] + d[7] * g[14]); qb1[3] += (float)(d[3] * g[7] + d[7] * g[15]);
          }
          for (; x < win; x++) { b1 += (float)(dd[x] * g0[x * 2]); b2 += (float)(dd[x] * g0[x * 2 + 1]); }
        }
        b1 += (qb0[0] + qb1[0]) + (qb0[2] + qb1[2]); b2 += (qb0[1] + qb1[1]) + (qb0[3] + qb1[3]);
        b1 *= FLT_SCALE; b2 *= FLT_SCALE;
        const float ddx = (float)((A12 * b2 - A22 * b1) * Dd), ddy = (float)((A12 * b1 - A11 * b2) * Dd);
        nx += ddx; ny += ddy;
        next_pts[p * 2] = nx + half; next_pts[p * 2 + 1] = ny + half;
        if ((double)ddx * ddx + (double)ddy * ddy <= eps2) break;
        if (j > 0 && fabsf(ddx + pdx) < 0.01f && fabsf(ddy + pdy) < 0.01f) {
          next_pts[p * 2] -= ddx * 0.5f; next_pts[p * 2 + 1] -= ddy * 0.5f;
          break;
        }
        pdx = ddx; pdy = ddy;
      }
      if (status[p] && level == 0) { /* the err computation's range check (err is always requested from Python) */
        const float fx = next_pts[p * 2] - half, fy = next_pts[p * 2 + 1] - half;
        const int ix = (int)floorf(fx), iy = (int)floorf(fy);
        if (ix < -win || ix >= lw || iy < -win || iy >= lh) status[p] = 0;
      }
    }
  }
  for (int l = 0; l <= levels; l++) { free(Dp[l]); if (l > 0) { free(Ip[l]); free(Jp[l]); } }
  free(IWin); free(DWin);
}

void classicref_pyr_lk(const uint8_t* I0, const uint8_t* J0, int h, int w, const float* prev_pts, int n, int win, int max_level,
                       int max_iter, double eps, float* next_pts, uint8_t* status) {
  pyr_lk_body(I0, J0, h, w, prev_pts, n, win, max_level, max_iter, eps, next_pts, status, 0);
}

void classicref_pyr_lk_exact(const uint8_t* I0, const uint8_t* J0, int h, int w, const float* prev_pts, int n, int win, int max_level,
                             int max_iter, double eps, float* next_pts, uint8_t* status) {
  pyr_lk_body(I0, J0, h, w, prev_pts, n, win, max_level, max_iter, eps, next_pts, status, 1);
}
