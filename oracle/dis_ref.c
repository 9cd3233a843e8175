/*
 * dis_ref.c -- plain-C restatement of cv2.DISOpticalFlow as the reference configures it.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): never linked into libvstab.so.
 *
 * Reference call sites: nodes/video_stabilizer_flow.py:76-87 (_create_flow_backend:
 * PRESET_MEDIUM + finest_scale 2, patch 8, stride 4, spatial propagation) and :140
 * (backend.calc(prev, curr, None)).  The arithmetic lives in the un-vendored dependency
 * opencv-python-headless (pinned range >=4.8,<5; 4.13.0.92 installed): this file restates the
 * published algorithm of modules/video/src/dis_flow.cpp and variational_refinement.cpp
 * (Kroeger et al., "Fast Optical Flow using Dense Inverse Search", ECCV 2016; Brox et al. 2004
 * for the refinement) in the operation order of OpenCV's 128-bit SIMD code path, so that the
 * result can be compared bit-for-bit with the wheel.  Pinned in tests/test_oracle_dis.py
 * against live cv2 and against tests/golden/dis_*.npz.
 *
 * Stages (per pyramid level, coarse to fine):
 *   1. INTER_AREA pyramid of both frames, Sobel 3x3 gradients of I0, I1 padded 16 px (replicate)
 *   2. per-patch structure tensor sums (8x8 patches on a stride-4 grid)
 *   3. patch inverse search: 2 passes (raster / reverse raster) in 8 fixed horizontal stripes,
 *      spatial propagation from left/up (right/down) neighbours, <=12 inverse-compositional
 *      steps with mean-normalised residuals
 *   4. densification weighted by 1/max(1,|I1(x+u)-I0(x)|)
 *   5. variational refinement: 5 fixed-point iterations x 5 red-black SOR sweeps (omega 1.6)
 *   6. bilinear x2 upsampling of the flow to the next level; final x4 upsampling
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DIS_EPS 0.001f
#define DIS_INF 1e10f
#define BORDER 16

typedef struct {
  int finest_scale;  /* 2 */
  int patch_size;    /* 8 (the SIMD summation order below assumes 8) */
  int patch_stride;  /* 4 */
  int gd_iter;       /* 25 */
  int vr_iter;       /* 5 */
  float vr_alpha;    /* 20 */
  float vr_delta;    /* 5 */
  float vr_gamma;    /* 10 */
  float vr_epsilon;  /* 0.01 */
  int use_mean_norm; /* 1 */
  int use_spatial;   /* 1 */
  int sor_iter;      /* 5 */
  float omega;       /* 1.6 */
  int reserved;
} dis_params;

void disref_default_params(dis_params* p) {
  p->finest_scale = 2;
  p->patch_size = 8;
  p->patch_stride = 4;
  p->gd_iter = 25;
  p->vr_iter = 5;
  p->vr_alpha = 20.f;
  p->vr_delta = 5.f;
  p->vr_gamma = 10.f;
  p->vr_epsilon = 0.01f;
  p->use_mean_norm = 1;
  p->use_spatial = 1;
  p->sor_iter = 5;
  p->omega = 1.6f;
  p->reserved = 0;
}

/* ------------------------------------------------------------------ INTER_AREA (uint8) ---- */

typedef struct { int di, si; float alpha; } area_ent;

static int area_tab(int ssize, int dsize, double scale, area_ent* tab) {
  int k = 0;
  for (int dx = 0; dx < dsize; dx++) {
    double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    double cell = scale < ssize - fsx1 ? scale : ssize - fsx1;
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    if (sx2 > ssize - 1) sx2 = ssize - 1;
    if (sx1 > sx2) sx1 = sx2;
    if (sx1 - fsx1 > 1e-3) { tab[k].di = dx; tab[k].si = sx1 - 1; tab[k++].alpha = (float)((sx1 - fsx1) / cell); }
    for (int sx = sx1; sx < sx2; sx++) { tab[k].di = dx; tab[k].si = sx; tab[k++].alpha = (float)(1.0 / cell); }
    if (fsx2 - sx2 > 1e-3) {
      double a = fsx2 - sx2; if (a > 1.) a = 1.; if (a > cell) a = cell;
      tab[k].di = dx; tab[k].si = sx2; tab[k++].alpha = (float)(a / cell);
    }
  }
  return k;
}

static inline uint8_t sat_u8_rint(float v) {
  long r = lrintf(v);
  return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

void disref_resize_area_u8(const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw) {
  if (sh == dh && sw == dw) { memcpy(dst, src, (size_t)sh * sw); return; }
  double scale_x = 1. / ((double)dw / sw), scale_y = 1. / ((double)dh / sh);
  int ix = (int)lrint(scale_x), iy = (int)lrint(scale_y);
  if (fabs(scale_x - ix) < 2.220446049250313e-16 && fabs(scale_y - iy) < 2.220446049250313e-16) {
    float sc = 1.f / (float)(ix * iy);
    for (int y = 0; y < dh; y++)
      for (int x = 0; x < dw; x++) {
        int sum = 0;
        for (int j = 0; j < iy; j++)
          for (int i = 0; i < ix; i++) sum += src[(size_t)(y * iy + j) * sw + x * ix + i];
        dst[(size_t)y * dw + x] = (ix == 2 && iy == 2) ? (uint8_t)((sum + 2) >> 2) : sat_u8_rint((float)sum * sc);
      }
    return;
  }
  area_ent* xt = (area_ent*)malloc(sizeof(area_ent) * (sw * 2 + 4));
  area_ent* yt = (area_ent*)malloc(sizeof(area_ent) * (sh * 2 + 4));
  int nx = area_tab(sw, dw, scale_x, xt), ny = area_tab(sh, dh, scale_y, yt);
  float* buf = (float*)malloc(sizeof(float) * dw);
  float* sum = (float*)calloc(dw, sizeof(float));
  int prev_dy = yt[0].di;
  for (int j = 0; j < ny; j++) {
    float beta = yt[j].alpha;
    int dy = yt[j].di, sy = yt[j].si;
    const uint8_t* S = src + (size_t)sy * sw;
    for (int x = 0; x < dw; x++) buf[x] = 0.f;
    for (int k = 0; k < nx; k++) buf[xt[k].di] += S[xt[k].si] * xt[k].alpha;
    if (dy != prev_dy) {
      for (int x = 0; x < dw; x++) { dst[(size_t)prev_dy * dw + x] = sat_u8_rint(sum[x]); sum[x] = beta * buf[x]; }
      prev_dy = dy;
    } else {
      for (int x = 0; x < dw; x++) sum[x] += beta * buf[x];
    }
  }
  for (int x = 0; x < dw; x++) dst[(size_t)prev_dy * dw + x] = sat_u8_rint(sum[x]);
  free(xt); free(yt); free(buf); free(sum);
}

/* --------------------------------------------------------------- bilinear resize (f32) ---- */

/* cv::resize(src, dst, INTER_LINEAR) for float32, `cn` interleaved channels, as the wheel runs it:
 *   mode 0  OpenCV's own code (used for the 2-channel final flow, IPP has no C2 variant):
 *           f = (float)((d+.5)*scale-.5); s = floor(f); f -= s;  out = S0*(1-f) + S1*f, h then v
 *   mode 2  IPP path taken for 1-channel float images (the per-level Ux / Uy upsampling):
 *           fraction computed in double then cast to float, out = fma(S1 - S0, f, S0), h then v
 * Both were identified black-box against cv2 4.13.0.92 and are bit-exact (tests/test_oracle_dis.py). */
void disref_resize_linear_f32(const float* src, int sh, int sw, int cn, float* dst, int dh, int dw, int mode) {
  double scale_x = mode == 2 ? (double)sw / dw : 1. / ((double)dw / sw);
  double scale_y = mode == 2 ? (double)sh / dh : 1. / ((double)dh / sh);
  int* xofs = (int*)malloc(sizeof(int) * dw);
  float* xa = (float*)malloc(sizeof(float) * dw);
  for (int dx = 0; dx < dw; dx++) {
    int sx; float fx;
    if (mode == 2) {
      double f = (dx + 0.5) * scale_x - 0.5;
      sx = (int)floor(f);
      f -= sx;
      if (sx < 0) { f = 0; sx = 0; }
      if (sx >= sw - 1) { f = 0; sx = sw - 1; }
      fx = (float)f;
    } else {
      fx = (float)((dx + 0.5) * scale_x - 0.5);
      sx = (int)floorf(fx);
      fx -= sx;
      if (sx < 0) { fx = 0; sx = 0; }
      if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    }
    xofs[dx] = sx;
    xa[dx] = fx;
  }
  float* rows[2];
  rows[0] = (float*)malloc(sizeof(float) * dw * cn);
  rows[1] = (float*)malloc(sizeof(float) * dw * cn);
  for (int dy = 0; dy < dh; dy++) {
    int sy; float fy;
    if (mode == 2) {
      double f = (dy + 0.5) * scale_y - 0.5;
      sy = (int)floor(f);
      f -= sy;
      if (sy < 0) { f = 0; sy = 0; }
      if (sy >= sh - 1) { f = 0; sy = sh - 1; }
      fy = (float)f;
    } else {
      fy = (float)((dy + 0.5) * scale_y - 0.5);
      sy = (int)floorf(fy);
      fy -= sy;
    }
    for (int k = 0; k < 2; k++) {
      int yy = sy + k;
      if (yy < 0) yy = 0;
      if (yy > sh - 1) yy = sh - 1;
      const float* S = src + (size_t)yy * sw * cn;
      for (int dx = 0; dx < dw; dx++)
        for (int c = 0; c < cn; c++) {
          int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
          float p0 = S[sx * cn + c], p1 = S[sx1 * cn + c], f = xa[dx];
          rows[k][dx * cn + c] = mode == 2 ? fmaf(p1 - p0, f, p0) : p0 * (1.f - f) + p1 * f;
        }
    }
    float* D = dst + (size_t)dy * dw * cn;
    for (int x = 0; x < dw * cn; x++)
      D[x] = mode == 2 ? fmaf(rows[1][x] - rows[0][x], fy, rows[0][x]) : rows[0][x] * (1.f - fy) + rows[1][x] * fy;
  }
  free(xofs); free(xa); free(rows[0]); free(rows[1]);
}

/* ------------------------------------------------------------------------- level data ---- */

typedef struct {
  int w, h;
  uint8_t *I0, *I1, *I1ext; /* I1ext: (h+32) x (w+32) */
  int16_t *I0x, *I0y;
  float *Ux, *Uy;
} dis_level;

static inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
  return p;
}

/* cv::spatialGradient, ksize 3, BORDER_DEFAULT (reflect-101): Sobel in int16. */
static void spatial_gradient(const uint8_t* I, int h, int w, int16_t* gx, int16_t* gy) {
  for (int y = 0; y < h; y++) {
    int ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
    for (int x = 0; x < w; x++) {
      int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
      int a = I[ym * w + xm], b = I[ym * w + x], c = I[ym * w + xp];
      int d = I[y * w + xm], f = I[y * w + xp];
      int g = I[yp * w + xm], hh = I[yp * w + x], i = I[yp * w + xp];
      gx[y * w + x] = (int16_t)((c + 2 * f + i) - (a + 2 * d + g));
      gy[y * w + x] = (int16_t)((g + 2 * hh + i) - (a + 2 * b + c));
    }
  }
}

static void make_border(const uint8_t* I, int h, int w, uint8_t* E) {
  int we = w + 2 * BORDER;
  for (int y = 0; y < h + 2 * BORDER; y++) {
    int sy = y - BORDER; if (sy < 0) sy = 0; if (sy > h - 1) sy = h - 1;
    for (int x = 0; x < we; x++) {
      int sx = x - BORDER; if (sx < 0) sx = 0; if (sx > w - 1) sx = w - 1;
      E[y * we + x] = I[sy * w + sx];
    }
  }
}

/* ------------------------------------------------------------- structure tensor sums ---- */

static void structure_tensor(const dis_level* L, int psz, int pstr, int ws, int hs, float* xx, float* yy, float* xy,
                             float* sx_, float* sy_) {
  int w = L->w, h = L->h;
  float* axx = (float*)malloc(sizeof(float) * h * ws * 5);
  float *ayy = axx + h * ws, *axy = ayy + h * ws, *ax = axy + h * ws, *ay = ax + h * ws;
  for (int i = 0; i < h; i++) {
    float s_xx = 0, s_yy = 0, s_xy = 0, s_x = 0, s_y = 0;
    const int16_t* xr = L->I0x + i * w; const int16_t* yr = L->I0y + i * w;
    for (int j = 0; j < psz; j++) {
      s_xx += xr[j] * xr[j]; s_yy += yr[j] * yr[j]; s_xy += xr[j] * yr[j]; s_x += xr[j]; s_y += yr[j];
    }
    axx[i * ws] = s_xx; ayy[i * ws] = s_yy; axy[i * ws] = s_xy; ax[i * ws] = s_x; ay[i * ws] = s_y;
    int js = 1;
    for (int j = psz; j < w; j++) {
      s_xx += (xr[j] * xr[j] - xr[j - psz] * xr[j - psz]);
      s_yy += (yr[j] * yr[j] - yr[j - psz] * yr[j - psz]);
      s_xy += (xr[j] * yr[j] - xr[j - psz] * yr[j - psz]);
      s_x += (xr[j] - xr[j - psz]);
      s_y += (yr[j] - yr[j - psz]);
      if ((j - psz + 1) % pstr == 0) {
        axx[i * ws + js] = s_xx; ayy[i * ws + js] = s_yy; axy[i * ws + js] = s_xy; ax[i * ws + js] = s_x; ay[i * ws + js] = s_y;
        js++;
      }
    }
  }
  float* c = (float*)calloc(ws * 5, sizeof(float));
  float *cxx = c, *cyy = c + ws, *cxy = c + 2 * ws, *cx = c + 3 * ws, *cy = c + 4 * ws;
  for (int i = 0; i < psz; i++)
    for (int j = 0; j < ws; j++) {
      cxx[j] += axx[i * ws + j]; cyy[j] += ayy[i * ws + j]; cxy[j] += axy[i * ws + j]; cx[j] += ax[i * ws + j]; cy[j] += ay[i * ws + j];
    }
  for (int j = 0; j < ws; j++) { xx[j] = cxx[j]; yy[j] = cyy[j]; xy[j] = cxy[j]; sx_[j] = cx[j]; sy_[j] = cy[j]; }
  int is = 1;
  for (int i = psz; i < h; i++) {
    for (int j = 0; j < ws; j++) {
      cxx[j] += (axx[i * ws + j] - axx[(i - psz) * ws + j]);
      cyy[j] += (ayy[i * ws + j] - ayy[(i - psz) * ws + j]);
      cxy[j] += (axy[i * ws + j] - axy[(i - psz) * ws + j]);
      cx[j] += (ax[i * ws + j] - ax[(i - psz) * ws + j]);
      cy[j] += (ay[i * ws + j] - ay[(i - psz) * ws + j]);
    }
    if ((i - psz + 1) % pstr == 0) {
      for (int j = 0; j < ws; j++) {
        xx[is * ws + j] = cxx[j]; yy[is * ws + j] = cyy[j]; xy[is * ws + j] = cxy[j]; sx_[is * ws + j] = cx[j]; sy_[is * ws + j] = cy[j];
      }
      is++;
    }
  }
  free(axx); free(c);
}

/* ----------------------------------------------------------------- patch functions ---- */
/* 8x8 patch, OpenCV's v_float32x4 order: lane l accumulates pixels l and l+4 of each row,
 * v_reduce_sum = (a0 + a2) + (a1 + a3). */

static inline float reduce4(const float* a) { return (a[0] + a[2]) + (a[1] + a[3]); }

static float patch_ssd_meannorm(const uint8_t* I0, const uint8_t* I1, int s0, int s1, float w00, float w01, float w10, float w11) {
  float sd[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
  for (int r = 0; r < 8; r++) {
    const uint8_t* a = I1 + r * s1; const uint8_t* b = a + s1; const uint8_t* z = I0 + r * s0;
    for (int l = 0; l < 4; l++) {
      float dl = w00 * (float)a[l] + w01 * (float)a[l + 1] + w10 * (float)b[l] + w11 * (float)b[l + 1] - (float)z[l];
      float dr = w00 * (float)a[l + 4] + w01 * (float)a[l + 5] + w10 * (float)b[l + 4] + w11 * (float)b[l + 5] - (float)z[l + 4];
      sq[l] = sq[l] + (dl * dl + dr * dr);
      sd[l] = sd[l] + (dl + dr);
    }
  }
  float sum_diff = reduce4(sd), sum_sq = reduce4(sq);
  return sum_sq - sum_diff * sum_diff / 64.f;
}

static float patch_ssd(const uint8_t* I0, const uint8_t* I1, int s0, int s1, float w00, float w01, float w10, float w11) {
  float sq[4] = {0, 0, 0, 0};
  for (int r = 0; r < 8; r++) {
    const uint8_t* a = I1 + r * s1; const uint8_t* b = a + s1; const uint8_t* z = I0 + r * s0;
    for (int l = 0; l < 4; l++) {
      float dl = w00 * (float)a[l] + w01 * (float)a[l + 1] + w10 * (float)b[l] + w11 * (float)b[l + 1] - (float)z[l];
      float dr = w00 * (float)a[l + 4] + w01 * (float)a[l + 5] + w10 * (float)b[l + 4] + w11 * (float)b[l + 5] - (float)z[l + 4];
      sq[l] = sq[l] + (dl * dl + dr * dr);
    }
  }
  return reduce4(sq);
}

static float patch_process_meannorm(float* dUx, float* dUy, const uint8_t* I0, const uint8_t* I1, const int16_t* gx,
                                    const int16_t* gy, int s0, int s1, float w00, float w01, float w10, float w11,
                                    float xgs, float ygs) {
  float sd[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0}, mx[4] = {0, 0, 0, 0}, my[4] = {0, 0, 0, 0};
  for (int r = 0; r < 8; r++) {
    const uint8_t* a = I1 + r * s1; const uint8_t* b = a + s1; const uint8_t* z = I0 + r * s0;
    const int16_t* px = gx + r * s0; const int16_t* py = gy + r * s0;
    for (int l = 0; l < 4; l++) {
      float dl = w00 * (float)a[l] + w01 * (float)a[l + 1] + w10 * (float)b[l] + w11 * (float)b[l + 1] - (float)z[l];
      float dr = w00 * (float)a[l + 4] + w01 * (float)a[l + 5] + w10 * (float)b[l + 4] + w11 * (float)b[l + 5] - (float)z[l + 4];
      mx[l] = mx[l] + (dl * (float)px[l] + dr * (float)px[l + 4]);
      my[l] = my[l] + (dl * (float)py[l] + dr * (float)py[l + 4]);
      sq[l] = sq[l] + (dl * dl + dr * dr);
      sd[l] = sd[l] + (dl + dr);
    }
  }
  float sum_diff = reduce4(sd), sum_sq = reduce4(sq), smx = reduce4(mx), smy = reduce4(my);
  *dUx = smx - sum_diff * xgs / 64.f;
  *dUy = smy - sum_diff * ygs / 64.f;
  return sum_sq - sum_diff * sum_diff / 64.f;
}

static float patch_process(float* dUx, float* dUy, const uint8_t* I0, const uint8_t* I1, const int16_t* gx,
                           const int16_t* gy, int s0, int s1, float w00, float w01, float w10, float w11) {
  float sq[4] = {0, 0, 0, 0}, mx[4] = {0, 0, 0, 0}, my[4] = {0, 0, 0, 0};
  for (int r = 0; r < 8; r++) {
    const uint8_t* a = I1 + r * s1; const uint8_t* b = a + s1; const uint8_t* z = I0 + r * s0;
    const int16_t* px = gx + r * s0; const int16_t* py = gy + r * s0;
    for (int l = 0; l < 4; l++) {
      float dl = w00 * (float)a[l] + w01 * (float)a[l + 1] + w10 * (float)b[l] + w11 * (float)b[l + 1] - (float)z[l];
      float dr = w00 * (float)a[l + 4] + w01 * (float)a[l + 5] + w10 * (float)b[l + 4] + w11 * (float)b[l + 5] - (float)z[l + 4];
      mx[l] = mx[l] + (dl * (float)px[l] + dr * (float)px[l + 4]);
      my[l] = my[l] + (dl * (float)py[l] + dr * (float)py[l + 4]);
      sq[l] = sq[l] + (dl * dl + dr * dr);
    }
  }
  *dUx = reduce4(mx);
  *dUy = reduce4(my);
  return reduce4(sq);
}

/* ------------------------------------------------------------- patch inverse search ---- */

typedef struct { float w00, w01, w10, w11; int iy, ix; } bil;

static inline bil bil_weights(float i, float j, float Ux, float Uy, int w, int h, int psz) {
  float i_lo = BORDER - psz + 1.0f, i_hi = BORDER + h - 1.0f;
  float j_lo = BORDER - psz + 1.0f, j_hi = BORDER + w - 1.0f;
  float iI = fminf(fmaxf(i + Uy + BORDER, i_lo), i_hi);
  float jI = fminf(fmaxf(j + Ux + BORDER, j_lo), j_hi);
  float di = iI - floorf(iI), dj = jI - floorf(jI);
  bil b;
  b.w11 = di * dj; b.w10 = di * (1 - dj); b.w01 = (1 - di) * dj; b.w00 = (1 - di) * (1 - dj);
  b.iy = (int)iI; b.ix = (int)jI;
  return b;
}

static void patch_inverse_search(const dis_level* L, const dis_params* P, int ws, int hs, float* Sx, float* Sy,
                                 const float* xx, const float* yy, const float* xy, const float* xs, const float* ys) {
  const int psz = P->patch_size, pstr = P->patch_stride, psz2 = psz / 2;
  const int w = L->w, h = L->h, we = w + 2 * BORDER;
  const int nstripes = 8; /* fixed when spatial propagation is on */
  const int stripe_sz = (int)ceil(hs / (double)nstripes);
  const int num_iter = P->use_spatial ? 2 : 1;
  const int inner = (int)floor(P->gd_iter / (float)num_iter);
  for (int n = 0; n < nstripes; n++) {
    for (int iter = 0; iter < num_iter; iter++) {
      int dir, s_is, e_is, s_js, e_js, s_i, s_j;
      int lo = n * stripe_sz < hs ? n * stripe_sz : hs, hi = (n + 1) * stripe_sz < hs ? (n + 1) * stripe_sz : hs;
      if (iter % 2 == 0) { dir = 1; s_is = lo; e_is = hi; s_js = 0; e_js = ws; s_i = s_is * pstr; s_j = 0; }
      else { dir = -1; s_is = hi - 1; e_is = lo - 1; s_js = ws - 1; e_js = -1; s_i = s_is * pstr; s_j = (ws - 1) * pstr; }
      int i = s_i;
      for (int is = s_is; dir * is < dir * e_is; is += dir) {
        int j = s_j;
        for (int js = s_js; dir * js < dir * e_js; js += dir) {
          const int k = is * ws + js;
          if (iter == 0) { Sx[k] = L->Ux[(i + psz2) * w + j + psz2]; Sy[k] = L->Uy[(i + psz2) * w + j + psz2]; }
          const uint8_t* I0p = L->I0 + i * w + j;
          float min_ssd = DIS_INF, cur_ssd;
#define SSD_AT(dst, ux, uy)                                                                         \
  do {                                                                                              \
    bil b_ = bil_weights((float)i, (float)j, (ux), (uy), w, h, psz);                                \
    const uint8_t* I1p = L->I1ext + b_.iy * we + b_.ix;                                             \
    dst = P->use_mean_norm ? patch_ssd_meannorm(I0p, I1p, w, we, b_.w00, b_.w01, b_.w10, b_.w11)    \
                           : patch_ssd(I0p, I1p, w, we, b_.w00, b_.w01, b_.w10, b_.w11);            \
  } while (0)
          if (P->use_spatial) {
            SSD_AT(min_ssd, Sx[k], Sy[k]);
            if (dir * js > dir * s_js) {
              SSD_AT(cur_ssd, Sx[k - dir], Sy[k - dir]);
              if (cur_ssd < min_ssd) { min_ssd = cur_ssd; Sx[k] = Sx[k - dir]; Sy[k] = Sy[k - dir]; }
            }
            if (dir * is > dir * s_is) {
              SSD_AT(cur_ssd, Sx[k - dir * ws], Sy[k - dir * ws]);
              if (cur_ssd < min_ssd) { min_ssd = cur_ssd; Sx[k] = Sx[k - dir * ws]; Sy[k] = Sy[k - dir * ws]; }
            }
          }
          float cur_Ux = Sx[k], cur_Uy = Sy[k];
          float detH = xx[k] * yy[k] - xy[k] * xy[k];
          if (fabsf(detH) < DIS_EPS) detH = DIS_EPS;
          float invH11 = yy[k] / detH, invH12 = -xy[k] / detH, invH22 = xx[k] / detH;
          float prev_ssd = DIS_INF, ssd;
          float xgs = xs[k], ygs = ys[k];
          for (int t = 0; t < inner; t++) {
            float dUx, dUy;
            bil b = bil_weights((float)i, (float)j, cur_Ux, cur_Uy, w, h, psz);
            const uint8_t* I1p = L->I1ext + b.iy * we + b.ix;
            if (P->use_mean_norm)
              ssd = patch_process_meannorm(&dUx, &dUy, I0p, I1p, L->I0x + i * w + j, L->I0y + i * w + j, w, we, b.w00,
                                           b.w01, b.w10, b.w11, xgs, ygs);
            else
              ssd = patch_process(&dUx, &dUy, I0p, I1p, L->I0x + i * w + j, L->I0y + i * w + j, w, we, b.w00, b.w01,
                                  b.w10, b.w11);
            float dx = invH11 * dUx + invH12 * dUy;
            float dy = invH12 * dUx + invH22 * dUy;
            cur_Ux -= dx; cur_Uy -= dy;
            if (ssd >= prev_ssd) break;
            prev_ssd = ssd;
          }
          {
            float ex = cur_Ux - Sx[k], ey = cur_Uy - Sy[k];
            /* cv::norm(Vec2f) accumulates in double */
            double nrm = sqrt((double)ex * ex + (double)ey * ey);
            if (nrm <= psz) { Sx[k] = cur_Ux; Sy[k] = cur_Uy; }
          }
          j += dir * pstr;
        }
        i += dir * pstr;
      }
    }
  }
#undef SSD_AT
}

/* -------------------------------------------------------------------- densification ---- */

static void densify(const dis_level* L, const dis_params* P, int ws, int hs, const float* Sx, const float* Sy) {
  const int psz = P->patch_size, pstr = P->patch_stride, w = L->w, h = L->h;
  int s_is = 0, e_is = -1;
  (void)hs;
  for (int i = 0; i < h; i++) {
    if (i % pstr == 0 && i + psz <= h) e_is++;
    if (i - psz >= 0 && (i - psz) % pstr == 0 && s_is < e_is) s_is++;
    int s_js = 0, e_js = -1;
    for (int j = 0; j < w; j++) {
      if (j % pstr == 0 && j + psz <= w) e_js++;
      if (j - psz >= 0 && (j - psz) % pstr == 0 && s_js < e_js) s_js++;
      float sum_coef = 0.f, sum_Ux = 0.f, sum_Uy = 0.f;
      for (int is = s_is; is <= e_is; is++)
        for (int js = s_js; js <= e_js; js++) {
          float sx = Sx[is * ws + js], sy = Sy[is * ws + js];
          float j_m = fminf(fmaxf(j + sx, 0.0f), w - 1.0f - DIS_EPS);
          float i_m = fminf(fmaxf(i + sy, 0.0f), h - 1.0f - DIS_EPS);
          int j_l = (int)j_m, j_u = j_l + 1, i_l = (int)i_m, i_u = i_l + 1;
          float diff = (j_m - j_l) * (i_m - i_l) * L->I1[i_u * w + j_u] + (j_u - j_m) * (i_m - i_l) * L->I1[i_u * w + j_l] +
                       (j_m - j_l) * (i_u - i_m) * L->I1[i_l * w + j_u] + (j_u - j_m) * (i_u - i_m) * L->I1[i_l * w + j_l] -
                       L->I0[i * w + j];
          float coef = 1 / fmaxf(1.0f, fabsf(diff));
          sum_Ux += coef * sx; sum_Uy += coef * sy; sum_coef += coef;
        }
      L->Ux[i * w + j] = sum_Ux / sum_coef;
      L->Uy[i * w + j] = sum_Uy / sum_coef;
    }
  }
}

/* ------------------------------------------------------------ variational refinement ---- */

/* cv::remap(src f32, mapX, mapY, INTER_LINEAR, BORDER_REPLICATE): 1/32-px fixed point. */
static void remap_linear_replicate(const float* src, int h, int w, const float* ux, const float* uy, float* dst) {
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      float mx = x + ux[y * w + x], my = y + uy[y * w + x];
      int ix = (int)lrintf(mx * 32.f), iy = (int)lrintf(my * 32.f);
      int sx = ix >> 5, sy = iy >> 5, ax = ix & 31, ay = iy & 31;
      if (sx < -32768) sx = -32768; if (sx > 32767) sx = 32767;
      if (sy < -32768) sy = -32768; if (sy > 32767) sy = 32767;
      float fx1 = ax * (1.f / 32.f), fy1 = ay * (1.f / 32.f), fx0 = 1.f - fx1, fy0 = 1.f - fy1;
      float w00 = fy0 * fx0, w01 = fy0 * fx1, w10 = fy1 * fx0, w11 = fy1 * fx1;
      int x0 = sx < 0 ? 0 : (sx > w - 1 ? w - 1 : sx), x1 = sx + 1 < 0 ? 0 : (sx + 1 > w - 1 ? w - 1 : sx + 1);
      int y0 = sy < 0 ? 0 : (sy > h - 1 ? h - 1 : sy), y1 = sy + 1 < 0 ? 0 : (sy + 1 > h - 1 ? h - 1 : sy + 1);
      dst[y * w + x] = src[y0 * w + x0] * w00 + src[y0 * w + x1] * w01 + src[y1 * w + x0] * w10 + src[y1 * w + x1] * w11;
    }
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* central differences without the 1/2 factor (cv::Sobel ksize 1), replicate border */
static void ddx(const float* s, int h, int w, float* d) {
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) d[y * w + x] = s[y * w + clampi(x + 1, 0, w - 1)] - s[y * w + clampi(x - 1, 0, w - 1)];
}
static void ddy(const float* s, int h, int w, float* d) {
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) d[y * w + x] = s[clampi(y + 1, 0, h - 1) * w + x] - s[clampi(y - 1, 0, h - 1) * w + x];
}

/* VariationalRefinement::calcUV restated on plain row-major buffers.  The red/black buffers of
 * OpenCV are a storage layout; what matters is (a) the order in which the smoothness terms are
 * accumulated into A11/A22/b1/b2, (b) red (x+y even) before black in every sweep, (c) dW = 0
 * outside the image, (d) the rightmost column / bottom row add no forward smoothness term. */
void disref_variational_refinement(const float* I0f, const float* I1f, int h, int w, float* W_u, float* W_v,
                                   float alpha, float delta, float gamma, float epsilon, int fp_iter, int sor_iter,
                                   float omega) {
  const size_t n = (size_t)h * w;
  float* buf = (float*)calloc(n * 22, sizeof(float));
  float *warped = buf, *avg = buf + n, *Iz = buf + 2 * n, *Ix = buf + 3 * n, *Iy = buf + 4 * n, *Ixx = buf + 5 * n,
        *Ixy = buf + 6 * n, *Iyy = buf + 7 * n, *Ixz = buf + 8 * n, *Iyz = buf + 9 * n, *A11 = buf + 10 * n,
        *A12 = buf + 11 * n, *A22 = buf + 12 * n, *b1 = buf + 13 * n, *b2 = buf + 14 * n, *wgt = buf + 15 * n,
        *tu = buf + 16 * n, *tv = buf + 17 * n, *du = buf + 18 * n, *dv = buf + 19 * n, *u0 = buf + 20 * n, *v0 = buf + 21 * n;
  remap_linear_replicate(I1f, h, w, W_u, W_v, warped);
  for (size_t k = 0; k < n; k++) { avg[k] = 0.5f * (I0f[k] + warped[k]); Iz[k] = warped[k] - I0f[k]; }
  ddx(avg, h, w, Ix); ddy(avg, h, w, Iy);
  ddx(Ix, h, w, Ixx); ddy(Ix, h, w, Ixy); ddy(Iy, h, w, Iyy);
  ddx(Iz, h, w, Ixz); ddy(Iz, h, w, Iyz);
  memcpy(u0, W_u, n * sizeof(float)); memcpy(v0, W_v, n * sizeof(float));
  memcpy(tu, W_u, n * sizeof(float)); memcpy(tv, W_v, n * sizeof(float));
  const float zeta2 = 0.1f * 0.1f, eps2 = epsilon * epsilon, gamma2 = gamma / 2, delta2 = delta / 2, alpha2 = alpha / 2;
  for (int it = 0; it < fp_iter; it++) {
    /* data term */
    for (size_t k = 0; k < n; k++) {
      float ix = Ix[k], iy = Iy[k], iz = Iz[k], ixx = Ixx[k], ixy = Ixy[k], iyy = Iyy[k], ixz = Ixz[k], iyz = Iyz[k];
      float dU = du[k], dV = dv[k];
      float derivNorm = ix * ix + iy * iy + zeta2;
      float Ik1z = iz + ix * dU + iy * dV;
      float weight = (delta2 / sqrtf(Ik1z * Ik1z / derivNorm + eps2)) / derivNorm;
      float a11 = weight * (ix * ix) + zeta2;
      float a12 = weight * (ix * iy);
      float a22 = weight * (iy * iy) + zeta2;
      float bb1 = -weight * (iz * ix);
      float bb2 = -weight * (iz * iy);
      derivNorm = ixx * ixx + ixy * ixy + zeta2;
      float derivNorm2 = iyy * iyy + ixy * ixy + zeta2;
      float Ik1zx = ixz + ixx * dU + ixy * dV;
      float Ik1zy = iyz + ixy * dU + iyy * dV;
      weight = gamma2 / sqrtf(Ik1zx * Ik1zx / derivNorm + Ik1zy * Ik1zy / derivNorm2 + eps2);
      a11 += weight * (ixx * ixx / derivNorm + ixy * ixy / derivNorm2);
      a12 += weight * (ixx * ixy / derivNorm + ixy * iyy / derivNorm2);
      a22 += weight * (ixy * ixy / derivNorm + iyy * iyy / derivNorm2);
      bb1 += -weight * (ixx * ixz / derivNorm + ixy * iyz / derivNorm2);
      bb2 += -weight * (ixy * ixz / derivNorm + iyy * iyz / derivNorm2);
      A11[k] = a11; A12[k] = a12; A22[k] = a22; b1[k] = bb1; b2[k] = bb2;
    }
    /* smoothness weights (forward differences of the current flow tu/tv, replicate border) */
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        size_t k = (size_t)y * w + x;
        size_t kr = (size_t)y * w + (x + 1 < w ? x + 1 : x), kd = (size_t)(y + 1 < h ? y + 1 : y) * w + x;
        float ux = tu[kr] - tu[k], vx = tv[kr] - tv[k], uy = tu[kd] - tu[k], vy = tv[kd] - tv[k];
        wgt[k] = alpha2 / sqrtf(ux * ux + vx * vx + uy * uy + vy * vy + eps2);
      }
    /* horizontal pass: red pixels first, then black; each adds to itself and to its right neighbour */
    for (int pass = 0; pass < 2; pass++)
      for (int y = 0; y < h; y++)
        for (int x = (y + pass) & 1; x < w - 1; x += 2) {
          size_t k = (size_t)y * w + x;
          float wv = wgt[k];
          float ux = wv * (u0[k + 1] - u0[k]), vx = wv * (v0[k + 1] - v0[k]);
          b1[k] += ux; A11[k] += wv; b2[k] += vx; A22[k] += wv;
          b1[k + 1] -= ux; A11[k + 1] += wv; b2[k + 1] -= vx; A22[k + 1] += wv;
        }
    /* vertical pass over rows 0..h-2 (the last row has no row below and is skipped) */
    for (int pass = 0; pass < 2; pass++)
      for (int y = 0; y < h - 1; y++)
        for (int x = (y + pass) & 1; x < w; x += 2) {
          size_t k = (size_t)y * w + x;
          float wv = wgt[k];
          float uy = wv * (u0[k + w] - u0[k]), vy = wv * (v0[k + w] - v0[k]);
          b1[k] += uy; A11[k] += wv; b2[k] += vy; A22[k] += wv;
          b1[k + w] -= uy; A11[k + w] += wv; b2[k + w] -= vy; A22[k + w] += wv;
        }
    /* red-black SOR on dW (zero outside the image) */
    for (int s = 0; s < sor_iter; s++)
      for (int pass = 0; pass < 2; pass++)
        for (int y = 0; y < h; y++)
          for (int x = (y + pass) & 1; x < w; x += 2) {
            size_t k = (size_t)y * w + x;
            float wl = x > 0 ? wgt[k - 1] : 0.f, dul = x > 0 ? du[k - 1] : 0.f, dvl = x > 0 ? dv[k - 1] : 0.f;
            float dur = x + 1 < w ? du[k + 1] : 0.f, dvr = x + 1 < w ? dv[k + 1] : 0.f;
            float wu = y > 0 ? wgt[k - w] : 0.f, duu = y > 0 ? du[k - w] : 0.f, dvu = y > 0 ? dv[k - w] : 0.f;
            float dud = y + 1 < h ? du[k + w] : 0.f, dvd = y + 1 < h ? dv[k + w] : 0.f;
            float wc = wgt[k];
            float sigmaU = wl * dul + wc * dur + wu * duu + wc * dud;
            float sigmaV = wl * dvl + wc * dvr + wu * dvu + wc * dvd;
            du[k] += omega * ((sigmaU + b1[k] - dv[k] * A12[k]) / A11[k] - du[k]);
            dv[k] += omega * ((sigmaV + b2[k] - du[k] * A12[k]) / A22[k] - dv[k]);
          }
    for (size_t k = 0; k < n; k++) { tu[k] = u0[k] + du[k]; tv[k] = v0[k] + dv[k]; }
  }
  memcpy(W_u, tu, n * sizeof(float)); memcpy(W_v, tv, n * sizeof(float));
  free(buf);
}

/* ---------------------------------------------------------------------------- driver ---- */

int disref_coarsest_scale(int h, int w, int psz) {
  int mx = w > h ? w : h, mn = w < h ? w : h;
  int a = (int)(log(mx / (4.0 * psz)) / log(2.0) + 0.5);
  int b = mn / psz > 0 ? (int)(log((double)(mn / psz)) / log(2.0)) : -1; /* log(0): cv2 ends up below zero too */
  return a < b ? a : b;
}

/* Runs stages 2-4 (and 5 if vr_iter > 0) on one level whose I0/I1/Ux/Uy are filled in. */
static void run_level(dis_level* L, const dis_params* P) {
  const int psz = P->patch_size, pstr = P->patch_stride;
  const int ws = 1 + (L->w - psz) / pstr, hs = 1 + (L->h - psz) / pstr;
  float* t = (float*)malloc(sizeof(float) * ws * hs * 7);
  float *xx = t, *yy = t + ws * hs, *xy = t + 2 * ws * hs, *xs = t + 3 * ws * hs, *ys = t + 4 * ws * hs,
        *Sx = t + 5 * ws * hs, *Sy = t + 6 * ws * hs;
  structure_tensor(L, psz, pstr, ws, hs, xx, yy, xy, xs, ys);
  patch_inverse_search(L, P, ws, hs, Sx, Sy, xx, yy, xy, xs, ys);
  densify(L, P, ws, hs, Sx, Sy);
  if (P->vr_iter > 0) {
    size_t n = (size_t)L->w * L->h;
    float* f = (float*)malloc(sizeof(float) * n * 2);
    for (size_t k = 0; k < n; k++) { f[k] = L->I0[k]; f[n + k] = L->I1[k]; }
    disref_variational_refinement(f, f + n, L->h, L->w, L->Ux, L->Uy, P->vr_alpha, P->vr_delta, P->vr_gamma,
                                  P->vr_epsilon, P->vr_iter, P->sor_iter, P->omega);
    free(f);
  }
  free(t);
}

/* flow: [h][w][2] float32.  stage_stop: 0 = full algorithm; k > 0 = return after processing
 * only the coarsest k levels' worth of work is not supported -- use dis_params knobs instead. */
static int calc_scales(const uint8_t* I0, const uint8_t* I1, int h, int w, const dis_params* P, int finest, int coarsest,
                       float* flow) {
  dis_level* Ls = (dis_level*)calloc(coarsest + 1, sizeof(dis_level));
  int fraction = 1;
  for (int i = 0; i <= coarsest; i++) {
    if (i >= finest) {
      dis_level* L = &Ls[i];
      if (i == finest) { L->h = h / fraction; L->w = w / fraction; }
      else { L->h = Ls[i - 1].h / 2; L->w = Ls[i - 1].w / 2; }
      size_t n = (size_t)L->h * L->w;
      L->I0 = (uint8_t*)malloc(n); L->I1 = (uint8_t*)malloc(n);
      L->I1ext = (uint8_t*)malloc((size_t)(L->h + 2 * BORDER) * (L->w + 2 * BORDER));
      L->I0x = (int16_t*)malloc(n * 2); L->I0y = (int16_t*)malloc(n * 2);
      L->Ux = (float*)calloc(n, 4); L->Uy = (float*)calloc(n, 4);
      if (i == finest) {
        disref_resize_area_u8(I0, h, w, L->I0, L->h, L->w);
        disref_resize_area_u8(I1, h, w, L->I1, L->h, L->w);
      } else {
        disref_resize_area_u8(Ls[i - 1].I0, Ls[i - 1].h, Ls[i - 1].w, L->I0, L->h, L->w);
        disref_resize_area_u8(Ls[i - 1].I1, Ls[i - 1].h, Ls[i - 1].w, L->I1, L->h, L->w);
      }
      make_border(L->I1, L->h, L->w, L->I1ext);
      spatial_gradient(L->I0, L->h, L->w, L->I0x, L->I0y);
    }
    fraction *= 2;
  }
  for (int i = coarsest; i >= finest; i--) {
    run_level(&Ls[i], P);
    if (i > finest) {
      disref_resize_linear_f32(Ls[i].Ux, Ls[i].h, Ls[i].w, 1, Ls[i - 1].Ux, Ls[i - 1].h, Ls[i - 1].w, 2);
      disref_resize_linear_f32(Ls[i].Uy, Ls[i].h, Ls[i].w, 1, Ls[i - 1].Uy, Ls[i - 1].h, Ls[i - 1].w, 2);
      size_t n = (size_t)Ls[i - 1].h * Ls[i - 1].w;
      for (size_t k = 0; k < n; k++) { Ls[i - 1].Ux[k] *= 2; Ls[i - 1].Uy[k] *= 2; }
    }
  }
  {
    dis_level* L = &Ls[finest];
    size_t n = (size_t)L->h * L->w;
    float* U = (float*)malloc(sizeof(float) * n * 2);
    for (size_t k = 0; k < n; k++) { U[2 * k] = L->Ux[k]; U[2 * k + 1] = L->Uy[k]; }
    if (L->h == h && L->w == w) memcpy(flow, U, sizeof(float) * n * 2);
    else disref_resize_linear_f32(U, L->h, L->w, 2, flow, h, w, 0);
    const float mul = (float)(1 << finest);
    for (size_t k = 0; k < (size_t)h * w * 2; k++) flow[k] *= mul;
    free(U);
  }
  for (int i = finest; i <= coarsest; i++) {
    free(Ls[i].I0); free(Ls[i].I1); free(Ls[i].I1ext); free(Ls[i].I0x); free(Ls[i].I0y); free(Ls[i].Ux); free(Ls[i].Uy);
  }
  free(Ls);
  return 0;
}

/* Which pyramid levels calc() works on.  dis_flow.cpp computes the coarsest scale from the frame size; when it
 * falls below the configured finest scale (frames under ~91 px on the long or 32 px on the short side with the
 * reference's finest_scale 2) autoSelectPatchSizeAndScales() replaces both: coarsest = max(0, floor(log2(2 w /
 * (5 * patch)))) in float, finest = max(coarsest - 2, 0), patch size 8.  Returns -1 where cv2 raises
 * ("The input image must have either width or height >= 12"). */
int disref_select_scales(int h, int w, const dis_params* P, int* finest, int* coarsest) {
  int c = disref_coarsest_scale(h, w, P->patch_size);
  int f = P->finest_scale;
  if (c < 0) return -1;
  if (c < f) {
    c = (int)floorf(log2f((2.0f * (float)w) / (5.0f * 8.0f)));
    if (c < 0) c = 0;
    f = c - 2 > 0 ? c - 2 : 0;
  }
  *finest = f; *coarsest = c;
  return 0;
}

/* cv2 reads outside its coarsest level (and usually crashes) when that level is smaller than one patch. */
static int levels_fit(int h, int w, int psz, int finest, int coarsest) {
  int lh = h >> finest, lw = w >> finest;
  for (int i = finest; i < coarsest; i++) { lh /= 2; lw /= 2; }
  return lh >= psz && lw >= psz;
}

/* Explicit levels (probing aid for the selection rule above). */
int disref_calc_scales(const uint8_t* I0, const uint8_t* I1, int h, int w, const dis_params* P, int finest, int coarsest,
                       float* flow) {
  if (finest < 0 || coarsest < finest) return -1;
  if (!levels_fit(h, w, P->patch_size, finest, coarsest)) return -2;
  return calc_scales(I0, I1, h, w, P, finest, coarsest, flow);
}

int disref_calc(const uint8_t* I0, const uint8_t* I1, int h, int w, const dis_params* P, float* flow) {
  int finest, coarsest;
  if (disref_select_scales(h, w, P, &finest, &coarsest)) return -1;
  if (!levels_fit(h, w, P->patch_size, finest, coarsest)) return -2;
  return calc_scales(I0, I1, h, w, P, finest, coarsest, flow);
}
