// classic.cu -- K5 + K6: Shi-Tomasi corners and pyramidal Lucas-Kanade, batched over a clip.
//
// Replaces nodes/video_stabilizer_classic.py:76-83 (cv2.goodFeaturesToTrack: 400 corners, quality
// 0.01, minDistance 7, blockSize 21) and :88-100 (cv2.calcOpticalFlowPyrLK: 31x31 window, 3 pyramid
// levels, 50 iterations / eps 0.01, status filter) for every pair (i, i+1) of the clip.
//
// GFTT arithmetic follows cv2 4.13 bit for bit (see oracle/classic_ref.c): fused Sobel forms,
// unnormalised 21x21 box filter with DOUBLE running sums (one thread per row, then one per
// column -- the running sums are sequential by construction, the parallelism is rows x frames),
// min-eigenvalue, threshold at 0.01 * max, 3x3 local maxima, 64-bit (value, address) keys sorted by
// a per-frame bitonic network, and the greedy minimum-distance pass as one warp per frame (the 9
// neighbour cells of the 7-px grid are probed by 9 lanes).  LK is one warp per (pair, feature):
// fixed-point bilinear windows exactly like OpenCV (14-bit weights, x32 intensities), 2x2 normal
// equations reduced with warp shuffles.  Compiled with -fmad=false; fmaf() only where cv2 fuses.
#include "common.cuh"

#include <stdlib.h>

#include <float.h>
#include <math.h>

namespace {

constexpr int kBlock = 21;
constexpr int kRadius = kBlock / 2;
constexpr int kWin = 31;
constexpr int kMaxLevel = 3;
constexpr int kMaxIter = 50;
constexpr int kCellCap = 4;
constexpr int kCell = 7;
constexpr float kMinDist2 = 49.0f;

__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
  return p;
}

// ---- GFTT ---------------------------------------------------------------------------------------

// Covariance of the Sobel derivatives, written TRANSPOSED: covT[f][x][y][3].  The horizontal box-filter
// pass below is a chain along x per row, so its parallelism is across rows; with x-major storage the 32 rows
// of a warp sit next to each other at every step of the chain.  (A 32x32 tile goes through shared memory:
// image reads coalesced along x, covariance writes coalesced along y.)
__global__ void __launch_bounds__(256) gftt_cov_kernel(const unsigned char* __restrict__ gray, int h, int w, float s, float s2,
                                                       float* __restrict__ covT /* [F][w][h][3] */) {
  __shared__ float tile[32][32 * 3 + 3];  // [x][3 y + c]; 99-float rows: both access patterns are conflict-free
  const int f = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  const unsigned char* img = gray + (size_t)f * h * w;
  const int x = x0 + lane;
  // A warp owns four consecutive rows: the six source rows they touch are read once (3 bytes each), their row
  // filters evaluated once, and the four pixels combine them -- the same operations per pixel as before.
  if (x < w && y0 + warp * 4 < h) {
    const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
    // row filter [s 2s s] and the central difference, rows y-1 .. y+4.  cv2's row filter fuses the chain in its
    // 32-pixel vector blocks; the last w % 32 pixels of a row go through its scalar loop, which does not
    // (found black-box: the eigenvalue map differed from the wheel by an ulp right of column (w / 32) * 32).
    const bool scalar_tail = x0 >= (w / 32) * 32;  // block-uniform: tiles are 32 columns wide
    float sm[6], df[6];
#pragma unroll
    for (int r = 0; r < 6; r++) {
      const int yy = reflect101(min(y0 + warp * 4 - 1 + r, h), h);  // rows past y+1 of the last valid pixel are never used
      const unsigned char* row = img + yy * w;
      const int ia = row[xm], ib = row[x], ic = row[xp];
      const float a = (float)ia, b = (float)ib, c = (float)ic;
      sm[r] = scalar_tail ? (s * a + s2 * b) + s * c : fmaf(s, c, fmaf(s2, b, s * a));
      df[r] = (float)(ic - ia);
    }
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
      const int yl = warp * 4 + rr;
      if (y0 + yl < h) {
        const float dy = sm[rr + 2] - sm[rr];
        const float dx = fmaf(df[rr] + df[rr + 2], s, df[rr + 1] * s2);
        tile[lane][yl * 3] = dx * dx;
        tile[lane][yl * 3 + 1] = dx * dy;
        tile[lane][yl * 3 + 2] = dy * dy;
      }
    }
  }
  __syncthreads();
  const int nvalid = min(32, h - y0) * 3;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int xl = warp * 4 + k;
    if (x0 + xl < w) {
      float* dst = covT + (((size_t)f * w + x0 + xl) * h + y0) * 3;
      for (int e = lane; e < nvalid; e += 32) dst[e] = tile[xl][e];
    }
  }
}

// horizontal running sums in double (cv::boxFilter's sliding sum, add for add): one lane per (frame, row),
// one warp per 32 consecutive rows.  Loads from covT are coalesced across the rows at every step; the sums
// of 32 steps are collected in a shared-memory tile and leave row-major (rows[f][y][x][3]) as coalesced
// 768-byte row segments, which is the layout the vertical pass streams.
constexpr int kRowTileStride = 32 * 3 + 1;  // doubles per x-slot of the output tile
#ifndef VSTAB_BOX_XS
#define VSTAB_BOX_XS 8
#endif
constexpr int kRowXs = VSTAB_BOX_XS;  // x-slots per tile.  Measured on 241 x 960x540: 32 slots (99 KB per CTA, 8 warps/SM) 3.09 ms, 16: 2.43 ms, 8 (25 KB): 1.44 ms, 4: 2.05 ms
__global__ void __launch_bounds__(128) gftt_box_rows_kernel(const float* __restrict__ covT, int F, int h, int w,
                                                            double* __restrict__ rows /* [F][h][w][3] */) {
  extern __shared__ double s_rows[];  // [4 warps][kRowXs x-slots][kRowTileStride]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ytiles = (h + 31) / 32;
  const int t = blockIdx.x * 4 + warp;
  if (t >= F * ytiles) return;  // whole warp
  const int f = t / ytiles, y0 = (t - f * ytiles) * 32;
  const int y = min(y0 + lane, h - 1);  // lanes past the last row repeat it and are never written out
  const size_t stride = (size_t)h * 3;
  const float* c = covT + (size_t)f * w * stride + (size_t)y * 3;
  double* tile = s_rows + (size_t)warp * kRowXs * kRowTileStride;
  double a0 = 0, a1 = 0, a2 = 0;
  for (int i = 0; i < kBlock; i++) {
    const float* q = c + (size_t)reflect101(i - kRadius, w) * stride;
    a0 += (double)q[0];
    a1 += (double)q[1];
    a2 += (double)q[2];
  }
  for (int x0 = 0; x0 < w; x0 += kRowXs) {
    const int cnt = min(kRowXs, w - x0);
    if (x0 - 1 - kRadius >= 0 && x0 + kRowXs - 1 + kRadius < w) {
      // no reflection in this chunk: the loads of four steps are issued before the chain consumes them
#pragma unroll 2
      for (int xs = 0; xs < kRowXs; xs += 4) {
        float in[4][3], out[4][3];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float* qn = c + (size_t)(x0 + xs + q + kRadius) * stride;
          const float* qo = c + (size_t)(x0 + xs + q - 1 - kRadius) * stride;
          in[q][0] = qn[0]; in[q][1] = qn[1]; in[q][2] = qn[2];
          out[q][0] = qo[0]; out[q][1] = qo[1]; out[q][2] = qo[2];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          a0 += (double)in[q][0] - (double)out[q][0];
          a1 += (double)in[q][1] - (double)out[q][1];
          a2 += (double)in[q][2] - (double)out[q][2];
          double* o = tile + (xs + q) * kRowTileStride + lane * 3;
          o[0] = a0; o[1] = a1; o[2] = a2;
        }
      }
    } else {
      for (int xs = 0; xs < cnt; xs++) {
        const int x = x0 + xs;
        if (x > 0) {
          const float* qn = c + (size_t)reflect101(x + kBlock - 1 - kRadius, w) * stride;
          const float* qo = c + (size_t)reflect101(x - 1 - kRadius, w) * stride;
          a0 += (double)qn[0] - (double)qo[0];
          a1 += (double)qn[1] - (double)qo[1];
          a2 += (double)qn[2] - (double)qo[2];
        }
        double* o = tile + xs * kRowTileStride + lane * 3;
        o[0] = a0; o[1] = a1; o[2] = a2;
      }
    }
    __syncwarp();
    const int n = cnt * 3;
    const int rmax = min(32, h - y0);
    for (int r = 0; r < rmax; r++) {
      double* dst = rows + (((size_t)f * h + y0 + r) * w + x0) * 3;
      for (int e = lane; e < n; e += 32) dst[e] = tile[(e / 3) * kRowTileStride + r * 3 + (e % 3)];
    }
    __syncwarp();
  }
}

// vertical running sums in double + minimum eigenvalue: one thread per (frame, column)
// The grid is 8 blocks per SM walking the (frame, column) items in order, NOT one thread per item: with every frame in
// flight at once the 21-row windows of all frames (116 MB for 241 frames) fall out of L2 and the row that leaves the
// window is read from DRAM a second time.  Measured on 241 x 960x540: one thread per item 5.88 GB read, 1.42 ms;
// 4 blocks per SM 3.88 GB but 1.93 ms (too few loads in flight); 8 blocks per SM 5.17 GB, 1.35 ms.
__global__ void __launch_bounds__(128) gftt_box_cols_kernel(const double* __restrict__ rows, int F, int h, int w, float* __restrict__ eig,
                                                            unsigned int* __restrict__ max_bits /* [F], float bits of the max (eig > 0) */) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < F * w; idx += gridDim.x * blockDim.x) {
    const int f = idx / w, x = idx - f * w;
    const double* R = rows + (size_t)f * h * w * 3 + (size_t)x * 3;
    const size_t rs = (size_t)w * 3;
    double s0 = 0, s1 = 0, s2 = 0;
    for (int i = 0; i < kBlock - 1; i++) {
      const int y = reflect101(i - kRadius, h);
      s0 += R[y * rs]; s1 += R[y * rs + 1]; s2 += R[y * rs + 2];
    }
    float best = 0.f;
    float* E = eig + (size_t)f * h * w + x;
    auto step = [&](int y, int yn, int yo) {
      const double t0 = s0 + R[yn * rs], t1 = s1 + R[yn * rs + 1], t2 = s2 + R[yn * rs + 2];
      const float a = (float)t0 * 0.5f, b = (float)t1, c = (float)t2 * 0.5f;
      const float e = (a + c) - sqrtf((a - c) * (a - c) + b * b);
      E[(size_t)y * w] = e;
      best = fmaxf(best, e);
      s0 = t0 - R[yo * rs]; s1 = t1 - R[yo * rs + 1]; s2 = t2 - R[yo * rs + 2];
    };
    // rows whose window touches the border reflect; the interior is unrolled so that the loads of four steps are in flight
    const int y_lo = min(kRadius, h), y_hi = max(y_lo, h - (kBlock - 1 - kRadius));
    for (int y = 0; y < y_lo; y++) step(y, reflect101(y + kBlock - 1 - kRadius, h), reflect101(y - kRadius, h));
#pragma unroll 4
    for (int y = y_lo; y < y_hi; y++) step(y, y + kBlock - 1 - kRadius, y - kRadius);
    for (int y = y_hi; y < h; y++) step(y, reflect101(y + kBlock - 1 - kRadius, h), reflect101(y - kRadius, h));
    atomicMax(max_bits + f, __float_as_uint(best));  // non-negative floats order like their bit patterns
  }
}

// thresholded 3x3 local maxima -> 64-bit keys (value bits << 32 | pixel index), appended per frame
__global__ void __launch_bounds__(256) gftt_candidates_kernel(const float* __restrict__ eig, int h, int w,
                                                              const unsigned int* __restrict__ max_bits, double quality,
                                                              unsigned long long* __restrict__ keys, int cap,
                                                              int* __restrict__ counts) {
  const int f = blockIdx.z;
  const int x = 1 + blockIdx.x * 32 + (threadIdx.x & 31);
  const int y0 = 1 + (blockIdx.y * 8 + (threadIdx.x >> 5)) * 8;  // 8 rows per thread: the grid was bound by block launches
  if (y0 >= h - 1) return;  // whole warp (a warp is one row of the block)
  const bool live = x < w - 1;
  const float* e = eig + (size_t)f * h * w;
  const float thr = (float)((double)__uint_as_float(max_bits[f]) * quality);
  // All candidates of a frame bump ONE counter: a warp reserves its slots with a single atomic (the keys are
  // sorted afterwards, so the slot order does not matter).  Lanes past the right edge stay in the loop for the ballot.
  // 3x3 dilation from per-row maxima: 3 loads per row instead of 9 per pixel (max is exact).
  // The ten rows a thread needs are loaded up front (30 loads in flight); with one row per loop iteration every
  // iteration waited for DRAM before its ballot.
  const float* ex = e + (live ? x : 1);
  float hm[10], ce[10];
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const int y = min(y0 - 1 + r, h - 1);
    const float* q = ex + (size_t)y * w;
    const float l = q[-1], m = q[0], rr = q[1];
    ce[r] = m;
    hm[r] = fmaxf(fmaxf(l, m), rr);
  }
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int y = y0 + r;
    const bool row_ok = y < h - 1;  // warp-uniform
    const float v = ce[r + 1];
    const bool cand = live && row_ok && (v > thr) && (v == fmaxf(fmaxf(hm[r], hm[r + 1]), hm[r + 2]));
    const unsigned votes = __ballot_sync(0xffffffffu, cand);
    if (votes) {
      const int lane = threadIdx.x & 31, leader = __ffs(votes) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(counts + f, __popc(votes));
      base = __shfl_sync(0xffffffffu, base, leader);
      const int slot = base + __popc(votes & ((1u << lane) - 1u));
      if (cand && slot < cap) keys[(size_t)f * cap + slot] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)(y * w + x);
    }
  }
}

// descending bitonic sort of each frame's keys in global memory (one CTA per frame)
__global__ void __launch_bounds__(1024) gftt_sort_kernel(unsigned long long* __restrict__ keys, int cap, const int* __restrict__ counts) {
  const int f = blockIdx.x;
  unsigned long long* k = keys + (size_t)f * cap;
  const int n = min(counts[f], cap);
  int m = 1;
  while (m < n) m <<= 1;
  for (int i = n + threadIdx.x; i < m; i += blockDim.x) k[i] = 0ull;  // pads sort to the end
  __syncthreads();
  for (int size = 2; size <= m; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const unsigned long long a = k[i], b = k[j];
          const bool desc = (i & size) == 0;
          if ((a < b) == desc) { k[i] = b; k[j] = a; }
        }
      }
      __syncthreads();
    }
}

// greedy minimum-distance selection: one warp per frame, 9 lanes probe the 3x3 neighbour cells
__global__ void gftt_select_kernel(const unsigned long long* __restrict__ keys, int cap, const int* __restrict__ counts, int h, int w,
                                   int max_corners, short* __restrict__ cells /* [F][gh*gw][kCellCap][2] */,
                                   unsigned char* __restrict__ cell_cnt /* [F][gh*gw], zeroed */, float* __restrict__ feats /* [F][max][2] */,
                                   int* __restrict__ n_feats) {
  const int f = blockIdx.x;
  const int lane = threadIdx.x;
  const int gw = (w + kCell - 1) / kCell, gh = (h + kCell - 1) / kCell;
  short* cl = cells + (size_t)f * gw * gh * kCellCap * 2;
  unsigned char* cc = cell_cnt + (size_t)f * gw * gh;
  const unsigned long long* k = keys + (size_t)f * cap;
  const int n = min(counts[f], cap);
  int accepted = 0;
  for (int i = 0; i < n && accepted < max_corners; i++) {
    const int idx = (int)(k[i] & 0xffffffffu);
    const int y = idx / w, x = idx - y * w;
    const int xc = x / kCell, yc = y / kCell;
    bool bad = false;
    if (lane < 9) {
      const int cx = xc + lane % 3 - 1, cy = yc + lane / 3 - 1;
      if (cx >= 0 && cy >= 0 && cx < gw && cy < gh) {
        const int c = cy * gw + cx;
        const int cn = cc[c];
        for (int j = 0; j < cn; j++) {
          const float dx = (float)(x - cl[(c * kCellCap + j) * 2]), dy = (float)(y - cl[(c * kCellCap + j) * 2 + 1]);
          if (dx * dx + dy * dy < kMinDist2) bad = true;
        }
      }
    }
    if (__any_sync(0xffffffffu, bad)) continue;
    if (lane == 0) {
      const int c = yc * gw + xc;
      const int cn = cc[c];
      if (cn < kCellCap) {
        cl[(c * kCellCap + cn) * 2] = (short)x;
        cl[(c * kCellCap + cn) * 2 + 1] = (short)y;
        cc[c] = (unsigned char)(cn + 1);
      }
      feats[((size_t)f * max_corners + accepted) * 2] = (float)x;
      feats[((size_t)f * max_corners + accepted) * 2 + 1] = (float)y;
    }
    accepted++;
    __syncwarp();
  }
  if (lane == 0) n_feats[f] = accepted;
}

// ---- pyramidal Lucas-Kanade ---------------------------------------------------------------------

// cv::pyrDown for 8-bit images: [1 4 6 4 1] x [1 4 6 4 1], (sum + 128) >> 8, reflect-101
__global__ void __launch_bounds__(256) pyr_down_kernel(const unsigned char* __restrict__ src, int sh, int sw,
                                                       unsigned char* __restrict__ dst, int dh, int dw) {
  const int f = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y0 = (blockIdx.y * 8 + (threadIdx.x >> 5)) * 8;  // 8 rows per thread (kRowsPerThread, declared below)
  if (x >= dw || y0 >= dh) return;
  const unsigned char* s = src + (size_t)f * sh * sw;
  int cx[5];
#pragma unroll
  for (int k = 0; k < 5; k++) cx[k] = reflect101(2 * x + k - 2, sw);
  const int wk[5] = {1, 4, 6, 4, 1};
  for (int y = y0; y < min(y0 + 8, dh); y++) {
    int sum = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
      const unsigned char* r = s + (size_t)reflect101(2 * y + j - 2, sh) * sw;
      const int row = r[cx[0]] + 4 * r[cx[1]] + 6 * r[cx[2]] + 4 * r[cx[3]] + r[cx[4]];
      sum += wk[j] * row;
    }
    dst[((size_t)f * dh + y) * dw + x] = (unsigned char)((sum + 128) >> 8);
  }
}

// Scharr derivatives (dx, dy) interleaved int16, reflect-101 inside the image
constexpr int kRowsPerThread = 8;  // one-pixel-per-thread grids of these small kernels were bound by block launches, not by work

__global__ void __launch_bounds__(256) scharr_kernel(const unsigned char* __restrict__ img, int h, int w, short2* __restrict__ d) {
  const int f = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y0 = (blockIdx.y * 8 + (threadIdx.x >> 5)) * kRowsPerThread;
  if (x >= w || y0 >= h) return;
  const unsigned char* I = img + (size_t)f * h * w;
  const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
  for (int y = y0; y < min(y0 + kRowsPerThread, h); y++) {
    const unsigned char* r0 = I + (size_t)reflect101(y - 1, h) * w;
    const unsigned char* r1 = I + (size_t)y * w;
    const unsigned char* r2 = I + (size_t)reflect101(y + 1, h) * w;
    const int t0p = (r0[xp] + r2[xp]) * 3 + r1[xp] * 10, t0m = (r0[xm] + r2[xm]) * 3 + r1[xm] * 10;
    const int t1m = r2[xm] - r0[xm], t1c = r2[x] - r0[x], t1p = r2[xp] - r0[xp];
    d[((size_t)f * h + y) * w + x] = make_short2((short)(t0p - t0m), (short)((t1p + t1m) * 3 + t1c * 10));
  }
}

constexpr int kLkPad = 32;  // >= kWin + 1: every tap of a window the tracker accepts lies inside the padded copy
struct LkLevel {
  const unsigned char* img;  // [F][h][w]
  const unsigned char* ext;  // [F][h + 2 kLkPad][w + 2 kLkPad]: img with its reflect-101 border written out
  const short2* deriv;       // [F][h][w]
  int h, w;
};

// ext[y][x] = img[reflect101(y - kLkPad)][reflect101(x - kLkPad)]: the search windows of the tracker hang over the
// frame border for most features of the coarse levels (a 31x31 window in a 120x67 image); with the border written
// out once, every Gauss-Newton iteration walks its window without reflecting coordinates.
__global__ void __launch_bounds__(256) lk_pad_kernel(const unsigned char* __restrict__ img, int h, int w, unsigned char* __restrict__ ext) {
  // four consecutive bytes per thread: one 32-bit load + store wherever the four source bytes are consecutive and aligned
  const int f = blockIdx.z;
  const int x = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int pw = w + 2 * kLkPad, ph = h + 2 * kLkPad;
  if (x >= pw || y >= ph) return;
  const unsigned char* row = img + ((size_t)f * h + reflect101(y - kLkPad, h)) * w;
  unsigned char* dst = ext + ((size_t)f * ph + y) * pw + x;
  const int sx = x - kLkPad;
  if (((w | pw) & 3) == 0 && sx >= 0 && sx + 3 < w && ((((size_t)f * h) * w) & 3) == 0) {
    *reinterpret_cast<unsigned int*>(dst) = *reinterpret_cast<const unsigned int*>(row + sx);
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (x + k < pw) dst[k] = row[reflect101(sx + k, w)];
  }
}
struct LkPyr {
  LkLevel lv[kMaxLevel + 1];
  int levels;  // highest level index
};

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// Per-warp shared memory of the tracker: the template window (I, Ix, Iy as int16, [y][x][3]) and a double-buffered
// row of float terms through which the lanes hand their products to the lanes that own a summation chain.
constexpr int kLkWinBytes = (kWin * kWin * 3 * 2 + 15) & ~15;   // 5776
constexpr int kLkRowFloats = 2 * 3 * 40;                         // A sums: 2 buffers x 3 quantities x 40 slots
constexpr int kLkWarpBytes = kLkWinBytes + kLkRowFloats * 4;     // 6736

// tail + ((q0 + q2) + (q1 + q3)): the final fold of cv2's lane accumulators; chain lanes first..first+3 and `tail`.
__device__ __forceinline__ float lk_fold(float acc, int first, int tail) {
  const float q0 = __shfl_sync(0xffffffffu, acc, first), q1 = __shfl_sync(0xffffffffu, acc, first + 1);
  const float q2 = __shfl_sync(0xffffffffu, acc, first + 2), q3 = __shfl_sync(0xffffffffu, acc, first + 3);
  const float t = __shfl_sync(0xffffffffu, acc, tail);
  return t + ((q0 + q2) + (q1 + q3));
}

// One warp per (pair, feature); lane x owns column x of the 31x31 window and walks down its rows (lane 31 idles).
//
// The window sums are float sums of integer terms, and above 2^24 their value depends on the order of the additions.
// cv2's order (its 128-bit SIMD code, found black-box, see oracle/classic_ref.c pyr_lk_body cv_order = 1): of every
// window row the first 24 columns go to four lane accumulators -- structure tensor: lane = x mod 4; mismatch vector:
// int32 pair sums d[x] g[x] + d[x+4] g[x+4] of the 8-column blocks, lane = x mod 4 -- and columns 24..30 to one scalar
// accumulator that runs on across the rows; at the end tail + ((l0 + l2) + (l1 + l3)).  These are 5 sequential chains
// per sum (4 x 186 + 217 terms for A, 4 x 93 + 217 for b).  Here every lane computes the term of its column, the
// terms of a row go through a small shared-memory row buffer, and 15 (A11, A12, A22) / 10 (b1, b2) lanes each own
// one chain and add their 3..7 terms of the row in order.  Same terms, same order, same bits as the wheel.
__global__ void __launch_bounds__(256) lk_track_kernel(LkPyr P, int n_pairs, int max_corners, const float* __restrict__ feats,
                                                       const int* __restrict__ n_feats, float* __restrict__ prev_out,
                                                       float* __restrict__ curr_out) {
  extern __shared__ __align__(16) unsigned char s_lk[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = blockIdx.x * 8 + warp;
  if (gid >= n_pairs * max_corners) return;
  const int pair = gid / max_corners, fi = gid - pair * max_corners;
  const float qnan = __int_as_float(0x7fc00000);
  float* po = prev_out + (size_t)gid * 2;
  float* co = curr_out + (size_t)gid * 2;
  if (fi >= n_feats[pair]) {
    if (lane == 0) { po[0] = qnan; po[1] = qnan; co[0] = qnan; co[1] = qnan; }
    return;
  }
  const float fx = feats[((size_t)pair * max_corners + fi) * 2], fy = feats[((size_t)pair * max_corners + fi) * 2 + 1];
  short* win = reinterpret_cast<short*>(s_lk + warp * kLkWarpBytes);
  float* rb = reinterpret_cast<float*>(s_lk + warp * kLkWarpBytes + kLkWinBytes);
  const float half = (kWin - 1) * 0.5f;
  const float FLT_SCALE = 1.f / (1 << 20);
  const int W_BITS = 14;
  const int x = lane;
  const bool col_on = x < kWin;
  // where this column's terms go in the row buffers, and which chain (if any) this lane adds up
  const int a_slot = x < 24 ? (x & 3) * 8 + (x >> 2) : 32 + (x - 24);          // A: 4 chains x 6 columns, tail x 7
  const int b_slot = x < 24 ? (x & 3) * 4 + (x >> 3) : 16 + (x - 24);          // b: 4 chains x 3 blocks, tail x 7
  const bool b_store = col_on && (x >= 24 || (x & 7) < 4);
  const bool a_chain = lane < 15, a_tail = lane >= 12;
  const int a_base = (a_tail ? lane - 12 : lane >> 2) * 40 + (a_tail ? 32 : (lane & 3) * 8);
  const bool b_chain = lane < 10, b_tail = lane >= 8;
  const int b_base = (b_tail ? lane - 8 : lane >> 2) * 24 + (b_tail ? 16 : (lane & 3) * 4);
  bool status = true;
  float nx = 0.f, ny = 0.f;
  for (int level = P.levels; level >= 0; level--) {
    const LkLevel L = P.lv[level];
    const int pw = L.w + 2 * kLkPad;
    const unsigned char* Jext = L.ext + (size_t)(pair + 1) * (L.h + 2 * kLkPad) * pw;
    const short2* D = L.deriv + (size_t)pair * L.h * L.w;
    const int lh = L.h, lw = L.w;
    float px = fx * (float)(1. / (1 << level)), py = fy * (float)(1. / (1 << level));
    if (level == P.levels) { nx = px; ny = py; } else { nx = nx * 2.f; ny = ny * 2.f; }
    px -= half; py -= half;
    const int ipx = (int)floorf(px), ipy = (int)floorf(py);
    if (ipx < -kWin || ipx >= lw || ipy < -kWin || ipy >= lh) { if (level == 0) status = false; continue; }
    float a = px - ipx, b = py - ipy;
    int iw00 = __float2int_rn((1.f - a) * (1.f - b) * (1 << W_BITS)), iw01 = __float2int_rn(a * (1.f - b) * (1 << W_BITS));
    int iw10 = __float2int_rn((1.f - a) * b * (1 << W_BITS)), iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
    float A11, A12, A22;
    __syncwarp();
    {
      // template window: intensities from the padded copy (no reflection), derivatives zero outside the frame.
      // Walking down a column, the lower two taps of a row are the upper two of the next one.
      const int cx = col_on ? x : 0;
      const unsigned char* q = L.ext + (size_t)pair * (lh + 2 * kLkPad) * pw + (ipy + kLkPad) * pw + (ipx + kLkPad) + cx;
      const int sx = ipx + cx;
      const bool in_x0 = sx >= 0 && sx < lw, in_x1 = sx + 1 >= 0 && sx + 1 < lw;
      const short2 zero = make_short2(0, 0);
      int i0 = q[0], i1 = q[1];
      short2 d0 = (in_x0 && ipy >= 0 && ipy < lh) ? D[ipy * lw + sx] : zero;
      short2 d1 = (in_x1 && ipy >= 0 && ipy < lh) ? D[ipy * lw + sx + 1] : zero;
      float acc = 0.f;
      short* wq = win + cx * 3;
      for (int y = 0; y < kWin; y++) {
        q += pw;
        const int sy1 = ipy + y + 1;
        const bool in_y1 = sy1 >= 0 && sy1 < lh;
        const int j0 = q[0], j1 = q[1];
        const short2 e0 = (in_x0 && in_y1) ? D[sy1 * lw + sx] : zero;
        const short2 e1 = (in_x1 && in_y1) ? D[sy1 * lw + sx + 1] : zero;
        const int ival = descale(i0 * iw00 + i1 * iw01 + j0 * iw10 + j1 * iw11, W_BITS - 5);
        const int ixval = descale(d0.x * iw00 + d1.x * iw01 + e0.x * iw10 + e1.x * iw11, W_BITS);
        const int iyval = descale(d0.y * iw00 + d1.y * iw01 + e0.y * iw10 + e1.y * iw11, W_BITS);
        i0 = j0; i1 = j1; d0 = e0; d1 = e1;
        float* rrow = rb + (y & 1) * 120;
        if (col_on) {
          wq[0] = (short)ival; wq[1] = (short)ixval; wq[2] = (short)iyval;
          rrow[a_slot] = (float)(ixval * ixval);
          rrow[40 + a_slot] = (float)(ixval * iyval);
          rrow[80 + a_slot] = (float)(iyval * iyval);
        }
        wq += kWin * 3;
        __syncwarp();
        if (a_chain) {
          const float4 v0 = *reinterpret_cast<const float4*>(rrow + a_base);
          const float4 v1 = *reinterpret_cast<const float4*>(rrow + a_base + 4);
          acc += v0.x; acc += v0.y; acc += v0.z; acc += v0.w; acc += v1.x; acc += v1.y;
          if (a_tail) acc += v1.z;
        }
      }
      A11 = lk_fold(acc, 0, 12) * FLT_SCALE;
      A12 = lk_fold(acc, 4, 13) * FLT_SCALE;
      A22 = lk_fold(acc, 8, 14) * FLT_SCALE;
    }
    __syncwarp();
    float Dd = A11 * A22 - A12 * A12;
    const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * kWin * kWin);
    if (minEig < 1e-4f || Dd < FLT_EPSILON) { if (level == 0) status = false; continue; }
    Dd = 1.f / Dd;
    float wx = nx - half, wy = ny - half;
    float pdx = 0, pdy = 0;
    for (int j = 0; j < kMaxIter; j++) {
      const int inx = (int)floorf(wx), iny = (int)floorf(wy);
      if (inx < -kWin || inx >= lw || iny < -kWin || iny >= lh) { if (level == 0) status = false; break; }
      a = wx - inx; b = wy - iny;
      iw00 = __float2int_rn((1.f - a) * (1.f - b) * (1 << W_BITS)); iw01 = __float2int_rn(a * (1.f - b) * (1 << W_BITS));
      iw10 = __float2int_rn((1.f - a) * b * (1 << W_BITS)); iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
      float b1, b2;
      {
        const int cx = col_on ? x : 0;
        const unsigned char* q = Jext + (iny + kLkPad) * pw + (inx + kLkPad) + cx;
        int i0 = q[0], i1 = q[1];
        const short* wp = win + cx * 3;
        float acc = 0.f;
        for (int y = 0; y < kWin; y++) {
          q += pw;
          const int j0 = q[0], j1 = q[1];
          const int diff = descale(i0 * iw00 + i1 * iw01 + j0 * iw10 + j1 * iw11, W_BITS - 5) - wp[0];
          i0 = j0; i1 = j1;
          const int p1 = diff * wp[1], p2 = diff * wp[2];
          wp += kWin * 3;
          // columns 0..23: int32 pair sums of the columns x and x + 4 of an 8-column block, held by the lower column
          const int s1 = p1 + __shfl_down_sync(0xffffffffu, p1, 4), s2 = p2 + __shfl_down_sync(0xffffffffu, p2, 4);
          float* rrow = rb + (y & 1) * 48;
          if (b_store) {
            rrow[b_slot] = (float)(x < 24 ? s1 : p1);
            rrow[24 + b_slot] = (float)(x < 24 ? s2 : p2);
          }
          __syncwarp();
          if (b_chain) {
            const float4 v0 = *reinterpret_cast<const float4*>(rrow + b_base);
            const float4 v1 = *reinterpret_cast<const float4*>(rrow + b_base + 4);
            acc += v0.x; acc += v0.y; acc += v0.z;
            if (b_tail) { acc += v0.w; acc += v1.x; acc += v1.y; acc += v1.z; }
          }
        }
        b1 = lk_fold(acc, 0, 8) * FLT_SCALE;
        b2 = lk_fold(acc, 4, 9) * FLT_SCALE;
      }
      const float ddx = (A12 * b2 - A22 * b1) * Dd, ddy = (A12 * b1 - A11 * b2) * Dd;
      wx += ddx; wy += ddy;
      nx = wx + half; ny = wy + half;
      if ((double)ddx * ddx + (double)ddy * ddy <= 1e-4) break;
      if (j > 0 && fabsf(ddx + pdx) < 0.01f && fabsf(ddy + pdy) < 0.01f) {
        nx -= ddx * 0.5f; ny -= ddy * 0.5f;
        break;
      }
      pdx = ddx; pdy = ddy;
    }
    if (status && level == 0) {
      const int ix = (int)floorf(nx - half), iy = (int)floorf(ny - half);
      if (ix < -kWin || ix >= lw || iy < -kWin || iy >= lh) status = false;
    }
  }
  if (lane == 0) {
    po[0] = fx; po[1] = fy;
    co[0] = status ? nx : qnan;
    co[1] = status ? ny : qnan;
  }
}

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

// Streams, events and workspace of one pass.  Two slots alternate when a clip is tracked in several passes, so that the
// tracker of pass k (issue-bound) runs under the corner detector of pass k + 1 (bandwidth-bound).
struct PassSlot {
  cudaStream_t main, pyr;               // corner detector + tracker | pyramids, derivatives, padded copies
  cudaEvent_t fork, join, corners;      // pass started | pyramids done | corner detector done (releases the next pass)
  unsigned char* base;                  // this slot's share of the handle's workspace
};

// slot == nullptr: only report the workspace one pass of n_frames needs
int gftt_lk_run(vstab_handle* hnd, const uint8_t* gray_dev, int n_frames, int height, int width, int max_corners,
                float* prev_dev, float* curr_dev, int32_t* detected_dev, const PassSlot* slot, size_t* bytes_out);

}  // namespace

// Pairs per pass: corner scratch, LK pyramids, Scharr derivatives and padded copies are sized for one pass
// (~26 MB per 960x540 frame), so the workspace stays bounded however long the clip is -- like dis_run's chunks.
// Frame p0 + kPairsPerPass is the last frame of one pass and the first of the next; its pyramid is built twice.
static const int kPairsPerPass = 256;

extern "C" int vstab_gftt_lk(vstab_handle* hnd, const uint8_t* gray_dev, int n_frames, int height, int width, int max_corners,
                             float* prev_dev, float* curr_dev, int32_t* detected_dev, void* stream) {
  if (!hnd) return vstab_fail(nullptr, VSTAB_ERR_INVALID, "vstab_gftt_lk: null handle");
  if (!gray_dev || !prev_dev || !curr_dev || !detected_dev || n_frames < 0 || height < 3 || width < 3 || max_corners <= 0)
    return vstab_fail(hnd, VSTAB_ERR_INVALID, "vstab_gftt_lk: bad argument");
  if (n_frames < 2) return VSTAB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  VSTAB_ENTER(hnd);
  const size_t npx = (size_t)height * width;
  const int P = n_frames - 1;
  int per_pass = kPairsPerPass;
  const char* env_pass = getenv("VSTAB_LK_PAIRS_PER_PASS");  // tests: several passes on a short clip
  if (env_pass && atoi(env_pass) > 0) per_pass = atoi(env_pass);
  // (a 240-pair clip split into two overlapping halves was measured: 14.8 ms against 13.6 ms in one pass -- the corner
  // detector's kernels are chains whose latency does not halve with the frame count -- so only clips beyond one pass use the slots)
  const int passes = (P + per_pass - 1) / per_pass;
  const char* env_pipe = getenv("VSTAB_LK_PIPELINE");
  const bool pipelined = passes >= 2 && !(env_pipe && atoi(env_pipe) == 0);
  size_t bytes = 0;
  int rc = gftt_lk_run(hnd, gray_dev, (P < per_pass ? P : per_pass) + 1, height, width, max_corners, prev_dev, curr_dev, detected_dev, nullptr, &bytes);
  if (rc != VSTAB_OK) return rc;
  void* wsp = nullptr;
  rc = vstab_workspace(hnd, bytes * (pipelined ? 2 : 1), &wsp);
  if (rc != VSTAB_OK) return rc;
  rc = vstab_aux_streams(hnd, pipelined ? 3 : 1);
  if (rc != VSTAB_OK) return rc;
  PassSlot slots[2];
  slots[0] = {st, hnd->aux_stream[0], hnd->level_event[0], hnd->level_event[1], hnd->stagger_event[0], (unsigned char*)wsp};
  if (pipelined) {
    slots[1] = {hnd->aux_stream[1], hnd->aux_stream[2], hnd->level_event[2], hnd->level_event[3], hnd->stagger_event[1], (unsigned char*)wsp + bytes};
    VSTAB_CUDA(hnd, cudaEventRecord(hnd->fork_event, st));  // the second slot's stream starts behind whatever `st` holds already
    VSTAB_CUDA(hnd, cudaStreamWaitEvent(slots[1].main, hnd->fork_event, 0));
  }
  for (int k = 0, p0 = 0; p0 < P; ++k, p0 += per_pass) {
    const int pairs = (P - p0) < per_pass ? (P - p0) : per_pass;
    const PassSlot& slot = slots[pipelined ? (k & 1) : 0];
    // pass k starts its corner detector when pass k - 1 has finished its own (they would only share the bandwidth);
    // from there on it runs next to the tracker of pass k - 1
    if (pipelined && k > 0) VSTAB_CUDA(hnd, cudaStreamWaitEvent(slot.main, slots[(k - 1) & 1].corners, 0));
    rc = gftt_lk_run(hnd, gray_dev + p0 * npx, pairs + 1, height, width, max_corners, prev_dev + (size_t)p0 * max_corners * 2,
                     curr_dev + (size_t)p0 * max_corners * 2, detected_dev + p0, &slot, nullptr);
    if (rc != VSTAB_OK) return rc;
  }
  if (pipelined) {
    VSTAB_CUDA(hnd, cudaEventRecord(hnd->join_event[1], slots[1].main));
    VSTAB_CUDA(hnd, cudaStreamWaitEvent(st, hnd->join_event[1], 0));
  }
  return VSTAB_OK;
}

namespace {

int gftt_lk_run(vstab_handle* hnd, const uint8_t* gray_dev, int n_frames, int height, int width, int max_corners,
                float* prev_dev, float* curr_dev, int32_t* detected_dev, const PassSlot* slot, size_t* bytes_out) {
  const int h = height, w = width;
  // frames per pass of the corner detector: every running-sum / selection kernel is a chain per row, column or
  // frame, so a pass costs its chain latency whatever the frame count (20.7 MB of scratch per 960x540 frame)
  const int kChunk = (n_frames - 1) < 256 ? (n_frames - 1) : 256;
  int cap = 1024;  // per-frame candidate capacity: power of two >= h*w/4 (a 3x3 local maximum needs its own 2x2 block)
  while (cap < (h * w) / 4) cap <<= 1;
  const int gw = (w + kCell - 1) / kCell, gh = (h + kCell - 1) / kCell;
  const int P = n_frames - 1;

  // ---- workspace ----
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  const size_t o_cov = take(sizeof(float) * 3 * (size_t)kChunk * h * w);
  const size_t o_rows = take(sizeof(double) * 3 * (size_t)kChunk * h * w);
  const size_t o_eig = take(sizeof(float) * (size_t)kChunk * h * w);
  const size_t o_max = take(sizeof(unsigned) * kChunk);
  const size_t o_keys = take(sizeof(unsigned long long) * (size_t)kChunk * cap);
  const size_t o_cnt = take(sizeof(int) * kChunk);
  const size_t o_cells = take(sizeof(short) * 2 * kCellCap * (size_t)kChunk * gw * gh);
  const size_t o_ccnt = take((size_t)kChunk * gw * gh);
  const size_t o_feats = take(sizeof(float) * 2 * (size_t)P * max_corners);
  LkPyr pyr;
  size_t o_img[kMaxLevel + 1], o_der[kMaxLevel + 1], o_ext[kMaxLevel + 1];
  int lh = h, lw = w;
  pyr.levels = 0;
  for (int l = 0; l <= kMaxLevel; l++) {
    if (l > 0) {
      const int dh = (lh + 1) / 2, dw = (lw + 1) / 2;
      if (dw <= kWin || dh <= kWin) break;
      lh = dh; lw = dw;
      o_img[l] = take((size_t)n_frames * lh * lw);
      pyr.levels = l;
    }
    pyr.lv[l].h = lh; pyr.lv[l].w = lw;
    o_der[l] = take(sizeof(short2) * (size_t)n_frames * lh * lw);
    o_ext[l] = take((size_t)n_frames * (lh + 2 * kLkPad) * (lw + 2 * kLkPad));
  }
  if (!slot) {
    *bytes_out = off;
    return VSTAB_OK;
  }
  cudaStream_t st = slot->main;
  unsigned char* base = slot->base;
  float* cov = (float*)(base + o_cov);
  double* rows = (double*)(base + o_rows);
  float* eig = (float*)(base + o_eig);
  unsigned* maxb = (unsigned*)(base + o_max);
  unsigned long long* keys = (unsigned long long*)(base + o_keys);
  int* cnt = (int*)(base + o_cnt);
  short* cells = (short*)(base + o_cells);
  unsigned char* ccnt = base + o_ccnt;
  float* feats = (float*)(base + o_feats);

  const double scale_d = 1.0 / ((double)(1 << 2) * kBlock * 255.0);
  const float s = (float)scale_d, s2 = s * 2.0f;
  // the pass starts here for both of its streams (everything before it on `st`, e.g. the slot's previous tracker, is done with the workspace)
  VSTAB_CUDA(hnd, cudaEventRecord(slot->fork, st));
  // ---- corners of frames 0 .. n-2 ----
  for (int f0 = 0; f0 < P; f0 += kChunk) {
    const int F = (P - f0) < kChunk ? (P - f0) : kChunk;
    const unsigned char* g = gray_dev + (size_t)f0 * h * w;
    dim3 gp(vstab_ceil_div(w, 32), vstab_ceil_div(h, 32), F);
    gftt_cov_kernel<<<gp, 256, 0, st>>>(g, h, w, s, s2, cov);
    VSTAB_LAUNCH_CHECK(hnd, "gftt_cov_kernel");
    {
      const size_t rows_smem = sizeof(double) * 4 * kRowXs * kRowTileStride;  // the tiles of 4 warps
      VSTAB_CUDA(hnd, cudaFuncSetAttribute(gftt_box_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rows_smem));
      gftt_box_rows_kernel<<<vstab_ceil_div(F * vstab_ceil_div(h, 32), 4), 128, rows_smem, st>>>(cov, F, h, w, rows);
    }
    VSTAB_LAUNCH_CHECK(hnd, "gftt_box_rows_kernel");
    VSTAB_CUDA(hnd, cudaMemsetAsync(maxb, 0, sizeof(unsigned) * kChunk, st));
    VSTAB_CUDA(hnd, cudaMemsetAsync(cnt, 0, sizeof(int) * kChunk, st));
    VSTAB_CUDA(hnd, cudaMemsetAsync(ccnt, 0, (size_t)kChunk * gw * gh, st));
    gftt_box_cols_kernel<<<min(vstab_ceil_div(F * w, 128), hnd->sm_count * 8), 128, 0, st>>>(rows, F, h, w, eig, maxb);
    VSTAB_LAUNCH_CHECK(hnd, "gftt_box_cols_kernel");
    dim3 gc(vstab_ceil_div(w - 2, 32), vstab_ceil_div(h - 2, 64), F);
    gftt_candidates_kernel<<<gc, 256, 0, st>>>(eig, h, w, maxb, 0.01, keys, cap, cnt);
    VSTAB_LAUNCH_CHECK(hnd, "gftt_candidates_kernel");
    gftt_sort_kernel<<<F, 1024, 0, st>>>(keys, cap, cnt);
    VSTAB_LAUNCH_CHECK(hnd, "gftt_sort_kernel");
    gftt_select_kernel<<<F, 32, 0, st>>>(keys, cap, cnt, h, w, max_corners, cells, ccnt, feats + (size_t)f0 * max_corners * 2,
                                          detected_dev + f0);
    VSTAB_LAUNCH_CHECK(hnd, "gftt_select_kernel");
  }
  // ---- pyramids + Scharr derivatives of every frame: independent of the corners, so they run on a helper stream
  // next to the corner detector (enqueued above on the caller's stream) and join it before the tracker ----
  VSTAB_CUDA(hnd, cudaEventRecord(slot->corners, st));
  cudaStream_t ps = slot->pyr;
  VSTAB_CUDA(hnd, cudaStreamWaitEvent(ps, slot->fork, 0));
  pyr.lv[0].img = gray_dev;
  for (int l = 0; l <= pyr.levels; l++) {
    if (l > 0) {
      unsigned char* dst = base + o_img[l];
      dim3 gd(vstab_ceil_div(pyr.lv[l].w, 32), vstab_ceil_div(pyr.lv[l].h, 64), n_frames);
      pyr_down_kernel<<<gd, 256, 0, ps>>>(pyr.lv[l - 1].img, pyr.lv[l - 1].h, pyr.lv[l - 1].w, dst, pyr.lv[l].h, pyr.lv[l].w);
      VSTAB_LAUNCH_CHECK(hnd, "pyr_down_kernel");
      pyr.lv[l].img = dst;
    }
    short2* der = (short2*)(base + o_der[l]);
    dim3 gs(vstab_ceil_div(pyr.lv[l].w, 32), vstab_ceil_div(pyr.lv[l].h, kRowsPerThread * 8), n_frames);
    scharr_kernel<<<gs, 256, 0, ps>>>(pyr.lv[l].img, pyr.lv[l].h, pyr.lv[l].w, der);
    VSTAB_LAUNCH_CHECK(hnd, "scharr_kernel");
    pyr.lv[l].deriv = der;
    unsigned char* ext = base + o_ext[l];
    dim3 ge(vstab_ceil_div(pyr.lv[l].w + 2 * kLkPad, 128), vstab_ceil_div(pyr.lv[l].h + 2 * kLkPad, 8), n_frames);
    lk_pad_kernel<<<ge, 256, 0, ps>>>(pyr.lv[l].img, pyr.lv[l].h, pyr.lv[l].w, ext);
    VSTAB_LAUNCH_CHECK(hnd, "lk_pad_kernel");
    pyr.lv[l].ext = ext;
  }
  VSTAB_CUDA(hnd, cudaEventRecord(slot->join, ps));
  VSTAB_CUDA(hnd, cudaStreamWaitEvent(st, slot->join, 0));
  // ---- track ----
  const size_t smem = (size_t)8 * kLkWarpBytes;
  VSTAB_CUDA(hnd, cudaFuncSetAttribute(lk_track_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lk_track_kernel<<<vstab_ceil_div(P * max_corners, 8), 256, smem, st>>>(pyr, P, max_corners, feats, detected_dev, prev_dev, curr_dev);
  VSTAB_LAUNCH_CHECK(hnd, "lk_track_kernel");
  return VSTAB_OK;
}

}  // namespace
