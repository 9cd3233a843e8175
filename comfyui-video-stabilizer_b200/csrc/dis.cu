// dis.cu -- K3 + K4: Dense Inverse Search optical flow, batched over all frame pairs of a clip.
//
// Replaces nodes/video_stabilizer_flow.py:76-87 (_create_flow_backend) and :140
// (cv2.DISOpticalFlow.calc) plus the 8-px grid sampling of :141-147.  The algorithm is OpenCV's
// DIS (Kroeger et al. 2016) in the configuration the reference sets: finest scale 2, 8x8 patches
// on a stride-4 grid, 2 x 12 inverse-compositional iterations with spatial propagation inside 8
// fixed horizontal stripes, mean-normalised residuals, densification, and 5 x 5 red-black SOR
// variational refinement (alpha 20, delta 5, gamma 10, eps 0.01, omega 1.6).
//
// Design for B200: every frame pair of the clip is independent, so each stage is ONE launch over
// all pairs (a 121-frame clip = 120 pairs keeps ~1000 warps in flight even at the 30x16 coarsest
// level).  Pyramids, gradients and structure tensors are built once per FRAME and shared by the
// two pairs that use it.  The sequential raster-order propagation of the patch search is run as
// an anti-diagonal wavefront: one warp per (pair, stripe), 8 patch rows x 4 lanes, where the 4
// lanes of a quad reproduce OpenCV's 4-wide SIMD accumulators and their reduction order
// ((a0+a2)+(a1+a3)), so the float results are bit-identical to the CPU reference and so are all
// the discrete decisions (candidate choice, early stop).  This file is compiled with
// -fmad=false: no multiply-add contraction anywhere except where the reference itself fuses.
//
// Schedule of one call (dis_run): the pyramid chain on the caller's stream; gradients / bordered copies / structure
// tensors per level on a preparation stream, coarsest level first; the pairs in four groups on forked streams, each
// running patch search -> densification -> variational refinement -> x2 upsampling per level and waiting for a level's
// preparation event right before it searches that level.  The refinement of a level is ONE launch: all 19 planes in one
// CTA's shared memory for the two coarsest levels (vr_fused_kernel<true>), the SOR state in shared memory / registers
// with row bands of a cluster exchanging halo rows through distributed shared memory where the bands fit
// (vr_resident_kernel), a cluster streaming its planes through L2 otherwise (vr_fused_kernel<false>).
#include "common.cuh"

#include <math.h>
#include <stdlib.h>

namespace {

//@emul-begin (tests/emul/vr_emul.cpp compiles the marked regions for the host)
constexpr int kBorder = 16;
constexpr int kPatch = 8;
constexpr int kStride = 4;
constexpr int kStripes = 8;
constexpr int kFinest = 2;
constexpr int kGdIter = 25;
constexpr int kVrIter = 5;
constexpr int kSorIter = 5;
constexpr float kEps = 0.001f;
constexpr float kInf = 1e10f;
constexpr float kAlpha = 20.f, kDelta = 5.f, kGamma = 10.f, kEpsilon = 0.01f, kOmega = 1.6f;
constexpr int kMaxLevels = 8;
#ifndef VSTAB_DIS_GROUPS_DEFAULT
#define VSTAB_DIS_GROUPS_DEFAULT 4
#endif
#ifndef VSTAB_DIS_STAGGER_DEFAULT
#define VSTAB_DIS_STAGGER_DEFAULT 0
#endif
#ifndef VSTAB_VR_ONCHIP_DEFAULT
#define VSTAB_VR_ONCHIP_DEFAULT 0
#endif
#ifndef VSTAB_VR_RESIDENT_DEFAULT
#define VSTAB_VR_RESIDENT_DEFAULT 2
#endif
#ifndef VSTAB_VR_RESIDENT_THREADS_DEFAULT
#define VSTAB_VR_RESIDENT_THREADS_DEFAULT 1024
#endif
#ifndef VSTAB_VR_MIN_CTAS
#define VSTAB_VR_MIN_CTAS 5
#endif

struct Level {
  int w, h, ws, hs;
  // per frame
  unsigned char* I;    // [F][h][w]
  unsigned char* Iext; // [F][h+32][w+32]
  short* Ix;           // [F][h][w]
  short* Iy;
  float* T;            // [F][5][hs][ws]  xx, yy, xy, x, y
  // per pair
  float* Ux;           // [P][h][w]
  float* Uy;
  float* Sx;           // [P][hs][ws]
  float* Sy;
};
//@emul-end

__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
  return p;
}

// Sobel 3x3 (cv::spatialGradient, reflect-101) + 16-px replicate border of the same image.
__global__ void __launch_bounds__(256) grad_border_kernel(const unsigned char* __restrict__ I, int h, int w,
                                                          short* __restrict__ gx, short* __restrict__ gy,
                                                          unsigned char* __restrict__ E) {
  const int f = blockIdx.z;
  const unsigned char* img = I + (size_t)f * h * w;
  const int we = w + 2 * kBorder, he = h + 2 * kBorder;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x < we && y < he) {
    const int sy = min(max(y - kBorder, 0), h - 1), sx = min(max(x - kBorder, 0), w - 1);
    E[(size_t)f * he * we + (size_t)y * we + x] = img[sy * w + sx];
  }
  if (x < w && y < h) {
    const int ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
    const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
    const int a = img[ym * w + xm], b = img[ym * w + x], c = img[ym * w + xp];
    const int d = img[y * w + xm], e = img[y * w + xp];
    const int g = img[yp * w + xm], hh = img[yp * w + x], i = img[yp * w + xp];
    gx[(size_t)f * h * w + y * w + x] = (short)((c + 2 * e + i) - (a + 2 * d + g));
    gy[(size_t)f * h * w + y * w + x] = (short)((g + 2 * hh + i) - (a + 2 * b + c));
  }
}

// Structure tensor, horizontal running sums: one thread per (frame, row); float running sums in
// OpenCV's order (precomputeStructureTensor) so that large sums round identically.
__global__ void tensor_rows_kernel(const short* __restrict__ gx, const short* __restrict__ gy, int F, int h, int w,
                                   int ws, float* __restrict__ aux /* [F][5][h][ws] */) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= F * h) return;
  const int f = idx / h, i = idx - f * h;
  const short* xr = gx + ((size_t)f * h + i) * w;
  const short* yr = gy + ((size_t)f * h + i) * w;
  float* o = aux + (size_t)f * 5 * h * ws + (size_t)i * ws;
  const size_t plane = (size_t)h * ws;
  float sxx = 0, syy = 0, sxy = 0, sx = 0, sy = 0;
  for (int j = 0; j < kPatch; j++) {
    const int a = xr[j], b = yr[j];
    sxx += (float)(a * a); syy += (float)(b * b); sxy += (float)(a * b); sx += (float)a; sy += (float)b;
  }
  o[0] = sxx; o[plane] = syy; o[2 * plane] = sxy; o[3 * plane] = sx; o[4 * plane] = sy;
  int js = 1;
  for (int j = kPatch; j < w; j++) {
    const int a = xr[j], b = yr[j], a0 = xr[j - kPatch], b0 = yr[j - kPatch];
    sxx += (float)(a * a - a0 * a0); syy += (float)(b * b - b0 * b0); sxy += (float)(a * b - a0 * b0);
    sx += (float)(a - a0); sy += (float)(b - b0);
    if ((j - kPatch + 1) % kStride == 0) {
      o[js] = sxx; o[plane + js] = syy; o[2 * plane + js] = sxy; o[3 * plane + js] = sx; o[4 * plane + js] = sy;
      js++;
    }
  }
}

// Structure tensor, vertical running sums: one thread per (frame, component, patch column).
__global__ void tensor_cols_kernel(const float* __restrict__ aux, int F, int h, int ws, int hs,
                                   float* __restrict__ T /* [F][5][hs][ws] */) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= F * 5 * ws) return;
  const int fc = idx / ws, j = idx - fc * ws;  // fc = frame * 5 + component
  const float* a = aux + (size_t)fc * h * ws + j;
  float* o = T + (size_t)fc * hs * ws + j;
  float s = 0.f;
  for (int i = 0; i < kPatch; i++) s += a[(size_t)i * ws];
  o[0] = s;
  int is = 1;
  for (int i = kPatch; i < h; i++) {
    s += (a[(size_t)i * ws] - a[(size_t)(i - kPatch) * ws]);
    if ((i - kPatch + 1) % kStride == 0) {
      o[(size_t)is * ws] = s;
      is++;
    }
  }
}

// ---- patch inverse search ----------------------------------------------------------------------

struct Bil {
  float w00, w01, w10, w11;
  int off;  // offset of the top-left sample in the bordered image
};

// Integer <-> float conversions without the XU pipe.  I2F / F2I / FRND run at 16 lanes per clock per SM, and the
// patch search converts a 9 x 4 texel window per evaluation: ncu showed that pipe 96 % busy on the slowest SM
// (profiles/r02_dis_ncu.txt).  The classic 2^23 bias does the same conversions exactly on the FMA / ALU pipes.
__device__ __forceinline__ float u8_to_float(unsigned v) { return __int_as_float(0x4B000000u | v) - 8388608.0f; }   // 0 <= v < 2^23
__device__ __forceinline__ float s16_to_float(int v) { return __int_as_float(0x4B400000 + v) - 12582912.0f; }      // |v| < 2^22
// floorf(x) and (int)x for 0 <= x < 2^22: x + 2^23 leaves rint(x) in the mantissa; step back when that rounded up
__device__ __forceinline__ float floor_pos(float x, int& as_int) {
  const float t = x + 8388608.0f;
  const float r = t - 8388608.0f;
  const bool up = r > x;
  as_int = (__float_as_int(t) & 0x7FFFFF) - (up ? 1 : 0);
  return up ? r - 1.0f : r;
}

__device__ __forceinline__ Bil bil_weights(float i, float j, float Ux, float Uy, int w, int h, int we) {
  const float i_lo = kBorder - kPatch + 1.0f, i_hi = kBorder + h - 1.0f;
  const float j_lo = kBorder - kPatch + 1.0f, j_hi = kBorder + w - 1.0f;
  const float iI = fminf(fmaxf(i + Uy + kBorder, i_lo), i_hi);   // >= 9: floor == truncation
  const float jI = fminf(fmaxf(j + Ux + kBorder, j_lo), j_hi);
  int ii, ji;
  const float di = iI - floor_pos(iI, ii), dj = jI - floor_pos(jI, ji);
  Bil b;
  b.w11 = di * dj;
  b.w10 = di * (1 - dj);
  b.w01 = (1 - di) * dj;
  b.w00 = (1 - di) * (1 - dj);
  b.off = ii * we + ji;
  return b;
}

// (a0 + a2) + (a1 + a3) across the 4 lanes of a quad; every lane gets the same bits.
__device__ __forceinline__ float quad_reduce(float v, unsigned mask) {
  const float t = v + __shfl_xor_sync(mask, v, 2);
  return t + __shfl_xor_sync(mask, t, 1);
}

// ---- packed float32x2 multiplies (Blackwell FMUL2): two IEEE round-to-nearest products per instruction.
// A lane owns pixel columns l and l + 4 of a patch, and OpenCV's code does the same operation on both, so every
// product of the inner loop is issued once for the pair -- same bits, fewer instructions in a kernel whose cost is the
// length of one warp's instruction stream.  Only the PRODUCTS are packed: ptxas 12.9 contracts mul.rn.f32x2 +
// add.rn.f32x2 into a fused FFMA2 even under -fmad=false and with explicit .rn (checked with cuobjdump), which would
// change the rounding, so the sums stay scalar FADDs, which it leaves alone.
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 pack2(float x, float y) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void unpack2(f2 a, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ float hsum2(f2 a) {
  float x, y;
  unpack2(a, x, y);
  return __fadd_rn(x, y);
}

// The 9 x (2 + 2) texels of I1 a lane needs for one patch position, converted to float once and kept in registers:
// consecutive Gauss-Newton steps of a patch move it by a fraction of a pixel, so the integer offset -- and with it
// every texel -- usually stays the same and only the four bilinear weights change.
struct I1Cache {
  f2 lo[9];   // (I1[r][l],     I1[r][l + 4])
  f2 hi[9];   // (I1[r][l + 1], I1[r][l + 5])
  int off;    // offset of the cached window in the bordered image, -1 = empty
};

__device__ __forceinline__ void i1_load(I1Cache& c, const unsigned char* __restrict__ I1e, int s1, int off, int l) {
  if (off == c.off) return;
  const unsigned char* p = I1e + off + l;
#pragma unroll
  for (int r = 0; r < 9; r++) {
    c.lo[r] = pack2(u8_to_float(p[r * s1]), u8_to_float(p[r * s1 + 4]));
    c.hi[r] = pack2(u8_to_float(p[r * s1 + 1]), u8_to_float(p[r * s1 + 5]));
  }
  c.off = off;
}

// One 8x8 patch evaluation by a quad: lane l handles pixel columns l and l+4 of every row.
// kind 0: mean-normalised SSD only; kind 1: SSD + gradient-weighted sums (processPatchMeanNorm).
// The template patch (I0 and its gradients) is the same for every evaluation of a patch, so the
// caller converts it to float once (z / gx / gy: .x = column l, .y = column l+4).
template <int KIND>
__device__ __forceinline__ float patch_eval(const f2 (&z)[8], const f2 (&gxv)[8], const f2 (&gyv)[8], const I1Cache& c,
                                            const Bil& b, float xgs, float ygs, float& dUx, float& dUy, unsigned mask) {
  float sd = 0.f, sq = 0.f, mx = 0.f, my = 0.f;
  const f2 w00 = pack2(b.w00, b.w00), w01 = pack2(b.w01, b.w01), w10 = pack2(b.w10, b.w10), w11 = pack2(b.w11, b.w11);
#pragma unroll
  for (int r = 0; r < 8; r++) {
    // (dl, dr) = ((w00 a + w01 a') + w10 b) + w11 b' - z in OpenCV's order: four packed products, scalar sums
    float p0x, p0y, p1x, p1y, p2x, p2y, p3x, p3y, zx, zy;
    unpack2(mul2(w00, c.lo[r]), p0x, p0y);
    unpack2(mul2(w01, c.hi[r]), p1x, p1y);
    unpack2(mul2(w10, c.lo[r + 1]), p2x, p2y);
    unpack2(mul2(w11, c.hi[r + 1]), p3x, p3y);
    unpack2(z[r], zx, zy);
    const float dl = __fsub_rn(__fadd_rn(__fadd_rn(__fadd_rn(p0x, p1x), p2x), p3x), zx);
    const float dr = __fsub_rn(__fadd_rn(__fadd_rn(__fadd_rn(p0y, p1y), p2y), p3y), zy);
    const f2 d = pack2(dl, dr);
    if (KIND == 1) {
      mx = __fadd_rn(mx, hsum2(mul2(d, gxv[r])));
      my = __fadd_rn(my, hsum2(mul2(d, gyv[r])));
    }
    sq = __fadd_rn(sq, hsum2(mul2(d, d)));
    sd = __fadd_rn(sd, __fadd_rn(dl, dr));
  }
  const float sum_diff = quad_reduce(sd, mask);
  const float sum_sq = quad_reduce(sq, mask);
  if (KIND == 1) {
    const float smx = quad_reduce(mx, mask), smy = quad_reduce(my, mask);
    dUx = smx - sum_diff * xgs / 64.f;
    dUy = smy - sum_diff * ygs / 64.f;
  }
  return sum_sq - sum_diff * sum_diff / 64.f;
}

// Anti-diagonal wavefront over the patch columns of OpenCV's 8 fixed stripes: a quad (4 lanes = OpenCV's 4
// SIMD accumulators) owns one patch row of one stripe.  A stripe has ceil(hs / 8) patch rows -- 4 at the
// finest level of a 960x540 working image, 2 / 1 / 1 above it -- so one warp (8 quads) carries `spw` =
// 8 / rows stripes of the same pair at once: every lane works at every level, and the pair needs 8 / spw
// warps instead of 8.  Those warps are spread over `cpp` CTAs of `wpc` warps (grid = P * cpp) so that the
// 148 SMs see one or two warps per scheduler: the search is a chain of dependent patch evaluations and what
// it costs is the issue latency of ONE warp's instruction stream, not throughput.
// STAGED keeps everything the chain touches in shared memory (both images and the sparse flow of the pair;
// the bordered I1 of a 240x135 level is 45 KB), which takes the global load round trips out of every link
// of the chain.  Same arithmetic either way.
#ifndef VSTAB_PS_MIN_CTAS
#define VSTAB_PS_MIN_CTAS 1
#endif
template <bool STAGED>
__global__ void __launch_bounds__(256, VSTAB_PS_MIN_CTAS) patch_search_kernel(Level L, int P, int spw, int cpp) {
  extern __shared__ __align__(16) unsigned char ps_smem[];
  const int pair = blockIdx.x / cpp;
  const int warp_in_pair = (blockIdx.x - pair * cpp) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int slot = lane >> 2;  // quad index 0..7
  const int l = lane & 3;
  const int w = L.w, h = L.h, ws = L.ws, hs = L.hs, we = w + 2 * kBorder;
  const int stripe_sz = (hs + kStripes - 1) / kStripes;  // ceil(hs / 8)  (<= 8 for working sizes <= 960)
  const int sub = slot / stripe_sz;                       // which of the warp's stripes this quad serves
  const int row_in_stripe = slot - sub * stripe_sz;
  const int stripe = warp_in_pair * spw + sub;
  const bool lane_on = sub < spw && stripe < kStripes;
  const int lo = lane_on ? min(stripe * stripe_sz, hs) : 0, hi = lane_on ? min((stripe + 1) * stripe_sz, hs) : 0;
  const int rows = hi - lo;

  const size_t npx = (size_t)h * w;
  const size_t tp = (size_t)hs * ws;
  const unsigned char* I0 = L.I + (size_t)pair * npx;                                     // frame `pair`
  const unsigned char* I1e = L.Iext + (size_t)(pair + 1) * (h + 2 * kBorder) * we;        // frame `pair + 1`
  volatile float* Sx = L.Sx + (size_t)pair * tp;
  volatile float* Sy = L.Sy + (size_t)pair * tp;
  if (STAGED) {
    const int n1 = (h + 2 * kBorder) * we, n0 = h * w;
    const int n1p = (n1 + 15) & ~15, n0p = (n0 + 15) & ~15;
    unsigned char* s1 = ps_smem;
    unsigned char* s0 = ps_smem + n1p;
    float* ssx = reinterpret_cast<float*>(ps_smem + n1p + n0p);
    float* ssy = ssx + tp;
    // 16-byte copies where the frame's offset allows it, bytes otherwise
    if ((reinterpret_cast<uintptr_t>(I1e) & 15) == 0) {
      for (int v = threadIdx.x; v < n1 / 16; v += blockDim.x) reinterpret_cast<uint4*>(s1)[v] = reinterpret_cast<const uint4*>(I1e)[v];
      for (int v = (n1 & ~15) + threadIdx.x; v < n1; v += blockDim.x) s1[v] = I1e[v];
    } else {
      for (int v = threadIdx.x; v < n1; v += blockDim.x) s1[v] = I1e[v];
    }
    if ((reinterpret_cast<uintptr_t>(I0) & 15) == 0) {
      for (int v = threadIdx.x; v < n0 / 16; v += blockDim.x) reinterpret_cast<uint4*>(s0)[v] = reinterpret_cast<const uint4*>(I0)[v];
      for (int v = (n0 & ~15) + threadIdx.x; v < n0; v += blockDim.x) s0[v] = I0[v];
    } else {
      for (int v = threadIdx.x; v < n0; v += blockDim.x) s0[v] = I0[v];
    }
    __syncthreads();
    I1e = s1;
    I0 = s0;
    Sx = ssx;
    Sy = ssy;
  }
  if (warp_in_pair * spw >= kStripes) return;  // whole warp exits together (after the only CTA-wide barrier)

  const short* gx = L.Ix + (size_t)pair * npx;
  const short* gy = L.Iy + (size_t)pair * npx;
  const float* T = L.T + (size_t)pair * 5 * hs * ws;
  const float* Ux = L.Ux + (size_t)pair * npx;
  const float* Uy = L.Uy + (size_t)pair * npx;
  const int inner = kGdIter / 2;  // floor(25 / 2 passes) = 12
  const unsigned qmask = 0xFu << (lane & ~3);

  for (int iter = 0; iter < 2; iter++) {
    const int dir = iter == 0 ? 1 : -1;
    const int nsteps = ws + stripe_sz - 1;  // same for every quad of the warp (a short last stripe idles)
    for (int step = 0; step < nsteps; step++) {
      const int c = step - row_in_stripe;  // column counted in processing order
      const bool active = row_in_stripe < rows && c >= 0 && c < ws;
      if (active) {
        const int is = dir == 1 ? lo + row_in_stripe : hi - 1 - row_in_stripe;
        const int js = dir == 1 ? c : ws - 1 - c;
        const int i = is * kStride, j = js * kStride;
        const int k = is * ws + js;
        float sx, sy;
        if (iter == 0) {
          sx = Ux[(i + kPatch / 2) * w + j + kPatch / 2];
          sy = Uy[(i + kPatch / 2) * w + j + kPatch / 2];
        } else {
          sx = Sx[k];
          sy = Sy[k];
        }
        const unsigned char* I0p = I0 + i * w + j;
        f2 z[8], gxv[8], gyv[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
          z[r] = pack2(u8_to_float(I0p[r * w + l]), u8_to_float(I0p[r * w + l + 4]));
          gxv[r] = pack2(s16_to_float(gx[(i + r) * w + j + l]), s16_to_float(gx[(i + r) * w + j + l + 4]));
          gyv[r] = pack2(s16_to_float(gy[(i + r) * w + j + l]), s16_to_float(gy[(i + r) * w + j + l + 4]));
        }
        float dux, duy;
        I1Cache i1c;
        i1c.off = -1;
        // spatial propagation: own / previous column / previous row of this pass
        Bil b = bil_weights((float)i, (float)j, sx, sy, w, h, we);
        i1_load(i1c, I1e, we, b.off, l);
        float min_ssd = patch_eval<0>(z, gxv, gyv, i1c, b, 0.f, 0.f, dux, duy, qmask);
        if (c > 0) {
          const float cx = Sx[k - dir], cy = Sy[k - dir];
          b = bil_weights((float)i, (float)j, cx, cy, w, h, we);
          i1_load(i1c, I1e, we, b.off, l);
          const float s = patch_eval<0>(z, gxv, gyv, i1c, b, 0.f, 0.f, dux, duy, qmask);
          if (s < min_ssd) { min_ssd = s; sx = cx; sy = cy; }
        }
        if (row_in_stripe > 0) {
          const float cx = Sx[k - dir * ws], cy = Sy[k - dir * ws];
          b = bil_weights((float)i, (float)j, cx, cy, w, h, we);
          i1_load(i1c, I1e, we, b.off, l);
          const float s = patch_eval<0>(z, gxv, gyv, i1c, b, 0.f, 0.f, dux, duy, qmask);
          if (s < min_ssd) { min_ssd = s; sx = cx; sy = cy; }
        }
        float cur_Ux = sx, cur_Uy = sy;
        const float xx = T[k], yy = T[tp + k], xy = T[2 * tp + k];
        const float xgs = T[3 * tp + k], ygs = T[4 * tp + k];
        float detH = xx * yy - xy * xy;
        if (fabsf(detH) < kEps) detH = kEps;
        const float invH11 = yy / detH, invH12 = -xy / detH, invH22 = xx / detH;
        float prev_ssd = kInf;
        for (int t = 0; t < inner; t++) {
          b = bil_weights((float)i, (float)j, cur_Ux, cur_Uy, w, h, we);
          i1_load(i1c, I1e, we, b.off, l);
          const float ssd = patch_eval<1>(z, gxv, gyv, i1c, b, xgs, ygs, dux, duy, qmask);
          const float dx = invH11 * dux + invH12 * duy;
          const float dy = invH12 * dux + invH22 * duy;
          cur_Ux -= dx;
          cur_Uy -= dy;
          if (ssd >= prev_ssd) break;
          prev_ssd = ssd;
        }
        const float ex = cur_Ux - sx, ey = cur_Uy - sy;
        const double nrm = sqrt((double)ex * ex + (double)ey * ey);
        if (nrm <= (double)kPatch) { sx = cur_Ux; sy = cur_Uy; }
        if (l == 0) { Sx[k] = sx; Sy[k] = sy; }
      }
      __syncwarp();
    }
  }
  if (STAGED) {  // the densification reads the sparse flow from global memory
    float* gSx = L.Sx + (size_t)pair * tp;
    float* gSy = L.Sy + (size_t)pair * tp;
    const int s_lo = min(warp_in_pair * spw * stripe_sz, hs), s_hi = min((warp_in_pair + 1) * spw * stripe_sz, hs);
    for (int k = s_lo * ws + lane; k < s_hi * ws; k += 32) {
      gSx[k] = Sx[k];
      gSy[k] = Sy[k];
    }
  }
}

// ---- densification -----------------------------------------------------------------------------

__global__ void __launch_bounds__(256) densify_kernel(Level L, int P) {
  const int pair = blockIdx.z;
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const int i = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int w = L.w, h = L.h, ws = L.ws, hs = L.hs;
  if (j >= w || i >= h) return;
  const size_t npx = (size_t)h * w;
  const unsigned char* I0 = L.I + (size_t)pair * npx;
  const unsigned char* I1 = L.I + (size_t)(pair + 1) * npx;
  const float* Sx = L.Sx + (size_t)pair * hs * ws;
  const float* Sy = L.Sy + (size_t)pair * hs * ws;
  // set of patches overlapping (i, j): OpenCV's incremental start/end bookkeeping in closed form
  const int e_is = min(i / kStride, hs - 1), e_js = min(j / kStride, ws - 1);
  const int s_is = min(i >= kPatch ? (i - kPatch) / kStride + 1 : 0, e_is);
  const int s_js = min(j >= kPatch ? (j - kPatch) / kStride + 1 : 0, e_js);
  float sum_coef = 0.f, sum_Ux = 0.f, sum_Uy = 0.f;
  for (int is = s_is; is <= e_is; is++)
    for (int js = s_js; js <= e_js; js++) {
      const float sx = Sx[is * ws + js], sy = Sy[is * ws + js];
      const float j_m = fminf(fmaxf(j + sx, 0.0f), w - 1.0f - kEps);
      const float i_m = fminf(fmaxf(i + sy, 0.0f), h - 1.0f - kEps);
      const int j_l = (int)j_m, j_u = j_l + 1, i_l = (int)i_m, i_u = i_l + 1;
      const float diff = (j_m - j_l) * (i_m - i_l) * I1[i_u * w + j_u] + (j_u - j_m) * (i_m - i_l) * I1[i_u * w + j_l] +
                         (j_m - j_l) * (i_u - i_m) * I1[i_l * w + j_u] + (j_u - j_m) * (i_u - i_m) * I1[i_l * w + j_l] -
                         I0[i * w + j];
      const float coef = 1 / fmaxf(1.0f, fabsf(diff));
      sum_Ux += coef * sx;
      sum_Uy += coef * sy;
      sum_coef += coef;
    }
  L.Ux[(size_t)pair * npx + i * w + j] = sum_Ux / sum_coef;
  L.Uy[(size_t)pair * npx + i * w + j] = sum_Uy / sum_coef;
}

// ---- variational refinement --------------------------------------------------------------------
//@emul-begin

struct VrBuf {  // all [P][h][w] float
  float *avg, *Iz, *Ix, *Iy, *Ixx, *Ixy, *Iyy, *Ixz, *Iyz;
  float *A11, *A12, *A22, *b1, *b2, *wgt, *tu, *tv, *du, *dv;
};

__device__ __forceinline__ float at_clamped(const float* p, size_t base, int y, int x, int h, int w) {
  return p[base + (size_t)min(max(y, 0), h - 1) * w + min(max(x, 0), w - 1)];
}

#define ADD_OWN_H() if (own_h) { bb1 += own_ux; a11 += wc; bb2 += own_vx; a22 += wc; }
#define ADD_LEFT_H() if (left_h) { bb1 -= left_ux; a11 += wl; bb2 -= left_vx; a22 += wl; }
#define ADD_OWN_V() if (own_v) { bb1 += own_uy; a11 += wc; bb2 += own_vy; a22 += wc; }
#define ADD_UP_V() if (up_v) { bb1 -= up_uy; a11 += wu; bb2 -= up_vy; a22 += wu; }

// All 5 fixed-point iterations x (weights, linear system, 5 red + 5 black SOR half-sweeps, update)
// plus the warp / derivative prologue run in ONE kernel.  A frame pair is owned by one thread-block
// cluster (1, 2, 4 or 8 CTAs depending on the level size, each CTA a band of rows); the phases are
// separated by cluster barriers (barrier.cluster release/acquire orders the global-memory traffic
// between the CTAs of the cluster), so the ~70 dependent launches per level collapse into one.

template <bool STATE = true>
__device__ __forceinline__ void vr_px_warp(const Level& L, const VrBuf& B, int pair, int x, int y) {
  const int w = L.w, h = L.h;
  const size_t k = (size_t)pair * h * w + (size_t)y * w + x;
  const unsigned char* I0 = L.I + (size_t)pair * h * w;
  const unsigned char* I1 = L.I + (size_t)(pair + 1) * h * w;
  const float mx = x + L.Ux[k], my = y + L.Uy[k];
  const int ix = __float2int_rn(mx * 32.f), iy = __float2int_rn(my * 32.f);
  int sx = ix >> 5, sy = iy >> 5;
  const int ax = ix & 31, ay = iy & 31;
  sx = min(max(sx, -32768), 32767);
  sy = min(max(sy, -32768), 32767);
  const float fx1 = ax * (1.f / 32.f), fy1 = ay * (1.f / 32.f), fx0 = 1.f - fx1, fy0 = 1.f - fy1;
  const float w00 = fy0 * fx0, w01 = fy0 * fx1, w10 = fy1 * fx0, w11 = fy1 * fx1;
  const int x0 = min(max(sx, 0), w - 1), x1 = min(max(sx + 1, 0), w - 1);
  const int y0 = min(max(sy, 0), h - 1), y1 = min(max(sy + 1, 0), h - 1);
  const float warped = (float)I1[y0 * w + x0] * w00 + (float)I1[y0 * w + x1] * w01 + (float)I1[y1 * w + x0] * w10 +
                       (float)I1[y1 * w + x1] * w11;
  const float i0 = (float)I0[y * w + x];
  B.avg[k] = 0.5f * (i0 + warped);
  B.Iz[k] = warped - i0;
  if (STATE) {  // the resident variant keeps du / dv in shared memory and forms tu / tv on the fly
    B.tu[k] = L.Ux[k];
    B.tv[k] = L.Uy[k];
    B.du[k] = 0.f;
    B.dv[k] = 0.f;
  }
}

__device__ __forceinline__ void vr_px_deriv1(const Level& L, const VrBuf& B, int pair, int x, int y) {
  const int w = L.w, h = L.h;
  const size_t base = (size_t)pair * h * w, k = base + (size_t)y * w + x;
  B.Ix[k] = at_clamped(B.avg, base, y, x + 1, h, w) - at_clamped(B.avg, base, y, x - 1, h, w);
  B.Iy[k] = at_clamped(B.avg, base, y + 1, x, h, w) - at_clamped(B.avg, base, y - 1, x, h, w);
  B.Ixz[k] = at_clamped(B.Iz, base, y, x + 1, h, w) - at_clamped(B.Iz, base, y, x - 1, h, w);
  B.Iyz[k] = at_clamped(B.Iz, base, y + 1, x, h, w) - at_clamped(B.Iz, base, y - 1, x, h, w);
}

__device__ __forceinline__ void vr_px_deriv2(const Level& L, const VrBuf& B, int pair, int x, int y) {
  const int w = L.w, h = L.h;
  const size_t base = (size_t)pair * h * w, k = base + (size_t)y * w + x;
  B.Ixx[k] = at_clamped(B.Ix, base, y, x + 1, h, w) - at_clamped(B.Ix, base, y, x - 1, h, w);
  B.Ixy[k] = at_clamped(B.Ix, base, y + 1, x, h, w) - at_clamped(B.Ix, base, y - 1, x, h, w);
  B.Iyy[k] = at_clamped(B.Iy, base, y + 1, x, h, w) - at_clamped(B.Iy, base, y - 1, x, h, w);
}

__device__ __forceinline__ void vr_px_weight(const Level& L, const VrBuf& B, int pair, int x, int y) {
  const int w = L.w, h = L.h;
  const size_t base = (size_t)pair * h * w, k = base + (size_t)y * w + x;
  const size_t kr = base + (size_t)y * w + min(x + 1, w - 1), kd = base + (size_t)min(y + 1, h - 1) * w + x;
  const float ux = B.tu[kr] - B.tu[k], vx = B.tv[kr] - B.tv[k], uy = B.tu[kd] - B.tu[k], vy = B.tv[kd] - B.tv[k];
  const float eps2 = kEpsilon * kEpsilon;
  B.wgt[k] = (kAlpha / 2) / sqrtf(ux * ux + vx * vx + uy * uy + vy * vy + eps2);
}

__device__ __forceinline__ void vr_px_system(const Level& L, const VrBuf& B, int pair, int x, int y) {
  const int w = L.w, h = L.h;
  const size_t k = (size_t)pair * h * w + (size_t)y * w + x;
  const float zeta2 = 0.1f * 0.1f, eps2 = kEpsilon * kEpsilon, gamma2 = kGamma / 2, delta2 = kDelta / 2;
  const float ix = B.Ix[k], iy = B.Iy[k], iz = B.Iz[k], ixx = B.Ixx[k], ixy = B.Ixy[k], iyy = B.Iyy[k];
  const float ixz = B.Ixz[k], iyz = B.Iyz[k], dU = B.du[k], dV = B.dv[k];
  float derivNorm = ix * ix + iy * iy + zeta2;
  const float Ik1z = iz + ix * dU + iy * dV;
  float weight = (delta2 / sqrtf(Ik1z * Ik1z / derivNorm + eps2)) / derivNorm;
  float a11 = weight * (ix * ix) + zeta2;
  float a12 = weight * (ix * iy);
  float a22 = weight * (iy * iy) + zeta2;
  float bb1 = -weight * (iz * ix);
  float bb2 = -weight * (iz * iy);
  derivNorm = ixx * ixx + ixy * ixy + zeta2;
  const float derivNorm2 = iyy * iyy + ixy * ixy + zeta2;
  const float Ik1zx = ixz + ixx * dU + ixy * dV;
  const float Ik1zy = iyz + ixy * dU + iyy * dV;
  weight = gamma2 / sqrtf(Ik1zx * Ik1zx / derivNorm + Ik1zy * Ik1zy / derivNorm2 + eps2);
  a11 += weight * (ixx * ixx / derivNorm + ixy * ixy / derivNorm2);
  a12 += weight * (ixx * ixy / derivNorm + ixy * iyy / derivNorm2);
  a22 += weight * (ixy * ixy / derivNorm + iyy * iyy / derivNorm2);
  bb1 += -weight * (ixx * ixz / derivNorm + ixy * iyz / derivNorm2);
  bb2 += -weight * (ixy * ixz / derivNorm + iyy * iyz / derivNorm2);
  const float* u0 = L.Ux;
  const float* v0 = L.Uy;
  const bool red = ((x + y) & 1) == 0;
  const float wc = B.wgt[k];
  const bool own_h = x < w - 1, left_h = x > 0;
  float own_ux = 0.f, own_vx = 0.f, left_ux = 0.f, left_vx = 0.f, wl = 0.f;
  if (own_h) { own_ux = wc * (u0[k + 1] - u0[k]); own_vx = wc * (v0[k + 1] - v0[k]); }
  if (left_h) { wl = B.wgt[k - 1]; left_ux = wl * (u0[k] - u0[k - 1]); left_vx = wl * (v0[k] - v0[k - 1]); }
  if (red) { ADD_OWN_H(); ADD_LEFT_H(); } else { ADD_LEFT_H(); ADD_OWN_H(); }
  const bool own_v = y < h - 1, up_v = y > 0;
  float own_uy = 0.f, own_vy = 0.f, up_uy = 0.f, up_vy = 0.f, wu = 0.f;
  if (own_v) { own_uy = wc * (u0[k + w] - u0[k]); own_vy = wc * (v0[k + w] - v0[k]); }
  if (up_v) { wu = B.wgt[k - w]; up_uy = wu * (u0[k] - u0[k - w]); up_vy = wu * (v0[k] - v0[k - w]); }
  if (red) { ADD_OWN_V(); ADD_UP_V(); } else { ADD_UP_V(); ADD_OWN_V(); }
  B.A11[k] = a11; B.A12[k] = a12; B.A22[k] = a22; B.b1[k] = bb1; B.b2[k] = bb2;
}

__device__ __forceinline__ void vr_px_sor(const Level& L, const VrBuf& B, int pair, int x, int y) {
  const int w = L.w, h = L.h;
  const size_t k = (size_t)pair * h * w + (size_t)y * w + x;
  const float wl = x > 0 ? B.wgt[k - 1] : 0.f, dul = x > 0 ? B.du[k - 1] : 0.f, dvl = x > 0 ? B.dv[k - 1] : 0.f;
  const float dur = x + 1 < w ? B.du[k + 1] : 0.f, dvr = x + 1 < w ? B.dv[k + 1] : 0.f;
  const float wu = y > 0 ? B.wgt[k - w] : 0.f, duu = y > 0 ? B.du[k - w] : 0.f, dvu = y > 0 ? B.dv[k - w] : 0.f;
  const float dud = y + 1 < h ? B.du[k + w] : 0.f, dvd = y + 1 < h ? B.dv[k + w] : 0.f;
  const float wc = B.wgt[k];
  const float sigmaU = wl * dul + wc * dur + wu * duu + wc * dud;
  const float sigmaV = wl * dvl + wc * dvr + wu * dvu + wc * dvd;
  float du = B.du[k], dv = B.dv[k];
  du += kOmega * ((sigmaU + B.b1[k] - dv * B.A12[k]) / B.A11[k] - du);
  dv += kOmega * ((sigmaV + B.b2[k] - du * B.A12[k]) / B.A22[k] - dv);
  B.du[k] = du;
  B.dv[k] = dv;
}

#ifndef VSTAB_HOST_EMUL
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
#define VSTAB_DYNAMIC_SMEM(name) extern __shared__ __align__(16) float name[]
#endif

// ONCHIP: the level is small enough for all 19 planes of a pair to live in the shared memory of ONE CTA
// (19 * h * w * 4 bytes: 39 KB at 30x17, 155 KB at 60x34): 1024 threads, __syncthreads between the
// phases, and every plane access is a shared-memory access -- the phases of these levels are pure
// latency, so that is what they cost.  The plane pointers are rebased so that the per-pixel code,
// which indexes [pair][y][x], lands in the CTA's own copy.
template <bool ONCHIP>
__global__ void __launch_bounds__(ONCHIP ? 1024 : 256, ONCHIP ? 1 : VSTAB_VR_MIN_CTAS) vr_fused_kernel(Level L, VrBuf B, int cluster_size) {
  VSTAB_DYNAMIC_SMEM(vr_smem);
  const int pair = blockIdx.x / cluster_size;
  const int crank = blockIdx.x % cluster_size;
  const int w = L.w, h = L.h;
  if (ONCHIP) {
    float** planes = reinterpret_cast<float**>(&B);
    const size_t px = (size_t)h * w;
#pragma unroll
    for (int k = 0; k < 19; k++) planes[k] = vr_smem + k * px - (size_t)pair * px;
  }
  const int rows_per = (h + cluster_size - 1) / cluster_size;
  const int r0 = min(crank * rows_per, h), r1 = min(r0 + rows_per, h);
  const int npx = (r1 - r0) * w;
  const int half_w = (w + 1) >> 1;
  const int nhalf = (r1 - r0) * half_w;
#define FOR_PX(...)                                               \
  for (int i = threadIdx.x; i < npx; i += blockDim.x) {           \
    const int y = r0 + i / w, x = i - (i / w) * w;                \
    __VA_ARGS__;                                                  \
  }
  FOR_PX(vr_px_warp(L, B, pair, x, y))
  if (ONCHIP) __syncthreads(); else cluster_barrier();
  FOR_PX(vr_px_deriv1(L, B, pair, x, y))
  if (ONCHIP) __syncthreads(); else cluster_barrier();
  FOR_PX(vr_px_deriv2(L, B, pair, x, y))
  for (int it = 0; it < kVrIter; it++) {
    FOR_PX(vr_px_weight(L, B, pair, x, y))
    if (ONCHIP) __syncthreads(); else cluster_barrier();
    FOR_PX(vr_px_system(L, B, pair, x, y))
    __syncthreads();  // the SOR sweeps map pixels to threads differently (checkerboard halves)
    for (int s = 0; s < kSorIter; s++)
      for (int colour = 0; colour < 2; colour++) {
        for (int i = threadIdx.x; i < nhalf; i += blockDim.x) {
          const int y = r0 + i / half_w;
          const int x = 2 * (i - (i / half_w) * half_w) + ((y + colour) & 1);
          if (x < w) vr_px_sor(L, B, pair, x, y);
        }
        if (ONCHIP) __syncthreads(); else cluster_barrier();
      }
    {
      const int last = it == kVrIter - 1;
      FOR_PX({
        const size_t k = (size_t)pair * h * w + (size_t)y * w + x;
        const float u = L.Ux[k] + B.du[k], v = L.Uy[k] + B.dv[k];
        B.tu[k] = u;
        B.tv[k] = v;
        if (last) { L.Ux[k] = u; L.Uy[k] = v; }
      })
    }
    if (ONCHIP) __syncthreads(); else cluster_barrier();
  }
#undef FOR_PX
}

// ---- variational refinement, resident variant ----------------------------------------------------
//
// The SOR state of a band of <= 4096 checkerboard cells per colour lives in the shared memory of ONE 1024-thread CTA:
// du, dv, the smoothness weights and three of the five system coefficients as colour-split planes
// [colour][band row + 1 halo row above and below][column / 2 + 1 pad cell left and right], A11 and A22 of a thread's own
// eight pixels in registers.  A half-sweep then touches no global memory at all (the cluster version above streams
// 19 loads per update through L2 at ~700 cycles each); the derivative planes a pixel needs once per outer iteration
// stay in global memory (written and read by the same thread).  A pair is one CTA (<= 8192 px, e.g. 120x67) or a cluster
// of 2 / 4 / 8 CTAs, each a band of rows; after a half-sweep the first and last row of a band are pushed into the halo
// rows of the neighbouring CTAs through distributed shared memory, and barrier.cluster orders the pushes.
//
// Same per-pixel arithmetic as vr_px_* above, expression by expression (this file is compiled with -fmad=false):
//  * the planes have a compile-time size, so every operand of an update is one LDS at [cell + immediate];
//  * pad cells, halo rows outside the image and dead cells (odd widths) hold +0 in du, dv and the weight plane and are
//    never written: `x > 0 ? w[k-1] * du[k-1] : 0 * 0` and its three siblings become plain loads (0 * 0 = +0 either way);
//  * the quotients of the gradient-constancy term that do not depend on du / dv -- (Ixx Ixx / n1 + Ixy Ixy / n2) and its
//    four siblings -- are formed once per pixel instead of in each of the five outer iterations (11 of the 16 IEEE
//    divisions of a system evaluation), by the same expressions on the same operands.
constexpr int kResCells = 4096;   // checkerboard cells per colour and CTA: 1024 threads x 4 or 512 threads x 8
constexpr int kResColour = 4416;  // floats per colour of a plane: (rows + 2 halo rows) x (cells per row + 2 pad cells) fits
constexpr int kResPlane = 2 * kResColour;
constexpr int kResSmemFloats = 6 * kResPlane;  // du, dv, weight, A12, b1, b2

#ifndef VSTAB_HOST_EMUL
__device__ __forceinline__ void st_cluster(float* local, unsigned rank, float v) {
  unsigned a = (unsigned)__cvta_generic_to_shared(local), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(r), "f"(v) : "memory");
}
#endif

template <bool CLUSTER, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) vr_resident_kernel(Level L, VrBuf B, int cluster_size) {
  constexpr int kResThreads = THREADS, kResPpt = kResCells / THREADS;
  constexpr int DU = 0, DV = kResPlane, WG = 2 * kResPlane, A12 = 3 * kResPlane, B1 = 4 * kResPlane, B2 = 5 * kResPlane;
  VSTAB_DYNAMIC_SMEM(rs);
  const int pair = blockIdx.x / cluster_size;
  const int crank = blockIdx.x % cluster_size;
  const int w = L.w, h = L.h, half_w = (w + 1) >> 1;
  const int rstride = half_w + 2;  // cells per plane row: one pad cell on either side
  const int rows_per = (h + cluster_size - 1) / cluster_size;
  const int r0 = min(crank * rows_per, h), r1 = min(r0 + rows_per, h), rows = r1 - r0;
  const int ncell = rows * half_w;
  const size_t pbase = (size_t)pair * h * w;
  const unsigned pbase32 = (unsigned)pbase;  // 32-bit indices: one IMAD.WIDE per plane access (a chunk of 256 pairs of a 960x540 level is 1.3e8 px)
  const float* __restrict__ u0 = L.Ux + pbase;
  const float* __restrict__ v0 = L.Uy + pbase;
  // the five quotient sums live in planes the streaming kernel uses for other things
  float* const q11 = B.A11;
  float* const q12 = B.A12;
  float* const q22 = B.A22;
  float* const qb1 = B.b1;
  float* const qb2 = B.b2;
#define RES_SYNC() do { if (CLUSTER) cluster_barrier(); else __syncthreads(); } while (0)

  // own cells: cell i = tid + THREADS m of the band -> (row, column pair); the same cell in both colours.
  // packed: bits 0..12 index inside a colour (halo row and pad cell included), 13 parity of the image row, 14 / 15 the
  //         cell is a pixel of the image in colour 0 / 1, 16..22 row of the band, 23 / 24 the row is the first / last of
  //         the band and has a band above / below it (its values are pushed into that band's halo row)
  unsigned own_s[kResPpt];
#pragma unroll
  for (int m = 0; m < kResPpt; m++) {
    const int i = threadIdx.x + kResThreads * m;
    own_s[m] = 0;
    if (i < ncell) {
      const int ly = i / half_w, j = i - ly * half_w, y = r0 + ly;
      const int live0 = 2 * j + (y & 1) < w, live1 = 2 * j + 1 - (y & 1) < w;
      const int first = CLUSTER && ly == 0 && crank > 0, last = CLUSTER && ly == rows - 1 && r1 < h;
      own_s[m] = (unsigned)((ly + 1) * rstride + j + 1) | ((y & 1) << 13) | (live0 << 14) | (live1 << 15) | (ly << 16) | (first << 23) | (last << 24);
    }
  }
  // column pair of a cell from its plane index and band row
#define RES_J(cs_, ly_) ((int)((cs_) & 8191) - ((ly_) + 1) * rstride - 1)

  // phases without per-cell register state walk the band with a rolled loop (cell -> row / column pair by division)
#define RES_FOR_CELLS(...)                                              \
  _Pragma("unroll 1") for (int c = 0; c < 2; c++)                       \
  _Pragma("unroll 1") for (int i = threadIdx.x; i < ncell; i += kResThreads) { \
    const int ly = i / half_w, j = i - ly * half_w, y = r0 + ly;        \
    const int q = (y + c) & 1, x = 2 * j + q;                           \
    if (x < w) {                                                        \
      const int sidx = (ly + 1) * rstride + j + 1;                      \
      const int ci = c * kResColour + sidx, oi = (1 - c) * kResColour + sidx; \
      const unsigned k = (unsigned)(y * w + x);                         \
      (void)ci; (void)oi; (void)k;                                      \
      __VA_ARGS__;                                                      \
    }                                                                   \
  }

  for (int i = threadIdx.x; i < 3 * kResPlane; i += kResThreads) rs[i] = 0.f;  // du, dv, weights: pads and halo rows too
  RES_FOR_CELLS(vr_px_warp<false>(L, B, pair, x, y))
  RES_SYNC();
  RES_FOR_CELLS(vr_px_deriv1(L, B, pair, x, y))
  RES_SYNC();
  RES_FOR_CELLS({
    vr_px_deriv2(L, B, pair, x, y);
    const unsigned gk = pbase32 + k;
    const float zeta2 = 0.1f * 0.1f;
    const float ixx = B.Ixx[gk], ixy = B.Ixy[gk], iyy = B.Iyy[gk], ixz = B.Ixz[gk], iyz = B.Iyz[gk];
    const float derivNorm = ixx * ixx + ixy * ixy + zeta2;
    const float derivNorm2 = iyy * iyy + ixy * ixy + zeta2;
    q11[gk] = (ixx * ixx / derivNorm + ixy * ixy / derivNorm2);
    q12[gk] = (ixx * ixy / derivNorm + ixy * iyy / derivNorm2);
    q22[gk] = (ixy * ixy / derivNorm + iyy * iyy / derivNorm2);
    qb1[gk] = (ixx * ixz / derivNorm + ixy * iyz / derivNorm2);
    qb2[gk] = (ixy * ixz / derivNorm + iyy * iyz / derivNorm2);
  })

  float a11r[2][kResPpt], a22r[2][kResPpt];
  for (int it = 0; it < kVrIter; it++) {
    // ---- smoothness weights (vr_px_weight) from tu = u0 + du, tv = v0 + dv of the cell, its right and its lower neighbour
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
      for (int m = 0; m < kResPpt; m++) {
        if (kResThreads * m >= ncell) break;  // small levels: no thread has a cell in this slot (same answer in every thread)
        unsigned cs_ = own_s[m];
        asm volatile("" : "+r"(cs_));
        if ((cs_ >> (14 + c)) & 1) {
          const int sidx = cs_ & 8191, ly = (cs_ >> 16) & 127, y = r0 + ly, j = RES_J(cs_, ly), q = (y + c) & 1, x = 2 * j + q;
          const int ci = c * kResColour + sidx, oi = (1 - c) * kResColour + sidx;
          const unsigned k = (unsigned)(y * w + x);
          const bool has_r = x + 1 < w, has_d = y + 1 < h;
          const unsigned kr = has_r ? k + 1 : k, kd = has_d ? k + w : k;
          const int sr = has_r ? oi + q : ci, sd = has_d ? oi + rstride : ci;
          float tu = u0[k], tv = v0[k], tur = u0[kr], tvr = v0[kr], tud = u0[kd], tvd = v0[kd];
          if (it > 0) {
            tu = tu + rs[DU + ci]; tv = tv + rs[DV + ci];
            tur = tur + rs[DU + sr]; tvr = tvr + rs[DV + sr];
            tud = tud + rs[DU + sd]; tvd = tvd + rs[DV + sd];
          }
          const float ux = tur - tu, vx = tvr - tv, uy = tud - tu, vy = tvd - tv;
          const float eps2 = kEpsilon * kEpsilon;
          const float wgt = (kAlpha / 2) / sqrtf(ux * ux + vx * vx + uy * uy + vy * vy + eps2);
          rs[WG + ci] = wgt;
          if (CLUSTER && (cs_ & (1u << 24))) st_cluster(&rs[WG + c * kResColour + j + 1], crank + 1, wgt);  // last row -> top halo of the band below
        }
        asm volatile("" ::: "memory");
      }
    RES_SYNC();
    // ---- linear system (vr_px_system): A11, A22 stay in registers, A12, b1, b2 in shared memory
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
      for (int m = 0; m < kResPpt; m++) {
        if (kResThreads * m >= ncell) break;  // small levels: no thread has a cell in this slot (same answer in every thread)
        unsigned cs_ = own_s[m];
        asm volatile("" : "+r"(cs_));  // decode here, not hoisted out of the iteration loop for all cells at once
        if ((cs_ >> (14 + c)) & 1) {
          const int sidx = cs_ & 8191, ly = (cs_ >> 16) & 127, y = r0 + ly, q = (y + c) & 1, x = 2 * RES_J(cs_, ly) + q;
          const int ci = c * kResColour + sidx, oi = (1 - c) * kResColour + sidx;
          const unsigned k = (unsigned)(y * w + x), gk = pbase32 + k;
          const float zeta2 = 0.1f * 0.1f, eps2 = kEpsilon * kEpsilon, gamma2 = kGamma / 2, delta2 = kDelta / 2;
          const float ix = B.Ix[gk], iy = B.Iy[gk], iz = B.Iz[gk], ixx = B.Ixx[gk], ixy = B.Ixy[gk], iyy = B.Iyy[gk];
          const float ixz = B.Ixz[gk], iyz = B.Iyz[gk], dU = rs[DU + ci], dV = rs[DV + ci];
          float derivNorm = ix * ix + iy * iy + zeta2;
          const float Ik1z = iz + ix * dU + iy * dV;
          float weight = (delta2 / sqrtf(Ik1z * Ik1z / derivNorm + eps2)) / derivNorm;
          float a11 = weight * (ix * ix) + zeta2;
          float a12 = weight * (ix * iy);
          float a22 = weight * (iy * iy) + zeta2;
          float bb1 = -weight * (iz * ix);
          float bb2 = -weight * (iz * iy);
          derivNorm = ixx * ixx + ixy * ixy + zeta2;
          const float derivNorm2 = iyy * iyy + ixy * ixy + zeta2;
          const float Ik1zx = ixz + ixx * dU + ixy * dV;
          const float Ik1zy = iyz + ixy * dU + iyy * dV;
          weight = gamma2 / sqrtf(Ik1zx * Ik1zx / derivNorm + Ik1zy * Ik1zy / derivNorm2 + eps2);
          a11 += weight * q11[gk];
          a12 += weight * q12[gk];
          a22 += weight * q22[gk];
          bb1 += -weight * qb1[gk];
          bb2 += -weight * qb2[gk];
          const bool red = ((x + y) & 1) == 0;
          const float wc = rs[WG + ci];
          const bool own_h = x < w - 1, left_h = x > 0;
          float own_ux = 0.f, own_vx = 0.f, left_ux = 0.f, left_vx = 0.f, wl = 0.f;
          if (own_h) { own_ux = wc * (u0[k + 1] - u0[k]); own_vx = wc * (v0[k + 1] - v0[k]); }
          if (left_h) { wl = rs[WG + oi + q - 1]; left_ux = wl * (u0[k] - u0[k - 1]); left_vx = wl * (v0[k] - v0[k - 1]); }
          if (red) { ADD_OWN_H(); ADD_LEFT_H(); } else { ADD_LEFT_H(); ADD_OWN_H(); }
          const bool own_v = y < h - 1, up_v = y > 0;
          float own_uy = 0.f, own_vy = 0.f, up_uy = 0.f, up_vy = 0.f, wu = 0.f;
          if (own_v) { own_uy = wc * (u0[k + w] - u0[k]); own_vy = wc * (v0[k + w] - v0[k]); }
          if (up_v) { wu = rs[WG + oi - rstride]; up_uy = wu * (u0[k] - u0[k - w]); up_vy = wu * (v0[k] - v0[k - w]); }
          if (red) { ADD_OWN_V(); ADD_UP_V(); } else { ADD_UP_V(); ADD_OWN_V(); }
          a11r[c][m] = a11;
          a22r[c][m] = a22;
          rs[A12 + ci] = a12;
          rs[B1 + ci] = bb1;
          rs[B2 + ci] = bb2;
        }
        asm volatile("" ::: "memory");  // one cell at a time: keeps the unrolled cells from piling their loads up in registers
      }
    // the sweeps read what this thread itself just wrote and du / dv, which nobody has touched since the last barrier
    for (int s = 0; s < kSorIter; s++) {
#pragma unroll
      for (int c = 0; c < 2; c++) {
#pragma unroll
        for (int m = 0; m < kResPpt; m++) {
          if (kResThreads * m >= ncell) break;
          unsigned cs_ = own_s[m];
          asm volatile("" : "+r"(cs_));
          if ((cs_ >> (14 + c)) & 1) {
            const int co = c * kResColour, oo = (1 - c) * kResColour;
            float* const p = rs + (cs_ & 8191);           // the cell; planes and colours are immediates from here
            const float* const ph = p + (((cs_ >> 13) + c) & 1);  // other colour, same row: left neighbour at ph[-1], right at ph[0]
            const float* const pu = p - rstride;
            const float* const pd = p + rstride;
            const float wl = ph[WG + oo - 1], dul = ph[DU + oo - 1], dvl = ph[DV + oo - 1];
            const float dur = ph[DU + oo], dvr = ph[DV + oo];
            const float wu = pu[WG + oo], duu = pu[DU + oo], dvu = pu[DV + oo];
            const float dud = pd[DU + oo], dvd = pd[DV + oo];
            const float wc = p[WG + co];
            const float sigmaU = wl * dul + wc * dur + wu * duu + wc * dud;
            const float sigmaV = wl * dvl + wc * dvr + wu * dvu + wc * dvd;
            float du = p[DU + co], dv = p[DV + co];
            du += kOmega * ((sigmaU + p[B1 + co] - dv * p[A12 + co]) / a11r[c][m] - du);
            dv += kOmega * ((sigmaV + p[B2 + co] - du * p[A12 + co]) / a22r[c][m] - dv);
            p[DU + co] = du;
            p[DV + co] = dv;
            if (CLUSTER && (cs_ & (3u << 23))) {  // 2 of a band's ~34 rows
              const int j1 = RES_J(cs_, (cs_ >> 16) & 127) + 1;
              if (cs_ & (1u << 23)) {  // first row of the band -> bottom halo of the band above
                st_cluster(&rs[DU + co + (rows_per + 1) * rstride + j1], crank - 1, du);
                st_cluster(&rs[DV + co + (rows_per + 1) * rstride + j1], crank - 1, dv);
              }
              if (cs_ & (1u << 24)) {  // last row -> top halo of the band below
                st_cluster(&rs[DU + co + j1], crank + 1, du);
                st_cluster(&rs[DV + co + j1], crank + 1, dv);
              }
            }
          }
          asm volatile("" ::: "memory");
        }
        RES_SYNC();
      }
    }
  }
  // total flow of the level = u0 + du (the update phase of the last iteration)
  RES_FOR_CELLS({
    L.Ux[pbase32 + k] = u0[k] + rs[DU + ci];
    L.Uy[pbase32 + k] = v0[k] + rs[DV + ci];
  })
#undef RES_FOR_CELLS
#undef RES_J
#undef RES_SYNC
}

//@emul-end
// ---- flow upsampling ---------------------------------------------------------------------------

// next level = 2 * resize(linear) in the arithmetic of the 1-channel float path of the wheel (IPP):
// fraction in double -> float, fma(S1 - S0, f, S0), horizontal then vertical.
__device__ __forceinline__ void ipp_coord(int d, double scale, int ssize, int& s, float& f) {
  double v = (d + 0.5) * scale - 0.5;
  int si = (int)floor(v);
  v -= si;
  if (si < 0) { v = 0; si = 0; }
  if (si >= ssize - 1) { v = 0; si = ssize - 1; }
  s = si;
  f = (float)v;
}

__global__ void __launch_bounds__(256) upsample_kernel(Level src, Level dst) {
  const int pair = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dst.w || y >= dst.h) return;
  int sx, sy;
  float fx, fy;
  ipp_coord(x, (double)src.w / dst.w, src.w, sx, fx);
  ipp_coord(y, (double)src.h / dst.h, src.h, sy, fy);
  const int sx1 = min(sx + 1, src.w - 1), sy1 = min(sy + 1, src.h - 1);
  const size_t sb = (size_t)pair * src.h * src.w;
  const size_t dk = (size_t)pair * dst.h * dst.w + (size_t)y * dst.w + x;
#define UP1(F)                                                                             \
  {                                                                                        \
    const float p00 = src.F[sb + sy * src.w + sx], p01 = src.F[sb + sy * src.w + sx1];     \
    const float p10 = src.F[sb + sy1 * src.w + sx], p11 = src.F[sb + sy1 * src.w + sx1];   \
    const float r0 = fmaf(p01 - p00, fx, p00), r1 = fmaf(p11 - p10, fx, p10);              \
    dst.F[dk] = fmaf(r1 - r0, fy, r0) * 2.f;                                               \
  }
  UP1(Ux)
  UP1(Uy)
#undef UP1
}

// final x(2^finest) upsampling in OpenCV's own 2-channel float path:
// f = (float)((d+.5)*scale-.5); s = floor(f); f -= s; out = S0*(1-f) + S1*f, horizontal then vertical
__device__ __forceinline__ void cv_coord_x(int d, double scale, int ssize, int& s, float& f) {
  f = (float)((d + 0.5) * scale - 0.5);
  s = (int)floorf(f);
  f -= s;
  if (s < 0) { f = 0; s = 0; }
  if (s >= ssize - 1) { f = 0; s = ssize - 1; }
}

__device__ __forceinline__ void final_sample(const Level& L, int pair, int x, int y, int W, int H, float mul, float& ox,
                                             float& oy) {
  int sx, sy;
  float fx, fy;
  cv_coord_x(x, 1. / ((double)W / L.w), L.w, sx, fx);
  fy = (float)((y + 0.5) * (1. / ((double)H / L.h)) - 0.5);
  sy = (int)floorf(fy);
  fy -= sy;
  const int y0 = min(max(sy, 0), L.h - 1), y1 = min(max(sy + 1, 0), L.h - 1);
  const int sx1 = min(sx + 1, L.w - 1);
  const size_t b = (size_t)pair * L.h * L.w;
#define FIN1(F, out)                                                                     \
  {                                                                                      \
    const float r0 = L.F[b + y0 * L.w + sx] * (1.f - fx) + L.F[b + y0 * L.w + sx1] * fx; \
    const float r1 = L.F[b + y1 * L.w + sx] * (1.f - fx) + L.F[b + y1 * L.w + sx1] * fx; \
    out = (r0 * (1.f - fy) + r1 * fy) * mul;                                             \
  }
  FIN1(Ux, ox)
  FIN1(Uy, oy)
#undef FIN1
}

__global__ void __launch_bounds__(256) final_flow_kernel(Level L, int W, int H, float mul, float* __restrict__ flow) {
  const int pair = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  float ox, oy;
  if (L.w == W && L.h == H) {
    const size_t k = (size_t)pair * H * W + (size_t)y * W + x;
    ox = L.Ux[k] * mul;
    oy = L.Uy[k] * mul;
  } else {
    final_sample(L, pair, x, y, W, H, mul, ox, oy);
  }
  float* o = flow + (((size_t)pair * H + y) * W + x) * 2;
  o[0] = ox;
  o[1] = oy;
}

__global__ void __launch_bounds__(256) final_grid_kernel(Level L, int W, int H, float mul, int step, int gw, int gh,
                                                         float* __restrict__ grid) {
  const int pair = blockIdx.z;
  const int gx = blockIdx.x * 32 + (threadIdx.x & 31);
  const int gy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (gx >= gw || gy >= gh) return;
  float ox, oy;
  if (L.w == W && L.h == H) {
    const size_t k = (size_t)pair * H * W + (size_t)(gy * step) * W + gx * step;
    ox = L.Ux[k] * mul;
    oy = L.Uy[k] * mul;
  } else {
    final_sample(L, pair, gx * step, gy * step, W, H, mul, ox, oy);
  }
  float* o = grid + (((size_t)pair * gh + gy) * gw + gx) * 2;
  o[0] = ox;
  o[1] = oy;
}

__global__ void zero_kernel(float* p, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

int coarsest_scale(int h, int w) {
  const int mx = w > h ? w : h, mn = w < h ? w : h;
  const int a = (int)(log(mx / (4.0 * kPatch)) / log(2.0) + 0.5);
  const int b = mn / kPatch > 0 ? (int)(log((double)(mn / kPatch)) / log(2.0)) : -1;
  return a < b ? a : b;
}

// The pyramid levels cv2.DISOpticalFlow::calc works on, given the finest scale the backend object
// currently holds.  calc() derives the coarsest scale from the frame size; when that falls below the
// finest scale (with the reference's finest scale 2: frames under ~91 px on the long or 32 px on the
// short side) autoSelectPatchSizeAndScales() replaces both -- coarsest = max(0, floor(log2(2 w / (5 *
// patch)))) in float, finest = max(coarsest - 2, 0) -- and LEAVES the new finest scale on the object,
// so the later pairs of a clip (same object, nodes/video_stabilizer_flow.py:312) see another state than
// the first one and may use fewer levels (90x50: levels 2..0 for pair 0, 1..0 afterwards).
struct Scales { int finest, coarsest; };

bool select_scales(int h, int w, int finest_state, Scales* out) {
  int c = coarsest_scale(h, w), f = finest_state;
  if (c < 0) return false;  // cv2: "The input image must have either width or height >= 12"
  if (c < f) {
    c = (int)floorf(log2f((2.0f * (float)w) / (5.0f * (float)kPatch)));
    if (c < 0) c = 0;
    f = c - 2 > 0 ? c - 2 : 0;
  }
  out->finest = f;
  out->coarsest = c;
  return true;
}

const char* check_scales(int h, int w, Scales s) {
  if (s.coarsest >= kMaxLevels) return "vstab_dis_flow: too many pyramid levels";
  int lh = h >> s.finest, lw = w >> s.finest;
  if ((lh - kPatch) / kStride + 1 > 8 * kStripes) return "vstab_dis_flow: more than 64 patch rows at the finest level (working image taller than 960 px, or a narrow portrait frame that cv2 computes at full resolution)";
  for (int i = s.finest; i < s.coarsest; i++) { lh /= 2; lw /= 2; }
  // cv2 itself reads outside the level (and usually crashes) when its coarsest level is smaller than a patch
  if (lh < kPatch || lw < kPatch) return "vstab_dis_flow: the coarsest pyramid level is smaller than one 8x8 patch (cv2 crashes on this size)";
  return nullptr;
}

int dis_run(vstab_handle* hnd, const uint8_t* gray_dev, int n_frames, int height, int width, Scales sc,
            float* flow_dev, float* grid_dev, int grid_step, cudaStream_t st);

}  // namespace

extern "C" int vstab_dis_flow_at(vstab_handle* hnd, const uint8_t* gray_dev, int n_frames, int height, int width,
                                 int first_pair_index, float* flow_dev, float* grid_dev, int grid_step, void* stream) {
  if (!hnd) return vstab_fail(nullptr, VSTAB_ERR_INVALID, "vstab_dis_flow: null handle");
  if (!gray_dev || n_frames < 0 || height <= 0 || width <= 0 || first_pair_index < 0)
    return vstab_fail(hnd, VSTAB_ERR_INVALID, "vstab_dis_flow: bad argument");
  if (grid_dev && grid_step <= 0) return vstab_fail(hnd, VSTAB_ERR_INVALID, "vstab_dis_flow: grid_step must be > 0");
  if (n_frames < 2) return VSTAB_OK;
  Scales first, later;
  if (!select_scales(height, width, kFinest, &first) || !select_scales(height, width, first.finest, &later))
    return vstab_fail(hnd, VSTAB_ERR_UNSUPPORTED, "vstab_dis_flow: frame smaller than 12 px (cv2 raises on this size)");
  if (const char* why = check_scales(height, width, first)) return vstab_fail(hnd, VSTAB_ERR_UNSUPPORTED, why);
  if (const char* why = check_scales(height, width, later)) return vstab_fail(hnd, VSTAB_ERR_UNSUPPORTED, why);
  cudaStream_t st = (cudaStream_t)stream;
  VSTAB_ENTER(hnd);
  const bool split = first_pair_index == 0 && (first.finest != later.finest || first.coarsest != later.coarsest);
  if (!split) return dis_run(hnd, gray_dev, n_frames, height, width, first_pair_index == 0 ? first : later, flow_dev, grid_dev, grid_step, st);
  // pair 0 is the call that found the backend object in its configured state
  int rc = dis_run(hnd, gray_dev, 2, height, width, first, flow_dev, grid_dev, grid_step, st);
  if (rc != VSTAB_OK || n_frames == 2) return rc;
  const size_t npx = (size_t)height * width;
  const size_t gpx = grid_dev ? (size_t)vstab_ceil_div(width, grid_step) * vstab_ceil_div(height, grid_step) : 0;
  return dis_run(hnd, gray_dev + npx, n_frames - 1, height, width, later, flow_dev ? flow_dev + npx * 2 : nullptr,
                 grid_dev ? grid_dev + gpx * 2 : nullptr, grid_step, st);
}

extern "C" int vstab_dis_flow(vstab_handle* hnd, const uint8_t* gray_dev, int n_frames, int height, int width,
                              float* flow_dev, float* grid_dev, int grid_step, void* stream) {
  return vstab_dis_flow_at(hnd, gray_dev, n_frames, height, width, 0, flow_dev, grid_dev, grid_step, stream);
}

namespace {

// The same level restricted to the pairs [p0, ...): per-frame planes start at frame p0, per-pair planes at pair p0.
Level level_from_pair(const Level& L, int p0) {
  Level o = L;
  const size_t npx = (size_t)L.h * L.w, tp = (size_t)L.hs * L.ws;
  o.I += p0 * npx;
  o.Iext += (size_t)p0 * (L.h + 2 * kBorder) * (L.w + 2 * kBorder);
  o.Ix += p0 * npx;
  o.Iy += p0 * npx;
  o.T += p0 * 5 * tp;
  o.Ux += p0 * npx;
  o.Uy += p0 * npx;
  o.Sx += p0 * tp;
  o.Sy += p0 * tp;
  return o;
}

int env_int(const char* name, int fallback) {
  const char* e = getenv(name);
  return e && atoi(e) > 0 ? atoi(e) : fallback;
}

// Cluster size of the refinement launch: a pair is one cluster, each CTA a band of rows with about 1024 pixels
// (4096: 4.66 ms, 1024: 4.44 ms, 256: 4.49 ms for 120 pairs at 960x540).  Round 2 measured the alternatives the
// occupancy calculator suggests -- 120 pairs in clusters of 8 are 960 CTAs of which only 86 clusters are resident
// at once, clusters of 6 or 7 all fit -- and they are SLOWER (6: 5.11 ms, 7: 4.53 ms, 8: 4.26 ms): the launch is
// bound by the throughput of the resident CTAs, not by the second wave (profiles/README.md, round 2).
int vr_cluster_size(int px) {
  const int forced = env_int("VSTAB_VR_CLUSTER", 0);
  if (forced) return forced > 8 ? 8 : forced;
  int cl = 1;
  while (cl < 8 && px > 1024 * cl) cl <<= 1;
  return cl;
}

// Cluster size of the resident refinement (vr_resident_kernel): the smallest of 1 / 2 / 4 / 8 row bands whose
// checkerboard cells fit the CTA (4096 per colour) and whose six planes fit its shared memory; 0 = use the streaming
// cluster kernel.  VSTAB_VR_RESIDENT: 0 = off, 1 = single-CTA levels only, 2 (default) = clusters too.
int vr_resident_cluster(const vstab_handle* hnd, int h, int w) {
  const char* e = getenv("VSTAB_VR_RESIDENT");
  const int mode = e ? atoi(e) : VSTAB_VR_RESIDENT_DEFAULT;
  if (mode <= 0 || h > 1023 || w > 1022 || kResSmemFloats * sizeof(float) > (size_t)hnd->max_smem_optin) return 0;
  const int half_w = (w + 1) / 2;
  for (int cl = 1; cl <= 8; cl <<= 1) {
    const int rows_per = (h + cl - 1) / cl;
    // 127 rows per band: the row of a cell travels in 7 bits of the packed cell word
    if (rows_per <= 127 && rows_per * half_w <= kResCells && (rows_per + 2) * (half_w + 2) <= kResColour) return (cl == 1 || mode >= 2) ? cl : 0;
  }
  return 0;
}

// One pyramid level of one pair group on one stream: patch search, densification, refinement, x2 upsampling.
int dis_level(vstab_handle* hnd, const Level& Lc, const Level* finer, const VrBuf& B, int P, cudaStream_t st,
              cudaEvent_t after_search = nullptr) {
  dim3 gp(vstab_ceil_div(Lc.w, 32), vstab_ceil_div(Lc.h, 8), P);
  {
    const int stripe_sz = (Lc.hs + kStripes - 1) / kStripes;
    // Stripes carried by one warp.  Packing 8 / rows stripes into a warp halves the instruction count of the
    // launch (342 M -> 158 M at the finest level) but NOT its duration: the cost is the length of one warp's
    // instruction stream at ~0.31 IPC, and a warp with 8 busy quads runs every Gauss-Newton loop as long as its
    // slowest quad.  Measured on 120 pairs (scripts/dis_sweep.py): one stripe per warp 4.00 ms, packed 4.25 ms.
    int spw = 1;
    if (env_int("VSTAB_PS_PACK", 0)) spw = stripe_sz <= 8 ? 8 / stripe_sz : 1;
    if (spw < 1) spw = 1;
    const int warps_per_pair = (kStripes + spw - 1) / spw;
    int wpc = env_int("VSTAB_PS_WPC", 8);                   // warps per CTA
    if (wpc > warps_per_pair) wpc = warps_per_pair;
    const int cpp = (warps_per_pair + wpc - 1) / wpc;       // CTAs per pair
    const size_t n1 = (size_t)(Lc.h + 2 * kBorder) * (Lc.w + 2 * kBorder), n0 = (size_t)Lc.h * Lc.w;
    const size_t ps_bytes = ((n1 + 15) & ~(size_t)15) + ((n0 + 15) & ~(size_t)15) + 2 * sizeof(float) * Lc.hs * Lc.ws;
    if (ps_bytes <= (size_t)hnd->max_smem_optin) {
      VSTAB_CUDA(hnd, cudaFuncSetAttribute(patch_search_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps_bytes));
      patch_search_kernel<true><<<P * cpp, wpc * 32, ps_bytes, st>>>(Lc, P, spw, cpp);
    } else {
      patch_search_kernel<false><<<P * cpp, wpc * 32, 0, st>>>(Lc, P, spw, cpp);
    }
  }
  VSTAB_LAUNCH_CHECK(hnd, "patch_search_kernel");
  if (after_search) VSTAB_CUDA(hnd, cudaEventRecord(after_search, st));
  densify_kernel<<<gp, 256, 0, st>>>(Lc, P);
  VSTAB_LAUNCH_CHECK(hnd, "densify_kernel");
  // variational refinement: one cluster-fused launch per level
  {
    const int px = Lc.w * Lc.h;
    const size_t onchip_bytes = (size_t)19 * px * sizeof(float);
    const char* onchip_env = getenv("VSTAB_VR_ONCHIP");
    const bool prefer_resident = (onchip_env ? atoi(onchip_env) == 0 : VSTAB_VR_ONCHIP_DEFAULT == 0) && vr_resident_cluster(hnd, Lc.h, Lc.w) == 1;
    if (onchip_bytes <= (size_t)hnd->max_smem_optin && !prefer_resident) {
      VSTAB_CUDA(hnd, cudaFuncSetAttribute(vr_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)onchip_bytes));
      vr_fused_kernel<true><<<P, 1024, onchip_bytes, st>>>(Lc, B, 1);
      VSTAB_LAUNCH_CHECK(hnd, "vr_fused_kernel");
    } else if (const int rcl = vr_resident_cluster(hnd, Lc.h, Lc.w)) {
      const size_t bytes = (size_t)kResSmemFloats * sizeof(float);
      const int threads = env_int("VSTAB_VR_RESIDENT_THREADS", VSTAB_VR_RESIDENT_THREADS_DEFAULT) == 512 ? 512 : 1024;
      void (*kernel)(Level, VrBuf, int) =
          rcl == 1 ? (threads == 512 ? vr_resident_kernel<false, 512> : vr_resident_kernel<false, 1024>)
                   : (threads == 512 ? vr_resident_kernel<true, 512> : vr_resident_kernel<true, 1024>);
      VSTAB_CUDA(hnd, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(P * rcl));
      cfg.blockDim = dim3(threads);
      cfg.dynamicSmemBytes = bytes;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = rcl;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = rcl > 1 ? 1 : 0;
      VSTAB_CUDA(hnd, cudaLaunchKernelEx(&cfg, kernel, Lc, B, rcl));
      hnd->launches++;
    } else {
      const int cl = vr_cluster_size(px);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(P * cl));
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cl;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      VSTAB_CUDA(hnd, cudaLaunchKernelEx(&cfg, vr_fused_kernel<false>, Lc, B, cl));
      hnd->launches++;
    }
  }
  if (finer) {
    dim3 gu(vstab_ceil_div(finer->w, 32), vstab_ceil_div(finer->h, 8), P);
    upsample_kernel<<<gu, 256, 0, st>>>(Lc, *finer);
    VSTAB_LAUNCH_CHECK(hnd, "upsample_kernel");
  }
  return VSTAB_OK;
}

// Pairs are processed in chunks so the workspace stays bounded (about 6.5 MB per pair at 960x540).
// Inside a chunk the pairs are split into up to 4 groups that run the coarse-to-fine chain on their own
// streams (forked from and joined back into the caller's stream with events): the patch search of one
// group -- one or two warps per scheduler, long tails on slow pairs -- runs under the refinement of another.
int dis_run(vstab_handle* hnd, const uint8_t* gray_dev, int n_frames, int height, int width, Scales sc,
            float* flow_dev, float* grid_dev, int grid_step, cudaStream_t st) {
  const int finest = sc.finest, coarsest = sc.coarsest;

  const int kChunk = 256;  // pairs per pass
  for (int p0 = 0; p0 < n_frames - 1; p0 += kChunk) {
    const int P = (n_frames - 1 - p0) < kChunk ? (n_frames - 1 - p0) : kChunk;
    const int F = P + 1;
    const uint8_t* gray = gray_dev + (size_t)p0 * height * width;

    // ---- carve the workspace ----
    Level L[kMaxLevels];
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    size_t o_I[kMaxLevels], o_E[kMaxLevels], o_gx[kMaxLevels], o_gy[kMaxLevels], o_T[kMaxLevels], o_Ux[kMaxLevels],
        o_Uy[kMaxLevels], o_Sx[kMaxLevels], o_Sy[kMaxLevels];
    int fraction = 1 << finest;
    for (int i = finest; i <= coarsest; i++) {
      if (i == finest) { L[i].h = height / fraction; L[i].w = width / fraction; }
      else { L[i].h = L[i - 1].h / 2; L[i].w = L[i - 1].w / 2; }
      L[i].ws = 1 + (L[i].w - kPatch) / kStride;
      L[i].hs = 1 + (L[i].h - kPatch) / kStride;
      const size_t n = (size_t)L[i].h * L[i].w, t = (size_t)L[i].hs * L[i].ws;
      o_I[i] = take(n * F);
      o_E[i] = take((size_t)(L[i].h + 2 * kBorder) * (L[i].w + 2 * kBorder) * F);
      o_gx[i] = take(n * 2 * F);
      o_gy[i] = take(n * 2 * F);
      o_T[i] = take(t * 5 * 4 * F);
      o_Ux[i] = take(n * 4 * P);
      o_Uy[i] = take(n * 4 * P);
      o_Sx[i] = take(t * 4 * P);
      o_Sy[i] = take(t * 4 * P);
    }
    const size_t nf = (size_t)L[finest].h * L[finest].w;
    const size_t o_aux = take((size_t)L[finest].h * L[finest].ws * 5 * 4 * F);
    size_t o_vr[19];
    for (int k = 0; k < 19; k++) o_vr[k] = take(nf * 4 * P);
    void* wsp = nullptr;
    int rc = vstab_workspace(hnd, off, &wsp);
    if (rc != VSTAB_OK) return rc;
    unsigned char* base = (unsigned char*)wsp;
    for (int i = finest; i <= coarsest; i++) {
      L[i].I = base + o_I[i];
      L[i].Iext = base + o_E[i];
      L[i].Ix = (short*)(base + o_gx[i]);
      L[i].Iy = (short*)(base + o_gy[i]);
      L[i].T = (float*)(base + o_T[i]);
      L[i].Ux = (float*)(base + o_Ux[i]);
      L[i].Uy = (float*)(base + o_Uy[i]);
      L[i].Sx = (float*)(base + o_Sx[i]);
      L[i].Sy = (float*)(base + o_Sy[i]);
    }
    float* aux = (float*)(base + o_aux);
    VrBuf B;
    {
      float** f = (float**)&B;
      for (int k = 0; k < 19; k++) f[k] = (float*)(base + o_vr[k]);
    }

    // ---- per-frame pyramid, gradients, bordered copies, structure tensors ----
    // The pyramid is a chain of INTER_AREA halvings on the caller's stream.  Gradients, bordered copies and structure
    // tensors of a level are needed when the coarse-to-fine chain reaches that level -- the finest one, which is most of
    // this work (~0.2 of 0.3 ms at 120 pairs), more than a millisecond after the coarsest -- so they run on a
    // preparation stream, coarsest level first, under the (latency-bound) coarse levels of the pair groups; a group
    // waits for a level's event right before it searches that level.  VSTAB_DIS_PREP_ASYNC=0: everything on one stream.
    for (int i = finest; i <= coarsest; i++) {
      if (i == finest) rc = vstab_area_u8(hnd, gray, F, height, width, L[i].I, L[i].h, L[i].w, st);
      else rc = vstab_area_u8(hnd, L[i - 1].I, F, L[i - 1].h, L[i - 1].w, L[i].I, L[i].h, L[i].w, st);
      if (rc != VSTAB_OK) return rc;
    }
    const char* prep_env = getenv("VSTAB_DIS_PREP_ASYNC");
    const bool prep_async = !(prep_env && atoi(prep_env) == 0) && coarsest - finest + 1 <= VSTAB_MAX_LEVEL_EVENTS;
    cudaStream_t prep = st;
    if (prep_async) {
      rc = vstab_aux_streams(hnd, VSTAB_MAX_AUX_STREAMS);
      if (rc != VSTAB_OK) return rc;
      if (!hnd->pyramid_event) VSTAB_CUDA(hnd, cudaEventCreateWithFlags(&hnd->pyramid_event, cudaEventDisableTiming));
      while (hnd->n_level_events < VSTAB_MAX_LEVEL_EVENTS) {
        VSTAB_CUDA(hnd, cudaEventCreateWithFlags(&hnd->level_event[hnd->n_level_events], cudaEventDisableTiming));
        hnd->n_level_events++;
      }
      prep = hnd->aux_stream[VSTAB_MAX_AUX_STREAMS - 1];  // the pair groups use aux_stream[0 .. G - 2]
      VSTAB_CUDA(hnd, cudaEventRecord(hnd->pyramid_event, st));
      VSTAB_CUDA(hnd, cudaStreamWaitEvent(prep, hnd->pyramid_event, 0));
    }
    for (int i = coarsest; i >= finest; i--) {
      dim3 ge(vstab_ceil_div(L[i].w + 2 * kBorder, 32), vstab_ceil_div(L[i].h + 2 * kBorder, 8), F);
      grad_border_kernel<<<ge, 256, 0, prep>>>(L[i].I, L[i].h, L[i].w, L[i].Ix, L[i].Iy, L[i].Iext);
      VSTAB_LAUNCH_CHECK(hnd, "grad_border_kernel");
      tensor_rows_kernel<<<vstab_ceil_div(F * L[i].h, 128), 128, 0, prep>>>(L[i].Ix, L[i].Iy, F, L[i].h, L[i].w, L[i].ws, aux);
      VSTAB_LAUNCH_CHECK(hnd, "tensor_rows_kernel");
      tensor_cols_kernel<<<vstab_ceil_div(F * 5 * L[i].ws, 128), 128, 0, prep>>>(aux, F, L[i].h, L[i].ws, L[i].hs, L[i].T);
      VSTAB_LAUNCH_CHECK(hnd, "tensor_cols_kernel");
      if (prep_async) VSTAB_CUDA(hnd, cudaEventRecord(hnd->level_event[i - finest], prep));
    }

    // ---- coarse-to-fine, pair groups on their own streams ----
    {
      const size_t n = (size_t)L[coarsest].h * L[coarsest].w * P;
      zero_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L[coarsest].Ux, n);
      zero_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L[coarsest].Uy, n);
      hnd->launches += 2;
    }
    int G = env_int("VSTAB_DIS_GROUPS", VSTAB_DIS_GROUPS_DEFAULT);
    if (G > VSTAB_MAX_AUX_STREAMS) G = VSTAB_MAX_AUX_STREAMS;  // the last helper stream prepares the levels
    if (G > P / 16) G = P / 16 > 0 ? P / 16 : 1;  // small batches: one chain
    cudaStream_t gs[VSTAB_MAX_AUX_STREAMS + 1];
    gs[0] = st;
    if (G > 1) {
      rc = vstab_aux_streams(hnd, G - 1);
      if (rc != VSTAB_OK) return rc;
      VSTAB_CUDA(hnd, cudaEventRecord(hnd->fork_event, st));
      for (int g = 1; g < G; g++) {
        gs[g] = hnd->aux_stream[g - 1];
        VSTAB_CUDA(hnd, cudaStreamWaitEvent(gs[g], hnd->fork_event, 0));
      }
    }
    // Groups that start together run in lockstep -- patch search under patch search, refinement under refinement --
    // and gain nothing.  So group g + 1 is released only when group g has finished the patch search of level
    // `stagger` levels above the finest: from then on the (latency-bound, 2 warps per scheduler) finest-level search
    // of one group sits under the (throughput-bound) finest-level refinement of the group before it.
    const int stagger = env_int("VSTAB_DIS_STAGGER", VSTAB_DIS_STAGGER_DEFAULT) - 1;  // env 0 / unset-to-0 = off, k = release after PS of level finest + k - 1
    for (int g = 0; g < G; g++) {
      const int a = (int)((long long)P * g / G), b = (int)((long long)P * (g + 1) / G);
      if (b <= a) continue;
      for (int i = coarsest; i >= finest; i--) {
        const Level Lg = level_from_pair(L[i], a);
        Level finer_g;
        if (i > finest) finer_g = level_from_pair(L[i - 1], a);
        VrBuf Bg = B;
        {
          float** f = (float**)&Bg;
          for (int k = 0; k < 19; k++) f[k] += (size_t)a * nf;  // a group's scratch is fixed: groups sit at different levels at the same time
        }
        cudaEvent_t after_search = nullptr;
        if (stagger >= 0 && g + 1 < G && i == (finest + stagger < coarsest ? finest + stagger : coarsest)) after_search = hnd->stagger_event[g];
        if (prep_async) VSTAB_CUDA(hnd, cudaStreamWaitEvent(gs[g], hnd->level_event[i - finest], 0));
        rc = dis_level(hnd, Lg, i > finest ? &finer_g : nullptr, Bg, b - a, gs[g], after_search);
        if (rc != VSTAB_OK) return rc;
        if (after_search) VSTAB_CUDA(hnd, cudaStreamWaitEvent(gs[g + 1], after_search, 0));
      }
    }
    for (int g = 1; g < G; g++) {
      VSTAB_CUDA(hnd, cudaEventRecord(hnd->join_event[g - 1], gs[g]));
      VSTAB_CUDA(hnd, cudaStreamWaitEvent(st, hnd->join_event[g - 1], 0));
    }
    const float mul = (float)(1 << finest);
    if (flow_dev) {
      dim3 gf(vstab_ceil_div(width, 32), vstab_ceil_div(height, 8), P);
      final_flow_kernel<<<gf, 256, 0, st>>>(L[finest], width, height, mul, flow_dev + (size_t)p0 * height * width * 2);
      VSTAB_LAUNCH_CHECK(hnd, "final_flow_kernel");
    }
    if (grid_dev) {
      const int gw = vstab_ceil_div(width, grid_step), gh = vstab_ceil_div(height, grid_step);
      dim3 gg(vstab_ceil_div(gw, 32), vstab_ceil_div(gh, 8), P);
      final_grid_kernel<<<gg, 256, 0, st>>>(L[finest], width, height, mul, grid_step, gw, gh,
                                            grid_dev + (size_t)p0 * gh * gw * 2);
      VSTAB_LAUNCH_CHECK(hnd, "final_grid_kernel");
    }
  }
  return VSTAB_OK;
}

}  // namespace
