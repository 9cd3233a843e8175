// gray.cu -- K1 + K2: RGB float32 -> uint8 luma -> INTER_AREA working image, one read of the
// source.  Also exports the uint8 -> uint8 INTER_AREA used by the DIS pyramid.
//
// Replaces nodes/stabilizer_utils.py:236-242 (_make_gray) and :271-276
// (_make_gray_for_estimation) of the reference.  Arithmetic restated from cv2 4.13 and pinned
// bit-exact in tests:
//   luma      Y = fma(B, .114f, fma(R, .299f, G * .587f))          (cv2's AVX2 RGB2Gray order)
//   u8        trunc(clip(Y * 255, 0, 255))                          (stabilizer_utils.py:242)
//   area x2   (a + b + c + d + 2) >> 2
//   area KxL  rint(float(sum) * (1.f / (K*L)))   half-to-even       (integer scales)
//   general   float32 separable coverage weights, accumulated in cv2's order:
//             buf = sum_k S[sx_k] * alpha_k  (k ascending),  out = beta_0*buf_0 + beta_1*buf_1...
#include "area.cuh"
#include "common.cuh"

#include <math.h>

#include <vector>

namespace {

struct RgbLuma {
  const float* __restrict__ p;  // one frame, [h][w][3]
  int w;
  __device__ __forceinline__ int operator()(int y, int x) const {
    const float* q = p + ((size_t)y * w + x) * 3;
    const float r = __ldg(q), g = __ldg(q + 1), b = __ldg(q + 2);
    const float yv = __fmaf_rn(b, 0.114f, __fmaf_rn(r, 0.299f, __fmul_rn(g, 0.587f)));
    const float s = __fmul_rn(yv, 255.0f);
    return (int)fminf(fmaxf(s, 0.0f), 255.0f);  // truncating cast, like ndarray.astype(uint8)
  }
};

struct U8Plane {
  const unsigned char* __restrict__ p;
  int w;
  __device__ __forceinline__ int operator()(int y, int x) const { return p[(size_t)y * w + x]; }
};

template <class Src>
__global__ void __launch_bounds__(256) area_copy_kernel(Src src0, size_t src_frame_stride, int h,
                                                        int w, unsigned char* __restrict__ dst) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  Src s = src0;
  s.p += src_frame_stride * blockIdx.z;
  dst[((size_t)blockIdx.z * h + y) * w + x] = (unsigned char)s(y, x);
}

template <class Src>
__global__ void __launch_bounds__(256)
    area_int_kernel(Src src0, size_t src_frame_stride, int kx, int ky, float scale, int dh, int dw,
                    unsigned char* __restrict__ dst) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dw || y >= dh) return;
  Src s = src0;
  s.p += src_frame_stride * blockIdx.z;
  int sum = 0;
  for (int j = 0; j < ky; ++j)
    for (int i = 0; i < kx; ++i) sum += s(y * ky + j, x * kx + i);
  int v;
  if (kx == 2 && ky == 2)
    v = (sum + 2) >> 2;
  else
    v = __float2int_rn(__fmul_rn((float)sum, scale));
  dst[((size_t)blockIdx.z * dh + y) * dw + x] = (unsigned char)min(max(v, 0), 255);
}

template <class Src>
__global__ void __launch_bounds__(256)
    area_general_kernel(Src src0, size_t src_frame_stride, const vstab_area_tab xt,
                        const vstab_area_tab yt, int dh, int dw, unsigned char* __restrict__ dst) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dw || y >= dh) return;
  Src s = src0;
  s.p += src_frame_stride * blockIdx.z;
  const int xb = xt.start[x], xe = xt.start[x + 1];
  const int yb = yt.start[y], ye = yt.start[y + 1];
  float sum = 0.f;
  for (int j = yb; j < ye; ++j) {
    const int sy = yt.si[j];
    float buf = 0.f;
    for (int k = xb; k < xe; ++k) buf = __fadd_rn(buf, __fmul_rn((float)s(sy, xt.si[k]), xt.alpha[k]));
    const float t = __fmul_rn(yt.alpha[j], buf);
    sum = (j == yb) ? t : __fadd_rn(sum, t);
  }
  const int v = __float2int_rn(sum);
  dst[((size_t)blockIdx.z * dh + y) * dw + x] = (unsigned char)min(max(v, 0), 255);
}

// cv::computeResizeAreaTab restated (modules/imgproc/src/resize.cpp), per-destination CSR.
void build_area_tab(int ssize, int dsize, double scale, std::vector<int>& start,
                    std::vector<int>& si, std::vector<float>& alpha) {
  start.assign(dsize + 1, 0);
  si.clear();
  alpha.clear();
  for (int dx = 0; dx < dsize; ++dx) {
    start[dx] = (int)si.size();
    const double fsx1 = dx * scale;
    const double fsx2 = fsx1 + scale;
    const double cell = fmin(scale, ssize - fsx1);
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    sx2 = sx2 < ssize - 1 ? sx2 : ssize - 1;
    sx1 = sx1 < sx2 ? sx1 : sx2;
    if (sx1 - fsx1 > 1e-3) {
      si.push_back(sx1 - 1);
      alpha.push_back((float)((sx1 - fsx1) / cell));
    }
    for (int sx = sx1; sx < sx2; ++sx) {
      si.push_back(sx);
      alpha.push_back((float)(1.0 / cell));
    }
    if (fsx2 - sx2 > 1e-3) {
      si.push_back(sx2);
      alpha.push_back((float)(fmin(fmin(fsx2 - sx2, 1.), cell) / cell));
    }
  }
  start[dsize] = (int)si.size();
}

}  // namespace

// Host side of vstab_area_tab: one small device allocation per (ssize, dsize), cached on the handle.
int vstab_area_tab_get(vstab_handle* h, int ssize, int dsize, vstab_area_tab* out) {
  for (int i = 0; i < h->n_area_cache; ++i)
    if (h->area_cache[i].ssize == ssize && h->area_cache[i].dsize == dsize) {
      *out = h->area_cache[i].tab;
      return VSTAB_OK;
    }
  int slot = h->n_area_cache;
  if (slot >= VSTAB_AREA_CACHE) {
    // full (a long-lived process that has seen many frame sizes): recycle the slots round robin.
    // Launches that still read the old table may be in flight on any stream, hence the drain.
    slot = h->area_evict;
    h->area_evict = (h->area_evict + 1) % VSTAB_AREA_CACHE;
    VSTAB_CUDA(h, cudaDeviceSynchronize());
    VSTAB_CUDA(h, cudaFree(h->area_cache[slot].dev));
    h->area_cache[slot].dev = nullptr;
    h->area_cache[slot].ssize = h->area_cache[slot].dsize = -1;
  }
  std::vector<int> start, si;
  std::vector<float> alpha;
  const double inv_scale = (double)dsize / ssize;
  const double scale = 1. / inv_scale;
  build_area_tab(ssize, dsize, scale, start, si, alpha);
  const size_t n = si.size();
  std::vector<unsigned char> host(sizeof(int) * (dsize + 1) + (sizeof(int) + sizeof(float)) * n);
  unsigned char* base = host.data();
  memcpy(base, start.data(), sizeof(int) * (dsize + 1));
  memcpy(base + sizeof(int) * (dsize + 1), si.data(), sizeof(int) * n);
  memcpy(base + sizeof(int) * (dsize + 1 + n), alpha.data(), sizeof(float) * n);
  void* dev = nullptr;
  VSTAB_CUDA(h, cudaMalloc(&dev, host.size()));
  VSTAB_CUDA(h, cudaMemcpy(dev, host.data(), host.size(), cudaMemcpyHostToDevice));
  unsigned char* d = (unsigned char*)dev;
  vstab_area_cache_entry& e = h->area_cache[slot];
  if (slot == h->n_area_cache) h->n_area_cache++;
  e.ssize = ssize;
  e.dsize = dsize;
  e.dev = dev;
  e.tab.start = (const int*)d;
  e.tab.si = (const int*)(d + sizeof(int) * (dsize + 1));
  e.tab.alpha = (const float*)(d + sizeof(int) * (dsize + 1 + n));
  *out = e.tab;
  return VSTAB_OK;
}

namespace {

template <class Src>
int launch_area(vstab_handle* h, Src src, size_t src_frame_stride, int n, int sh, int sw,
                unsigned char* dst, int dh, int dw, cudaStream_t st) {
  dim3 grid(vstab_ceil_div(dw, 32), vstab_ceil_div(dh, 8), n);
  if (dh == sh && dw == sw) {
    area_copy_kernel<Src><<<grid, 256, 0, st>>>(src, src_frame_stride, sh, sw, dst);
    VSTAB_LAUNCH_CHECK(h, "area_copy_kernel");
    return VSTAB_OK;
  }
  if (dw > sw || dh > sh)
    return vstab_fail(h, VSTAB_ERR_UNSUPPORTED, "INTER_AREA upscaling is not on the hot path");
  const double scale_x = 1. / ((double)dw / sw), scale_y = 1. / ((double)dh / sh);
  const int iscale_x = (int)lrint(scale_x), iscale_y = (int)lrint(scale_y);
  const bool fast = fabs(scale_x - iscale_x) < 2.220446049250313e-16 &&
                    fabs(scale_y - iscale_y) < 2.220446049250313e-16;
  if (fast) {
    const float scale = 1.f / (float)(iscale_x * iscale_y);
    area_int_kernel<Src><<<grid, 256, 0, st>>>(src, src_frame_stride, iscale_x, iscale_y, scale, dh, dw, dst);
    VSTAB_LAUNCH_CHECK(h, "area_int_kernel");
    return VSTAB_OK;
  }
  // general path: coverage tables are cached on the handle per (source, destination) length
  vstab_area_tab xt, yt;
  int rc = vstab_area_tab_get(h, sw, dw, &xt);
  if (rc != VSTAB_OK) return rc;
  rc = vstab_area_tab_get(h, sh, dh, &yt);
  if (rc != VSTAB_OK) return rc;
  rc = vstab_area_tab_get(h, sw, dw, &xt);  // a full cache may just have recycled xt's slot for yt
  if (rc != VSTAB_OK) return rc;
  area_general_kernel<Src><<<grid, 256, 0, st>>>(src, src_frame_stride, xt, yt, dh, dw, dst);
  VSTAB_LAUNCH_CHECK(h, "area_general_kernel");
  return VSTAB_OK;
}

}  // namespace

int vstab_area_u8(vstab_handle* h, const unsigned char* src, int n, int sh, int sw,
                  unsigned char* dst, int dh, int dw, cudaStream_t st) {
  U8Plane s{src, sw};
  return launch_area(h, s, (size_t)sh * sw, n, sh, sw, dst, dh, dw, st);
}

extern "C" int vstab_working_size(int width, int height, int* work_w, int* work_h) {
  if (width <= 0 || height <= 0 || !work_w || !work_h) return VSTAB_ERR_INVALID;
  // nodes/stabilizer_utils.py:248-268: cap the longest side at 960, Python round() (half-even)
  const int max_side = 960;
  const int longest = width > height ? width : height;
  *work_w = width;
  *work_h = height;
  if (longest <= max_side) return VSTAB_OK;
  const double scale = max_side / (double)longest;
  int sw = (int)nearbyint(width * scale), sh = (int)nearbyint(height * scale);
  sw = sw < 1 ? 1 : sw;
  sh = sh < 1 ? 1 : sh;
  if (sw >= width || sh >= height) return VSTAB_OK;
  *work_w = sw;
  *work_h = sh;
  return VSTAB_OK;
}

extern "C" int vstab_gray_working(vstab_handle* h, const float* rgb_dev, int n, int height,
                                  int width, uint8_t* gray_dev, int work_h, int work_w,
                                  void* stream) {
  if (!h) return vstab_fail(nullptr, VSTAB_ERR_INVALID, "vstab_gray_working: null handle");
  if (!rgb_dev || !gray_dev || n < 0 || height <= 0 || width <= 0 || work_h <= 0 || work_w <= 0)
    return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_gray_working: bad argument");
  if (n == 0) return VSTAB_OK;
  if (n > 65535) return vstab_fail(h, VSTAB_ERR_UNSUPPORTED, "vstab_gray_working: n > 65535 frames per call");
  VSTAB_ENTER(h);
  RgbLuma s{rgb_dev, width};
  return launch_area(h, s, (size_t)height * width * 3, n, height, width, gray_dev, work_h, work_w,
                     (cudaStream_t)stream);
}
