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

// Range flags of the input adapter (nodes/stabilizer_utils.py:96-147 _to_numpy_frame): a float frame whose max()
// exceeds 1.5 is a 0..255 frame and gets divided by 255.  max() > 1.5 <=> some element > 1.5, unless a NaN is present
// (numpy's max() is then NaN and the test is false): bit 0 = an element > 1.5 was seen, bit 1 = a NaN was seen;
// a frame is rescaled iff its flags == 1.
constexpr unsigned kFlagBig = 1u, kFlagNan = 2u;
__device__ __forceinline__ unsigned range_bits(float r, float g, float b) {
  unsigned f = (r > 1.5f || g > 1.5f || b > 1.5f) ? kFlagBig : 0u;
  if (r != r || g != g || b != b) f |= kFlagNan;
  return f;
}

// DIV255: the frame is a 0..255 frame: every channel is divided by 255 (IEEE, like numpy's `arr /= 255.0`) before the luma.
template <bool DIV255>
struct RgbLumaT {
  const float* __restrict__ p;  // one frame, [h][w][3]
  int w;
  __device__ __forceinline__ int load(int y, int x, unsigned& seen) const {
    const float* q = p + ((size_t)y * w + x) * 3;
    float r = __ldg(q), g = __ldg(q + 1), b = __ldg(q + 2);
    seen |= range_bits(r, g, b);
    if (DIV255) {
      r = __fdiv_rn(r, 255.0f);
      g = __fdiv_rn(g, 255.0f);
      b = __fdiv_rn(b, 255.0f);
    }
    const float yv = __fmaf_rn(b, 0.114f, __fmaf_rn(r, 0.299f, __fmul_rn(g, 0.587f)));
    const float s = __fmul_rn(yv, 255.0f);
    return (int)fminf(fmaxf(s, 0.0f), 255.0f);  // truncating cast, like ndarray.astype(uint8)
  }
  __device__ __forceinline__ int operator()(int y, int x) const {
    unsigned ignored = 0;
    return load(y, x, ignored);
  }
};
using RgbLuma = RgbLumaT<false>;

struct U8Plane {
  const unsigned char* __restrict__ p;
  int w;
  __device__ __forceinline__ int load(int y, int x, unsigned&) const { return p[(size_t)y * w + x]; }
  __device__ __forceinline__ int operator()(int y, int x) const { return p[(size_t)y * w + x]; }
};

// One destination pixel of the three INTER_AREA variants (shared by the grid kernels and by the redo pass of 0..255 frames)
template <class Src>
__device__ __forceinline__ int area_int_px(const Src& s, int kx, int ky, float scale, int y, int x, unsigned& seen) {
  int sum = 0;
  for (int j = 0; j < ky; ++j)
    for (int i = 0; i < kx; ++i) sum += s.load(y * ky + j, x * kx + i, seen);
  const int v = (kx == 2 && ky == 2) ? (sum + 2) >> 2 : __float2int_rn(__fmul_rn((float)sum, scale));
  return min(max(v, 0), 255);
}

template <class Src>
__device__ __forceinline__ int area_general_px(const Src& s, const vstab_area_tab& xt, const vstab_area_tab& yt, int y, int x, unsigned& seen) {
  const int xb = xt.start[x], xe = xt.start[x + 1];
  const int yb = yt.start[y], ye = yt.start[y + 1];
  float sum = 0.f;
  for (int j = yb; j < ye; ++j) {
    const int sy = yt.si[j];
    float buf = 0.f;
    for (int k = xb; k < xe; ++k) buf = __fadd_rn(buf, __fmul_rn((float)s.load(sy, xt.si[k], seen), xt.alpha[k]));
    const float t = __fmul_rn(yt.alpha[j], buf);
    sum = (j == yb) ? t : __fadd_rn(sum, t);
  }
  return min(max(__float2int_rn(sum), 0), 255);
}

template <class Src>
__global__ void __launch_bounds__(256) area_copy_kernel(Src src0, size_t src_frame_stride, int h,
                                                        int w, unsigned char* __restrict__ dst, unsigned* __restrict__ flags) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  Src s = src0;
  s.p += src_frame_stride * blockIdx.z;
  unsigned seen = 0;
  dst[((size_t)blockIdx.z * h + y) * w + x] = (unsigned char)s.load(y, x, seen);
  if (flags && seen) atomicOr(flags + blockIdx.z, seen);  // rare: only frames that are not plain 0..1 content
}

template <class Src>
__global__ void __launch_bounds__(256)
    area_int_kernel(Src src0, size_t src_frame_stride, int kx, int ky, float scale, int dh, int dw,
                    unsigned char* __restrict__ dst, unsigned* __restrict__ flags) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dw || y >= dh) return;
  Src s = src0;
  s.p += src_frame_stride * blockIdx.z;
  unsigned seen = 0;
  dst[((size_t)blockIdx.z * dh + y) * dw + x] = (unsigned char)area_int_px(s, kx, ky, scale, y, x, seen);
  if (flags && seen) atomicOr(flags + blockIdx.z, seen);
}

template <class Src>
__global__ void __launch_bounds__(256)
    area_general_kernel(Src src0, size_t src_frame_stride, const vstab_area_tab xt,
                        const vstab_area_tab yt, int dh, int dw, unsigned char* __restrict__ dst, unsigned* __restrict__ flags) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dw || y >= dh) return;
  Src s = src0;
  s.p += src_frame_stride * blockIdx.z;
  unsigned seen = 0;
  dst[((size_t)blockIdx.z * dh + y) * dw + x] = (unsigned char)area_general_px(s, xt, yt, y, x, seen);
  if (flags && seen) atomicOr(flags + blockIdx.z, seen);
}

// ---- the rare half of the fused input adapter: frames whose flags say "0..255 float content" ----------------------
// Both kernels are launched over ALL frames with a handful of blocks per frame; a block whose frame is ordinary 0..1
// content (flags != 1) returns at once, so the common case costs two near-empty launches and no host round trip.
constexpr int kRedoBlocks = 16, kDivideBlocks = 64;

// mode 0: copy, 1: integer scale, 2: general.  Working image of the rescaled frames, from the RAW values with the
// division by 255 applied on the fly (the in-place division below runs after this kernel, on the same stream).
__global__ void __launch_bounds__(256) gray_redo_kernel(const float* __restrict__ rgb, size_t frame_stride, int sw, int mode, int kx,
                                                        int ky, float scale, vstab_area_tab xt, vstab_area_tab yt, int dh, int dw,
                                                        const unsigned* __restrict__ flags, unsigned char* __restrict__ dst) {
  const int f = blockIdx.y;
  if (flags[f] != kFlagBig) return;
  RgbLumaT<true> s{rgb + frame_stride * f, sw};
  unsigned seen = 0;
  for (int idx = blockIdx.x * 256 + threadIdx.x; idx < dh * dw; idx += kRedoBlocks * 256) {
    const int y = idx / dw, x = idx - y * dw;
    int v;
    if (mode == 0) v = s.load(y, x, seen);
    else if (mode == 1) v = area_int_px(s, kx, ky, scale, y, x, seen);
    else v = area_general_px(s, xt, yt, y, x, seen);
    dst[((size_t)f * dh + y) * dw + x] = (unsigned char)v;
  }
}

__global__ void __launch_bounds__(256) range_divide_kernel(float* __restrict__ rgb, size_t frame_floats, const unsigned* __restrict__ flags) {
  const int f = blockIdx.y;
  if (flags[f] != kFlagBig) return;
  float* p = rgb + frame_floats * f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < frame_floats; i += (size_t)kDivideBlocks * 256) p[i] = __fdiv_rn(p[i], 255.0f);
}

// Adapter without a gray pass (Motion Apply): the flags of every frame in one read of the clip.
__global__ void __launch_bounds__(256) range_flag_kernel(const float* __restrict__ rgb, size_t frame_floats, unsigned* __restrict__ flags) {
  const int f = blockIdx.y;
  const float* p = rgb + frame_floats * f;
  unsigned seen = 0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < frame_floats; i += (size_t)gridDim.x * 256) {
    const float v = __ldg(p + i);
    seen |= range_bits(v, v, v);
  }
  if (seen) atomicOr(flags + f, seen);
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

// flags != nullptr (float RGB sources only): the range flags of every frame are gathered on the way, and `redo` is
// filled with what gray_redo_kernel needs to recompute the working image of a frame that turns out to be 0..255.
struct AreaRedo {
  int mode, kx, ky;
  float scale;
  vstab_area_tab xt, yt;
};

template <class Src>
int launch_area(vstab_handle* h, Src src, size_t src_frame_stride, int n, int sh, int sw,
                unsigned char* dst, int dh, int dw, cudaStream_t st, unsigned* flags = nullptr, AreaRedo* redo = nullptr) {
  dim3 grid(vstab_ceil_div(dw, 32), vstab_ceil_div(dh, 8), n);
  AreaRedo local = {};
  AreaRedo& r = redo ? *redo : local;
  if (dh == sh && dw == sw) {
    r.mode = 0;
    area_copy_kernel<Src><<<grid, 256, 0, st>>>(src, src_frame_stride, sh, sw, dst, flags);
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
    r.mode = 1; r.kx = iscale_x; r.ky = iscale_y; r.scale = scale;
    area_int_kernel<Src><<<grid, 256, 0, st>>>(src, src_frame_stride, iscale_x, iscale_y, scale, dh, dw, dst, flags);
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
  r.mode = 2; r.xt = xt; r.yt = yt;
  area_general_kernel<Src><<<grid, 256, 0, st>>>(src, src_frame_stride, xt, yt, dh, dw, dst, flags);
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

// K1 + K2 with the input adapter's range rule fused into the one read of the source.
extern "C" int vstab_gray_working_adapt(vstab_handle* h, float* rgb_dev, int n, int height, int width, uint8_t* gray_dev,
                                        int work_h, int work_w, uint32_t* flags_dev, void* stream) {
  if (!h) return vstab_fail(nullptr, VSTAB_ERR_INVALID, "vstab_gray_working_adapt: null handle");
  if (!rgb_dev || !gray_dev || !flags_dev || n < 0 || height <= 0 || width <= 0 || work_h <= 0 || work_w <= 0)
    return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_gray_working_adapt: bad argument");
  if (n == 0) return VSTAB_OK;
  if (n > 65535) return vstab_fail(h, VSTAB_ERR_UNSUPPORTED, "vstab_gray_working_adapt: n > 65535 frames per call");
  VSTAB_ENTER(h);
  cudaStream_t st = (cudaStream_t)stream;
  VSTAB_CUDA(h, cudaMemsetAsync(flags_dev, 0, sizeof(uint32_t) * n, st));
  RgbLuma s{rgb_dev, width};
  AreaRedo redo = {};
  const size_t stride = (size_t)height * width * 3;
  int rc = launch_area(h, s, stride, n, height, width, gray_dev, work_h, work_w, st, flags_dev, &redo);
  if (rc != VSTAB_OK) return rc;
  gray_redo_kernel<<<dim3(kRedoBlocks, n), 256, 0, st>>>(rgb_dev, stride, width, redo.mode, redo.kx, redo.ky, redo.scale, redo.xt, redo.yt,
                                                        work_h, work_w, flags_dev, gray_dev);
  VSTAB_LAUNCH_CHECK(h, "gray_redo_kernel");
  range_divide_kernel<<<dim3(kDivideBlocks, n), 256, 0, st>>>(rgb_dev, stride, flags_dev);
  VSTAB_LAUNCH_CHECK(h, "range_divide_kernel");
  return VSTAB_OK;
}

// The adapter alone (callers without an estimation pass: Motion Apply, the legacy inverse): flags, then the division.
extern "C" int vstab_range_normalize(vstab_handle* h, float* rgb_dev, int n, int height, int width, int channels,
                                     uint32_t* flags_dev, void* stream) {
  if (!h) return vstab_fail(nullptr, VSTAB_ERR_INVALID, "vstab_range_normalize: null handle");
  if (!rgb_dev || !flags_dev || n < 0 || height <= 0 || width <= 0 || channels <= 0)
    return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_range_normalize: bad argument");
  if (n == 0) return VSTAB_OK;
  if (n > 65535) return vstab_fail(h, VSTAB_ERR_UNSUPPORTED, "vstab_range_normalize: n > 65535 frames per call");
  VSTAB_ENTER(h);
  cudaStream_t st = (cudaStream_t)stream;
  VSTAB_CUDA(h, cudaMemsetAsync(flags_dev, 0, sizeof(uint32_t) * n, st));
  const size_t stride = (size_t)height * width * channels;
  int bx = (int)((stride + 256 * 16 - 1) / (256 * 16));  // ~16 elements per thread, capped so that n * bx blocks stay cheap
  const int cap = (h->sm_count * 16 + n - 1) / n;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  range_flag_kernel<<<dim3(bx, n), 256, 0, st>>>(rgb_dev, stride, flags_dev);
  VSTAB_LAUNCH_CHECK(h, "range_flag_kernel");
  range_divide_kernel<<<dim3(kDivideBlocks, n), 256, 0, st>>>(rgb_dev, stride, flags_dev);
  VSTAB_LAUNCH_CHECK(h, "range_divide_kernel");
  return VSTAB_OK;
}
