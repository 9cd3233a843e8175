// warp.cu -- fused inverse-map resampler (K10 + K11 + K12 of SURVEY.md section 2.1).
//
// One launch covers the whole batch: for every output pixel it evaluates, per shutter sample,
// the inverse-mapped source coordinate in double, quantises it to 1/32 px (round half to even)
// exactly like OpenCV's fixed-point remap, blends 2x2 (bilinear) or 4x4 (bicubic) taps with
// cv2's float32 weight tables and per-tap BORDER_CONSTANT, accumulates the samples in float32
// in sample order, and writes RGB (float4 stores through a per-warp shared-memory transpose),
// the padding mask and the per-frame padded-pixel count in the same pass.
//
// Replaces nodes/video_stabilizer_flow.py:560-588, nodes/video_stabilizer_classic.py:491-519,
// nodes/motion_apply.py:75-122 and :137-202 of the reference.
//
// Two kernels share the arithmetic:
//
//  * warp_stream_kernel (single sample, 16-byte aligned source rows: every node path except motion blur).
//    A plan pass computes per 64x16 output tile where its source box lies and what kind of tile it is.
//    The resampling kernel is persistent (4 CTAs of 8 warps per SM); tiles are handed out by a global
//    ticket so that the tiles in flight are neighbours in the frame and their halos hit in L2.  Each CTA
//    runs a two-stage shared-memory pipeline fed by the TMA engine: one cp.async.bulk.tensor per tile
//    (3-D descriptor over [frame][row][3*col], zero fill outside the frame, L2 evict_last) completing on an
//    mbarrier; the warp that finishes a tile last issues the load that reuses its stage.  Interior tiles
//    (footprint >= 1 px inside the source) run a branch-free bilinear or bicubic path, edge tiles (footprint
//    leaves the frame but fits the box) add cv2's border rules and the coverage test, everything else
//    (partial tiles, oversized footprints) takes the general path.  Output leaves as streaming
//    (evict-first) 16-byte stores through a per-warp shared-memory transpose.
//
//  * warp_fused_kernel<interp, blur> (motion blur with up to 33 samples, unaligned sources, VSTAB_STAGE_GLOBAL):
//    one CTA of 256 threads per 64x32 output tile; the bounding box of the four projected corners over all
//    samples is staged by one cp.async.bulk per source row on one mbarrier.  Taps outside the staged box fall
//    back to a global load, so staging is a pure optimisation and never changes results.  The blur
//    instantiation (2..33 samples) walks pixel-outer / sample-inner and keeps the 2x2 / 4x4 texel footprint in
//    registers across shutter samples (blur_tile); VSTAB_STAGE_GLOBAL keeps every tile on general_tile_body,
//    the reference those paths are tested against.
//
// In both, warp w owns rows w and w+8 of a 16-row group and lane l owns columns l and l+32 (stride-1
// lanes => conflict-free shared-memory gathers with a 3-word stride).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int TW = 64;
#ifndef VSTAB_WARP_GROUPS
#define VSTAB_WARP_GROUPS 2
#endif
constexpr int GROUPS = VSTAB_WARP_GROUPS;  // row groups of 16 output rows per CTA
constexpr int TH = 16 * GROUPS;
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
#ifndef VSTAB_WARP_MIN_CTAS
#define VSTAB_WARP_MIN_CTAS 4
#endif
constexpr int MIN_CTAS = VSTAB_WARP_MIN_CTAS;
constexpr int MAX_SAMPLES = 33;
constexpr int MINV_SLOTS = 34;  // 34*9*8 bytes keeps everything behind it 16-byte aligned
constexpr int SCRATCH_FLOATS_PER_WARP = 2 * TW * 3;  // two rows of RGB

struct WarpParams {
  const float* __restrict__ src;
  const float* __restrict__ fwd;
  float* __restrict__ dst;
  float* __restrict__ mask;
  unsigned int* __restrict__ pad_count;
  int n, sh, sw, oh, ow;
  int samples;
  int mask_rule;
  const unsigned char* __restrict__ rules;  // VSTAB_MASK_RULE_AUTO: rule of every (frame, sample), else nullptr
  int stage_mode;
  int stage_capacity;  // floats available for the staged source tile
  int vec_store;       // ow % 4 == 0 && dst 16B aligned
  int vec_load;        // sw % 4 == 0 && src 16B aligned
  int vec_mask;        // ow % 4 == 0 && mask 16B aligned
  float border[3];
};

__constant__ float c_cubic_tab[32][4];

// Mask rule of one (frame, shutter sample): the launch-wide rule, or what mask_rule_auto_kernel chose for it.
__device__ __forceinline__ int sample_rule(const WarpParams& p, int frame_idx, int s) {
  return p.rules ? (int)p.rules[(size_t)frame_idx * p.samples + s] : p.mask_rule;
}

// VSTAB_MASK_RULE_AUTO: which rule cv2 itself applies to one warpPerspective(ones, M, INTER_NEAREST) call.
// The wheel's IPP path handles the destination in `stripes` horizontal stripes -- min(cv2.getNumThreads(),
// ceil(W' H' / 2^14)) of them, stripe s = rows [(s H' + S/2) / S, ((s+1) H' + S/2) / S) -- and when the source frame
// (forward-mapped quad of its pixel centres) misses ONE of them entirely the whole call falls back to OpenCV's own code,
// which rounds the source coordinate before the range test (Rule C).  Found black-box and pinned against the live wheel
// in tests/test_oracle_resample.py (random warps + boundary sweeps, 0 mismatches); oracle: resample_np.auto_rule.
__device__ __forceinline__ int clip_halfplane(double (&px)[10], double (&py)[10], int n, double a, double b, double c) {
  double qx[10], qy[10];
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const int j = i + 1 == n ? 0 : i + 1;
    const double fp = a * px[i] + b * py[i] + c, fq = a * px[j] + b * py[j] + c;
    if (fp >= 0.0 && m < 10) { qx[m] = px[i]; qy[m] = py[i]; ++m; }
    if ((fp >= 0.0) != (fq >= 0.0) && m < 10) {
      const double t = fp / (fp - fq);
      qx[m] = px[i] + t * (px[j] - px[i]);
      qy[m] = py[i] + t * (py[j] - py[i]);
      ++m;
    }
  }
  for (int i = 0; i < m; ++i) { px[i] = qx[i]; py[i] = qy[i]; }
  return m;
}

__global__ void mask_rule_auto_kernel(const float* __restrict__ fwd, int count, int sh, int sw, int oh, int ow, int threads,
                                      unsigned char* __restrict__ rules) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float* m = fwd + (size_t)i * 9;
  const double cx[4] = {0.0, (double)(sw - 1), (double)(sw - 1), 0.0}, cy[4] = {0.0, 0.0, (double)(sh - 1), (double)(sh - 1)};
  double qx[4], qy[4];
  for (int k = 0; k < 4; ++k) {
    const double X = (double)m[0] * cx[k] + (double)m[1] * cy[k] + (double)m[2];
    const double Y = (double)m[3] * cx[k] + (double)m[4] * cy[k] + (double)m[5];
    const double W = (double)m[6] * cx[k] + (double)m[7] * cy[k] + (double)m[8];
    qx[k] = X / W;
    qy[k] = Y / W;
  }
  long long area_stripes = ((long long)ow * oh + 16383) / 16384;
  int S = threads < area_stripes ? threads : (int)area_stripes;
  if (S < 1) S = 1;
  int rule = VSTAB_MASK_RULE_P;
  for (int s = 0; s < S && rule == VSTAB_MASK_RULE_P; ++s) {
    const int r0 = (int)(((long long)s * oh + S / 2) / S), r1 = (int)(((long long)(s + 1) * oh + S / 2) / S);
    if (r1 <= r0) continue;
    double px[10], py[10];
    for (int k = 0; k < 4; ++k) { px[k] = qx[k]; py[k] = qy[k]; }
    int n = 4;
    n = clip_halfplane(px, py, n, 1.0, 0.0, 0.0);                       // x >= 0
    if (n) n = clip_halfplane(px, py, n, -1.0, 0.0, (double)(ow - 1));  // x <= W' - 1
    if (n) n = clip_halfplane(px, py, n, 0.0, 1.0, -(double)r0);        // y >= r0
    if (n) n = clip_halfplane(px, py, n, 0.0, -1.0, (double)(r1 - 1));  // y <= r1 - 1
    if (n == 0) rule = VSTAB_MASK_RULE_C;
  }
  rules[i] = (unsigned char)rule;
}

struct StagedTile {
  const float* smem;  // staged box, row pitch = pitch floats
  int x0, y0, x1, y1; // inclusive box in source pixel coordinates (already clipped to the image)
  int pitch;
  bool active;
};

// RGB of one tap (3 consecutive floats).
__device__ __forceinline__ void fetch_rgb(const WarpParams& p, const float* __restrict__ frame,
                                          const StagedTile& t, int y, int x, float& r, float& g,
                                          float& b) {
  if ((unsigned)x >= (unsigned)p.sw || (unsigned)y >= (unsigned)p.sh) {
    r = p.border[0];
    g = p.border[1];
    b = p.border[2];
    return;
  }
  if (t.active && x >= t.x0 && x <= t.x1 && y >= t.y0 && y <= t.y1) {
    const float* s = t.smem + (y - t.y0) * t.pitch + (x - t.x0) * 3;
    r = s[0];
    g = s[1];
    b = s[2];
    return;
  }
  const float* s = frame + ((size_t)y * p.sw + x) * 3;
  r = __ldg(s);
  g = __ldg(s + 1);
  b = __ldg(s + 2);
}

// ---- bulk asynchronous copies (TMA engine, 1-D form) completing on an mbarrier ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
#ifndef VSTAB_WAIT_SLEEP
#define VSTAB_WAIT_SLEEP 1
#endif
__device__ __forceinline__ void mbar_wait_spin(unsigned long long* bar, unsigned phase) {
  unsigned done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(phase)
                 : "memory");
  } while (!done);
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// wait with a suspend-time hint: the warp sleeps in hardware instead of burning issue slots
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long* bar, unsigned phase) {
  unsigned done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(phase), "r"(20000u)
                 : "memory");
  } while (!done);
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  if (VSTAB_WAIT_SLEEP) mbar_wait_sleep(bar, phase); else mbar_wait_spin(bar, phase);
}


// Interior tile, single sample, bilinear: every tap is inside the staged box and every pixel is
// covered, so the loop body is coordinates -> 12 shared-memory loads -> blend, nothing else.
// Column / row products of the inverse matrix are hoisted (2 columns x 2 rows per thread).
template <bool AFFINE, bool VEC, bool WAIT, int G, int PITCH>
__device__ __forceinline__ void interior_tile(const double* __restrict__ s_minv, const float* __restrict__ tile0, int pitch_rt,
                                              int tx0, int ty0, int warp, int lane, int tid, int ow,
                                              float* __restrict__ scratch, float* __restrict__ dst_tile,
                                              float* __restrict__ mask_tile, int vec_mask, unsigned long long* bar,
                                              bool bulk_pending) {
  const int pitch = PITCH ? PITCH : pitch_rt;
  const double m0 = s_minv[0], m1 = s_minv[1], m2 = s_minv[2], m3 = s_minv[3], m4 = s_minv[4], m5 = s_minv[5];
  const double m6 = s_minv[6], m7 = s_minv[7], m8 = s_minv[8];
  const double sc_affine = (m8 != 0.0) ? __ddiv_rn(32.0, m8) : 0.0;
  double ax[2], ay[2], aw[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double dx = (double)(tx0 + lane + 32 * k);
    ax[k] = __dmul_rn(m0, dx);
    ay[k] = __dmul_rn(m3, dx);
    if (!AFFINE) aw[k] = __dmul_rn(m6, dx);
  }
  int ixs[4 * G], iys[4 * G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const double dy = (double)(ty0 + g * 16 + warp + NWARPS * rr);
      const double bx = __dmul_rn(m1, dy), by = __dmul_rn(m4, dy);
      double bw = 0.0;
      if (!AFFINE) bw = __dmul_rn(m7, dy);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const double X = __dadd_rn(__dadd_rn(ax[cc], bx), m2);
        const double Y = __dadd_rn(__dadd_rn(ay[cc], by), m5);
        double sc = sc_affine;
        if (!AFFINE) {
          const double W = __dadd_rn(__dadd_rn(aw[cc], bw), m8);
          sc = (W != 0.0) ? __ddiv_rn(32.0, W) : 0.0;
        }
        // |coordinates| < 2^15 here, so cv2's INT_MIN/INT_MAX and short saturation are no-ops
        ixs[g * 4 + rr * 2 + cc] = __double2int_rn(__dmul_rn(X, sc));
        iys[g * 4 + rr * 2 + cc] = __double2int_rn(__dmul_rn(Y, sc));
      }
    }
  }
  if (mask_tile) {  // fully covered tile: mask = 0; 64 x 16*GROUPS floats = GROUPS 16-byte stores per thread
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float* mrow = mask_tile + (size_t)(g * 16 + (tid >> 4)) * ow + (tid & 15) * 4;
      if (vec_mask) {
        *reinterpret_cast<float4*>(mrow) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        mrow[0] = 0.f; mrow[1] = 0.f; mrow[2] = 0.f; mrow[3] = 0.f;
      }
    }
  }
  // the source box has been streaming into shared memory meanwhile
  if (WAIT) {
    if (bulk_pending) mbar_wait(bar, 0); else __syncthreads();
  }
#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int ix = ixs[g * 4 + rr * 2 + cc], iy = iys[g * 4 + rr * 2 + cc];
        const float fx1 = (float)(ix & 31) * 0.03125f, fy1 = (float)(iy & 31) * 0.03125f;
        const float fx0 = 1.0f - fx1, fy0 = 1.0f - fy1;
        const float w00 = __fmul_rn(fy0, fx0), w01 = __fmul_rn(fy0, fx1);
        const float w10 = __fmul_rn(fy1, fx0), w11 = __fmul_rn(fy1, fx1);
        const float* s0 = tile0 + (iy >> 5) * pitch + (ix >> 5) * 3;
        const float* s1 = s0 + pitch;
        float v[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
          v[ch] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s0[ch], w00), __fmul_rn(s0[3 + ch], w01)), __fmul_rn(s1[ch], w10)),
                            __fmul_rn(s1[3 + ch], w11));
        if (VEC) {
          float* o = scratch + rr * (TW * 3) + (lane + cc * 32) * 3;
          o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
        } else {
          float* o = dst_tile + ((size_t)(g * 16 + warp + rr * NWARPS) * ow + lane + cc * 32) * 3;
          o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
        }
      }
    }
    if (VEC) {
      __syncwarp();
#pragma unroll
      for (int q3 = 0; q3 < 3; ++q3) {
        const int q = lane + q3 * 32;
        const int rr = q >= 48 ? 1 : 0, qi = q - rr * 48;
        const float4 val = *reinterpret_cast<const float4*>(scratch + rr * (TW * 3) + qi * 4);
        *reinterpret_cast<float4*>(dst_tile + (size_t)(g * 16 + warp + rr * NWARPS) * ow * 3 + qi * 4) = val;
      }
      __syncwarp();
    }
  }
}

// remapBicubic with the whole 4x4 footprint inside the source, in cv2's own order of additions (found black-box,
// oracle/resample_np.py warp_np(cubic_rows=True), bit-exact against the wheel): the four products of a tap row are
// summed left to right, then the row sums are added top to bottom -- not one running sum over the 16 taps.
// t: first tap (row k1, column k2, channel ch at t[k1 * row_stride + k2 * 3 + ch]).
__device__ __forceinline__ void cubic_rows_sum(const float* __restrict__ t, int row_stride, const float (&wx)[4], const float (&wy)[4],
                                               float (&v)[3]) {
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
    const float* r = t + k1 * row_stride;
    const float w0 = __fmul_rn(wy[k1], wx[0]), w1 = __fmul_rn(wy[k1], wx[1]);
    const float w2 = __fmul_rn(wy[k1], wx[2]), w3 = __fmul_rn(wy[k1], wx[3]);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float row = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[ch], w0), __fmul_rn(r[3 + ch], w1)), __fmul_rn(r[6 + ch], w2)),
                                  __fmul_rn(r[9 + ch], w3));
      v[ch] = k1 == 0 ? row : __fadd_rn(v[ch], row);
    }
  }
}

// General tile: any sample count, both interpolations, border handling, coverage mask and padded
// count.  Uses the staged box where it can and global loads elsewhere; only warp-level
// synchronisation inside, so it serves both the one-tile-per-CTA kernel and the streaming kernel.
template <int INTERP, int G>
__device__ __forceinline__ void general_tile_body(const WarpParams& p, const float* __restrict__ frame, int frame_idx, int tx0, int ty0,
                                             const double* __restrict__ s_minv, const StagedTile& tile,
                                             const float* __restrict__ s_cubic, float* __restrict__ scratch, int warp, int lane) {
  const int S = p.samples;
  const float* __restrict__ t_smem = tile.smem;
  // ---- per-pixel resampling ------------------------------------------------------------------
  // Sample-outer loop: the per-sample row/column products of the inverse matrix are hoisted out of
  // the 4 pixels a thread owns; the float32 accumulation per pixel is still in sample order.
  const float fS = (float)S;
  unsigned int padded = 0;
  const double x_hi = (double)(p.sw - 1), y_hi = (double)(p.sh - 1);
  const bool t_on = tile.active;
  const int t_x0 = tile.x0, t_y0 = tile.y0, t_x1 = tile.x1, t_y1 = tile.y1, t_pitch = tile.pitch;
  const double dxs[2] = {(double)(tx0 + lane), (double)(tx0 + lane + 32)};

  for (int g = 0; g < G; ++g) {  // row groups of 16 output rows
  const int tyg = ty0 + g * 16;
  float acc[4][3];
  int cover[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    acc[q][0] = acc[q][1] = acc[q][2] = 0.f;
    cover[q] = 0;
  }
  const double dys[2] = {(double)(tyg + warp), (double)(tyg + warp + NWARPS)};
  for (int s = 0; s < S; ++s) {
    const double* m = s_minv + s * 9;
    const double m2 = m[2], m5 = m[5], m8 = m[8];
    double ax[2], ay[2], aw[2], bx[2], by[2], bw[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      ax[k] = __dmul_rn(m[0], dxs[k]);
      ay[k] = __dmul_rn(m[3], dxs[k]);
      aw[k] = __dmul_rn(m[6], dxs[k]);
      bx[k] = __dmul_rn(m[1], dys[k]);
      by[k] = __dmul_rn(m[4], dys[k]);
      bw[k] = __dmul_rn(m[7], dys[k]);
    }
    // affine maps (last row 0 0 w): W is the same double for every pixel => one division per CTA
    const bool affine = (m[6] == 0.0) && (m[7] == 0.0);
    const double sc_affine = (m8 != 0.0) ? __ddiv_rn(32.0, m8) : 0.0;
    const int rule = sample_rule(p, frame_idx, s);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int rr = q >> 1, cc = q & 1;
      const int oy = tyg + warp + rr * NWARPS, ox = tx0 + lane + cc * 32;
      if (oy >= p.oh || ox >= p.ow) continue;
      const double X = __dadd_rn(__dadd_rn(ax[cc], bx[rr]), m2);
      const double Y = __dadd_rn(__dadd_rn(ay[cc], by[rr]), m5);
      double W, sc;
      if (affine) {
        W = m8;
        sc = sc_affine;
      } else {
        W = __dadd_rn(__dadd_rn(aw[cc], bw[rr]), m8);
        sc = (W != 0.0) ? __ddiv_rn(32.0, W) : 0.0;
      }
      // -- coverage (INTER_NEAREST ones warp): fl(X/W), fl(Y/W) inside the closed source rectangle.
      //    X * fl(1/W) decides everything that is not within 1e-6 px of a boundary; only those
      //    pixels pay for the exact divisions.
      {
        bool ok;
        const double rw = sc * 0.03125;  // fl(32/W)/32 == fl(1/W): scaling by 2^-5 is exact
        const double qx = X * rw, qy = Y * rw;
        const double band = 1e-6;
        if (rule == VSTAB_MASK_RULE_P && qx > band && qx < x_hi - band && qy > band && qy < y_hi - band) {
          ok = true;
        } else if (rule == VSTAB_MASK_RULE_P && (qx < -band || qx > x_hi + band || qy < -band || qy > y_hi + band)) {
          ok = false;
        } else {
          double cxs = __ddiv_rn(X, W), cys = __ddiv_rn(Y, W);
          if (rule == VSTAB_MASK_RULE_C) {
            cxs = rint(cxs);
            cys = rint(cys);
          }
          ok = (cxs >= 0.0) && (cxs <= x_hi) && (cys >= 0.0) && (cys <= y_hi);
        }
        cover[q] += ok ? 1 : 0;
      }
      // -- 1/32-px fixed-point source coordinate
      double fx = __dmul_rn(X, sc), fy = __dmul_rn(Y, sc);
      fx = fmax(-2147483648.0, fmin(2147483647.0, fx));
      fy = fmax(-2147483648.0, fmin(2147483647.0, fy));
      const int ix = __double2int_rn(fx), iy = __double2int_rn(fy);
      int sx = ix >> 5, sy = iy >> 5;
      sx = max(-32768, min(32767, sx));
      sy = max(-32768, min(32767, sy));
      const int fxi = ix & 31, fyi = iy & 31;
      float vr, vg, vb;
      if (INTERP == VSTAB_INTERP_BILINEAR) {
        const float fx1 = (float)fxi * 0.03125f, fy1 = (float)fyi * 0.03125f;
        const float fx0 = 1.0f - fx1, fy0 = 1.0f - fy1;
        const float w00 = __fmul_rn(fy0, fx0), w01 = __fmul_rn(fy0, fx1);
        const float w10 = __fmul_rn(fy1, fx0), w11 = __fmul_rn(fy1, fx1);
        float r0, g0, b0, r1, g1, b1, r2, g2, b2, r3, g3, b3;
        bool have = true;
        if (t_on && sx >= t_x0 && sx < t_x1 && sy >= t_y0 && sy < t_y1) {
          // whole 2x2 footprint inside the staged (in-image) box: 12 conflict-free LDS
          const float* s0 = t_smem + (sy - t_y0) * t_pitch + (sx - t_x0) * 3;
          const float* s1 = s0 + t_pitch;
          r0 = s0[0]; g0 = s0[1]; b0 = s0[2]; r1 = s0[3]; g1 = s0[4]; b1 = s0[5];
          r2 = s1[0]; g2 = s1[1]; b2 = s1[2]; r3 = s1[3]; g3 = s1[4]; b3 = s1[5];
        } else if (sx >= p.sw || sx + 1 < 0 || sy >= p.sh || sy + 1 < 0) {
          // remapBilinear: footprint entirely outside the source => the border colour itself
          have = false;
          r0 = g0 = b0 = r1 = g1 = b1 = r2 = g2 = b2 = r3 = g3 = b3 = 0.f;
        } else {
          fetch_rgb(p, frame, tile, sy, sx, r0, g0, b0);
          fetch_rgb(p, frame, tile, sy, sx + 1, r1, g1, b1);
          fetch_rgb(p, frame, tile, sy + 1, sx, r2, g2, b2);
          fetch_rgb(p, frame, tile, sy + 1, sx + 1, r3, g3, b3);
        }
        if (have) {
          vr = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0, w00), __fmul_rn(r1, w01)), __fmul_rn(r2, w10)), __fmul_rn(r3, w11));
          vg = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(g0, w00), __fmul_rn(g1, w01)), __fmul_rn(g2, w10)), __fmul_rn(g3, w11));
          vb = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(b0, w00), __fmul_rn(b1, w01)), __fmul_rn(b2, w10)), __fmul_rn(b3, w11));
        } else {
          vr = p.border[0];
          vg = p.border[1];
          vb = p.border[2];
        }
      } else {
        float wx[4], wy[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          wx[k] = s_cubic[fxi * 4 + k];
          wy[k] = s_cubic[fyi * 4 + k];
        }
        const int bxs = sx - 1, bys = sy - 1;
        if (bxs >= 0 && bxs < p.sw - 3 && bys >= 0 && bys < p.sh - 3) {
          // footprint inside the source: cv2's row sums (cubic_rows_sum)
          float v3[3];
          if (t_on && bxs >= t_x0 && bxs + 3 <= t_x1 && bys >= t_y0 && bys + 3 <= t_y1) {
            cubic_rows_sum(t_smem + (bys - t_y0) * t_pitch + (bxs - t_x0) * 3, t_pitch, wx, wy, v3);
          } else {
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) {
              float row[3] = {0.f, 0.f, 0.f};
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) {
                const float w = __fmul_rn(wy[k1], wx[k2]);
                float r, g, b;
                fetch_rgb(p, frame, tile, bys + k1, bxs + k2, r, g, b);
                row[0] = k2 == 0 ? __fmul_rn(r, w) : __fadd_rn(row[0], __fmul_rn(r, w));
                row[1] = k2 == 0 ? __fmul_rn(g, w) : __fadd_rn(row[1], __fmul_rn(g, w));
                row[2] = k2 == 0 ? __fmul_rn(b, w) : __fadd_rn(row[2], __fmul_rn(b, w));
              }
#pragma unroll
              for (int ch = 0; ch < 3; ++ch) v3[ch] = k1 == 0 ? row[ch] : __fadd_rn(v3[ch], row[ch]);
            }
          }
          vr = v3[0];
          vg = v3[1];
          vb = v3[2];
        } else {
          // remapBicubic border branch: cv + sum over in-range taps of (S - cv) * w
          vr = p.border[0];
          vg = p.border[1];
          vb = p.border[2];
          if (!(bxs >= p.sw || bxs + 3 < 0 || bys >= p.sh || bys + 3 < 0)) {
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) {
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) {
                const int yy = bys + k1, xx = bxs + k2;
                if ((unsigned)xx < (unsigned)p.sw && (unsigned)yy < (unsigned)p.sh) {
                  const float w = __fmul_rn(wy[k1], wx[k2]);
                  float r, g, b;
                  fetch_rgb(p, frame, tile, yy, xx, r, g, b);
                  vr = __fadd_rn(vr, __fmul_rn(__fsub_rn(r, p.border[0]), w));
                  vg = __fadd_rn(vg, __fmul_rn(__fsub_rn(g, p.border[1]), w));
                  vb = __fadd_rn(vb, __fmul_rn(__fsub_rn(b, p.border[2]), w));
                }
              }
            }
          }
        }
      }
      if (S == 1) {  // _warp_with_matrices copies the warp; only the blur path accumulates from zero
        acc[q][0] = vr;
        acc[q][1] = vg;
        acc[q][2] = vb;
      } else {
        acc[q][0] = __fadd_rn(acc[q][0], vr);
        acc[q][1] = __fadd_rn(acc[q][1], vg);
        acc[q][2] = __fadd_rn(acc[q][2], vb);
      }
    }
  }

#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int rr = q >> 1, cc = q & 1;
    const int oy = tyg + warp + rr * NWARPS, ox = tx0 + lane + cc * 32;
    if (oy < p.oh && ox < p.ow) {
      if (S > 1) {
        acc[q][0] = __fdiv_rn(acc[q][0], fS);
        acc[q][1] = __fdiv_rn(acc[q][1], fS);
        acc[q][2] = __fdiv_rn(acc[q][2], fS);
      }
      // padding mask: 1 - (coverage > .5) for one sample, 1 - count/S for blur; <1e-3 -> 0
      float mval;
      if (S == 1) {
        mval = cover[q] ? 0.0f : 1.0f;
      } else {
        mval = __fsub_rn(1.0f, __fdiv_rn((float)cover[q], fS));
        if (mval < 1e-3f) mval = 0.0f;
      }
      if (mval > 1e-3f) ++padded;
      if (p.mask) p.mask[((size_t)frame_idx * p.oh + oy) * p.ow + ox] = mval;
      if (!p.vec_store) {
        float* d = p.dst + (((size_t)frame_idx * p.oh + oy) * p.ow + ox) * 3;
        d[0] = acc[q][0];
        d[1] = acc[q][1];
        d[2] = acc[q][2];
      }
    }
    if (p.vec_store) {
      float* sc = scratch + rr * (TW * 3) + (lane + cc * 32) * 3;
      sc[0] = acc[q][0];
      sc[1] = acc[q][1];
      sc[2] = acc[q][2];
    }
  }

  if (p.vec_store) {
    __syncwarp();
    // 2 rows x 192 floats = 96 float4 per warp, 3 per lane, fully coalesced 16-byte stores.
    const int valid_floats = (min(tx0 + TW, p.ow) - tx0) * 3;  // multiple of 4 when ow % 4 == 0
#pragma unroll
    for (int q3 = 0; q3 < 3; ++q3) {
      const int q = lane + q3 * 32;
      const int rr = q / 48, qi = q - rr * 48;
      const int oy = tyg + warp + rr * NWARPS;
      if (oy < p.oh && qi * 4 < valid_floats) {
        const float4 v = *reinterpret_cast<const float4*>(scratch + rr * (TW * 3) + qi * 4);
        float* d = p.dst + (((size_t)frame_idx * p.oh + oy) * p.ow + tx0) * 3 + qi * 4;
        *reinterpret_cast<float4*>(d) = v;
      }
    }
  }
  if (p.vec_store) __syncwarp();
  }  // row groups

  if (p.pad_count) {
    for (int o = 16; o > 0; o >>= 1) padded += __shfl_down_sync(0xffffffffu, padded, o);
    if (lane == 0 && padded) atomicAdd(p.pad_count + frame_idx, padded);
  }
}

template <int INTERP, int G>
__device__ __forceinline__ void general_tile(const WarpParams& p, const float* __restrict__ frame, int frame_idx, int tx0, int ty0,
                                             const double* __restrict__ s_minv, const StagedTile& tile,
                                             const float* __restrict__ s_cubic, float* __restrict__ scratch, int warp, int lane) {
  general_tile_body<INTERP, G>(p, frame, frame_idx, tx0, ty0, s_minv, tile, s_cubic, scratch, warp, lane);
}
// out-of-line copy for the streaming kernel: border tiles are the minority there and must not
// dictate the register allocation of the tile loop
template <int INTERP, int G>
__device__ __noinline__ void general_tile_call(const WarpParams& p, const float* __restrict__ frame, int frame_idx, int tx0, int ty0,
                                               const double* __restrict__ s_minv, const StagedTile& tile,
                                               const float* __restrict__ s_cubic, float* __restrict__ scratch, int warp, int lane) {
  general_tile_body<INTERP, G>(p, frame, frame_idx, tx0, ty0, s_minv, tile, s_cubic, scratch, warp, lane);
}

// Motion-blur tiles (2..33 shutter samples).  Consecutive shutter samples of a pixel land a fraction of a
// pixel apart, so the 2x2 / 4x4 texel footprint stays in registers across samples and is re-read only when
// its integer origin moves: pixel-outer, sample-inner.  What is left per sample is the coordinate (double),
// the weights and the blend -- the same operations in the same order as general_tile_body, so the same bits.
//   EDGE = false: the footprint of every sample lies >= 1 px inside the source and inside the staged box:
//                 every tap is in shared memory, every sample covers the pixel (mask 0, nothing padded).
//   EDGE = true:  everything else (frame border, partial tiles, unstaged footprints).  The reload applies
//                 cv2's border rules once per footprint: bilinear substitutes the border colour per tap and
//                 returns it outright when the footprint is outside; bicubic near the border is
//                 cv + sum over in-range taps of (S - cv) * w, cached as (S - cv) with zeros for the taps
//                 that are skipped (x + 0*w == x for the sums that occur: they start at +0 or cv).
template <int INTERP, bool EDGE>
__device__ __forceinline__ int blur_reload(const WarpParams& p, const float* __restrict__ frame, const StagedTile& tile, int sx, int sy,
                                           float* __restrict__ t) {
  constexpr int NT = (INTERP == VSTAB_INTERP_BILINEAR) ? 2 : 4;
  constexpr int OFF = (INTERP == VSTAB_INTERP_BILINEAR) ? 0 : 1;
  const int bx = sx - OFF, by = sy - OFF;
  const bool staged = tile.active && bx >= tile.x0 && bx + NT - 1 <= tile.x1 && by >= tile.y0 && by + NT - 1 <= tile.y1;
  if (!EDGE || staged) {  // the staged box is clipped to the source, so these taps are all in the image
    const float* s0 = tile.smem + (by - tile.y0) * tile.pitch + (bx - tile.x0) * 3;
#pragma unroll
    for (int k1 = 0; k1 < NT; ++k1) {
#pragma unroll
      for (int j = 0; j < NT * 3; ++j) t[k1 * NT * 3 + j] = s0[k1 * tile.pitch + j];
    }
    return 0;
  }
  if (INTERP == VSTAB_INTERP_BILINEAR) {
    if (sx >= p.sw || sx + 1 < 0 || sy >= p.sh || sy + 1 < 0) return 2;  // remapBilinear: the border colour itself
#pragma unroll
    for (int k = 0; k < 4; ++k) fetch_rgb(p, frame, tile, sy + (k >> 1), sx + (k & 1), t[k * 3], t[k * 3 + 1], t[k * 3 + 2]);
    return 0;
  }
  const bool in_image = bx >= 0 && bx < p.sw - 3 && by >= 0 && by < p.sh - 3;
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
      const int yy = by + k1, xx = bx + k2;
      float r = 0.f, g = 0.f, b = 0.f;
      if (in_image) {
        fetch_rgb(p, frame, tile, yy, xx, r, g, b);
      } else if ((unsigned)xx < (unsigned)p.sw && (unsigned)yy < (unsigned)p.sh) {
        fetch_rgb(p, frame, tile, yy, xx, r, g, b);
        r = __fsub_rn(r, p.border[0]);
        g = __fsub_rn(g, p.border[1]);
        b = __fsub_rn(b, p.border[2]);
      }
      t[(k1 * 4 + k2) * 3 + 0] = r;
      t[(k1 * 4 + k2) * 3 + 1] = g;
      t[(k1 * 4 + k2) * 3 + 2] = b;
    }
  }
  return in_image ? 0 : 1;  // 1: remapBicubic's border branch, the sum starts at the border colour
}

template <int INTERP, bool AFFINE, bool EDGE, int G>
__device__ __forceinline__ void blur_tile(const WarpParams& p, const float* __restrict__ frame, int frame_idx, int tx0, int ty0,
                                          const double* __restrict__ s_minv, const double* __restrict__ s_pk, const StagedTile& tile,
                                          const float* __restrict__ s_cubic, float* __restrict__ scratch, int warp, int lane, int tid) {
  constexpr int NT = (INTERP == VSTAB_INTERP_BILINEAR) ? 2 : 4;   // taps per axis
  const int S = p.samples;
  const float fS = (float)S;
  if (!EDGE && p.mask) {  // every sample covers every pixel: mask = 1 - S/S = 0
    float* mask_tile = p.mask + ((size_t)frame_idx * p.oh + ty0) * p.ow + tx0;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float* mrow = mask_tile + (size_t)(g * 16 + (tid >> 4)) * p.ow + (tid & 15) * 4;
      if (p.vec_mask) {
        *reinterpret_cast<float4*>(mrow) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        mrow[0] = 0.f; mrow[1] = 0.f; mrow[2] = 0.f; mrow[3] = 0.f;
      }
    }
  }
  unsigned int padded = 0;
  const double x_hi = (double)(p.sw - 1), y_hi = (double)(p.sh - 1);
#pragma unroll 1
  for (int g = 0; g < G; ++g) {
    const int tyg = ty0 + g * 16;
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      const int rr = q >> 1, cc = q & 1;
      const int oy = tyg + warp + rr * NWARPS, ox = tx0 + lane + cc * 32;
      const bool live = !EDGE || (oy < p.oh && ox < p.ow);
      float ar = 0.f, ag = 0.f, ab = 0.f;
      if (live) {
        const double dx = (double)ox, dy = (double)oy;
        float t[NT * NT * 3];
        int csx = INT_MIN, csy = INT_MIN, mode = 0, cover = 0;
#pragma unroll 1
        for (int s = 0; s < S; ++s) {
          double X, Y, W, sc;
          if (AFFINE) {
            const double2* pk = reinterpret_cast<const double2*>(s_pk + s * 8);  // m0 m1 | m2 m3 | m4 m5 | 32/m8 m8
            const double2 a = pk[0], b = pk[1], c = pk[2], d = pk[3];
            X = __dadd_rn(__dadd_rn(__dmul_rn(a.x, dx), __dmul_rn(a.y, dy)), b.x);
            Y = __dadd_rn(__dadd_rn(__dmul_rn(b.y, dx), __dmul_rn(c.x, dy)), c.y);
            sc = d.x;
            W = d.y;
          } else {
            const double* m = s_minv + s * 9;
            X = __dadd_rn(__dadd_rn(__dmul_rn(m[0], dx), __dmul_rn(m[1], dy)), m[2]);
            Y = __dadd_rn(__dadd_rn(__dmul_rn(m[3], dx), __dmul_rn(m[4], dy)), m[5]);
            W = __dadd_rn(__dadd_rn(__dmul_rn(m[6], dx), __dmul_rn(m[7], dy)), m[8]);
            sc = (W != 0.0) ? __ddiv_rn(32.0, W) : 0.0;
          }
          double fx = __dmul_rn(X, sc), fy = __dmul_rn(Y, sc);
          if (EDGE) {
            // coverage (INTER_NEAREST ones warp), as in general_tile_body: X * fl(1/W) decides everything
            // that is not within 1e-6 px of a boundary; only those pixels pay for the exact divisions
            bool ok;
            const double rw = sc * 0.03125;
            const double qx = X * rw, qy = Y * rw;
            const double band = 1e-6;
            const int rule = sample_rule(p, frame_idx, s);
            if (rule == VSTAB_MASK_RULE_P && qx > band && qx < x_hi - band && qy > band && qy < y_hi - band) {
              ok = true;
            } else if (rule == VSTAB_MASK_RULE_P && (qx < -band || qx > x_hi + band || qy < -band || qy > y_hi + band)) {
              ok = false;
            } else {
              double cxs = __ddiv_rn(X, W), cys = __ddiv_rn(Y, W);
              if (rule == VSTAB_MASK_RULE_C) {
                cxs = rint(cxs);
                cys = rint(cys);
              }
              ok = (cxs >= 0.0) && (cxs <= x_hi) && (cys >= 0.0) && (cys <= y_hi);
            }
            cover += ok ? 1 : 0;
            fx = fmax(-2147483648.0, fmin(2147483647.0, fx));
            fy = fmax(-2147483648.0, fmin(2147483647.0, fy));
          }
          // interior tiles: |coordinates| < 2^15, cv2's INT_MIN/INT_MAX and short saturation are no-ops
          const int ix = __double2int_rn(fx), iy = __double2int_rn(fy);
          int sx = ix >> 5, sy = iy >> 5;
          if (EDGE) {
            sx = max(-32768, min(32767, sx));
            sy = max(-32768, min(32767, sy));
          }
          if (sx != csx || sy != csy) {
            mode = blur_reload<INTERP, EDGE>(p, frame, tile, sx, sy, t);
            csx = sx;
            csy = sy;
          }
          float vr, vg, vb;
          if (INTERP == VSTAB_INTERP_BILINEAR) {
            const float fx1 = (float)(ix & 31) * 0.03125f, fy1 = (float)(iy & 31) * 0.03125f;
            const float fx0 = 1.0f - fx1, fy0 = 1.0f - fy1;
            const float w00 = __fmul_rn(fy0, fx0), w01 = __fmul_rn(fy0, fx1);
            const float w10 = __fmul_rn(fy1, fx0), w11 = __fmul_rn(fy1, fx1);
            vr = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[0], w00), __fmul_rn(t[3], w01)), __fmul_rn(t[6], w10)), __fmul_rn(t[9], w11));
            vg = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[1], w00), __fmul_rn(t[4], w01)), __fmul_rn(t[7], w10)), __fmul_rn(t[10], w11));
            vb = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[2], w00), __fmul_rn(t[5], w01)), __fmul_rn(t[8], w10)), __fmul_rn(t[11], w11));
            if (EDGE && mode == 2) {
              vr = p.border[0];
              vg = p.border[1];
              vb = p.border[2];
            }
          } else {
            const float4 wx4 = *reinterpret_cast<const float4*>(s_cubic + (ix & 31) * 4);
            const float4 wy4 = *reinterpret_cast<const float4*>(s_cubic + (iy & 31) * 4);
            const float wx[4] = {wx4.x, wx4.y, wx4.z, wx4.w}, wy[4] = {wy4.x, wy4.y, wy4.z, wy4.w};
            if (EDGE && mode) {
              // remapBicubic's border branch: one running sum that starts at the border colour (t holds S - cv, zeros for skipped taps)
              vr = p.border[0];
              vg = p.border[1];
              vb = p.border[2];
#pragma unroll
              for (int k1 = 0; k1 < 4; ++k1) {
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                  const float w = __fmul_rn(wy[k1], wx[k2]);
                  vr = __fadd_rn(vr, __fmul_rn(t[(k1 * 4 + k2) * 3 + 0], w));
                  vg = __fadd_rn(vg, __fmul_rn(t[(k1 * 4 + k2) * 3 + 1], w));
                  vb = __fadd_rn(vb, __fmul_rn(t[(k1 * 4 + k2) * 3 + 2], w));
                }
              }
            } else {
              float v3[3];
              cubic_rows_sum(t, 12, wx, wy, v3);
              vr = v3[0];
              vg = v3[1];
              vb = v3[2];
            }
          }
          ar = __fadd_rn(ar, vr);
          ag = __fadd_rn(ag, vg);
          ab = __fadd_rn(ab, vb);
        }
        ar = __fdiv_rn(ar, fS);
        ag = __fdiv_rn(ag, fS);
        ab = __fdiv_rn(ab, fS);
        if (EDGE) {  // padding mask: 1 - count/S, < 1e-3 -> 0
          float mval = __fsub_rn(1.0f, __fdiv_rn((float)cover, fS));
          if (mval < 1e-3f) mval = 0.0f;
          if (mval > 1e-3f) ++padded;
          if (p.mask) p.mask[((size_t)frame_idx * p.oh + oy) * p.ow + ox] = mval;
        }
        if (!p.vec_store) {
          float* d = p.dst + (((size_t)frame_idx * p.oh + oy) * p.ow + ox) * 3;
          d[0] = ar;
          d[1] = ag;
          d[2] = ab;
        }
      }
      if (p.vec_store) {
        float* o = scratch + rr * (TW * 3) + (lane + cc * 32) * 3;
        o[0] = ar;
        o[1] = ag;
        o[2] = ab;
      }
    }
    if (p.vec_store) {
      __syncwarp();
      // 2 rows x 192 floats = 96 float4 per warp, 3 per lane, fully coalesced 16-byte stores
      const int valid_floats = (min(tx0 + TW, p.ow) - tx0) * 3;  // multiple of 4 when ow % 4 == 0
#pragma unroll
      for (int q3 = 0; q3 < 3; ++q3) {
        const int q = lane + q3 * 32;
        const int rr = q >= 48 ? 1 : 0, qi = q - rr * 48;
        const int oy = tyg + warp + rr * NWARPS;
        if (oy < p.oh && qi * 4 < valid_floats) {
          const float4 val = *reinterpret_cast<const float4*>(scratch + rr * (TW * 3) + qi * 4);
          *reinterpret_cast<float4*>(p.dst + (((size_t)frame_idx * p.oh + oy) * p.ow + tx0) * 3 + qi * 4) = val;
        }
      }
      __syncwarp();
    }
  }
  if (EDGE && p.pad_count) {
    for (int o = 16; o > 0; o >>= 1) padded += __shfl_down_sync(0xffffffffu, padded, o);
    if (lane == 0 && padded) atomicAdd(p.pad_count + frame_idx, padded);
  }
}

// BLUR = launches with 2..33 shutter samples: their staged box leaves room for 2 CTAs per SM anyway, so that
// instantiation trades occupancy for the registers of blur_interior_tile's texel cache.
template <int INTERP, bool BLUR>
__global__ void __launch_bounds__(NTHREADS, BLUR ? 2 : MIN_CTAS) warp_fused_kernel(const WarpParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_minv = reinterpret_cast<double*>(smem_raw);                       // [34][9], 16B multiple
  double* s_pk = s_minv + MINV_SLOTS * 9;                                     // [34][8] affine samples, packed for 16-byte loads
  float* s_scratch = reinterpret_cast<float*>(s_pk + MINV_SLOTS * 8);         // [8][384]
  int* s_box = reinterpret_cast<int*>(s_scratch + NWARPS * SCRATCH_FLOATS_PER_WARP);  // 8 ints
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_box + 6);  // mbarrier of the bulk copies
  float* s_cubic = reinterpret_cast<float*>(s_box + 8);                       // [32][4] bicubic coefficients
  float* s_tile = s_cubic + 128;                                              // staged source box

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int frame_idx = blockIdx.z;
  const int tx0 = blockIdx.x * TW;
  const int ty0 = blockIdx.y * TH;
  const int S = p.samples;
  const float* __restrict__ frame = p.src + (size_t)frame_idx * p.sh * p.sw * 3;

  if (INTERP == VSTAB_INTERP_BICUBIC && tid < 128) s_cubic[tid] = c_cubic_tab[tid >> 2][tid & 3];
  if (tid == NTHREADS - 1) mbar_init(s_bar, 1);  // made visible by the __syncthreads below

  // ---- inverse matrices + source footprint of the tile (4 corners x S samples) ---------------
  StagedTile tile;
  tile.smem = s_tile;
  tile.active = false;
  tile.x0 = tile.y0 = 0;
  tile.x1 = tile.y1 = -1;
  tile.pitch = 0;
  bool interior = false;
  bool blur_interior = false;  // BLUR: every sample's footprint inside the source and the staged box
  bool projective = false, any_projective = false;  // BLUR: some shutter sample has a projective last row
  bool bulk_pending = false;   // CTA-uniform: the staged box arrives through bulk copies
  const int txe = min(tx0 + TW, p.ow) - 1;
  const int tye = min(ty0 + TH, p.oh) - 1;
  if (S == 1) {
    // single sample: warp 0 inverts, projects the 4 corners (lane & 3) and reduces with shuffles;
    // no atomics and one barrier less than the general path
    if (warp == 0) {
      double mi[9];
      vstab_invert3(p.fwd + (size_t)frame_idx * 9, mi);
      const double cx = (lane & 1) ? (double)txe : (double)tx0;
      const double cy = (lane & 2) ? (double)tye : (double)ty0;
      const double X = mi[0] * cx + mi[1] * cy + mi[2];
      const double Y = mi[3] * cx + mi[4] * cy + mi[5];
      const double W = mi[6] * cx + mi[7] * cy + mi[8];
      const double sx = X / W, sy = Y / W;
      const bool bad = !(W > 1e-12) || !(fabs(sx) < 1e8) || !(fabs(sy) < 1e8);
      const int fx = bad ? 0 : (int)floor(sx), fy = bad ? 0 : (int)floor(sy);
      const int mnx = __reduce_min_sync(0xffffffffu, fx), mny = __reduce_min_sync(0xffffffffu, fy);
      const int mxx = __reduce_max_sync(0xffffffffu, fx), mxy = __reduce_max_sync(0xffffffffu, fy);
      const bool any_bad = __any_sync(0xffffffffu, bad);
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) s_minv[k] = mi[k];
        s_box[0] = mnx; s_box[1] = mny; s_box[2] = mxx; s_box[3] = mxy; s_box[4] = any_bad ? 1 : 0;
      }
    }
  } else {
    if (tid < S) {
      double* m = s_minv + tid * 9;
      vstab_invert3(p.fwd + ((size_t)frame_idx * S + tid) * 9, m);
      if (BLUR) {
        double* pk = s_pk + tid * 8;
        pk[0] = m[0]; pk[1] = m[1]; pk[2] = m[2]; pk[3] = m[3]; pk[4] = m[4]; pk[5] = m[5];
        pk[6] = (m[8] != 0.0) ? __ddiv_rn(32.0, m[8]) : 0.0;  // general_tile_body's sc_affine
        pk[7] = m[8];
        projective = !((m[6] == 0.0) && (m[7] == 0.0));
      }
    }
    if (tid == 0) {
      s_box[0] = INT_MAX;  // min x
      s_box[1] = INT_MAX;  // min y
      s_box[2] = INT_MIN;  // max x
      s_box[3] = INT_MIN;  // max y
      s_box[4] = 0;        // degenerate flag
    }
    any_projective = __syncthreads_or(projective) != 0;
    if (p.stage_mode == VSTAB_STAGE_AUTO) {
      for (int k = tid; k < 4 * S; k += NTHREADS) {
        const double* m = s_minv + (k >> 2) * 9;
        const double cx = (k & 1) ? (double)txe : (double)tx0;
        const double cy = (k & 2) ? (double)tye : (double)ty0;
        const double X = m[0] * cx + m[1] * cy + m[2];
        const double Y = m[3] * cx + m[4] * cy + m[5];
        const double W = m[6] * cx + m[7] * cy + m[8];
        const double sx = X / W, sy = Y / W;
        // Convexity of the projected tile needs W to keep one sign; positive is the sane case.
        if (!(W > 1e-12) || !(fabs(sx) < 1e8) || !(fabs(sy) < 1e8)) {
          atomicOr(&s_box[4], 1);
        } else {
          atomicMin(&s_box[0], (int)floor(sx));
          atomicMin(&s_box[1], (int)floor(sy));
          atomicMax(&s_box[2], (int)floor(sx));
          atomicMax(&s_box[3], (int)floor(sy));
        }
      }
    }
  }
  __syncthreads();
  if (p.stage_mode == VSTAB_STAGE_AUTO) {
    if (s_box[4] == 0) {
      constexpr int LO = (INTERP == VSTAB_INTERP_BILINEAR) ? 1 : 2;
      constexpr int HI = (INTERP == VSTAB_INTERP_BILINEAR) ? 2 : 3;
      int bx0 = max(s_box[0] - LO, 0), by0 = max(s_box[1] - LO, 0);
      int bx1 = min(s_box[2] + HI, p.sw - 1), by1 = min(s_box[3] + HI, p.sh - 1);
      if (p.vec_load) {  // 16-byte granules: 4 pixels = 48 bytes keeps every row 16B aligned
        bx0 &= ~3;
        bx1 = min(bx1 | 3, p.sw - 1);
      }
      const int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;
      if (bw > 0 && bh > 0 && (long long)bw * 3 * bh <= (long long)p.stage_capacity) {
        tile.active = true;
        // Interior tile: the whole (unclipped) footprint lies >= 1 px inside the source, so every
        // tap is in the staged box and every pixel is covered: no per-tap or per-pixel tests.
        const bool inside = (tx0 + TW <= p.ow) && (ty0 + TH <= p.oh) && (s_box[0] - LO >= 1) && (s_box[1] - LO >= 1) &&
                            (s_box[2] + HI <= p.sw - 2) && (s_box[3] + HI <= p.sh - 2);
        interior = inside && (S == 1) && (INTERP == VSTAB_INTERP_BILINEAR);
        blur_interior = inside && BLUR && (S > 1);
        tile.x0 = bx0;
        tile.y0 = by0;
        tile.x1 = bx1;
        tile.y1 = by1;
        tile.pitch = bw * 3;
        const int row_floats = bw * 3;
        if (p.vec_load) {
          // one bulk copy (TMA engine) per source row, issued by warp 0; bw % 4 == 0 keeps every row a
          // 16-byte multiple at a 16-byte aligned address.  All of them complete on one mbarrier.
          if (warp == 0) {
            const unsigned row_bytes = (unsigned)row_floats * 4u;
            if (lane == 0) mbar_expect_tx(s_bar, row_bytes * (unsigned)bh);
            __syncwarp();
            for (int r = lane; r < bh; r += 32)
              bulk_copy_g2s(s_tile + r * row_floats, frame + ((size_t)(by0 + r) * p.sw + bx0) * 3, row_bytes, s_bar);
          }
          bulk_pending = true;
        } else {
          for (int r = warp; r < bh; r += NWARPS) {
            const float* g = frame + ((size_t)(by0 + r) * p.sw + bx0) * 3;
            float* s = s_tile + r * row_floats;
            for (int v = lane; v < row_floats; v += 32) s[v] = __ldg(g + v);
          }
        }
      }
    }
  }

  // ---- interior tiles: branch-free bilinear, single sample ---------------------------------------
  // (the 16-byte cp.async copies of the source box are still in flight: the interior path computes
  //  its coordinates and weights first and only then waits for them)
  if (interior) {
    float* scratch = s_scratch + warp * SCRATCH_FLOATS_PER_WARP;
    const bool affine = (s_minv[6] == 0.0) && (s_minv[7] == 0.0);
    float* dst_tile = p.dst + (((size_t)frame_idx * p.oh + ty0) * p.ow + tx0) * 3;
    float* mask_tile = p.mask ? p.mask + ((size_t)frame_idx * p.oh + ty0) * p.ow + tx0 : nullptr;
    const float* tile0 = s_tile - (tile.y0 * tile.pitch + tile.x0 * 3);  // so that tile0[sy*pitch + sx*3] is the texel
    if (p.vec_store) {
      if (affine) interior_tile<true, true, true, GROUPS, 0>(s_minv, tile0, tile.pitch, tx0, ty0, warp, lane, tid, p.ow, scratch, dst_tile, mask_tile, p.vec_mask, s_bar, bulk_pending);
      else interior_tile<false, true, true, GROUPS, 0>(s_minv, tile0, tile.pitch, tx0, ty0, warp, lane, tid, p.ow, scratch, dst_tile, mask_tile, p.vec_mask, s_bar, bulk_pending);
    } else {
      if (affine) interior_tile<true, false, true, GROUPS, 0>(s_minv, tile0, tile.pitch, tx0, ty0, warp, lane, tid, p.ow, scratch, dst_tile, mask_tile, p.vec_mask, s_bar, bulk_pending);
      else interior_tile<false, false, true, GROUPS, 0>(s_minv, tile0, tile.pitch, tx0, ty0, warp, lane, tid, p.ow, scratch, dst_tile, mask_tile, p.vec_mask, s_bar, bulk_pending);
    }
    return;
  }
  if (bulk_pending) mbar_wait(s_bar, 0);
  __syncthreads();  // also orders the scalar staging path and the s_minv / s_cubic writes

  // VSTAB_STAGE_GLOBAL keeps every tile on general_tile_body: the reference the cached paths are tested against
  if (BLUR && p.stage_mode == VSTAB_STAGE_AUTO) {
    float* scratch = s_scratch + warp * SCRATCH_FLOATS_PER_WARP;
    if (blur_interior) {
      if (!any_projective) blur_tile<INTERP, true, false, GROUPS>(p, frame, frame_idx, tx0, ty0, s_minv, s_pk, tile, s_cubic, scratch, warp, lane, tid);
      else blur_tile<INTERP, false, false, GROUPS>(p, frame, frame_idx, tx0, ty0, s_minv, s_pk, tile, s_cubic, scratch, warp, lane, tid);
    } else {
      if (!any_projective) blur_tile<INTERP, true, true, GROUPS>(p, frame, frame_idx, tx0, ty0, s_minv, s_pk, tile, s_cubic, scratch, warp, lane, tid);
      else blur_tile<INTERP, false, true, GROUPS>(p, frame, frame_idx, tx0, ty0, s_minv, s_pk, tile, s_cubic, scratch, warp, lane, tid);
    }
    return;
  }
  general_tile<INTERP, GROUPS>(p, frame, frame_idx, tx0, ty0, s_minv, tile, s_cubic, s_scratch + warp * SCRATCH_FLOATS_PER_WARP, warp, lane);
}

// ---- streaming variant (single sample, 16-byte aligned source rows) ---------------------------------
// Every node path except motion blur.  A plan kernel computes, per 64x16 output tile, where its source box
// lies and whether the tile is interior; the resampling kernel is persistent (4 CTAs per SM) and walks its
// tiles through a two-stage shared-memory pipeline: ONE cp.async.bulk.tensor (3-D TMA descriptor over
// [frame][row][3*col], zero fill outside the frame) brings the 72x24 source box of tile i+2 while the
// eight warps resample tiles i and i+1, so no warp waits for DRAM or for a per-tile prologue.  The warp
// that finishes a tile last issues the next load for that stage.
#ifndef VSTAB_TMA_HINT
#define VSTAB_TMA_HINT 1
#endif
#ifndef VSTAB_STREAM_STORES
#define VSTAB_STREAM_STORES 1
#endif
constexpr int S_G = 1;                 // row groups per streamed tile
constexpr int S_TH = 16 * S_G;         // 64 x 16 output pixels
constexpr int S_BW = 76;               // source box: tile + tap margin + alignment + rotation / zoom slack
#ifndef VSTAB_S_BH_EXTRA
#define VSTAB_S_BH_EXTRA 8
#endif
constexpr int S_BH = S_TH + VSTAB_S_BH_EXTRA;
constexpr int S_PITCH = S_BW * 3;      // floats per staged row (228 <= 256, the TMA box limit)
constexpr int S_BOX_BYTES = S_PITCH * S_BH * 4;
constexpr int S_STAGE_BYTES = (S_BOX_BYTES + 127) / 128 * 128;  // stage buffers stay 128-byte aligned
#ifndef VSTAB_S_STAGES
#define VSTAB_S_STAGES 2
#endif
#ifndef VSTAB_S_CTAS
#define VSTAB_S_CTAS 4
#endif
constexpr int S_STAGES = VSTAB_S_STAGES;
constexpr int S_CTAS = VSTAB_S_CTAS;


struct TilePlan {  // 16 bytes, read as one int4
  int ox;        // source pixel of the box origin, multiple of 4 (may be negative: the TMA engine zero-fills)
  int oy_flags;  // (oy << 4) | flags; flags bit 0: box loaded, bit 1: interior tile, bit 3: edge tile
  int txy;       // output pixel of the tile origin: tx0 | ty0 << 16
  int frame;
};

struct StageMeta {
  double minv[10];  // 80 bytes, written by a bulk copy
  int tx0, ty0, frame, flags, ox, oy;
  int pad[6];
};
static_assert(sizeof(StageMeta) == 128, "StageMeta layout");
static_assert(3 * S_STAGES + 1 <= 8 && S_STAGES * 12 + 4 <= 32, "ring of 8 slots, 32 bytes of barriers and counters");

template <int INTERP>
__global__ void __launch_bounds__(256) warp_plan_kernel(const float* __restrict__ fwd, unsigned ntiles, int gx, int gy, int ow, int oh,
                                                        int sw, int sh, TilePlan* __restrict__ plan, double* __restrict__ minv) {
  constexpr int LO = (INTERP == VSTAB_INTERP_BILINEAR) ? 1 : 2;
  constexpr int HI = (INTERP == VSTAB_INTERP_BILINEAR) ? 2 : 3;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const unsigned per_frame = (unsigned)gx * (unsigned)gy;
  const int frame = (int)(t / per_frame);
  const unsigned rem = t - (unsigned)frame * per_frame;
  const int tyi = (int)(rem / (unsigned)gx);
  const int tx0 = (int)(rem - (unsigned)tyi * (unsigned)gx) * TW, ty0 = tyi * S_TH;
  const int txe = min(tx0 + TW, ow) - 1, tye = min(ty0 + S_TH, oh) - 1;
  double mi[9];
  vstab_invert3(fwd + (size_t)frame * 9, mi);
  if (rem == 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k) minv[(size_t)frame * 10 + k] = mi[k];
    minv[(size_t)frame * 10 + 9] = 0.0;
  }
  int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
  bool bad = false;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double cx = (c & 1) ? (double)txe : (double)tx0;
    const double cy = (c & 2) ? (double)tye : (double)ty0;
    const double X = mi[0] * cx + mi[1] * cy + mi[2];
    const double Y = mi[3] * cx + mi[4] * cy + mi[5];
    const double W = mi[6] * cx + mi[7] * cy + mi[8];
    const double sx = X / W, sy = Y / W;
    // convexity of the projected tile needs W to keep one sign; positive is the sane case
    if (!(W > 1e-12) || !(fabs(sx) < 1e8) || !(fabs(sy) < 1e8)) {
      bad = true;
    } else {
      const int fx = (int)floor(sx), fy = (int)floor(sy);
      mnx = min(mnx, fx); mxx = max(mxx, fx);
      mny = min(mny, fy); mxy = max(mxy, fy);
    }
  }
  int ox = 0, oy = 0, flags = 0;
  if (!bad) {
    // the TMA engine wants the innermost coordinate on a 16-byte boundary: ox is a multiple of 4 pixels
    const int need_w = mxx - mnx + 1 + LO + HI, need_h = mxy - mny + 1 + LO + HI;
    ox = (mnx - LO - (need_w + 3 <= S_BW ? (S_BW - need_w - 3) / 2 : 0)) & ~3;
    oy = mny - LO - (need_h <= S_BH ? (S_BH - need_h) / 2 : 0);
    flags = 1;
    if ((mxx + HI <= ox + S_BW - 1) && need_h <= S_BH && (tx0 + TW <= ow) && (ty0 + S_TH <= oh) &&
        (mnx - LO >= 1) && (mny - LO >= 1) && (mxx + HI <= sw - 2) && (mxy + HI <= sh - 2))
      flags |= 2;  // interior tile (either interpolation): every tap in the box and in the frame, every pixel covered
    else if ((mxx + HI <= ox + S_BW - 1) && need_h <= S_BH && (tx0 + TW <= ow) && (ty0 + S_TH <= oh) &&
             (mnx - LO >= -30000) && (mny - LO >= -30000) && (mxx + HI <= 30000) && (mxy + HI <= 30000))
      flags |= 8;  // edge tile: the staged box holds every tap position, some of them outside the frame
    else if (INTERP == VSTAB_INTERP_BILINEAR && (tx0 + TW <= ow) && (ty0 + S_TH <= oh) && (mnx - LO >= -30000) && (mny - LO >= -30000) &&
             (mxx + HI <= 30000) && (mxy + HI <= 30000)) {
      // mixed tile: footprint larger than the box (minifying map).  The box sits on the middle of the footprint; pixels
      // whose taps fall outside it gather from global memory (stream_edge<MIXED>)
      ox = ((mnx + mxx) / 2 - S_BW / 2) & ~3;
      oy = (mny + mxy) / 2 - S_BH / 2;
      flags |= 2 | 8;
    }
  }
  TilePlan pl;
  pl.ox = ox;
  pl.oy_flags = oy * 16 + flags;
  pl.txy = tx0 | (ty0 << 16);
  pl.frame = frame;
  plan[t] = pl;
}

__device__ __forceinline__ void tma_load_box(void* smem_dst, const void* tmap, int c0, int c1, int c2, unsigned long long* bar) {
#if VSTAB_TMA_HINT
  // source rows are re-read by the neighbouring tiles (halo): keep them in L2 ahead of the write stream
  unsigned long long policy;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(policy));
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
#else
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
#endif
}

// One thread: publish the stage's meta block and start the copies of tile t (box + inverse matrix).
__device__ __forceinline__ void stream_issue(const void* tmap, int4 pl, const double* __restrict__ minv, StageMeta* meta, float* box,
                                             unsigned long long* full) {
  const int flags = pl.y & 15, oy = pl.y >> 4, frame = pl.w;
  meta->tx0 = pl.z & 0xffff;
  meta->ty0 = (int)((unsigned)pl.z >> 16);
  meta->frame = frame;
  meta->flags = flags;
  meta->ox = pl.x;
  meta->oy = oy;
  // arrive.expect_tx has release semantics: the meta block is visible to whoever acquires the phase
  const bool do_load = (flags & 1) != 0;
  mbar_expect_tx(full, 80u + (do_load ? (unsigned)S_BOX_BYTES : 0u));
  bulk_copy_g2s(meta->minv, minv + (size_t)frame * 10, 80u, full);
  if (do_load) tma_load_box(box, tmap, pl.x * 3, oy, frame, full);
}

// Interior tile of the streaming kernel (64x16, box already in shared memory, compile-time pitch):
// coordinates are computed pixel by pixel so that few values stay live across the tile loop.
template <int INTERP, bool AFFINE, bool VEC>
__device__ __forceinline__ void stream_interior(const double* __restrict__ s_minv, const float* __restrict__ tile0, int tx0, int ty0,
                                                int warp, int lane, int tid, int ow, float* __restrict__ scratch,
                                                float* __restrict__ dst_tile, float* __restrict__ mask_tile, int vec_mask,
                                                const float* __restrict__ s_cubic) {
  const double m0 = s_minv[0], m1 = s_minv[1], m2 = s_minv[2], m3 = s_minv[3], m4 = s_minv[4], m5 = s_minv[5];
  const double m8 = s_minv[8];
  const double sc_affine = (m8 != 0.0) ? __ddiv_rn(32.0, m8) : 0.0;
  if (mask_tile) {  // fully covered tile: mask = 0; 64 x 16 floats = one 16-byte store per thread
    float* mrow = mask_tile + (size_t)(tid >> 4) * ow + (tid & 15) * 4;
    if (vec_mask) {
      if (VSTAB_STREAM_STORES) __stcs(reinterpret_cast<float4*>(mrow), make_float4(0.f, 0.f, 0.f, 0.f));
      else *reinterpret_cast<float4*>(mrow) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      mrow[0] = 0.f; mrow[1] = 0.f; mrow[2] = 0.f; mrow[3] = 0.f;
    }
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const double dy = (double)(ty0 + warp + NWARPS * rr);
    const double bx = __dmul_rn(m1, dy), by = __dmul_rn(m4, dy);
    double bw = 0.0;
    if (!AFFINE) bw = __dmul_rn(s_minv[7], dy);
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const double dx = (double)(tx0 + lane + 32 * cc);
      const double X = __dadd_rn(__dadd_rn(__dmul_rn(m0, dx), bx), m2);
      const double Y = __dadd_rn(__dadd_rn(__dmul_rn(m3, dx), by), m5);
      double sc = sc_affine;
      if (!AFFINE) {
        const double W = __dadd_rn(__dadd_rn(__dmul_rn(s_minv[6], dx), bw), m8);
        sc = (W != 0.0) ? __ddiv_rn(32.0, W) : 0.0;
      }
      // |coordinates| < 2^15 here, so cv2's INT_MIN/INT_MAX and short saturation are no-ops
      const int ix = __double2int_rn(__dmul_rn(X, sc)), iy = __double2int_rn(__dmul_rn(Y, sc));
      float v[3];
      if (INTERP == VSTAB_INTERP_BILINEAR) {
        const float fx1 = (float)(ix & 31) * 0.03125f, fy1 = (float)(iy & 31) * 0.03125f;
        const float fx0 = 1.0f - fx1, fy0 = 1.0f - fy1;
        const float w00 = __fmul_rn(fy0, fx0), w01 = __fmul_rn(fy0, fx1);
        const float w10 = __fmul_rn(fy1, fx0), w11 = __fmul_rn(fy1, fx1);
        const float* s0 = tile0 + (iy >> 5) * S_PITCH + (ix >> 5) * 3;
        const float* s1 = s0 + S_PITCH;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
          v[ch] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s0[ch], w00), __fmul_rn(s0[3 + ch], w01)), __fmul_rn(s1[ch], w10)),
                            __fmul_rn(s1[3 + ch], w11));
      } else {
        // remapBicubic away from the border: row sums of S * (wy * wx), added top to bottom (cubic_rows_sum)
        const float4 wx4 = *reinterpret_cast<const float4*>(s_cubic + (ix & 31) * 4);
        const float4 wy4 = *reinterpret_cast<const float4*>(s_cubic + (iy & 31) * 4);
        const float wx[4] = {wx4.x, wx4.y, wx4.z, wx4.w}, wy[4] = {wy4.x, wy4.y, wy4.z, wy4.w};
        cubic_rows_sum(tile0 + ((iy >> 5) - 1) * S_PITCH + ((ix >> 5) - 1) * 3, S_PITCH, wx, wy, v);
      }
      float* o = VEC ? scratch + rr * (TW * 3) + (lane + cc * 32) * 3 : dst_tile + ((size_t)(warp + rr * NWARPS) * ow + lane + cc * 32) * 3;
      o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    }
  }
  if (VEC) {
    __syncwarp();
#pragma unroll
    for (int q3 = 0; q3 < 3; ++q3) {
      const int q = lane + q3 * 32;
      const int rr = q >= 48 ? 1 : 0, qi = q - rr * 48;
      const float4 val = *reinterpret_cast<const float4*>(scratch + rr * (TW * 3) + qi * 4);
      float4* o4 = reinterpret_cast<float4*>(dst_tile + (size_t)(warp + rr * NWARPS) * ow * 3 + qi * 4);
      if (VSTAB_STREAM_STORES) __stcs(o4, val); else *o4 = val;  // written once, never re-read by this launch
    }
    __syncwarp();
  }
}

// Edge tile of the streaming kernel: a full 64x16 tile whose footprint leaves the source frame but still
// fits the staged box.  Same arithmetic as stream_interior plus what the frame border needs: per-tap
// BORDER_CONSTANT substitution (the TMA engine zero-filled those texels), remapBilinear's "footprint
// entirely outside" rule, the coverage test of the padding mask and the padded-pixel count.
// MIXED (bilinear only): the tile's source footprint is LARGER than the staged box (the map minifies: zoom-out, or the
// far side of a projective correction).  Same arithmetic, but every pixel first asks whether its 2x2 footprint lies in
// the box -- then it is served from shared memory like an edge pixel -- and gathers its taps from global memory (L1)
// otherwise.  Before round 2 such tiles took the general path and a frame with 10 % minification ran at 2.5 TB/s, with
// 20 % at 1.7 TB/s (scripts/warp_persp_probe.py): the late frames of a camera-locked perspective clip.
template <int INTERP, bool AFFINE, bool VEC, bool MIXED = false>
__device__ __forceinline__ void stream_edge(const WarpParams& p, const double* __restrict__ s_minv, const float* __restrict__ tile0,
                                            int tx0, int ty0, int frame_idx, int warp, int lane, float* __restrict__ scratch,
                                            float* __restrict__ dst_tile, float* __restrict__ mask_tile, const float* __restrict__ s_cubic,
                                            int box_x0 = 0, int box_y0 = 0) {
  const double m0 = s_minv[0], m1 = s_minv[1], m2 = s_minv[2], m3 = s_minv[3], m4 = s_minv[4], m5 = s_minv[5];
  const double m8 = s_minv[8];
  const double sc_affine = (m8 != 0.0) ? __ddiv_rn(32.0, m8) : 0.0;
  const double x_hi = (double)(p.sw - 1), y_hi = (double)(p.sh - 1);
  const float br = p.border[0], bg = p.border[1], bb = p.border[2];
  const int ow = p.ow;
  const int rule = sample_rule(p, frame_idx, 0);
  unsigned padded = 0;
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const double dy = (double)(ty0 + warp + NWARPS * rr);
    const double bx = __dmul_rn(m1, dy), by = __dmul_rn(m4, dy);
    double bw = 0.0;
    if (!AFFINE) bw = __dmul_rn(s_minv[7], dy);
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const double dx = (double)(tx0 + lane + 32 * cc);
      const double X = __dadd_rn(__dadd_rn(__dmul_rn(m0, dx), bx), m2);
      const double Y = __dadd_rn(__dadd_rn(__dmul_rn(m3, dx), by), m5);
      double W = m8, sc = sc_affine;
      if (!AFFINE) {
        W = __dadd_rn(__dadd_rn(__dmul_rn(s_minv[6], dx), bw), m8);
        sc = (W != 0.0) ? __ddiv_rn(32.0, W) : 0.0;
      }
      // coverage (INTER_NEAREST ones warp, closed source rectangle): X * fl(1/W) decides everything that is
      // not within 1e-6 px of a boundary; only those pixels pay for the exact divisions
      bool ok;
      {
        const double rw = sc * 0.03125;
        const double qx = X * rw, qy = Y * rw;
        const double band = 1e-6;
        if (rule == VSTAB_MASK_RULE_P && qx > band && qx < x_hi - band && qy > band && qy < y_hi - band) {
          ok = true;
        } else if (rule == VSTAB_MASK_RULE_P && (qx < -band || qx > x_hi + band || qy < -band || qy > y_hi + band)) {
          ok = false;
        } else {
          double cxs = __ddiv_rn(X, W), cys = __ddiv_rn(Y, W);
          if (rule == VSTAB_MASK_RULE_C) {
            cxs = rint(cxs);
            cys = rint(cys);
          }
          ok = (cxs >= 0.0) && (cxs <= x_hi) && (cys >= 0.0) && (cys <= y_hi);
        }
      }
      const float mval = ok ? 0.0f : 1.0f;
      padded += ok ? 0u : 1u;
      if (mask_tile) __stcs(mask_tile + (size_t)(warp + rr * NWARPS) * ow + lane + cc * 32, mval);
      // the plan keeps |coordinates| < 2^15 for edge tiles: cv2's saturations are no-ops
      const int ix = __double2int_rn(__dmul_rn(X, sc)), iy = __double2int_rn(__dmul_rn(Y, sc));
      const int sx = ix >> 5, sy = iy >> 5;
      float v[3];
      if (INTERP == VSTAB_INTERP_BILINEAR) {
        const float fx1 = (float)(ix & 31) * 0.03125f, fy1 = (float)(iy & 31) * 0.03125f;
        const float fx0 = 1.0f - fx1, fy0 = 1.0f - fy1;
        const float w00 = __fmul_rn(fy0, fx0), w01 = __fmul_rn(fy0, fx1);
        const float w10 = __fmul_rn(fy1, fx0), w11 = __fmul_rn(fy1, fx1);
        const bool x0in = (unsigned)sx < (unsigned)p.sw, x1in = (unsigned)(sx + 1) < (unsigned)p.sw;
        const bool y0in = (unsigned)sy < (unsigned)p.sh, y1in = (unsigned)(sy + 1) < (unsigned)p.sh;
        const bool in00 = x0in && y0in, in01 = x1in && y0in, in10 = x0in && y1in, in11 = x1in && y1in;
        const bool any_in = (x0in || x1in) && (y0in || y1in);  // else: remapBilinear returns the border colour itself
        const float* s0 = tile0 + sy * S_PITCH + sx * 3;
        const float* s1 = s0 + S_PITCH;
        int step1 = S_PITCH;
        if (MIXED && !(sx >= box_x0 && sx + 1 <= box_x0 + S_BW - 1 && sy >= box_y0 && sy + 1 <= box_y0 + S_BH - 1)) {
          // footprint (partly) outside the staged box: the same four taps from the frame itself; taps outside the
          // frame are never dereferenced (in00..in11 select the border colour)
          s0 = p.src + ((size_t)frame_idx * p.sh + sy) * (size_t)p.sw * 3 + (ptrdiff_t)sx * 3;
          step1 = p.sw * 3;
          s1 = s0 + step1;
        }
        (void)step1;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float bc = ch == 0 ? br : (ch == 1 ? bg : bb);
          const float t00 = in00 ? s0[ch] : bc, t01 = in01 ? s0[3 + ch] : bc;
          const float t10 = in10 ? s1[ch] : bc, t11 = in11 ? s1[3 + ch] : bc;
          const float r = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00, w00), __fmul_rn(t01, w01)), __fmul_rn(t10, w10)), __fmul_rn(t11, w11));
          v[ch] = any_in ? r : bc;
        }
      } else {
        // remapBicubic: all 16 taps inside the frame -> row sums (cubic_rows_sum); otherwise the border branch,
        // cv + sum over the in-range taps of (S - cv) * w (nothing in range leaves cv itself)
        const float4 wx4 = *reinterpret_cast<const float4*>(s_cubic + (ix & 31) * 4);
        const float4 wy4 = *reinterpret_cast<const float4*>(s_cubic + (iy & 31) * 4);
        const float wx[4] = {wx4.x, wx4.y, wx4.z, wx4.w}, wy[4] = {wy4.x, wy4.y, wy4.z, wy4.w};
        const int bxs = sx - 1, bys = sy - 1;
        const float* s0 = tile0 + bys * S_PITCH + bxs * 3;
        if (bxs >= 0 && bxs < p.sw - 3 && bys >= 0 && bys < p.sh - 3) {
          cubic_rows_sum(s0, S_PITCH, wx, wy, v);
        } else {
          v[0] = br; v[1] = bg; v[2] = bb;
#pragma unroll
          for (int k1 = 0; k1 < 4; ++k1) {
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
              if ((unsigned)(bxs + k2) < (unsigned)p.sw && (unsigned)(bys + k1) < (unsigned)p.sh) {
                const float w = __fmul_rn(wy[k1], wx[k2]);
                const float* t = s0 + k1 * S_PITCH + k2 * 3;
                v[0] = __fadd_rn(v[0], __fmul_rn(__fsub_rn(t[0], br), w));
                v[1] = __fadd_rn(v[1], __fmul_rn(__fsub_rn(t[1], bg), w));
                v[2] = __fadd_rn(v[2], __fmul_rn(__fsub_rn(t[2], bb), w));
              }
            }
          }
        }
      }
      float* o = VEC ? scratch + rr * (TW * 3) + (lane + cc * 32) * 3 : dst_tile + ((size_t)(warp + rr * NWARPS) * ow + lane + cc * 32) * 3;
      o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    }
  }
  if (VEC) {
    __syncwarp();
#pragma unroll
    for (int q3 = 0; q3 < 3; ++q3) {
      const int q = lane + q3 * 32;
      const int rr = q >= 48 ? 1 : 0, qi = q - rr * 48;
      const float4 val = *reinterpret_cast<const float4*>(scratch + rr * (TW * 3) + qi * 4);
      __stcs(reinterpret_cast<float4*>(dst_tile + (size_t)(warp + rr * NWARPS) * ow * 3 + qi * 4), val);
    }
    __syncwarp();
  }
  if (p.pad_count) {
    padded = __reduce_add_sync(0xffffffffu, padded);
    if (lane == 0 && padded) atomicAdd(p.pad_count + frame_idx, padded);
  }
}

template <int INTERP>
__global__ void __launch_bounds__(NTHREADS, S_CTAS) warp_stream_kernel(const __grid_constant__ CUtensorMap tmap, const WarpParams p,
                                                                       const TilePlan* __restrict__ plan,
                                                                       const double* __restrict__ minv, unsigned* __restrict__ ticket, unsigned ntiles) {
  extern __shared__ __align__(128) unsigned char smem_stream[];
  float* s_stage = reinterpret_cast<float*>(smem_stream);                                        // [STAGES][S_BOX_BYTES]
  StageMeta* s_meta = reinterpret_cast<StageMeta*>(smem_stream + S_STAGES * S_STAGE_BYTES);      // [STAGES]
  unsigned long long* s_full = reinterpret_cast<unsigned long long*>(s_meta + S_STAGES);         // [STAGES]
  int* s_done = reinterpret_cast<int*>(s_full + S_STAGES);                                       // [STAGES] warps finished
  unsigned* s_ring_t = reinterpret_cast<unsigned*>(smem_stream + S_STAGES * S_STAGE_BYTES + S_STAGES * 128 + 32);  // [8] ticket per sequence slot
  int4* s_ring_p = reinterpret_cast<int4*>(smem_stream + S_STAGES * S_STAGE_BYTES + S_STAGES * 128 + 64);          // [8] plan entry per sequence slot
  float* s_cubic = reinterpret_cast<float*>(smem_stream + S_STAGES * S_STAGE_BYTES + S_STAGES * 128 + 192);  // [32][4]
  float* s_scratch = s_cubic + 128;                                                              // [8][384]

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  if (INTERP == VSTAB_INTERP_BICUBIC && tid < 128) s_cubic[tid] = c_cubic_tab[tid >> 2][tid & 3];
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < S_STAGES; ++k) {
      mbar_init(s_full + k, 1);
      s_done[k] = 0;
    }
    s_done[S_STAGES] = 0;  // empty slots in a row
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  // Tiles are handed out through a global ticket so that the tiles in flight at any moment are neighbours
  // in the frame (their halos hit in L2 whatever the relative speed of the CTAs).  Sequence slot q of this
  // CTA owns ring entry q % 8: its ticket T[q] and its plan entry P[q].  The warp that finishes tile j last
  // issues the load of slot j+S (S = stages) straight from the ring.  Refilling the ring costs two global
  // round trips (ticket, plan entry); warp j % 8 pays them at the end of tile j, for P[j+2S] and T[j+3S],
  // before it reports the tile as finished -- so the entries are published by the stage's next
  // arrive.expect_tx, S tiles before anybody reads them, and the issuing warp never waits on global memory.
  if (tid == 0) {
    const unsigned base = atomicAdd(ticket, 3u * S_STAGES);
#pragma unroll
    for (int k = 0; k < 3 * S_STAGES; ++k) s_ring_t[k] = base + k;
#pragma unroll
    for (int k = 0; k < 2 * S_STAGES; ++k)
      s_ring_p[k] = (base + k < ntiles) ? __ldg(reinterpret_cast<const int4*>(plan) + base + k) : make_int4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < S_STAGES; ++k) {
      if (base + k < ntiles) {
        stream_issue(&tmap, s_ring_p[k], minv, s_meta + k, s_stage + k * (S_STAGE_BYTES / 4), s_full + k);
      } else {
        const int run = s_done[S_STAGES] + 1;
        s_done[S_STAGES] = run;
        s_meta[k].flags = 4 | (run << 8);  // no tile for this slot
        mbar_arrive(s_full + k);
      }
    }
  }
  __syncthreads();

  float* scratch = s_scratch + warp * SCRATCH_FLOATS_PER_WARP;
  for (unsigned it = 0;; ++it) {
    const int stage = it % S_STAGES;
    mbar_wait_spin(s_full + stage, (it / S_STAGES) & 1);  // no suspend hint: a hinted wait sleeps its full 20 us when the phase is completed by a plain arrive
    const StageMeta* meta = s_meta + stage;
    const int flags = meta->flags;
    // The ring is refilled by different warps, so a CTA's tickets are not monotonic in slot order: an
    // exhausted slot (flag 4) can be followed by a live one.  The issuing thread counts the empty slots
    // in a row (bits 8.. of the flags); after 3S+1 of them every ticket this CTA ever drew has been served.
    if ((flags & 4) && (flags >> 8) > 3 * S_STAGES) break;
    const float* box = s_stage + stage * (S_STAGE_BYTES / 4);
    const int tx0 = meta->tx0, ty0 = meta->ty0, frame_idx = meta->frame;
    const int ox = meta->ox, oy = meta->oy;
    if (flags & 4) {
      // empty slot: nothing to resample, the stage still goes through the hand-over below
    } else if ((flags & 10) == 10) {
      if (INTERP == VSTAB_INTERP_BILINEAR) {  // the plan only marks bilinear tiles as mixed
        const bool affine = (meta->minv[6] == 0.0) && (meta->minv[7] == 0.0);
        float* dst_tile = p.dst + (((size_t)frame_idx * p.oh + ty0) * p.ow + tx0) * 3;
        float* mask_tile = p.mask ? p.mask + ((size_t)frame_idx * p.oh + ty0) * p.ow + tx0 : nullptr;
        const float* tile0 = box - (oy * S_PITCH + ox * 3);
        if (p.vec_store) {
          if (affine) stream_edge<VSTAB_INTERP_BILINEAR, true, true, true>(p, meta->minv, tile0, tx0, ty0, frame_idx, warp, lane, scratch, dst_tile, mask_tile, s_cubic, ox, oy);
          else stream_edge<VSTAB_INTERP_BILINEAR, false, true, true>(p, meta->minv, tile0, tx0, ty0, frame_idx, warp, lane, scratch, dst_tile, mask_tile, s_cubic, ox, oy);
        } else {
          if (affine) stream_edge<VSTAB_INTERP_BILINEAR, true, false, true>(p, meta->minv, tile0, tx0, ty0, frame_idx, warp, lane, scratch, dst_tile, mask_tile, s_cubic, ox, oy);
          else stream_edge<VSTAB_INTERP_BILINEAR, false, false, true>(p, meta->minv, tile0, tx0, ty0, frame_idx, warp, lane, scratch, dst_tile, mask_tile, s_cubic, ox, oy);
        }
      }
    } else if (flags & 2) {
      const bool affine = (meta->minv[6] == 0.0) && (meta->minv[7] == 0.0);
      float* dst_tile = p.dst + (((size_t)frame_idx * p.oh + ty0) * p.ow + tx0) * 3;
      float* mask_tile = p.mask ? p.mask + ((size_t)frame_idx * p.oh + ty0) * p.ow + tx0 : nullptr;
      const float* tile0 = box - (oy * S_PITCH + ox * 3);
      if (p.vec_store) {
        if (affine) stream_interior<INTERP, true, true>(meta->minv, tile0, tx0, ty0, warp, lane, tid, p.ow, scratch, dst_tile, mask_tile, p.vec_mask, s_cubic);
        else stream_interior<INTERP, false, true>(meta->minv, tile0, tx0, ty0, warp, lane, tid, p.ow, scratch, dst_tile, mask_tile, p.vec_mask, s_cubic);
      } else {
        if (affine) stream_interior<INTERP, true, false>(meta->minv, tile0, tx0, ty0, warp, lane, tid, p.ow, scratch, dst_tile, mask_tile, p.vec_mask, s_cubic);
        else stream_interior<INTERP, false, false>(meta->minv, tile0, tx0, ty0, warp, lane, tid, p.ow, scratch, dst_tile, mask_tile, p.vec_mask, s_cubic);
      }
    } else if (flags & 8) {
      const bool affine = (meta->minv[6] == 0.0) && (meta->minv[7] == 0.0);
      float* dst_tile = p.dst + (((size_t)frame_idx * p.oh + ty0) * p.ow + tx0) * 3;
      float* mask_tile = p.mask ? p.mask + ((size_t)frame_idx * p.oh + ty0) * p.ow + tx0 : nullptr;
      const float* tile0 = box - (oy * S_PITCH + ox * 3);
      if (p.vec_store) {
        if (affine) stream_edge<INTERP, true, true>(p, meta->minv, tile0, tx0, ty0, frame_idx, warp, lane, scratch, dst_tile, mask_tile, s_cubic);
        else stream_edge<INTERP, false, true>(p, meta->minv, tile0, tx0, ty0, frame_idx, warp, lane, scratch, dst_tile, mask_tile, s_cubic);
      } else {
        if (affine) stream_edge<INTERP, true, false>(p, meta->minv, tile0, tx0, ty0, frame_idx, warp, lane, scratch, dst_tile, mask_tile, s_cubic);
        else stream_edge<INTERP, false, false>(p, meta->minv, tile0, tx0, ty0, frame_idx, warp, lane, scratch, dst_tile, mask_tile, s_cubic);
      }
    } else {
      StagedTile tile;
      tile.x0 = max(ox, 0);
      tile.y0 = max(oy, 0);
      tile.x1 = min(ox + S_BW - 1, p.sw - 1);
      tile.y1 = min(oy + S_BH - 1, p.sh - 1);
      tile.pitch = S_PITCH;
      tile.active = (flags & 1) && tile.x1 >= tile.x0 && tile.y1 >= tile.y0;
      tile.smem = box + ((tile.y0 - oy) * S_PITCH + (tile.x0 - ox) * 3);
      const float* frame = p.src + (size_t)frame_idx * p.sh * p.sw * 3;
      general_tile<INTERP, S_G>(p, frame, frame_idx, tx0, ty0, meta->minv, tile, s_cubic, scratch, warp, lane);
    }
    __syncwarp();
    if (lane == 0) {
      if (warp == (int)(it & 7)) {  // this tile's ring refill (blocks this warp for one global round trip)
        const unsigned tq = s_ring_t[(it + 2 * S_STAGES) & 7];
        int4 pl = make_int4(0, 0, 0, 0);
        unsigned tn = 0xffffffffu;
        if (tq < ntiles) {
          pl = __ldg(reinterpret_cast<const int4*>(plan) + tq);
          tn = atomicAdd(ticket, 1u);
        }
        s_ring_p[(it + 2 * S_STAGES) & 7] = pl;
        s_ring_t[(it + 3 * S_STAGES) & 7] = tn;
      }
      // The last warp to finish with this stage starts the load of the tile that reuses it.
      // No memory fence here: a fence would also wait for this warp's global stores to drain (about a
      // microsecond on the critical path of the next load).  Everything the counter orders lives in shared
      // memory -- the stage (already read into registers) and the ring -- and the SM performs one warp's
      // shared-memory operations in program order.
      asm volatile("" ::: "memory");
      const int done = atomicAdd(s_done + stage, 1);
      if (done == NWARPS - 1) {
        s_done[stage] = 0;
        if (s_ring_t[(it + S_STAGES) & 7] < ntiles) {
          s_done[S_STAGES] = 0;
          stream_issue(&tmap, s_ring_p[(it + S_STAGES) & 7], minv, s_meta + stage, s_stage + stage * (S_STAGE_BYTES / 4), s_full + stage);
        } else {
          const int run = s_done[S_STAGES] + 1;  // only the (serialised) issuing threads touch this counter
          s_done[S_STAGES] = run;
          s_meta[stage].flags = 4 | (run << 8);
          mbar_arrive(s_full + stage);
        }
      }
    }
  }
}

// AND-reduction of INTER_NEAREST coverage over n matrices (crop solvers).
__global__ void __launch_bounds__(256) common_coverage_kernel(const float* __restrict__ fwd, int n,
                                                              int sh, int sw, int oh, int ow,
                                                              int mask_rule, const unsigned char* __restrict__ rules,
                                                              unsigned char* __restrict__ common) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_minv = reinterpret_cast<double*>(smem_raw);  // [chunk][9]
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
  const int oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  bool all_ok = true;
  constexpr int CHUNK = 64;
  for (int base = 0; base < n; base += CHUNK) {
    const int cnt = min(CHUNK, n - base);
    __syncthreads();
    if ((int)threadIdx.x < cnt)
      vstab_invert3(fwd + (size_t)(base + threadIdx.x) * 9, s_minv + threadIdx.x * 9);
    __syncthreads();
    if (ox < ow && oy < oh) {
      const double dx = (double)ox, dy = (double)oy;
      for (int k = 0; k < cnt; ++k) {
        const double* m = s_minv + k * 9;
        const double X = __dadd_rn(__dadd_rn(__dmul_rn(m[0], dx), __dmul_rn(m[1], dy)), m[2]);
        const double Y = __dadd_rn(__dadd_rn(__dmul_rn(m[3], dx), __dmul_rn(m[4], dy)), m[5]);
        const double W = __dadd_rn(__dadd_rn(__dmul_rn(m[6], dx), __dmul_rn(m[7], dy)), m[8]);
        double cxs = __ddiv_rn(X, W), cys = __ddiv_rn(Y, W);
        if ((rules ? (int)rules[base + k] : mask_rule) == VSTAB_MASK_RULE_C) {
          cxs = rint(cxs);
          cys = rint(cys);
        }
        all_ok = all_ok && (cxs >= 0.0) && (cxs <= (double)(sw - 1)) && (cys >= 0.0) &&
                 (cys <= (double)(sh - 1));
      }
    }
  }
  if (ox < ow && oy < oh) common[(size_t)oy * ow + ox] = all_ok ? 1 : 0;
}

// Per-frame INTER_NEAREST coverage as bytes (crop solvers: the closing + bounding box below).
__global__ void __launch_bounds__(256) coverage_u8_kernel(const float* __restrict__ fwd, int sh, int sw, int oh, int ow,
                                                          int mask_rule, const unsigned char* __restrict__ rules,
                                                          unsigned char* __restrict__ cov) {
  __shared__ double s_m[9];
  const int f = blockIdx.z;
  if (threadIdx.x == 0) vstab_invert3(fwd + (size_t)f * 9, s_m);
  __syncthreads();
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
  const int oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= ow || oy >= oh) return;
  const double dx = (double)ox, dy = (double)oy;
  const double X = __dadd_rn(__dadd_rn(__dmul_rn(s_m[0], dx), __dmul_rn(s_m[1], dy)), s_m[2]);
  const double Y = __dadd_rn(__dadd_rn(__dmul_rn(s_m[3], dx), __dmul_rn(s_m[4], dy)), s_m[5]);
  const double W = __dadd_rn(__dadd_rn(__dmul_rn(s_m[6], dx), __dmul_rn(s_m[7], dy)), s_m[8]);
  double cxs = __ddiv_rn(X, W), cys = __ddiv_rn(Y, W);
  if ((rules ? (int)rules[f] : mask_rule) == VSTAB_MASK_RULE_C) {
    cxs = rint(cxs);
    cys = rint(cys);
  }
  const bool ok = (cxs >= 0.0) && (cxs <= (double)(sw - 1)) && (cys >= 0.0) && (cys <= (double)(sh - 1));
  cov[((size_t)f * oh + oy) * ow + ox] = ok ? 1 : 0;
}

// Bounding box of erode3x3(dilate3x3(coverage)) per frame (cv2 morphology ignores out-of-image
// neighbours): bbox[f] = {xmin, ymin, xmax, ymax}, initialised to {INT_MAX, INT_MAX, -1, -1}.
__global__ void __launch_bounds__(256) closed_bbox_kernel(const unsigned char* __restrict__ cov, int oh, int ow,
                                                          int* __restrict__ bbox) {
  const int f = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const unsigned char* c = cov + (size_t)f * oh * ow;
  bool closed = false;
  if (x < ow && y < oh) {
    closed = true;
    for (int qy = max(y - 1, 0); qy <= min(y + 1, oh - 1) && closed; ++qy)
      for (int qx = max(x - 1, 0); qx <= min(x + 1, ow - 1) && closed; ++qx) {
        bool any = false;
        for (int ry = max(qy - 1, 0); ry <= min(qy + 1, oh - 1) && !any; ++ry)
          for (int rx = max(qx - 1, 0); rx <= min(qx + 1, ow - 1) && !any; ++rx) any = c[(size_t)ry * ow + rx] != 0;
        closed = any;
      }
  }
  int xmin = closed ? x : INT_MAX, ymin = closed ? y : INT_MAX, xmax = closed ? x : -1, ymax = closed ? y : -1;
  for (int o = 16; o > 0; o >>= 1) {
    xmin = min(xmin, __shfl_down_sync(0xffffffffu, xmin, o));
    ymin = min(ymin, __shfl_down_sync(0xffffffffu, ymin, o));
    xmax = max(xmax, __shfl_down_sync(0xffffffffu, xmax, o));
    ymax = max(ymax, __shfl_down_sync(0xffffffffu, ymax, o));
  }
  if ((threadIdx.x & 31) == 0 && xmax >= 0) {
    atomicMin(bbox + f * 4 + 0, xmin);
    atomicMin(bbox + f * 4 + 1, ymin);
    atomicMax(bbox + f * 4 + 2, xmax);
    atomicMax(bbox + f * 4 + 3, ymax);
  }
}

__global__ void bbox_init_kernel(int* bbox, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    bbox[i * 4 + 0] = INT_MAX;
    bbox[i * 4 + 1] = INT_MAX;
    bbox[i * 4 + 2] = -1;
    bbox[i * 4 + 3] = -1;
  }
}

bool g_cubic_tab_ready[64] = {false};

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (libcuda is not linked)
typedef CUresult (*tensor_map_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
tensor_map_encode_fn tensor_map_encoder() {
  static tensor_map_encode_fn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (tensor_map_encode_fn)f;
  }();
  return fn;
}

void build_cubic_tab(float tab[32][4]) {
  // cv::interpolateCubic with A = -0.75, float32 arithmetic in this exact order
  const volatile float A = -0.75f;
  for (int i = 0; i < 32; ++i) {
    volatile float x = (float)i * (1.0f / 32.0f);
    volatile float x1 = x + 1.0f;
    volatile float t0 = A * x1;
    volatile float t1 = t0 - 5.0f * A;
    volatile float t2 = t1 * x1;
    volatile float t3 = t2 + 8.0f * A;
    volatile float t4 = t3 * x1;
    volatile float c0 = t4 - 4.0f * A;
    volatile float u0 = (A + 2.0f) * x;
    volatile float u1 = u0 - (A + 3.0f);
    volatile float u2 = u1 * x;
    volatile float u3 = u2 * x;
    volatile float c1 = u3 + 1.0f;
    volatile float xm = 1.0f - x;
    volatile float v0 = (A + 2.0f) * xm;
    volatile float v1 = v0 - (A + 3.0f);
    volatile float v2 = v1 * xm;
    volatile float v3 = v2 * xm;
    volatile float c2 = v3 + 1.0f;
    volatile float w0 = 1.0f - c0;
    volatile float w1 = w0 - c1;
    volatile float c3 = w1 - c2;
    tab[i][0] = c0;
    tab[i][1] = c1;
    tab[i][2] = c2;
    tab[i][3] = c3;
  }
}

}  // namespace

namespace {
// Splits the ABI's mask_rule argument and, for VSTAB_MASK_RULE_AUTO, fills the handle's per-matrix rule table.
// *rules_out stays nullptr for the two fixed rules.
int resolve_mask_rule(vstab_handle* h, const float* fwd_dev, int count, int src_h, int src_w, int out_h, int out_w,
                      int mask_rule, cudaStream_t st, int* rule_out, const unsigned char** rules_out) {
  const int rule = mask_rule & 0xff, threads = mask_rule >> 8;
  *rule_out = rule;
  *rules_out = nullptr;
  if (rule == VSTAB_MASK_RULE_P || rule == VSTAB_MASK_RULE_C) {
    if (threads != 0) return vstab_fail(h, VSTAB_ERR_INVALID, "mask_rule: a thread count only goes with VSTAB_MASK_RULE_AUTO");
    return VSTAB_OK;
  }
  if (rule != VSTAB_MASK_RULE_AUTO || threads < 1) return vstab_fail(h, VSTAB_ERR_INVALID, "unknown mask rule (AUTO needs VSTAB_MASK_RULE_AUTO_THREADS(t >= 1))");
  if (threads == 1 || count == 0) {  // one stripe only misses the source when the whole output does: Rule P
    *rule_out = VSTAB_MASK_RULE_P;
    return VSTAB_OK;
  }
  if ((size_t)count > h->rules_bytes) {
    VSTAB_CUDA(h, cudaDeviceSynchronize());
    if (h->rules) cudaFree(h->rules);
    h->rules = nullptr;
    h->rules_bytes = 0;
    const size_t want = (size_t)count + 4096;
    if (cudaMalloc(&h->rules, want) != cudaSuccess) return vstab_fail(h, VSTAB_ERR_NOMEM, "mask rule table: cudaMalloc failed");
    h->rules_bytes = want;
  }
  mask_rule_auto_kernel<<<vstab_ceil_div(count, 128), 128, 0, st>>>(fwd_dev, count, src_h, src_w, out_h, out_w, threads,
                                                                     (unsigned char*)h->rules);
  VSTAB_LAUNCH_CHECK(h, "mask_rule_auto_kernel");
  *rule_out = VSTAB_MASK_RULE_P;
  *rules_out = (const unsigned char*)h->rules;
  return VSTAB_OK;
}
}  // namespace

extern "C" int vstab_warp_fused(vstab_handle* h, const float* src_dev, int n, int src_h, int src_w,
                                const float* fwd_dev, int samples, int interp, int out_h,
                                int out_w, const float* border_host, int mask_rule,
                                int stage_mode, float* dst_dev, float* mask_dev,
                                uint32_t* pad_count_dev, void* stream) {
  if (!h) return vstab_fail(nullptr, VSTAB_ERR_INVALID, "vstab_warp_fused: null handle");
  if (!src_dev || !fwd_dev || !dst_dev || !border_host)
    return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_warp_fused: null pointer argument");
  if (n < 0 || src_h <= 0 || src_w <= 0 || out_h <= 0 || out_w <= 0)
    return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_warp_fused: bad dimensions");
  if (samples < 1 || samples > MAX_SAMPLES)
    return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_warp_fused: samples must be in 1..33");
  if (interp != VSTAB_INTERP_BILINEAR && interp != VSTAB_INTERP_BICUBIC)
    return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_warp_fused: unknown interpolation");
  {
    const int r = mask_rule & 0xff;
    if (r != VSTAB_MASK_RULE_P && r != VSTAB_MASK_RULE_C && r != VSTAB_MASK_RULE_AUTO)
      return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_warp_fused: unknown mask rule");
  }
  if (src_w > 32767 || src_h > 32767)
    return vstab_fail(h, VSTAB_ERR_UNSUPPORTED, "vstab_warp_fused: source larger than 32767 px");
  if (n == 0) return VSTAB_OK;
  if (n > 65535) return vstab_fail(h, VSTAB_ERR_UNSUPPORTED, "vstab_warp_fused: n > 65535 frames per call");
  cudaStream_t st = (cudaStream_t)stream;
  VSTAB_ENTER(h);

  if (!g_cubic_tab_ready[h->device & 63]) {
    float tab[32][4];
    build_cubic_tab(tab);
    VSTAB_CUDA(h, cudaMemcpyToSymbolAsync(c_cubic_tab, tab, sizeof(tab), 0, cudaMemcpyHostToDevice, st));
    g_cubic_tab_ready[h->device & 63] = true;
  }
  if (pad_count_dev) VSTAB_CUDA(h, cudaMemsetAsync(pad_count_dev, 0, sizeof(uint32_t) * n, st));

  WarpParams p;
  p.src = src_dev;
  p.fwd = fwd_dev;
  p.dst = dst_dev;
  p.mask = mask_dev;
  p.pad_count = pad_count_dev;
  p.n = n;
  p.sh = src_h;
  p.sw = src_w;
  p.oh = out_h;
  p.ow = out_w;
  p.samples = samples;
  {
    const unsigned char* rules = nullptr;
    const int rc = resolve_mask_rule(h, fwd_dev, n * samples, src_h, src_w, out_h, out_w, mask_rule, st, &p.mask_rule, &rules);
    if (rc != VSTAB_OK) return rc;
    p.rules = rules;
  }
  p.stage_mode = stage_mode;
  p.vec_store = (out_w % 4 == 0) && (((uintptr_t)dst_dev & 15) == 0);
  p.vec_load = (src_w % 4 == 0) && (((uintptr_t)src_dev & 15) == 0);
  p.vec_mask = (out_w % 4 == 0) && (((uintptr_t)mask_dev & 15) == 0);
  p.border[0] = border_host[0];
  p.border[1] = border_host[1];
  p.border[2] = border_host[2];

  const size_t fixed = sizeof(double) * MINV_SLOTS * (9 + 8) + sizeof(float) * NWARPS * SCRATCH_FLOATS_PER_WARP + sizeof(int) * 8 + sizeof(float) * 128;
  // Staged source box: (TW + margin) x (TH + margin) pixels for near-identity maps; blur and
  // bicubic get a larger box.  Degenerate footprints gather from global memory instead.
  int box_w = TW + 8, box_h = TH + 8;
  if (samples > 1) {
    box_w += 24;
    box_h += 24;
  }
  size_t stage_bytes = (size_t)box_w * box_h * 3 * sizeof(float);
  if (stage_mode == VSTAB_STAGE_GLOBAL) stage_bytes = 16;
  p.stage_capacity = (int)(stage_bytes / sizeof(float));
  const size_t smem = fixed + stage_bytes;

  dim3 grid(vstab_ceil_div(out_w, TW), vstab_ceil_div(out_h, TH), n);
  if (grid.y > 65535) return vstab_fail(h, VSTAB_ERR_UNSUPPORTED, "vstab_warp_fused: output too tall");

  // Single-sample launches with 16-byte aligned source rows (every node path except motion blur) go
  // through the streaming kernel: plan pass + persistent CTAs fed by the TMA engine.
  static const bool stream_off = [] {
    const char* e = getenv("VSTAB_WARP_STREAM");
    return e && e[0] == '0';
  }();
  const unsigned long long s_tiles = (unsigned long long)grid.x * vstab_ceil_div(out_h, S_TH) * n;
  if (samples == 1 && p.vec_load && stage_mode == VSTAB_STAGE_AUTO && s_tiles < (1ull << 31) && src_w >= S_BW && src_h >= S_BH && out_w <= 65535 && out_h <= 65535 &&
      !stream_off && tensor_map_encoder()) {
    const int gx = (int)grid.x, gy = vstab_ceil_div(out_h, S_TH);
    CUtensorMap tmap;
    const cuuint64_t gdim[3] = {(cuuint64_t)src_w * 3, (cuuint64_t)src_h, (cuuint64_t)n};
    const cuuint64_t gstride[2] = {(cuuint64_t)src_w * 12, (cuuint64_t)src_w * 12 * (cuuint64_t)src_h};
    const cuuint32_t box[3] = {(cuuint32_t)S_PITCH, (cuuint32_t)S_BH, 1};
    const cuuint32_t estride[3] = {1, 1, 1};
    static const CUtensorMapL2promotion promo = [] {
      const char* e = getenv("VSTAB_TMA_PROMO");
      const int v = e ? atoi(e) : 2;
      return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }();
    const CUresult cr = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)src_dev, gdim, gstride, box, estride,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr == CUDA_SUCCESS) {
      const size_t plan_bytes = ((size_t)s_tiles * sizeof(TilePlan) + 15) & ~(size_t)15;
      const size_t need = plan_bytes + (size_t)n * 80 + 16;
      if (need > h->plan_bytes) {
        VSTAB_CUDA(h, cudaDeviceSynchronize());  // the old block may still be read by enqueued work
        if (h->plan) cudaFree(h->plan);
        h->plan = nullptr;
        h->plan_bytes = 0;
        if (cudaMalloc(&h->plan, need + need / 4) != cudaSuccess) return vstab_fail(h, VSTAB_ERR_NOMEM, "vstab_warp_fused: plan allocation failed");
        h->plan_bytes = need + need / 4;
      }
      TilePlan* plan = (TilePlan*)h->plan;
      double* minv = (double*)((char*)h->plan + plan_bytes);
      unsigned* ticket = (unsigned*)((char*)h->plan + plan_bytes + (size_t)n * 80);
      VSTAB_CUDA(h, cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
      const size_t s_smem = (size_t)S_STAGES * S_STAGE_BYTES + S_STAGES * 128 + 192 + sizeof(float) * 128 +
                            sizeof(float) * NWARPS * SCRATCH_FLOATS_PER_WARP;
      unsigned ctas = (unsigned)h->sm_count * S_CTAS;
      if ((unsigned long long)ctas > s_tiles) ctas = (unsigned)s_tiles;
      const unsigned plan_blocks = (unsigned)((s_tiles + 255) / 256);
      p.stage_capacity = S_BOX_BYTES / 4;
      if (interp == VSTAB_INTERP_BILINEAR) {
        warp_plan_kernel<VSTAB_INTERP_BILINEAR><<<plan_blocks, 256, 0, st>>>(fwd_dev, (unsigned)s_tiles, gx, gy, out_w, out_h, src_w, src_h, plan, minv);
        VSTAB_LAUNCH_CHECK(h, "warp_plan_kernel");
        VSTAB_CUDA(h, cudaFuncSetAttribute(warp_stream_kernel<VSTAB_INTERP_BILINEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s_smem));
        warp_stream_kernel<VSTAB_INTERP_BILINEAR><<<ctas, NTHREADS, s_smem, st>>>(tmap, p, plan, minv, ticket, (unsigned)s_tiles);
      } else {
        warp_plan_kernel<VSTAB_INTERP_BICUBIC><<<plan_blocks, 256, 0, st>>>(fwd_dev, (unsigned)s_tiles, gx, gy, out_w, out_h, src_w, src_h, plan, minv);
        VSTAB_LAUNCH_CHECK(h, "warp_plan_kernel");
        VSTAB_CUDA(h, cudaFuncSetAttribute(warp_stream_kernel<VSTAB_INTERP_BICUBIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s_smem));
        warp_stream_kernel<VSTAB_INTERP_BICUBIC><<<ctas, NTHREADS, s_smem, st>>>(tmap, p, plan, minv, ticket, (unsigned)s_tiles);
      }
      VSTAB_LAUNCH_CHECK(h, "warp_stream_kernel");
      return VSTAB_OK;
    }
  }
#define VSTAB_LAUNCH_FUSED(INTERP_, BLUR_)                                                                              \
  do {                                                                                                                 \
    VSTAB_CUDA(h, cudaFuncSetAttribute(warp_fused_kernel<INTERP_, BLUR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       (int)smem));                                                                    \
    warp_fused_kernel<INTERP_, BLUR_><<<grid, NTHREADS, smem, st>>>(p);                                                \
  } while (0)
  if (interp == VSTAB_INTERP_BILINEAR) {
    if (samples > 1) VSTAB_LAUNCH_FUSED(VSTAB_INTERP_BILINEAR, true);
    else VSTAB_LAUNCH_FUSED(VSTAB_INTERP_BILINEAR, false);
  } else {
    if (samples > 1) VSTAB_LAUNCH_FUSED(VSTAB_INTERP_BICUBIC, true);
    else VSTAB_LAUNCH_FUSED(VSTAB_INTERP_BICUBIC, false);
  }
#undef VSTAB_LAUNCH_FUSED
  VSTAB_LAUNCH_CHECK(h, "warp_fused_kernel");
  return VSTAB_OK;
}

extern "C" int vstab_common_coverage(vstab_handle* h, const float* fwd_dev, int n, int src_h,
                                     int src_w, int out_h, int out_w, int mask_rule,
                                     uint8_t* common_dev, void* stream) {
  if (!h) return vstab_fail(nullptr, VSTAB_ERR_INVALID, "vstab_common_coverage: null handle");
  if (!fwd_dev || !common_dev || n < 0 || src_h <= 0 || src_w <= 0 || out_h <= 0 || out_w <= 0)
    return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_common_coverage: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  VSTAB_ENTER(h);
  int rule = mask_rule;
  const unsigned char* rules = nullptr;
  {
    const int rc = resolve_mask_rule(h, fwd_dev, n, src_h, src_w, out_h, out_w, mask_rule, st, &rule, &rules);
    if (rc != VSTAB_OK) return rc;
  }
  dim3 grid(vstab_ceil_div(out_w, 32), vstab_ceil_div(out_h, 8));
  common_coverage_kernel<<<grid, 256, sizeof(double) * 64 * 9, st>>>(fwd_dev, n, src_h, src_w, out_h,
                                                                    out_w, rule, rules, common_dev);
  VSTAB_LAUNCH_CHECK(h, "common_coverage_kernel");
  return VSTAB_OK;
}

extern "C" int vstab_coverage_bbox(vstab_handle* h, const float* fwd_dev, int n, int src_h, int src_w, int out_h, int out_w,
                                   int mask_rule, int32_t* bbox_dev, void* stream) {
  if (!h) return vstab_fail(nullptr, VSTAB_ERR_INVALID, "vstab_coverage_bbox: null handle");
  if (!fwd_dev || !bbox_dev || n < 0 || src_h <= 0 || src_w <= 0 || out_h <= 0 || out_w <= 0)
    return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_coverage_bbox: bad argument");
  if (n == 0) return VSTAB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  VSTAB_ENTER(h);
  const int chunk = 64;
  void* ws = nullptr;
  int rc = vstab_workspace(h, (size_t)chunk * out_h * out_w, &ws);
  if (rc != VSTAB_OK) return rc;
  unsigned char* cov = (unsigned char*)ws;
  int rule = mask_rule;
  const unsigned char* rules = nullptr;
  rc = resolve_mask_rule(h, fwd_dev, n, src_h, src_w, out_h, out_w, mask_rule, st, &rule, &rules);
  if (rc != VSTAB_OK) return rc;
  bbox_init_kernel<<<vstab_ceil_div(n, 128), 128, 0, st>>>(bbox_dev, n);
  VSTAB_LAUNCH_CHECK(h, "bbox_init_kernel");
  for (int f0 = 0; f0 < n; f0 += chunk) {
    const int F = (n - f0) < chunk ? (n - f0) : chunk;
    dim3 grid(vstab_ceil_div(out_w, 32), vstab_ceil_div(out_h, 8), F);
    coverage_u8_kernel<<<grid, 256, 0, st>>>(fwd_dev + (size_t)f0 * 9, src_h, src_w, out_h, out_w, rule, rules ? rules + f0 : nullptr, cov);
    VSTAB_LAUNCH_CHECK(h, "coverage_u8_kernel");
    closed_bbox_kernel<<<grid, 256, 0, st>>>(cov, out_h, out_w, bbox_dev + (size_t)f0 * 4);
    VSTAB_LAUNCH_CHECK(h, "closed_bbox_kernel");
  }
  return VSTAB_OK;
}

// ---- binary padding mask as bytes (host results only) ------------------------------------------------------------
// A single-sample resampling writes mask values that are exactly 0.0f or 1.0f.  Results that go back to the host carry
// them as one byte per pixel over the link (a quarter of the float32 mask: 0.75 GB less per 121 x 1080p clip, ~10 % of the
// end-to-end time on a PCIe 5 x16 link) and are widened to the float32 MASK on the host while the frames are still being
// copied.  Any other value raises `odd` (the caller then fails loudly: a soft mask must travel as float32).
__global__ void __launch_bounds__(256) mask_pack_u8_kernel(const float* __restrict__ mask, size_t n, unsigned char* __restrict__ out,
                                                           unsigned int* __restrict__ odd) {
  const size_t n16 = n / 16;
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const float4* src = reinterpret_cast<const float4*>(mask) + 4 * i;
    unsigned int w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 v = __ldcs(src + k);
      bad |= (v.x != 0.f && v.x != 1.f) || (v.y != 0.f && v.y != 1.f) || (v.z != 0.f && v.z != 1.f) || (v.w != 0.f && v.w != 1.f);
      w[k] = (v.x != 0.f ? 1u : 0u) | (v.y != 0.f ? 1u << 8 : 0u) | (v.z != 0.f ? 1u << 16 : 0u) | (v.w != 0.f ? 1u << 24 : 0u);
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (size_t i = n16 * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = mask[i];
    bad |= v != 0.f && v != 1.f;
    out[i] = v != 0.f ? 1 : 0;
  }
  if (bad) atomicOr(odd, 1u);
}

extern "C" int vstab_mask_pack_u8(vstab_handle* h, const float* mask_dev, size_t n, uint8_t* out_dev, uint32_t* odd_dev, void* stream) {
  if (!h) return vstab_fail(nullptr, VSTAB_ERR_INVALID, "vstab_mask_pack_u8: null handle");
  if (!mask_dev || !out_dev || !odd_dev) return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_mask_pack_u8: bad argument");
  if (((uintptr_t)mask_dev & 15) || ((uintptr_t)out_dev & 15))
    return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_mask_pack_u8: buffers must be 16-byte aligned");
  if (n == 0) return VSTAB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  VSTAB_ENTER(h);
  const size_t want = (n / 16 + 255) / 256 + 1;
  const int blocks = (int)(want < (size_t)h->sm_count * 8 ? want : (size_t)h->sm_count * 8);
  mask_pack_u8_kernel<<<blocks, 256, 0, st>>>(mask_dev, n, out_dev, odd_dev);
  VSTAB_LAUNCH_CHECK(h, "mask_pack_u8_kernel");
  return VSTAB_OK;
}
