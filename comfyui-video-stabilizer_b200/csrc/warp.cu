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
// Tiling: a CTA of 256 threads owns a 64x32 output tile (two groups of 16 rows).  In a group warp w
// owns rows w and w+8, lane l owns columns l and l+32 (stride-1 lanes => conflict-free
// shared-memory gathers, 3-word stride).  The source footprint of the tile (bounding box of the
// four projected corners over all samples, plus the tap margin) is staged once into shared memory
// by the TMA engine: one cp.async.bulk per source row, all completing on one mbarrier, issued by
// warp 0 while the other warps already compute their coordinates.  Taps that fall outside the
// staged box (degenerate maps) fall back to a global load, so the staging is a pure optimisation
// and never changes results.  Interior tiles (footprint >= 1 px inside the source, one sample,
// bilinear) take a branch-free path without any per-tap or per-pixel tests.
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
  int stage_mode;
  int stage_capacity;  // floats available for the staged source tile
  int vec_store;       // ow % 4 == 0 && dst 16B aligned
  int vec_load;        // sw % 4 == 0 && src 16B aligned
  int vec_mask;        // ow % 4 == 0 && mask 16B aligned
  float border[3];
};

__constant__ float c_cubic_tab[32][4];

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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
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
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  unsigned done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(phase)
                 : "memory");
  } while (!done);
}

__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// Interior tile, single sample, bilinear: every tap is inside the staged box and every pixel is
// covered, so the loop body is coordinates -> 12 shared-memory loads -> blend, nothing else.
// Column / row products of the inverse matrix are hoisted (2 columns x 2 rows per thread).
template <bool AFFINE, bool VEC>
__device__ __forceinline__ void interior_tile(const double* __restrict__ s_minv, const float* __restrict__ tile0, int pitch,
                                              int tx0, int ty0, int warp, int lane, int tid, int ow,
                                              float* __restrict__ scratch, float* __restrict__ dst_tile,
                                              float* __restrict__ mask_tile, int vec_mask, unsigned long long* bar,
                                              bool bulk_pending) {
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
  int ixs[4 * GROUPS], iys[4 * GROUPS];
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
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
    for (int g = 0; g < GROUPS; ++g) {
      float* mrow = mask_tile + (size_t)(g * 16 + (tid >> 4)) * ow + (tid & 15) * 4;
      if (vec_mask) {
        *reinterpret_cast<float4*>(mrow) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        mrow[0] = 0.f; mrow[1] = 0.f; mrow[2] = 0.f; mrow[3] = 0.f;
      }
    }
  }
  // the source box has been streaming into shared memory meanwhile
  if (bulk_pending) mbar_wait(bar, 0); else __syncthreads();
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
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

template <int INTERP>
__global__ void __launch_bounds__(NTHREADS, MIN_CTAS) warp_fused_kernel(const WarpParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_minv = reinterpret_cast<double*>(smem_raw);                       // [34][9], 16B multiple
  float* s_scratch = reinterpret_cast<float*>(s_minv + MINV_SLOTS * 9);       // [8][384]
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
  bool bulk_pending = false;  // CTA-uniform: the staged box arrives through bulk copies
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
    if (tid < S) vstab_invert3(p.fwd + ((size_t)frame_idx * S + tid) * 9, s_minv + tid * 9);
    if (tid == 0) {
      s_box[0] = INT_MAX;  // min x
      s_box[1] = INT_MAX;  // min y
      s_box[2] = INT_MIN;  // max x
      s_box[3] = INT_MIN;  // max y
      s_box[4] = 0;        // degenerate flag
    }
    __syncthreads();
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
        interior = (S == 1) && (INTERP == VSTAB_INTERP_BILINEAR) && (tx0 + TW <= p.ow) && (ty0 + TH <= p.oh) &&
                   (s_box[0] - LO >= 1) && (s_box[1] - LO >= 1) && (s_box[2] + HI <= p.sw - 2) &&
                   (s_box[3] + HI <= p.sh - 2);
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
      if (affine) interior_tile<true, true>(s_minv, tile0, tile.pitch, tx0, ty0, warp, lane, tid, p.ow, scratch, dst_tile, mask_tile, p.vec_mask, s_bar, bulk_pending);
      else interior_tile<false, true>(s_minv, tile0, tile.pitch, tx0, ty0, warp, lane, tid, p.ow, scratch, dst_tile, mask_tile, p.vec_mask, s_bar, bulk_pending);
    } else {
      if (affine) interior_tile<true, false>(s_minv, tile0, tile.pitch, tx0, ty0, warp, lane, tid, p.ow, scratch, dst_tile, mask_tile, p.vec_mask, s_bar, bulk_pending);
      else interior_tile<false, false>(s_minv, tile0, tile.pitch, tx0, ty0, warp, lane, tid, p.ow, scratch, dst_tile, mask_tile, p.vec_mask, s_bar, bulk_pending);
    }
    return;
  }
  if (bulk_pending) mbar_wait(s_bar, 0);
  __syncthreads();  // also orders the scalar staging path and the s_minv / s_cubic writes

  // ---- per-pixel resampling ------------------------------------------------------------------
  // Sample-outer loop: the per-sample row/column products of the inverse matrix are hoisted out of
  // the 4 pixels a thread owns; the float32 accumulation per pixel is still in sample order.
  const float fS = (float)S;
  float* scratch = s_scratch + warp * SCRATCH_FLOATS_PER_WARP;
  unsigned int padded = 0;
  const double x_hi = (double)(p.sw - 1), y_hi = (double)(p.sh - 1);
  const bool t_on = tile.active;
  const int t_x0 = tile.x0, t_y0 = tile.y0, t_x1 = tile.x1, t_y1 = tile.y1, t_pitch = tile.pitch;
  const double dxs[2] = {(double)(tx0 + lane), (double)(tx0 + lane + 32)};

  for (int g = 0; g < GROUPS; ++g) {  // row groups of 16 output rows
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
        if (p.mask_rule == VSTAB_MASK_RULE_P && qx > band && qx < x_hi - band && qy > band && qy < y_hi - band) {
          ok = true;
        } else if (p.mask_rule == VSTAB_MASK_RULE_P && (qx < -band || qx > x_hi + band || qy < -band || qy > y_hi + band)) {
          ok = false;
        } else {
          double cxs = __ddiv_rn(X, W), cys = __ddiv_rn(Y, W);
          if (p.mask_rule == VSTAB_MASK_RULE_C) {
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
          const float* s0 = s_tile + (sy - t_y0) * t_pitch + (sx - t_x0) * 3;
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
          vr = vg = vb = 0.f;
          if (t_on && bxs >= t_x0 && bxs + 3 <= t_x1 && bys >= t_y0 && bys + 3 <= t_y1) {
            const float* s0 = s_tile + (bys - t_y0) * t_pitch + (bxs - t_x0) * 3;
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) {
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) {
                const float w = __fmul_rn(wy[k1], wx[k2]);
                const float* t = s0 + k1 * t_pitch + k2 * 3;
                vr = __fadd_rn(vr, __fmul_rn(t[0], w));
                vg = __fadd_rn(vg, __fmul_rn(t[1], w));
                vb = __fadd_rn(vb, __fmul_rn(t[2], w));
              }
            }
          } else {
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) {
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) {
                const float w = __fmul_rn(wy[k1], wx[k2]);
                float r, g, b;
                fetch_rgb(p, frame, tile, bys + k1, bxs + k2, r, g, b);
                vr = __fadd_rn(vr, __fmul_rn(r, w));
                vg = __fadd_rn(vg, __fmul_rn(g, w));
                vb = __fadd_rn(vb, __fmul_rn(b, w));
              }
            }
          }
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

// AND-reduction of INTER_NEAREST coverage over n matrices (crop solvers).
__global__ void __launch_bounds__(256) common_coverage_kernel(const float* __restrict__ fwd, int n,
                                                              int sh, int sw, int oh, int ow,
                                                              int mask_rule,
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
        if (mask_rule == VSTAB_MASK_RULE_C) {
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
                                                          int mask_rule, unsigned char* __restrict__ cov) {
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
  if (mask_rule == VSTAB_MASK_RULE_C) {
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
  if (mask_rule != VSTAB_MASK_RULE_P && mask_rule != VSTAB_MASK_RULE_C)
    return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_warp_fused: unknown mask rule");
  if (src_w > 32767 || src_h > 32767)
    return vstab_fail(h, VSTAB_ERR_UNSUPPORTED, "vstab_warp_fused: source larger than 32767 px");
  if (n == 0) return VSTAB_OK;
  if (n > 65535) return vstab_fail(h, VSTAB_ERR_UNSUPPORTED, "vstab_warp_fused: n > 65535 frames per call");
  cudaStream_t st = (cudaStream_t)stream;
  VSTAB_CUDA(h, cudaSetDevice(h->device));

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
  p.mask_rule = mask_rule;
  p.stage_mode = stage_mode;
  p.vec_store = (out_w % 4 == 0) && (((uintptr_t)dst_dev & 15) == 0);
  p.vec_load = (src_w % 4 == 0) && (((uintptr_t)src_dev & 15) == 0);
  p.vec_mask = (out_w % 4 == 0) && (((uintptr_t)mask_dev & 15) == 0);
  p.border[0] = border_host[0];
  p.border[1] = border_host[1];
  p.border[2] = border_host[2];

  const size_t fixed = sizeof(double) * MINV_SLOTS * 9 + sizeof(float) * NWARPS * SCRATCH_FLOATS_PER_WARP + sizeof(int) * 8 + sizeof(float) * 128;
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
  if (interp == VSTAB_INTERP_BILINEAR) {
    VSTAB_CUDA(h, cudaFuncSetAttribute(warp_fused_kernel<VSTAB_INTERP_BILINEAR>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    warp_fused_kernel<VSTAB_INTERP_BILINEAR><<<grid, NTHREADS, smem, st>>>(p);
  } else {
    VSTAB_CUDA(h, cudaFuncSetAttribute(warp_fused_kernel<VSTAB_INTERP_BICUBIC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    warp_fused_kernel<VSTAB_INTERP_BICUBIC><<<grid, NTHREADS, smem, st>>>(p);
  }
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
  VSTAB_CUDA(h, cudaSetDevice(h->device));
  dim3 grid(vstab_ceil_div(out_w, 32), vstab_ceil_div(out_h, 8));
  common_coverage_kernel<<<grid, 256, sizeof(double) * 64 * 9, st>>>(fwd_dev, n, src_h, src_w, out_h,
                                                                    out_w, mask_rule, common_dev);
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
  VSTAB_CUDA(h, cudaSetDevice(h->device));
  const int chunk = 64;
  void* ws = nullptr;
  int rc = vstab_workspace(h, (size_t)chunk * out_h * out_w, &ws);
  if (rc != VSTAB_OK) return rc;
  unsigned char* cov = (unsigned char*)ws;
  bbox_init_kernel<<<vstab_ceil_div(n, 128), 128, 0, st>>>(bbox_dev, n);
  VSTAB_LAUNCH_CHECK(h, "bbox_init_kernel");
  for (int f0 = 0; f0 < n; f0 += chunk) {
    const int F = (n - f0) < chunk ? (n - f0) : chunk;
    dim3 grid(vstab_ceil_div(out_w, 32), vstab_ceil_div(out_h, 8), F);
    coverage_u8_kernel<<<grid, 256, 0, st>>>(fwd_dev + (size_t)f0 * 9, src_h, src_w, out_h, out_w, mask_rule, cov);
    VSTAB_LAUNCH_CHECK(h, "coverage_u8_kernel");
    closed_bbox_kernel<<<grid, 256, 0, st>>>(cov, out_h, out_w, bbox_dev + (size_t)f0 * 4);
    VSTAB_LAUNCH_CHECK(h, "closed_bbox_kernel");
  }
  return VSTAB_OK;
}
