// Host emulation of the two variational-refinement kernels of csrc/dis.cu (test infrastructure, CPU only).
//
// The build script cuts the regions of dis.cu marked //@emul-begin ... //@emul-end into vr_regions.inc; this file
// supplies stand-ins for the CUDA execution model -- one host thread per CUDA thread, a pthread barrier for
// __syncthreads / barrier.cluster, plain arrays for (distributed) shared memory -- and runs the SAME kernel source on
// random levels: vr_fused_kernel<false> (the cluster kernel whose GPU results are pinned bit-exact to cv2) and
// vr_resident_kernel (shared-memory / register resident) must leave identical bits in Ux / Uy.
// Compiled with -ffp-contract=off: sqrtf, / and the unfused products round like the device code built with -fmad=false.
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#define VSTAB_HOST_EMUL 1
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n)

struct EmulDim { unsigned x, y, z; };
static thread_local EmulDim threadIdx, blockIdx, blockDim;
static thread_local float* tl_smem = nullptr;     // this CTA's shared memory
static thread_local float** tl_cluster = nullptr; // shared memory of every CTA of the cluster, by rank
static pthread_barrier_t* g_barrier = nullptr;

using std::max;
using std::min;
static inline int __float2int_rn(float v) { return (int)lrintf(v); }  // round-to-nearest-even, the default mode
static inline void __syncthreads() { pthread_barrier_wait(g_barrier); }  // the whole cluster: a superset of the CTA
static inline void cluster_barrier() { pthread_barrier_wait(g_barrier); }
static inline void st_cluster(float* local, unsigned rank, float v) { tl_cluster[rank][local - tl_smem] = v; }
#define VSTAB_DYNAMIC_SMEM(name) float* name = tl_smem

namespace {
#include "vr_regions.inc"
}

typedef void (*Kernel)(Level, VrBuf, int);

static void launch(Kernel kernel, Level L, VrBuf B, int cluster, int threads, size_t smem_floats) {
  std::vector<std::vector<float>> smem(cluster, std::vector<float>(smem_floats + 1, 123.f));
  std::vector<float*> ptrs(cluster);
  for (int c = 0; c < cluster; c++) ptrs[c] = smem[c].data();
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, cluster * threads);
  g_barrier = &bar;
  std::vector<std::thread> pool;
  pool.reserve((size_t)cluster * threads);
  for (int c = 0; c < cluster; c++)
    for (int t = 0; t < threads; t++)
      pool.emplace_back([=, &ptrs]() {
        threadIdx = {(unsigned)t, 0, 0};
        blockIdx = {(unsigned)c, 0, 0};  // pair 0
        blockDim = {(unsigned)threads, 1, 1};
        tl_smem = ptrs[c];
        tl_cluster = const_cast<float**>(ptrs.data());
        kernel(L, B, cluster);
      });
  for (auto& th : pool) th.join();
  pthread_barrier_destroy(&bar);
}

static uint32_t rng_state = 12345;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }
static float frand() { return (float)rnd() / 16777216.f; }

static int run_case(int w, int h, int cl_old, int cl_new, int threads_new) {
  const size_t n = (size_t)w * h;
  std::vector<unsigned char> I(2 * n);
  // a smooth texture plus noise, second frame shifted: real gradients and a flow that is not trivial
  for (int f = 0; f < 2; f++)
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        const float v = 128.f + 60.f * sinf(0.21f * (x + 1.3f * f)) * cosf(0.17f * (y - 0.8f * f)) + 30.f * sinf(0.05f * x * y * 0.1f) + 12.f * (frand() - 0.5f);
        I[f * n + (size_t)y * w + x] = (unsigned char)std::min(255.f, std::max(0.f, v));
      }
  std::vector<float> ux0(n), uy0(n);
  for (size_t i = 0; i < n; i++) { ux0[i] = 1.3f + 0.8f * (frand() - 0.5f); uy0[i] = -0.8f + 0.8f * (frand() - 0.5f); }
  ux0[3] = -0.0f;  // signed zeros take the same route in both kernels
  uy0[n - 2] = 0.0f;

  auto run = [&](bool resident, std::vector<float>& ux, std::vector<float>& uy) {
    ux = ux0; uy = uy0;
    std::vector<std::vector<float>> planes(19, std::vector<float>(n, resident ? 777.f : -555.f));  // garbage: nothing may depend on it
    Level L = {};
    L.w = w; L.h = h;
    L.I = I.data(); L.Ux = ux.data(); L.Uy = uy.data();
    VrBuf B;
    float** f = (float**)&B;
    for (int k = 0; k < 19; k++) f[k] = planes[k].data();
    if (!resident) {
      launch(vr_fused_kernel<false>, L, B, cl_old, 256, 0);
    } else {
      const int rows_per = (h + cl_new - 1) / cl_new, half_w = (w + 1) / 2;
      if (rows_per > 127 || rows_per * half_w > kResCells || (rows_per + 2) * (half_w + 2) > kResColour) { printf("case %dx%d cl %d does not fit the resident kernel\n", w, h, cl_new); exit(2); }
      const size_t floats = kResSmemFloats;
      Kernel k = cl_new == 1 ? (threads_new == 512 ? vr_resident_kernel<false, 512> : vr_resident_kernel<false, 1024>)
                             : (threads_new == 512 ? vr_resident_kernel<true, 512> : vr_resident_kernel<true, 1024>);
      launch(k, L, B, cl_new, threads_new, floats);
    }
  };
  std::vector<float> ax, ay, bx, by;
  run(false, ax, ay);
  run(true, bx, by);
  size_t bad = 0, first = n;
  double moved = 0;
  for (size_t i = 0; i < n; i++) {
    if (memcmp(&ax[i], &bx[i], 4) || memcmp(&ay[i], &by[i], 4)) { if (!bad) first = i; bad++; }
    moved += fabs(ax[i] - ux0[i]) + fabs(ay[i] - uy0[i]);
  }
  printf("%dx%d cluster %d -> resident cluster %d x %d threads: %zu of %zu pixels differ%s, mean |update| %.4f\n", w, h, cl_old, cl_new,
         threads_new, bad, n, bad ? " (FAIL)" : "", moved / (2 * n));
  if (bad) printf("  first at (%zu, %zu): %.9g %.9g vs %.9g %.9g\n", first % w, first / w, ax[first], ay[first], bx[first], by[first]);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  int fails = 0;
  if (argc == 6) return run_case(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]));
  fails += run_case(120, 67, 8, 1, 512);    // level 3 of a 960x540 working image: one CTA per pair
  fails += run_case(61, 34, 2, 1, 512);     // odd width
  fails += run_case(101, 57, 4, 2, 512);    // two bands, odd sizes
  fails += run_case(240, 135, 8, 4, 512);   // finest level: four bands
  fails += run_case(50, 33, 2, 8, 512);     // more bands than needed: thin and empty bands
  return fails ? 1 : 0;
}
