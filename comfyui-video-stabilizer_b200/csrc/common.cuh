// common.cuh -- handle, error plumbing and small device helpers shared by the vstab kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vstab.h"

#include "area.cuh"

#define VSTAB_AREA_CACHE 32
#define VSTAB_MAX_AUX_STREAMS 4
#define VSTAB_MAX_LEVEL_EVENTS 8
struct vstab_area_cache_entry {
  int ssize, dsize;
  void* dev;
  vstab_area_tab tab;
};

struct vstab_handle {
  int device;
  int sm_count;
  int max_smem_optin;
  uint64_t launches;
  char err[512];
  // grow-only device workspace (DIS pyramids, fit scratch)
  void* ws;
  size_t ws_bytes;
  // grow-only tile plan of the streaming resampler
  void* plan;
  size_t plan_bytes;
  vstab_area_cache_entry area_cache[VSTAB_AREA_CACHE];
  int n_area_cache;
  int area_evict;  // next slot to recycle once the cache is full
  // grow-only per-(frame, sample) mask rules of VSTAB_MASK_RULE_AUTO launches
  void* rules;
  size_t rules_bytes;
  // helper streams for work that forks from / joins back into the caller's stream inside one call
  // (DIS pair groups); created on first use
  cudaStream_t aux_stream[VSTAB_MAX_AUX_STREAMS];
  cudaEvent_t join_event[VSTAB_MAX_AUX_STREAMS];
  cudaEvent_t stagger_event[VSTAB_MAX_AUX_STREAMS];
  cudaEvent_t fork_event;
  int n_aux;
  // DIS: per-level "gradients and structure tensors are ready" events of the preparation stream, created on first use
  cudaEvent_t level_event[VSTAB_MAX_LEVEL_EVENTS];
  cudaEvent_t pyramid_event;
  int n_level_events;
};

extern char g_vstab_err[512];

static inline int vstab_fail(vstab_handle* h, int code, const char* fmt, const char* a = "",
                             const char* b = "") {
  char* dst = h ? h->err : g_vstab_err;
  snprintf(dst, 512, fmt, a, b);
  if (h) snprintf(g_vstab_err, 512, "%s", dst);
  return code;
}

#define VSTAB_CUDA(h, expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return vstab_fail((h), VSTAB_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define VSTAB_LAUNCH_CHECK(h, name)                                                      \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess)                                                               \
      return vstab_fail((h), VSTAB_ERR_CUDA, "launch of %s failed: %s", name,            \
                        cudaGetErrorString(_e));                                         \
    (h)->launches++;                                                                     \
  } while (0)

// Entry of every compute call: select the handle's device and drop any error another library left
// sticky on this host thread, so that the launch checks below only ever report our own launches.
#define VSTAB_ENTER(h)                            \
  do {                                            \
    VSTAB_CUDA((h), cudaSetDevice((h)->device));  \
    (void)cudaGetLastError();                     \
  } while (0)

// Returns a device workspace of at least `bytes` (grow-only; synchronises only when growing).
int vstab_workspace(vstab_handle* h, size_t bytes, void** out);

// Makes sure aux_stream[0..n) / join_event[0..n) / fork_event exist (n <= VSTAB_MAX_AUX_STREAMS).
int vstab_aux_streams(vstab_handle* h, int n);

static inline int vstab_ceil_div(int a, int b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// 3x3 inverse by cofactors in double, same operation order as cv::invert.
// Explicit _rn intrinsics: no FMA contraction, so host (numpy) and device agree bit for bit.
__device__ __forceinline__ void vstab_invert3(const float* __restrict__ m32, double* __restrict__ o) {
  const double a = m32[0], b = m32[1], c = m32[2];
  const double d = m32[3], e = m32[4], f = m32[5];
  const double g = m32[6], hh = m32[7], i = m32[8];
  auto mul = [](double x, double y) { return __dmul_rn(x, y); };
  auto sub = [](double x, double y) { return __dsub_rn(x, y); };
  const double c0 = sub(mul(e, i), mul(f, hh));
  const double c1 = sub(mul(d, i), mul(f, g));
  const double c2 = sub(mul(d, hh), mul(e, g));
  const double det = __dadd_rn(sub(mul(a, c0), mul(b, c1)), mul(c, c2));
  if (det == 0.0) {
    for (int k = 0; k < 9; ++k) o[k] = 0.0;
    return;
  }
  const double r = __ddiv_rn(1.0, det);
  o[0] = mul(c0, r);
  o[1] = mul(sub(mul(c, hh), mul(b, i)), r);
  o[2] = mul(sub(mul(b, f), mul(c, e)), r);
  o[3] = mul(sub(mul(f, g), mul(d, i)), r);
  o[4] = mul(sub(mul(a, i), mul(c, g)), r);
  o[5] = mul(sub(mul(c, d), mul(a, f)), r);
  o[6] = mul(c2, r);
  o[7] = mul(sub(mul(b, g), mul(a, hh)), r);
  o[8] = mul(sub(mul(a, e), mul(b, d)), r);
}
#endif
