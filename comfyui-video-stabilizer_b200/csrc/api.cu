// api.cu -- handle lifecycle and error plumbing of libvstab.so (see include/vstab.h).
#include "common.cuh"

#include <math.h>
#include <stdlib.h>

char g_vstab_err[512] = "";

extern "C" int vstab_abi_version(void) { return VSTAB_ABI_VERSION; }

extern "C" int vstab_create(int device, vstab_handle** out) {
  if (!out) return vstab_fail(nullptr, VSTAB_ERR_INVALID, "vstab_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess)
    return vstab_fail(nullptr, VSTAB_ERR_CUDA, "vstab_create: no CUDA device (%s)", cudaGetErrorString(e));
  if (device < 0 || device >= count)
    return vstab_fail(nullptr, VSTAB_ERR_INVALID, "vstab_create: device index out of range");
  vstab_handle* h = (vstab_handle*)calloc(1, sizeof(vstab_handle));
  if (!h) return vstab_fail(nullptr, VSTAB_ERR_NOMEM, "vstab_create: host allocation failed");
  h->device = device;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    free(h);
    return vstab_fail(nullptr, VSTAB_ERR_CUDA, "vstab_create: cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  }
  if (prop.major < 10) {
    free(h);
    return vstab_fail(nullptr, VSTAB_ERR_UNSUPPORTED,
                      "vstab_create: libvstab is built for sm_100a (B200) only; device is %s", prop.name);
  }
  h->sm_count = prop.multiProcessorCount;
  h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  *out = h;
  return VSTAB_OK;
}

extern "C" void vstab_destroy(vstab_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->ws) cudaFree(h->ws);
  if (h->plan) cudaFree(h->plan);
  if (h->rules) cudaFree(h->rules);
  for (int i = 0; i < h->n_aux; ++i) {
    cudaStreamDestroy(h->aux_stream[i]);
    cudaEventDestroy(h->join_event[i]);
    cudaEventDestroy(h->stagger_event[i]);
  }
  if (h->fork_event) cudaEventDestroy(h->fork_event);
  for (int i = 0; i < h->n_level_events; ++i) cudaEventDestroy(h->level_event[i]);
  if (h->pyramid_event) cudaEventDestroy(h->pyramid_event);
  for (int i = 0; i < h->n_area_cache; ++i)
    if (h->area_cache[i].dev) cudaFree(h->area_cache[i].dev);
  free(h);
}

extern "C" const char* vstab_last_error(const vstab_handle* h) { return h ? h->err : g_vstab_err; }

extern "C" uint64_t vstab_launch_count(const vstab_handle* h) { return h ? h->launches : 0; }

// Host helper, no GPU involved: glibc libm over an array.  The trajectory solve between the estimation and the
// resampler runs in float64 on the host like the reference (nodes/stabilizer_utils.py:300-358 uses math.atan2 / log /
// exp / cos / sin per frame); at 1000-2000 frames per clip those per-element Python calls were the largest part of it
// on every rank.  Python's math module calls these very libm functions, so the results carry the same bits.
extern "C" int vstab_host_libm(int op, const double* a, const double* b, double* out, int n) {
  if (!a || !out || n < 0 || (op == VSTAB_LIBM_ATAN2 && !b)) return VSTAB_ERR_INVALID;
  switch (op) {
    case VSTAB_LIBM_ATAN2: for (int i = 0; i < n; ++i) out[i] = atan2(a[i], b[i]); break;
    case VSTAB_LIBM_LOG: for (int i = 0; i < n; ++i) out[i] = log(a[i]); break;
    case VSTAB_LIBM_EXP: for (int i = 0; i < n; ++i) out[i] = exp(a[i]); break;
    case VSTAB_LIBM_COS: for (int i = 0; i < n; ++i) out[i] = cos(a[i]); break;
    case VSTAB_LIBM_SIN: for (int i = 0; i < n; ++i) out[i] = sin(a[i]); break;
    default: return VSTAB_ERR_INVALID;
  }
  return VSTAB_OK;
}

int vstab_aux_streams(vstab_handle* h, int n) {
  if (n > VSTAB_MAX_AUX_STREAMS) return vstab_fail(h, VSTAB_ERR_INVALID, "vstab_aux_streams: too many helper streams");
  if (!h->fork_event) VSTAB_CUDA(h, cudaEventCreateWithFlags(&h->fork_event, cudaEventDisableTiming));
  while (h->n_aux < n) {
    VSTAB_CUDA(h, cudaStreamCreateWithFlags(&h->aux_stream[h->n_aux], cudaStreamNonBlocking));
    VSTAB_CUDA(h, cudaEventCreateWithFlags(&h->join_event[h->n_aux], cudaEventDisableTiming));
    VSTAB_CUDA(h, cudaEventCreateWithFlags(&h->stagger_event[h->n_aux], cudaEventDisableTiming));
    h->n_aux++;
  }
  while (h->n_level_events < VSTAB_MAX_LEVEL_EVENTS) {
    VSTAB_CUDA(h, cudaEventCreateWithFlags(&h->level_event[h->n_level_events], cudaEventDisableTiming));
    h->n_level_events++;
  }
  if (!h->pyramid_event) VSTAB_CUDA(h, cudaEventCreateWithFlags(&h->pyramid_event, cudaEventDisableTiming));
  return VSTAB_OK;
}

int vstab_workspace(vstab_handle* h, size_t bytes, void** out) {
  if (bytes > h->ws_bytes) {
    // grow-only; the old block may still be in use by enqueued work, so drain first
    VSTAB_CUDA(h, cudaDeviceSynchronize());
    if (h->ws) cudaFree(h->ws);
    h->ws = nullptr;
    h->ws_bytes = 0;
    const size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&h->ws, want);
    if (e != cudaSuccess) return vstab_fail(h, VSTAB_ERR_NOMEM, "workspace cudaMalloc failed: %s", cudaGetErrorString(e));
    h->ws_bytes = want;
  }
  *out = h->ws;
  return VSTAB_OK;
}
