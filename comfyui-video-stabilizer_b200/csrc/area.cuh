// area.cuh -- INTER_AREA coverage tables (cv::computeResizeAreaTab restated) and the uint8
// area-resize launcher shared by gray.cu and dis.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

struct vstab_handle;

// CSR over destination coordinates: entries start[d] .. start[d+1]-1 give (si, alpha).
struct vstab_area_tab {
  const int* start;
  const int* si;
  const float* alpha;
};

// Looks the (ssize -> dsize) table up in the handle's cache, building and uploading it on a miss
// (one synchronous copy the first time a size pair is seen).
int vstab_area_tab_get(vstab_handle* h, int ssize, int dsize, vstab_area_tab* out);

// dst[n][dh][dw] = cv2.resize(src[n][sh][sw], (dw, dh), INTER_AREA), uint8, bit-exact.
int vstab_area_u8(vstab_handle* h, const unsigned char* src, int n, int sh, int sw,
                  unsigned char* dst, int dh, int dw, cudaStream_t st);
