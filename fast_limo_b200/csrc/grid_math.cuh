// Cell arithmetic shared by the index build (map_index.cu) and the search (match_kernel.cu).
// Both MUST use these exact expressions: the search's exactness argument relies on a point and a
// query being binned by the same float computation (see knn_search() in match_kernel.cu).
#pragma once
#include <cuda_runtime.h>

namespace flimo {

// Position in cell units (float).  Written with explicit round-to-nearest intrinsics so that the
// result does not depend on the translation unit's -fmad setting.
__device__ __forceinline__ float cell_units(float x, float origin, float inv_cell) {
  return __fmul_rn(__fadd_rn(x, -origin), inv_cell);
}

// Clamped integer cell coordinate.  NaN maps to cell 0.
__device__ __forceinline__ int cell_coord(float x, float origin, float inv_cell, int n) {
  const float u = cell_units(x, origin, inv_cell);
  int c = (u >= 0.f) ? (int)fminf(floorf(u), 2.0e9f) : 0;   // also sends NaN to 0
  return c < n - 1 ? c : n - 1;
}

}  // namespace flimo
