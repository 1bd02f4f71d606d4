// Shared host/device declarations of libflimo_cuda (B200 / sm_100a registration hot path).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace flimo {

// Uniform search grid over the map's bounding box.  Cell (ix,iy,iz) has linear index
// (iz*ny + iy)*nx + ix; map points are stored sorted by that index (x fastest), so the cells
// [ix0..ix1] of one (iy,iz) row are ONE contiguous run of points:
//   [cell_start[row + ix0], cell_start[row + ix1 + 1]).
struct GridDesc {
  float ox, oy, oz;   // lower corner
  float inv_cell;     // 1 / cell (float)
  float cell;         // cell side in metres
  int nx, ny, nz;
};

// Constants of one measurement pass.  Built on the host from the double filter state exactly the
// way the reference builds them (State.cpp:38-55,136-172; Localizer.cpp:554-555).
struct PoseConsts {
  float R_wb[9], t_wb[3];         // State::get_RT()             (float-cast quaternion)
  float Rinv_wb[9], tinv_wb[3];   // State::get_RT_inv()
  float Rinv_LI[9], tinv_LI[3];   // State::get_extr_RT_inv()
  float Rd_wb_inv[9];             // s.rot.conjugate().toRotationMatrix().cast<float>()
  float Rd_LI_inv[9];             // s.offset_R_L_I.conjugate().toRotationMatrix().cast<float>()
};

constexpr int kTileQueries = 128;     // queries per CTA tile (== threads per CTA)
constexpr int kTriEntries = 91;       // upper triangle of the 13x13 outer product of [row(12), z]
constexpr int kPartialStride = 96;    // doubles per partial record

struct MatchParams {
  const float4* scan;          // body-frame points; .w carries the original scan index (bit pattern)
  const float4* map;           // world-frame map points sorted by cell
  const uint32_t* cell_start;  // nx*ny*nz + 1 prefix offsets
  GridDesc g;
  PoseConsts pc;
  int q_begin, q_end;          // slice of the scan handled by this launch
  float max_dist_f;            // smallest float >= MAX_DIST_PLANE  (d2_5 < MAX_DIST_PLANE test)
  float plane_thr;             // (float)PLANE_THRESHOLD
  int estimate_extrinsics;
  uint32_t orig_limit;         // rows contribute only if original index < orig_limit
  double* partials;            // [n_tiles][96] scratch
  unsigned int* ticket;        // last-CTA-done counter (self resetting)
  double* out96;               // packed result (see flimo.h)
  float* dbg16;                // optional per-point record [n][16], indexed by original index
  uint8_t* valid_by_orig;      // optional accepted flag per original index
};

// map_index.cu
struct MapIndex {
  float4* pts = nullptr;        // sorted map points
  uint32_t* cell_start = nullptr;
  size_t n_pts = 0, cap_pts = 0;
  size_t n_cells = 0, cap_cells = 0;
  GridDesc g{};
  // scratch
  uint32_t *keys = nullptr, *keys_alt = nullptr, *vals = nullptr, *vals_alt = nullptr;
  float4* pts_alt = nullptr;
  size_t cap_scratch = 0;
  void* cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  float* bbox = nullptr;        // 6 floats on device
};

// Rebuilds the index from `n` device points (float4, any order) stored in idx.pts_alt[0..n):
// computes the bounding box, picks the grid, sorts by cell, builds cell_start.
// cell <= 0 selects the side automatically from the point density.  Returns cudaError_t.
cudaError_t map_index_build(MapIndex& idx, size_t n, float cell, size_t max_cells, cudaStream_t st,
                            uint64_t* launches);
cudaError_t map_index_reserve(MapIndex& idx, size_t n_pts);
void map_index_free(MapIndex& idx);

// Packs strided xyz (device) into float4 with w = 0, dropping NaN points; writes the count.
cudaError_t pack_points(const void* d_src, size_t n, size_t stride_bytes, float4* dst, unsigned int* d_count,
                        cudaStream_t st);
// Packs a scan: float4 with w = original index bits (no NaN filtering: the reference matches them
// and they simply fail the kNN gate).
cudaError_t pack_scan(const void* d_src, size_t n, size_t stride_bytes, float4* dst, cudaStream_t st);
cudaError_t sort_scan_morton(float4* scan, float4* tmp, size_t n, void** cub_tmp, size_t* cub_tmp_bytes,
                             uint32_t** keys, size_t* keys_cap, cudaStream_t st, uint64_t* launches);
cudaError_t transform_scan(const float4* scan, size_t n, const PoseConsts& pc, float* d_out_xyz, cudaStream_t st);

// match_kernel.cu
cudaError_t launch_match(const MatchParams& p, cudaStream_t st);
int match_num_tiles(int n_queries);

}  // namespace flimo
