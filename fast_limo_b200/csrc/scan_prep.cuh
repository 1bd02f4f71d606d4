// Declarations of the device scan preparation (scan_prep.cu): filters, time sort, deskew, voxel grid.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/flimo.h"

namespace flimo {

struct PrepDev {                 // kernel-parameter form of flimo_prep_cfg (+ the sweep reference time)
  int crop_active, dist_active, rate_active, fov_active;
  float crop_min[3], crop_max[3];
  float min_dist;                // static_cast<float>(Config::filters::min_dist), Localizer.cpp:274
  int rate_value;
  float fov_angle;
  int sensor_type, end_of_sweep;
  double sweep_ref_time;
};

struct DeskewDev {
  double offset;                 // IMU / LiDAR time offset (Localizer.cpp:797-801)
  int n_frames;
  float T_l2b[16];               // extr.lidar2baselink_T, row-major
  float Tinv[16];                // last_state.get_RT_inv(), row-major
};

struct VoxelDev {                // device-resident voxel grid description (pcl::VoxelGrid members)
  unsigned int mn[3], mx[3];     // bounding box, order-preserving float encoding
  int min_b[3], mul[3];
  float inv_leaf;
  int overflow;
};

struct PrepBuffers {
  unsigned char* raw = nullptr;          // raw cloud, 32-byte fast_limo::Point records
  uint32_t* u32 = nullptr;               // 10 scratch arrays of `cap` uint32
  unsigned long long* keys64 = nullptr;  // 2 x cap sort keys
  double* t_sorted = nullptr;            // point times, sorted order
  float4 *world = nullptr, *xt2 = nullptr, *vox_out = nullptr;
  flimo_frame* frames = nullptr;
  size_t frames_cap = 0;
  void* cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  unsigned char* small = nullptr;        // counters + VoxelDev
  size_t cap = 0;
  const uint32_t* order = nullptr;       // time-sorted indices into raw (inside u32)
  size_t n_sorted = 0;
};

cudaError_t prep_reserve(PrepBuffers& b, size_t n);
// sensor_msgs/PointCloud2 payload (device copy) -> 32-byte fast_limo::Point records in b.raw
cudaError_t prep_decode_msg(PrepBuffers& b, const unsigned char* d_msg, size_t n, size_t point_step, const flimo_msg_layout& L,
                            cudaStream_t st, uint64_t* launches);
void prep_free(PrepBuffers& b);
cudaError_t prep_filter_sort(PrepBuffers& b, size_t n, const PrepDev& c, cudaStream_t st, uint32_t* n_kept, double* t_last,
                             uint64_t* launches);
cudaError_t prep_deskew(PrepBuffers& b, const DeskewDev& d, const flimo_frame* h_frames, bool keep_world, cudaStream_t st,
                        uint64_t* launches);
cudaError_t prep_voxel(PrepBuffers& b, const float4* in, uint32_t cap_n, const uint32_t* d_n, float leaf, cudaStream_t st,
                       uint32_t* n_out_host, bool* passthrough, uint64_t* launches);

}  // namespace flimo
