// Device side of fast_limo::Localizer::updatePointCloud (fast_limo/Modules/Localizer.cpp:245-377) over
// libflimo_cuda: the stages between the raw LiDAR message and the iterated update, with the reference's
// names.  The IMU-rate bookkeeping (propagated_buffer, integrateImu, time offset, :797-805) stays in the
// reference's Localizer; this class is what its scan thread calls instead of the PCL / OpenMP code.
//
//   Localizer::updatePointCloud                      here
//   :262-302  NaN / crop / dist / rate / FoV         filter_and_sort(raw_pc, time_stamp)   -> n kept, time of the last point
//   :744-789  sort by point time                       (same call)
//   :805      frames = integrateImu(...)             host (reference code), passed to deskew()
//   :822-843  per-point deskew                       deskew(frames, last_state pose, lidar2baselink_T, offset) -> pc2match bound
//   :313-321  voxel grid                               (same call, if voxel_active)
//   :333      iterated update                        Mapper::update
//   :361-377  world cloud + Mapper::add              Mapper::scan_to_world / Mapper::add
//
// Types are templates so that the header compiles without PCL / Eigen: PointCloudPtr is
// pcl::PointCloud<fast_limo::Point>::Ptr (32-byte points), FiltersT is fast_limo::Config::Filters.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

#include "../flimo.h"

namespace fast_limo_gpu {

class Localizer {
 public:
  explicit Localizer(flimo_handle h) : h_(h) {}

  // Localizer::init (:57-61) + set_sensor_type: FiltersT has the reference's field names.
  template <typename FiltersT>
  void set_filters(const FiltersT& f, int sensor_type, bool end_of_sweep) {
    cfg_ = flimo_prep_cfg{};
    cfg_.crop_active = f.crop_active;
    cfg_.dist_active = f.dist_active;
    cfg_.rate_active = f.rate_active;
    cfg_.fov_active = f.fov_active;
    cfg_.voxel_active = f.voxel_active;
    for (int i = 0; i < 3; ++i) {
      cfg_.cropBoxMin[i] = f.cropBoxMin[i];
      cfg_.cropBoxMax[i] = f.cropBoxMax[i];
    }
    cfg_.min_dist = f.min_dist;
    cfg_.rate_value = f.rate_value;
    cfg_.fov_angle = f.fov_angle;
    cfg_.leafSize = f.leafSize[0];                 // the reference passes leafSize[0] for all three axes (:61)
    cfg_.sensor_type = sensor_type;
    cfg_.end_of_sweep = end_of_sweep ? 1 : 0;
  }

  // :262-302 + :744-789.  Returns the size of the filtered cloud; t_last = time of its last point (:797,:802).
  template <typename PointCloudPtr>
  std::size_t filter_and_sort(PointCloudPtr& raw_pc, double time_stamp, double& t_last) {
    static_assert(sizeof(raw_pc->points[0]) == 32, "fast_limo::Point is 32 bytes");
    std::size_t n = 0;
    check(flimo_prep_filter_sort(h_, raw_pc->points.data(), raw_pc->points.size(), time_stamp, &cfg_, &n, &t_last));
    return n;
  }

  // :822-843 (+ :313-321).  frames: fast_limo::State members as flimo_frame; returns pc2match's size.
  std::size_t deskew(const std::vector<flimo_frame>& frames, const float last_q_xyzw[4], const float last_p[3],
                     const float lidar2baselink_T_rowmajor[16], double offset) {
    std::size_t n = 0;
    check(flimo_prep_deskew(h_, frames.data(), static_cast<int>(frames.size()), last_q_xyzw, last_p, lidar2baselink_T_rowmajor,
                            offset, &n));
    return n;
  }

  // get_deskewed_pointcloud / get_pc2match_pointcloud (Localizer.cpp:119-137): xyz1 float4 per point.
  std::vector<float> get_cloud(int what /*1 world, 2 deskewed_Xt2, 3 pc2match*/) {
    std::size_t n = 0;
    check(flimo_prep_get(h_, what, nullptr, 0, &n));
    std::vector<float> out(4 * n);
    if (n) check(flimo_prep_get(h_, what, out.data(), n, &n));
    return out;
  }

 private:
  void check(int rc) {
    if (rc != FLIMO_OK) throw std::runtime_error(std::string("libflimo_cuda: ") + flimo_last_error(h_));
  }
  flimo_handle h_;
  flimo_prep_cfg cfg_{};
};

}  // namespace fast_limo_gpu
