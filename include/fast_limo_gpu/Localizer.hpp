// Device side of fast_limo::Localizer::updatePointCloud (fast_limo/Modules/Localizer.cpp:245-377) over
// libflimo_cuda: the stages between the raw LiDAR message and the iterated update, with the reference's
// names, plus the IMU-rate host algebra (prediction, propagated_buffer, integrateImu).  IMU calibration, the
// IMU -> base-link transform and the time offset (:401-531, :697-728, :797-801) stay in the reference's
// Localizer; this class is what its two callbacks call instead of the Eigen / PCL / OpenMP code.
//
//   Localizer::updatePointCloud                      here
//   :262-302  NaN / crop / dist / rate / FoV         filter_and_sort(raw_pc, time_stamp)   -> n kept, time of the last point
//   :744-789  sort by point time                       (same call)
//   :583-608  propagateImu(imu) (IMU callback)       propagateImu(x, P, imu, cov)  -> predict + push on the ring
//   :805      frames = integrateImu(...)             integrateImu(prev_scan_stamp, scan_stamp), passed to deskew()
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

  // Localizer::propagateImu(const IMUmeas&) (:583-608): x26 / P529 are _iKFoM's x_ and P_ (flat, as in flimo.h),
  // IKFoMT is fast_limo::Config::iKFoM (cov_gyro, cov_acc, cov_bias_gyro, cov_bias_acc).  ImuT is
  // fast_limo::IMUmeas (stamp, dt, ang_vel, lin_accel with operator[]).
  template <typename ImuT, typename IKFoMT>
  void propagateImu(double x26[26], double P529[529], const ImuT& imu, const IKFoMT& ikfom) {
    flimo_imu m{};
    m.stamp = imu.stamp;
    m.dt = imu.dt;
    for (int i = 0; i < 3; ++i) {
      m.ang_vel[i] = imu.ang_vel[i];
      m.lin_accel[i] = imu.lin_accel[i];
    }
    const double cov4[4] = {ikfom.cov_gyro, ikfom.cov_acc, ikfom.cov_bias_gyro, ikfom.cov_bias_acc};
    check(flimo_ekf_predict(h_, x26, P529, &m, cov4));
  }

  // Localizer::integrateImu (:855-871).  Empty = "not enough propagated states" (the first scan); throws when
  // the IMU side has not reached end_time yet (the reference waits on cv_prop_stamp there, :880-887).
  std::vector<flimo_frame> integrateImu(double start_time, double end_time) {
    std::size_t n = 0;
    check(flimo_propagated_frames(h_, start_time, end_time, nullptr, 0, &n));
    std::vector<flimo_frame> frames(n);
    if (n) check(flimo_propagated_frames(h_, start_time, end_time, frames.data(), n, &n));
    return frames;
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
