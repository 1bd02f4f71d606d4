// fast_limo::Localizer with the reference's call surface (fast_limo/Modules/Localizer.hpp:139-180, singleton :206-209)
// over libflimo_cuda: the ROS wrapper's call sites (src/main.cpp:16-93,178-206) compile against this header unchanged.
//
//   updateIMU        : IMU -> base_link transform, stand-still calibration, bias correction, esekf::predict and the ring of
//                      propagated states (host algebra inside the library: flimo_ekf_predict)
//   updatePointCloud : filters / time sort / deskew / voxel grid on the device (flimo_prep_*), the iterated update entirely on
//                      the device (flimo_update), world cloud + Mapper::add on the device (flimo_map_add_scan)
// What the reference keeps in an esekfom::esekf object (_iKFoM: x_, P_) is kept here as the flat state26 / P529 of flimo.h.
#pragma once
#include <mutex>

#include "fast_limo/Common.hpp"
#include "fast_limo/Modules/Mapper.hpp"
#include "fast_limo/Objects/Match.hpp"
#include "fast_limo/Objects/State.hpp"
#include "fast_limo/Utils/Config.hpp"

class fast_limo::Localizer {
 public:
  pcl::PointCloud<PointType>::Ptr pc2match;   // pointcloud to match in Xt2 (last_state) frame

  Localizer();
  void init(Config& cfg);

  // Callbacks
  void updateIMU(IMUmeas& raw_imu);
  void updatePointCloud(pcl::PointCloud<PointType>::Ptr& raw_pc, double time_stamp);

  // Get output
  pcl::PointCloud<PointType>::Ptr get_pointcloud();
  pcl::PointCloud<PointType>::Ptr get_finalraw_pointcloud();
  pcl::PointCloud<PointType>::ConstPtr get_orig_pointcloud();
  pcl::PointCloud<PointType>::ConstPtr get_deskewed_pointcloud();
  pcl::PointCloud<PointType>::Ptr get_pc2match_pointcloud();
  Matches& get_matches();

  State getWorldState();   // state in body/base_link frame
  State getBodyState();    // state in LiDAR frame
  std::vector<double> getPoseCovariance();
  std::vector<double> getTwistCovariance();
  double get_propagate_time();

  void get_cpu_stats(float& comput_time, float& max_comput_time, float& mean_comput_time, float& cpu_cores, float& cpu_load,
                     float& cpu_max_load, float& ram_usage);
  bool is_calibrated();
  void set_sensor_type(uint8_t type);
  fast_limo::SensorType get_sensor_type();

  // iKFoM measurement model, reference form (Localizer.cpp:537-577): rows from materialised matches.  The device path never
  // builds H; this is the call the reference's use-ikfom.cpp makes and the unit tests compare with the kernel's sums.
  void calculate_H(const state_ikfom&, const Matches&, Eigen::MatrixXd& H, Eigen::VectorXd& h);

  // Backpropagation
  void propagateImu(const IMUmeas& imu);

  // the filter state as flimo.h lays it out (extension: lets harnesses compare poses exactly)
  const double* state26() const { return x_; }
  const double* covariance529() const { return P_; }
  int last_passes() const { return last_passes_; }
  // extension for harnesses: start from a known pose / velocity instead of the origin at rest (call after init)
  void set_initial_state(const double p[3], const double q_xyzw[4], const double v[3]);

  static Localizer& getInstance() {
    static Localizer* loc = new Localizer();
    return *loc;
  }

 private:
  Localizer(const Localizer&) = delete;
  Localizer& operator=(const Localizer&) = delete;

  void init_iKFoM_state();
  IMUmeas imu2baselink(IMUmeas& imu);
  state_ikfom get_x() const;
  void check(int rc) const;

  std::mutex mtx_ikfom;
  double x_[26], P_[529];
  State state;
  Extrinsics extr;
  SensorType sensor = SensorType::UNKNOWN;
  IMUmeas last_imu;
  Config config;
  Matches matches;
  flimo_prep_cfg prep_{};

  pcl::PointCloud<PointType>::ConstPtr original_scan, deskewed_scan;
  pcl::PointCloud<PointType>::Ptr final_raw_scan, final_scan;

  double scan_stamp = 0.0, prev_scan_stamp = 0.0;
  double imu_stamp = 0.0, prev_imu_stamp = 0.0, first_imu_stamp = 0.0, last_propagate_time_ = 0.0, imu_calib_time_ = 3.0;
  double gravity_ = 9.81;
  bool imu_calibrated_ = false;
  bool have_imu_ = false;
  int last_passes_ = 0;
  Eigen::Matrix3f imu_accel_sm_;
  // calibration accumulators (function-local statics in the reference, Localizer.cpp:414-417,707)
  int calib_samples_ = 0;
  Eigen::Vector3f gyro_avg_, accel_avg_, ang_vel_cg_prev_;
  bool have_prev_ang_vel_ = false;
  // stats
  float cpu_time_ = 0.f, cpu_max_time_ = 0.f, cpu_mean_time_ = 0.f;
  unsigned long n_scans_ = 0;
};
