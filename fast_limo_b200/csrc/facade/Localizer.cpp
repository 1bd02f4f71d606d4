// fast_limo::Localizer over libflimo_cuda: the sequencing of the reference's two callbacks
// (fast_limo/Modules/Localizer.cpp:245-399 updatePointCloud, :401-531 updateIMU) with every data-parallel stage on the B200.
#include "fast_limo/Modules/Localizer.hpp"

#include <chrono>
#include <cstring>
#include <iostream>
#include <stdexcept>

namespace fast_limo {

namespace {
Eigen::Matrix3f map3(const std::vector<float>& v) {                 // Eigen::Map<Matrix3f>(data, 3, 3): column-major
  Eigen::Matrix3f M = Eigen::Matrix3f::Identity();
  if (v.size() >= 9)
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) M(r, c) = v[3 * c + r];
  return M;
}
Eigen::Vector3f map3v(const std::vector<float>& v) { return v.size() >= 3 ? Eigen::Vector3f(v[0], v[1], v[2]) : Eigen::Vector3f(); }
pcl::PointCloud<PointType>::Ptr cloud_from_xyz4(const std::vector<float>& xyz4) {
  auto pc = fast_limo::make_shared<pcl::PointCloud<PointType>>();
  pc->points.resize(xyz4.size() / 4);
  for (size_t i = 0; i < pc->points.size(); ++i) pc->points[i] = PointType(xyz4[4 * i], xyz4[4 * i + 1], xyz4[4 * i + 2]);
  pc->width = (std::uint32_t)pc->points.size();
  return pc;
}
}  // namespace

Localizer::Localizer() {
  std::memset(x_, 0, sizeof(x_));
  std::memset(P_, 0, sizeof(P_));
  x_[6] = x_[10] = 1.0;
  pc2match = fast_limo::make_shared<pcl::PointCloud<PointType>>();
  final_raw_scan = fast_limo::make_shared<pcl::PointCloud<PointType>>();
  final_scan = fast_limo::make_shared<pcl::PointCloud<PointType>>();
  original_scan = fast_limo::make_shared<pcl::PointCloud<PointType>>();
  deskewed_scan = fast_limo::make_shared<pcl::PointCloud<PointType>>();
  imu_accel_sm_ = Eigen::Matrix3f::Identity();
}

void Localizer::check(int rc) const {
  if (rc != FLIMO_OK) throw std::runtime_error(std::string("libflimo_cuda: ") + flimo_last_error(Mapper::getInstance().gpu()));
}

void Localizer::init(Config& cfg) {                                 // Localizer.cpp:35-117
  config = cfg;
  Mapper& map = Mapper::getInstance();
  map.set_num_threads(config.num_threads);
  map.set_config(config.ikfom.mapping, config.ikfom.estimate_extrinsics, config.gpu_device);
  if (config.ikfom.LIMITS.size() < 23) config.ikfom.LIMITS.resize(23, config.ikfom.LIMITS.empty() ? 0.001 : config.ikfom.LIMITS[0]);

  // filters (Localizer.cpp:57-61: the crop box is negative, the leaf is leafSize[0] on all axes)
  prep_ = flimo_prep_cfg{};
  prep_.crop_active = config.filters.crop_active && config.filters.cropBoxMin.size() >= 3 && config.filters.cropBoxMax.size() >= 3;
  for (int i = 0; i < 3 && prep_.crop_active; ++i) {
    prep_.cropBoxMin[i] = config.filters.cropBoxMin[i];
    prep_.cropBoxMax[i] = config.filters.cropBoxMax[i];
  }
  prep_.dist_active = config.filters.dist_active;
  prep_.min_dist = config.filters.min_dist;
  prep_.rate_active = config.filters.rate_active;
  prep_.rate_value = config.filters.rate_value > 0 ? config.filters.rate_value : 1;
  prep_.fov_active = config.filters.fov_active;
  prep_.fov_angle = config.filters.fov_angle;
  prep_.voxel_active = config.filters.voxel_active && !config.filters.leafSize.empty();
  prep_.leafSize = config.filters.leafSize.empty() ? 0.f : config.filters.leafSize[0];
  prep_.end_of_sweep = config.end_of_sweep ? 1 : 0;
  set_sensor_type((uint8_t)config.sensor_type);
  prep_.sensor_type = config.sensor_type;

  // IMU intrinsics, extrinsics (stored transposed at init: Localizer.cpp:72-86)
  imu_accel_sm_ = map3(config.intrinsics.imu_sm);
  state.b.accel = map3v(config.intrinsics.accel_bias);
  state.b.gyro = map3v(config.intrinsics.gyro_bias);
  extr.imu2baselink.t = map3v(config.extrinsics.imu2baselink_t);
  extr.imu2baselink.R = map3(config.extrinsics.imu2baselink_R).transpose();
  extr.lidar2baselink.t = map3v(config.extrinsics.lidar2baselink_t);
  extr.lidar2baselink.R = map3(config.extrinsics.lidar2baselink_R).transpose();
  for (Extrinsics::SE3* se : {&extr.imu2baselink, &extr.lidar2baselink}) {
    Eigen::Matrix4f& T = se == &extr.imu2baselink ? extr.imu2baselink_T : extr.lidar2baselink_T;
    T = Eigen::Matrix4f::Identity();
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) T(r, c) = se->R(r, c);
      T(r, 3) = se->t(r);
    }
  }
  check(flimo_propagated_clear(map.gpu()));
  if (!(config.gravity_align || config.calibrate_accel || config.calibrate_gyro)) {   // no automatic calibration
    imu_calibrated_ = true;
    init_iKFoM_state();
  }
  imu_calib_time_ = config.imu_calib_time;
}

void Localizer::init_iKFoM_state() {                                // Localizer.cpp:672-694
  const Eigen::Quaternionf qLI(extr.lidar2baselink.R);
  // S2<double, 98090, 10000, 1> rescales (0, 0, -gravity_) to the length 9.809 (use-ikfom.hpp:8, S2.hpp:123-126)
  const double g = std::fabs(gravity_) * (9.809 / std::fabs(gravity_));
  const double x[26] = {state.p(0), state.p(1), state.p(2), state.q.x(), state.q.y(), state.q.z(), state.q.w(),
                        qLI.x(), qLI.y(), qLI.z(), qLI.w(), extr.lidar2baselink.t(0), extr.lidar2baselink.t(1), extr.lidar2baselink.t(2),
                        0.0, 0.0, 0.0, state.b.gyro(0), state.b.gyro(1), state.b.gyro(2), state.b.accel(0), state.b.accel(1), state.b.accel(2),
                        0.0, 0.0, -g};
  std::memcpy(x_, x, sizeof(x_));
  std::memset(P_, 0, sizeof(P_));
  for (int i = 0; i < 23; ++i) P_[i * 23 + i] = 1.0;
  for (int i = 6; i < 12; ++i) P_[i * 23 + i] = 0.000001;
  for (int i = 15; i < 18; ++i) P_[i * 23 + i] = 0.00001;
  for (int i = 18; i < 21; ++i) P_[i * 23 + i] = 0.0001;
  for (int i = 21; i < 23; ++i) P_[i * 23 + i] = 0.000001;
}

void Localizer::set_initial_state(const double p[3], const double q_xyzw[4], const double v[3]) {
  state.p = Eigen::Vector3f((float)p[0], (float)p[1], (float)p[2]);
  state.q = Eigen::Quaternionf((float)q_xyzw[3], (float)q_xyzw[0], (float)q_xyzw[1], (float)q_xyzw[2]);
  init_iKFoM_state();
  for (int i = 0; i < 3; ++i) {                       // the filter state itself keeps the full precision of the arguments
    x_[i] = p[i];
    x_[14 + i] = v[i];
  }
  for (int i = 0; i < 4; ++i) x_[3 + i] = q_xyzw[i];
}

state_ikfom Localizer::get_x() const {
  state_ikfom s;
  s.pos = Eigen::Vector3d(x_[0], x_[1], x_[2]);
  s.rot = Eigen::Quaterniond(x_[6], x_[3], x_[4], x_[5]);
  s.offset_R_L_I = Eigen::Quaterniond(x_[10], x_[7], x_[8], x_[9]);
  s.offset_T_L_I = Eigen::Vector3d(x_[11], x_[12], x_[13]);
  s.vel = Eigen::Vector3d(x_[14], x_[15], x_[16]);
  s.bg = Eigen::Vector3d(x_[17], x_[18], x_[19]);
  s.ba = Eigen::Vector3d(x_[20], x_[21], x_[22]);
  s.grav = Eigen::Vector3d(x_[23], x_[24], x_[25]);
  return s;
}

// ---- getters (Localizer.cpp:119-244) -----------------------------------------------------------------------------------
pcl::PointCloud<PointType>::Ptr Localizer::get_pointcloud() { return final_scan; }
pcl::PointCloud<PointType>::Ptr Localizer::get_finalraw_pointcloud() { return final_raw_scan; }
pcl::PointCloud<PointType>::ConstPtr Localizer::get_orig_pointcloud() { return original_scan; }
pcl::PointCloud<PointType>::ConstPtr Localizer::get_deskewed_pointcloud() { return deskewed_scan; }
pcl::PointCloud<PointType>::Ptr Localizer::get_pc2match_pointcloud() { return pc2match; }
Matches& Localizer::get_matches() { return matches; }
bool Localizer::is_calibrated() { return imu_calibrated_; }
void Localizer::set_sensor_type(uint8_t type) { sensor = type < 5 ? static_cast<SensorType>(type) : SensorType::UNKNOWN; }
SensorType Localizer::get_sensor_type() { return sensor; }
double Localizer::get_propagate_time() { return last_propagate_time_; }

State Localizer::getBodyState() {
  if (!is_calibrated()) return State();
  State out(get_x());
  out.w = last_imu.ang_vel;
  out.a = last_imu.lin_accel;
  out.time = imu_stamp;
  out.p += out.pLI;                                               // position in LiDAR frame
  out.q *= out.qLI;                                               // attitude in LiDAR frame
  out.v = out.q.toRotationMatrix().transpose() * out.v;           // local velocity vector
  return out;
}

State Localizer::getWorldState() {
  if (!is_calibrated()) return State();
  State out(get_x());
  out.w = last_imu.ang_vel;
  out.a = last_imu.lin_accel;
  out.time = imu_stamp;
  out.v = out.q.toRotationMatrix().transpose() * out.v;
  return out;
}

void Localizer::get_cpu_stats(float& comput_time, float& max_comput_time, float& mean_comput_time, float& cpu_cores, float& cpu_load,
                              float& cpu_max_load, float& ram_usage) {
  comput_time = cpu_time_;
  max_comput_time = cpu_max_time_;
  mean_comput_time = cpu_mean_time_;
  cpu_cores = cpu_load = cpu_max_load = ram_usage = 0.f;          // the debug board's /proc statistics are not mirrored
}

std::vector<double> Localizer::getPoseCovariance() {              // 6x6 [orientation, position], column-major like Eigen::Map
  std::vector<double> cov(36, 0.0);
  if (!is_calibrated()) return cov;
  auto P = [this](int r, int c) { return P_[r * 23 + c]; };
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      cov[c * 6 + r] = P(3 + r, 3 + c);
      cov[(3 + c) * 6 + r] = P(3 + r, c);
      cov[c * 6 + 3 + r] = P(r, 3 + c);
      cov[(3 + c) * 6 + 3 + r] = P(r, c);
    }
  return cov;
}

std::vector<double> Localizer::getTwistCovariance() {             // as written in the reference: block (6,6) and cov_gyro
  std::vector<double> cov(36, 0.0);
  if (!is_calibrated()) return cov;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) cov[c * 6 + r] = P_[(6 + r) * 23 + 6 + c];
  for (int i = 3; i < 6; ++i) cov[i * 6 + i] = config.ikfom.cov_gyro;
  return cov;
}

// ---- IMU callback (Localizer.cpp:401-531) ---------------------------------------------------------------------------------
IMUmeas Localizer::imu2baselink(IMUmeas& imu) {                   // Localizer.cpp:696-731 (float arithmetic, dt fallback 1/200 s)
  IMUmeas out;
  double dt = imu.stamp - prev_imu_stamp;
  if (dt == 0. || dt > 0.1) dt = 1.0 / 200.0;
  const Eigen::Vector3f ang_vel_cg = extr.imu2baselink.R * imu.ang_vel;
  if (!have_prev_ang_vel_) {
    ang_vel_cg_prev_ = ang_vel_cg;
    have_prev_ang_vel_ = true;
  }
  Eigen::Vector3f lin_accel_cg = extr.imu2baselink.R * imu.lin_accel;
  const Eigen::Vector3f lever = -extr.imu2baselink.t;
  lin_accel_cg = lin_accel_cg + ((ang_vel_cg - ang_vel_cg_prev_) / (float)dt).cross(lever) + ang_vel_cg.cross(ang_vel_cg.cross(lever));
  ang_vel_cg_prev_ = ang_vel_cg;
  out.ang_vel = ang_vel_cg;
  out.lin_accel = lin_accel_cg;
  out.dt = dt;
  out.stamp = imu.stamp;
  Eigen::Quaternionf q(extr.imu2baselink.R);
  q.normalize();
  out.q = q * imu.q;
  prev_imu_stamp = imu.stamp;
  return out;
}

void Localizer::updateIMU(IMUmeas& raw_imu) {
  imu_stamp = raw_imu.stamp;
  IMUmeas imu = imu2baselink(raw_imu);
  if (first_imu_stamp == 0.0) first_imu_stamp = imu.stamp;

  if (!imu_calibrated_) {                                          // stand-still calibration (:412-509)
    if ((imu.stamp - first_imu_stamp) < imu_calib_time_) {
      ++calib_samples_;
      gyro_avg_ += imu.ang_vel;
      accel_avg_ += imu.lin_accel;
      return;
    }
    if (calib_samples_ > 0) {
      gyro_avg_ /= (float)calib_samples_;
      accel_avg_ /= (float)calib_samples_;
    }
    Eigen::Vector3f grav_vec(0.f, 0.f, (float)gravity_);
    state.q = imu.q;
    if (config.gravity_align) {
      grav_vec = (accel_avg_ - state.b.accel).normalized() * (float)std::fabs(gravity_);
      state.q = Eigen::Quaternionf::FromTwoVectors(grav_vec, Eigen::Vector3f(0.f, 0.f, (float)gravity_));
      state.g = grav_vec;
    }
    if (config.calibrate_accel) state.b.accel = accel_avg_ - grav_vec;
    if (config.calibrate_gyro) state.b.gyro = gyro_avg_;
    state.q.normalize();
    init_iKFoM_state();
    imu_calibrated_ = true;
    return;
  }

  // calibrated: intrinsic correction, remember the sample, propagate the filter (:512-528)
  imu.lin_accel = (imu_accel_sm_ * imu.lin_accel) - state.b.accel;
  imu.ang_vel = imu.ang_vel - state.b.gyro;
  last_imu = imu;
  have_imu_ = true;
  propagateImu(imu);
}

void Localizer::propagateImu(const IMUmeas& imu) {                // Localizer.cpp:583-608
  flimo_imu m{};
  m.stamp = imu.stamp;
  m.dt = imu.dt;
  for (int i = 0; i < 3; ++i) {
    m.ang_vel[i] = imu.ang_vel(i);
    m.lin_accel[i] = imu.lin_accel(i);
  }
  const double cov4[4] = {config.ikfom.cov_gyro, config.ikfom.cov_acc, config.ikfom.cov_bias_gyro, config.ikfom.cov_bias_acc};
  std::lock_guard<std::mutex> lock(mtx_ikfom);
  check(flimo_ekf_predict(Mapper::getInstance().gpu(), x_, P_, &m, cov4));   // esekf::predict + push on the propagated ring
  last_propagate_time_ = imu.stamp;
}

// ---- LiDAR callback (Localizer.cpp:245-399) ---------------------------------------------------------------------------------
void Localizer::updatePointCloud(pcl::PointCloud<PointType>::Ptr& raw_pc, double time_stamp) {
  const auto t_start = std::chrono::steady_clock::now();
  if (!raw_pc || raw_pc->points.size() < 1) {
    std::cout << "FAST_LIMO::Raw PointCloud is empty!\n";
    return;
  }
  if (!imu_calibrated_) return;
  if (!have_imu_) {
    std::cout << "FAST_LIMO::IMU buffer is empty!\n";
    return;
  }
  Mapper& map = Mapper::getInstance();
  flimo_handle h = map.gpu();

  // NaN / crop / distance / rate / FoV filters and the time sort (:262-302, :744-789), on the device
  size_t n_kept = 0;
  double t_last = 0.0;
  check(flimo_prep_filter_sort(h, raw_pc->points.data(), raw_pc->points.size(), time_stamp, &prep_, &n_kept, &t_last));
  if (n_kept < 1) return;
  if (config.debug) {                                               // original_scan: the filtered cloud, LiDAR frame
    std::vector<std::uint32_t> order(n_kept);
    size_t n = 0;
    check(flimo_prep_get(h, 0, order.data(), order.size(), &n));
    auto orig = fast_limo::make_shared<pcl::PointCloud<PointType>>();
    orig->points.reserve(n);
    for (size_t i = 0; i < n; ++i) orig->points.push_back(raw_pc->points[order[i]]);
    original_scan = orig;
  }
  double offset = 0.0;
  if (config.time_offset) {                                         // :797-801
    offset = imu_stamp - t_last - 1.e-4;
    if (offset > 0.0) offset = 0.0;
  }
  scan_stamp = t_last + offset;                                     // :805

  // frames of the sweep from the propagated states (:808, integrateImu :855-871)
  size_t n_frames = 0;
  int rc = flimo_propagated_frames(h, prev_scan_stamp, scan_stamp, nullptr, 0, &n_frames);
  if (rc != FLIMO_OK) {                                             // IMU behind the scan: the reference waits on cv_prop_stamp (:880-887)
    std::cout << "FAST_LIMO::propagated states do not reach the end of the scan yet\n";
    prev_scan_stamp = scan_stamp;
    return;
  }
  std::vector<flimo_frame> frames(n_frames);
  if (n_frames) check(flimo_propagated_frames(h, prev_scan_stamp, scan_stamp, frames.data(), n_frames, &n_frames));
  size_t n_pc2match = 0;
  if (n_frames < 1) {                                               // first scan (prev_scan_stamp = 0): :807-814
    std::cout << "FAST_LIMO::deskewPointCloud(): no frames obtained from IMU propagation!\n";
  } else {
    const float lq[4] = {(float)x_[3], (float)x_[4], (float)x_[5], (float)x_[6]}, lp[3] = {(float)x_[0], (float)x_[1], (float)x_[2]};
    float T[16];
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) T[4 * r + c] = extr.lidar2baselink_T(r, c);
    check(flimo_prep_deskew(h, frames.data(), (int)n_frames, lq, lp, T, offset, &n_pc2match));   // :822-843 (+ voxel grid :313-321)
  }

  if (n_pc2match > 1) {
    int passes = 0;
    {
      std::lock_guard<std::mutex> lock(mtx_ikfom);                  // :326-353
      check(flimo_update(h, x_, P_, config.ikfom.MAX_NUM_ITERS, config.ikfom.LIMITS.data(), 0.001 /*LiDAR noise*/, 5.0 /*degeneracy*/, &passes));
      map.matches.clear();
      State corrected(get_x());
      if (config.calibrate_gyro) corrected.b.gyro = state.b.gyro;    // calibrated biases / gravity stay constant
      if (config.calibrate_accel) corrected.b.accel = state.b.accel;
      if (config.gravity_align) corrected.g = state.g;
      state = corrected;
      state.w = last_imu.ang_vel;
      state.a = last_imu.lin_accel;
    }
    last_passes_ = passes;
    extr.lidar2baselink_T = state.get_extr_RT();                    // :356
    // world cloud + Mapper::add without leaving the device (:361, :377)
    check(flimo_map_add_scan(h, x_, scan_stamp));
    if (config.debug) {                                             // clouds for the wrapper's debug topics
      std::vector<float> xyz4;
      size_t n = 0;
      check(flimo_prep_get(h, 3, nullptr, 0, &n));
      xyz4.resize(4 * n);
      if (n) check(flimo_prep_get(h, 3, xyz4.data(), n, &n));
      pc2match = cloud_from_xyz4(xyz4);
      check(flimo_prep_get(h, 1, nullptr, 0, &n));
      xyz4.resize(4 * n);
      if (n) check(flimo_prep_get(h, 1, xyz4.data(), n, &n));
      deskewed_scan = cloud_from_xyz4(xyz4);
      check(flimo_prep_get(h, 2, nullptr, 0, &n));                  // final_raw_scan: deskewed cloud without the voxel grid, world frame
      xyz4.resize(4 * n);
      if (n) check(flimo_prep_get(h, 2, xyz4.data(), n, &n));
      const Eigen::Matrix4f RT = state.get_RT();
      auto fr = cloud_from_xyz4(xyz4);
      for (PointType& p : fr->points) {
        const float x = p.x, y = p.y, z = p.z;
        p.x = RT(0, 0) * x + RT(0, 1) * y + RT(0, 2) * z + RT(0, 3);
        p.y = RT(1, 0) * x + RT(1, 1) * y + RT(1, 2) * z + RT(1, 3);
        p.z = RT(2, 0) * x + RT(2, 1) * y + RT(2, 2) * z + RT(2, 3);
      }
      final_raw_scan = fr;
    }
    {                                                               // final_scan = pc2match in the world frame (:371)
      size_t n = 0;
      check(flimo_scan_to_world(h, x_, nullptr, 0, &n));
      std::vector<float> xyz(3 * (n ? n : 1));
      check(flimo_scan_to_world(h, x_, xyz.data(), n, &n));
      auto fs = fast_limo::make_shared<pcl::PointCloud<PointType>>();
      fs->points.resize(n);
      for (size_t i = 0; i < n; ++i) fs->points[i] = PointType(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
      final_scan = fs;
    }
  } else {
    std::cout << "-------------- FAST_LIMO::NULL ITERATION --------------\n";
  }
  const float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_start).count();
  cpu_time_ = ms;
  cpu_max_time_ = ms > cpu_max_time_ ? ms : cpu_max_time_;
  cpu_mean_time_ = (cpu_mean_time_ * (float)n_scans_ + ms) / (float)(n_scans_ + 1);
  ++n_scans_;
  prev_scan_stamp = scan_stamp;                                     // :398
}

// ---- measurement model, reference form (Localizer.cpp:537-577) -------------------------------------------------------------
void Localizer::calculate_H(const state_ikfom& s, const Matches& m, Eigen::MatrixXd& H, Eigen::VectorXd& h) {
  const int N = (int)m.size() > config.ikfom.mapping.MAX_NUM_MATCHES ? config.ikfom.mapping.MAX_NUM_MATCHES : (int)m.size();
  H = Eigen::MatrixXd::Zero(N, 12);
  h.resize(N);
  const State S(s);
  const Eigen::Matrix3f R = S.q.toRotationMatrix(), RLI = S.qLI.toRotationMatrix();
  const Eigen::Matrix3f Rinv = R.transpose(), RLIinv = RLI.transpose();
  const Eigen::Matrix3f Rd_inv = s.rot.conjugate().toRotationMatrix().cast<float>();          // :554-555: double conjugate -> float
  const Eigen::Matrix3f RdLI_inv = s.offset_R_L_I.conjugate().toRotationMatrix().cast<float>();
  for (int i = 0; i < N; ++i) {
    const Match& match = m[i];
    const Eigen::Vector3f pg = match.get_global_point();
    const Eigen::Vector3f p_imu = Rinv * (pg - S.p);                // T_wb^-1 * p_global
    const Eigen::Vector3f p_lidar = RLIinv * (p_imu - S.pLI);       // T_LI^-1 * p_imu
    const Eigen::Vector4f n4 = match.plane.get_normal();
    const Eigen::Vector3f n(n4(0), n4(1), n4(2));
    const Eigen::Vector3f C = Rd_inv * n;
    const Eigen::Vector3f A = p_imu.cross(C);
    const Eigen::Vector3f B = p_lidar.cross(RdLI_inv * C);
    for (int k = 0; k < 3; ++k) {
      H(i, k) = n(k);
      H(i, 3 + k) = A(k);
      if (config.ikfom.estimate_extrinsics) {
        H(i, 6 + k) = B(k);
        H(i, 9 + k) = C(k);
      }
    }
    h(i) = -match.dist;
  }
  matches = m;                                                      // :575-576 (debug copy for the markers)
}

}  // namespace fast_limo
