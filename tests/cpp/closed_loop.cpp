// Closed loop through fast_limo::Localizer / fast_limo::Mapper (include/fast_limo/**, libfast_limo.so) exactly the way the
// ROS wrapper drives them (src/main.cpp:16-93): Localizer::getInstance().init(config), then updateIMU for every IMU message
// and updatePointCloud for every LiDAR message of a recorded stream.
//
// usage: closed_loop <stream.bin> <poses.bin> [device]
// stream.bin (written by tests/test_gpu_cpp_closed_loop.py from the synthetic stream):
//   header : u64 n_scans, f64 leaf, f64 min_dist, i32 max_iters, i32 sensor_type, f64 p0[3], f64 q0[4] (xyzw), f64 v0[3]
//   scan   : f64 stamp, u64 n_imu, n_imu x {f64 stamp, f32 acc[3], f32 gyro[3]}, u64 n_points, n_points x fast_limo::Point (32 B)
// poses.bin: per scan f64 state26[26], f64 P_diag[23], f64 map_size, f64 passes
#include <cstdio>
#include <cstring>
#include <vector>

#include "fast_limo/Modules/Localizer.hpp"
#include "fast_limo/Modules/Mapper.hpp"

template <typename T>
static bool rd(std::FILE* f, T* v, size_t n = 1) { return std::fread(v, sizeof(T), n, f) == n; }

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::FILE* f = std::fopen(argv[1], "rb");
  std::FILE* o = std::fopen(argv[2], "wb");
  if (!f || !o) return 3;
  unsigned long long n_scans = 0;
  double leaf = 0, min_dist = 0;
  int max_iters = 3, sensor_type = 1;
  double p0[3], q0[4], v0[3];
  if (!rd(f, &n_scans) || !rd(f, &leaf) || !rd(f, &min_dist) || !rd(f, &max_iters) || !rd(f, &sensor_type) || !rd(f, p0, 3) || !rd(f, q0, 4) ||
      !rd(f, v0, 3))
    return 4;

  fast_limo::Localizer& loc = fast_limo::Localizer::getInstance();
  fast_limo::Mapper& map = fast_limo::Mapper::getInstance();
  fast_limo::Config config;                                   // what load_config (src/main.cpp:101-168) fills from the YAML
  config.sensor_type = sensor_type;
  config.filters.cropBoxMin = {-1.f, -1.f, -1.f};
  config.filters.cropBoxMax = {1.f, 1.f, 1.f};
  config.filters.crop_active = true;
  config.filters.dist_active = true;
  config.filters.min_dist = min_dist;
  config.filters.voxel_active = leaf > 0;
  config.filters.leafSize = {(float)leaf, (float)leaf, (float)leaf};
  config.extrinsics.imu2baselink_t = {0.f, 0.f, 0.f};
  config.extrinsics.imu2baselink_R = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
  config.extrinsics.lidar2baselink_t = {0.f, 0.f, 0.f};
  config.extrinsics.lidar2baselink_R = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
  config.intrinsics.accel_bias = {0.f, 0.f, 0.f};
  config.intrinsics.gyro_bias = {0.f, 0.f, 0.f};
  config.intrinsics.imu_sm = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
  config.ikfom.MAX_NUM_ITERS = max_iters;
  config.ikfom.LIMITS = std::vector<double>(23, 0.001);
  config.ikfom.estimate_extrinsics = true;
  config.ikfom.mapping.MAX_NUM_MATCHES = 1 << 18;
  config.ikfom.mapping.MAX_NUM_PC2MATCH = 1 << 18;
  config.gpu_device = argc > 3 ? std::atoi(argv[3]) : 0;
  loc.init(config);
  if (!loc.is_calibrated()) return 5;
  loc.set_initial_state(p0, q0, v0);

  for (unsigned long long k = 0; k < n_scans; ++k) {
    double stamp = 0;
    unsigned long long n_imu = 0, n_pts = 0;
    if (!rd(f, &stamp) || !rd(f, &n_imu)) return 6;
    for (unsigned long long i = 0; i < n_imu; ++i) {
      double t;
      float acc[3], gyro[3];
      if (!rd(f, &t) || !rd(f, acc, 3) || !rd(f, gyro, 3)) return 7;
      fast_limo::IMUmeas imu;                                   // tf_limo::fromROStoLimo (ROSutils.hpp)
      imu.stamp = t;
      imu.lin_accel = Eigen::Vector3f(acc[0], acc[1], acc[2]);
      imu.ang_vel = Eigen::Vector3f(gyro[0], gyro[1], gyro[2]);
      loc.updateIMU(imu);
    }
    if (!rd(f, &n_pts)) return 8;
    pcl::PointCloud<PointType>::Ptr pc_(fast_limo::make_shared<pcl::PointCloud<PointType>>());
    pc_->points.resize(n_pts);
    if (n_pts && !rd(f, pc_->points.data(), n_pts)) return 9;
    loc.updatePointCloud(pc_, stamp);
    // what the wrapper publishes after a scan
    const fast_limo::State ws = loc.getWorldState();
    const std::vector<double> pose_cov = loc.getPoseCovariance();
    (void)ws;
    (void)pose_cov;
    (void)loc.get_pointcloud();
    double rec[26 + 23 + 2];
    std::memcpy(rec, loc.state26(), 26 * sizeof(double));
    for (int i = 0; i < 23; ++i) rec[26 + i] = loc.covariance529()[i * 23 + i];
    rec[49] = (double)map.size();
    rec[50] = (double)loc.last_passes();
    std::fwrite(rec, sizeof(double), 51, o);
  }
  std::fclose(f);
  std::fclose(o);
  std::printf("closed loop ok: %llu scans, map %d points\n", n_scans, map.size());
  return 0;
}
