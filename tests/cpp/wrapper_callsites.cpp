// Type-check of the ROS wrapper's fast_limo call sites (src/main.cpp:16-93,178-206, ROSutils.hpp member accesses)
// against include/fast_limo/** — compiled (not run) by tests/test_cpp_facade.py with plain g++, no ROS / PCL / Eigen.
#include "fast_limo/Modules/Localizer.hpp"
#include "fast_limo/Modules/Mapper.hpp"

double lidar_callback_body(pcl::PointCloud<PointType>::Ptr& pc_, double stamp) {          // src/main.cpp:16-65
  fast_limo::Localizer& loc = fast_limo::Localizer::getInstance();
  fast_limo::SensorType st = loc.get_sensor_type();
  (void)st;
  loc.updatePointCloud(pc_, stamp);
  pcl::PointCloud<PointType>::Ptr a = loc.get_pointcloud();
  pcl::PointCloud<PointType>::ConstPtr b = loc.get_orig_pointcloud();
  pcl::PointCloud<PointType>::ConstPtr c = loc.get_deskewed_pointcloud();
  pcl::PointCloud<PointType>::Ptr d = loc.get_pc2match_pointcloud();
  pcl::PointCloud<PointType>::Ptr e = loc.get_finalraw_pointcloud();
  Matches& m = loc.get_matches();
  double s = (double)(a->points.size() + b->points.size() + c->points.size() + d->points.size() + e->points.size());
  for (fast_limo::Match& match : m) {                                                       // ROSutils.hpp:216-252 (markers)
    const Eigen::Vector3f g = match.get_global_point();
    const Eigen::Vector4f n = match.plane.get_normal();
    s += g(0) + n(3) + match.dist + (match.lisanAlGaib() ? 1.0 : 0.0);
  }
  return s;
}

double imu_callback_body(fast_limo::IMUmeas& imu) {                                       // src/main.cpp:67-95
  fast_limo::Localizer& loc = fast_limo::Localizer::getInstance();
  loc.updateIMU(imu);
  fast_limo::State w = loc.getWorldState(), b = loc.getBodyState();
  std::vector<double> pc = loc.getPoseCovariance(), tc = loc.getTwistCovariance();
  // tf_limo::fromLimoToROS (ROSutils.hpp:56-110): the members it reads
  return w.p(0) + w.p(1) + w.p(2) + w.q.x() + w.q.y() + w.q.z() + w.q.w() + w.v(0) + w.w(2) + w.a(1) + w.b.gyro(0) + w.b.accel(2) +
         w.g(2) + w.pLI(0) + w.qLI.w() + w.time + b.p(0) + pc[0] + tc[35];
}

int main_body() {                                                                         // src/main.cpp:178-206
  fast_limo::Localizer& loc = fast_limo::Localizer::getInstance();
  fast_limo::Mapper& map = fast_limo::Mapper::getInstance();
  fast_limo::Config config;
  config.topics.lidar = "/velodyne_points";
  config.ikfom.mapping.octree.bucket_size = 2;
  loc.init(config);
  float a, b, c, d, e, f, g;
  loc.get_cpu_stats(a, b, c, d, e, f, g);
  loc.set_sensor_type(1);
  return map.size() + (map.exists() ? 1 : 0) + (int)map.last_time() + (loc.is_calibrated() ? 1 : 0) + (int)loc.get_propagate_time();
}

void hook_body(state_ikfom& s, pcl::PointCloud<PointType>::Ptr& pc, Eigen::MatrixXd& H, Eigen::VectorXd& h) {   // use-ikfom.cpp:10-31
  fast_limo::Mapper& MAP = fast_limo::Mapper::getInstance();
  fast_limo::Localizer& LOC = fast_limo::Localizer::getInstance();
  Matches matches = MAP.match(fast_limo::State(s), pc);
  LOC.calculate_H(s, matches, H, h);
  MAP.matches.clear();
}
