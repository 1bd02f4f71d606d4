// fast_limo::Config with the reference's field names (fast_limo/Utils/Config.hpp:23-95): the ROS wrapper's load_config
// (src/main.cpp:101-168) fills it from the YAML files unchanged.
#pragma once
#include "fast_limo/Common.hpp"

struct fast_limo::Config {
  struct Topics { std::string lidar, imu; } topics;
  struct Extrinsics { std::vector<float> imu2baselink_t, imu2baselink_R, lidar2baselink_t, lidar2baselink_R; } extrinsics;
  struct Intrinsics { std::vector<float> accel_bias, gyro_bias, imu_sm; } intrinsics;
  struct Filters {
    std::vector<float> cropBoxMin, cropBoxMax; bool crop_active = false;      // crop box (negative)
    std::vector<float> leafSize; bool voxel_active = false;                   // voxel grid
    double min_dist = 0.0; bool dist_active = false;                          // norm filter
    int rate_value = 1; bool rate_active = false;                             // keep every n-th point
    float fov_angle = 0.f; bool fov_active = false;                           // field of view
  } filters;
  struct iKFoM {
    struct Mapping {
      int NUM_MATCH_POINTS = 5, MAX_NUM_MATCHES = 2000, MAX_NUM_PC2MATCH = 10000;
      double MAX_DIST_PLANE = 2.0, PLANE_THRESHOLD = 5.e-2;
      struct Octree { int bucket_size = 2; float min_extent = 0.2f; bool downsampling = true; } octree;
    } mapping;
    int MAX_NUM_ITERS = 3;
    std::vector<double> LIMITS;
    bool estimate_extrinsics = false;
    double cov_gyro = 6.e-4, cov_acc = 1.e-2, cov_bias_gyro = 1.e-5, cov_bias_acc = 3.e-4;
  } ikfom;
  bool gravity_align = false, calibrate_accel = false, calibrate_gyro = false, time_offset = false, end_of_sweep = false;
  bool debug = false, verbose = false;
  int sensor_type = 1, num_threads = 1;
  double imu_calib_time = 3.0;
  int gpu_device = 0;          // extension: CUDA ordinal used by Mapper (defaults keep existing YAMLs valid)
};
