// Compiles the C++ façade against include/flimo.h with plain g++ (no CUDA, no Eigen/PCL) and drives
// the host-only part of the ABI: a host-only handle must run the pass-wise filter state machine and
// must refuse every GPU entry point loudly.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "fast_limo_gpu/Localizer.hpp"
#include "fast_limo_gpu/Mapper.hpp"

struct Pt {            // fast_limo::Point layout (32 bytes)
  float x, y, z, pad, intensity, t0;
  double ts;
};
struct Cloud {
  std::vector<Pt> points;
};
struct Octree { int bucket_size = 2; float min_extent = 0.2f; bool downsampling = true; };
struct Mapping {
  int NUM_MATCH_POINTS = 5, MAX_NUM_MATCHES = 2000, MAX_NUM_PC2MATCH = 10000;
  double MAX_DIST_PLANE = 2.0, PLANE_THRESHOLD = 0.05;
  Octree octree;
};

int main() {
  static_assert(sizeof(Pt) == 32, "fast_limo::Point is 32 bytes");
  auto& M = fast_limo_gpu::Mapper::getInstance();
  Mapping cfg;
  M.set_config(cfg, true, /*device=*/-1);          // host-only handle (no GPU in the build container)
  if (M.exists() || M.size() != 0 || M.last_time() != -1.0) return 1;
  auto pc = std::make_shared<Cloud>();
  pc->points.resize(4);
  bool threw = false;
  try { M.add(pc, 0.0); } catch (const std::exception& e) { threw = std::strstr(e.what(), "host-only") != nullptr; }
  if (!threw) return 2;
  // the scan front end (Localizer side) must refuse a host-only handle as loudly
  struct Filters {
    std::vector<float> cropBoxMin{-1, -1, -1}, cropBoxMax{1, 1, 1}, leafSize{0.5f, 0.5f, 0.5f};
    bool crop_active = true, voxel_active = true, dist_active = true, rate_active = false, fov_active = false;
    double min_dist = 3.0;
    int rate_value = 1;
    float fov_angle = 3.0f;
  } filters;
  fast_limo_gpu::Localizer L(M.handle());
  L.set_filters(filters, /*sensor_type=*/1, /*end_of_sweep=*/false);
  threw = false;
  double t_last = 0.0;
  try { L.filter_and_sort(pc, 100.0, t_last); } catch (const std::exception& e) { threw = std::strstr(e.what(), "host-only") != nullptr; }
  if (!threw) return 8;
  threw = false;
  try { M.prefetch_scan(pc); } catch (const std::exception& e) { threw = std::strstr(e.what(), "host-only") != nullptr; }
  if (!threw) return 9;
  // filter state machine from C++
  double x[26] = {0}, P[529] = {0}, lim[23], HTH[144] = {0}, HTh[12] = {0};
  x[6] = 1.0; x[10] = 1.0; x[25] = -9.809;
  for (int i = 0; i < 23; ++i) { P[i * 23 + i] = 1.0; lim[i] = 0.001; }
  for (int i = 0; i < 12; ++i) { HTH[i * 12 + i] = 1000.0 + i; HTh[i] = 0.01 * (i + 1); }
  flimo_handle h = M.handle();
  if (flimo_ekf_begin(h, x, P, 2, lim, 0.001, 5.0) != FLIMO_OK) return 3;
  int done = 0, passes = 0;
  while (!done) {
    double cur[26];
    if (flimo_ekf_state(h, cur) != FLIMO_OK) return 4;
    if (flimo_ekf_step(h, HTH, HTh, 500, &done) != FLIMO_OK) return 5;
    ++passes;
  }
  if (flimo_ekf_end(h, x, P) != FLIMO_OK) return 6;
  if (passes < 1 || passes > 3 || !(std::fabs(x[0]) > 1e-9) || !(P[0] < 1.0)) return 7;
  // IMU rate: 40 predictions at 200 Hz, then the frames of a scan interval (host algebra, works without a device)
  struct Imu { double stamp, dt; float ang_vel[3], lin_accel[3]; } imu{0.0, 0.005, {0.0f, 0.0f, 0.2f}, {0.5f, 0.0f, 9.809f}};
  struct IKFoM { double cov_gyro = 6e-4, cov_acc = 1e-2, cov_bias_gyro = 1e-5, cov_bias_acc = 3e-4; } ikfom;
  const double P00_before = P[0];
  for (int i = 1; i <= 40; ++i) {
    imu.stamp = 0.005 * i;
    L.propagateImu(x, P, imu, ikfom);
  }
  if (!(P[0] > P00_before) || !(x[14] > 0.09 && x[14] < 0.11)) return 10;       // covariance grows; v_x = 0.5 * 0.2 s
  const std::vector<flimo_frame> frames = L.integrateImu(0.0512, 0.1512);
  if (frames.size() != 22 || frames.front().time != 0.05 || frames.back().time != 0.155) return 11;
  if (!L.integrateImu(0.0, 0.1).empty()) return 12;                              // first scan: nothing before t = 0
  threw = false;
  try { L.integrateImu(0.1, 0.5); } catch (const std::exception& e) { threw = std::strstr(e.what(), "IMU behind") != nullptr; }
  if (!threw) return 13;
  std::printf("facade ok passes=%d x0=%.3e P00=%.3e\n", passes, x[0], P[0]);
  return 0;
}
