// fast_limo::Mapper call surface (fast_limo/Modules/Mapper.hpp:49-60) over libflimo_cuda.
//
// Drop-in façade for the registration hot path: same method names, argument meaning and error
// behaviour (prints nothing, never throws; an empty map yields no matches) as the reference class,
// with the GPU doing Mapper::match + Localizer::calculate_H + H^T H / H^T h in one call.
// On a ROS machine `PointCloudT` is pcl::PointCloud<fast_limo::Point> (32-byte points,
// fast_limo/Common.hpp:100-113) and `StateT` is fast_limo::State built from state_ikfom; here they are
// template parameters so that this header compiles without PCL/Eigen (neither is installed in the
// build image).  Requirements on the types:
//   PointCloudT: `.points` contiguous container of PODs that start with float x, y, z
//   state14    : pos[3], rot xyzw[4], offset_R_L_I xyzw[4], offset_T_L_I[3] (state_ikfom order)
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>

#include "../flimo.h"

namespace fast_limo_gpu {

// Reduced-form result of one h_share_model evaluation (what esekfom.hpp:1722-1729 consumes).
struct NormalEquations {
  double HTH[144];
  double HTh[12];
  std::int64_t n_valid = 0;   // Mapper::match(...).size()
  std::int64_t n_rows = 0;    // rows calculate_H would have produced (first-N cap applied)
  double sum_sq_res = 0.0;
};

class Mapper {
 public:
  // Mapper::getInstance(): one process-wide instance, like the reference.
  static Mapper& getInstance() {
    static Mapper* mapper = new Mapper();
    return *mapper;
  }

  void set_num_threads(int) {}   // kept for source compatibility (OpenMP thread count; unused on the GPU)

  // Mapper::set_config(const Config::iKFoM::Mapping&): MappingT needs the reference's field names.
  template <typename MappingT>
  void set_config(const MappingT& cfg, bool estimate_extrinsics = true, int device = 0) {
    flimo_cfg c;
    flimo_cfg_default(&c);
    c.NUM_MATCH_POINTS = cfg.NUM_MATCH_POINTS;
    c.MAX_NUM_MATCHES = cfg.MAX_NUM_MATCHES;
    c.MAX_NUM_PC2MATCH = cfg.MAX_NUM_PC2MATCH;
    c.MAX_DIST_PLANE = cfg.MAX_DIST_PLANE;
    c.PLANE_THRESHOLD = cfg.PLANE_THRESHOLD;
    c.octree_bucket_size = cfg.octree.bucket_size;
    c.octree_min_extent = cfg.octree.min_extent;
    c.octree_downsampling = cfg.octree.downsampling ? 1 : 0;
    c.estimate_extrinsics = estimate_extrinsics ? 1 : 0;
    reset(c, device);
  }

  bool exists() { return h_ && flimo_map_exists(h_) != 0; }
  int size() {
    std::size_t n = 0;
    if (h_) flimo_map_size(h_, &n);
    return static_cast<int>(n);
  }
  double last_time() { return h_ ? flimo_map_last_time(h_) : -1.0; }

  // Mapper::add(pc, time)
  template <typename PointCloudPtr>
  void add(PointCloudPtr& pc, double time) {
    if (pc->points.size() < 1) return;
    check(flimo_map_add(h_, reinterpret_cast<const float*>(pc->points.data()), pc->points.size(),
                        sizeof(pc->points[0]), time));
  }

  // Binds Localizer::pc2match for the passes of one update (use-ikfom.cpp:18).
  template <typename PointCloudPtr>
  void bind_scan(PointCloudPtr& pc) {
    check(flimo_scan_set(h_, reinterpret_cast<const float*>(pc->points.data()), pc->points.size(), sizeof(pc->points[0])));
  }

  // Optional: start the host-to-device copy of the NEXT scan; it overlaps the current registration and the
  // next bind_scan of the same cloud picks it up.
  template <typename PointCloudPtr>
  void prefetch_scan(PointCloudPtr& pc) {
    check(flimo_scan_prefetch(h_, reinterpret_cast<const float*>(pc->points.data()), pc->points.size(), sizeof(pc->points[0])));
  }

  // pcl::transformPointCloud(*pc2match, *final_scan, state.get_RT()) (Localizer.cpp:361-374) on the device copy.
  std::size_t scan_to_world(const double state14[14], float* out_xyz, std::size_t cap_points) {
    std::size_t n = 0;
    check(flimo_scan_to_world(h_, state14, out_xyz, cap_points, &n));
    return n;
  }

  // Mapper::match(State, pc) + Localizer::calculate_H + H^T H / H^T h for the bound scan.
  NormalEquations match(const double state14[14]) {
    NormalEquations ne;
    check(flimo_match_reduce(h_, state14, ne.HTH, ne.HTh, &ne.n_valid, &ne.n_rows, &ne.sum_sq_res));
    return ne;
  }

  // esekf::update_iterated_dyn_share_modified(R, D, solve_time) on the bound scan.
  int update(double state26[26], double P529[529], int max_num_iters, const double limits23[23], double R = 0.001,
             double D = 5.0) {
    int passes = 0;
    check(flimo_update(h_, state26, P529, max_num_iters, limits23, R, D, &passes));
    return passes;
  }

  flimo_handle handle() { return h_; }

 private:
  Mapper() = default;
  Mapper(const Mapper&) = delete;
  Mapper& operator=(const Mapper&) = delete;
  void reset(const flimo_cfg& c, int device) {
    if (h_) flimo_destroy(h_);
    h_ = nullptr;
    const int rc = flimo_create(&c, device, &h_);
    if (rc != FLIMO_OK) throw std::runtime_error(std::string("flimo_create: ") + flimo_last_error(nullptr));
  }
  void check(int rc) {
    if (rc != FLIMO_OK) throw std::runtime_error(std::string("libflimo_cuda: ") + flimo_last_error(h_));
  }
  flimo_handle h_ = nullptr;
};

}  // namespace fast_limo_gpu
