// fast_limo::Mapper over libflimo_cuda (reference: fast_limo/Modules/Mapper.cpp:23-114).
#include "fast_limo/Modules/Mapper.hpp"

#include <stdexcept>

namespace fast_limo {

Mapper::Mapper() {                                     // Mapper.cpp:23-32: defaults until set_config is called
  config = Config::iKFoM::Mapping();
}

Mapper::~Mapper() {
  if (h_) flimo_destroy(h_);
}

void Mapper::check(int rc) const {
  if (rc != FLIMO_OK) throw std::runtime_error(std::string("libflimo_cuda: ") + flimo_last_error(h_));
}

void Mapper::set_num_threads(int n) { num_threads_ = n < 1 ? 1 : n; }

void Mapper::set_config(const Config::iKFoM::Mapping& cfg) { set_config(cfg, estimate_extrinsics_, device_); }

void Mapper::set_config(const Config::iKFoM::Mapping& cfg, bool estimate_extrinsics, int device) {
  config = cfg;
  estimate_extrinsics_ = estimate_extrinsics;
  device_ = device;
  flimo_cfg c;
  flimo_cfg_default(&c);
  c.NUM_MATCH_POINTS = cfg.NUM_MATCH_POINTS;
  c.MAX_NUM_MATCHES = cfg.MAX_NUM_MATCHES;
  c.MAX_NUM_PC2MATCH = cfg.MAX_NUM_PC2MATCH;
  c.MAX_DIST_PLANE = cfg.MAX_DIST_PLANE;
  c.PLANE_THRESHOLD = cfg.PLANE_THRESHOLD;
  c.estimate_extrinsics = estimate_extrinsics ? 1 : 0;
  c.octree_bucket_size = cfg.octree.bucket_size;           // accepted and ignored, like the reference (Octree.hpp:178-180)
  c.octree_min_extent = cfg.octree.min_extent;
  c.octree_downsampling = cfg.octree.downsampling ? 1 : 0;
  if (h_) flimo_destroy(h_);
  h_ = nullptr;
  const int rc = flimo_create(&c, device, &h_);
  if (rc != FLIMO_OK) throw std::runtime_error(std::string("libflimo_cuda: ") + flimo_last_error(nullptr));
}

bool Mapper::exists() { return h_ && flimo_map_exists(h_) != 0; }

int Mapper::size() {
  size_t n = 0;
  if (h_) check(flimo_map_size(h_, &n));
  return (int)n;
}

double Mapper::last_time() { return h_ ? flimo_map_last_time(h_) : -1.0; }

// Mapper.cpp:59-86: world points, 5-NN planes and distances of the first MAX_NUM_PC2MATCH points; the good ones, in order.
Matches Mapper::match(State s, pcl::PointCloud<PointType>::Ptr& pc) {
  matches.clear();
  if (!exists() || !pc || pc->points.empty()) return matches;           // Mapper.cpp:61
  check(flimo_scan_set(h_, &pc->points[0].x, pc->points.size(), sizeof(PointType)));
  const double st14[14] = {s.p(0), s.p(1), s.p(2), s.q.x(), s.q.y(), s.q.z(), s.q.w(),
                           s.qLI.x(), s.qLI.y(), s.qLI.z(), s.qLI.w(), s.pLI(0), s.pLI(1), s.pLI(2)};
  size_t n = 0;
  check(flimo_match_debug(h_, st14, nullptr, 0, &n));
  std::vector<float> rec(16 * (n ? n : 1));
  check(flimo_match_debug(h_, st14, rec.data(), n, &n));
  matches.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    const float* r = &rec[16 * i];
    if (!(r[8] > 0.5f)) continue;                                        // Match::lisanAlGaib()
    const PointType& pl = pc->points[i];
    matches.emplace_back(Eigen::Vector3f(r[0], r[1], r[2]), Eigen::Vector3f(pl.x, pl.y, pl.z),
                         Plane(Eigen::Vector4f(r[3], r[4], r[5], r[6]), true), r[7]);
  }
  return matches;
}

void Mapper::add(pcl::PointCloud<PointType>::Ptr& pc, double time) {   // Mapper.cpp:88-96
  if (!pc || pc->points.size() < 1) return;
  if (!h_) throw std::runtime_error("fast_limo::Mapper: set_config has not been called");
  check(flimo_map_add(h_, &pc->points[0].x, pc->points.size(), sizeof(PointType), time));
}

}  // namespace fast_limo
