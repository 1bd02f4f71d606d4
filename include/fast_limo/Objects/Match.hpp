// fast_limo::Plane / fast_limo::Match as data carriers (fast_limo/Objects/Plane.hpp:24-57, Match.hpp:24-46): on the
// B200 path the fit happens inside the measurement kernel; these records are what Mapper::match / Localizer::get_matches
// hand to the wrapper's marker code (ROSutils.hpp:216-252).
#pragma once
#include "fast_limo/Common.hpp"
#include "fast_limo/Objects/State.hpp"

class fast_limo::Plane {
 public:
  Plane() = default;
  Plane(const Eigen::Vector4f& abcd, bool ok) : n_ABCD(abcd), is_plane(ok) {}
  Eigen::Vector4f get_normal() const { return n_ABCD; }
  bool good_fit() const { return is_plane; }
  float dist2plane(const Eigen::Vector3f& p) const { return n_ABCD(0) * p(0) + n_ABCD(1) * p(1) + n_ABCD(2) * p(2) + n_ABCD(3); }
  Eigen::Vector4f n_ABCD;
  bool is_plane = false;
};

class fast_limo::Match {
 public:
  fast_limo::Plane plane;
  float dist = 0.f;
  Match() = default;
  Match(const Eigen::Vector3f& p_global_, const Eigen::Vector3f& p_local_, const fast_limo::Plane& H, float d)
      : plane(H), dist(d), p_global(p_global_), p_local(p_local_) {}
  bool lisanAlGaib() const { return plane.good_fit(); }
  Eigen::Vector4f get_4Dglobal() const { return Eigen::Vector4f(p_global(0), p_global(1), p_global(2), 1.f); }
  Eigen::Vector4f get_4Dlocal() const { return Eigen::Vector4f(p_local(0), p_local(1), p_local(2), 1.f); }
  Eigen::Vector3f get_global_point() const { return p_global; }
  Eigen::Vector3f get_local_point() const { return p_local; }

 private:
  Eigen::Vector3f p_global, p_local;
};
