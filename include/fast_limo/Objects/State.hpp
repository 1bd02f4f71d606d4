// fast_limo::State (fast_limo/Objects/State.hpp:23-66): the float mirror of the filter state the wrapper publishes.
#pragma once
#include "fast_limo/Common.hpp"

class fast_limo::State {
 public:
  Eigen::Vector3f p;        // position, world frame
  Eigen::Quaternionf q;     // orientation, world frame
  Eigen::Vector3f v, g, w, a;
  Eigen::Quaternionf qLI;   // offsets (LiDAR -> base_link)
  Eigen::Vector3f pLI;
  double time = 0.0;
  struct IMUbias { Eigen::Vector3f gyro, accel; } b;

  State() { g = Eigen::Vector3f(0.f, 0.f, -9.807f); }
  State(const state_ikfom& s) : State(s, 0.0) {}                            // State.cpp:38-55: double -> float casts
  State(const state_ikfom& s, double t) : time(t) {
    p = s.pos.cast<float>(); q = s.rot.cast<float>(); v = s.vel.cast<float>(); g = s.grav.cast<float>();
    qLI = s.offset_R_L_I.cast<float>(); pLI = s.offset_T_L_I.cast<float>();
    b.gyro = s.bg.cast<float>(); b.accel = s.ba.cast<float>();
  }
  State(const state_ikfom& s, double t, Eigen::Vector3f a_, Eigen::Vector3f w_) : State(s, t) { a = a_; w = w_; }

  Eigen::Matrix4f get_RT() const { return rt(q, p); }                       // State.cpp:136-143
  Eigen::Matrix4f get_extr_RT() const { return rt(qLI, pLI); }              // State.cpp:155-162

 private:
  static Eigen::Matrix4f rt(const Eigen::Quaternionf& q_, const Eigen::Vector3f& t_) {
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    const Eigen::Matrix3f R = q_.toRotationMatrix();
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) T(r, c) = R(r, c);
      T(r, 3) = t_(r);
    }
    return T;
  }
};
