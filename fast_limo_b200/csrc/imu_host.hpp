// IMU-rate host side of the filter (product code, float64, no dependencies): the prediction step and the
// ring of propagated states that the deskew stage reads.
//
//   esekf::predict                      IKFoM_toolkit/esekfom/esekfom.hpp:279-384
//   process model f, df/dx, df/dw       IKFoM/use-ikfom.cpp:46-91
//   Q and the push into the ring        fast_limo/Modules/Localizer.cpp:583-608 (propagateImu)
//   State(state_ikfom, t, a, w)         fast_limo/Objects/State.cpp:38-69
//   integrateImu / propagatedFromTimeRange   Localizer.cpp:855-913, ring capacity :54
//
// The reference evaluates predict() generically (lists of vect / SO3 / S2 sub-states, dense 24x23 and 24x12
// Jacobians).  For this state (use-ikfom.hpp:12-21) with this process model most of that is structurally zero;
// what is left is written out block by block:
//
//   f     = [ vel | gyro - bg | 0 | 0 | R (acc - ba) + grav | 0 | 0 | 0 ]                    (flattened, 24)
//   x    <- x (+) f dt :  pos += vel dt;  rot <- rot * Exp((gyro - bg) dt);  vel += (R (acc - ba) + grav) dt
//   F     = I + dt * [ pos,vel: I | rot,bg: -A | vel,rot: -R [acc - ba]x | vel,ba: -R | vel,grav: Mx ],
//           F[grav,grav] = Nx_yy(grav) * Mx(grav, 0),     A = A_matrix(-(gyro - bg) dt)
//   G     = dt * [ rot,ng: -A | vel,na: -R | bg,nbg: I | ba,nba: I ]
//   P    <- F P F^T + G Q G^T
//
// Reproduced on purpose: F[rot,rot] stays the identity — the reference builds that block from an exponential
// with scale scalar(1/2) == 0 (integer division, esekfom.hpp:312), likewise the S2 one (:344); the offset
// rotation and gravity have zero rate, so their (+) is the identity map.
#pragma once
#include <cstddef>
#include <vector>

#include "ekf_host.hpp"

namespace flimo {
namespace ekf {

struct ProcessNoise {   // Config::iKFoM covariances in the order Localizer.cpp:589-592 puts them on Q's diagonal
  double gyro, acc, bias_gyro, bias_acc;
};

inline V3 rot_apply(const Mat<3, 3>& R, const V3& v) {
  return {R(0, 0) * v[0] + R(0, 1) * v[1] + R(0, 2) * v[2], R(1, 0) * v[0] + R(1, 1) * v[1] + R(1, 2) * v[2],
          R(2, 0) * v[0] + R(2, 1) * v[1] + R(2, 2) * v[2]};
}

// One esekf::predict(dt, Q, {acc, gyro}) on (x, P).
inline void predict(State& x, Mat<N, N>& P, const V3& acc, const V3& gyro, double dt, const ProcessNoise& q) {
  const Mat<3, 3> R = rotmat(x.rot);                     // all Jacobians use the state BEFORE the step (:280-285)
  const V3 omega = {gyro[0] - x.bg[0], gyro[1] - x.bg[1], gyro[2] - x.bg[2]};
  const V3 a_b = {acc[0] - x.ba[0], acc[1] - x.ba[1], acc[2] - x.ba[2]};
  const V3 a_w = rot_apply(R, a_b);
  const V3 grav_before = x.grav;
  const Mat<3, 2> Mx = s2_Mx(grav_before, 0.0, 0.0);     // df_dx's gravity block and x_before.S2_Mx (:345)

  // x (+) f dt  (build_manifold.hpp:195-197: every sub-state's oplus in declaration order)
  for (int i = 0; i < 3; ++i) x.pos[i] += dt * x.vel[i];
  x.rot = qmul(x.rot, qexp(omega, dt / 2));
  for (int i = 0; i < 3; ++i) x.vel[i] += dt * (a_w[i] + grav_before[i]);

  const Mat<3, 3> A = A_matrix({-omega[0] * dt, -omega[1] * dt, -omega[2] * dt});
  const Mat<3, 3> RK = R * skew(a_b);

  Mat<N, N> F = Mat<N, N>::identity();
  Mat<N, 12> G = Mat<N, 12>::zero();
  for (int i = 0; i < 3; ++i) {
    F(i, 12 + i) += dt;                                  // d pos / d vel
    for (int j = 0; j < 3; ++j) {
      F(3 + i, 15 + j) += dt * -A(i, j);                 // d rot / d bg   = A * (-I)
      F(12 + i, 3 + j) += dt * -RK(i, j);                // d vel / d rot  = -R [acc - ba]x
      F(12 + i, 18 + j) += dt * -R(i, j);                // d vel / d ba
      G(3 + i, j) = dt * -A(i, j);                       // rot <- gyro noise
      G(12 + i, 3 + j) = dt * -R(i, j);                  // vel <- accelerometer noise
    }
    for (int j = 0; j < 2; ++j) F(12 + i, 21 + j) += dt * Mx(i, j);   // d vel / d grav
    G(15 + i, 6 + i) = dt;
    G(18 + i, 9 + i) = dt;
  }
  const Mat<2, 2> Fg = s2_Nx_yy(x.grav) * Mx;            // Nx of the state after the step * I * Mx (:343-354)
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) F(21 + i, 21 + j) = Fg(i, j);

  const double qd[12] = {q.gyro, q.gyro, q.gyro, q.acc, q.acc, q.acc, q.bias_gyro, q.bias_gyro, q.bias_gyro,
                         q.bias_acc, q.bias_acc, q.bias_acc};
  const Mat<N, N> FP = F * P;
  Mat<N, N> Pn = FP * F.T();
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) {
      double s = 0.0;
      for (int k = 0; k < 12; ++k) s += G(r, k) * qd[k] * G(c, k);
      Pn(r, c) += s;
    }
  P = Pn;
}

// fast_limo::State as the deskew stage reads it (flimo_frame has the same members in the same order).
struct PropagatedState {
  double time;
  float q[4];
  float p[3], v[3], w[3], a[3], bg[3], ba[3], g[3];
};

inline PropagatedState make_propagated(const State& x, double stamp, const float lin_accel[3], const float ang_vel[3]) {
  PropagatedState s;
  s.time = stamp;
  for (int i = 0; i < 4; ++i) s.q[i] = static_cast<float>(x.rot[i]);
  for (int i = 0; i < 3; ++i) {
    s.p[i] = static_cast<float>(x.pos[i]);
    s.v[i] = static_cast<float>(x.vel[i]);
    s.g[i] = static_cast<float>(x.grav[i]);
    s.bg[i] = static_cast<float>(x.bg[i]);
    s.ba[i] = static_cast<float>(x.ba[i]);
    s.w[i] = ang_vel[i];
    s.a[i] = lin_accel[i];
  }
  return s;
}

// boost::circular_buffer<State> propagated_buffer used with push_front: index 0 is the newest state, a push
// into a full ring drops the oldest.
class PropagatedRing {
 public:
  explicit PropagatedRing(std::size_t capacity = 2000) : buf_(capacity), head_(0), size_(0) {}
  void clear() { head_ = size_ = 0; }
  std::size_t size() const { return size_; }
  std::size_t capacity() const { return buf_.size(); }
  void push_front(const PropagatedState& s) {
    head_ = (head_ + buf_.size() - 1) % buf_.size();
    buf_[head_] = s;
    if (size_ < buf_.size()) ++size_;
  }
  const PropagatedState& at(std::size_t i) const { return buf_[(head_ + i) % buf_.size()]; }   // 0 = newest

  // propagatedFromTimeRange + integrateImu.  Returns -1 when the newest state is older than end_time (the
  // reference blocks on a condition variable there), otherwise the number of frames (0: "not enough
  // propagated states"), written oldest first: one state before start_time up to the first one at or
  // after end_time.
  long frames(double start_time, double end_time, std::vector<PropagatedState>& out) const {
    out.clear();
    if (size_ == 0 || at(0).time < end_time) return -1;
    std::size_t it = 1, last = 0;
    while (it < size_ && at(it).time >= end_time) last = it++;
    while (it < size_ && at(it).time >= start_time) ++it;
    if (it == size_) return 0;
    for (std::size_t i = it + 1; i-- > last;) out.push_back(at(i));
    return static_cast<long>(out.size());
  }

 private:
  std::vector<PropagatedState> buf_;
  std::size_t head_, size_;
};

}  // namespace ekf
}  // namespace flimo
