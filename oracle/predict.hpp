// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// CPU restatement of the IMU-rate side of the filter, kept in the reference's generic shape: dense process
// Jacobians, lists of vect / SO3 / S2 sub-states, the same loop order.  All float64.  Parity status: UNPINNED
// by the reference (no tests / golden vectors; Eigen and Boost are absent, so the original cannot be compiled
// here); tests/test_imu_predict.py cross-checks it against an independent numpy / scipy evaluation.
//
// Follows (paths relative to /root/reference/include):
//   IKFoM/IKFoM_toolkit/esekfom/esekfom.hpp:279-384   esekf::predict                      -> predict()
//   IKFoM/use-ikfom.cpp:46-91                         get_f, df_dx, df_dw                 -> process_f / _fx / _fw
//   IKFoM/IKFoM_toolkit/mtk/build_manifold.hpp:109-111,195-197   sub-state lists, compound oplus
//   IKFoM/IKFoM_toolkit/mtk/types/vect.hpp:124-126, SOn.hpp:242-245, S2.hpp:129-134      oplus of each type
//   fast_limo/Modules/Localizer.cpp:583-608           propagateImu(imu): Q, predict, push -> propagate()
//   fast_limo/Objects/State.cpp:38-69                 State(state_ikfom, t, a, w)         -> Frame
//   fast_limo/Modules/Localizer.cpp:855-913           integrateImu, propagatedFromTimeRange -> frames_in_range()
#pragma once
#include <deque>
#include <vector>

#include "ekf.hpp"
#include "prep.hpp"

namespace orc {

constexpr int NFLAT = 24;   // flattened dimension: every sub-state 3 wide (S2 is stored as a 3-vector)
constexpr int NNOISE = 12;

struct SubState { int idx, dim, dof; };   // position in the error state, in the flattened state, width
// declaration order of use-ikfom.hpp:12-21
static const SubState kVect[] = {{0, 0, 3}, {9, 9, 3}, {12, 12, 3}, {15, 15, 3}, {18, 18, 3}};
static const SubState kSO3[] = {{3, 3, 3}, {6, 6, 3}};
static const SubState kS2[] = {{21, 21, 2}};

inline void process_f(const EkfState& s, const double acc[3], const double gyro[3], double f[NFLAT]) {
  for (int i = 0; i < NFLAT; ++i) f[i] = 0;
  double R[9];
  quat_R(s.rot, R);
  const double a[3] = {acc[0] - s.ba[0], acc[1] - s.ba[1], acc[2] - s.ba[2]};
  for (int i = 0; i < 3; ++i) {
    f[i] = s.vel[i];
    f[3 + i] = gyro[i] - s.bg[i];
    f[12 + i] = (R[3 * i] * a[0] + R[3 * i + 1] * a[1] + R[3 * i + 2] * a[2]) + s.grav[i];
  }
}

inline void process_fx(const EkfState& s, const double acc[3], double fx[NFLAT * NDOF]) {
  for (int i = 0; i < NFLAT * NDOF; ++i) fx[i] = 0;
  double R[9], K[9], RK[9], Mx[6];
  quat_R(s.rot, R);
  const double a[3] = {acc[0] - s.ba[0], acc[1] - s.ba[1], acc[2] - s.ba[2]};
  hat(a, K);
  matmul(R, K, RK, 3, 3, 3);
  const double zero2[2] = {0, 0};
  s2_Mx(s.grav, zero2, Mx);
  for (int i = 0; i < 3; ++i) {
    fx[i * NDOF + 12 + i] = 1;
    fx[(3 + i) * NDOF + 15 + i] = -1;
    for (int j = 0; j < 3; ++j) {
      fx[(12 + i) * NDOF + 3 + j] = -RK[3 * i + j];
      fx[(12 + i) * NDOF + 18 + j] = -R[3 * i + j];
    }
    for (int j = 0; j < 2; ++j) fx[(12 + i) * NDOF + 21 + j] = Mx[2 * i + j];
  }
}

inline void process_fw(const EkfState& s, double fw[NFLAT * NNOISE]) {
  for (int i = 0; i < NFLAT * NNOISE; ++i) fw[i] = 0;
  double R[9];
  quat_R(s.rot, R);
  for (int i = 0; i < 3; ++i) {
    fw[(3 + i) * NNOISE + i] = -1;
    fw[(15 + i) * NNOISE + 6 + i] = 1;
    fw[(18 + i) * NNOISE + 9 + i] = 1;
    for (int j = 0; j < 3; ++j) fw[(12 + i) * NNOISE + 3 + j] = -R[3 * i + j];
  }
}

// compound oplus: x <- x (+) f * dt, sub-state by sub-state
inline void state_oplus(EkfState& x, const double f[NFLAT], double dt) {
  for (int i = 0; i < 3; ++i) x.pos[i] += dt * f[i];
  for (const SubState& s : kSO3) {
    double* q = s.idx == 3 ? x.rot : x.offR;
    double e[4], r[4];
    e[3] = mtk_exp(e, f + s.dim, dt / 2);
    quat_mul(q, e, r);
    for (int i = 0; i < 4; ++i) q[i] = r[i];
  }
  for (int i = 0; i < 3; ++i) {
    x.offT[i] += dt * f[9 + i];
    x.vel[i] += dt * f[12 + i];
    x.bg[i] += dt * f[15 + i];
    x.ba[i] += dt * f[18 + i];
  }
  {
    double e[4], R[9];
    e[3] = mtk_exp(e, f + 21, dt / 2);
    quat_R(e, R);
    const double v[3] = {x.grav[0], x.grav[1], x.grav[2]};
    for (int i = 0; i < 3; ++i) x.grav[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
  }
}

// esekf::predict.  Q is the full 12 x 12 process noise covariance (row-major).
inline void predict(EkfState& x, double* P, const double acc[3], const double gyro[3], double dt, const double* Q) {
  double f[NFLAT];
  std::vector<double> fx(NFLAT * NDOF), fw(NFLAT * NNOISE), fx_fin(NDOF * NDOF, 0.0), fw_fin(NDOF * NNOISE, 0.0);
  process_f(x, acc, gyro, f);
  process_fx(x, acc, fx.data());
  process_fw(x, fw.data());
  const EkfState before = x;
  state_oplus(x, f, dt);

  std::vector<double> F1(NDOF * NDOF, 0.0);
  for (int i = 0; i < NDOF; ++i) F1[i * NDOF + i] = 1;
  for (const SubState& s : kVect)
    for (int j = 0; j < s.dof; ++j) {
      for (int i = 0; i < NDOF; ++i) fx_fin[(s.idx + j) * NDOF + i] = fx[(s.dim + j) * NDOF + i];
      for (int i = 0; i < NNOISE; ++i) fw_fin[(s.idx + j) * NNOISE + i] = fw[(s.dim + j) * NNOISE + i];
    }
  for (const SubState& s : kSO3) {
    double seg[3], e[4], E[9], A[9];
    for (int i = 0; i < 3; ++i) seg[i] = -1 * f[s.dim + i] * dt;
    e[3] = mtk_exp(e, seg, 0.0);   // scalar(1/2): integer division => scale 0 => identity (esekfom.hpp:312)
    quat_R(e, E);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) F1[(s.idx + i) * NDOF + s.idx + j] = E[3 * i + j];
    A_matrix(seg, A);
    for (int c = 0; c < NDOF; ++c)
      for (int i = 0; i < 3; ++i) {
        double v = 0;
        for (int k = 0; k < 3; ++k) v += A[3 * i + k] * fx[(s.dim + k) * NDOF + c];
        fx_fin[(s.idx + i) * NDOF + c] = v;
      }
    for (int c = 0; c < NNOISE; ++c)
      for (int i = 0; i < 3; ++i) {
        double v = 0;
        for (int k = 0; k < 3; ++k) v += A[3 * i + k] * fw[(s.dim + k) * NNOISE + c];
        fw_fin[(s.idx + i) * NNOISE + c] = v;
      }
  }
  for (const SubState& s : kS2) {
    double seg[3], e[4], E[9], Nx[6], Mx[6], Hb[9], A[9], At[9];
    for (int i = 0; i < 3; ++i) seg[i] = f[s.dim + i] * dt;
    e[3] = mtk_exp(e, seg, 0.0);   // scalar(1/2) again (:344)
    quat_R(e, E);
    s2_Nx_yy(x.grav, Nx);
    const double zero2[2] = {0, 0};
    s2_Mx(before.grav, zero2, Mx);
    double NE[6], NEM[4];
    matmul(Nx, E, NE, 2, 3, 3);
    matmul(NE, Mx, NEM, 2, 3, 2);
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) F1[(s.idx + i) * NDOF + s.idx + j] = NEM[2 * i + j];
    hat(before.grav, Hb);
    A_matrix(seg, A);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) At[3 * i + j] = A[3 * j + i];
    double T1[6], T2[6];
    matmul(NE, Hb, T1, 2, 3, 3);
    matmul(T1, At, T2, 2, 3, 3);
    for (int i = 0; i < 6; ++i) T2[i] = -T2[i];
    for (int c = 0; c < NDOF; ++c)
      for (int i = 0; i < 2; ++i) {
        double v = 0;
        for (int k = 0; k < 3; ++k) v += T2[3 * i + k] * fx[(s.dim + k) * NDOF + c];
        fx_fin[(s.idx + i) * NDOF + c] = v;
      }
    for (int c = 0; c < NNOISE; ++c)
      for (int i = 0; i < 2; ++i) {
        double v = 0;
        for (int k = 0; k < 3; ++k) v += T2[3 * i + k] * fw[(s.dim + k) * NNOISE + c];
        fw_fin[(s.idx + i) * NNOISE + c] = v;
      }
  }
  for (int i = 0; i < NDOF * NDOF; ++i) F1[i] += fx_fin[i] * dt;
  for (double& v : fw_fin) v *= dt;
  std::vector<double> FP(NDOF * NDOF), Ft(NDOF * NDOF), Pn(NDOF * NDOF), GQ(NDOF * NNOISE), Gt(NNOISE * NDOF), N2(NDOF * NDOF);
  for (int i = 0; i < NDOF; ++i)
    for (int j = 0; j < NDOF; ++j) Ft[i * NDOF + j] = F1[j * NDOF + i];
  for (int i = 0; i < NDOF; ++i)
    for (int j = 0; j < NNOISE; ++j) Gt[j * NDOF + i] = fw_fin[i * NNOISE + j];
  matmul(F1.data(), P, FP.data(), NDOF, NDOF, NDOF);
  matmul(FP.data(), Ft.data(), Pn.data(), NDOF, NDOF, NDOF);
  matmul(fw_fin.data(), Q, GQ.data(), NDOF, NNOISE, NNOISE);
  matmul(GQ.data(), Gt.data(), N2.data(), NDOF, NNOISE, NDOF);
  for (int i = 0; i < NDOF * NDOF; ++i) P[i] = Pn[i] + N2[i];
}

// fast_limo::State built from the filter state (float members): prep.hpp's Frame.
inline Frame make_frame(const EkfState& s, double t, const float a[3], const float w[3]) {
  Frame fr;
  fr.time = t;
  for (int i = 0; i < 4; ++i) fr.q[i] = static_cast<float>(s.rot[i]);
  for (int i = 0; i < 3; ++i) {
    fr.p[i] = static_cast<float>(s.pos[i]);
    fr.v[i] = static_cast<float>(s.vel[i]);
    fr.g[i] = static_cast<float>(s.grav[i]);
    fr.bg[i] = static_cast<float>(s.bg[i]);
    fr.ba[i] = static_cast<float>(s.ba[i]);
    fr.a[i] = a[i];
    fr.w[i] = w[i];
  }
  return fr;
}

// The IMU-rate object: filter state + propagated_buffer (front = newest, capacity 2000).
struct Propagator {
  EkfState x;
  std::vector<double> P = std::vector<double>(NDOF * NDOF, 0.0);
  std::deque<Frame> buffer;
  size_t capacity = 2000;

  void propagate(double stamp, double dt, const float lin_accel[3], const float ang_vel[3], const double cov4[4]) {
    double Q[NNOISE * NNOISE] = {0};
    for (int i = 0; i < NNOISE; ++i) Q[i * NNOISE + i] = cov4[i / 3];
    const double acc[3] = {lin_accel[0], lin_accel[1], lin_accel[2]}, gyro[3] = {ang_vel[0], ang_vel[1], ang_vel[2]};
    predict(x, P.data(), acc, gyro, dt, Q);
    if (buffer.size() == capacity) buffer.pop_back();
    buffer.push_front(make_frame(x, stamp, lin_accel, ang_vel));
  }

  // -1: the reference would wait for newer IMU data; 0: "not enough propagated states"; else the frames,
  // forward in time.
  long frames_in_range(double start_time, double end_time, std::vector<Frame>& out) const {
    out.clear();
    if (buffer.empty() || buffer.front().time < end_time) return -1;
    auto it = buffer.begin();
    auto last = it;
    ++it;
    while (it != buffer.end() && it->time >= end_time) {
      last = it;
      ++it;
    }
    while (it != buffer.end() && it->time >= start_time) ++it;
    if (it == buffer.end()) return 0;
    ++it;
    // reverse range [reverse(it), reverse(last)) == forward elements [last, it) visited backwards
    for (auto r = std::deque<Frame>::const_reverse_iterator(it); r != std::deque<Frame>::const_reverse_iterator(last); ++r)
      out.push_back(*r);
    return static_cast<long>(out.size());
  }
};

}  // namespace orc
