// Pose constants of one measurement pass, float32 exactly as the reference builds them; shared by the host
// (flimo_api.cu) and the device-side update (match_kernel.cu: the CTA that finishes a pass derives the next pose).
#pragma once
#include "ekf_host.hpp"      // FLIMO_HD
#include "flimo_dev.cuh"

namespace flimo {

// --- pose constants, float32 exactly as the reference builds them -----------------------------
template <typename T>
FLIMO_HD inline void quat_matrix(const T q[4], T R[9]) {   // Eigen QuaternionBase::toRotationMatrix
  const T tx = T(2) * q[0], ty = T(2) * q[1], tz = T(2) * q[2];
  const T twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const T txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const T tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = T(1) - (tyy + tzz); R[1] = txy - twz;          R[2] = txz + twy;
  R[3] = txy + twz;          R[4] = T(1) - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;          R[7] = tyz + twx;          R[8] = T(1) - (txx + tyy);
}

// [R|t]^-1 = [R^T | -R^T t] with Eigen's fixed-size evaluation order (State.cpp:145-153)
FLIMO_HD inline void rigid_inverse(const float R[9], const float t[3], float Ri[9], float ti[3]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Ri[3 * r + c] = R[3 * c + r];
  for (int r = 0; r < 3; ++r) {
#if defined(__CUDA_ARCH__)
    const float a = (-Ri[3 * r]) * t[0], b = (-Ri[3 * r + 1]) * t[1], c = (-Ri[3 * r + 2]) * t[2];   // --fmad=false: no contraction
    const float bc = b + c;
#else
    const volatile float a = (-Ri[3 * r]) * t[0], b = (-Ri[3 * r + 1]) * t[1], c = (-Ri[3 * r + 2]) * t[2];
    const volatile float bc = b + c;
#endif
    ti[r] = a + bc;
  }
}

FLIMO_HD inline void make_pose(const double s[14], PoseConsts& pc) {
  // State::State(const state_ikfom&) casts (State.cpp:38-55)
  float q[4], qLI[4], p[3], pLI[3];
  for (int i = 0; i < 3; ++i) p[i] = (float)s[i];
  for (int i = 0; i < 4; ++i) q[i] = (float)s[3 + i];
  for (int i = 0; i < 4; ++i) qLI[i] = (float)s[7 + i];
  for (int i = 0; i < 3; ++i) pLI[i] = (float)s[11 + i];
  quat_matrix<float>(q, pc.R_wb);
  for (int i = 0; i < 3; ++i) pc.t_wb[i] = p[i];
  rigid_inverse(pc.R_wb, p, pc.Rinv_wb, pc.tinv_wb);
  float R_LI[9];
  quat_matrix<float>(qLI, R_LI);
  rigid_inverse(R_LI, pLI, pc.Rinv_LI, pc.tinv_LI);
  // Localizer.cpp:554-555: conjugate of the DOUBLE quaternion -> matrix -> cast<float>
  double qc[4] = {-s[3], -s[4], -s[5], s[6]}, Rd[9];
  quat_matrix<double>(qc, Rd);
  for (int i = 0; i < 9; ++i) pc.Rd_wb_inv[i] = (float)Rd[i];
  double qc2[4] = {-s[7], -s[8], -s[9], s[10]};
  quat_matrix<double>(qc2, Rd);
  for (int i = 0; i < 9; ++i) pc.Rd_LI_inv[i] = (float)Rd[i];
}

// make_pose in four independent parts (the filter kernel runs them on four warps): identical results.
FLIMO_HD inline void make_pose_part(const double s[14], PoseConsts& pc, int part) {
  if (part == 0) {
    float q[4], p[3];
    for (int i = 0; i < 3; ++i) p[i] = (float)s[i];
    for (int i = 0; i < 4; ++i) q[i] = (float)s[3 + i];
    float R[9];
    quat_matrix<float>(q, R);
    for (int i = 0; i < 9; ++i) pc.R_wb[i] = R[i];
    for (int i = 0; i < 3; ++i) pc.t_wb[i] = p[i];
    rigid_inverse(R, p, pc.Rinv_wb, pc.tinv_wb);
  } else if (part == 1) {
    float qLI[4], pLI[3], R_LI[9];
    for (int i = 0; i < 4; ++i) qLI[i] = (float)s[7 + i];
    for (int i = 0; i < 3; ++i) pLI[i] = (float)s[11 + i];
    quat_matrix<float>(qLI, R_LI);
    rigid_inverse(R_LI, pLI, pc.Rinv_LI, pc.tinv_LI);
  } else if (part == 2) {
    double qc[4] = {-s[3], -s[4], -s[5], s[6]}, Rd[9];
    quat_matrix<double>(qc, Rd);
    for (int i = 0; i < 9; ++i) pc.Rd_wb_inv[i] = (float)Rd[i];
  } else {
    double qc2[4] = {-s[7], -s[8], -s[9], s[10]}, Rd[9];
    quat_matrix<double>(qc2, Rd);
    for (int i = 0; i < 9; ++i) pc.Rd_LI_inv[i] = (float)Rd[i];
  }
}

}  // namespace flimo
