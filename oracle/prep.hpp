// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the scan preparation that feeds the
// registration hot path:
//   point filters        fast_limo/Modules/Localizer.cpp:262-302 (+ isInRange :866-869)
//   deskewPointCloud     fast_limo/Modules/Localizer.cpp:733-853
//   State::update(t)     fast_limo/Objects/State.cpp:76-119, get_RT :136-143, get_RT_inv :145-153
//   binary_search_tailored   fast_limo/Utils/Algorithms.hpp:25-38
//   voxel grid           fast_limo/Modules/Localizer.cpp:313-321 -> pcl::VoxelGrid<PointT>::applyFilter
//                        (PCL 1.10, filters/include/pcl/filters/impl/voxel_grid.hpp:211-400; PCL is a
//                        third-party dependency that is NOT in /root/reference — find_package(PCL 1.8),
//                        CMakeLists.txt:14 — its published algorithm is restated here)
//
// Parity status: UNPINNED.  The reference has no tests for this code and cannot be built here.  Details
// the reference leaves to its libraries and that are therefore fixed by convention in this restatement:
//   * std::partial_sort_copy (Localizer.cpp:789) and std::sort inside pcl::VoxelGrid are not stable; here
//     points of equal time / equal voxel keep their input order (stable sorts);
//   * Eigen's evaluation order of the fixed-size float products is restated coefficient-wise
//     (sum over k = 0..n-1, left to right), no FMA (the reference builds for baseline x86-64);
//   * libm's sinf/cosf/atan2f are what this machine provides.
// The product (fast_limo_b200/csrc/scan_prep.cu) never includes this file.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

namespace orc {

// fast_limo::Point (Common.hpp:100-113): 32 bytes, xyz + pad, intensity at 16, time union at 24.
struct RawPoint {
  float x, y, z, pad;
  float intensity;
  float pad2;
  union {
    uint32_t t;        // OUSTER: ns since start of scan
    float time;        // VELODYNE: s since start of scan
    double timestamp;  // HESAI: absolute s; LIVOX: absolute ns
  };
};
static_assert(sizeof(RawPoint) == 32, "fast_limo::Point is 32 bytes");

struct PrepCfg {
  int32_t crop_active, dist_active, rate_active, fov_active;
  float crop_min[3], crop_max[3];
  double min_dist;         // Config::filters::min_dist (cast to float once, Localizer.cpp:274)
  int32_t rate_value;
  float fov_angle;
  int32_t sensor_type;     // 0 OUSTER, 1 VELODYNE, 2 HESAI, 3 LIVOX (Localizer.cpp:747-777)
  int32_t end_of_sweep;
  int32_t voxel_active;
  float leaf;              // voxel_filter.setLeafSize(leafSize[0] x3) (Localizer.cpp:61)
};

// fast_limo::State, the members State::update / get_RT read (State.hpp).
struct Frame {
  double time;
  float q[4];   // x y z w
  float p[3], v[3], w[3], a[3], bg[3], ba[3], g[3];
};

// ---- filters (Localizer.cpp:262-302) ------------------------------------------------------------
// Returns the indices (into raw) of the points of input_pc, in order.
inline std::vector<uint32_t> prep_filter(const RawPoint* raw, size_t n, const PrepCfg& c) {
  std::vector<uint32_t> kept;
  const float min_dist = static_cast<float>(c.min_dist);
  long idx = 0;   // boost::adaptors::indexed(): position in the cloud AFTER NaN removal and crop
  for (size_t i = 0; i < n; ++i) {
    const RawPoint& p = raw[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;   // removeNaNFromPointCloud :265
    if (c.crop_active) {                                                                // CropBox, setNegative(true) :57,268-271
      const bool outside = (p.x < c.crop_min[0] || p.y < c.crop_min[1] || p.z < c.crop_min[2]) ||
                           (p.x > c.crop_max[0] || p.y > c.crop_max[1] || p.z > c.crop_max[2]);
      if (!outside) continue;
    }
    const long my = idx++;
    bool ok = true;
    if (c.dist_active) {
      const float nrm = std::sqrt(p.x * p.x + (p.y * p.y + p.z * p.z));                  // Vector3f::norm()
      ok = ok && (nrm > min_dist);
    }
    if (c.rate_active) ok = ok && (my % c.rate_value == 0);
    if (c.fov_active) ok = ok && (std::fabs(std::atan2(p.y, p.x)) < c.fov_angle);        // isInRange :866-869
    if (ok) kept.push_back((uint32_t)i);
  }
  return kept;
}

// extract_point_time (Localizer.cpp:747-777) relative to sweep_ref_time.
inline double prep_point_time(const RawPoint& p, const PrepCfg& c, double sweep_ref_time) {
  switch (c.sensor_type) {
    case 0: return c.end_of_sweep ? sweep_ref_time - p.t * 1e-9f : sweep_ref_time + p.t * 1e-9f;
    case 1: return c.end_of_sweep ? sweep_ref_time - p.time : sweep_ref_time + p.time;
    case 2: return p.timestamp;
    default: return p.timestamp * 1e-9f;
  }
}

// Sort by time (Localizer.cpp:744-789), stable.  `order` holds indices into raw.
inline void prep_sort(const RawPoint* raw, std::vector<uint32_t>& order, const PrepCfg& c) {
  auto cmp = [&](uint32_t ia, uint32_t ib) {
    const RawPoint &a = raw[ia], &b = raw[ib];
    switch (c.sensor_type) {
      case 0: return c.end_of_sweep ? a.t > b.t : a.t < b.t;
      case 1: return c.end_of_sweep ? a.time > b.time : a.time < b.time;
      default: return a.timestamp < b.timestamp;
    }
  };
  std::stable_sort(order.begin(), order.end(), cmp);
}

// ---- small float algebra in Eigen's fixed-size evaluation order -----------------------------------
inline void quat_to_R(const float q[4], float R[9]) {   // Eigen::Quaternionf::toRotationMatrix
  const float tx = 2.f * q[0], ty = 2.f * q[1], tz = 2.f * q[2];
  const float twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const float txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const float tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1.f - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.f - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.f - (txx + tyy);
}

// Eigen::Quaternionf(Matrix3f) — QuaternionBase::operator=(MatrixBase), quat_product order x y z w
inline void R_to_quat(const float m[9], float q[4]) {
  float t = m[0] + (m[4] + m[8]);                  // trace() = diagonal().sum(): Eigen's unrolled redux of 3 is c0 + (c1 + c2)
  if (t > 0.f) {
    t = std::sqrt(t + 1.0f);
    q[3] = 0.5f * t;
    t = 0.5f / t;
    q[0] = (m[7] - m[5]) * t;
    q[1] = (m[2] - m[6]) * t;
    q[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0f);
    q[i] = 0.5f * t;
    t = 0.5f / t;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
}

inline void quat_mul(const float a[4], const float b[4], float o[4]) {
  // Eigen quat_product<Architecture::SSE, float> (Geometry/arch/Geometry_SSE.h): the operation order
  // of the vectorised product the reference's x86-64 build uses
  o[0] = (a[0] * b[3] - a[2] * b[1]) + (a[1] * b[2] + a[3] * b[0]);
  o[1] = (a[1] * b[3] - a[0] * b[2]) + (a[2] * b[0] + a[3] * b[1]);
  o[2] = (a[2] * b[3] - a[1] * b[0]) + (a[0] * b[1] + a[3] * b[2]);
  o[3] = (a[3] * b[3] - a[0] * b[0]) + -(a[2] * b[2] + a[1] * b[1]);
}

inline void quat_rotate(const float q[4], const float v[3], float o[3]) {   // QuaternionBase::_transformVector
  float uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  const float c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  for (int i = 0; i < 3; ++i) o[i] = (v[i] + q[3] * uv[i]) + c[i];
}

// State::update(t) (State.cpp:76-119): integrates pose over dt = t - time with the stored IMU sample.
inline void frame_update(Frame& s, double t) {
  const double dt = t - s.time;
  const float w[3] = {s.w[0] - s.bg[0], s.w[1] - s.bg[1], s.w[2] - s.bg[2]};
  const float w_norm = std::sqrt(w[0] * w[0] + (w[1] * w[1] + w[2] * w[2]));
  float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (w_norm > 1.e-7) {
    const float r[3] = {w[0] / w_norm, w[1] / w_norm, w[2] / w_norm};
    const float K[9] = {0.f, -r[2], r[1], r[2], 0.f, -r[0], -r[1], r[0], 0.f};
    const float r_ang = (float)(w_norm * dt);
    const float sn = std::sin(r_ang);
    const float c1 = (float)(1.0 - std::cos(r_ang));     // double scalar, converted to float by Eigen's scalar promotion
    float cK[9], KK[9];
    for (int i = 0; i < 9; ++i) cK[i] = c1 * K[i];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) KK[3 * i + j] = (cK[3 * i] * K[j] + cK[3 * i + 1] * K[3 + j]) + cK[3 * i + 2] * K[6 + j];
    for (int i = 0; i < 9; ++i) R[i] = R[i] + (sn * K[i] + KK[i]);
  }
  const float ab[3] = {s.a[0] - s.ba[0], s.a[1] - s.ba[1], s.a[2] - s.ba[2]};
  float a0[3];
  quat_rotate(s.q, ab, a0);
  for (int i = 0; i < 3; ++i) a0[i] += s.g[i];
  float qu[4], qn[4];
  R_to_quat(R, qu);
  quat_mul(s.q, qu, qn);
  std::memcpy(s.q, qn, sizeof(qn));
  const float dtf = (float)dt;
  for (int i = 0; i < 3; ++i) s.p[i] = s.p[i] + (s.v[i] * dtf + ((0.5f * a0[i]) * dtf) * dtf);
  for (int i = 0; i < 3; ++i) s.v[i] = s.v[i] + a0[i] * dtf;
}

// binary_search_tailored (Algorithms.hpp:25-38)
inline int frame_search(const Frame* f, int n, double t) {
  int low = 0, high = n - 1;
  while (high >= low) {
    const int mid = (low + high) / 2;
    if (f[mid].time > t) high = mid - 1; else low = mid + 1;
  }
  return high < 0 ? 0 : high;
}

inline void mat4_mul(const float A[16], const float B[16], float C[16]) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      C[4 * i + j] = ((A[4 * i] * B[j] + A[4 * i + 1] * B[4 + j]) + A[4 * i + 2] * B[8 + j]) + A[4 * i + 3] * B[12 + j];
}
inline void mat4_vec(const float A[16], const float v[4], float o[4]) {
  for (int i = 0; i < 4; ++i) o[i] = ((A[4 * i] * v[0] + A[4 * i + 1] * v[1]) + A[4 * i + 2] * v[2]) + A[4 * i + 3] * v[3];
}
inline void rt_of(const float q[4], const float p[3], float T[16]) {   // State::get_RT
  float R[9];
  quat_to_R(q, R);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
    T[4 * i + 3] = p[i];
  }
  T[12] = T[13] = T[14] = 0.f;
  T[15] = 1.f;
}
inline void rt_inv_of(const float q[4], const float p[3], float T[16]) {   // State::get_RT_inv
  float R[9];
  quat_to_R(q, R);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * j + i];
    T[4 * i + 3] = (-R[i] * p[0] + -R[3 + i] * p[1]) + -R[6 + i] * p[2];
  }
  T[12] = T[13] = T[14] = 0.f;
  T[15] = 1.f;
}

// The per-point loop of deskewPointCloud (Localizer.cpp:822-843).  order = time-sorted indices into raw;
// out_world / out_xt2: 4 floats per point (xyz + 1).
inline void prep_deskew(const RawPoint* raw, const std::vector<uint32_t>& order, const PrepCfg& c, double sweep_ref_time,
                        double offset, const Frame* frames, int n_frames, const float last_q[4], const float last_p[3],
                        const float T_l2b[16], float* out_world, float* out_xt2) {
  float Tinv[16];
  rt_inv_of(last_q, last_p, Tinv);
  for (size_t k = 0; k < order.size(); ++k) {
    const RawPoint& p = raw[order[k]];
    const double t = prep_point_time(p, c, sweep_ref_time) + offset;
    Frame X0 = frames[frame_search(frames, n_frames, t)];
    frame_update(X0, t);
    float RT[16], T[16];
    rt_of(X0.q, X0.p, RT);
    mat4_mul(RT, T_l2b, T);
    const float v[4] = {p.x, p.y, p.z, 1.f};
    float wv[4], bv[4];
    mat4_vec(T, v, wv);
    mat4_vec(Tinv, wv, bv);
    std::memcpy(out_world + 4 * k, wv, sizeof(wv));
    std::memcpy(out_xt2 + 4 * k, bv, sizeof(bv));
  }
}

// pcl::VoxelGrid<PointT>::applyFilter on a dense cloud, downsample_all_data (xyz centroid only is kept).
// in: n points (4 floats each).  Returns centroids, 4 floats each (w = 1), ascending voxel index.
inline std::vector<float> prep_voxel(const float* in, size_t n, float leaf) {
  std::vector<float> out;
  if (n == 0) return out;
  float mn[3] = {in[0], in[1], in[2]}, mx[3] = {in[0], in[1], in[2]};
  for (size_t i = 1; i < n; ++i)
    for (int a = 0; a < 3; ++a) {
      mn[a] = std::min(mn[a], in[4 * i + a]);
      mx[a] = std::max(mx[a], in[4 * i + a]);
    }
  const float inv = 1.0f / leaf;
  const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1,
                dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)INT32_MAX) {          // "Leaf size is too small": output = input
    out.assign(in, in + 4 * n);
    return out;
  }
  int min_b[3], max_b[3], div_b[3];
  for (int a = 0; a < 3; ++a) {
    min_b[a] = (int)std::floor(mn[a] * inv);
    max_b[a] = (int)std::floor(mx[a] * inv);
    div_b[a] = max_b[a] - min_b[a] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  std::vector<uint32_t> key(n), idx(n);
  for (size_t i = 0; i < n; ++i) {
    int ijk[3];
    for (int a = 0; a < 3; ++a) ijk[a] = (int)(std::floor(in[4 * i + a] * inv) - (float)min_b[a]);
    key[i] = (uint32_t)(ijk[0] * mul[0] + ijk[1] * mul[1] + ijk[2] * mul[2]);
    idx[i] = (uint32_t)i;
  }
  std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
  size_t i = 0;
  while (i < n) {
    size_t j = i;
    float s[3] = {0.f, 0.f, 0.f};
    while (j < n && key[idx[j]] == key[idx[i]]) {
      for (int a = 0; a < 3; ++a) s[a] += in[4 * idx[j] + a];
      ++j;
    }
    const float cnt = (float)(j - i);
    out.push_back(s[0] / cnt);
    out.push_back(s[1] / cnt);
    out.push_back(s[2] / cnt);
    out.push_back(1.f);
    i = j;
  }
  return out;
}

}  // namespace orc
