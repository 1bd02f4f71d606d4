// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// CPU restatement of the per-point part of the registration hot path:
//   body->world transform, exact kNN, 5-point plane fit, plane gates, signed
//   point-to-plane distance, Jacobian row.  Parity status: UNPINNED by the reference
//   (no tests/golden vectors there; Eigen absent here so its own code cannot be built).
//
// Follows (paths relative to /root/reference/include):
//   fast_limo/Objects/State.cpp:38-55      double state -> float State        -> FState
//   fast_limo/Objects/State.cpp:136-172    get_RT / get_RT_inv / get_extr_*   -> FState::prepare()
//   fast_limo/Modules/Mapper.cpp:59-86     Mapper::match                      -> match_scan()
//   fast_limo/Modules/Mapper.cpp:100-114   Mapper::match_plane                -> match_one()
//   fast_limo/Objects/Plane.cpp:23-31,41-48  gates (>=k neighbours, d2_k < MAX_DIST_PLANE)
//   fast_limo/Objects/Plane.cpp:80-105     estimate_plane (colPivHouseholderQr().solve)
//                                          -> colpiv_qr_solve() restating Eigen 3.3's
//                                             ColPivHouseholderQR::computeInPlace/_solve_impl
//   fast_limo/Objects/Plane.cpp:107-114    plane_eval
//   fast_limo/Objects/Match.cpp:23-28      dist = n.p + d
//   fast_limo/Modules/Localizer.cpp:537-577 calculate_H
//
// Arithmetic (float32, no FMA contraction; build with -ffp-contract=off).  Where Eigen's
// evaluation order is fixed by its fixed-size unrollers it is reproduced:
//   * 4x4 * vec4 (vectorised coeff-based product): ((m0*x + m1*y) + m2*z) + m3*w
//   * 3x3 * vec3 and Vector3f::squaredNorm(): a0 + (a1 + a2)
//   * quaternion -> matrix: Eigen's QuaternionBase::toRotationMatrix formula
// Inside the dynamic 5x3 QR Eigen's reduction order depends on SIMD width and buffer
// alignment and cannot be pinned; plain left-to-right sums are used (documented unpinned).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "ioctree.hpp"

namespace orc {

struct MatchCfg {
  int k = 5;                      // NUM_MATCH_POINTS
  long max_pc2match = 10000;      // MAX_NUM_PC2MATCH
  long max_matches = 2000;        // MAX_NUM_MATCHES
  double max_dist_plane = 2.0;    // compared against the k-th SQUARED distance (Plane.cpp:47)
  double plane_threshold = 0.05;  // PLANE_THRESHOLD
  int estimate_extrinsics = 1;
  int num_threads = 1;
};

// x,y,z,w order everywhere.
template <typename T>
inline void quat_to_mat(const T q[4], T R[9]) {
  const T x = q[0], y = q[1], z = q[2], w = q[3];
  const T tx = T(2) * x, ty = T(2) * y, tz = T(2) * z;
  const T twx = tx * w, twy = ty * w, twz = tz * w;
  const T txx = tx * x, txy = ty * x, txz = tz * x;
  const T tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = T(1) - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = T(1) - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = T(1) - (txx + tyy);
}

// Row-major 3x3 times vec3 with Eigen's fixed-3 reduction order.
inline void mat3_vec(const float M[9], const float v[3], float out[3]) {
  for (int r = 0; r < 3; ++r) {
    const float a = M[3 * r] * v[0], b = M[3 * r + 1] * v[1], c = M[3 * r + 2] * v[2];
    out[r] = a + (b + c);
  }
}

// Affine (R|t) applied the way Eigen evaluates Matrix4f * Vector4f(x,y,z,1).
inline void affine_apply(const float R[9], const float t[3], const float p[3], float out[3]) {
  for (int r = 0; r < 3; ++r) {
    float acc = R[3 * r] * p[0];
    acc = R[3 * r + 1] * p[1] + acc;
    acc = R[3 * r + 2] * p[2] + acc;
    out[r] = t[r] * 1.0f + acc;
  }
}

// Float view of the filter state plus the constant matrices the hot loops use.
struct FState {
  float q[4], p[3], qLI[4], pLI[3];          // State.cpp:38-55 (cast of the double state)
  float R_wb[9], t_wb[3];                    // get_RT
  float Rinv_wb[9], tinv_wb[3];              // get_RT_inv
  float Rinv_LI[9], tinv_LI[3];              // get_extr_RT_inv
  float Rd_wb_inv[9], Rd_LI_inv[9];          // Localizer.cpp:554-555: double quat -> matrix -> cast

  // state = pos[3], rot[4], offR[4], offT[3]   (doubles)
  void prepare(const double pos[3], const double rot[4], const double offR[4], const double offT[3]) {
    for (int i = 0; i < 4; ++i) {
      q[i] = static_cast<float>(rot[i]);
      qLI[i] = static_cast<float>(offR[i]);
    }
    for (int i = 0; i < 3; ++i) {
      p[i] = static_cast<float>(pos[i]);
      pLI[i] = static_cast<float>(offT[i]);
    }
    quat_to_mat<float>(q, R_wb);
    for (int i = 0; i < 3; ++i) t_wb[i] = p[i];
    inverse_of(R_wb, p, Rinv_wb, tinv_wb);
    float R_LI[9];
    quat_to_mat<float>(qLI, R_LI);
    inverse_of(R_LI, pLI, Rinv_LI, tinv_LI);
    double qc[4] = {-rot[0], -rot[1], -rot[2], rot[3]}, Rd[9];
    quat_to_mat<double>(qc, Rd);
    for (int i = 0; i < 9; ++i) Rd_wb_inv[i] = static_cast<float>(Rd[i]);
    double qc2[4] = {-offR[0], -offR[1], -offR[2], offR[3]};
    quat_to_mat<double>(qc2, Rd);
    for (int i = 0; i < 9; ++i) Rd_LI_inv[i] = static_cast<float>(Rd[i]);
  }

 private:
  // Tinv = [R^T | -R^T p]  (State.cpp:145-153)
  static void inverse_of(const float R[9], const float p[3], float Rt[9], float tinv[3]) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Rt[3 * r + c] = R[3 * c + r];
    float neg[9];
    for (int i = 0; i < 9; ++i) neg[i] = -Rt[i];
    mat3_vec(neg, p, tinv);
  }
};

// Solve min ||A x - b|| for a rows x 3 float system the way Eigen's
// ColPivHouseholderQR does (A is row-major rows x 3, destroyed).  rows <= 16.
inline void colpiv_qr_solve(float* A, int rows, const float* b_in, float x[3]) {
  const int cols = 3;
  const int size = rows < cols ? rows : cols;
  auto at = [&](int r, int c) -> float& { return A[r * cols + c]; };
  float hcoef[3] = {0, 0, 0};
  int perm[3] = {0, 1, 2};
  float norm_upd[3], norm_dir[3];
  for (int c = 0; c < cols; ++c) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s = s + at(r, c) * at(r, c);
    norm_dir[c] = std::sqrt(s);
    norm_upd[c] = norm_dir[c];
  }
  const float eps = 1.1920929e-07f;  // FLT_EPSILON
  float maxn = norm_upd[0];
  for (int c = 1; c < cols; ++c)
    if (norm_upd[c] > maxn) maxn = norm_upd[c];
  const float th = maxn * eps;
  const float threshold_helper = (th * th) / static_cast<float>(rows);
  const float downdate_thr = std::sqrt(eps);
  int nonzero_pivots = size;

  for (int k = 0; k < size; ++k) {
    int big = k;
    float bign = norm_upd[k];
    for (int c = k + 1; c < cols; ++c)
      if (norm_upd[c] > bign) {
        bign = norm_upd[c];
        big = c;
      }
    const float big_sq = bign * bign;
    if (nonzero_pivots == size && big_sq < threshold_helper * static_cast<float>(rows - k)) nonzero_pivots = k;
    if (big != k) {
      for (int r = 0; r < rows; ++r) {
        float tmp = at(r, k);
        at(r, k) = at(r, big);
        at(r, big) = tmp;
      }
      float t1 = norm_upd[k]; norm_upd[k] = norm_upd[big]; norm_upd[big] = t1;
      float t2 = norm_dir[k]; norm_dir[k] = norm_dir[big]; norm_dir[big] = t2;
      int t3 = perm[k]; perm[k] = perm[big]; perm[big] = t3;
    }
    // Householder reflector of column k below (and including) the diagonal.
    float tail_sq = 0.f;
    for (int r = k + 1; r < rows; ++r) tail_sq = tail_sq + at(r, k) * at(r, k);
    const float c0 = at(k, k);
    float tau, beta;
    if (tail_sq <= FLT_MIN) {
      tau = 0.f;
      beta = c0;
      for (int r = k + 1; r < rows; ++r) at(r, k) = 0.f;
    } else {
      beta = std::sqrt(c0 * c0 + tail_sq);
      if (c0 >= 0.f) beta = -beta;
      const float den = c0 - beta;
      for (int r = k + 1; r < rows; ++r) at(r, k) = at(r, k) / den;
      tau = (beta - c0) / beta;
    }
    at(k, k) = beta;
    hcoef[k] = tau;
    // Apply H_k to the trailing columns.
    if (tau != 0.f) {
      for (int c = k + 1; c < cols; ++c) {
        float tmp = 0.f;
        for (int r = k + 1; r < rows; ++r) tmp = tmp + at(r, k) * at(r, c);
        tmp = tmp + at(k, c);
        at(k, c) = at(k, c) - tau * tmp;
        for (int r = k + 1; r < rows; ++r) at(r, c) = at(r, c) - tmp * (tau * at(r, k));
      }
    }
    // LAPACK-style column-norm downdate (LAWN 176).
    for (int c = k + 1; c < cols; ++c) {
      if (norm_upd[c] != 0.f) {
        float temp = std::fabs(at(k, c)) / norm_upd[c];
        temp = (1.f + temp) * (1.f - temp);
        temp = temp < 0.f ? 0.f : temp;
        const float ratio = norm_upd[c] / norm_dir[c];
        const float temp2 = temp * (ratio * ratio);
        if (temp2 <= downdate_thr) {
          float s = 0.f;
          for (int r = k + 1; r < rows; ++r) s = s + at(r, c) * at(r, c);
          norm_dir[c] = std::sqrt(s);
          norm_upd[c] = norm_dir[c];
        } else {
          norm_upd[c] = norm_upd[c] * std::sqrt(temp);
        }
      }
    }
  }

  // c = Q^T b, then back-substitute R x = c on the leading nonzero_pivots block.
  float c[16];
  for (int r = 0; r < rows; ++r) c[r] = b_in[r];
  x[0] = x[1] = x[2] = 0.f;
  if (nonzero_pivots == 0) return;
  for (int k = 0; k < nonzero_pivots; ++k) {
    const float tau = hcoef[k];
    if (tau == 0.f) continue;
    float tmp = 0.f;
    for (int r = k + 1; r < rows; ++r) tmp = tmp + at(r, k) * c[r];
    tmp = tmp + c[k];
    c[k] = c[k] - tau * tmp;
    for (int r = k + 1; r < rows; ++r) c[r] = c[r] - tmp * (tau * at(r, k));
  }
  for (int i = nonzero_pivots - 1; i >= 0; --i) {
    c[i] = c[i] / at(i, i);
    for (int j = 0; j < i; ++j) c[j] = c[j] - c[i] * at(j, i);
  }
  for (int i = 0; i < nonzero_pivots; ++i) x[perm[i]] = c[i];
}

struct PointMatch {
  float g[3];        // world point
  float l[3];        // body point
  float n[4];        // plane (A,B,C,D), meaningful only when the fit ran
  float dist;        // signed point-to-plane distance
  float nn_d2[8];    // sorted neighbour squared distances (debug / parity)
  int nn_cnt;
  bool good;         // plane accepted (Match::lisanAlGaib)
};

// Mapper::match_plane + Plane ctor + Match ctor for one point.
inline PointMatch match_one(const IOctree& map, const MatchCfg& cfg, const float g[3], const float l[3]) {
  PointMatch m;
  std::memset(&m, 0, sizeof(m));
  for (int i = 0; i < 3; ++i) {
    m.g[i] = g[i];
    m.l[i] = l[i];
  }
  std::vector<float> d2(cfg.k);   // heap-allocated per query as in the reference (Mapper.cpp:103-111)
  std::vector<P3> nb(cfg.k);
  const int got = map.knn(P3{g[0], g[1], g[2]}, cfg.k, nb.data(), d2.data());
  m.nn_cnt = got;
  for (int i = 0; i < got && i < 8; ++i) m.nn_d2[i] = d2[i];
  m.good = false;
  // Plane.cpp:23-31: default-constructed n_ABCD is uninitialised in the reference when a
  // gate fails; dist is then garbage but never used (match is not chosen).  We emit zeros.
  if (got < cfg.k) return m;                                                   // enough_points
  if (!(static_cast<double>(d2[got - 1]) < cfg.max_dist_plane)) return m;       // close_enough
  std::vector<float> A(static_cast<size_t>(got) * 3), b(got, -1.0f);             // dynamic matrices (Plane.cpp:82-83)
  for (int j = 0; j < got; ++j) {
    A[3 * j] = nb[j].x;
    A[3 * j + 1] = nb[j].y;
    A[3 * j + 2] = nb[j].z;
  }
  float x[3];
  colpiv_qr_solve(A.data(), got, b.data(), x);
  const float nrm = std::sqrt(x[0] * x[0] + (x[1] * x[1] + x[2] * x[2]));
  m.n[0] = x[0] / nrm;
  m.n[1] = x[1] / nrm;
  m.n[2] = x[2] / nrm;
  m.n[3] = static_cast<float>(1.0 / static_cast<double>(nrm));
  const float thr = static_cast<float>(cfg.plane_threshold);
  bool ok = true;
  for (int j = 0; j < got; ++j) {
    const float res = ((m.n[0] * nb[j].x + m.n[1] * nb[j].y) + m.n[2] * nb[j].z) + m.n[3];
    if (std::fabs(res) > thr) {
      ok = false;
      break;
    }
  }
  m.good = ok;
  m.dist = ((m.n[0] * g[0] + m.n[1] * g[1]) + m.n[2] * g[2]) + m.n[3];
  return m;
}

// Mapper::match: first min(N, MAX_NUM_PC2MATCH) points, OpenMP over points, then an
// order-preserving serial compaction of the accepted matches.
inline void match_scan(const IOctree& map, const MatchCfg& cfg, const FState& s, const float* scan, size_t n,
                       size_t stride, std::vector<PointMatch>& all, std::vector<PointMatch>& chosen) {
  all.clear();
  chosen.clear();
  if (map.size() == 0) return;   // Mapper::exists() (Mapper.cpp:47-49,61)
  const long cap = cfg.max_pc2match;
  const long n_q = (static_cast<long>(n) > cap) ? cap : static_cast<long>(n);
  all.resize(n_q);
#pragma omp parallel for num_threads(cfg.num_threads)
  for (long i = 0; i < n_q; ++i) {
    const float* p = scan + static_cast<size_t>(i) * stride;
    float g[3];
    affine_apply(s.R_wb, s.t_wb, p, g);
    all[i] = match_one(map, cfg, g, p);
  }
  for (long j = 0; j < n_q; ++j)
    if (all[j].good) chosen.push_back(all[j]);
}

// Localizer::calculate_H: rows for the first min(N_v, MAX_NUM_MATCHES) chosen matches.
// H is row-major n_rows x 12 (double), h = -dist.
inline long jacobian_rows(const MatchCfg& cfg, const FState& s, const std::vector<PointMatch>& chosen, double* H,
                          double* h) {
  const long nv = static_cast<long>(chosen.size());
  const long n_rows = nv > cfg.max_matches ? cfg.max_matches : nv;
#pragma omp parallel for num_threads(cfg.num_threads)
  for (long i = 0; i < n_rows; ++i) {
    const PointMatch m = chosen[i];
    float p_imu[3], p_lid[3];
    affine_apply(s.Rinv_wb, s.tinv_wb, m.g, p_imu);
    affine_apply(s.Rinv_LI, s.tinv_LI, p_imu, p_lid);
    const float nrm[3] = {m.n[0], m.n[1], m.n[2]};
    float C[3], RC[3];
    mat3_vec(s.Rd_wb_inv, nrm, C);
    mat3_vec(s.Rd_LI_inv, C, RC);
    const float B[3] = {p_lid[1] * RC[2] - p_lid[2] * RC[1], p_lid[2] * RC[0] - p_lid[0] * RC[2],
                        p_lid[0] * RC[1] - p_lid[1] * RC[0]};
    const float Av[3] = {p_imu[1] * C[2] - p_imu[2] * C[1], p_imu[2] * C[0] - p_imu[0] * C[2],
                         p_imu[0] * C[1] - p_imu[1] * C[0]};
    double* row = H + 12 * i;
    for (int c = 0; c < 12; ++c) row[c] = 0.0;
    row[0] = nrm[0]; row[1] = nrm[1]; row[2] = nrm[2];
    row[3] = Av[0];  row[4] = Av[1];  row[5] = Av[2];
    if (cfg.estimate_extrinsics) {
      row[6] = B[0]; row[7] = B[1]; row[8] = B[2];
      row[9] = C[0]; row[10] = C[1]; row[11] = C[2];
    }
    h[i] = -static_cast<double>(m.dist);
  }
  return n_rows;
}

// esekfom.hpp:1723,1727 — HTH = H^T H (12x12, row-major) and H^T h (12), plain loops.
inline void normal_equations(const double* H, const double* h, long n, double HTH[144], double HTh[12]) {
  for (int i = 0; i < 144; ++i) HTH[i] = 0.0;
  for (int i = 0; i < 12; ++i) HTh[i] = 0.0;
  for (long r = 0; r < n; ++r) {
    const double* row = H + 12 * r;
    for (int i = 0; i < 12; ++i) {
      for (int j = 0; j < 12; ++j) HTH[12 * i + j] += row[i] * row[j];
      HTh[i] += row[i] * h[r];
    }
  }
}

}  // namespace orc
