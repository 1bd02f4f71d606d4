// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// CPU restatement of the IKFoM iterated error-state Kalman update that consumes the
// registration Jacobian.  All float64.  Parity status: UNPINNED by the reference (no
// tests/golden vectors; Eigen/Boost absent so the original cannot be compiled here).
// Eigen internals that are not reproducible (blocked LU order, GEMM order,
// EigenSolver's eigenvector basis) are replaced by plain algorithms; differences are
// O(1e-15) relative except in the degenerate-geometry branch, which is basis dependent
// in the reference and documented as unpinned.
//
// Follows (paths relative to /root/reference/include/IKFoM):
//   use-ikfom.hpp:6-21                       state layout: pos rot offR offT vel bg ba grav(S2)
//   IKFoM_toolkit/esekfom/esekfom.hpp:1620-1823  update_iterated_dyn_share_modified -> iterated_update()
//   IKFoM_toolkit/mtk/src/mtkmath.hpp:143-175    cos_sinc_sqrt
//   IKFoM_toolkit/mtk/src/mtkmath.hpp:236-247    A_matrix
//   IKFoM_toolkit/mtk/src/mtkmath.hpp:250-257    exp
//   IKFoM_toolkit/mtk/src/mtkmath.hpp:269-289    log
//   IKFoM_toolkit/mtk/types/SOn.hpp:233-239,284-297  SO3 boxplus/boxminus/exp/log
//   IKFoM_toolkit/mtk/types/S2.hpp:136-167       S2 boxplus/boxminus
//   IKFoM_toolkit/mtk/types/S2.hpp:179-232       S2_Bx (chart type 1 branch :216-231)
//   IKFoM_toolkit/mtk/types/S2.hpp:259-280       S2_Nx_yy, S2_Mx  (scalar(1/2) == 0 at :277)
//   IKFoM_toolkit/mtk/types/vect.hpp:117-122     vect boxplus/boxminus
//   IKFoM_toolkit/mtk/build_manifold.hpp:192-200 compound boxplus/boxminus in declaration order
#pragma once
#include <cmath>
#include <cstring>
#include <functional>
#include <vector>

namespace orc {

constexpr int NDOF = 23;
constexpr double S2_LEN = 98090.0 / 10000.0;  // MTK::S2<double, 98090, 10000, 1>
constexpr double MTK_TOL = 1e-11;

// quaternions are x,y,z,w
struct EkfState {
  double pos[3], rot[4], offR[4], offT[3], vel[3], bg[3], ba[3], grav[3];
};
static_assert(sizeof(EkfState) == 26 * sizeof(double), "flat layout");

// ---------- small dense helpers (row-major) ----------
inline void matmul(const double* A, const double* B, double* C, int n, int k, int m) {  // C[n x m] = A[n x k] B[k x m]
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) {
      double s = 0;
      for (int t = 0; t < k; ++t) s += A[i * k + t] * B[t * m + j];
      C[i * m + j] = s;
    }
}

// In-place inverse by LU with partial pivoting (Gauss-Jordan on [A|I]).  Returns false if singular.
inline bool invert(double* A, int n) {
  std::vector<double> M(static_cast<size_t>(n) * 2 * n, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) M[i * 2 * n + j] = A[i * n + j];
    M[i * 2 * n + n + i] = 1.0;
  }
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = std::fabs(M[c * 2 * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(M[r * 2 * n + c]) > best) {
        best = std::fabs(M[r * 2 * n + c]);
        piv = r;
      }
    if (best == 0.0) return false;
    if (piv != c)
      for (int j = 0; j < 2 * n; ++j) std::swap(M[c * 2 * n + j], M[piv * 2 * n + j]);
    const double d = M[c * 2 * n + c];
    for (int j = 0; j < 2 * n; ++j) M[c * 2 * n + j] /= d;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = M[r * 2 * n + c];
      if (f == 0.0) continue;
      for (int j = 0; j < 2 * n; ++j) M[r * 2 * n + j] -= f * M[c * 2 * n + j];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) A[i * n + j] = M[i * 2 * n + n + j];
  return true;
}

// Cyclic Jacobi eigen-decomposition of a symmetric n x n matrix: A = V diag(w) V^T (columns of V).
inline void jacobi_eig(const double* Ain, int n, double* w, double* V) {
  std::vector<double> A(Ain, Ain + n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        const double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq;
          V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
}

// ---------- manifold pieces ----------
inline void hat(const double v[3], double M[9]) {
  M[0] = 0;     M[1] = -v[2]; M[2] = v[1];
  M[3] = v[2];  M[4] = 0;     M[5] = -v[0];
  M[6] = -v[1]; M[7] = v[0];  M[8] = 0;
}

inline void cos_sinc_sqrt(double x2, double& c, double& sinc) {
  const double eps = 2.220446049250313e-16;
  const double taylor_n_bound = std::sqrt(std::sqrt(eps));
  if (x2 >= taylor_n_bound) {
    const double x = std::sqrt(x2);
    c = std::cos(x);
    sinc = std::sin(x) / x;
    return;
  }
  static const double inv[] = {1 / 3., 1 / 4., 1 / 5., 1 / 6., 1 / 7., 1 / 8., 1 / 9.};
  double cosi = 1., si = 1.;
  double term = -1 / 2. * x2;
  for (int i = 0; i < 3; ++i) {
    cosi += term;
    term *= inv[2 * i];
    si += term;
    term *= -inv[2 * i + 1] * x2;
  }
  c = cosi;
  sinc = si;
}

// MTK::exp: returns w, writes vector part.
inline double mtk_exp(double out_vec[3], const double v[3], double scale) {
  const double n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  double c, sinc;
  cos_sinc_sqrt(scale * scale * n2, c, sinc);
  const double mult = sinc * scale;
  for (int i = 0; i < 3; ++i) out_vec[i] = mult * v[i];
  return c;
}

inline void quat_mul(const double a[4], const double b[4], double r[4]) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
  r[3] = aw * bw - ax * bx - ay * by - az * bz;
  r[0] = aw * bx + ax * bw + ay * bz - az * by;
  r[1] = aw * by + ay * bw + az * bx - ax * bz;
  r[2] = aw * bz + az * bw + ax * by - ay * bx;
}

inline void quat_R(const double q[4], double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

inline void so3_boxplus(double q[4], const double d[3]) {
  double e[4];
  e[3] = mtk_exp(e, d, 0.5);
  double r[4];
  quat_mul(q, e, r);
  std::memcpy(q, r, sizeof(r));
}

// res = log(other^-1 * self)
inline void so3_boxminus(const double self[4], const double other[4], double res[3]) {
  const double oc[4] = {-other[0], -other[1], -other[2], other[3]};
  double r[4];
  quat_mul(oc, self, r);
  double nv = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (nv < MTK_TOL) nv = MTK_TOL;           // plus_minus_periodicity == true path
  const double s = 2.0 / nv * std::atan(nv / r[3]);
  for (int i = 0; i < 3; ++i) res[i] = s * r[i];
}

inline void A_matrix(const double v[3], double A[9]) {
  const double sq = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  const double n = std::sqrt(sq);
  for (int i = 0; i < 9; ++i) A[i] = (i % 4 == 0);
  if (n < MTK_TOL) return;
  double K[9], K2[9];
  hat(v, K);
  matmul(K, K, K2, 3, 3, 3);
  const double a = (1 - std::cos(n)) / sq, b = (1 - std::sin(n) / n) / sq;
  for (int i = 0; i < 9; ++i) A[i] += a * K[i] + b * K2[i];
}

// S2 chart type 1 (S2.hpp:216-231).  Bx is 3x2 row-major.
inline void s2_Bx(const double v[3], double Bx[6]) {
  const double L = S2_LEN;
  if (v[0] + L > MTK_TOL) {
    const double d = L + v[0];
    Bx[0] = -v[1];                 Bx[1] = -v[2];
    Bx[2] = L - v[1] * v[1] / d;   Bx[3] = -v[2] * v[1] / d;
    Bx[4] = -v[2] * v[1] / d;      Bx[5] = L - v[2] * v[2] / d;
    for (int i = 0; i < 6; ++i) Bx[i] /= L;
  } else {
    for (int i = 0; i < 6; ++i) Bx[i] = 0;
    Bx[3] = -1;  // res(1,1)
    Bx[4] = 1;   // res(2,0)
  }
}

inline void s2_boxplus(double v[3], const double d[2]) {
  double Bx[6];
  s2_Bx(v, Bx);
  const double Bu[3] = {Bx[0] * d[0] + Bx[1] * d[1], Bx[2] * d[0] + Bx[3] * d[1], Bx[4] * d[0] + Bx[5] * d[1]};
  double e[4];
  e[3] = mtk_exp(e, Bu, 0.5);
  double R[9];
  quat_R(e, R);
  const double r[3] = {R[0] * v[0] + R[1] * v[1] + R[2] * v[2], R[3] * v[0] + R[4] * v[1] + R[5] * v[2],
                       R[6] * v[0] + R[7] * v[1] + R[8] * v[2]};
  v[0] = r[0]; v[1] = r[1]; v[2] = r[2];
}

inline void cross3(const double a[3], const double b[3], double r[3]) {
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}

inline void s2_boxminus(const double self[3], const double other[3], double res[2]) {
  double cr[3];
  cross3(self, other, cr);
  const double v_sin = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
  const double v_cos = self[0] * other[0] + self[1] * other[1] + self[2] * other[2];
  const double theta = std::atan2(v_sin, v_cos);
  if (v_sin < MTK_TOL) {
    if (std::fabs(theta) > MTK_TOL) {
      res[0] = 3.1415926;
      res[1] = 0;
    } else {
      res[0] = res[1] = 0;
    }
    return;
  }
  double Bx[6];
  s2_Bx(other, Bx);
  double oc[3];
  cross3(other, self, oc);  // hat(other) * self
  const double f = theta / v_sin;
  res[0] = f * (Bx[0] * oc[0] + Bx[2] * oc[1] + Bx[4] * oc[2]);
  res[1] = f * (Bx[1] * oc[0] + Bx[3] * oc[1] + Bx[5] * oc[2]);
}

// Nx (2x3) = 1/len^2 * Bx^T * hat(v)
inline void s2_Nx_yy(const double v[3], double Nx[6]) {
  double Bx[6], H[9];
  s2_Bx(v, Bx);
  hat(v, H);
  const double f = 1 / S2_LEN / S2_LEN;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += Bx[k * 2 + i] * H[k * 3 + j];
      Nx[i * 3 + j] = f * s;
    }
}

// Mx (3x2): -hat(v) Bx when |delta| small; else -exp0 * hat(v) * A(Bu)^T * Bx with exp0 == I
inline void s2_Mx(const double v[3], const double delta[2], double Mx[6]) {
  double Bx[6], H[9];
  s2_Bx(v, Bx);
  hat(v, H);
  const double dn = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1]);
  if (dn < MTK_TOL) {
    double T[6];
    matmul(H, Bx, T, 3, 3, 2);
    for (int i = 0; i < 6; ++i) Mx[i] = -T[i];
    return;
  }
  const double Bu[3] = {Bx[0] * delta[0] + Bx[1] * delta[1], Bx[2] * delta[0] + Bx[3] * delta[1],
                        Bx[4] * delta[0] + Bx[5] * delta[1]};
  double e[4];
  e[3] = mtk_exp(e, Bu, 0.0);  // scalar(1/2) is integer division in the reference => 0
  double E[9], A[9], At[9], T1[9], T2[9], T3[6];
  quat_R(e, E);
  A_matrix(Bu, A);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) At[i * 3 + j] = A[j * 3 + i];
  for (int i = 0; i < 9; ++i) E[i] = -E[i];
  matmul(E, H, T1, 3, 3, 3);
  matmul(T1, At, T2, 3, 3, 3);
  matmul(T2, Bx, T3, 3, 3, 2);
  for (int i = 0; i < 6; ++i) Mx[i] = T3[i];
}

inline void state_boxplus(EkfState& x, const double d[NDOF]) {
  for (int i = 0; i < 3; ++i) x.pos[i] += d[i];
  so3_boxplus(x.rot, d + 3);
  so3_boxplus(x.offR, d + 6);
  for (int i = 0; i < 3; ++i) x.offT[i] += d[9 + i];
  for (int i = 0; i < 3; ++i) x.vel[i] += d[12 + i];
  for (int i = 0; i < 3; ++i) x.bg[i] += d[15 + i];
  for (int i = 0; i < 3; ++i) x.ba[i] += d[18 + i];
  s2_boxplus(x.grav, d + 21);
}

inline void state_boxminus(const EkfState& x, const EkfState& o, double d[NDOF]) {
  for (int i = 0; i < 3; ++i) d[i] = x.pos[i] - o.pos[i];
  so3_boxminus(x.rot, o.rot, d + 3);
  so3_boxminus(x.offR, o.offR, d + 6);
  for (int i = 0; i < 3; ++i) d[9 + i] = x.offT[i] - o.offT[i];
  for (int i = 0; i < 3; ++i) d[12 + i] = x.vel[i] - o.vel[i];
  for (int i = 0; i < 3; ++i) d[15 + i] = x.bg[i] - o.bg[i];
  for (int i = 0; i < 3; ++i) d[18 + i] = x.ba[i] - o.ba[i];
  s2_boxminus(x.grav, o.grav, d + 21);
}

// rows [idx, idx+b) <- J * rows ; then cols [idx, idx+b) <- cols * J^T, on an n x n matrix.
inline void project_rows(double* P, int n, int idx, int b, const double* J) {
  for (int c = 0; c < n; ++c) {
    double t[3];
    for (int i = 0; i < b; ++i) {
      double s = 0;
      for (int k = 0; k < b; ++k) s += J[i * b + k] * P[(idx + k) * n + c];
      t[i] = s;
    }
    for (int i = 0; i < b; ++i) P[(idx + i) * n + c] = t[i];
  }
}
inline void project_cols(double* P, int n, int idx, int b, const double* J) {
  for (int r = 0; r < n; ++r) {
    double t[3];
    for (int i = 0; i < b; ++i) {
      double s = 0;
      for (int k = 0; k < b; ++k) s += P[r * n + idx + k] * J[i * b + k];
      t[i] = s;
    }
    for (int i = 0; i < b; ++i) P[r * n + idx + i] = t[i];
  }
}

// Measurement model: fills H (N x 12 row-major) and h (N), returns N.
using MeasModel = std::function<long(const EkfState&, std::vector<double>&, std::vector<double>&)>;

struct PassTrace {
  EkfState x_after;
  double dx[NDOF];      // dx_ (before the degeneracy filter)
  long n_rows;
  double HTH[144], HTh[12];
};

// esekf::update_iterated_dyn_share_modified.  P is 23x23 row-major, updated in place.
// Returns the number of passes executed.
inline int iterated_update(EkfState& x, double* P, int maximum_iter, const double limit[NDOF], double Rn, double D,
                           const MeasModel& model, std::vector<PassTrace>* trace = nullptr) {
  const int n = NDOF;
  int t = 0, passes = 0;
  const EkfState x_prop = x;
  std::vector<double> P_prop(P, P + n * n);
  std::vector<double> H, h;
  double K_h[NDOF], K_x[NDOF * NDOF], dx_new[NDOF];
  std::memset(dx_new, 0, sizeof(dx_new));

  for (int i = -1; i < maximum_iter; ++i) {
    const long N = model(x, H, h);
    ++passes;
    double dx[NDOF];
    state_boxminus(x, x_prop, dx);
    std::memcpy(dx_new, dx, sizeof(dx));
    std::memcpy(P, P_prop.data(), sizeof(double) * n * n);

    for (int idx : {3, 6}) {
      double A[9], J[9];
      A_matrix(dx + idx, A);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) J[r * 3 + c] = A[c * 3 + r];
      double tmp[3];
      for (int r = 0; r < 3; ++r) tmp[r] = J[r * 3] * dx_new[idx] + J[r * 3 + 1] * dx_new[idx + 1] + J[r * 3 + 2] * dx_new[idx + 2];
      for (int r = 0; r < 3; ++r) dx_new[idx + r] = tmp[r];
      project_rows(P, n, idx, 3, J);
      project_cols(P, n, idx, 3, J);
    }
    {
      const int idx = 21;
      double Nx[6], Mx[6], J2[4];
      s2_Nx_yy(x.grav, Nx);
      s2_Mx(x_prop.grav, dx + idx, Mx);
      matmul(Nx, Mx, J2, 2, 3, 2);
      const double a = J2[0] * dx_new[idx] + J2[1] * dx_new[idx + 1], b = J2[2] * dx_new[idx] + J2[3] * dx_new[idx + 1];
      dx_new[idx] = a;
      dx_new[idx + 1] = b;
      project_rows(P, n, idx, 2, J2);
      project_cols(P, n, idx, 2, J2);
    }

    double HTH[144], HTh[12];
    for (int q = 0; q < 144; ++q) HTH[q] = 0;   // reference leaves HTH uninitialised when N < 23; restated as 0
    for (int q = 0; q < 12; ++q) HTh[q] = 0;
    if (n > N) {
      // K = P Hc^T (Hc P Hc^T / R + I)^-1 / R
      std::vector<double> Hc(static_cast<size_t>(N) * n, 0.0), PHt(static_cast<size_t>(n) * N), S(static_cast<size_t>(N) * N);
      for (long r = 0; r < N; ++r)
        for (int c = 0; c < 12; ++c) Hc[r * n + c] = H[r * 12 + c];
      for (int r = 0; r < n; ++r)
        for (long c = 0; c < N; ++c) {
          double s = 0;
          for (int k = 0; k < n; ++k) s += P[r * n + k] * Hc[c * n + k];
          PHt[r * N + c] = s;
        }
      for (long r = 0; r < N; ++r)
        for (long c = 0; c < N; ++c) {
          double s = 0;
          for (int k = 0; k < n; ++k) s += Hc[r * n + k] * PHt[k * N + c];
          S[r * N + c] = s / Rn + (r == c ? 1.0 : 0.0);
        }
      if (N > 0) invert(S.data(), static_cast<int>(N));
      std::vector<double> K(static_cast<size_t>(n) * N);
      for (int r = 0; r < n; ++r)
        for (long c = 0; c < N; ++c) {
          double s = 0;
          for (long k = 0; k < N; ++k) s += PHt[r * N + k] * S[k * N + c];
          K[r * N + c] = s / Rn;
        }
      for (int r = 0; r < n; ++r) {
        double s = 0;
        for (long k = 0; k < N; ++k) s += K[r * N + k] * h[k];
        K_h[r] = s;
        for (int c = 0; c < n; ++c) {
          double s2 = 0;
          for (long k = 0; k < N; ++k) s2 += K[r * N + k] * Hc[k * n + c];
          K_x[r * n + c] = s2;
        }
      }
    } else {
      std::vector<double> Pt(P, P + n * n);
      for (auto& v : Pt) v /= Rn;
      invert(Pt.data(), n);
      for (long r = 0; r < N; ++r) {
        const double* row = &H[r * 12];
        for (int a = 0; a < 12; ++a) {
          for (int b = 0; b < 12; ++b) HTH[a * 12 + b] += row[a] * row[b];
          HTh[a] += row[a] * h[r];
        }
      }
      for (int a = 0; a < 12; ++a)
        for (int b = 0; b < 12; ++b) Pt[a * n + b] += HTH[a * 12 + b];
      invert(Pt.data(), n);  // P_inv
      for (int r = 0; r < n; ++r) {
        double s = 0;
        for (int k = 0; k < 12; ++k) s += Pt[r * n + k] * HTh[k];
        K_h[r] = s;
        for (int c = 0; c < n; ++c) K_x[r * n + c] = 0;
        for (int c = 0; c < 12; ++c) {
          double s2 = 0;
          for (int k = 0; k < 12; ++k) s2 += Pt[r * n + k] * HTH[k * 12 + c];
          K_x[r * n + c] = s2;
        }
      }
    }

    double dx_[NDOF];
    for (int r = 0; r < n; ++r) {
      double s = K_h[r];
      for (int c = 0; c < n; ++c) s += (K_x[r * n + c] - (r == c ? 1.0 : 0.0)) * dx_new[c];
      dx_[r] = s;
    }

    // Degeneracy filter on the pose block (esekfom.hpp:1736-1744).
    double dxn[NDOF];
    std::memcpy(dxn, dx_, sizeof(dx_));
    {
      double B[36], w[6], V[36];
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) B[r * 6 + c] = HTH[r * 12 + c];
      jacobi_eig(B, 6, w, V);
      double prod = 1;
      for (int k = 0; k < 6; ++k) prod *= w[k];
      if (prod < 1e-20)
        for (int r = 0; r < 6; ++r)
          for (int c = 0; c < 6; ++c) V[r * 6 + c] = (r == c);
      double Ssel[36];
      std::memcpy(Ssel, V, sizeof(V));
      for (int k = 0; k < 6; ++k)
        if (w[k] < D)
          for (int c = 0; c < 6; ++c) Ssel[k * 6 + c] = 0;   // zeroes ROW k, as the reference does
      double Vi[36];
      std::memcpy(Vi, V, sizeof(V));
      invert(Vi, 6);
      double T[36];
      matmul(Vi, Ssel, T, 6, 6, 6);
      for (int r = 0; r < 6; ++r) {
        double s = 0;
        for (int c = 0; c < 6; ++c) s += T[r * 6 + c] * dx_[c];
        dxn[r] = s;
      }
    }

    state_boxplus(x, dxn);
    bool converge = true;
    for (int r = 0; r < n; ++r)
      if (std::fabs(dx_[r]) > limit[r]) {
        converge = false;
        break;
      }
    if (converge) ++t;

    if (trace) {
      PassTrace tr;
      tr.x_after = x;
      std::memcpy(tr.dx, dx_, sizeof(dx_));
      tr.n_rows = N;
      std::memcpy(tr.HTH, HTH, sizeof(HTH));
      std::memcpy(tr.HTh, HTh, sizeof(HTh));
      trace->push_back(tr);
    }

    if (t > 1 || i == maximum_iter - 1) {
      std::vector<double> L(P, P + n * n);
      for (int idx : {3, 6}) {
        double A[9], J[9];
        A_matrix(dx_ + idx, A);
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) J[r * 3 + c] = A[c * 3 + r];
        // L rows taken from P (not from L)
        for (int c = 0; c < n; ++c)
          for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += J[r * 3 + k] * P[(idx + k) * n + c];
            L[(idx + r) * n + c] = s;
          }
        for (int c = 0; c < 12; ++c) {
          double tcol[3];
          for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += J[r * 3 + k] * K_x[(idx + k) * n + c];
            tcol[r] = s;
          }
          for (int r = 0; r < 3; ++r) K_x[(idx + r) * n + c] = tcol[r];
        }
        project_cols(L.data(), n, idx, 3, J);
        project_cols(P, n, idx, 3, J);
      }
      {
        const int idx = 21;
        double Nx[6], Mx[6], J2[4];
        s2_Nx_yy(x.grav, Nx);
        s2_Mx(x_prop.grav, dx_ + idx, Mx);
        matmul(Nx, Mx, J2, 2, 3, 2);
        for (int c = 0; c < n; ++c)
          for (int r = 0; r < 2; ++r) {
            double s = 0;
            for (int k = 0; k < 2; ++k) s += J2[r * 2 + k] * P[(idx + k) * n + c];
            L[(idx + r) * n + c] = s;
          }
        for (int c = 0; c < 12; ++c) {
          double tcol[2];
          for (int r = 0; r < 2; ++r) {
            double s = 0;
            for (int k = 0; k < 2; ++k) s += J2[r * 2 + k] * K_x[(idx + k) * n + c];
            tcol[r] = s;
          }
          for (int r = 0; r < 2; ++r) K_x[(idx + r) * n + c] = tcol[r];
        }
        project_cols(L.data(), n, idx, 2, J2);
        project_cols(P, n, idx, 2, J2);
      }
      std::vector<double> Pn(static_cast<size_t>(n) * n);
      for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
          double s = 0;
          for (int k = 0; k < 12; ++k) s += K_x[r * n + k] * P[k * n + c];
          Pn[r * n + c] = L[r * n + c] - s;
        }
      std::memcpy(P, Pn.data(), sizeof(double) * n * n);
      return passes;
    }
  }
  return passes;
}

}  // namespace orc
