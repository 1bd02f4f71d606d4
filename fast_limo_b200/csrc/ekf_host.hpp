// Host side of the iterated error-state Kalman update (product code, float64, no dependencies).
//
// The reference runs esekf::update_iterated_dyn_share_modified (IKFoM_toolkit/esekfom/esekfom.hpp:
// 1620-1823) with Eigen on the host; the north star keeps this O(23^3) algebra on the host and moves
// only the N-sized work (h_share_model + H^T H + H^T h) to the GPU.  Eigen/Boost are not available
// in this build environment, so the few pieces of MTK (manifold toolkit) the update needs are
// restated here on plain fixed-size arrays:
//   state layout / compound (+) (-)   IKFoM/use-ikfom.hpp:12-21, mtk/build_manifold.hpp:192-200
//   SO3 (+),(-), exp, log             mtk/types/SOn.hpp:233-239,284-297; mtk/src/mtkmath.hpp:143-175,250-289
//   S2  (+),(-), Bx, Nx_yy, Mx        mtk/types/S2.hpp:136-167,216-231,259-280  (S2<double,98090,10000,1>)
//   A_matrix                          mtk/src/mtkmath.hpp:236-247
//
// The update is exposed as a pass-wise state machine (begin / state / step / end) so that the
// measurement pass — and, with several GPUs, the all-reduce of its 96 doubles — sits between
// `state()` and `step()`.
//
// Differences from the reference, all documented in DESIGN.md:
//  * the measurement enters in reduced form (HTH, HTh, n_rows).  For n_rows >= 23 this is the
//    reference's own branch (esekfom.hpp:1722-1729).  For n_rows < 23 the reference uses the
//    algebraically identical gain K = P H^T (H P H^T / R + I)^-1 / R (:1701-1709) and then reads an
//    UNINITIALISED HTH in the degeneracy filter (:1736); here the information form is used for the
//    gain and HTH := 0 in the filter (=> no pose update), which is what the CPU oracle does too.
//  * Eigen::EigenSolver's eigenvector basis (degenerate geometry only) is replaced by a symmetric
//    Jacobi decomposition; when all six eigenvalues exceed D the filter is the identity either way.
#pragma once
#include <cmath>
#include <math.h>
#include <cstdint>
#include <cstring>

// The math below is shared with the device-side update (csrc/ekf_step.hpp, run by the last CTA of the
// registration kernel): every helper is __host__ __device__ when compiled by nvcc.
#if defined(__CUDACC__)
#define FLIMO_HD __host__ __device__
#else
#define FLIMO_HD
#endif
#if defined(__CUDA_ARCH__)
#define FLIMO_UNROLL _Pragma("unroll")
#else
#define FLIMO_UNROLL
#endif

namespace flimo {
namespace ekf {

constexpr int N = 23;          // error-state dimension
constexpr double kTol = 1e-11; // MTK::tolerance<double>()
constexpr double kS2Len = 98090.0 / 10000.0;

template <int Rw, int Cl>
struct Mat {
  double a[Rw * Cl];
  FLIMO_HD double& operator()(int r, int c) { return a[r * Cl + c]; }
  FLIMO_HD double operator()(int r, int c) const { return a[r * Cl + c]; }
  FLIMO_HD static Mat zero() {
    Mat m;
    for (int i = 0; i < Rw * Cl; ++i) m.a[i] = 0.0;
    return m;
  }
  FLIMO_HD static Mat identity() {
    Mat m = zero();
    for (int i = 0; i < (Rw < Cl ? Rw : Cl); ++i) m(i, i) = 1.0;
    return m;
  }
  FLIMO_HD Mat<Cl, Rw> T() const {
    Mat<Cl, Rw> t;
    for (int r = 0; r < Rw; ++r)
      for (int c = 0; c < Cl; ++c) t(c, r) = (*this)(r, c);
    return t;
  }
};
template <int A, int B, int C>
FLIMO_HD inline Mat<A, C> operator*(const Mat<A, B>& x, const Mat<B, C>& y) {
  Mat<A, C> o;
  for (int r = 0; r < A; ++r)
    for (int c = 0; c < C; ++c) {
      double s = 0.0;
      for (int k = 0; k < B; ++k) s += x(r, k) * y(k, c);
      o(r, c) = s;
    }
  return o;
}
template <int K>
struct Vec {                         // plain aggregate (std::array is not usable in device code)
  double v[K];
  FLIMO_HD double& operator[](int i) { return v[i]; }
  FLIMO_HD double operator[](int i) const { return v[i]; }
  FLIMO_HD double* data() { return v; }
  FLIMO_HD const double* data() const { return v; }
};
using V3 = Vec<3>;
using Q4 = Vec<4>;   // x y z w

struct State {
  V3 pos;
  Q4 rot;
  Q4 offR;
  V3 offT, vel, bg, ba, grav;
  FLIMO_HD void load(const double* f) {
    for (int i = 0; i < 3; ++i) { pos[i] = f[i]; offT[i] = f[11 + i]; vel[i] = f[14 + i]; bg[i] = f[17 + i]; ba[i] = f[20 + i]; grav[i] = f[23 + i]; }
    for (int i = 0; i < 4; ++i) { rot[i] = f[3 + i]; offR[i] = f[7 + i]; }
  }
  FLIMO_HD void store(double* f) const {
    for (int i = 0; i < 3; ++i) { f[i] = pos[i]; f[11 + i] = offT[i]; f[14 + i] = vel[i]; f[17 + i] = bg[i]; f[20 + i] = ba[i]; f[23 + i] = grav[i]; }
    for (int i = 0; i < 4; ++i) { f[3 + i] = rot[i]; f[7 + i] = offR[i]; }
  }
};

FLIMO_HD inline V3 cross(const V3& a, const V3& b) {
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
FLIMO_HD inline double dot(const V3& a, const V3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
FLIMO_HD inline double norm(const V3& a) { return sqrt(dot(a, a)); }

FLIMO_HD inline Mat<3, 3> skew(const V3& v) {
  Mat<3, 3> m = Mat<3, 3>::zero();
  m(0, 1) = -v[2]; m(0, 2) = v[1];
  m(1, 0) = v[2];  m(1, 2) = -v[0];
  m(2, 0) = -v[1]; m(2, 1) = v[0];
  return m;
}

FLIMO_HD inline Q4 qmul(const Q4& a, const Q4& b) {
  return {a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1],
          a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2],
          a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0],
          a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]};
}

FLIMO_HD inline Mat<3, 3> rotmat(const Q4& q) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  Mat<3, 3> R;
  R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz;       R(0, 2) = txz + twy;
  R(1, 0) = txy + twz;       R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
  R(2, 0) = txz - twy;       R(2, 1) = tyz + twx;       R(2, 2) = 1 - (txx + tyy);
  return R;
}

// cos(sqrt(x2)), sin(sqrt(x2))/sqrt(x2) with the Taylor branch of mtkmath.hpp:143-175
FLIMO_HD inline void cos_sinc(double x2, double& c, double& sc) {
  const double bound = 1.220703125e-4;   // sqrt(sqrt(epsilon)) = 2^-13 exactly (mtkmath.hpp:146)
  if (x2 >= bound) {
    const double x = sqrt(x2);
    c = cos(x);
    sc = sin(x) / x;
    return;
  }
  const double inv[7] = {1 / 3., 1 / 4., 1 / 5., 1 / 6., 1 / 7., 1 / 8., 1 / 9.};
  double ci = 1., si = 1., term = -1 / 2. * x2;
  for (int i = 0; i < 3; ++i) {
    ci += term;
    term *= inv[2 * i];
    si += term;
    term *= -inv[2 * i + 1] * x2;
  }
  c = ci;
  sc = si;
}

// quaternion exp(scale * v) in MTK's convention (mtkmath.hpp:250-257)
FLIMO_HD inline Q4 qexp(const V3& v, double scale) {
  double c, sc;
  cos_sinc(scale * scale * dot(v, v), c, sc);
  const double m = sc * scale;
  return {m * v[0], m * v[1], m * v[2], c};
}

FLIMO_HD inline void so3_plus(Q4& q, const V3& d) { q = qmul(q, qexp(d, 0.5)); }
FLIMO_HD inline V3 so3_minus(const Q4& a, const Q4& b) {   // log(b^-1 a), SOn.hpp:237-239 + mtkmath.hpp:269-289
  const Q4 r = qmul({-b[0], -b[1], -b[2], b[3]}, a);
  double nv = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (nv < kTol) nv = kTol;
  const double s = 2.0 / nv * atan(nv / r[3]);
  return {s * r[0], s * r[1], s * r[2]};
}

FLIMO_HD inline Mat<3, 3> A_matrix(const V3& v) {
  const double sq = dot(v, v), n = sqrt(sq);
  Mat<3, 3> A = Mat<3, 3>::identity();
  if (n < kTol) return A;
  const Mat<3, 3> K = skew(v), K2 = K * K;
  const double a = (1 - cos(n)) / sq, b = (1 - sin(n) / n) / sq;
  for (int i = 0; i < 9; ++i) A.a[i] += a * K.a[i] + b * K2.a[i];
  return A;
}

FLIMO_HD inline Mat<3, 2> s2_Bx(const V3& v) {   // chart type 1 (S2.hpp:216-231)
  Mat<3, 2> B = Mat<3, 2>::zero();
  const double L = kS2Len;
  if (v[0] + L > kTol) {
    const double d = L + v[0];
    B(0, 0) = -v[1];               B(0, 1) = -v[2];
    B(1, 0) = L - v[1] * v[1] / d; B(1, 1) = -v[2] * v[1] / d;
    B(2, 0) = -v[2] * v[1] / d;    B(2, 1) = L - v[2] * v[2] / d;
    for (int i = 0; i < 6; ++i) B.a[i] /= L;
  } else {
    B(1, 1) = -1;
    B(2, 0) = 1;
  }
  return B;
}

FLIMO_HD inline void s2_plus(V3& v, double d0, double d1) {
  const Mat<3, 2> B = s2_Bx(v);
  const V3 Bu = {B(0, 0) * d0 + B(0, 1) * d1, B(1, 0) * d0 + B(1, 1) * d1, B(2, 0) * d0 + B(2, 1) * d1};
  const Mat<3, 3> R = rotmat(qexp(Bu, 0.5));
  v = {R(0, 0) * v[0] + R(0, 1) * v[1] + R(0, 2) * v[2], R(1, 0) * v[0] + R(1, 1) * v[1] + R(1, 2) * v[2],
       R(2, 0) * v[0] + R(2, 1) * v[1] + R(2, 2) * v[2]};
}

FLIMO_HD inline void s2_minus(const V3& a, const V3& b, double& r0, double& r1) {
  const double v_sin = norm(cross(a, b)), v_cos = dot(a, b);
  const double theta = atan2(v_sin, v_cos);
  if (v_sin < kTol) {
    r0 = (fabs(theta) > kTol) ? 3.1415926 : 0.0;
    r1 = 0.0;
    return;
  }
  const Mat<3, 2> B = s2_Bx(b);
  const V3 w = cross(b, a);
  const double f = theta / v_sin;
  r0 = f * (B(0, 0) * w[0] + B(1, 0) * w[1] + B(2, 0) * w[2]);
  r1 = f * (B(0, 1) * w[0] + B(1, 1) * w[1] + B(2, 1) * w[2]);
}

FLIMO_HD inline Mat<2, 3> s2_Nx_yy(const V3& v) {
  Mat<2, 3> Nx = s2_Bx(v).T() * skew(v);
  for (int i = 0; i < 6; ++i) Nx.a[i] *= 1 / kS2Len / kS2Len;
  return Nx;
}

FLIMO_HD inline Mat<3, 2> s2_Mx(const V3& v, double d0, double d1) {
  const Mat<3, 2> B = s2_Bx(v);
  Mat<3, 3> negH = skew(v);
  for (int i = 0; i < 9; ++i) negH.a[i] = -negH.a[i];
  if (sqrt(d0 * d0 + d1 * d1) < kTol) return negH * B;
  const V3 Bu = {B(0, 0) * d0 + B(0, 1) * d1, B(1, 0) * d0 + B(1, 1) * d1, B(2, 0) * d0 + B(2, 1) * d1};
  // the reference builds this rotation with scale scalar(1/2) == 0 (integer division, S2.hpp:277):
  const Mat<3, 3> E = rotmat(qexp(Bu, 0.0));
  return (E * negH) * (A_matrix(Bu).T() * B);
}

FLIMO_HD inline void boxplus(State& x, const double* d) {
  for (int i = 0; i < 3; ++i) x.pos[i] += d[i];
  so3_plus(x.rot, {d[3], d[4], d[5]});
  so3_plus(x.offR, {d[6], d[7], d[8]});
  for (int i = 0; i < 3; ++i) {
    x.offT[i] += d[9 + i];
    x.vel[i] += d[12 + i];
    x.bg[i] += d[15 + i];
    x.ba[i] += d[18 + i];
  }
  s2_plus(x.grav, d[21], d[22]);
}

FLIMO_HD inline void boxminus(const State& x, const State& y, double* d) {
  for (int i = 0; i < 3; ++i) d[i] = x.pos[i] - y.pos[i];
  const V3 r = so3_minus(x.rot, y.rot), r2 = so3_minus(x.offR, y.offR);
  for (int i = 0; i < 3; ++i) {
    d[3 + i] = r[i];
    d[6 + i] = r2[i];
    d[9 + i] = x.offT[i] - y.offT[i];
    d[12 + i] = x.vel[i] - y.vel[i];
    d[15 + i] = x.bg[i] - y.bg[i];
    d[18 + i] = x.ba[i] - y.ba[i];
  }
  s2_minus(x.grav, y.grav, d[21], d[22]);
}

// Inverse of an n x n row-major matrix by Gauss-Jordan elimination with partial (row) pivoting on the
// augmented system [A | I]; every inner loop runs over a contiguous row, so the host compiler
// vectorises it (the 23x23 case takes a few microseconds).  Returns false on a zero / non-finite pivot.
template <int n>
FLIMO_HD inline bool invert(Mat<n, n>& M) {
  double w[n][2 * n];
  for (int r = 0; r < n; ++r) {
    for (int c = 0; c < n; ++c) {
      w[r][c] = M(r, c);
      w[r][n + c] = (r == c) ? 1.0 : 0.0;
    }
  }
  for (int c = 0; c < n; ++c) {
    int p = c;
    double best = fabs(w[c][c]);
    for (int r = c + 1; r < n; ++r) {
      const double v = fabs(w[r][c]);
      if (v > best) {
        best = v;
        p = r;
      }
    }
    if (!(best > 0.0) || best > 1.7e308) return false;   // zero, NaN or Inf pivot
    if (p != c)
      for (int k = 0; k < 2 * n; ++k) {
        const double t = w[c][k];
        w[c][k] = w[p][k];
        w[p][k] = t;
      }
    const double inv = 1.0 / w[c][c];
    for (int k = 0; k < 2 * n; ++k) w[c][k] *= inv;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = w[r][c];
      if (f == 0.0) continue;
      for (int k = 0; k < 2 * n; ++k) w[r][k] -= f * w[c][k];
    }
  }
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) M(r, c) = w[r][n + c];
  return true;
}

// true when every eigenvalue of the symmetric 6x6 S exceeds `floor` by a safe margin: S - floor*I has
// a Cholesky factorisation whose pivots are all comfortably positive.
FLIMO_HD inline bool all_eigs_above(const Mat<6, 6>& S, double floor) {
  double L[6][6];
  double scale = 0.0;
FLIMO_UNROLL
  for (int i = 0; i < 6; ++i) scale = fmax(scale, fabs(S(i, i)));
  const double tiny = 1e-9 * (scale + fabs(floor)) + 1e-300;
  bool ok = true;
FLIMO_UNROLL
  for (int j = 0; j < 6; ++j) {
    double d = S(j, j) - floor;
FLIMO_UNROLL
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
    if (!(d > tiny)) ok = false;
    const double dj = sqrt(ok ? d : 1.0);
    L[j][j] = dj;
    const double inv = 1.0 / dj;
FLIMO_UNROLL
    for (int i = j + 1; i < 6; ++i) {
      double v = S(i, j);
FLIMO_UNROLL
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = v * inv;
    }
  }
  return ok;
}

// Symmetric Jacobi eigen-decomposition (6x6), columns of V are eigenvectors.
FLIMO_HD inline void sym_eig6(const Mat<6, 6>& S, double w[6], Mat<6, 6>& V) {
  Mat<6, 6> A = S;
  V = Mat<6, 6>::identity();
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0;
    for (int i = 0; i < 6; ++i)
      for (int j = i + 1; j < 6; ++j) off += A(i, j) * A(i, j);
    if (off < 1e-300) break;
    for (int p = 0; p < 6; ++p)
      for (int q = p + 1; q < 6; ++q) {
        if (A(p, q) == 0.0) continue;
        const double th = (A(q, q) - A(p, p)) / (2 * A(p, q));
        const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
        const double c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 6; ++k) {
          const double x = A(k, p), y = A(k, q);
          A(k, p) = c * x - s * y;
          A(k, q) = s * x + c * y;
        }
        for (int k = 0; k < 6; ++k) {
          const double x = A(p, k), y = A(q, k);
          A(p, k) = c * x - s * y;
          A(q, k) = s * x + c * y;
        }
        for (int k = 0; k < 6; ++k) {
          const double x = V(k, p), y = V(k, q);
          V(k, p) = c * x - s * y;
          V(k, q) = s * x + c * y;
        }
      }
  }
  for (int i = 0; i < 6; ++i) w[i] = A(i, i);
}

// Fast exit of the degeneracy filter: every eigenvalue of the pose block HTH[0:6, 0:6] is safely above both
// thresholds of the filter (D and, for the product test, 1e-20^(1/6)); the reference then computes
// V^-1 * V * dx = dx, so the decomposition can be skipped.
FLIMO_HD inline bool all_eigs_above6(const double* HTH144, double floor) {
  Mat<6, 6> S6;
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) S6(r, c) = HTH144[r * 12 + c];
  return all_eigs_above(S6, floor);
}

// The filter itself (esekfom.hpp:1736-1744): dxn[0:6] = V^-1 * (V with the rows of small eigenvalues zeroed) * dxk[0:6].
// have_hth = false reproduces the N < 23 branch (HTH := 0, see the header comment).  Returns false if V is singular
// (dxn[0:6] is then left equal to dxk[0:6]).
FLIMO_HD inline bool degeneracy_filter(const double* HTH144, bool have_hth, double D, const double* dxk, double* dxn) {
  Mat<6, 6> S6 = Mat<6, 6>::zero();
  if (have_hth)
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) S6(r, c) = HTH144[r * 12 + c];
  double w[6];
  Mat<6, 6> V;
  sym_eig6(S6, w, V);
  double prod = 1;
  for (int i = 0; i < 6; ++i) prod *= w[i];
  if (prod < 1e-20) V = Mat<6, 6>::identity();
  Mat<6, 6> Sel = V;
  for (int k = 0; k < 6; ++k)
    if (w[k] < D)
      for (int c = 0; c < 6; ++c) Sel(k, c) = 0.0;    // row k, as in the reference
  Mat<6, 6> Vi = V;
  if (!invert<6>(Vi)) return false;
  const Mat<6, 6> T = Vi * Sel;
  for (int r = 0; r < 6; ++r) {
    double s = 0;
    for (int c = 0; c < 6; ++c) s += T(r, c) * dxk[c];
    dxn[r] = s;
  }
  return true;
}

// Left/right congruence of a square matrix by a small block Jacobian J at [idx, idx+B).
template <int B>
FLIMO_HD inline void rows_by(Mat<N, N>& P, int idx, const Mat<B, B>& J, const Mat<N, N>& src) {
  for (int c = 0; c < N; ++c) {
    double t[B];
    for (int i = 0; i < B; ++i) {
      double s = 0;
      for (int k = 0; k < B; ++k) s += J(i, k) * src(idx + k, c);
      t[i] = s;
    }
    for (int i = 0; i < B; ++i) P(idx + i, c) = t[i];
  }
}
template <int B>
FLIMO_HD inline void cols_by_T(Mat<N, N>& P, int idx, const Mat<B, B>& J) {
  for (int r = 0; r < N; ++r) {
    double t[B];
    for (int i = 0; i < B; ++i) {
      double s = 0;
      for (int k = 0; k < B; ++k) s += P(r, idx + k) * J(i, k);
      t[i] = s;
    }
    for (int i = 0; i < B; ++i) P(r, idx + i) = t[i];
  }
}

class IteratedUpdate {
 public:
  void begin(const double* x26, const double* P529, int max_iter, const double* limit23, double R, double D) {
    x_.load(x26);
    x_prop_ = x_;
    std::memcpy(P_prop_.a, P529, sizeof(P_prop_.a));
    P_ = P_prop_;
    std::memcpy(limit_, limit23, sizeof(limit_));
    max_iter_ = max_iter;
    R_ = R;
    D_ = D;
    iter_ = -1;
    conv_count_ = 0;
    done_ = (iter_ >= max_iter_);   // max_iter < 0: the reference's loop body never runs
    failed_ = false;
    passes_ = 0;
  }
  bool done() const { return done_; }
  int passes() const { return passes_; }
  void state(double* x26) const { x_.store(x26); }
  void end(double* x26, double* P529) const {
    x_.store(x26);
    std::memcpy(P529, P_.a, sizeof(P_.a));
  }

  // One pass of esekfom.hpp:1652-1819 given the reduced measurement.  Returns done().
  bool step(const double* HTH144, const double* HTh12, int64_t n_rows) {
    if (done_) return true;
    ++passes_;
    double dx[N], dx_new[N];
    boxminus(x_, x_prop_, dx);
    std::memcpy(dx_new, dx, sizeof(dx));
    P_ = P_prop_;
    for (int idx : {3, 6}) {
      const Mat<3, 3> J = A_matrix({dx[idx], dx[idx + 1], dx[idx + 2]}).T();
      double t[3];
      for (int i = 0; i < 3; ++i) t[i] = J(i, 0) * dx_new[idx] + J(i, 1) * dx_new[idx + 1] + J(i, 2) * dx_new[idx + 2];
      for (int i = 0; i < 3; ++i) dx_new[idx + i] = t[i];
      rows_by<3>(P_, idx, J, P_);
      cols_by_T<3>(P_, idx, J);
    }
    {
      const Mat<2, 2> J = s2_Nx_yy(x_.grav) * s2_Mx(x_prop_.grav, dx[21], dx[22]);
      const double a = J(0, 0) * dx_new[21] + J(0, 1) * dx_new[22], b = J(1, 0) * dx_new[21] + J(1, 1) * dx_new[22];
      dx_new[21] = a;
      dx_new[22] = b;
      rows_by<2>(P_, 21, J, P_);
      cols_by_T<2>(P_, 21, J);
    }

    // gain in information form (esekfom.hpp:1722-1729).  Only the first 12 columns of
    //   P_inv = ((P/R)^-1 + U HTH U^T)^-1,  U = [I12; 0],
    // are ever used.  They satisfy  P_inv U = (P U / R) (I12 + HTH P11 / R)^-1  (multiply the defining
    // equation by P/R and restrict to the first 12 rows), which needs ONE 12x12 inverse instead of the
    // reference's two 23x23 ones and is better conditioned.  reference_form_ evaluates it the reference's way.
    double K_h[N];
    Mat<N, N> K_x = Mat<N, N>::zero();
    Mat<N, 12> X;                                    // P_inv[:, 0:12]
    if (reference_form_) {
      Mat<N, N> Pi = P_;
      for (double& v : Pi.a) v /= R_;
      bool ok = invert<N>(Pi);
      for (int r = 0; r < 12; ++r)
        for (int c = 0; c < 12; ++c) Pi(r, c) += HTH144[r * 12 + c];
      ok = ok && invert<N>(Pi);   // P_inv
      if (!ok) return fail_singular();
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < 12; ++c) X(r, c) = Pi(r, c);
    } else {
      Mat<12, 12> M;
      for (int r = 0; r < 12; ++r)
        for (int c = 0; c < 12; ++c) {
          double s2 = (r == c) ? 1.0 : 0.0;
          for (int k = 0; k < 12; ++k) s2 += HTH144[r * 12 + k] * (P_(k, c) / R_);
          M(r, c) = s2;
        }
      if (!invert<12>(M)) return fail_singular();
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < 12; ++c) {
          double s2 = 0;
          for (int k = 0; k < 12; ++k) s2 += (P_(r, k) / R_) * M(k, c);
          X(r, c) = s2;
        }
    }
    for (int r = 0; r < N; ++r) {
      double s = 0;
      for (int k = 0; k < 12; ++k) s += X(r, k) * HTh12[k];
      K_h[r] = s;
      for (int c = 0; c < 12; ++c) {
        double s2 = 0;
        for (int k = 0; k < 12; ++k) s2 += X(r, k) * HTH144[k * 12 + c];
        K_x(r, c) = s2;
      }
    }
    double dxk[N];
    for (int r = 0; r < N; ++r) {
      double s = K_h[r];
      for (int c = 0; c < N; ++c) s += (K_x(r, c) - (r == c ? 1.0 : 0.0)) * dx_new[c];
      dxk[r] = s;
    }

    // degeneracy filter on the pose block (esekfom.hpp:1736-1744)
    double dxn[N];
    std::memcpy(dxn, dxk, sizeof(dxk));
    if (!(n_rows >= N && all_eigs_above6(HTH144, std::fmax(D_, 1e-3)))) degeneracy_filter(HTH144, n_rows >= N, D_, dxk, dxn);

    boxplus(x_, dxn);
    bool converge = true;
    for (int r = 0; r < N; ++r)
      if (std::fabs(dxk[r]) > limit_[r]) {
        converge = false;
        break;
      }
    if (converge) ++conv_count_;
    std::memcpy(last_dx_, dxk, sizeof(dxk));

    if (conv_count_ > 1 || iter_ == max_iter_ - 1 || force_final_) {
      Mat<N, N> L = P_;
      for (int idx : {3, 6}) {
        const Mat<3, 3> J = A_matrix({dxk[idx], dxk[idx + 1], dxk[idx + 2]}).T();
        rows_by<3>(L, idx, J, P_);          // rows of L come from P (esekfom.hpp:1776-1778)
        for (int c = 0; c < 12; ++c) {
          double t[3];
          for (int i = 0; i < 3; ++i) t[i] = J(i, 0) * K_x(idx, c) + J(i, 1) * K_x(idx + 1, c) + J(i, 2) * K_x(idx + 2, c);
          for (int i = 0; i < 3; ++i) K_x(idx + i, c) = t[i];
        }
        cols_by_T<3>(L, idx, J);
        cols_by_T<3>(P_, idx, J);
      }
      {
        const Mat<2, 2> J = s2_Nx_yy(x_.grav) * s2_Mx(x_prop_.grav, dxk[21], dxk[22]);
        rows_by<2>(L, 21, J, P_);
        for (int c = 0; c < 12; ++c) {
          const double a = J(0, 0) * K_x(21, c) + J(0, 1) * K_x(22, c), b = J(1, 0) * K_x(21, c) + J(1, 1) * K_x(22, c);
          K_x(21, c) = a;
          K_x(22, c) = b;
        }
        cols_by_T<2>(L, 21, J);
        cols_by_T<2>(P_, 21, J);
      }
      Mat<N, N> Pn;
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
          double s = 0;
          for (int k = 0; k < 12; ++k) s += K_x(r, k) * P_(k, c);
          Pn(r, c) = L(r, c) - s;
        }
      P_ = Pn;
      done_ = true;
      return true;
    }
    ++iter_;
    if (iter_ >= max_iter_) done_ = true;   // unreachable (the branch above fires at max_iter-1); defensive
    return done_;
  }

  // The last pass of an update whose earlier passes ran elsewhere (device-resident registration): continue from the
  // state the last pass was evaluated at and take the final branch (esekfom.hpp:1764-1819) whatever the local
  // convergence test says — the caller has already decided that this pass ends the loop.
  bool finish(const double* x_eval26, const double* HTH144, const double* HTh12, int64_t n_rows) {
    x_.load(x_eval26);
    done_ = false;
    force_final_ = true;
    const bool d = step(HTH144, HTh12, n_rows);
    force_final_ = false;
    return d;
  }

  const double* last_dx() const { return last_dx_; }
  bool failed() const { return failed_; }   // a singular / non-finite system was met: x and P are the propagated ones

 private:
  // A singular pivot (NaN/Inf in HTH or P, degenerate covariance): the update is abandoned, state and covariance stay
  // at the propagated values and the caller reports FLIMO_ERR_STATE.
  bool fail_singular() {
    x_ = x_prop_;
    P_ = P_prop_;
    failed_ = true;
    done_ = true;
    return true;
  }

 private:
  State x_, x_prop_;
  Mat<N, N> P_, P_prop_;
  double limit_[N];
  double last_dx_[N];
  int max_iter_ = 0, iter_ = -1, conv_count_ = 0, passes_ = 0;
  double R_ = 0.001, D_ = 5.0;
  bool done_ = true, failed_ = false, force_final_ = false;
 public:
  bool reference_form_ = false;   // true: form the gain with the reference's two 23x23 inversions (FLIMO_EKF_REFERENCE_FORM=1)
 private:
};

}  // namespace ekf
}  // namespace flimo
