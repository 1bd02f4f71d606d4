// fast_limo/Common.hpp for the B200 build of the registration path.
//
// The reference's Common.hpp (fast_limo/Common.hpp:66-172) pulls in Eigen, PCL and Boost and defines the point type,
// IMUmeas, Extrinsics and the container typedefs.  On a ROS machine those libraries exist: define
// FLIMO_WITH_PCL_EIGEN and this header includes them and nothing below replaces anything.  In this repository's
// environment none of them is installed, so the few pieces of their interfaces that the fast_limo::Localizer /
// fast_limo::Mapper call surface and the ROS wrapper's call sites (src/main.cpp:16-93,178-206, ROSutils.hpp) touch are
// provided as minimal stand-ins with the same names and member access (Eigen::Vector3f::operator(), x(), y(), z();
// Eigen::Quaternionf::x() .. w(), toRotationMatrix(); pcl::PointCloud<T>::points / Ptr / ConstPtr; ...).
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#ifdef FLIMO_WITH_PCL_EIGEN
#include <Eigen/Dense>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#else
namespace Eigen {
template <typename T, int R, int C>
struct Matrix {                                   // fixed size, column-major like Eigen's default
  T d[R * C];
  Matrix() { for (T& v : d) v = T(0); }
  Matrix(T x, T y, T z) { static_assert(R * C == 3, "3-vector"); d[0] = x; d[1] = y; d[2] = z; }
  Matrix(T x, T y, T z, T w) { static_assert(R * C == 4, "4-vector"); d[0] = x; d[1] = y; d[2] = z; d[3] = w; }
  T& operator()(int i) { return d[i]; }
  T operator()(int i) const { return d[i]; }
  T& operator[](int i) { return d[i]; }
  T operator[](int i) const { return d[i]; }
  T& operator()(int r, int c) { return d[c * R + r]; }
  T operator()(int r, int c) const { return d[c * R + r]; }
  T x() const { return d[0]; }
  T y() const { return d[1]; }
  T z() const { return d[2]; }
  T* data() { return d; }
  const T* data() const { return d; }
  static constexpr int rows() { return R; }
  static constexpr int cols() { return C; }
  static Matrix Identity() { Matrix m; for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1); return m; }
  static Matrix Zero() { return Matrix(); }
  T norm() const { T s = 0; for (const T& v : d) s += v * v; return std::sqrt(s); }
  template <typename U> Matrix<U, R, C> cast() const { Matrix<U, R, C> o; for (int i = 0; i < R * C; ++i) o.d[i] = (U)d[i]; return o; }
  Matrix<T, C, R> transpose() const { Matrix<T, C, R> o; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) o(c, r) = (*this)(r, c); return o; }
  Matrix operator+(const Matrix& b) const { Matrix o; for (int i = 0; i < R * C; ++i) o.d[i] = d[i] + b.d[i]; return o; }
  Matrix operator-(const Matrix& b) const { Matrix o; for (int i = 0; i < R * C; ++i) o.d[i] = d[i] - b.d[i]; return o; }
  Matrix operator-() const { Matrix o; for (int i = 0; i < R * C; ++i) o.d[i] = -d[i]; return o; }
  Matrix operator*(T s) const { Matrix o; for (int i = 0; i < R * C; ++i) o.d[i] = d[i] * s; return o; }
  Matrix operator/(T s) const { Matrix o; for (int i = 0; i < R * C; ++i) o.d[i] = d[i] / s; return o; }
  Matrix& operator+=(const Matrix& b) { for (int i = 0; i < R * C; ++i) d[i] += b.d[i]; return *this; }
  Matrix& operator/=(T s) { for (int i = 0; i < R * C; ++i) d[i] /= s; return *this; }
  template <int K> Matrix<T, R, K> operator*(const Matrix<T, C, K>& b) const {
    Matrix<T, R, K> o;
    for (int r = 0; r < R; ++r)
      for (int k = 0; k < K; ++k) { T s = 0; for (int c = 0; c < C; ++c) s += (*this)(r, c) * b(c, k); o(r, k) = s; }
    return o;
  }
  Matrix cross(const Matrix& b) const {
    static_assert(R * C == 3, "3-vector");
    return Matrix(d[1] * b.d[2] - d[2] * b.d[1], d[2] * b.d[0] - d[0] * b.d[2], d[0] * b.d[1] - d[1] * b.d[0]);
  }
  T dot(const Matrix& b) const { T s = 0; for (int i = 0; i < R * C; ++i) s += d[i] * b.d[i]; return s; }
  Matrix normalized() const { return *this / norm(); }
};
using Vector3f = Matrix<float, 3, 1>;
using Vector4f = Matrix<float, 4, 1>;
using Vector3d = Matrix<double, 3, 1>;
using Matrix3f = Matrix<float, 3, 3>;
using Matrix4f = Matrix<float, 4, 4>;
using Matrix3d = Matrix<double, 3, 3>;

template <typename T>
struct Quaternion {                               // coefficient order of the constructor as in Eigen: (w, x, y, z)
  T qx = 0, qy = 0, qz = 0, qw = 1;
  Quaternion() = default;
  Quaternion(T w, T x, T y, T z) : qx(x), qy(y), qz(z), qw(w) {}
  explicit Quaternion(const Matrix<T, 3, 3>& R) {  // rotation matrix -> quaternion (Shepperd)
    const T tr = R(0, 0) + R(1, 1) + R(2, 2);
    if (tr > 0) {
      const T s = std::sqrt(tr + 1) * 2;
      qw = s / 4; qx = (R(2, 1) - R(1, 2)) / s; qy = (R(0, 2) - R(2, 0)) / s; qz = (R(1, 0) - R(0, 1)) / s;
    } else if (R(0, 0) > R(1, 1) && R(0, 0) > R(2, 2)) {
      const T s = std::sqrt(1 + R(0, 0) - R(1, 1) - R(2, 2)) * 2;
      qw = (R(2, 1) - R(1, 2)) / s; qx = s / 4; qy = (R(0, 1) + R(1, 0)) / s; qz = (R(0, 2) + R(2, 0)) / s;
    } else if (R(1, 1) > R(2, 2)) {
      const T s = std::sqrt(1 + R(1, 1) - R(0, 0) - R(2, 2)) * 2;
      qw = (R(0, 2) - R(2, 0)) / s; qx = (R(0, 1) + R(1, 0)) / s; qy = s / 4; qz = (R(1, 2) + R(2, 1)) / s;
    } else {
      const T s = std::sqrt(1 + R(2, 2) - R(0, 0) - R(1, 1)) * 2;
      qw = (R(1, 0) - R(0, 1)) / s; qx = (R(0, 2) + R(2, 0)) / s; qy = (R(1, 2) + R(2, 1)) / s; qz = s / 4;
    }
  }
  T x() const { return qx; }
  T y() const { return qy; }
  T z() const { return qz; }
  T w() const { return qw; }
  T& x() { return qx; }
  T& y() { return qy; }
  T& z() { return qz; }
  T& w() { return qw; }
  void normalize() { const T n = std::sqrt(qx * qx + qy * qy + qz * qz + qw * qw); qx /= n; qy /= n; qz /= n; qw /= n; }
  Quaternion conjugate() const { return Quaternion(qw, -qx, -qy, -qz); }
  template <typename U> Quaternion<U> cast() const { return Quaternion<U>((U)qw, (U)qx, (U)qy, (U)qz); }
  Quaternion operator*(const Quaternion& b) const {
    return Quaternion(qw * b.qw - qx * b.qx - qy * b.qy - qz * b.qz, qw * b.qx + qx * b.qw + qy * b.qz - qz * b.qy,
                      qw * b.qy + qy * b.qw + qz * b.qx - qx * b.qz, qw * b.qz + qz * b.qw + qx * b.qy - qy * b.qx);
  }
  Quaternion& operator*=(const Quaternion& b) { *this = *this * b; return *this; }
  Matrix<T, 3, 3> toRotationMatrix() const {      // Eigen's evaluation order (QuaternionBase::toRotationMatrix)
    Matrix<T, 3, 3> R;
    const T tx = T(2) * qx, ty = T(2) * qy, tz = T(2) * qz;
    const T twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    R(0, 0) = T(1) - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = T(1) - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = T(1) - (txx + tyy);
    return R;
  }
  static Quaternion FromTwoVectors(const Matrix<T, 3, 1>& a, const Matrix<T, 3, 1>& b) {
    const Matrix<T, 3, 1> v0 = a.normalized(), v1 = b.normalized();
    const T c = v0.dot(v1);
    if (c < T(-1) + T(1e-6)) return Quaternion(0, 1, 0, 0);       // opposite vectors: any orthogonal axis (rare; Eigen uses an SVD here)
    const Matrix<T, 3, 1> axis = v0.cross(v1);
    const T s = std::sqrt((T(1) + c) * T(2)), invs = T(1) / s;
    return Quaternion(s * T(0.5), axis(0) * invs, axis(1) * invs, axis(2) * invs);
  }
};
using Quaternionf = Quaternion<float>;
using Quaterniond = Quaternion<double>;

struct MatrixXd {                                  // dynamic, column-major; what Localizer::calculate_H fills
  int r = 0, c = 0;
  std::vector<double> d;
  MatrixXd() = default;
  MatrixXd(int rows, int cols) : r(rows), c(cols), d((size_t)rows * cols, 0.0) {}
  static MatrixXd Zero(int rows, int cols) { return MatrixXd(rows, cols); }
  void resize(int rows, int cols) { r = rows; c = cols; d.assign((size_t)rows * cols, 0.0); }
  int rows() const { return r; }
  int cols() const { return c; }
  double& operator()(int i, int j) { return d[(size_t)j * r + i]; }
  double operator()(int i, int j) const { return d[(size_t)j * r + i]; }
};
struct VectorXd {
  std::vector<double> d;
  VectorXd() = default;
  explicit VectorXd(int n) : d((size_t)n, 0.0) {}
  void resize(int n) { d.assign((size_t)n, 0.0); }
  int size() const { return (int)d.size(); }
  double& operator()(int i) { return d[i]; }
  double operator()(int i) const { return d[i]; }
};
}  // namespace Eigen

namespace pcl {
template <typename PointT>
struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 1;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
};
struct PointXYZ { float x = 0, y = 0, z = 0, pad = 1.f; };
}  // namespace pcl
#endif  // FLIMO_WITH_PCL_EIGEN

namespace fast_limo {
class State;
class Plane;
class Match;
class Mapper;
class Localizer;
struct Config;

enum class SensorType { OUSTER, VELODYNE, HESAI, LIVOX, UNKNOWN };   // Common.hpp:87-93

#ifndef FLIMO_WITH_PCL_EIGEN
// fast_limo::Point (Common.hpp:100-113): xyz + padding, intensity, per-sensor time union — 32 bytes, 16-byte aligned
struct alignas(16) Point {
  Point() : x(0.f), y(0.f), z(0.f), pad_(1.f), intensity(0.f), timestamp(0.0) {}
  Point(float x_, float y_, float z_) : x(x_), y(y_), z(z_), pad_(1.f), intensity(0.f), timestamp(0.0) {}
  float x, y, z, pad_;
  float intensity;
  union {
    std::uint32_t t;   // (Ouster) ns since the beginning of the scan
    float time;        // (Velodyne) s since the beginning of the scan
    double timestamp;  // (Hesai) absolute s; (Livox) absolute s * 1e9
  };
};
static_assert(sizeof(Point) == 32, "fast_limo::Point is 32 bytes");
#endif

struct Extrinsics {                                // Common.hpp:115-124
  struct SE3 {
    Eigen::Vector3f t;
    Eigen::Matrix3f R;
  };
  SE3 imu2baselink, lidar2baselink;
  Eigen::Matrix4f imu2baselink_T, lidar2baselink_T;
};

struct IMUmeas {                                   // Common.hpp:126-132
  double stamp = 0.0;
  double dt = 0.0;   // difference between this and the previous measurement
  Eigen::Vector3f ang_vel, lin_accel;
  Eigen::Quaternionf q;
};

template <typename T> using shared_ptr = std::shared_ptr<T>;
template <typename T, typename... Args> std::shared_ptr<T> make_shared(Args&&... args) { return std::make_shared<T>(std::forward<Args>(args)...); }
}  // namespace fast_limo

// IKFoM/use-ikfom.hpp:12-21 — the filter state as the hook sees it (double; declaration order = flimo.h's state26)
struct state_ikfom {
  Eigen::Vector3d pos;
  Eigen::Quaterniond rot, offset_R_L_I;
  Eigen::Vector3d offset_T_L_I, vel, bg, ba, grav;
};

typedef fast_limo::Point PointType;
typedef pcl::PointXYZ MapPoint;
typedef std::vector<pcl::PointXYZ> MapPoints;
typedef std::vector<fast_limo::Match> Matches;
typedef std::vector<fast_limo::State> States;
