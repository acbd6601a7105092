// c2g_shims.h — the few Eigen / PCL types that appear in the signatures of the reference's ContourManager / ContourDB
// surface.  When the real libraries are installed the real headers are used and this file adds nothing;
// in this repository's build image they are absent, so minimal stand-ins with the same names and the member functions
// the cont2_batch_bin_test loop touches are provided.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <vector>

#if defined(__has_include) && __has_include(<Eigen/Geometry>) && !defined(C2G_FORCE_SHIMS)
#include <Eigen/Core>
#include <Eigen/Geometry>
#define C2G_HAVE_EIGEN 1
typedef Eigen::Matrix<float, 2, 1> V2F;
typedef Eigen::Matrix<float, 2, 2> M2F;
typedef Eigen::Matrix<double, 2, 1> V2D;
typedef Eigen::Matrix<double, 2, 2> M2D;
#else
#define C2G_HAVE_EIGEN 0
namespace c2g_shim {
template <typename T>
struct Vec2 {
  T d[2] = {0, 0};
  Vec2() {}
  Vec2(T a, T b) { d[0] = a; d[1] = b; }
  T &x() { return d[0]; }
  T &y() { return d[1]; }
  const T &x() const { return d[0]; }
  const T &y() const { return d[1]; }
  T &operator()(int i) { return d[i]; }
  const T &operator()(int i) const { return d[i]; }
  T norm() const { return std::sqrt(d[0] * d[0] + d[1] * d[1]); }
  Vec2 operator-(const Vec2 &o) const { return Vec2(d[0] - o.d[0], d[1] - o.d[1]); }
  Vec2 operator+(const Vec2 &o) const { return Vec2(d[0] + o.d[0], d[1] + o.d[1]); }
};
template <typename T>
struct Mat2 {  // column-major like Eigen
  T d[4] = {0, 0, 0, 0};
  T &operator()(int r, int c) { return d[c * 2 + r]; }
  const T &operator()(int r, int c) const { return d[c * 2 + r]; }
  const T *data() const { return d; }
};
}  // namespace c2g_shim
typedef c2g_shim::Vec2<float> V2F;
typedef c2g_shim::Mat2<float> M2F;
typedef c2g_shim::Vec2<double> V2D;
typedef c2g_shim::Mat2<double> M2D;
namespace Eigen {
// Isometry2d stand-in: 3x3 homogeneous matrix with the calls used by the harness (operator(), translation(), rotate,
// pretranslate, setIdentity, inverse, operator*).
struct Isometry2d {
  double m[3][3];
  Isometry2d() { setIdentity(); }
  static Isometry2d Identity() { return Isometry2d(); }
  void setIdentity() {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) m[r][c] = r == c ? 1.0 : 0.0;
  }
  double &operator()(int r, int c) { return m[r][c]; }
  const double &operator()(int r, int c) const { return m[r][c]; }
  V2D translation() const { return V2D(m[0][2], m[1][2]); }
  void rotate(double a) {  // linear = linear * R(a)
    const double c = std::cos(a), s = std::sin(a);
    const double l00 = m[0][0], l01 = m[0][1], l10 = m[1][0], l11 = m[1][1];
    m[0][0] = l00 * c + l01 * s;
    m[0][1] = -l00 * s + l01 * c;
    m[1][0] = l10 * c + l11 * s;
    m[1][1] = -l10 * s + l11 * c;
  }
  void pretranslate(const V2D &t) {
    m[0][2] += t.x();
    m[1][2] += t.y();
  }
  Isometry2d inverse() const {
    Isometry2d r;
    r.m[0][0] = m[0][0];
    r.m[0][1] = m[1][0];
    r.m[1][0] = m[0][1];
    r.m[1][1] = m[1][1];
    r.m[0][2] = -(r.m[0][0] * m[0][2] + r.m[0][1] * m[1][2]);
    r.m[1][2] = -(r.m[1][0] * m[0][2] + r.m[1][1] * m[1][2]);
    return r;
  }
  Isometry2d operator*(const Isometry2d &b) const {
    Isometry2d r;
    for (int i = 0; i < 2; ++i) {
      for (int j = 0; j < 2; ++j) r.m[i][j] = m[i][0] * b.m[0][j] + m[i][1] * b.m[1][j];
      r.m[i][2] = (m[i][0] * b.m[0][2] + m[i][1] * b.m[1][2]) + m[i][2];
    }
    return r;
  }
};
}  // namespace Eigen
#endif

#if defined(__has_include) && __has_include(<pcl/point_cloud.h>) && !defined(C2G_FORCE_SHIMS)
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#else
namespace pcl {
struct PointXYZ {
  float x = 0, y = 0, z = 0;
};
struct PCLHeader {
  uint64_t stamp = 0;
};
template <typename PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  std::vector<PointT> points;
  PCLHeader header;
  size_t size() const { return points.size(); }
  void reserve(size_t n) { points.reserve(n); }
  void push_back(const PointT &p) { points.push_back(p); }
};
}  // namespace pcl
#endif
