// pcl_compat.hpp -- the slice of PCL / Eigen that the reference's caller
// (examples/run_global_registration.cpp:103-108, 202-251) touches, for builds WITHOUT PCL/Eigen.
//
// Define QB200_USE_REAL_PCL to compile the shim classes (quatro.hpp, fpfh_manager.hpp) against the
// real <pcl/...> and <Eigen/...> headers instead: pcl::PointXYZ is the same 16-byte record either way,
// so clouds are handed to the C-ABI without conversion.
#pragma once

#ifdef QB200_USE_REAL_PCL
#include <Eigen/Core>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/registration/registration.h>
namespace qbcompat {
template <class T> using shared_ptr = typename pcl::PointCloud<pcl::PointXYZ>::Ptr::element_type*;  // unused
}
#else

#include <array>
#include <cstddef>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

namespace Eigen {
// Column-major fixed/dynamic matrices with just the accessors the example needs.
struct Matrix4d {
  double m[16];
  static Matrix4d Identity() {
    Matrix4d r;
    for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.0 : 0.0;
    return r;
  }
  double& operator()(int r, int c) { return m[4 * c + r]; }
  double operator()(int r, int c) const { return m[4 * c + r]; }
  double* data() { return m; }
  const double* data() const { return m; }
};
struct Matrix3d {
  double m[9];
  static Matrix3d Identity() {
    Matrix3d r;
    for (int i = 0; i < 9; ++i) r.m[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return r;
  }
  double& operator()(int r, int c) { return m[3 * c + r]; }
  double operator()(int r, int c) const { return m[3 * c + r]; }
};
struct Vector3d {
  double v[3];
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
};
// stand-in for Eigen::Matrix<double, 3, Eigen::Dynamic>
struct Matrix3Xd {
  std::vector<double> d;
  void resize(int rows, int cols) { (void)rows; d.assign(3 * (size_t)cols, 0.0); }
  long cols() const { return (long)(d.size() / 3); }
  double& operator()(int r, long c) { return d[3 * (size_t)c + r]; }
  double operator()(int r, long c) const { return d[3 * (size_t)c + r]; }
};
}  // namespace Eigen

namespace pcl {

struct alignas(16) PointXYZ {
  float x = 0.f, y = 0.f, z = 0.f, data3 = 1.f;  // PCL pads PointXYZ to 16 bytes with data[3] = 1
  PointXYZ() = default;
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), data3(1.f) {}
};
static_assert(sizeof(PointXYZ) == 16, "pcl::PointXYZ must be a 16-byte record");

struct FPFHSignature33 {
  float histogram[33];
  static int descriptorSize() { return 33; }
};
static_assert(sizeof(FPFHSignature33) == 132, "pcl::FPFHSignature33 is 33 packed floats");

struct Normal {
  float normal_x, normal_y, normal_z, curvature;
};

template <class PointT>
class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  unsigned width = 0, height = 1;
  bool is_dense = true;

  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = 0; }
  void reserve(size_t n) { points.reserve(n); }
  void resize(size_t n) { points.resize(n); width = (unsigned)n; }
  void push_back(const PointT& p) { points.push_back(p); width = (unsigned)points.size(); }
  PointT& at(size_t i) { return points.at(i); }
  const PointT& at(size_t i) const { return points.at(i); }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
  typename std::vector<PointT>::iterator begin() { return points.begin(); }
  typename std::vector<PointT>::iterator end() { return points.end(); }
  typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
  typename std::vector<PointT>::const_iterator end() const { return points.end(); }
  PointCloud operator+(const PointCloud& o) const {
    PointCloud r = *this;
    r.points.insert(r.points.end(), o.points.begin(), o.points.end());
    r.width = (unsigned)r.points.size();
    return r;
  }
};

// pcl::Registration: only the members Quatro<> touches (include/quatro.hpp:80-104 of the reference).
template <class PointSource, class PointTarget, class Scalar = float>
class Registration {
 public:
  using PointCloudSource = PointCloud<PointSource>;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = PointCloud<PointTarget>;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  using Matrix4 = Eigen::Matrix4d;

  virtual ~Registration() = default;
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) { target_ = cloud; }
  PointCloudSourceConstPtr getInputSource() const { return input_; }
  PointCloudTargetConstPtr getInputTarget() const { return target_; }
  const std::string& getClassName() const { return reg_name_; }
  inline void setMaximumIterations(int n) { max_iterations_ = n; }

 protected:
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;
  std::string reg_name_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  int nr_iterations_ = 0, max_iterations_ = 10;
  bool converged_ = false;
};

inline void transformPointCloud(const PointCloud<PointXYZ>& in, PointCloud<PointXYZ>& out, const Eigen::Matrix4d& T) {
  out.points.resize(in.points.size());
  for (size_t i = 0; i < in.points.size(); ++i) {
    const PointXYZ& p = in.points[i];
    out.points[i] = PointXYZ((float)(T(0, 0) * p.x + T(0, 1) * p.y + T(0, 2) * p.z + T(0, 3)),
                             (float)(T(1, 0) * p.x + T(1, 1) * p.y + T(1, 2) * p.z + T(1, 3)),
                             (float)(T(2, 0) * p.x + T(2, 1) * p.y + T(2, 2) * p.z + T(2, 3)));
  }
  out.width = (unsigned)out.points.size();
}

namespace io {
// pcl::io::savePCDFile / loadPCDFile for xyz clouds (ASCII, PCD v0.7): what FPFHManager::save/loadFeaturePair need
inline int savePCDFile(const std::string& name, const PointCloud<PointXYZ>& c) {
  FILE* f = std::fopen(name.c_str(), "w");
  if (!f) return -1;
  std::fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
                  "WIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA ascii\n", c.points.size(), c.points.size());
  for (const PointXYZ& p : c.points) std::fprintf(f, "%.9g %.9g %.9g\n", p.x, p.y, p.z);
  std::fclose(f);
  return 0;
}
inline int loadPCDFile(const std::string& name, PointCloud<PointXYZ>& c) {
  FILE* f = std::fopen(name.c_str(), "r");
  if (!f) return -1;
  char line[256];
  size_t n = 0;
  bool data = false;
  while (std::fgets(line, sizeof(line), f)) {
    if (std::sscanf(line, "POINTS %zu", &n) == 1) continue;
    if (std::string(line).rfind("DATA ascii", 0) == 0) { data = true; break; }
  }
  c.clear();
  if (!data) { std::fclose(f); return -1; }
  float x, y, z;
  for (size_t i = 0; i < n && std::fscanf(f, "%f %f %f", &x, &y, &z) == 3; ++i) c.push_back(PointXYZ(x, y, z));
  std::fclose(f);
  return c.points.size() == n ? 0 : -1;
}
}  // namespace io

}  // namespace pcl

#define PCL_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
#endif  // QB200_USE_REAL_PCL

using PointType = pcl::PointXYZ;  // include/utility.h:109 of the reference
