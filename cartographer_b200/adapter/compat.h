// Layout-compatible stand-ins for the handful of Cartographer types that appear
// in the signatures of the scan matchers.  Used ONLY when the real headers are
// not on the include path (this image has no Eigen/glog/abseil/protobuf, so the
// reference cannot be compiled here); inside a Cartographer checkout the
// adapter includes the real headers instead (see the __has_include switch in
// scan_matchers_b200.h and INTEGRATION.md).
//
// Only the members the adapter touches are provided.
#ifndef CSM_ADAPTER_COMPAT_H_
#define CSM_ADAPTER_COMPAT_H_

#include <cstdint>
#include <vector>

namespace cartographer {

namespace transform {
// transform/rigid_transform.h:34-103 (Rigid2<double>): translation + Rotation2D angle.
class Rigid2d {
 public:
  struct Vector { double x_, y_; double x() const { return x_; } double y() const { return y_; } };
  struct Rotation2D { double angle_; double angle() const { return angle_; } };
  Rigid2d() : t_{0., 0.}, r_{0.} {}
  Rigid2d(const Vector& t, double rotation) : t_(t), r_{rotation} {}
  static Rigid2d Identity() { return Rigid2d(); }
  const Vector& translation() const { return t_; }
  Rotation2D rotation() const { return r_; }
 private:
  Vector t_;
  Rotation2D r_;
};
}  // namespace transform

namespace sensor {
// sensor/rangefinder_point.h:31-33 and sensor/point_cloud.h:33-92.
struct RangefinderPoint { struct { float v[3]; float x() const { return v[0]; } float y() const { return v[1]; } float z() const { return v[2]; } } position; };
class PointCloud {
 public:
  void push_back(const RangefinderPoint& p) { points_.push_back(p); }
  size_t size() const { return points_.size(); }
  const std::vector<RangefinderPoint>& points() const { return points_; }
 private:
  std::vector<RangefinderPoint> points_;
};
}  // namespace sensor

namespace transform {
// transform/rigid_transform.h:116-196 (Rigid3<double>) and Eigen::Quaterniond as used
// in the 3D matcher's signatures.
struct Quaterniond {
  double w_, x_, y_, z_;
  double w() const { return w_; } double x() const { return x_; }
  double y() const { return y_; } double z() const { return z_; }
};
class Rigid3d {
 public:
  struct Vector { double v[3]; double x() const { return v[0]; } double y() const { return v[1]; } double z() const { return v[2]; } };
  Rigid3d() : t_{{0., 0., 0.}}, q_{1., 0., 0., 0.} {}
  Rigid3d(const Vector& t, const Quaterniond& q) : t_(t), q_(q) {}
  const Vector& translation() const { return t_; }
  const Quaterniond& rotation() const { return q_; }
 private:
  Vector t_;
  Quaterniond q_;
};
}  // namespace transform

namespace mapping {
// mapping/3d/hybrid_grid.h:468-526: what the adapter needs is iteration over the
// non-zero voxels (HybridGrid::Iterator / begin()-end(), :413-460) and resolution().
class HybridGrid {
 public:
  struct Voxel { int x, y, z; uint16_t value; };
  explicit HybridGrid(float resolution) : resolution_(resolution) {}
  float resolution() const { return resolution_; }
  int grid_size() const { return grid_size_; }
  void Set(int x, int y, int z, uint16_t value) {
    voxels_.push_back(Voxel{x, y, z, value});
    while (x < -(grid_size_ >> 1) || x >= (grid_size_ >> 1) || y < -(grid_size_ >> 1) ||
           y >= (grid_size_ >> 1) || z < -(grid_size_ >> 1) || z >= (grid_size_ >> 1))
      grid_size_ <<= 1;
  }
  const std::vector<Voxel>& voxels() const { return voxels_; }
 private:
  float resolution_;
  int grid_size_ = 128;
  std::vector<Voxel> voxels_;
};
// mapping/trajectory_node.h:45-63 (the fields the 3D matcher reads).
struct TrajectoryNodeData {
  transform::Quaterniond gravity_alignment{1., 0., 0., 0.};
  sensor::PointCloud high_resolution_point_cloud;
  sensor::PointCloud low_resolution_point_cloud;
  std::vector<float> rotational_scan_matcher_histogram;   // Eigen::VectorXf
};
}  // namespace mapping

namespace mapping {
// mapping/2d/xy_index.h:34-45, mapping/2d/map_limits.h:40-95.
struct CellLimits { int num_x_cells = 0; int num_y_cells = 0; };
class MapLimits {
 public:
  struct Vector2d { double x_, y_; double x() const { return x_; } double y() const { return y_; } };
  MapLimits(double resolution, double max_x, double max_y, const CellLimits& c)
      : resolution_(resolution), max_{max_x, max_y}, cell_limits_(c) {}
  double resolution() const { return resolution_; }
  const Vector2d& max() const { return max_; }
  const CellLimits& cell_limits() const { return cell_limits_; }
 private:
  double resolution_;
  Vector2d max_;
  CellLimits cell_limits_;
};
// mapping/2d/grid_2d.h:37-141 — the read-only accessors the matcher ctor uses.
class Grid2D {
 public:
  Grid2D(const MapLimits& limits, float min_cost, float max_cost, std::vector<uint16_t> cells)
      : limits_(limits), min_(min_cost), max_(max_cost), cells_(std::move(cells)) {}
  const MapLimits& limits() const { return limits_; }
  float GetMinCorrespondenceCost() const { return min_; }
  float GetMaxCorrespondenceCost() const { return max_; }
  // protected in the reference (grid_2d.h:98-100); the real adapter is a friend or
  // uses ToProto().cells() — see INTEGRATION.md.
  const std::vector<uint16_t>& correspondence_cost_cells() const { return cells_; }
 private:
  MapLimits limits_;
  float min_, max_;
  std::vector<uint16_t> cells_;
};

namespace scan_matching {
namespace proto {
// proto/scan_matching/fast_correlative_scan_matcher_options_2d.proto
class FastCorrelativeScanMatcherOptions2D {
 public:
  double linear_search_window() const { return linear_search_window_; }
  double angular_search_window() const { return angular_search_window_; }
  int branch_and_bound_depth() const { return branch_and_bound_depth_; }
  void set_linear_search_window(double v) { linear_search_window_ = v; }
  void set_angular_search_window(double v) { angular_search_window_ = v; }
  void set_branch_and_bound_depth(int v) { branch_and_bound_depth_ = v; }
 private:
  double linear_search_window_ = 0., angular_search_window_ = 0.;
  int branch_and_bound_depth_ = 1;
};
// proto/scan_matching/real_time_correlative_scan_matcher_options.proto
class RealTimeCorrelativeScanMatcherOptions {
 public:
  double linear_search_window() const { return lin_; }
  double angular_search_window() const { return ang_; }
  double translation_delta_cost_weight() const { return wt_; }
  double rotation_delta_cost_weight() const { return wr_; }
  void set_linear_search_window(double v) { lin_ = v; }
  void set_angular_search_window(double v) { ang_ = v; }
  void set_translation_delta_cost_weight(double v) { wt_ = v; }
  void set_rotation_delta_cost_weight(double v) { wr_ = v; }
 private:
  double lin_ = 0., ang_ = 0., wt_ = 0., wr_ = 0.;
};
// proto/scan_matching/fast_correlative_scan_matcher_options_3d.proto
class FastCorrelativeScanMatcherOptions3D {
 public:
  int branch_and_bound_depth() const { return o.branch_and_bound_depth; }
  int full_resolution_depth() const { return o.full_resolution_depth; }
  double min_rotational_score() const { return o.min_rotational_score; }
  double min_low_resolution_score() const { return o.min_low_resolution_score; }
  double linear_xy_search_window() const { return o.linear_xy_search_window; }
  double linear_z_search_window() const { return o.linear_z_search_window; }
  double angular_search_window() const { return o.angular_search_window; }
  struct { int branch_and_bound_depth = 8, full_resolution_depth = 3;
           double min_rotational_score = 0.77, min_low_resolution_score = 0.55,
                  linear_xy_search_window = 5., linear_z_search_window = 1.,
                  angular_search_window = 0.2617993877991494; } o;
};
}  // namespace proto
}  // namespace scan_matching
}  // namespace mapping
}  // namespace cartographer

#endif  // CSM_ADAPTER_COMPAT_H_
