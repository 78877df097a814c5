// Drop-in C++ classes with the reference's names, namespaces and signatures that
// forward to the C ABI (include/csm_abi.h).  A Cartographer checkout replaces the
// bodies of
//   mapping/internal/2d/scan_matching/fast_correlative_scan_matcher_2d.{h,cc}
//   mapping/internal/2d/scan_matching/real_time_correlative_scan_matcher_2d.{h,cc}
// with this header; callers — ConstraintBuilder2D
// (constraints/constraint_builder_2d.cc:213-215, 226-228) and
// LocalTrajectoryBuilder2D::ScanMatch (internal/2d/local_trajectory_builder_2d.cc:77-82)
// — compile unchanged.
//
// Error convention: the reference CHECK-aborts on programmer errors; the adapter
// keeps that contract by aborting with the C ABI's error string (the ABI itself
// never aborts).  "No pose above min_score" returns false and leaves the outputs
// untouched, exactly like fast_correlative_scan_matcher_2d.cc:253-261.
#ifndef CSM_ADAPTER_SCAN_MATCHERS_B200_H_
#define CSM_ADAPTER_SCAN_MATCHERS_B200_H_

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/csm_abi.h"

#if defined(__has_include) && __has_include("cartographer/mapping/2d/grid_2d.h")
#include "cartographer/mapping/2d/grid_2d.h"
#include "cartographer/mapping/proto/scan_matching/fast_correlative_scan_matcher_options_2d.pb.h"
#include "cartographer/mapping/proto/scan_matching/real_time_correlative_scan_matcher_options.pb.h"
#include "cartographer/sensor/point_cloud.h"
#include "cartographer/transform/rigid_transform.h"
#define CSM_ADAPTER_REAL_CARTOGRAPHER 1
#else
#include "compat.h"
#endif

namespace cartographer {
namespace mapping {
namespace scan_matching {

namespace b200_internal {
inline void Check(csm_status s) {
  if (s != CSM_OK) {
    std::fprintf(stderr, "libcsm_b200: %s\n", csm_last_error_string());
    std::abort();  // glog CHECK semantics of the reference
  }
}
inline std::vector<float> Flatten(const sensor::PointCloud& point_cloud) {
  std::vector<float> xyz;
  xyz.reserve(3 * point_cloud.size());
  for (const auto& p : point_cloud.points()) {
    xyz.push_back(p.position.x());
    xyz.push_back(p.position.y());
    xyz.push_back(p.position.z());
  }
  return xyz;
}
}  // namespace b200_internal

// fast_correlative_scan_matcher_2d.h:112-136
class FastCorrelativeScanMatcher2D {
 public:
  FastCorrelativeScanMatcher2D(const Grid2D& grid,
                               const proto::FastCorrelativeScanMatcherOptions2D& options,
                               int device = 0)
      : options_(options) {
    const MapLimits& l = grid.limits();
    b200_internal::Check(csm_stack2d_create(
        grid.correspondence_cost_cells().data(), l.cell_limits().num_x_cells,
        l.cell_limits().num_y_cells, l.resolution(), l.max().x(), l.max().y(),
        grid.GetMinCorrespondenceCost(), grid.GetMaxCorrespondenceCost(),
        options.branch_and_bound_depth(), device, &stack_));
  }
  ~FastCorrelativeScanMatcher2D() { csm_stack2d_destroy(stack_); }
  FastCorrelativeScanMatcher2D(const FastCorrelativeScanMatcher2D&) = delete;
  FastCorrelativeScanMatcher2D& operator=(const FastCorrelativeScanMatcher2D&) = delete;

  bool Match(const transform::Rigid2d& initial_pose_estimate,
             const sensor::PointCloud& point_cloud, float min_score, float* score,
             transform::Rigid2d* pose_estimate) const {
    const double init[3] = {initial_pose_estimate.translation().x(),
                            initial_pose_estimate.translation().y(),
                            initial_pose_estimate.rotation().angle()};
    return Run(init, 0, point_cloud, min_score, score, pose_estimate);
  }

  bool MatchFullSubmap(const sensor::PointCloud& point_cloud, float min_score, float* score,
                       transform::Rigid2d* pose_estimate) const {
    const double init[3] = {0., 0., 0.};
    return Run(init, 1, point_cloud, min_score, score, pose_estimate);
  }

  const csm_stack2d* stack() const { return stack_; }  // for the batched ConstraintBuilder

 private:
  bool Run(const double init[3], int full, const sensor::PointCloud& point_cloud,
           float min_score, float* score, transform::Rigid2d* pose_estimate) const {
    if (score == nullptr || pose_estimate == nullptr) std::abort();  // CHECK(:232-233)
    const std::vector<float> xyz = b200_internal::Flatten(point_cloud);
    int32_t found = 0;
    float s = 0.f;
    double pose[3] = {0., 0., 0.};
    b200_internal::Check(csm_match2d(stack_, xyz.data(), static_cast<int32_t>(point_cloud.size()),
                                     init, full, options_.linear_search_window(),
                                     options_.angular_search_window(), min_score, &found, &s,
                                     pose, nullptr));
    if (!found) return false;
    *score = s;
    *pose_estimate = transform::Rigid2d({pose[0], pose[1]}, pose[2]);
    return true;
  }
  const proto::FastCorrelativeScanMatcherOptions2D options_;
  csm_stack2d* stack_ = nullptr;
};

// real_time_correlative_scan_matcher_2d.h:53-85 (ProbabilityGrid path)
class RealTimeCorrelativeScanMatcher2D {
 public:
  explicit RealTimeCorrelativeScanMatcher2D(
      const proto::RealTimeCorrelativeScanMatcherOptions& options, int device = 0)
      : options_(options), device_(device) {}

  double Match(const transform::Rigid2d& initial_pose_estimate,
               const sensor::PointCloud& point_cloud, const Grid2D& grid,
               transform::Rigid2d* pose_estimate) const {
    if (pose_estimate == nullptr) std::abort();  // CHECK(:121)
    const MapLimits& l = grid.limits();
    const std::vector<float> xyz = b200_internal::Flatten(point_cloud);
    const double init[3] = {initial_pose_estimate.translation().x(),
                            initial_pose_estimate.translation().y(),
                            initial_pose_estimate.rotation().angle()};
    double score = 0., pose[3] = {0., 0., 0.};
    b200_internal::Check(csm_rt_match2d(
        grid.correspondence_cost_cells().data(), l.cell_limits().num_x_cells,
        l.cell_limits().num_y_cells, l.resolution(), l.max().x(), l.max().y(), xyz.data(),
        static_cast<int32_t>(point_cloud.size()), init, options_.linear_search_window(),
        options_.angular_search_window(), options_.translation_delta_cost_weight(),
        options_.rotation_delta_cost_weight(), device_, &score, pose, nullptr));
    *pose_estimate = transform::Rigid2d({pose[0], pose[1]}, pose[2]);
    return score;
  }

 private:
  const proto::RealTimeCorrelativeScanMatcherOptions options_;
  int device_;
};

}  // namespace scan_matching
}  // namespace mapping
}  // namespace cartographer

#endif  // CSM_ADAPTER_SCAN_MATCHERS_B200_H_
