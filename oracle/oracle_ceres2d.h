// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.h).
// CeresScanMatcher2D (mapping/internal/2d/scan_matching/ceres_scan_matcher_2d.{h,cc}),
// the refinement ConstraintBuilder2D runs on every found match
// (mapping/internal/constraints/constraint_builder_2d.cc:245-249).
//
// PARITY UNPINNED AGAINST CERES ITSELF.  The solver is a third-party dependency that is
// absent from /root/reference and from this image: Ceres Solver, pinned by the reference at
// commit 58c5edae2f7c4d2533fe8a975c1f5f0b892dfd3e (bazel/repositories.bzl:136-142;
// package.xml:48 libceres-dev).  What is restated here:
//   * the reference's own cost functors, line by line (occupied_space_cost_function_2d.cc:
//     31-106, translation_delta_cost_functor_2d.h:41-45, rotation_delta_cost_functor_2d.h:
//     40-43) including the kPadding = INT_MAX / 4 shift of the interpolation coordinates;
//   * Ceres' published algorithms for everything the reference's call reaches with
//     Solver::Options defaults + {DENSE_QR, use_nonmonotonic_steps, max_num_iterations}
//     (ceres_scan_matcher_2d.cc:52-57, common/internal/ceres_solver_options.cc:37-44):
//     BiCubicInterpolator / CubicHermiteSpline (Catmull-Rom, ceres/cubic_interpolation.h),
//     the forward-mode derivative of the functor, the trust-region loop with the
//     Levenberg-Marquardt strategy, Jacobi column scaling, the function / gradient / parameter
//     tolerances and the non-monotonic step acceptance of Conn, Gould & Toint, Alg. 10.1.2.
//     The 3-parameter damped least-squares step is solved through the normal equations
//     (Ceres: Householder QR of the stacked [J; D] — same minimiser, different rounding).
// Pinned to what the reference holds for this path: ceres_scan_matcher_2d_test.cc:40-97
// (four known-answer starts, pose within 1e-2, final cost within 1e-2 of 0) and
// occupied_space_cost_function_2d_test.cc:31-51 (residual == kMaxProbability exactly) —
// tests/test_oracle_golden_ceres2d.py.
#ifndef ORACLE_CERES2D_H_
#define ORACLE_CERES2D_H_

#include <vector>

#include "oracle_2d.h"

namespace oracle {

// proto/scan_matching/ceres_scan_matcher_options_2d.proto + common/proto/ceres_solver_options.proto
struct CeresScanMatcherOptions2D {
  double occupied_space_weight = 20.;  // configuration_files/pose_graph.lua:30-39
  double translation_weight = 10.;
  double rotation_weight = 1.;
  bool use_nonmonotonic_steps = true;
  int max_num_iterations = 10;
};

enum CeresTermination {
  kCeresNoConvergence = 0,        // max_num_iterations reached
  kCeresFunctionTolerance = 1,
  kCeresGradientTolerance = 2,
  kCeresParameterTolerance = 3,
  kCeresMinTrustRegionRadius = 4,
  kCeresInvalidSteps = 5,         // max_num_consecutive_invalid_steps reached (FAILURE)
};

struct CeresSummary2D {
  double initial_cost = 0., final_cost = 0.;
  int iterations = 0;              // trust-region iterations after the initial evaluation
  int num_successful_steps = 0;
  int termination = kCeresNoConvergence;
};

// residuals (n + 3) and the row-major (n + 3) x 3 Jacobian of the three residual blocks at
// `pose` = {x, y, theta}; `jacobian` may be null (then the plain-double evaluation path).
void EvaluateCeresResiduals2D(const ProbabilityGrid& grid, const PointCloud& cloud,
                              const CeresScanMatcherOptions2D& options,
                              const double target_translation[2], double target_angle,
                              const double pose[3], std::vector<double>* residuals,
                              std::vector<double>* jacobian);

// CeresScanMatcher2D::Match (ceres_scan_matcher_2d.cc:62-107)
void CeresMatch2D(const ProbabilityGrid& grid, const PointCloud& cloud,
                  const CeresScanMatcherOptions2D& options, const double target_translation[2],
                  const double initial_pose[3], double pose_estimate[3],
                  CeresSummary2D* summary);

}  // namespace oracle

#endif  // ORACLE_CERES2D_H_
