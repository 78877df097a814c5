// Post-match refinement: CeresScanMatcher2D::Match
// (mapping/internal/2d/scan_matching/ceres_scan_matcher_2d.cc:62-107), the call
// ConstraintBuilder2D makes on every found match (constraint_builder_2d.cc:245-249).
//
// Ceres itself is not linked (it is a third-party dependency of the reference, absent here,
// pinned there at commit 58c5edae, bazel/repositories.bzl:136-142).  What runs instead is a
// batched solver for exactly the problem that call builds — 3 parameters {x, y, theta},
// n + 3 residuals — one CTA per match, the whole trust-region loop on the device:
//   * residuals of the reference's cost functors (occupied_space_cost_function_2d.cc:42-69,
//     translation_delta_cost_functor_2d.h:41-45, rotation_delta_cost_functor_2d.h:40-43),
//     the occupied-space term through Ceres' BiCubicInterpolator (Catmull-Rom splines over the
//     4 x 4 cells around the point) with the kPadding = INT_MAX / 4 coordinate shift the
//     reference applies (:61-66, :72) — the shift quantises the interpolation coordinate to
//     2^-23 of a cell, so it has to be reproduced;
//   * threads stride the scan points, accumulate cost = 1/2 |r|^2, g = J^T r and H = J^T J
//     (10 doubles) and the block reduces them; thread 0 then does what Ceres' trust-region
//     minimiser does with Solver::Options defaults + {DENSE_QR, use_nonmonotonic_steps,
//     max_num_iterations} (ceres_scan_matcher_2d.cc:52-57): Jacobi column scaling fixed at the
//     first Jacobian, Levenberg-Marquardt damping D^2 = clamp(diag(Js^T Js)) / radius, the
//     3 x 3 damped normal equations (Ceres: Householder QR of [Js; D] — the same minimiser),
//     step quality against the model decrease, non-monotonic acceptance (Conn, Gould &
//     Toint, Alg. 10.1.2), radius update, and the function / gradient / parameter tolerances;
//     the lowest-cost iterate is what is returned, as Ceres does under non-monotonic steps.
// Doubles throughout, compiled without FMA contraction (Makefile: -fmad=false for this
// file) so that every per-point value is the one the oracle's restatement computes; sums are
// block-tree ordered, cos / sin come from the device's double-precision routines (<= 2 ulp).
#include <algorithm>
#include <climits>
#include <cmath>

// CSM_REFINE_DEVICE_ONLY: tests/emulation/refine2d_emulation.cc includes the kernels below
// into a CPU harness (SIMT shims, one std::thread per CUDA thread) to check their control
// flow against the oracle where no GPU is present; the library itself never defines it.
#ifndef CSM_REFINE_DEVICE_ONLY
#include "rtgrid.cuh"
#endif

namespace csm {

constexpr int kRefThreads = 256;
constexpr int kRefWarps = kRefThreads / 32;
constexpr int kPadding = INT_MAX / 4;   // occupied_space_cost_function_2d.cc:72

struct RefJobDev {
  const uint16_t* cells;   // the submap grid on the device (row pitch `pitch`)
  int nx, ny, pitch, n;
  long long xyz_off;       // first float of the job's cloud in the upload buffer
  double resolution, max_x, max_y;
  double target[2];        // target_translation
  double init[3];          // initial_pose_estimate {x, y, angle}
};

struct RefOpts {
  double occupied_space_weight, translation_weight, rotation_weight;
  int use_nonmonotonic_steps, max_num_iterations;
  float k_scale, cost_bias, max_cost;   // value -> correspondence cost (value_conversion_tables.cc:29-37)
};

struct RefResultDev {
  double pose[3];
  double initial_cost, final_cost;
  int iterations, num_successful_steps, termination, pad;
};

// GridArrayAdapter::GetValue (occupied_space_cost_function_2d.cc:78-87)
__device__ __forceinline__ double PaddedValue(const RefJobDev& J, const RefOpts& P, int row,
                                              int column) {
  const int iy = row - kPadding, ix = column - kPadding;
  if (ix < 0 || iy < 0 || ix >= J.nx || iy >= J.ny) return static_cast<double>(P.max_cost);
  const int value = __ldg(J.cells + static_cast<size_t>(iy) * J.pitch + ix) & 0x7fff;
  const float cost = value == 0 ? P.max_cost
                                : __fadd_rn(__fmul_rn(__int2float_rn(value), P.k_scale), P.cost_bias);
  return static_cast<double>(cost);
}

// ceres/cubic_interpolation.h, CubicHermiteSpline (Catmull-Rom)
template <bool kValue, bool kDeriv>
__device__ __forceinline__ void CubicHermiteSpline(double p0, double p1, double p2, double p3,
                                                   double x, double* f, double* dfdx) {
  const double a = 0.5 * (-p0 + 3.0 * p1 - 3.0 * p2 + p3);
  const double b = 0.5 * (2.0 * p0 - 5.0 * p1 + 4.0 * p2 - p3);
  const double c = 0.5 * (-p0 + p2);
  const double d = p1;
  if (kValue) *f = d + x * (c + x * (b + x * a));
  if (kDeriv) *dfdx = c + x * (2.0 * b + 3.0 * a * x);
}

// BiCubicInterpolator::Evaluate
template <bool kJac>
__device__ __forceinline__ void BiCubic(const RefJobDev& J, const RefOpts& P, double r, double c,
                                        double* f, double* dfdr, double* dfdc) {
  const int row = static_cast<int>(floor(r));
  const int col = static_cast<int>(floor(c));
  double fr[4], dfr[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int rr = row - 1 + k;
    CubicHermiteSpline<true, kJac>(PaddedValue(J, P, rr, col - 1), PaddedValue(J, P, rr, col),
                                   PaddedValue(J, P, rr, col + 1), PaddedValue(J, P, rr, col + 2),
                                   c - col, &fr[k], &dfr[k]);
  }
  CubicHermiteSpline<true, kJac>(fr[0], fr[1], fr[2], fr[3], r - row, f, dfdr);
  if (kJac) CubicHermiteSpline<true, false>(dfr[0], dfr[1], dfr[2], dfr[3], r - row, dfdc, nullptr);
}

// One occupied-space residual (and its Jacobian row) at pose x with cs = cos, sn = sin of x[2].
template <bool kJac>
__device__ __forceinline__ void PointResidual(const RefJobDev& J, const RefOpts& P,
                                              double scaling, const double* x, double cs,
                                              double sn, double px, double py, double* res,
                                              double* jrow) {
  const double wx = (cs * px + (-sn) * py) + x[0] * 1.0;
  const double wy = (sn * px + cs * py) + x[1] * 1.0;
  double f, dfdr = 0., dfdc = 0.;
  if (!kJac) {
    const double r = (J.max_x - wx) / J.resolution - 0.5 + static_cast<double>(kPadding);
    const double c = (J.max_y - wy) / J.resolution - 0.5 + static_cast<double>(kPadding);
    BiCubic<false>(J, P, r, c, &f, nullptr, nullptr);
    *res = scaling * f;
    return;
  }
  // on dual numbers the division by the resolution multiplies by its reciprocal
  const double inverse_resolution = 1.0 / J.resolution;
  const double r = (J.max_x - wx) * inverse_resolution - 0.5 + static_cast<double>(kPadding);
  const double c = (J.max_y - wy) * inverse_resolution - 0.5 + static_cast<double>(kPadding);
  BiCubic<true>(J, P, r, c, &f, &dfdr, &dfdc);
  *res = scaling * f;
  const double dwx[3] = {1.0, 0.0, (-sn) * px + (-cs) * py};
  const double dwy[3] = {0.0, 1.0, cs * px + (-sn) * py};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double dr = -dwx[k] * inverse_resolution, dc = -dwy[k] * inverse_resolution;
    jrow[k] = scaling * (dfdr * dr + dfdc * dc);
  }
}

// acc = {sum r^2, g0..2, h xx xy xt yy yt tt}; every thread returns the block total.
template <int kN>
__device__ __forceinline__ void BlockSum(double* acc, double (*s_part)[10], double* s_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kN; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_part[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kN) {
    double v = 0.;
#pragma unroll
    for (int w = 0; w < kRefWarps; ++w) v += s_part[w][threadIdx.x];
    s_tot[threadIdx.x] = v;
  }
  __syncthreads();
}

// cost, gradient and normal matrix of all three residual blocks at x (block-wide; the
// result lands in s_tot: {cost, g[3], h[6]}).
template <bool kJac>
__device__ __forceinline__ void EvaluateAt(const RefJobDev& J, const RefOpts& P,
                                           const float* __restrict__ xyz, const double* x,
                                           double (*s_part)[10], double* s_tot) {
  const double scaling = P.occupied_space_weight / sqrt(static_cast<double>(J.n));
  double sn, cs;
  sincos(x[2], &sn, &cs);
  double acc[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) acc[k] = 0.;
  for (int i = threadIdx.x; i < J.n; i += kRefThreads) {
    const double px = static_cast<double>(xyz[3 * static_cast<size_t>(i)]);
    const double py = static_cast<double>(xyz[3 * static_cast<size_t>(i) + 1]);
    double res, jr[3];
    PointResidual<kJac>(J, P, scaling, x, cs, sn, px, py, &res, jr);
    acc[0] += res * res;
    if (kJac) {
      acc[1] += jr[0] * res;
      acc[2] += jr[1] * res;
      acc[3] += jr[2] * res;
      acc[4] += jr[0] * jr[0];
      acc[5] += jr[0] * jr[1];
      acc[6] += jr[0] * jr[2];
      acc[7] += jr[1] * jr[1];
      acc[8] += jr[1] * jr[2];
      acc[9] += jr[2] * jr[2];
    }
  }
  BlockSum<kJac ? 10 : 1>(acc, s_part, s_tot);
  if (threadIdx.x == 0) {
    // translation / rotation priors (the rotation prior is on the INITIAL angle,
    // ceres_scan_matcher_2d.cc:94-97)
    const double wt = P.translation_weight, wr = P.rotation_weight;
    const double r0 = wt * (x[0] - J.target[0]), r1 = wt * (x[1] - J.target[1]);
    const double r2 = wr * (x[2] - J.init[2]);
    double sq = s_tot[0];
    sq += r0 * r0;
    sq += r1 * r1;
    sq += r2 * r2;
    s_tot[0] = 0.5 * sq;
    if (kJac) {
      s_tot[1] += wt * r0;
      s_tot[2] += wt * r1;
      s_tot[3] += wr * r2;
      s_tot[4] += wt * wt;
      s_tot[7] += wt * wt;
      s_tot[9] += wr * wr;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ bool SolveSpd3(const double* a, const double* b, double* y) {
  if (!(a[0] > 0.)) return false;
  const double l00 = sqrt(a[0]);
  const double l10 = a[1] / l00, l20 = a[2] / l00;
  const double l11sq = a[3] - l10 * l10;
  if (!(l11sq > 0.)) return false;
  const double l11 = sqrt(l11sq);
  const double l21 = (a[4] - l20 * l10) / l11;
  const double l22sq = a[5] - l20 * l20 - l21 * l21;
  if (!(l22sq > 0.)) return false;
  const double l22 = sqrt(l22sq);
  const double z0 = b[0] / l00;
  const double z1 = (b[1] - l10 * z0) / l11;
  const double z2 = (b[2] - l20 * z0 - l21 * z1) / l22;
  y[2] = z2 / l22;
  y[1] = (z1 - l21 * y[2]) / l11;
  y[0] = (z0 - l10 * y[1] - l20 * y[2]) / l00;
  return isfinite(y[0]) && isfinite(y[1]) && isfinite(y[2]);
}

__device__ __forceinline__ double Norm3(const double* v) {
  return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}

enum { kCmdEvalCandidate = 0, kCmdAccept = 1, kCmdRejected = 2, kCmdDone = 3 };
enum {
  kTermNoConvergence = 0, kTermFunctionTolerance = 1, kTermGradientTolerance = 2,
  kTermParameterTolerance = 3, kTermMinRadius = 4, kTermInvalidSteps = 5
};

__global__ void __launch_bounds__(kRefThreads)
k_ceres_match2d(const RefJobDev* __restrict__ jobs, RefOpts P, const float* __restrict__ cloud,
                RefResultDev* __restrict__ results) {
  __shared__ double s_part[kRefWarps][10];
  __shared__ double s_tot[10];
  __shared__ double s_pose[3];
  __shared__ int s_cmd;
  const RefJobDev J = jobs[blockIdx.x];
  const float* __restrict__ xyz = cloud + J.xyz_off;

  // Solver::Options defaults the reference leaves untouched
  const double kInitialRadius = 1e4, kMaxRadius = 1e16, kMinRadius = 1e-32;
  const double kMinRelativeDecrease = 1e-3, kMinLmDiagonal = 1e-6, kMaxLmDiagonal = 1e32;
  const int kMaxConsecutiveInvalidSteps = 5;
  const double kFunctionTolerance = 1e-6, kGradientTolerance = 1e-10, kParameterTolerance = 1e-8;
  const int max_nonmonotonic = P.use_nonmonotonic_steps ? 5 : 0;

  // ---- state of the minimiser (meaningful in thread 0 only) -----------------------
  double x[3] = {J.init[0], J.init[1], J.init[2]};
  double best[3] = {x[0], x[1], x[2]};
  double g[3], h[6], scale[3] = {1., 1., 1.};
  double x_cost, x_norm, minimum_cost, initial_cost;
  double radius = kInitialRadius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  double diagonal[3] = {0., 0., 0.};
  double current_cost, reference_cost, candidate_cost_ev, ev_minimum_cost;
  double acc_reference = 0., acc_candidate = 0.;
  int num_nonmonotonic = 0, num_invalid = 0, iteration = 0, successful = 0;
  int termination = kTermNoConvergence;
  bool last_step_successful = false;
  double cand[3] = {x[0], x[1], x[2]};
  double model_cost_change = 0.;

  EvaluateAt<true>(J, P, xyz, x, s_part, s_tot);
  x_cost = s_tot[0];
#pragma unroll
  for (int k = 0; k < 3; ++k) g[k] = s_tot[1 + k];
#pragma unroll
  for (int k = 0; k < 6; ++k) h[k] = s_tot[4 + k];
  x_norm = Norm3(x);
  initial_cost = minimum_cost = x_cost;
  current_cost = reference_cost = candidate_cost_ev = ev_minimum_cost = x_cost;
  scale[0] = 1.0 / (1.0 + sqrt(h[0]));
  scale[1] = 1.0 / (1.0 + sqrt(h[3]));
  scale[2] = 1.0 / (1.0 + sqrt(h[5]));
  __syncthreads();   // s_tot is overwritten by the next evaluation

  while (true) {
    // ---- thread 0: close the previous iteration, stopping tests, next trial step ----
    if (threadIdx.x == 0) {
      int cmd = kCmdEvalCandidate;
      while (true) {   // (repeats only after an invalid step)
        if (last_step_successful) {
          ++successful;
          if (x_cost < minimum_cost) {
            minimum_cost = x_cost;
            best[0] = x[0];
            best[1] = x[1];
            best[2] = x[2];
          }
          last_step_successful = false;
        }
        if (iteration >= P.max_num_iterations) { termination = kTermNoConvergence; cmd = kCmdDone; break; }
        const double gmax = fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2])));
        if (gmax <= kGradientTolerance) { termination = kTermGradientTolerance; cmd = kCmdDone; break; }
        if (radius <= kMinRadius) { termination = kTermMinRadius; cmd = kCmdDone; break; }
        ++iteration;
        const double hs[6] = {h[0] * scale[0] * scale[0], h[1] * scale[0] * scale[1],
                              h[2] * scale[0] * scale[2], h[3] * scale[1] * scale[1],
                              h[4] * scale[1] * scale[2], h[5] * scale[2] * scale[2]};
        const double gs[3] = {g[0] * scale[0], g[1] * scale[1], g[2] * scale[2]};
        if (!reuse_diagonal) {
          diagonal[0] = fmin(fmax(hs[0], kMinLmDiagonal), kMaxLmDiagonal);
          diagonal[1] = fmin(fmax(hs[3], kMinLmDiagonal), kMaxLmDiagonal);
          diagonal[2] = fmin(fmax(hs[5], kMinLmDiagonal), kMaxLmDiagonal);
        }
        const double a[6] = {hs[0] + diagonal[0] / radius, hs[1], hs[2],
                             hs[3] + diagonal[1] / radius, hs[4], hs[5] + diagonal[2] / radius};
        double y[3];
        bool valid = SolveSpd3(a, gs, y);
        reuse_diagonal = true;
        double step[3] = {0., 0., 0.};
        if (valid) {
          step[0] = -y[0];
          step[1] = -y[1];
          step[2] = -y[2];
          const double hs_step[3] = {hs[0] * step[0] + hs[1] * step[1] + hs[2] * step[2],
                                     hs[1] * step[0] + hs[3] * step[1] + hs[4] * step[2],
                                     hs[2] * step[0] + hs[4] * step[1] + hs[5] * step[2]};
          model_cost_change =
              -((step[0] * gs[0] + step[1] * gs[1] + step[2] * gs[2]) +
                0.5 * (step[0] * hs_step[0] + step[1] * hs_step[1] + step[2] * hs_step[2]));
          valid = !(model_cost_change < 0.0);
        }
        if (!valid) {
          if (++num_invalid >= kMaxConsecutiveInvalidSteps) { termination = kTermInvalidSteps; cmd = kCmdDone; break; }
          radius = radius / decrease_factor;
          decrease_factor *= 2.0;
          reuse_diagonal = false;
          continue;
        }
        num_invalid = 0;
        cand[0] = x[0] + step[0] * scale[0];
        cand[1] = x[1] + step[1] * scale[1];
        cand[2] = x[2] + step[2] * scale[2];
        break;
      }
      s_pose[0] = cand[0];
      s_pose[1] = cand[1];
      s_pose[2] = cand[2];
      s_cmd = cmd;
    }
    __syncthreads();
    if (s_cmd == kCmdDone) break;
    {
      const double p[3] = {s_pose[0], s_pose[1], s_pose[2]};
      __syncthreads();
      EvaluateAt<false>(J, P, xyz, p, s_part, s_tot);   // the candidate's cost, plain doubles
    }
    // ---- thread 0: tolerances on the trial step, step quality --------------------
    if (threadIdx.x == 0) {
      const double candidate_cost = s_tot[0];
      int cmd = kCmdRejected;
      const double diff[3] = {x[0] - cand[0], x[1] - cand[1], x[2] - cand[2]};
      if (Norm3(diff) <= kParameterTolerance * (x_norm + kParameterTolerance)) {
        termination = kTermParameterTolerance;   // the step is not taken
        cmd = kCmdDone;
      } else if (fabs(x_cost - candidate_cost) <= kFunctionTolerance * x_cost) {
        termination = kTermFunctionTolerance;    // the step is not taken
        cmd = kCmdDone;
      } else {
        const double relative_decrease = (current_cost - candidate_cost) / model_cost_change;
        const double historical_decrease =
            (reference_cost - candidate_cost) / (acc_reference + model_cost_change);
        const double step_quality = fmax(relative_decrease, historical_decrease);
        if (step_quality > kMinRelativeDecrease) {
          cmd = kCmdAccept;
          x[0] = cand[0];
          x[1] = cand[1];
          x[2] = cand[2];
          x_norm = Norm3(x);
          const double t = 2.0 * step_quality - 1.0;
          radius = radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
          radius = fmin(kMaxRadius, radius);
          decrease_factor = 2.0;
          reuse_diagonal = false;
          current_cost = candidate_cost;
          acc_candidate += model_cost_change;
          acc_reference += model_cost_change;
          if (current_cost < ev_minimum_cost) {
            ev_minimum_cost = current_cost;
            num_nonmonotonic = 0;
            candidate_cost_ev = current_cost;
            acc_candidate = 0.;
          } else {
            ++num_nonmonotonic;
            if (current_cost > candidate_cost_ev) {
              candidate_cost_ev = current_cost;
              acc_candidate = 0.;
            }
          }
          if (num_nonmonotonic == max_nonmonotonic) {
            reference_cost = candidate_cost_ev;
            acc_reference = acc_candidate;
          }
        } else {
          radius = radius / decrease_factor;
          decrease_factor *= 2.0;
          reuse_diagonal = true;
        }
      }
      s_cmd = cmd;
    }
    __syncthreads();
    if (s_cmd == kCmdDone) break;
    if (s_cmd == kCmdAccept) {
      const double p[3] = {s_pose[0], s_pose[1], s_pose[2]};
      __syncthreads();
      EvaluateAt<true>(J, P, xyz, p, s_part, s_tot);   // residuals + Jacobian at the new x
      if (threadIdx.x == 0) {
        x_cost = s_tot[0];
#pragma unroll
        for (int k = 0; k < 3; ++k) g[k] = s_tot[1 + k];
#pragma unroll
        for (int k = 0; k < 6; ++k) h[k] = s_tot[4 + k];
        last_step_successful = true;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    RefResultDev out;
    out.pose[0] = best[0];
    out.pose[1] = best[1];
    out.pose[2] = best[2];
    out.initial_cost = initial_cost;
    out.final_cost = minimum_cost;
    out.iterations = iteration;
    out.num_successful_steps = successful;
    out.termination = termination;
    out.pad = 0;
    results[blockIdx.x] = out;
  }
}

// Test hook: residuals (and Jacobian rows) of one job at one pose, one thread per residual.
__global__ void k_ceres_evaluate2d(RefJobDev J, RefOpts P, const float* __restrict__ xyz,
                                   double px, double py, double pt, double cs, double sn,
                                   int with_jacobian, double* __restrict__ residuals,
                                   double* __restrict__ jacobian) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double x[3] = {px, py, pt};
  const double scaling = P.occupied_space_weight / sqrt(static_cast<double>(J.n));
  if (i < J.n) {
    const double qx = static_cast<double>(xyz[3 * static_cast<size_t>(i)]);
    const double qy = static_cast<double>(xyz[3 * static_cast<size_t>(i) + 1]);
    double res, jr[3] = {0., 0., 0.};
    if (with_jacobian) {
      PointResidual<true>(J, P, scaling, x, cs, sn, qx, qy, &res, jr);
      jacobian[3 * static_cast<size_t>(i)] = jr[0];
      jacobian[3 * static_cast<size_t>(i) + 1] = jr[1];
      jacobian[3 * static_cast<size_t>(i) + 2] = jr[2];
    } else {
      PointResidual<false>(J, P, scaling, x, cs, sn, qx, qy, &res, jr);
    }
    residuals[i] = res;
  } else if (i == J.n) {
    const size_t n = J.n;
    residuals[n] = P.translation_weight * (x[0] - J.target[0]);
    residuals[n + 1] = P.translation_weight * (x[1] - J.target[1]);
    residuals[n + 2] = P.rotation_weight * (x[2] - J.init[2]);
    if (with_jacobian) {
      for (int k = 0; k < 9; ++k) jacobian[3 * n + k] = 0.;
      jacobian[3 * n + 0] = P.translation_weight;
      jacobian[3 * (n + 1) + 1] = P.translation_weight;
      jacobian[3 * (n + 2) + 2] = P.rotation_weight;
    }
  }
}

}  // namespace csm

#ifndef CSM_REFINE_DEVICE_ONLY
using namespace csm;

namespace {

csm_status FillOpts(const csm_ceres_options2d* o, RefOpts* P) {
  CSM_REQUIRE(o != nullptr, "null options");
  // CHECK_GT(..., 0.) in ceres_scan_matcher_2d.cc:72,87,92
  CSM_REQUIRE(o->occupied_space_weight > 0. && o->translation_weight > 0. &&
              o->rotation_weight > 0., "weights must be positive");
  CSM_REQUIRE(o->max_num_iterations > 0, "max_num_iterations");   // ceres_solver_options.cc:31
  P->occupied_space_weight = o->occupied_space_weight;
  P->translation_weight = o->translation_weight;
  P->rotation_weight = o->rotation_weight;
  P->use_nonmonotonic_steps = o->use_nonmonotonic_steps != 0;
  P->max_num_iterations = o->max_num_iterations;
  // probability_values.h:64-67 and value_conversion_tables.cc:29-37 evaluated in float
  const float kMinProbability = 0.1f;
  const float kMaxProbability = 1.f - kMinProbability;
  const float kMinCost = 1.f - kMaxProbability;
  const float kMaxCost = 1.f - kMinProbability;
  P->k_scale = (kMaxCost - kMinCost) / 32766.f;
  P->cost_bias = kMinCost - P->k_scale;
  P->max_cost = kMaxCost;
  return CSM_OK;
}

void FillJob(const csm_rt_grid2d* grid, int n, long long xyz_off, const double target[2],
             const double init[3], RefJobDev* j) {
  std::memset(j, 0, sizeof(*j));
  j->cells = grid->g.cells;
  j->nx = grid->g.nx;
  j->ny = grid->g.ny;
  j->pitch = grid->g.pitch;
  j->n = n;
  j->xyz_off = xyz_off;
  j->resolution = grid->g.resolution;
  j->max_x = grid->g.max_x;
  j->max_y = grid->g.max_y;
  j->target[0] = target[0];
  j->target[1] = target[1];
  j->init[0] = init[0];
  j->init[1] = init[1];
  j->init[2] = init[2];
}

}  // namespace

extern "C" {

csm_status csm_ceres_match2d_batch(const csm_ceres_job2d* jobs, int32_t num_jobs,
                                   const csm_ceres_options2d* options,
                                   csm_ceres_result2d* results, csm_stats* stats) {
  CSM_REQUIRE(jobs && results, "null pointer");
  CSM_REQUIRE(num_jobs >= 1, "empty batch");
  RefOpts P;
  CSM_TRY(FillOpts(options, &P));
  const int device = jobs[0].grid ? jobs[0].grid->ctx->device : -1;
  long long floats = 0;
  for (int j = 0; j < num_jobs; ++j) {
    CSM_REQUIRE(jobs[j].grid != nullptr && jobs[j].xyz != nullptr, "null pointer");
    CSM_REQUIRE(jobs[j].num_points >= 1, "empty point cloud");
    CSM_REQUIRE(jobs[j].grid->d_wcells == nullptr, "the grid must be a ProbabilityGrid");
    CSM_REQUIRE(jobs[j].grid->ctx->device == device, "grids of one batch share a device");
    floats += 3LL * jobs[j].num_points;
  }
  CSM_REQUIRE(floats < (1LL << 31), "batch too large");
  LaneGuard guard;
  CSM_TRY(AcquireLane(device, &guard));
  Ctx* ctx = guard.lane;
  CSM_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t off_jobs = (static_cast<size_t>(floats) * 4 + 255) / 256 * 256;
  const size_t up_bytes = off_jobs + sizeof(RefJobDev) * num_jobs;
  PinnedBuf& up = ctx->P("ref_upload");
  DevBuf& d_up = ctx->D("ref_upload");
  DevBuf& d_out = ctx->D("ref_results");
  PinnedBuf& rb = ctx->P("ref_readback");
  CSM_TRY(up.Reserve(up_bytes));
  CSM_TRY(d_up.Reserve(up_bytes));
  CSM_TRY(d_out.Reserve(sizeof(RefResultDev) * num_jobs));
  CSM_TRY(rb.Reserve(sizeof(RefResultDev) * num_jobs));
  char* h = up.as<char>();
  RefJobDev* hj = reinterpret_cast<RefJobDev*>(h + off_jobs);
  long long off = 0;
  for (int j = 0; j < num_jobs; ++j) {
    std::memcpy(reinterpret_cast<float*>(h) + off, jobs[j].xyz,
                sizeof(float) * 3 * static_cast<size_t>(jobs[j].num_points));
    FillJob(jobs[j].grid, jobs[j].num_points, off, jobs[j].target_translation,
            jobs[j].initial_pose, &hj[j]);
    off += 3LL * jobs[j].num_points;
  }
  CSM_CUDA(cudaEventRecord(ctx->ev0, s));
  CSM_CUDA(cudaMemcpyAsync(d_up.p, h, up_bytes, cudaMemcpyHostToDevice, s));
  ProfBegin(ctx);
  k_ceres_match2d<<<num_jobs, kRefThreads, 0, s>>>(
      reinterpret_cast<const RefJobDev*>(d_up.as<char>() + off_jobs), P, d_up.as<float>(),
      d_out.as<RefResultDev>());
  CSM_LAUNCH_CHECK();
  ProfEnd(ctx, "k_ceres_match2d", static_cast<double>(num_jobs));
  CSM_CUDA(cudaEventRecord(ctx->ev1, s));
  CSM_CUDA(cudaMemcpyAsync(rb.p, d_out.p, sizeof(RefResultDev) * num_jobs,
                           cudaMemcpyDeviceToHost, s));
  CSM_CUDA(cudaStreamSynchronize(s));
  const RefResultDev* r = rb.as<RefResultDev>();
  for (int j = 0; j < num_jobs; ++j) {
    csm_ceres_result2d& o = results[j];
    std::memset(&o, 0, sizeof(o));
    std::memcpy(o.pose_estimate, r[j].pose, sizeof(double) * 3);
    o.initial_cost = r[j].initial_cost;
    o.final_cost = r[j].final_cost;
    o.iterations = r[j].iterations;
    o.num_successful_steps = r[j].num_successful_steps;
    o.termination = r[j].termination;
  }
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    stats->device_ms = ms;
    stats->host_syncs = 1;
  }
  return CSM_OK;
}

csm_status csm_ceres_evaluate2d(const csm_rt_grid2d* grid, const float* xyz, int32_t num_points,
                                const csm_ceres_options2d* options,
                                const double target_translation[2], double target_angle,
                                const double pose[3], double* residuals, double* jacobian) {
  CSM_REQUIRE(grid && xyz && target_translation && pose && residuals, "null pointer");
  CSM_REQUIRE(num_points >= 1, "empty point cloud");
  CSM_REQUIRE(grid->d_wcells == nullptr, "the grid must be a ProbabilityGrid");
  RefOpts P;
  CSM_TRY(FillOpts(options, &P));
  LaneGuard guard;
  CSM_TRY(AcquireLane(grid->ctx->device, &guard));
  Ctx* ctx = guard.lane;
  CSM_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t n = static_cast<size_t>(num_points);
  DevBuf& d_xyz = ctx->D("ref_eval_xyz");
  DevBuf& d_res = ctx->D("ref_eval_res");
  DevBuf& d_jac = ctx->D("ref_eval_jac");
  CSM_TRY(d_xyz.Reserve(sizeof(float) * 3 * n));
  CSM_TRY(d_res.Reserve(sizeof(double) * (n + 3)));
  CSM_TRY(d_jac.Reserve(sizeof(double) * 3 * (n + 3)));
  CSM_CUDA(cudaMemcpyAsync(d_xyz.p, xyz, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, s));
  RefJobDev J;
  const double init[3] = {pose[0], pose[1], target_angle};   // init[2] carries the prior's angle
  FillJob(grid, num_points, 0, target_translation, init, &J);
  // the hook takes cos / sin from the host's libm so that its values can be compared bit for
  // bit with the oracle's; the solver kernel evaluates them on the device
  k_ceres_evaluate2d<<<static_cast<int>((n + 1 + 255) / 256), 256, 0, s>>>(
      J, P, d_xyz.as<float>(), pose[0], pose[1], pose[2], std::cos(pose[2]), std::sin(pose[2]),
      jacobian != nullptr, d_res.as<double>(), d_jac.as<double>());
  CSM_LAUNCH_CHECK();
  CSM_CUDA(cudaMemcpyAsync(residuals, d_res.p, sizeof(double) * (n + 3), cudaMemcpyDeviceToHost, s));
  if (jacobian)
    CSM_CUDA(cudaMemcpyAsync(jacobian, d_jac.p, sizeof(double) * 3 * (n + 3),
                             cudaMemcpyDeviceToHost, s));
  CSM_CUDA(cudaStreamSynchronize(s));
  return CSM_OK;
}

}  // extern "C"
#endif  // CSM_REFINE_DEVICE_ONLY
