// Post-match refinement in 3D: CeresScanMatcher3D::Match
// (mapping/internal/3d/scan_matching/ceres_scan_matcher_3d.cc:95-157) as ConstraintBuilder3D
// calls it on every found match (constraint_builder_3d.cc:265-275): occupied-space blocks for
// the (high-resolution cloud, high-resolution grid) and (low-resolution cloud, low-resolution
// grid) pairs, a translation prior and a rotation prior, over {translation[3], rotation[4]}
// with Ceres' QuaternionParameterization (only_optimize_yaw = false; no intensity grids —
// the constraint builder passes none).
//
// As in refine2d.cu, Ceres is not linked: one CTA per match runs the whole trust-region loop
// on the device.  Per point: Eigen's quaternion * vector + translation and its derivative by
// the 7 ambient parameters, the 8 voxel probabilities around the point, the smoothstep
// interpolation of interpolated_grid.h:49-96 on dual numbers over (x, y, z)
// (occupied_space_cost_function_3d.h:68-78), then the residual row in the 6-dimensional
// tangent space (ambient row times the parameterisation's 4 x 3 plus-Jacobian).  The block
// reduces cost, J^T r and the upper triangle of J^T J (28 doubles); thread 0 keeps the
// minimiser's state in shared memory and does the Levenberg-Marquardt step exactly as
// refine2d.cu does, with 6 parameters, x (+) delta = {t + dt, exp(dq) * q} and the gradient
// tolerance measured as |x - (x (+) -g)|_inf.  Doubles, no FMA contraction (-fmad=false).
#include <algorithm>
#include <cmath>

// see refine2d.cu: tests/emulation includes the device code below into a CPU harness
#ifndef CSM_REFINE_DEVICE_ONLY
#include "grid3d.cuh"
#endif

namespace csm {

constexpr int kRef3Threads = 256;
constexpr int kRef3Warps = kRef3Threads / 32;
constexpr int kRef3N = 6;                              // tangent-space parameters
constexpr int kRef3H = kRef3N * (kRef3N + 1) / 2;      // upper triangle of J^T J
constexpr int kRef3Acc = 1 + kRef3N + kRef3H;          // 28
constexpr int kRef3MaxClouds = 2;

struct Ref3Cloud {
  const uint16_t* vol;     // dense box of HybridGrid values (csm_grid3d)
  int lo[3], n[3];
  float resolution, k_scale, bias, min_probability;
  int npts;
  long long xyz_off;       // first float of the cloud in the upload buffer
};

struct Ref3JobDev {
  Ref3Cloud c[kRef3MaxClouds];
  int num_clouds, pad;
  double target_t[3];
  double init[7];          // {t xyz, q wxyz}; init + 3 is also the rotation prior's target
};

struct Ref3Opts {
  double occupied_space_weight[kRef3MaxClouds];
  double translation_weight, rotation_weight;
  int use_nonmonotonic_steps, max_num_iterations;
};

struct Ref3ResultDev {
  double pose[7];
  double initial_cost, final_cost;
  int iterations, num_successful_steps, termination, pad;
};

// ---- dual numbers over (x, y, z) with ceres/jet.h's arithmetic ------------------------
struct D3 {
  double a, v0, v1, v2;
};
__device__ __forceinline__ D3 Add(const D3& f, const D3& g) {
  return D3{f.a + g.a, f.v0 + g.v0, f.v1 + g.v1, f.v2 + g.v2};
}
__device__ __forceinline__ D3 Sub(const D3& f, const D3& g) {
  return D3{f.a - g.a, f.v0 - g.v0, f.v1 - g.v1, f.v2 - g.v2};
}
__device__ __forceinline__ D3 AddS(const D3& f, double s) { return D3{f.a + s, f.v0, f.v1, f.v2}; }
__device__ __forceinline__ D3 Mul(const D3& f, const D3& g) {
  return D3{f.a * g.a, f.a * g.v0 + f.v0 * g.a, f.a * g.v1 + f.v1 * g.a, f.a * g.v2 + f.v2 * g.a};
}
__device__ __forceinline__ D3 MulS(const D3& f, double s) {
  return D3{f.a * s, f.v0 * s, f.v1 * s, f.v2 * s};
}

// HybridGrid::GetCellIndex (mapping/3d/hybrid_grid.h:428-433) of a float coordinate
__device__ __forceinline__ int CellOf(float p, float resolution) {
  return static_cast<int>(lroundf(__fdiv_rn(p, resolution)));
}

// HybridGrid::GetProbability (hybrid_grid.h:521-523) as a double
__device__ __forceinline__ double Probability(const Ref3Cloud& G, int x, int y, int z) {
  const int ix = x - G.lo[0], iy = y - G.lo[1], iz = z - G.lo[2];
  int value = 0;
  if (static_cast<unsigned>(ix) < static_cast<unsigned>(G.n[0]) &&
      static_cast<unsigned>(iy) < static_cast<unsigned>(G.n[1]) &&
      static_cast<unsigned>(iz) < static_cast<unsigned>(G.n[2]))
    value = __ldg(G.vol + (static_cast<size_t>(iz) * G.n[1] + iy) * G.n[0] + ix) & 0x7fff;
  const float prob = value == 0 ? G.min_probability
                                : __fadd_rn(__fmul_rn(__int2float_rn(value), G.k_scale), G.bias);
  return static_cast<double>(prob);
}

// InterpolatedGrid<HybridGrid>::GetInterpolatedValue (interpolated_grid.h:49-96)
template <bool kDual>
__device__ __forceinline__ D3 Interpolate(const Ref3Cloud& G, double x, double y, double z) {
  // CenterOfLowerVoxel (:115-135)
  const float res = G.resolution;
  float cx = __fmul_rn(__int2float_rn(CellOf(static_cast<float>(x), res)), res);
  float cy = __fmul_rn(__int2float_rn(CellOf(static_cast<float>(y), res)), res);
  float cz = __fmul_rn(__int2float_rn(CellOf(static_cast<float>(z), res)), res);
  if (static_cast<double>(cx) > x) cx = __fsub_rn(cx, res);
  if (static_cast<double>(cy) > y) cy = __fsub_rn(cy, res);
  if (static_cast<double>(cz) > z) cz = __fsub_rn(cz, res);
  const double x1 = cx, y1 = cy, z1 = cz;
  const double x2 = __fadd_rn(cx, res), y2 = __fadd_rn(cy, res), z2 = __fadd_rn(cz, res);
  const int i = CellOf(cx, res), j = CellOf(cy, res), k = CellOf(cz, res);
  const double q111 = Probability(G, i, j, k), q112 = Probability(G, i, j, k + 1);
  const double q121 = Probability(G, i, j + 1, k), q122 = Probability(G, i, j + 1, k + 1);
  const double q211 = Probability(G, i + 1, j, k), q212 = Probability(G, i + 1, j, k + 1);
  const double q221 = Probability(G, i + 1, j + 1, k), q222 = Probability(G, i + 1, j + 1, k + 1);
  D3 nx, ny, nz;
  if (kDual) {
    const double ix = 1.0 / (x2 - x1), iy = 1.0 / (y2 - y1), iz = 1.0 / (z2 - z1);
    nx = D3{(x - x1) * ix, 1.0 * ix, 0.0 * ix, 0.0 * ix};
    ny = D3{(y - y1) * iy, 0.0 * iy, 1.0 * iy, 0.0 * iy};
    nz = D3{(z - z1) * iz, 0.0 * iz, 0.0 * iz, 1.0 * iz};
  } else {
    nx = D3{(x - x1) / (x2 - x1), 0., 0., 0.};
    ny = D3{(y - y1) / (y2 - y1), 0., 0., 0.};
    nz = D3{(z - z1) / (z2 - z1), 0., 0., 0.};
  }
  const D3 nxx = Mul(nx, nx), nxxx = Mul(nx, nxx);
  const D3 nyy = Mul(ny, ny), nyyy = Mul(ny, nyy);
  const D3 nzz = Mul(nz, nz), nzzz = Mul(nz, nzz);
  // (qa - qb) * n^3 * 2. + (qb - qa) * n^2 * 3. + qa, scalars first
  const D3 q11 = AddS(Add(MulS(MulS(nzzz, q111 - q112), 2.), MulS(MulS(nzz, q112 - q111), 3.)), q111);
  const D3 q12 = AddS(Add(MulS(MulS(nzzz, q121 - q122), 2.), MulS(MulS(nzz, q122 - q121), 3.)), q121);
  const D3 q21 = AddS(Add(MulS(MulS(nzzz, q211 - q212), 2.), MulS(MulS(nzz, q212 - q211), 3.)), q211);
  const D3 q22 = AddS(Add(MulS(MulS(nzzz, q221 - q222), 2.), MulS(MulS(nzz, q222 - q221), 3.)), q221);
  const D3 q1 = Add(Add(MulS(Mul(Sub(q11, q12), nyyy), 2.), MulS(Mul(Sub(q12, q11), nyy), 3.)), q11);
  const D3 q2 = Add(Add(MulS(Mul(Sub(q21, q22), nyyy), 2.), MulS(Mul(Sub(q22, q21), nyy), 3.)), q21);
  return Add(Add(MulS(Mul(Sub(q1, q2), nxxx), 2.), MulS(Mul(Sub(q2, q1), nxx), 3.)), q1);
}

__device__ __forceinline__ void Cross3(const double* a, const double* b, double* out) {
  out[0] = a[1] * b[2] - a[2] * b[1];
  out[1] = a[2] * b[0] - a[0] * b[2];
  out[2] = a[0] * b[1] - a[1] * b[0];
}

// ceres::QuaternionParameterization::ComputeJacobian (4 x 3, row-major)
__device__ __forceinline__ void PlusJacobian(const double* q, double* jac) {
  jac[0] = -q[1]; jac[1] = -q[2]; jac[2] = -q[3];
  jac[3] = q[0];  jac[4] = q[3];  jac[5] = -q[2];
  jac[6] = -q[3]; jac[7] = q[0];  jac[8] = q[1];
  jac[9] = q[2];  jac[10] = -q[1]; jac[11] = q[0];
}

__device__ __forceinline__ void ToLocal(const double* ambient, const double* pj, double* row) {
  row[0] = ambient[0];
  row[1] = ambient[1];
  row[2] = ambient[2];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    row[3 + k] = ambient[3] * pj[k] + ambient[4] * pj[3 + k] + ambient[5] * pj[6 + k] +
                 ambient[6] * pj[9 + k];
}

// x (+) delta (QuaternionParameterization::Plus on the rotation block)
__device__ __forceinline__ void Plus7(const double* x, const double* delta, double* out) {
  for (int k = 0; k < 3; ++k) out[k] = x[k] + delta[k];
  const double* d = delta + 3;
  const double norm_delta = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (norm_delta > 0.0) {
    const double sin_delta_by_delta = sin(norm_delta) / norm_delta;
    const double z[4] = {cos(norm_delta), sin_delta_by_delta * d[0], sin_delta_by_delta * d[1],
                         sin_delta_by_delta * d[2]};
    const double* w = x + 3;
    out[3] = z[0] * w[0] - z[1] * w[1] - z[2] * w[2] - z[3] * w[3];
    out[4] = z[0] * w[1] + z[1] * w[0] + z[2] * w[3] - z[3] * w[2];
    out[5] = z[0] * w[2] - z[1] * w[3] + z[2] * w[0] + z[3] * w[1];
    out[6] = z[0] * w[3] + z[1] * w[2] - z[2] * w[1] + z[3] * w[0];
  } else {
    for (int k = 3; k < 7; ++k) out[k] = x[k];
  }
}

// One occupied-space residual and (kJac) its tangent-space Jacobian row.
template <bool kJac>
__device__ __forceinline__ void PointResidual3(const Ref3Cloud& G, double scaling,
                                               const double* pose, const double* pj, double px,
                                               double py, double pz, double* res, double* row) {
  // Eigen QuaternionBase::_transformVector (q not normalised) + translation
  const double p[3] = {px, py, pz};
  const double w = pose[3];
  const double qv[3] = {pose[4], pose[5], pose[6]};
  double uv[3], c[3], world[3];
  Cross3(qv, p, uv);
  uv[0] += uv[0];
  uv[1] += uv[1];
  uv[2] += uv[2];
  Cross3(qv, uv, c);
#pragma unroll
  for (int k = 0; k < 3; ++k) world[k] = ((p[k] + w * uv[k]) + c[k]) + pose[k];
  const D3 f = Interpolate<kJac>(G, world[0], world[1], world[2]);
  *res = scaling * (1. - f.a);
  if (!kJac) return;
  double ambient[7];
  // d world / d t = I
  ambient[0] = scaling * (-((f.v0 * 1.0 + f.v1 * 0.0) + f.v2 * 0.0));
  ambient[1] = scaling * (-((f.v0 * 0.0 + f.v1 * 1.0) + f.v2 * 0.0));
  ambient[2] = scaling * (-((f.v0 * 0.0 + f.v1 * 0.0) + f.v2 * 1.0));
  // d world / d w = uv
  ambient[3] = scaling * (-((f.v0 * uv[0] + f.v1 * uv[1]) + f.v2 * uv[2]));
#pragma unroll
  for (int k = 0; k < 3; ++k) {   // d world / d qv[k] = w * duv + e_k x uv + qv x duv
    double e[3] = {0., 0., 0.};
    e[k] = 1.0;
    double duv[3], t1[3], t2[3];
    Cross3(e, p, duv);
    duv[0] += duv[0];
    duv[1] += duv[1];
    duv[2] += duv[2];
    Cross3(e, uv, t1);
    Cross3(qv, duv, t2);
    const double d0 = (w * duv[0] + t1[0]) + t2[0];
    const double d1 = (w * duv[1] + t1[1]) + t2[1];
    const double d2 = (w * duv[2] + t1[2]) + t2[2];
    ambient[4 + k] = scaling * (-((f.v0 * d0 + f.v1 * d1) + f.v2 * d2));
  }
  ToLocal(ambient, pj, row);
}

// The 3 + 3 prior residuals and their tangent-space rows (thread 0).
__device__ __forceinline__ void PriorRows(const Ref3JobDev& J, const Ref3Opts& P,
                                          const double* pose, const double* pj, double* res,
                                          double (*rows)[kRef3N]) {
  for (int k = 0; k < 3; ++k) {
    res[k] = P.translation_weight * (pose[k] - J.target_t[k]);
    for (int a = 0; a < kRef3N; ++a) rows[k][a] = 0.;
    rows[k][k] = P.translation_weight;
  }
  const double z[4] = {J.init[3], -J.init[4], -J.init[5], -J.init[6]};
  const double* w = pose + 3;
  const double delta[3] = {z[0] * w[1] + z[1] * w[0] + z[2] * w[3] - z[3] * w[2],
                           z[0] * w[2] - z[1] * w[3] + z[2] * w[0] + z[3] * w[1],
                           z[0] * w[3] + z[1] * w[2] - z[2] * w[1] + z[3] * w[0]};
  const double ddelta[3][4] = {{z[1], z[0], -z[3], z[2]},
                               {z[2], z[3], z[0], -z[1]},
                               {z[3], -z[2], z[1], z[0]}};
  for (int k = 0; k < 3; ++k) {
    res[3 + k] = P.rotation_weight * delta[k];
    double ambient[7] = {0., 0., 0., 0., 0., 0., 0.};
    for (int c = 0; c < 4; ++c) ambient[3 + c] = P.rotation_weight * ddelta[k][c];
    ToLocal(ambient, pj, rows[3 + k]);
  }
}

__device__ __forceinline__ int Tri6(int i, int j) {   // i <= j
  return i * kRef3N - i * (i - 1) / 2 + (j - i);
}

template <int kCount>
__device__ __forceinline__ void BlockSum3(double* acc, double (*s_part)[kRef3Acc],
                                          double* s_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kCount; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_part[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kCount) {
    double v = 0.;
#pragma unroll
    for (int w = 0; w < kRef3Warps; ++w) v += s_part[w][threadIdx.x];
    s_tot[threadIdx.x] = v;
  }
  __syncthreads();
}

// Block-wide evaluation at `pose`: s_tot = {cost, g[6], h[21]} (g, h only with kJac).
template <bool kJac>
__device__ __forceinline__ void EvaluateAt3(const Ref3JobDev& J, const Ref3Opts& P,
                                            const float* __restrict__ cloud, const double* pose,
                                            double (*s_part)[kRef3Acc], double* s_tot) {
  double pj[12];
  PlusJacobian(pose + 3, pj);
  double acc[kRef3Acc];
#pragma unroll
  for (int k = 0; k < kRef3Acc; ++k) acc[k] = 0.;
  for (int b = 0; b < J.num_clouds; ++b) {
    const Ref3Cloud& G = J.c[b];
    const float* __restrict__ xyz = cloud + G.xyz_off;
    const double scaling = P.occupied_space_weight[b] / sqrt(static_cast<double>(G.npts));
    for (int i = threadIdx.x; i < G.npts; i += kRef3Threads) {
      const double px = static_cast<double>(xyz[3 * static_cast<size_t>(i)]);
      const double py = static_cast<double>(xyz[3 * static_cast<size_t>(i) + 1]);
      const double pz = static_cast<double>(xyz[3 * static_cast<size_t>(i) + 2]);
      double res, row[kRef3N];
      PointResidual3<kJac>(G, scaling, pose, pj, px, py, pz, &res, row);
      acc[0] += res * res;
      if (kJac) {
#pragma unroll
        for (int a = 0; a < kRef3N; ++a) {
          acc[1 + a] += row[a] * res;
#pragma unroll
          for (int c = a; c < kRef3N; ++c) acc[1 + kRef3N + Tri6(a, c)] += row[a] * row[c];
        }
      }
    }
  }
  BlockSum3<kJac ? kRef3Acc : 1>(acc, s_part, s_tot);
  if (threadIdx.x == 0) {
    double res[6], rows[6][kRef3N];
    PriorRows(J, P, pose, pj, res, rows);
    double sq = s_tot[0];
    for (int k = 0; k < 6; ++k) sq += res[k] * res[k];
    s_tot[0] = 0.5 * sq;
    if (kJac) {
      for (int k = 0; k < 6; ++k)
        for (int a = 0; a < kRef3N; ++a) {
          s_tot[1 + a] += rows[k][a] * res[k];
          for (int c = a; c < kRef3N; ++c) s_tot[1 + kRef3N + Tri6(a, c)] += rows[k][a] * rows[k][c];
        }
    }
  }
  __syncthreads();
}

// Cholesky solve of A y = b, A symmetric positive definite (upper triangle, row-major)
__device__ __forceinline__ bool SolveSpd6(const double* a, const double* b, double* y) {
  double l[kRef3N][kRef3N];
  for (int i = 0; i < kRef3N; ++i)
    for (int j = 0; j < kRef3N; ++j) l[i][j] = 0.;
  for (int i = 0; i < kRef3N; ++i) {
    for (int j = 0; j <= i; ++j) {
      double s = a[Tri6(j, i)];
      for (int k = 0; k < j; ++k) s -= l[i][k] * l[j][k];
      if (i == j) {
        if (!(s > 0.)) return false;
        l[i][i] = sqrt(s);
      } else {
        l[i][j] = s / l[j][j];
      }
    }
  }
  double z[kRef3N];
  for (int i = 0; i < kRef3N; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= l[i][k] * z[k];
    z[i] = s / l[i][i];
  }
  for (int i = kRef3N - 1; i >= 0; --i) {
    double s = z[i];
    for (int k = i + 1; k < kRef3N; ++k) s -= l[k][i] * y[k];
    y[i] = s / l[i][i];
  }
  for (int i = 0; i < kRef3N; ++i)
    if (!isfinite(y[i])) return false;
  return true;
}

__device__ __forceinline__ double NormN(const double* v, int n) {
  double s = 0.;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return sqrt(s);
}

// State of the minimiser; lives in shared memory, touched by thread 0 only.
struct Solver3 {
  double x[7], best[7], cand[7];
  double g[kRef3N], h[kRef3H], scale[kRef3N], diagonal[kRef3N];
  double x_cost, x_norm, minimum_cost, initial_cost;
  double radius, decrease_factor;
  double current_cost, reference_cost, candidate_cost_ev, ev_minimum_cost;
  double acc_reference, acc_candidate, model_cost_change;
  int reuse_diagonal, last_step_successful;
  int num_nonmonotonic, num_invalid, iteration, successful, termination;
};

enum { kCmd3EvalCandidate = 0, kCmd3Accept = 1, kCmd3Rejected = 2, kCmd3Done = 3 };
enum {
  kTerm3NoConvergence = 0, kTerm3FunctionTolerance = 1, kTerm3GradientTolerance = 2,
  kTerm3ParameterTolerance = 3, kTerm3MinRadius = 4, kTerm3InvalidSteps = 5
};

__global__ void __launch_bounds__(kRef3Threads)
k_ceres_match3d(const Ref3JobDev* __restrict__ jobs, Ref3Opts P, const float* __restrict__ cloud,
                Ref3ResultDev* __restrict__ results) {
  __shared__ double s_part[kRef3Warps][kRef3Acc];
  __shared__ double s_tot[kRef3Acc];
  __shared__ double s_pose[7];
  __shared__ int s_cmd;
  __shared__ Solver3 S;
  __shared__ Ref3JobDev J;
  if (threadIdx.x == 0) J = jobs[blockIdx.x];
  __syncthreads();

  const double kInitialRadius = 1e4, kMaxRadius = 1e16, kMinRadius = 1e-32;
  const double kMinRelativeDecrease = 1e-3, kMinLmDiagonal = 1e-6, kMaxLmDiagonal = 1e32;
  const int kMaxConsecutiveInvalidSteps = 5;
  const double kFunctionTolerance = 1e-6, kGradientTolerance = 1e-10, kParameterTolerance = 1e-8;
  const int max_nonmonotonic = P.use_nonmonotonic_steps ? 5 : 0;

  {
    double p[7];
    for (int k = 0; k < 7; ++k) p[k] = J.init[k];
    EvaluateAt3<true>(J, P, cloud, p, s_part, s_tot);
  }
  if (threadIdx.x == 0) {
    for (int k = 0; k < 7; ++k) S.x[k] = S.best[k] = S.cand[k] = J.init[k];
    S.x_cost = s_tot[0];
    for (int a = 0; a < kRef3N; ++a) S.g[a] = s_tot[1 + a];
    for (int a = 0; a < kRef3H; ++a) S.h[a] = s_tot[1 + kRef3N + a];
    S.x_norm = NormN(S.x, 7);
    S.initial_cost = S.minimum_cost = S.x_cost;
    S.current_cost = S.reference_cost = S.candidate_cost_ev = S.ev_minimum_cost = S.x_cost;
    for (int a = 0; a < kRef3N; ++a) {
      S.scale[a] = 1.0 / (1.0 + sqrt(S.h[Tri6(a, a)]));
      S.diagonal[a] = 0.;
    }
    S.radius = kInitialRadius;
    S.decrease_factor = 2.0;
    S.reuse_diagonal = 0;
    S.last_step_successful = 0;
    S.acc_reference = S.acc_candidate = S.model_cost_change = 0.;
    S.num_nonmonotonic = S.num_invalid = S.iteration = S.successful = 0;
    S.termination = kTerm3NoConvergence;
  }
  __syncthreads();

  while (true) {
    if (threadIdx.x == 0) {
      int cmd = kCmd3EvalCandidate;
      while (true) {   // (repeats only after an invalid step)
        if (S.last_step_successful) {
          ++S.successful;
          if (S.x_cost < S.minimum_cost) {
            S.minimum_cost = S.x_cost;
            for (int k = 0; k < 7; ++k) S.best[k] = S.x[k];
          }
          S.last_step_successful = 0;
        }
        if (S.iteration >= P.max_num_iterations) { S.termination = kTerm3NoConvergence; cmd = kCmd3Done; break; }
        {
          double neg[kRef3N], moved[7], gmax = 0.;
          for (int a = 0; a < kRef3N; ++a) neg[a] = -S.g[a];
          Plus7(S.x, neg, moved);
          for (int k = 0; k < 7; ++k) gmax = fmax(gmax, fabs(S.x[k] - moved[k]));
          if (gmax <= kGradientTolerance) { S.termination = kTerm3GradientTolerance; cmd = kCmd3Done; break; }
        }
        if (S.radius <= kMinRadius) { S.termination = kTerm3MinRadius; cmd = kCmd3Done; break; }
        ++S.iteration;
        double hs[kRef3H], gs[kRef3N];
        for (int a = 0; a < kRef3N; ++a) {
          gs[a] = S.g[a] * S.scale[a];
          for (int c = a; c < kRef3N; ++c) hs[Tri6(a, c)] = S.h[Tri6(a, c)] * S.scale[a] * S.scale[c];
        }
        if (!S.reuse_diagonal)
          for (int a = 0; a < kRef3N; ++a)
            S.diagonal[a] = fmin(fmax(hs[Tri6(a, a)], kMinLmDiagonal), kMaxLmDiagonal);
        double am[kRef3H];
        for (int i = 0; i < kRef3H; ++i) am[i] = hs[i];
        for (int a = 0; a < kRef3N; ++a) am[Tri6(a, a)] = hs[Tri6(a, a)] + S.diagonal[a] / S.radius;
        double y[kRef3N];
        bool valid = SolveSpd6(am, gs, y);
        S.reuse_diagonal = 1;
        double step[kRef3N];
        for (int a = 0; a < kRef3N; ++a) step[a] = 0.;
        if (valid) {
          for (int a = 0; a < kRef3N; ++a) step[a] = -y[a];
          double sg = 0., shs = 0.;
          for (int a = 0; a < kRef3N; ++a) {
            sg += step[a] * gs[a];
            double row = 0.;
            for (int c = 0; c < kRef3N; ++c) row += hs[a <= c ? Tri6(a, c) : Tri6(c, a)] * step[c];
            shs += step[a] * row;
          }
          S.model_cost_change = -(sg + 0.5 * shs);
          valid = !(S.model_cost_change < 0.0);
        }
        if (!valid) {
          if (++S.num_invalid >= kMaxConsecutiveInvalidSteps) { S.termination = kTerm3InvalidSteps; cmd = kCmd3Done; break; }
          S.radius = S.radius / S.decrease_factor;
          S.decrease_factor *= 2.0;
          S.reuse_diagonal = 0;
          continue;
        }
        S.num_invalid = 0;
        double delta[kRef3N];
        for (int a = 0; a < kRef3N; ++a) delta[a] = step[a] * S.scale[a];
        Plus7(S.x, delta, S.cand);
        break;
      }
      for (int k = 0; k < 7; ++k) s_pose[k] = S.cand[k];
      s_cmd = cmd;
    }
    __syncthreads();
    if (s_cmd == kCmd3Done) break;
    {
      double p[7];
      for (int k = 0; k < 7; ++k) p[k] = s_pose[k];
      __syncthreads();
      EvaluateAt3<false>(J, P, cloud, p, s_part, s_tot);
    }
    if (threadIdx.x == 0) {
      const double candidate_cost = s_tot[0];
      int cmd = kCmd3Rejected;
      double diff[7];
      for (int k = 0; k < 7; ++k) diff[k] = S.x[k] - S.cand[k];
      if (NormN(diff, 7) <= kParameterTolerance * (S.x_norm + kParameterTolerance)) {
        S.termination = kTerm3ParameterTolerance;
        cmd = kCmd3Done;
      } else if (fabs(S.x_cost - candidate_cost) <= kFunctionTolerance * S.x_cost) {
        S.termination = kTerm3FunctionTolerance;
        cmd = kCmd3Done;
      } else {
        const double relative_decrease = (S.current_cost - candidate_cost) / S.model_cost_change;
        const double historical_decrease =
            (S.reference_cost - candidate_cost) / (S.acc_reference + S.model_cost_change);
        const double step_quality = fmax(relative_decrease, historical_decrease);
        if (step_quality > kMinRelativeDecrease) {
          cmd = kCmd3Accept;
          for (int k = 0; k < 7; ++k) S.x[k] = S.cand[k];
          S.x_norm = NormN(S.x, 7);
          const double t = 2.0 * step_quality - 1.0;
          S.radius = S.radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
          S.radius = fmin(kMaxRadius, S.radius);
          S.decrease_factor = 2.0;
          S.reuse_diagonal = 0;
          S.current_cost = candidate_cost;
          S.acc_candidate += S.model_cost_change;
          S.acc_reference += S.model_cost_change;
          if (S.current_cost < S.ev_minimum_cost) {
            S.ev_minimum_cost = S.current_cost;
            S.num_nonmonotonic = 0;
            S.candidate_cost_ev = S.current_cost;
            S.acc_candidate = 0.;
          } else {
            ++S.num_nonmonotonic;
            if (S.current_cost > S.candidate_cost_ev) {
              S.candidate_cost_ev = S.current_cost;
              S.acc_candidate = 0.;
            }
          }
          if (S.num_nonmonotonic == max_nonmonotonic) {
            S.reference_cost = S.candidate_cost_ev;
            S.acc_reference = S.acc_candidate;
          }
        } else {
          S.radius = S.radius / S.decrease_factor;
          S.decrease_factor *= 2.0;
          S.reuse_diagonal = 1;
        }
      }
      s_cmd = cmd;
    }
    __syncthreads();
    if (s_cmd == kCmd3Done) break;
    if (s_cmd == kCmd3Accept) {
      double p[7];
      for (int k = 0; k < 7; ++k) p[k] = s_pose[k];
      __syncthreads();
      EvaluateAt3<true>(J, P, cloud, p, s_part, s_tot);
      if (threadIdx.x == 0) {
        S.x_cost = s_tot[0];
        for (int a = 0; a < kRef3N; ++a) S.g[a] = s_tot[1 + a];
        for (int a = 0; a < kRef3H; ++a) S.h[a] = s_tot[1 + kRef3N + a];
        S.last_step_successful = 1;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    Ref3ResultDev out;
    for (int k = 0; k < 7; ++k) out.pose[k] = S.best[k];
    out.initial_cost = S.initial_cost;
    out.final_cost = S.minimum_cost;
    out.iterations = S.iteration;
    out.num_successful_steps = S.successful;
    out.termination = S.termination;
    out.pad = 0;
    results[blockIdx.x] = out;
  }
}

// Test hook: residuals (and tangent-space Jacobian rows) of one job at one pose.
__global__ void k_ceres_evaluate3d(const Ref3JobDev* __restrict__ job, Ref3Opts P,
                                   const float* __restrict__ cloud, const double* __restrict__ pose7,
                                   int with_jacobian, double* __restrict__ residuals,
                                   double* __restrict__ jacobian) {
  const Ref3JobDev& J = *job;
  double pose[7], pj[12];
  for (int k = 0; k < 7; ++k) pose[k] = pose7[k];
  PlusJacobian(pose + 3, pj);
  int total = 0;
  for (int b = 0; b < J.num_clouds; ++b) total += J.c[b].npts;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    int b = 0, local = i;
    while (local >= J.c[b].npts) {
      local -= J.c[b].npts;
      ++b;
    }
    const Ref3Cloud& G = J.c[b];
    const float* xyz = cloud + G.xyz_off;
    const double scaling = P.occupied_space_weight[b] / sqrt(static_cast<double>(G.npts));
    const double px = static_cast<double>(xyz[3 * static_cast<size_t>(local)]);
    const double py = static_cast<double>(xyz[3 * static_cast<size_t>(local) + 1]);
    const double pz = static_cast<double>(xyz[3 * static_cast<size_t>(local) + 2]);
    double res, row[kRef3N];
    if (with_jacobian) {
      PointResidual3<true>(G, scaling, pose, pj, px, py, pz, &res, row);
      for (int a = 0; a < kRef3N; ++a) jacobian[kRef3N * static_cast<size_t>(i) + a] = row[a];
    } else {
      PointResidual3<false>(G, scaling, pose, pj, px, py, pz, &res, row);
    }
    residuals[i] = res;
  } else if (i == total) {
    double res[6], rows[6][kRef3N];
    PriorRows(J, P, pose, pj, res, rows);
    for (int k = 0; k < 6; ++k) {
      residuals[total + k] = res[k];
      if (with_jacobian)
        for (int a = 0; a < kRef3N; ++a)
          jacobian[kRef3N * static_cast<size_t>(total + k) + a] = rows[k][a];
    }
  }
}

}  // namespace csm

#ifndef CSM_REFINE_DEVICE_ONLY
using namespace csm;

namespace {

csm_status FillOpts3(const csm_ceres_options3d* o, int max_clouds, Ref3Opts* P) {
  CSM_REQUIRE(o != nullptr, "null options");
  // CHECK_GT(..., 0.) in ceres_scan_matcher_3d.cc:114,143,148
  for (int b = 0; b < max_clouds; ++b)
    CSM_REQUIRE(o->occupied_space_weight[b] > 0., "occupied_space_weight must be positive");
  CSM_REQUIRE(o->translation_weight > 0. && o->rotation_weight > 0., "weights must be positive");
  CSM_REQUIRE(o->max_num_iterations > 0, "max_num_iterations");
  CSM_REQUIRE(o->only_optimize_yaw == 0,
              "only_optimize_yaw (YawOnlyQuaternionPlus) is not supported");
  for (int b = 0; b < kRef3MaxClouds; ++b) P->occupied_space_weight[b] = o->occupied_space_weight[b];
  P->translation_weight = o->translation_weight;
  P->rotation_weight = o->rotation_weight;
  P->use_nonmonotonic_steps = o->use_nonmonotonic_steps != 0;
  P->max_num_iterations = o->max_num_iterations;
  return CSM_OK;
}

csm_status CheckJob(const csm_ceres_job3d& j, int device) {
  CSM_REQUIRE(j.num_clouds >= 1 && j.num_clouds <= kRef3MaxClouds, "1 or 2 point clouds");
  for (int b = 0; b < j.num_clouds; ++b) {
    CSM_REQUIRE(j.grid[b] != nullptr && j.xyz[b] != nullptr, "null pointer");
    CSM_REQUIRE(j.num_points[b] >= 1, "empty point cloud");
    CSM_REQUIRE(j.grid[b]->ctx->device == device, "grids of one batch share a device");
  }
  return CSM_OK;
}

void FillJob3(const csm_ceres_job3d& j, long long* float_off, Ref3JobDev* d) {
  std::memset(d, 0, sizeof(*d));
  d->num_clouds = j.num_clouds;
  for (int b = 0; b < j.num_clouds; ++b) {
    const Grid3Dev& g = j.grid[b]->g;
    Ref3Cloud& c = d->c[b];
    c.vol = g.p;
    for (int a = 0; a < 3; ++a) {
      c.lo[a] = g.lo[a];
      c.n[a] = g.n[a];
    }
    c.resolution = g.resolution;
    c.k_scale = g.k_scale;
    c.bias = g.bias;
    c.min_probability = g.min_probability;
    c.npts = j.num_points[b];
    c.xyz_off = *float_off;
    *float_off += 3LL * j.num_points[b];
  }
  for (int k = 0; k < 3; ++k) d->target_t[k] = j.target_translation[k];
  for (int k = 0; k < 7; ++k) d->init[k] = j.initial_pose[k];
}

}  // namespace

extern "C" {

csm_status csm_ceres_match3d_batch(const csm_ceres_job3d* jobs, int32_t num_jobs,
                                   const csm_ceres_options3d* options,
                                   csm_ceres_result3d* results, csm_stats* stats) {
  CSM_REQUIRE(jobs && results, "null pointer");
  CSM_REQUIRE(num_jobs >= 1, "empty batch");
  CSM_REQUIRE(jobs[0].num_clouds >= 1 && jobs[0].grid[0] != nullptr, "null pointer");
  const int device = jobs[0].grid[0]->ctx->device;
  int max_clouds = 0;
  long long floats = 0;
  for (int j = 0; j < num_jobs; ++j) {
    CSM_TRY(CheckJob(jobs[j], device));
    max_clouds = std::max(max_clouds, jobs[j].num_clouds);
    for (int b = 0; b < jobs[j].num_clouds; ++b) floats += 3LL * jobs[j].num_points[b];
  }
  CSM_REQUIRE(floats < (1LL << 31), "batch too large");
  Ref3Opts P;
  CSM_TRY(FillOpts3(options, max_clouds, &P));
  LaneGuard guard;
  CSM_TRY(AcquireLane(device, &guard));
  Ctx* ctx = guard.lane;
  CSM_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t off_jobs = (static_cast<size_t>(floats) * 4 + 255) / 256 * 256;
  const size_t up_bytes = off_jobs + sizeof(Ref3JobDev) * num_jobs;
  PinnedBuf& up = ctx->P("ref3_upload");
  DevBuf& d_up = ctx->D("ref3_upload");
  DevBuf& d_out = ctx->D("ref3_results");
  PinnedBuf& rb = ctx->P("ref3_readback");
  CSM_TRY(up.Reserve(up_bytes));
  CSM_TRY(d_up.Reserve(up_bytes));
  CSM_TRY(d_out.Reserve(sizeof(Ref3ResultDev) * num_jobs));
  CSM_TRY(rb.Reserve(sizeof(Ref3ResultDev) * num_jobs));
  char* h = up.as<char>();
  Ref3JobDev* hj = reinterpret_cast<Ref3JobDev*>(h + off_jobs);
  long long off = 0;
  for (int j = 0; j < num_jobs; ++j) {
    for (int b = 0; b < jobs[j].num_clouds; ++b)
      std::memcpy(reinterpret_cast<float*>(h) + off + (b == 0 ? 0 : 3LL * jobs[j].num_points[0]),
                  jobs[j].xyz[b], sizeof(float) * 3 * static_cast<size_t>(jobs[j].num_points[b]));
    FillJob3(jobs[j], &off, &hj[j]);
  }
  CSM_CUDA(cudaEventRecord(ctx->ev0, s));
  CSM_CUDA(cudaMemcpyAsync(d_up.p, h, up_bytes, cudaMemcpyHostToDevice, s));
  ProfBegin(ctx);
  k_ceres_match3d<<<num_jobs, kRef3Threads, 0, s>>>(
      reinterpret_cast<const Ref3JobDev*>(d_up.as<char>() + off_jobs), P, d_up.as<float>(),
      d_out.as<Ref3ResultDev>());
  CSM_LAUNCH_CHECK();
  ProfEnd(ctx, "k_ceres_match3d", static_cast<double>(num_jobs));
  CSM_CUDA(cudaEventRecord(ctx->ev1, s));
  CSM_CUDA(cudaMemcpyAsync(rb.p, d_out.p, sizeof(Ref3ResultDev) * num_jobs,
                           cudaMemcpyDeviceToHost, s));
  CSM_CUDA(cudaStreamSynchronize(s));
  const Ref3ResultDev* r = rb.as<Ref3ResultDev>();
  for (int j = 0; j < num_jobs; ++j) {
    csm_ceres_result3d& o = results[j];
    std::memset(&o, 0, sizeof(o));
    std::memcpy(o.pose_estimate, r[j].pose, sizeof(double) * 7);
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

csm_status csm_ceres_evaluate3d(const csm_ceres_job3d* job, const csm_ceres_options3d* options,
                                const double pose[7], double* residuals, double* jacobian) {
  CSM_REQUIRE(job && pose && residuals, "null pointer");
  CSM_REQUIRE(job->num_clouds >= 1 && job->grid[0] != nullptr, "null pointer");
  const int device = job->grid[0]->ctx->device;
  CSM_TRY(CheckJob(*job, device));
  Ref3Opts P;
  CSM_TRY(FillOpts3(options, job->num_clouds, &P));
  LaneGuard guard;
  CSM_TRY(AcquireLane(device, &guard));
  Ctx* ctx = guard.lane;
  CSM_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  long long floats = 0;
  for (int b = 0; b < job->num_clouds; ++b) floats += 3LL * job->num_points[b];
  const size_t rows = static_cast<size_t>(floats / 3) + 6;
  const size_t off_job = (static_cast<size_t>(floats) * 4 + 255) / 256 * 256;
  const size_t off_pose = off_job + (sizeof(Ref3JobDev) + 255) / 256 * 256;
  const size_t up_bytes = off_pose + 64;
  PinnedBuf& up = ctx->P("ref3_eval_upload");
  DevBuf& d_up = ctx->D("ref3_eval_upload");
  DevBuf& d_res = ctx->D("ref3_eval_res");
  DevBuf& d_jac = ctx->D("ref3_eval_jac");
  CSM_TRY(up.Reserve(up_bytes));
  CSM_TRY(d_up.Reserve(up_bytes));
  CSM_TRY(d_res.Reserve(sizeof(double) * rows));
  CSM_TRY(d_jac.Reserve(sizeof(double) * kRef3N * rows));
  char* h = up.as<char>();
  long long off = 0;
  for (int b = 0; b < job->num_clouds; ++b)
    std::memcpy(reinterpret_cast<float*>(h) + (b == 0 ? 0 : 3LL * job->num_points[0]), job->xyz[b],
                sizeof(float) * 3 * static_cast<size_t>(job->num_points[b]));
  FillJob3(*job, &off, reinterpret_cast<Ref3JobDev*>(h + off_job));
  std::memcpy(h + off_pose, pose, sizeof(double) * 7);
  CSM_CUDA(cudaMemcpyAsync(d_up.p, h, up_bytes, cudaMemcpyHostToDevice, s));
  k_ceres_evaluate3d<<<static_cast<int>((rows + 255) / 256), 256, 0, s>>>(
      reinterpret_cast<const Ref3JobDev*>(d_up.as<char>() + off_job), P, d_up.as<float>(),
      reinterpret_cast<const double*>(d_up.as<char>() + off_pose), jacobian != nullptr,
      d_res.as<double>(), d_jac.as<double>());
  CSM_LAUNCH_CHECK();
  CSM_CUDA(cudaMemcpyAsync(residuals, d_res.p, sizeof(double) * rows, cudaMemcpyDeviceToHost, s));
  if (jacobian)
    CSM_CUDA(cudaMemcpyAsync(jacobian, d_jac.p, sizeof(double) * kRef3N * rows,
                             cudaMemcpyDeviceToHost, s));
  CSM_CUDA(cudaStreamSynchronize(s));
  return CSM_OK;
}

}  // extern "C"
#endif  // CSM_REFINE_DEVICE_ONLY
