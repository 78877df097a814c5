// RealTimeCorrelativeScanMatcher2D::Match on a ProbabilityGrid
// (cartographer/mapping/internal/2d/scan_matching/real_time_correlative_scan_matcher_2d.cc:61-176).
//
// The per-candidate score is an ORDERED float32 sum over the scan points
// (:64-72), so each candidate is summed by one thread in point order — that is
// what makes the result bit-identical to the reference.  The exp/hypot weight
// (:170-174) is evaluated in double on the host with libm, like every other
// transcendental of the path.
#include <algorithm>
#include <climits>
#include <cmath>

#include "engine2d.cuh"

namespace csm {

struct RtParams {
  int nx, ny, n, num_scans, lin, width;  // width = 2 * lin + 1
  float k_scale, cost_bias, max_cost, min_probability;
  // TSDF variant: value -> tsd / weight (tsd_value_converter.cc:24-34)
  float tsd_scale, tsd_bias, min_tsd, w_scale, w_bias, truncation;
};

// mapping/2d/probability_grid.cc:78-82 + probability_values.cc:29-37
__device__ __forceinline__ float GetProbability(const uint16_t* __restrict__ cells,
                                                const RtParams& P, int x, int y) {
  if (static_cast<unsigned>(x) >= static_cast<unsigned>(P.nx) ||
      static_cast<unsigned>(y) >= static_cast<unsigned>(P.ny))
    return P.min_probability;
  const int value = __ldg(cells + static_cast<size_t>(y) * P.nx + x) & 0x7fff;
  const float cost = value == 0 ? P.max_cost
                                : __fadd_rn(__fmul_rn(__int2float_rn(value), P.k_scale),
                                            P.cost_bias);
  return __fsub_rn(1.f, cost);
}

// One thread per candidate, candidates in the reference's generation order
// (scan-major, x outer, y inner; real_time...2d.cc:98-111).
__global__ void __launch_bounds__(128)
k_rt_score(const uint16_t* __restrict__ cells, const short2* __restrict__ dscan,
           const double* __restrict__ weight, RtParams P, float* __restrict__ scores) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_scan = P.width * P.width;
  if (c >= P.num_scans * per_scan) return;
  const int scan = c / per_scan;
  const int r = c - scan * per_scan;
  const int xo = -P.lin + r / P.width;
  const int yo = -P.lin + r % P.width;
  const short2* __restrict__ pts = dscan + static_cast<size_t>(scan) * P.n;
  float sum = 0.f;
  for (int p = 0; p < P.n; ++p) {
    const short2 q = pts[p];
    sum = __fadd_rn(sum, GetProbability(cells, P, q.x + xo, q.y + yo));
  }
  float score = __fdiv_rn(sum, __int2float_rn(P.n));
  // candidate.score *= exp(...): float *= double  (:170-174)
  score = __double2float_rn(__dmul_rn(static_cast<double>(score), weight[c]));
  scores[c] = score;
}

// TSDF variant (:38-59): score = sum(normalized_tsd * weight) / sum(weight), both sums
// ordered float sums over the scan points.
__global__ void __launch_bounds__(128)
k_rt_score_tsdf(const uint16_t* __restrict__ tsd_cells, const uint16_t* __restrict__ w_cells,
                const short2* __restrict__ dscan, const double* __restrict__ weight, RtParams P,
                float* __restrict__ scores) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_scan = P.width * P.width;
  if (c >= P.num_scans * per_scan) return;
  const int scan = c / per_scan;
  const int r = c - scan * per_scan;
  const int xo = -P.lin + r / P.width;
  const int yo = -P.lin + r % P.width;
  const short2* __restrict__ pts = dscan + static_cast<size_t>(scan) * P.n;
  float candidate_score = 0.f, summed_weight = 0.f;
  for (int p = 0; p < P.n; ++p) {
    const short2 q = pts[p];
    const int x = q.x + xo, y = q.y + yo;
    float tsd = P.min_tsd, w = 0.f;  // outside the limits (tsdf_2d.cc:94-95)
    if (static_cast<unsigned>(x) < static_cast<unsigned>(P.nx) &&
        static_cast<unsigned>(y) < static_cast<unsigned>(P.ny)) {
      const size_t flat = static_cast<size_t>(y) * P.nx + x;
      const int tv = __ldg(tsd_cells + flat) & 0x7fff, wv = __ldg(w_cells + flat) & 0x7fff;
      tsd = tv == 0 ? P.min_tsd : __fadd_rn(__fmul_rn(__int2float_rn(tv), P.tsd_scale), P.tsd_bias);
      w = wv == 0 ? 0.f : __fadd_rn(__fmul_rn(__int2float_rn(wv), P.w_scale), P.w_bias);
    }
    const float normalized = __fdiv_rn(__fsub_rn(P.truncation, fabsf(tsd)), P.truncation);
    candidate_score = __fadd_rn(candidate_score, __fmul_rn(normalized, w));
    summed_weight = __fadd_rn(summed_weight, w);
  }
  float score = summed_weight == 0.f ? 0.f : __fdiv_rn(candidate_score, summed_weight);
  score = __double2float_rn(__dmul_rn(static_cast<double>(score), weight[c]));
  scores[c] = score;
}

// std::max_element: first maximum in generation order (:142-143).
__global__ void __launch_bounds__(1024)
k_first_argmax(const float* __restrict__ scores, int count, int* __restrict__ best) {
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  float v = -INFINITY;
  int idx = INT_MAX;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const float x = scores[i];
    if (x > v || (x == v && i < idx)) { v = x; idx = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = v; s_i[threadIdx.x >> 5] = idx; }
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? s_v[threadIdx.x] : -INFINITY;
    idx = threadIdx.x < (blockDim.x >> 5) ? s_i[threadIdx.x] : INT_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (threadIdx.x == 0) *best = idx;
  }
}

}  // namespace csm

using namespace csm;

namespace {
struct HV3 { float x, y, z; };
inline HV3 HCross(const HV3& a, const HV3& b) {
  return HV3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Same Eigen formula as the device RotateRn (transform/rigid_transform.h:192-196).
inline HV3 HRotate(float qw, const HV3& qv, const HV3& v) {
  HV3 uv = HCross(qv, v);
  uv.x += uv.x; uv.y += uv.y; uv.z += uv.z;
  const HV3 c = HCross(qv, uv);
  HV3 r{(v.x + qw * uv.x) + c.x, (v.y + qw * uv.y) + c.y, (v.z + qw * uv.z) + c.z};
  r.x += 0.f; r.y += 0.f; r.z += 0.f;
  return r;
}
}  // namespace

static csm_status RtMatch(const uint16_t* cells, const uint16_t* weight_cells, float truncation,
                          float max_weight, int32_t nx, int32_t ny, double resolution,
                          double max_x, double max_y, const float* xyz, int32_t n,
                          const double initial_pose[3], double linear_window,
                          double angular_window, double w_t, double w_r, int32_t device,
                          double* score, double pose_estimate[3], csm_stats* stats) {
  CSM_REQUIRE(cells && xyz && initial_pose && score && pose_estimate, "null pointer");  // :121
  CSM_REQUIRE(nx >= 1 && ny >= 1 && n >= 1 && resolution > 0., "sizes");
  LaneGuard guard;
  CSM_TRY(AcquireLane(device, &guard));
  Ctx* ctx = guard.lane;
  CSM_CUDA(cudaSetDevice(device));
  cudaStream_t s = ctx->stream;

  // rotated_point_cloud (:123-127) on the host: SearchParameters needs its max range.
  const float yaw = static_cast<float>(initial_pose[2]);
  const float ha0 = 0.5f * yaw;
  const float s0 = std::sin(ha0);
  const HV3 q0{s0 * 0.f, s0 * 0.f, s0 * 1.f};
  const float q0w = std::cos(ha0);
  std::vector<float> rot(3 * static_cast<size_t>(n));
  float max_scan_range = 3.f * resolution;  // correlative_scan_matcher_2d.cc:34
  for (int i = 0; i < n; ++i) {
    const HV3 r = HRotate(q0w, q0, HV3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
    rot[3 * i] = r.x;
    rot[3 * i + 1] = r.y;
    rot[3 * i + 2] = r.z;
    const float range = std::sqrt(r.x * r.x + r.y * r.y);
    max_scan_range = std::max(range, max_scan_range);
  }
  const double kSafetyMargin = 1. - 1e-3;
  const double step = kSafetyMargin * std::acos(1. - (resolution * resolution) /
                                                         (2. * (max_scan_range * max_scan_range)));
  const int num_angular = static_cast<int>(std::ceil(angular_window / step));
  const int num_scans = 2 * num_angular + 1;
  const int lin = static_cast<int>(std::ceil(linear_window / resolution));
  const int width = 2 * lin + 1;
  const long long num_cand = static_cast<long long>(num_scans) * width * width;
  CSM_REQUIRE(num_scans > 0 && num_cand < (1LL << 28), "search window too large");

  std::vector<float> trig(2 * static_cast<size_t>(num_scans));
  std::vector<double> weight(num_cand);
  {
    double delta_theta = -num_angular * step;
    for (int k = 0; k < num_scans; ++k, delta_theta += step) {
      const float ha = 0.5f * static_cast<float>(delta_theta);
      trig[2 * k] = std::cos(ha);
      trig[2 * k + 1] = std::sin(ha);
      const double orientation = (k - num_angular) * step;  // Candidate2D ctor
      for (int xo = -lin; xo <= lin; ++xo)
        for (int yo = -lin; yo <= lin; ++yo) {
          const double cx = -yo * resolution, cy = -xo * resolution;
          const double e = std::hypot(cx, cy) * w_t + std::abs(orientation) * w_r;
          weight[(static_cast<size_t>(k) * width + (xo + lin)) * width + (yo + lin)] =
              std::exp(-(e * e));
        }
    }
  }

  // descriptors reused from the fast matcher's K2
  StackDev sd;
  std::memset(&sd, 0, sizeof(sd));
  sd.nx = nx;
  sd.ny = ny;
  sd.depth = 1;
  sd.resolution = resolution;
  sd.max_x = max_x;
  sd.max_y = max_y;
  DevBuf& d_sd = ctx->D("rt_stack");
  DevBuf& d_cells = ctx->D("rt_cells");
  DevBuf& d_wcells = ctx->D("rt_wcells");
  DevBuf& d_xyz = ctx->D("rt_xyz");
  DevBuf& d_trig = ctx->D("rt_trig");
  DevBuf& d_w = ctx->D("rt_weight");
  DevBuf& d_job = ctx->D("rt_job");
  DevBuf& d_sj = ctx->D("rt_scan_job");
  DevBuf& d_info = ctx->D("rt_info");
  DevBuf& d_dscan = ctx->D("rt_dscan");
  DevBuf& d_scores = ctx->D("rt_scores");
  DevBuf& d_misc = ctx->D("rt_misc");
  const size_t ncell = static_cast<size_t>(nx) * ny;
  CSM_TRY(d_sd.Reserve(sizeof(StackDev)));
  CSM_TRY(d_cells.Reserve(ncell * 2));
  if (weight_cells) CSM_TRY(d_wcells.Reserve(ncell * 2));
  CSM_TRY(d_xyz.Reserve(rot.size() * 4));
  CSM_TRY(d_trig.Reserve(trig.size() * 4));
  CSM_TRY(d_w.Reserve(weight.size() * 8));
  CSM_TRY(d_job.Reserve(sizeof(JobDev)));
  CSM_TRY(d_sj.Reserve(sizeof(int) * num_scans));
  CSM_TRY(d_info.Reserve(sizeof(ScanInfo) * num_scans));
  CSM_TRY(d_dscan.Reserve(sizeof(short2) * static_cast<size_t>(num_scans) * n));
  CSM_TRY(d_scores.Reserve(sizeof(float) * num_cand));
  CSM_TRY(d_misc.Reserve(64));

  JobDev jd;
  std::memset(&jd, 0, sizeof(jd));
  jd.stack = d_sd.as<StackDev>();
  jd.xyz = d_xyz.as<float>();
  jd.trig = d_trig.as<float2>();
  jd.n = n;
  jd.num_scans = num_scans;
  jd.lin = lin;
  jd.q0w = 1.f;  // the cloud is already rotated; identity is exact
  jd.tx = static_cast<float>(initial_pose[0]);  // Translation2f(double, double) (:135-137)
  jd.ty = static_cast<float>(initial_pose[1]);

  CSM_CUDA(cudaEventRecord(ctx->ev0, s));
  CSM_CUDA(cudaMemcpyAsync(d_sd.p, &sd, sizeof(sd), cudaMemcpyHostToDevice, s));
  CSM_CUDA(cudaMemcpyAsync(d_cells.p, cells, ncell * 2, cudaMemcpyHostToDevice, s));
  if (weight_cells)
    CSM_CUDA(cudaMemcpyAsync(d_wcells.p, weight_cells, ncell * 2, cudaMemcpyHostToDevice, s));
  CSM_CUDA(cudaMemcpyAsync(d_xyz.p, rot.data(), rot.size() * 4, cudaMemcpyHostToDevice, s));
  CSM_CUDA(cudaMemcpyAsync(d_trig.p, trig.data(), trig.size() * 4, cudaMemcpyHostToDevice, s));
  CSM_CUDA(cudaMemcpyAsync(d_w.p, weight.data(), weight.size() * 8, cudaMemcpyHostToDevice, s));
  CSM_CUDA(cudaMemcpyAsync(d_job.p, &jd, sizeof(jd), cudaMemcpyHostToDevice, s));
  CSM_CUDA(cudaMemsetAsync(d_sj.p, 0, sizeof(int) * num_scans, s));
  CSM_CUDA(cudaMemsetAsync(d_misc.p, 0, 64, s));
  CSM_TRY(LaunchDiscretize2D(s, d_job.as<JobDev>(), d_sj.as<int>(), num_scans,
                             d_dscan.as<short2>(), d_info.as<ScanInfo>(), 0,
                             d_misc.as<unsigned long long>()));
  RtParams P;
  P.nx = nx;
  P.ny = ny;
  P.n = n;
  P.num_scans = num_scans;
  P.lin = lin;
  P.width = width;
  {
    // probability_values.h:64-67 and .cc:29-37 evaluated in float
    const float kMinProbability = 0.1f;
    const float kMaxProbability = 1.f - kMinProbability;
    const float kMinCost = 1.f - kMaxProbability;
    const float kMaxCost = 1.f - kMinProbability;
    P.k_scale = (kMaxCost - kMinCost) / 32766.f;
    P.cost_bias = kMinCost - P.k_scale;
    P.max_cost = kMaxCost;
    P.min_probability = kMinProbability;
  }
  if (weight_cells) {
    // TSDValueConverter tables (tsd_value_converter.cc:24-34, value_conversion_tables.cc:29-37)
    const float min_tsd = -truncation;
    P.tsd_scale = (truncation - min_tsd) / 32766.f;
    P.tsd_bias = min_tsd - P.tsd_scale;
    P.min_tsd = min_tsd;
    P.w_scale = (max_weight - 0.f) / 32766.f;
    P.w_bias = 0.f - P.w_scale;
    P.truncation = truncation;
    k_rt_score_tsdf<<<static_cast<int>((num_cand + 127) / 128), 128, 0, s>>>(
        d_cells.as<uint16_t>(), d_wcells.as<uint16_t>(), d_dscan.as<short2>(), d_w.as<double>(),
        P, d_scores.as<float>());
  } else {
    k_rt_score<<<static_cast<int>((num_cand + 127) / 128), 128, 0, s>>>(
        d_cells.as<uint16_t>(), d_dscan.as<short2>(), d_w.as<double>(), P, d_scores.as<float>());
  }
  CSM_LAUNCH_CHECK();
  int* d_best = reinterpret_cast<int*>(d_misc.as<char>() + 32);
  k_first_argmax<<<1, 1024, 0, s>>>(d_scores.as<float>(), static_cast<int>(num_cand), d_best);
  CSM_LAUNCH_CHECK();
  CSM_CUDA(cudaEventRecord(ctx->ev1, s));
  int best = 0;
  CSM_CUDA(cudaMemcpyAsync(&best, d_best, sizeof(int), cudaMemcpyDeviceToHost, s));
  CSM_CUDA(cudaStreamSynchronize(s));
  float best_score = 0.f;
  CSM_CUDA(cudaMemcpy(&best_score, d_scores.as<float>() + best, sizeof(float),
                      cudaMemcpyDeviceToHost));
  const int scan = best / (width * width);
  const int r = best % (width * width);
  const int xo = -lin + r / width, yo = -lin + r % width;
  *score = best_score;
  pose_estimate[0] = initial_pose[0] + (-yo * resolution);
  pose_estimate[1] = initial_pose[1] + (-xo * resolution);
  pose_estimate[2] = initial_pose[2] + (scan - num_angular) * step;
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->candidates_scored = num_cand;
    stats->lowest_resolution_candidates = num_cand;
    stats->num_scans = num_scans;
    stats->best_scan_index = scan;
    stats->best_x_offset = xo;
    stats->best_y_offset = yo;
    stats->leaves_tied = 1;
    cudaEventElapsedTime(&stats->device_ms, ctx->ev0, ctx->ev1);
  }
  return CSM_OK;
}

extern "C" {

csm_status csm_rt_match2d(const uint16_t* cells, int32_t nx, int32_t ny, double resolution,
                          double max_x, double max_y, const float* xyz, int32_t n,
                          const double initial_pose[3], double linear_window,
                          double angular_window, double w_t, double w_r, int32_t device,
                          double* score, double pose_estimate[3], csm_stats* stats) {
  return RtMatch(cells, nullptr, 0.f, 0.f, nx, ny, resolution, max_x, max_y, xyz, n,
                 initial_pose, linear_window, angular_window, w_t, w_r, device, score,
                 pose_estimate, stats);
}

csm_status csm_rt_match2d_tsdf(const uint16_t* tsd_cells, const uint16_t* weight_cells,
                               int32_t nx, int32_t ny, double resolution, double max_x,
                               double max_y, float truncation_distance, float max_weight,
                               const float* xyz, int32_t n, const double initial_pose[3],
                               double linear_window, double angular_window, double w_t,
                               double w_r, int32_t device, double* score,
                               double pose_estimate[3], csm_stats* stats) {
  CSM_REQUIRE(weight_cells != nullptr, "null weight cells");
  CSM_REQUIRE(truncation_distance > 0.f && max_weight > 0.f, "TSDF parameters");
  return RtMatch(tsd_cells, weight_cells, truncation_distance, max_weight, nx, ny, resolution,
                 max_x, max_y, xyz, n, initial_pose, linear_window, angular_window, w_t, w_r,
                 device, score, pose_estimate, stats);
}

}  // extern "C"
