// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference algorithm (cartographer-project/cartographer
// @ 877157a0) for the correlative scan-matching hot path.  It is the parity
// checker for the CUDA engine in cartographer_b200/ and the CPU baseline that
// bench.py times (`cpu_baseline`, `--impl reference`).  Only tests/ (incl. the
// fixture generator tests/golden/make_golden.py), __graft_entry__.smoke() and the
// measurement scripts' CPU-baseline / checker legs (bench.py, benchmarks/) may use
// anything in this directory.  The product path (cartographer_b200/, libcsm_b200.so)
// never links, loads or calls it (tests/test_abi_cpu.py enforces that).
//
// Parity pin status: pinned against the reference's own known-answer tests
// (tests/test_oracle_golden_2d.py / _3d.py list every vector and its file:line) and
// frozen by the committed fixtures tests/golden/golden_v1.npz.  The real
// reference cannot be compiled in this image (Eigen, glog, abseil, protobuf,
// Ceres, Lua are absent and its headers need protoc-generated code), so there
// is no oracle/_ref; see DESIGN.md "Oracle".
//
// Every function cites the reference file:line it restates (paths relative to
// /root/reference/cartographer/).  Arithmetic types and operation order are
// kept exactly: float where the reference is float, double where it is double,
// no FMA contraction (build with -ffp-contract=off, no -march, no fast-math,
// mirroring cmake/functions.cmake:100-101).
#ifndef ORACLE_COMMON_H_
#define ORACLE_COMMON_H_

#include <cmath>
#include <cstdint>
#include <vector>

namespace oracle {

// common/port.h:40-42  RoundToInt = std::lround.
inline int RoundToInt(const float x) { return static_cast<int>(std::lround(x)); }
inline int RoundToInt(const double x) { return static_cast<int>(std::lround(x)); }

// common/math.h:31-40
template <typename T>
inline T Clamp(const T value, const T min, const T max) {
  if (value > max) return max;
  if (value < min) return min;
  return value;
}
// common/math.h:48-52
template <typename T>
inline T Pow2(T a) { return a * a; }

// mapping/probability_values.h:64-67 — evaluated in float exactly as written.
constexpr float kMinProbability = 0.1f;
constexpr float kMaxProbability = 1.f - kMinProbability;
constexpr float kMinCorrespondenceCost = 1.f - kMaxProbability;
constexpr float kMaxCorrespondenceCost = 1.f - kMinProbability;
constexpr uint16_t kUnknownProbabilityValue = 0;          // h:80
constexpr uint16_t kUnknownCorrespondenceValue = 0;       // h:81
constexpr uint16_t kUpdateMarker = 1u << 15;              // h:82

// mapping/probability_values.h:32-44
inline uint16_t BoundedFloatToValue(const float float_value,
                                    const float lower_bound,
                                    const float upper_bound) {
  const int value =
      RoundToInt((Clamp(float_value, lower_bound, upper_bound) - lower_bound) *
                 (32766.f / (upper_bound - lower_bound))) +
      1;
  return static_cast<uint16_t>(value);
}
// h:85-88, h:91-93
inline uint16_t CorrespondenceCostToValue(const float correspondence_cost) {
  return BoundedFloatToValue(correspondence_cost, kMinCorrespondenceCost,
                             kMaxCorrespondenceCost);
}
inline uint16_t ProbabilityToValue(const float probability) {
  return BoundedFloatToValue(probability, kMinProbability, kMaxProbability);
}
inline float ProbabilityToCorrespondenceCost(const float p) { return 1.f - p; }
inline float CorrespondenceCostToProbability(const float c) { return 1.f - c; }
inline float Odds(float p) { return p / (1.f - p); }                 // h:46-48
inline float ProbabilityFromOdds(const float o) { return o / (o + 1.f); }  // h:50-52

// mapping/value_conversion_tables.cc:29-37 / probability_values.cc:29-37
inline float SlowValueToBoundedFloat(const uint16_t value,
                                     const uint16_t unknown_value,
                                     const float unknown_result,
                                     const float lower_bound,
                                     const float upper_bound) {
  if (value == unknown_value) return unknown_result;
  const float kScale = (upper_bound - lower_bound) / 32766.f;
  return value * kScale + (lower_bound - kScale);
}

// mapping/value_conversion_tables.cc:39-51: 65536 entries, bit 15 masked.
inline std::vector<float> PrecomputeValueToBoundedFloat(
    const uint16_t unknown_value, const float unknown_result,
    const float lower_bound, const float upper_bound) {
  std::vector<float> result;
  result.reserve(65536);
  for (size_t value = 0; value != 65536; ++value) {
    result.push_back(SlowValueToBoundedFloat(
        static_cast<uint16_t>(value) & static_cast<uint16_t>(~kUpdateMarker),
        unknown_value, unknown_result, lower_bound, upper_bound));
  }
  return result;
}

// ---- Eigen semantics the path relies on -------------
struct Vec3f { float x, y, z; };
struct Quatf { float w, x, y, z; };

// Eigen 3.3 MatrixBase::cross for 3-vectors (scalar path).
inline Vec3f Cross(const Vec3f& a, const Vec3f& b) {
  return Vec3f{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z,
               a.x * b.y - a.y * b.x};
}
// Eigen 3.3 Quaternion::operator=(AngleAxis): ha = 0.5*angle; w = cos(ha);
// vec = sin(ha) * axis.   Here axis = UnitZ.
inline Quatf QuatFromAngleAxisZ(const float angle) {
  const float ha = 0.5f * angle;
  const float s = std::sin(ha);
  return Quatf{std::cos(ha), s * 0.f, s * 0.f, s * 1.f};
}
// Eigen 3.3 QuaternionBase::_transformVector:
//   uv = q.vec x v; uv += uv; return v + q.w * uv + q.vec x uv.
inline Vec3f Rotate(const Quatf& q, const Vec3f& v) {
  const Vec3f qv{q.x, q.y, q.z};
  Vec3f uv = Cross(qv, v);
  uv.x += uv.x; uv.y += uv.y; uv.z += uv.z;
  const Vec3f c = Cross(qv, uv);
  return Vec3f{(v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y,
               (v.z + q.w * uv.z) + c.z};
}
// transform/rigid_transform.h:192-196 with translation == Zero
// (Rigid3f::Rotation), applied per point by sensor/point_cloud.cc:56-64.
inline Vec3f RotateAsRigid3f(const Quatf& q, const Vec3f& v) {
  const Vec3f r = Rotate(q, v);
  return Vec3f{r.x + 0.f, r.y + 0.f, r.z + 0.f};
}

using PointCloud = std::vector<Vec3f>;

inline PointCloud TransformPointCloudRotZ(const PointCloud& cloud,
                                          const float angle) {
  const Quatf q = QuatFromAngleAxisZ(angle);
  PointCloud out;
  out.reserve(cloud.size());
  for (const Vec3f& p : cloud) out.push_back(RotateAsRigid3f(q, p));
  return out;
}

}  // namespace oracle

#endif  // ORACLE_COMMON_H_
