// FastCorrelativeScanMatcher2D on B200: precomputation-grid stack build,
// scan discretisation, candidate scoring and a batched, device-resident
// branch-and-bound.  Reference: cartographer/mapping/internal/2d/scan_matching/
// {correlative_scan_matcher_2d,fast_correlative_scan_matcher_2d}.{h,cc}.
//
// Exactness rules: every float/double expression whose
// value reaches an output is written with explicit round-to-nearest intrinsics
// (__fmul_rn ...) so ptxas can never contract it into an FMA — the reference
// build has no FMA (cmake/functions.cmake:100-101).  Transcendentals are
// evaluated on the host with libm, like the reference.
#include "engine2d.cuh"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <functional>

namespace csm {

// ---------------------------------------------------------------------------
// Small device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned FloatToOrdered(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float OrderedToFloat(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
static inline unsigned HostFloatToOrdered(float f) {
  unsigned u;
  std::memcpy(&u, &f, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static inline float HostOrderedToFloat(unsigned u) {
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

// PrecomputationGrid2D::ToScore(sum / float(N))   (fast...2d.h:74-76, .cc:328-329)
__device__ __forceinline__ float ToScore(const StackDev& st, int sum, int n) {
  const float mean = __fdiv_rn(__int2float_rn(sum), __int2float_rn(n));
  return __fadd_rn(st.min_score, __fmul_rn(mean, st.k255));
}

// PrecomputationGrid2D::GetValue (fast...2d.h:56-71) on the row-major level.
__device__ __forceinline__ int GetValue(const uint8_t* __restrict__ g, int wx, int wy, int lx,
                                        int ly) {
  if (static_cast<unsigned>(lx) >= static_cast<unsigned>(wx) ||
      static_cast<unsigned>(ly) >= static_cast<unsigned>(wy))
    return 0;
  return __ldg(g + static_cast<size_t>(ly) * wx + lx);
}

__device__ __forceinline__ int WarpSum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Word index of lattice cell (X, Y) (both already shifted by +1) of phase `ph` in the
// child-window array of a level (StackDev::win).
//   default      : row-major per phase — x-neighbours share 128-byte lines;
//   CSM_WIN_TILED: 8 x 4 word tiles (one line each) — x- AND y-neighbours share lines, so
//                  a compact blob of surviving parents touches about half as many lines.
#ifdef CSM_WIN_TILED
__host__ __device__ __forceinline__ long long WinPhaseWords(int jd, int ids) {
  return static_cast<long long>((jd + 3) >> 2) * ((ids + 7) >> 3) * 32;
}
__device__ __forceinline__ int WinCell(int X, int Y, int ids) {
  return ((((Y >> 2) * ((ids + 7) >> 3) + (X >> 3)) << 5) | ((Y & 3) << 3) | (X & 7));
}
#else
__host__ __device__ __forceinline__ long long WinPhaseWords(int jd, int ids) {
  return static_cast<long long>(jd) * ids;
}
__device__ __forceinline__ int WinCell(int X, int Y, int ids) { return Y * ids + X; }
#endif

// ---------------------------------------------------------------------------
// K1: precomputation grid stack
// ---------------------------------------------------------------------------
// Level 0: cells -> uint8 through the 64 Ki LUT
//   lut[v] = ComputeCellValue(1.f - |table[v]|)   (fast...2d.cc:110-111,163-169)
__global__ void k_stack_level0(const uint16_t* __restrict__ cells,
                               const uint8_t* __restrict__ lut, uint8_t* __restrict__ out,
                               int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = lut[cells[i]];
}

// Level h (window w = 2^h) from level h-1 (window w/2): a w x w window is the
// union of four (w/2) x (w/2) windows; max is exact and quantisation is
// monotone, so max-of-uint8 equals the reference's quantised float max
// (fast...2d.cc:108-160).  Coordinates are wide-grid local indices.
__global__ void k_stack_double(const uint8_t* __restrict__ prev, int pwx, int pwy,
                               uint8_t* __restrict__ out, int wx, int wy, int half) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= wx || y >= wy) return;
  auto at = [&](int px, int py) -> int {
    return (px >= 0 && py >= 0 && px < pwx && py < pwy) ? prev[static_cast<size_t>(py) * pwx + px]
                                                        : 0;
  };
  int v = max(max(at(x - half, y - half), at(x, y - half)), max(at(x - half, y), at(x, y)));
  out[static_cast<size_t>(y) * wx + x] = static_cast<uint8_t>(v);
}

// Four byte-shifted copies of the decimated lowest-resolution level
// (see StackDev::dec4).
__global__ void k_stack_decimate4(const uint8_t* __restrict__ lvl, int wx, int wy, int h,
                                  uint8_t* __restrict__ dec4, int lpad, int id, int jd,
                                  int ids) {
  const int s = 1 << h;
  const long long total = static_cast<long long>(s) * s * jd * ids;
  const long long n = 4LL * lpad;
  for (long long u = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; u < n;
       u += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(u / lpad);
    const long long t = u % lpad - 16 + k;  // index into D
    uint8_t v = 0;
    if (t >= 0 && t < total) {
      const int I = static_cast<int>(t % ids);
      long long r = t / ids;
      const int J = static_cast<int>(r % jd);
      r /= jd;
      const int ax = static_cast<int>(r % s);
      const int ay = static_cast<int>(r / s);
      const int x = s * I + ax, y = s * J + ay;
      if (I < id && x < wx && y < wy) v = lvl[static_cast<size_t>(y) * wx + x];
    }
    dec4[u] = v;
  }
}

// Child-window layout of parent level h (see StackDev::win): one 32-bit word per
// (phase, lattice cell) with the four children values of level h-1.
__global__ void k_stack_window(const uint8_t* __restrict__ lvl, int wx, int wy, int h,
                               unsigned* __restrict__ win, int jd, int ids) {
  const int S = 1 << h, s = S >> 1;
  const long long per_phase = WinPhaseWords(jd, ids);
  const long long total = static_cast<long long>(S) * S * per_phase;
  for (long long u = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; u < total;
       u += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ph = static_cast<int>(u / per_phase);
    const int c = static_cast<int>(u - ph * per_phase);
#ifdef CSM_WIN_TILED
    const int it = (ids + 7) >> 3;
    const int tile = c >> 5, in = c & 31;
    const int X = ((tile % it) << 3) | (in & 7), Y = ((tile / it) << 2) | (in >> 3);
#else
    const int X = c % ids, Y = c / ids;
#endif
    const int I = X - 1, J = Y - 1;
    const int ax = ph % S, ay = ph / S;
    const int x = S * I + ax, y = S * J + ay;
    auto at = [&](int px, int py) -> unsigned {
      return (px >= 0 && py >= 0 && px < wx && py < wy) ? lvl[static_cast<size_t>(py) * wx + px]
                                                        : 0u;
    };
    unsigned v = 0u;
    if (X < ids && Y < jd)
      v = at(x, y) | (at(x + s, y) << 8) | (at(x, y + s) << 16) | (at(x + s, y + s) << 24);
    win[u] = v;
  }
}

// ---------------------------------------------------------------------------
// K2: GenerateRotatedScans + DiscretizeScans + ShrinkToFit
// ---------------------------------------------------------------------------
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 CrossRn(const V3& a, const V3& b) {
  return V3{__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)),
            __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
            __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x))};
}
// Eigen QuaternionBase::_transformVector followed by "+ translation(0)"
// (transform/rigid_transform.h:192-196, sensor/point_cloud.cc:56-64).
__device__ __forceinline__ V3 RotateRn(float qw, const V3& qv, const V3& v) {
  V3 uv = CrossRn(qv, v);
  uv.x = __fadd_rn(uv.x, uv.x);
  uv.y = __fadd_rn(uv.y, uv.y);
  uv.z = __fadd_rn(uv.z, uv.z);
  const V3 c = CrossRn(qv, uv);
  V3 r{__fadd_rn(__fadd_rn(v.x, __fmul_rn(qw, uv.x)), c.x),
       __fadd_rn(__fadd_rn(v.y, __fmul_rn(qw, uv.y)), c.y),
       __fadd_rn(__fadd_rn(v.z, __fmul_rn(qw, uv.z)), c.z)};
  r.x = __fadd_rn(r.x, 0.f);
  r.y = __fadd_rn(r.y, 0.f);
  r.z = __fadd_rn(r.z, 0.f);
  return r;
}

// One CTA per (job, scan).  Writes the scan's cell indices and its LinearBounds
// after ShrinkToFit (correlative_scan_matcher_2d.cc:73-127), plus the number of
// lowest-resolution candidates per axis (fast...2d.cc:281-292).
__global__ void __launch_bounds__(128)
k_discretize(const JobDev* __restrict__ jobs, const int* __restrict__ scan_job,
             short2* __restrict__ dscan, ScanInfo* __restrict__ info, int shrink,
             unsigned long long* __restrict__ counters) {
  const int sg = blockIdx.x;
  const int j = scan_job[sg];
  const JobDev jb = jobs[j];
  const StackDev& st = *jb.stack;
  const int k = sg - jb.scan_base;
  const float2 cs = jb.trig[k];
  // Quaternionf(AngleAxisf(a, UnitZ)): w = cos(ha), vec = sin(ha) * (0, 0, 1)
  const V3 qk{__fmul_rn(cs.y, 0.f), __fmul_rn(cs.y, 0.f), __fmul_rn(cs.y, 1.f)};
  const V3 q0{jb.q0x, jb.q0y, jb.q0z};
  short2* out = dscan + jb.dscan_off + static_cast<long long>(k) * jb.n;
  int min_ix = INT_MAX, max_ix = INT_MIN, min_iy = INT_MAX, max_iy = INT_MIN;
  for (int p = threadIdx.x; p < jb.n; p += blockDim.x) {
    const V3 v{jb.xyz[3 * p], jb.xyz[3 * p + 1], jb.xyz[3 * p + 2]};
    const V3 r0 = RotateRn(jb.q0w, q0, v);   // rotated_point_cloud   (fast...2d.cc:236-239)
    const V3 r1 = RotateRn(cs.x, qk, r0);    // GenerateRotatedScans  (corr...2d.cc:93-109)
    // Affine2f(Translation2f) * v = (1*x + 0*y) + t                    (corr...2d.cc:120-121)
    const float px = __fadd_rn(__fadd_rn(__fmul_rn(1.f, r1.x), __fmul_rn(0.f, r1.y)), jb.tx);
    const float py = __fadd_rn(__fadd_rn(__fmul_rn(0.f, r1.x), __fmul_rn(1.f, r1.y)), jb.ty);
    // MapLimits::GetCellIndex in double                               (2d/map_limits.h:69-76)
    const double fx = __dsub_rn(__ddiv_rn(__dsub_rn(st.max_y, static_cast<double>(py)),
                                          st.resolution), 0.5);
    const double fy = __dsub_rn(__ddiv_rn(__dsub_rn(st.max_x, static_cast<double>(px)),
                                          st.resolution), 0.5);
    const int ix = static_cast<int>(llround(fx));
    const int iy = static_cast<int>(llround(fy));
    // cells are kept as 2 x int16 (grids are < 32 k cells per axis; points farther
    // than 30 k cells from the origin read as 0 for every candidate either way)
    out[p] = make_short2(static_cast<short>(max(-30000, min(30000, ix))),
                         static_cast<short>(max(-30000, min(30000, iy))));
    if (shrink && (abs(ix) > 30000 || abs(iy) > 30000)) counters[7] = 1ull;  // reported as an error
    min_ix = min(min_ix, ix);
    max_ix = max(max_ix, ix);
    min_iy = min(min_iy, iy);
    max_iy = max(max_iy, iy);
  }
  __shared__ int red[4][4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    min_ix = min(min_ix, __shfl_xor_sync(0xffffffffu, min_ix, o));
    max_ix = max(max_ix, __shfl_xor_sync(0xffffffffu, max_ix, o));
    min_iy = min(min_iy, __shfl_xor_sync(0xffffffffu, min_iy, o));
    max_iy = max(max_iy, __shfl_xor_sync(0xffffffffu, max_iy, o));
  }
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    red[warp][0] = min_ix;
    red[warp][1] = max_ix;
    red[warp][2] = min_iy;
    red[warp][3] = max_iy;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 4; ++w) {
      red[0][0] = min(red[0][0], red[w][0]);
      red[0][1] = max(red[0][1], red[w][1]);
      red[0][2] = min(red[0][2], red[w][2]);
      red[0][3] = max(red[0][3], red[w][3]);
    }
    ScanInfo si;
    si.job = j;
    si.min_x = -jb.lin;
    si.max_x = jb.lin;
    si.min_y = -jb.lin;
    si.max_y = jb.lin;
    if (shrink) {
      // min_bound = min(0, min(-xy)); max_bound = max(0, max(limits - 1 - xy))
      const int min_bx = min(0, -red[0][1]), max_bx = max(0, st.nx - 1 - red[0][0]);
      const int min_by = min(0, -red[0][3]), max_by = max(0, st.ny - 1 - red[0][2]);
      si.min_x = max(si.min_x, min_bx);
      si.max_x = min(si.max_x, max_bx);
      si.min_y = max(si.min_y, min_by);
      si.max_y = min(si.max_y, max_by);
    }
    const int step = 1 << (st.depth - 1);
    si.nxc = (si.max_x - si.min_x + step) / step;
    si.nyc = (si.max_y - si.min_y + step) / step;
    si.pad = 0;
    if (shrink && (si.nyc > jb.cap_y || static_cast<long long>(si.nxc) * jb.cap_y > jb.cap)) {
      // the host's a-priori slot bound must cover the lattice (reported as an error)
      counters[6] = 1ull;
      si.nxc = si.nyc = 0;
    }
    info[sg] = si;
    const unsigned long long slots = static_cast<unsigned long long>(si.nxc) * si.nyc;
    atomicAdd(&counters[0], slots);  // every lowest-resolution candidate gets scored
    atomicAdd(&counters[3], slots);
  }
}

csm_status LaunchDiscretize2D(cudaStream_t stream, const JobDev* jobs, const int* scan_job,
                              int total_scans, short2* dscan, ScanInfo* info, int shrink,
                              unsigned long long* counters) {
  k_discretize<<<total_scans, 128, 0, stream>>>(jobs, scan_job, dscan, info, shrink, counters);
  CSM_LAUNCH_CHECK();
  return CSM_OK;
}

// ---------------------------------------------------------------------------
// K3: candidate scoring
// ---------------------------------------------------------------------------
// Lowest-resolution pass, gather form: one warp per (scan, slot); slot = i*nyc+j
// follows the reference's generation order (x outer, y inner; fast...2d.cc:296-309).
__global__ void __launch_bounds__(256)
k_score_top_gather(const JobDev* __restrict__ jobs, const ScanInfo* __restrict__ info,
                   const short2* __restrict__ dscan, int* __restrict__ top_sum,
                   const long long* __restrict__ scan_slot_base, int total_scans,
                   long long total_slots) {
  const int lane = threadIdx.x & 31;
  const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (w >= total_slots) return;
  // scan = last index with scan_slot_base[scan] <= w
  int lo = 0, hi = total_scans - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (scan_slot_base[mid] <= w) lo = mid; else hi = mid - 1;
  }
  const int sg = lo;
  const int slot = static_cast<int>(w - scan_slot_base[sg]);
  const ScanInfo si = info[sg];
  if (slot >= si.nxc * si.nyc) return;
  const JobDev& jb = jobs[si.job];
  const StackDev& st = *jb.stack;
  const int h = st.depth - 1;
  const int w1 = (1 << h) - 1;
  const uint8_t* __restrict__ g = st.level[h];
  const int wx = st.wx[h], wy = st.wy[h];
  const short2* __restrict__ pts = dscan + jb.dscan_off +
                                 static_cast<long long>(sg - jb.scan_base) * jb.n;
  const int i = slot / si.nyc, jy = slot - i * si.nyc;
  const int ox = si.min_x + (i << h) + w1, oy = si.min_y + (jy << h) + w1;
  int sum = 0;
  for (int p = lane; p < jb.n; p += 32) {
    const short2 c = pts[p];
    sum += GetValue(g, wx, wy, c.x + ox, c.y + oy);
  }
  sum = WarpSum(sum);
  if (lane == 0) top_sum[scan_slot_base[sg] + slot] = sum;
}

// Lowest-resolution pass, small-lattice form (local search windows: a few dozen
// candidates per scan).  `lanes` consecutive lanes of a warp share one scan; a lane
// owns one quad = 4 x-consecutive candidates of one lattice row and fetches their
// cells for a scan point with ONE aligned 32-bit load from the decimated level
// (as k_score_top_dense), so a warp works on 32 / lanes scans at once and nearly
// every lane is busy.  Packed u16 sums are flushed every 256 points.
__global__ void __launch_bounds__(128)
k_score_top_small(const JobDev* __restrict__ jobs, const ScanInfo* __restrict__ info,
                  const short2* __restrict__ dscan, int* __restrict__ top_sum,
                  const long long* __restrict__ scan_slot_base, int total_scans, int lanes) {
  extern __shared__ __align__(16) int2 s_small[];  // [warp][scan of the warp][32 points]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int spw = 32 / lanes;                 // scans per warp
  const int sub = lane / lanes, ql = lane - sub * lanes;
  const long long gwarp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long sgl = gwarp * spw + sub;
  const bool has_scan = sub < spw && sgl < total_scans;
  const int sg = has_scan ? static_cast<int>(sgl) : 0;
  const ScanInfo si = info[sg];
  const JobDev& jb = jobs[si.job];
  const StackDev& st = *jb.stack;
  const int h = st.depth - 1;
  const int s1 = (1 << h) - 1;
  const int id = st.dec_id[h], jd = st.dec_jd[h], ids = st.dec_ids[h];
  const unsigned lpad1 = static_cast<unsigned>(st.dec_lpad[h]) - 1u;
  const uint8_t* __restrict__ dec = st.dec4[h] + 16;
  const int qr = (si.nxc + 3) >> 2;
  const bool owns = has_scan && ql < qr * si.nyc;   // this lane scores a quad
  const int jy = ql / qr, i0 = (ql - jy * qr) << 2;
  const int toff = jy * ids + i0;
  const int ox = si.min_x + s1, oy = si.min_y + s1;
  const short2* __restrict__ pts = dscan + jb.dscan_off +
                                 static_cast<long long>(sg - jb.scan_base) * jb.n;
  int2* __restrict__ s_pt = s_small + (warp * spw + (sub < spw ? sub : 0)) * 32;
  // all jobs of a batch have the same point count only per job: the warp walks the
  // longest of its scans, lanes of shorter scans idle
  int n_max = has_scan ? jb.n : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, o));
  unsigned sum0 = 0, sum1 = 0, sum2 = 0, sum3 = 0;
  unsigned a02 = 0, a13 = 0;
  for (int p0 = 0; p0 < n_max; p0 += 32) {
    // the scan's lanes stage {D index of the lattice origin, qy << 16 | qx} of 32 points
    __syncwarp();
    if (has_scan) {
      for (int t = ql; t < 32; t += lanes) {
        int2 d = make_int2(0, static_cast<int>(0x80008000u));  // never in range
        if (p0 + t < jb.n) {
          const short2 c = pts[p0 + t];
          const int bx = c.x + ox, by = c.y + oy;
          const int qx = bx >> h, qy = by >> h;  // floor division
          if (qx > -32000 && qx < 32000 && qy > -32000 && qy < 32000)
            d = make_int2(((((by & s1) << h) | (bx & s1)) * jd + qy) * ids + qx,
                          (qy << 16) | (qx & 0xffff));
        }
        s_pt[t] = d;
      }
    }
    __syncwarp();
    if (owns) {
#pragma unroll 8
      for (int t = 0; t < 32; t += 2) {
        const int4 d = *reinterpret_cast<const int4*>(s_pt + t);
        const int Ja = (d.y >> 16) + jy, ca = static_cast<short>(d.y & 0xffff) + i0 + 3;
        const int Jb = (d.w >> 16) + jy, cb = static_cast<short>(d.w & 0xffff) + i0 + 3;
        if (static_cast<unsigned>(Ja) < static_cast<unsigned>(jd) &&
            static_cast<unsigned>(ca) < static_cast<unsigned>(id + 3)) {
          const int a = d.x + toff;
          const unsigned k = static_cast<unsigned>(a) & 3u;
          const unsigned w = __ldg(reinterpret_cast<const unsigned*>(
              dec + static_cast<long long>(a) + static_cast<long long>(k * lpad1)));
          a02 += __byte_perm(w, 0u, 0x4240);  // bytes 0 and 2 in u16 lanes
          a13 += __byte_perm(w, 0u, 0x4341);  // bytes 1 and 3
        }
        if (static_cast<unsigned>(Jb) < static_cast<unsigned>(jd) &&
            static_cast<unsigned>(cb) < static_cast<unsigned>(id + 3)) {
          const int a = d.z + toff;
          const unsigned k = static_cast<unsigned>(a) & 3u;
          const unsigned w = __ldg(reinterpret_cast<const unsigned*>(
              dec + static_cast<long long>(a) + static_cast<long long>(k * lpad1)));
          a02 += __byte_perm(w, 0u, 0x4240);
          a13 += __byte_perm(w, 0u, 0x4341);
        }
      }
    }
    if (((p0 + 32) & 255) == 0 || p0 + 32 >= n_max) {  // flush the packed u16 sums
      sum0 += a02 & 0xffffu;
      sum2 += a02 >> 16;
      sum1 += a13 & 0xffffu;
      sum3 += a13 >> 16;
      a02 = a13 = 0u;
    }
  }
  if (!owns) return;
  int* __restrict__ out = top_sum + scan_slot_base[sg];
  const unsigned sums[4] = {sum0, sum1, sum2, sum3};
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (i0 + e < si.nxc) out[(i0 + e) * si.nyc + jy] = static_cast<int>(sums[e]);
}

// Lowest-resolution pass, dense form: one CTA per scan.  The candidates of one
// rotated scan form a lattice of stride s = 2^h, so for a scan point p the cells
// they read are one contiguous block of the decimated level (StackDev::dec4).
// A thread owns kQuads "quads" = 4 x-consecutive candidates of one lattice row
// and fetches their 4 cells with ONE aligned 32-bit load from the byte-shifted
// copy that matches the address' alignment; the four byte lanes are accumulated
// SIMD-in-register (2 x u16 per register, flushed every 256 points).  Point
// descriptors (tile base + lattice origin) are computed once per CTA into shared
// memory and broadcast to all threads.
constexpr int kDenseThreads = 128;
constexpr int kDenseChunk = 256;  // <= 257 so the packed u16 sums cannot overflow
template <int kQuads>
__global__ void __launch_bounds__(kDenseThreads)
k_score_top_dense(const JobDev* __restrict__ jobs, const ScanInfo* __restrict__ info,
                  const short2* __restrict__ dscan, int* __restrict__ top_sum,
                  const long long* __restrict__ scan_slot_base, int total_scans) {
  __shared__ int4 s_pt[kDenseChunk];  // {D index of lattice origin, qx, qy, 0}
  for (int sg = blockIdx.x; sg < total_scans; sg += gridDim.x) {
    const ScanInfo si = info[sg];
    const JobDev& jb = jobs[si.job];
    const StackDev& st = *jb.stack;
    const int h = st.depth - 1;
    const int s = 1 << h;
    const int id = st.dec_id[h], jd = st.dec_jd[h], ids = st.dec_ids[h];
    const unsigned lpad1 = static_cast<unsigned>(st.dec_lpad[h]) - 1u;
    const uint8_t* __restrict__ dec = st.dec4[h] + 16;
    const int qr = (si.nxc + 3) >> 2;       // quads per lattice row
    const int quads = qr * si.nyc;
    const short2* __restrict__ pts = dscan + jb.dscan_off +
                                   static_cast<long long>(sg - jb.scan_base) * jb.n;
    int* __restrict__ out = top_sum + scan_slot_base[sg];
    for (int u0 = 0; u0 < quads; u0 += kDenseThreads * kQuads) {
      int jy[kQuads], i0[kQuads], toff[kQuads];
      unsigned sum[kQuads][4];
#pragma unroll
      for (int r = 0; r < kQuads; ++r) {
        const int u = u0 + threadIdx.x + r * kDenseThreads;
        jy[r] = u / qr;
        i0[r] = (u - jy[r] * qr) << 2;
        if (u >= quads) jy[r] = 1 << 20;    // never in range
        toff[r] = jy[r] * ids + i0[r];
        sum[r][0] = sum[r][1] = sum[r][2] = sum[r][3] = 0u;
      }
      for (int p0 = 0; p0 < jb.n; p0 += kDenseChunk) {
        __syncthreads();
        for (int t = threadIdx.x; t < kDenseChunk; t += kDenseThreads) {
          const int p = p0 + t;
          int4 d = make_int4(0, -(1 << 24), -(1 << 24), 0);
          if (p < jb.n) {
            const short2 c = pts[p];
            const int bx = c.x + si.min_x + s - 1, by = c.y + si.min_y + s - 1;
            const int qx = bx >> h, qy = by >> h;          // floor division
            const int ax = bx & (s - 1), ay = by & (s - 1);
            if (qx > -(1 << 20) && qx < (1 << 20) && qy > -(1 << 20) && qy < (1 << 20))
              d = make_int4(((ay * s + ax) * jd + qy) * ids + qx, qx, qy, 0);
          }
          s_pt[t] = d;
        }
        __syncthreads();
        const int cnt = min(kDenseChunk, jb.n - p0);
        unsigned a02[kQuads], a13[kQuads];
#pragma unroll
        for (int r = 0; r < kQuads; ++r) a02[r] = a13[r] = 0u;
#pragma unroll 4
        for (int t = 0; t < cnt; ++t) {
          const int4 d = s_pt[t];
#pragma unroll
          for (int r = 0; r < kQuads; ++r) {
            const int J = d.z + jy[r];
            const int c3 = d.y + i0[r] + 3;   // column of the quad's last byte
            if (static_cast<unsigned>(J) < static_cast<unsigned>(jd) &&
                static_cast<unsigned>(c3) < static_cast<unsigned>(id + 3)) {
              const int a = d.x + toff[r];    // D index of the quad's first byte (>= -3)
              const unsigned k = static_cast<unsigned>(a) & 3u;
              const unsigned w = __ldg(reinterpret_cast<const unsigned*>(
                  dec + static_cast<long long>(a) + static_cast<long long>(k * lpad1)));
              a02[r] += __byte_perm(w, 0u, 0x4240);  // bytes 0 and 2 in u16 lanes
              a13[r] += __byte_perm(w, 0u, 0x4341);  // bytes 1 and 3
            }
          }
        }
#pragma unroll
        for (int r = 0; r < kQuads; ++r) {
          sum[r][0] += a02[r] & 0xffffu;
          sum[r][2] += a02[r] >> 16;
          sum[r][1] += a13[r] & 0xffffu;
          sum[r][3] += a13[r] >> 16;
        }
      }
#pragma unroll
      for (int r = 0; r < kQuads; ++r) {
        if (jy[r] < si.nyc) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (i0[r] + e < si.nxc) out[(i0[r] + e) * si.nyc + jy[r]] = static_cast<int>(sum[r][e]);
        }
      }
    }
  }
}

// ---- lowest-resolution pass, tile form ------------------------------------------
// For one scan point the candidates of a scan's lattice that fall inside the grid
// read exactly one tile of the decimated level: jd rows of ids bytes, contiguous in
// memory (phase (ax, ay) of StackDev::dec4).  MatchFullSubmap lattices are wider than
// that tile (most candidate/point pairs lie outside the grid), so instead of every
// candidate walking all points, ONE WARP walks the points of a scan and adds each
// point's tile into the lattice:
//   * lane l owns the tile words f = l - 1 + 32*it (fully coalesced 128 B loads);
//   * the word offset between tile and lattice, (qx, qy), only changes when the point
//     moves to another coarse cell; consecutive beams mostly stay in one, so the
//     lanes accumulate in REGISTERS (packed u16 pairs) across a run of points and add
//     the run into the 32-bit lattice in shared memory when (qx, qy) changes (or
//     after 256 points, before the u16 lanes can overflow);
//   * copy k = qx & 3 of dec4 aligns tile words with lattice quads.
// Integer sums are order independent, so the result equals k_score_top_dense's.
#ifndef CSM_TILE_THREADS
#define CSM_TILE_THREADS 128
#endif
constexpr int kTileThreads = CSM_TILE_THREADS;
constexpr int kTileWarps = kTileThreads / 32;
template <int kIters>
__global__ void __launch_bounds__(kTileThreads)
k_score_top_tile(const JobDev* __restrict__ jobs, const ScanInfo* __restrict__ info,
                 const short2* __restrict__ dscan, int* __restrict__ top_sum,
                 const long long* __restrict__ scan_slot_base, int total_scans, int lat_ints) {
  extern __shared__ __align__(16) int s_lat_all[];
  __shared__ __align__(16) int s_key_all[kTileWarps][32], s_off_all[kTileWarps][32];
  constexpr int kNoKey = 0x7fff7fff;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* __restrict__ s_lat = s_lat_all + warp * lat_ints;  // [nyc][qr] quads of 4 ints
  int* __restrict__ s_key = s_key_all[warp];
  int* __restrict__ s_off = s_off_all[warp];
  const int nwarps = gridDim.x * kTileWarps;
  for (int sg = blockIdx.x * kTileWarps + warp; sg < total_scans; sg += nwarps) {
    const ScanInfo si = info[sg];
    const JobDev& jb = jobs[si.job];
    const StackDev& st = *jb.stack;
    const int h = st.depth - 1;
    const int s1 = (1 << h) - 1;
    const int id = st.dec_id[h], jd = st.dec_jd[h], ids = st.dec_ids[h];
    const long long lpad = st.dec_lpad[h];
    const uint8_t* __restrict__ dec = st.dec4[h] + 16;
    const int rw = ids >> 2;         // words per tile row
    const int W = jd * rw + 1;       // tile words incl. the one before the tile (f = -1)
    const int qr = (si.nxc + 3) >> 2;
    const int lat_n = qr * 4 * si.nyc;
    const short2* __restrict__ pts = dscan + jb.dscan_off +
                                   static_cast<long long>(sg - jb.scan_base) * jb.n;
    int* __restrict__ out = top_sum + scan_slot_base[sg];
    for (int t = lane; t < lat_n; t += 32) s_lat[t] = 0;
    unsigned a02[kIters], a13[kIters];
#pragma unroll
    for (int it = 0; it < kIters; ++it) a02[it] = a13[it] = 0u;
    int cur_key = kNoKey;      // (qy << 16) | (qx & 0xffff) of the current run
    int run = 0;
    const uint8_t* cur_base = dec;
    // adds the run's register sums into the lattice
    // tile row and byte column (before the copy shift k) of this lane's words
    int f_row[kIters], f_col[kIters];
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int f = lane + 32 * it - 1;
      f_row[it] = (f + rw) / rw - 1;               // floor(f / rw) for f >= -1
      f_col[it] = (f - f_row[it] * rw) << 2;
      if (f + 1 >= W) f_row[it] = -(1 << 20);      // not a tile word: never valid
    }
    auto flush = [&]() {
      __syncwarp();
      const int qx = static_cast<short>(cur_key & 0xffff), qy = cur_key >> 16;
      const int k = qx & 3;
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        int r = f_row[it];
        int c0 = f_col[it] + k;                 // tile column of the word's first byte
        if (c0 + 3 >= ids) { ++r; c0 -= ids; }  // the word continues in the next row
        const int j = r - qy, i0 = c0 - qx;     // lattice row / first lattice column (multiple of 4)
        if (c0 < id && static_cast<unsigned>(r) < static_cast<unsigned>(jd) &&
            static_cast<unsigned>(j) < static_cast<unsigned>(si.nyc) &&
            static_cast<unsigned>(i0) < static_cast<unsigned>(qr << 2)) {
          int4* cell = reinterpret_cast<int4*>(s_lat) + (j * qr + (i0 >> 2));
          int4 v = *cell;
          v.x += static_cast<int>(a02[it] & 0xffffu);
          v.y += static_cast<int>(a13[it] & 0xffffu);
          v.z += static_cast<int>(a02[it] >> 16);
          v.w += static_cast<int>(a13[it] >> 16);
          *cell = v;
        }
        a02[it] = a13[it] = 0u;
      }
    };
    for (int p0 = 0; p0 < jb.n; p0 += 32) {
      int my_off = 0, my_key = kNoKey;
      if (p0 + lane < jb.n) {
        const short2 c = pts[p0 + lane];
        const int bx = c.x + si.min_x + s1, by = c.y + si.min_y + s1;
        const int qx = bx >> h, qy = by >> h;  // floor division
        if (qx > -32000 && qx < 32000 && qy > -32000 && qy < 32000) {
          my_off = ((((by & s1) << h) | (bx & s1)) * jd) * ids;
          my_key = (qy << 16) | (qx & 0xffff);
        }
      }
      __syncwarp();
      s_key[lane] = my_key;
      s_off[lane] = my_off;
      __syncwarp();
      for (int t = 0; t < 32; t += 4) {   // points past the end carry kNoKey
        const int4 key = *reinterpret_cast<const int4*>(s_key + t);
        const int4 off = *reinterpret_cast<const int4*>(s_off + t);
        if (key.x == cur_key && key.y == cur_key && key.z == cur_key && key.w == cur_key &&
            run <= 252 && cur_key != kNoKey) {
          // four points of the current run: 4 * kIters independent loads in flight
          run += 4;
          unsigned w[4][kIters];
#pragma unroll
          for (int it = 0; it < kIters; ++it) {
            const bool on = lane + 32 * it < W;
            w[0][it] = on ? __ldg(reinterpret_cast<const unsigned*>(cur_base + off.x) + 32 * it) : 0u;
            w[1][it] = on ? __ldg(reinterpret_cast<const unsigned*>(cur_base + off.y) + 32 * it) : 0u;
            w[2][it] = on ? __ldg(reinterpret_cast<const unsigned*>(cur_base + off.z) + 32 * it) : 0u;
            w[3][it] = on ? __ldg(reinterpret_cast<const unsigned*>(cur_base + off.w) + 32 * it) : 0u;
          }
#pragma unroll
          for (int it = 0; it < kIters; ++it) {
            a02[it] += (__byte_perm(w[0][it], 0u, 0x4240) + __byte_perm(w[1][it], 0u, 0x4240)) +
                       (__byte_perm(w[2][it], 0u, 0x4240) + __byte_perm(w[3][it], 0u, 0x4240));
            a13[it] += (__byte_perm(w[0][it], 0u, 0x4341) + __byte_perm(w[1][it], 0u, 0x4341)) +
                       (__byte_perm(w[2][it], 0u, 0x4341) + __byte_perm(w[3][it], 0u, 0x4341));
          }
          continue;
        }
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
          const int ku = s_key[t + u], ou = s_off[t + u];
          if (ku != cur_key || run == 256) {
            if (run) flush();
            cur_key = ku;
            run = 0;
            cur_base = dec + (cur_key & 3) * lpad - 4 + 4 * lane;
          }
          if (ku == kNoKey) continue;  // point cannot hit the grid / past the end
          ++run;
          const unsigned* __restrict__ q = reinterpret_cast<const unsigned*>(cur_base + ou);
#pragma unroll
          for (int it = 0; it < kIters; ++it) {
            if (lane + 32 * it < W) {
              const unsigned w = __ldg(q + 32 * it);
              a02[it] += __byte_perm(w, 0u, 0x4240);  // bytes 0 and 2 in u16 lanes
              a13[it] += __byte_perm(w, 0u, 0x4341);  // bytes 1 and 3
            }
          }
        }
      }
    }
    if (run) flush();
    __syncwarp();
    const int slots = si.nxc * si.nyc;
    for (int o = lane; o < slots; o += 32) {
      const int i = o / si.nyc, j = o - i * si.nyc;
      out[o] = s_lat[((j * qr + (i >> 2)) << 2) + (i & 3)];
    }
    __syncwarp();
  }
}

// Generic list scoring (test hook + tie resolution): one warp per candidate.
struct ListCand { int scan; int xo, yo, level; };
__global__ void __launch_bounds__(256)
k_score_list(const JobDev* __restrict__ jobs, const ScanInfo* __restrict__ info,
             const short2* __restrict__ dscan, const ListCand* __restrict__ cands, int count,
             int* __restrict__ sums, float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= count) return;
  const ListCand c = cands[warp];
  const JobDev& jb = jobs[info[c.scan].job];
  const StackDev& st = *jb.stack;
  const int w1 = (1 << c.level) - 1;
  const uint8_t* __restrict__ g = st.level[c.level];
  const int wx = st.wx[c.level], wy = st.wy[c.level];
  const short2* __restrict__ pts = dscan + jb.dscan_off +
                                 static_cast<long long>(c.scan - jb.scan_base) * jb.n;
  int sum = 0;
  for (int p = lane; p < jb.n; p += 32) {
    const short2 q = pts[p];
    sum += GetValue(g, wx, wy, q.x + c.xo + w1, q.y + c.yo + w1);
  }
  sum = WarpSum(sum);
  if (lane == 0) {
    if (sums) sums[warp] = sum;
    if (scores) scores[warp] = ToScore(st, sum, jb.n);
  }
}

// Scores the (up to) four children of a node at level h-1 with one warp.
// Slot t = 2*ix + iy (ix, iy in {0,1}) is the child at offset (ix*half, iy*half);
// increasing t is the reference's generation order (x offset outer, y offset
// inner), and a slot is valid unless it is clipped by the scan's max bound
// (fast...2d.cc:352-367).  One aligned word of the child-window layout
// (StackDev::win) holds all four children values for a scan point; four points per
// lane are in flight per iteration (the loop is latency-bound otherwise).
// Returns the valid mask; sums[t] is 0 if invalid.
__device__ __forceinline__ unsigned ScoreChildren(const StackDev& st, const ScanInfo& si,
                                                  const short2* __restrict__ pts, int n, int xo,
                                                  int yo, int h, int lane, int sums[4]) {
  const int half = 1 << (h - 1);
  const bool x2 = !(xo + half > si.max_x);
  const bool y2 = !(yo + half > si.max_y);
  const int S1 = (1 << h) - 1;
  const unsigned* __restrict__ win = st.win[h];
  const int jd = st.win_jd[h], ids = st.win_ids[h];
  const long long per_phase = WinPhaseWords(jd, ids);
  const int bx = xo + half - 1, by = yo + half - 1;
  int s00 = 0, s01 = 0, s10 = 0, s11 = 0;
  constexpr int kU = 4;
  for (int p0 = lane; p0 < n; p0 += 32 * kU * 64) {
    // packed u16 pairs: r0 = (ix 0, ix 1) of iy 0, r1 = same of iy 1; at most
    // 64 * kU = 256 points per lane between flushes (256 * 255 < 2^16)
    unsigned r0 = 0, r1 = 0;
    const int pend = min(n, p0 + 32 * kU * 64);
    for (int p = p0; p < pend; p += 32 * kU) {
      short2 q[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int pp = p + 32 * u;
        q[u] = pp < pend ? pts[pp] : make_short2(0, 0);
      }
      unsigned w[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int lx = q[u].x + bx, ly = q[u].y + by;
        const int Qx = (lx >> h) + 1, Qy = (ly >> h) + 1;
        w[u] = 0u;
        if (p + 32 * u < pend && static_cast<unsigned>(Qx) < static_cast<unsigned>(ids) &&
            static_cast<unsigned>(Qy) < static_cast<unsigned>(jd))
          w[u] = __ldg(win + (static_cast<long long>((ly & S1) << h | (lx & S1)) * per_phase +
                              WinCell(Qx, Qy, ids)));
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        r0 += __byte_perm(w[u], 0u, 0x4140);
        r1 += __byte_perm(w[u], 0u, 0x4342);
      }
    }
    s00 += r0 & 0xffffu;
    s10 += r0 >> 16;
    s01 += r1 & 0xffffu;
    s11 += r1 >> 16;
  }
  sums[0] = WarpSum(s00);
  sums[1] = y2 ? WarpSum(s01) : 0;
  sums[2] = x2 ? WarpSum(s10) : 0;
  sums[3] = (x2 && y2) ? WarpSum(s11) : 0;
  return 1u | (y2 ? 2u : 0u) | (x2 ? 4u : 0u) | ((x2 && y2) ? 8u : 0u);
}

// scan -> job and scan -> first lowest-resolution slot, one CTA per job.
__global__ void k_scan_tables(const JobDev* __restrict__ jobs, int* __restrict__ scan_job,
                              long long* __restrict__ scan_slot_base, unsigned* __restrict__ lb,
                              int* __restrict__ job_best) {
  const JobDev& d = jobs[blockIdx.x];
  if (threadIdx.x == 0) {
    // the bound starts at min_score: only scores > min_score are ever accepted (fast...2d.cc:253)
    lb[blockIdx.x] = FloatToOrdered(d.min_score);
    job_best[blockIdx.x] = 0;
  }
  for (int k = threadIdx.x; k < d.num_scans; k += blockDim.x) {
    scan_job[d.scan_base + k] = blockIdx.x;
    scan_slot_base[d.scan_base + k] = d.top_off + static_cast<long long>(k) * d.cap;
  }
}

// Per-job maximum of the lowest-resolution sums (one warp per scan -> atomicMax).
__global__ void __launch_bounds__(256)
k_job_best(const ScanInfo* __restrict__ info, const int* __restrict__ top_sum,
           const long long* __restrict__ scan_slot_base, int total_scans,
           int* __restrict__ job_best) {
  const int lane = threadIdx.x & 31;
  const int sg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (sg >= total_scans) return;
  const ScanInfo si = info[sg];
  const int slots = si.nxc * si.nyc;
  const int* __restrict__ ts = top_sum + scan_slot_base[sg];
  int best = 0;
  for (int s = lane; s < slots; s += 32) best = max(best, ts[s]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (lane == 0) atomicMax(&job_best[si.job], best);
}

// Greedy dives: one warp per scan starts at the scan's best lowest-resolution
// candidate and follows the best child down to a leaf.  Every leaf score is a
// valid lower bound of the job's optimum; the maximum over all scans seeds the
// branch-and-bound (the reference's DFS gets the same bound from its first dive,
// fast...2d.cc:335-378, only later).
__global__ void __launch_bounds__(256)
k_dive(const JobDev* __restrict__ jobs, const ScanInfo* __restrict__ info,
       const short2* __restrict__ dscan, const int* __restrict__ top_sum,
       const long long* __restrict__ scan_slot_base, int total_scans,
       const int* __restrict__ job_best, float dive_ratio,
       unsigned* __restrict__ lb, unsigned long long* __restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const int sg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (sg >= total_scans) return;
  const ScanInfo si = info[sg];
  const JobDev& jb = jobs[si.job];
  const StackDev& st = *jb.stack;
  const int slots = si.nxc * si.nyc;
  const int* __restrict__ ts = top_sum + scan_slot_base[sg];
  int best = -1, best_slot = 0;
  for (int s = lane; s < slots; s += 32) {
    const int v = ts[s];
    if (v > best) { best = v; best_slot = s; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int os = __shfl_xor_sync(0xffffffffu, best_slot, o);
    if (ov > best || (ov == best && os < best_slot)) { best = ov; best_slot = os; }
  }
  if (!(ToScore(st, best, jb.n) > jb.min_score)) return;
  // only scans whose best bound is close to the job's best bound are worth a dive
  // (any subset keeps the bound valid; this one keeps it tight at a fraction of the cost)
  if (static_cast<float>(best) < dive_ratio * static_cast<float>(job_best[si.job])) return;
  int h = st.depth - 1;
  const int i = best_slot / si.nyc, jy = best_slot - i * si.nyc;
  int xo = si.min_x + (i << h), yo = si.min_y + (jy << h);
  const short2* __restrict__ pts = dscan + jb.dscan_off +
                                 static_cast<long long>(sg - jb.scan_base) * jb.n;
  int leaf_sum = best;
  unsigned long long scored = 0;
  while (h > 0) {
    int sums[4];
    const unsigned valid = ScoreChildren(st, si, pts, jb.n, xo, yo, h, lane, sums);
    scored += __popc(valid);
    const int half = 1 << (h - 1);
    int b = 0, bs = sums[0];
#pragma unroll
    for (int t = 1; t < 4; ++t)
      if (((valid >> t) & 1u) && sums[t] > bs) { b = t; bs = sums[t]; }
    xo += (b >> 1) * half;
    yo += (b & 1) * half;
    leaf_sum = bs;
    --h;
  }
  if (lane == 0) {
    const float score = ToScore(st, leaf_sum, jb.n);
    if (score > jb.min_score) atomicMax(&lb[si.job], FloatToOrdered(score));
    atomicAdd(&counters[0], scored);
    atomicAdd(&counters[2], 1ull);
  }
}

// Pushes every lowest-resolution candidate that can still contain the optimum
// (score > min_score and score >= current bound) onto the top-level queue.  One
// CTA per scan walks the scan's lattice ROW by ROW (y outer, x inner) and appends
// the survivors in that order with a block-wide ordered compaction, so queue
// neighbours are lattice neighbours along x — the direction the lattice branch
// kernel coalesces over.
__global__ void __launch_bounds__(256)
k_filter_top(const JobDev* __restrict__ jobs, const ScanInfo* __restrict__ info,
             const int* __restrict__ top_sum, const long long* __restrict__ scan_slot_base,
             int total_scans, const unsigned* __restrict__ lb, Node* __restrict__ queue,
             int* __restrict__ qcount, int qcap, int* __restrict__ overflow) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int sg = blockIdx.x; sg < total_scans; sg += gridDim.x) {
    const ScanInfo si = info[sg];
    const JobDev& jb = jobs[si.job];
    const StackDev& st = *jb.stack;
    const int h = st.depth - 1;
    const int slots = si.nxc * si.nyc;
    const float bound = OrderedToFloat(lb[si.job]);
    const int* __restrict__ ts = top_sum + scan_slot_base[sg];
    for (int q0 = 0; q0 < slots; q0 += blockDim.x) {
      const int q = q0 + threadIdx.x;  // row-major lattice index
      bool keep = false;
      float score = 0.f;
      int i = 0, jy = 0;
      if (q < slots) {
        jy = q / si.nxc;
        i = q - jy * si.nxc;
        score = ToScore(st, ts[i * si.nyc + jy], jb.n);
        keep = score > jb.min_score && score >= bound;
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_warp[warp] = __popc(m);
      __syncthreads();
      if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) { const int c = s_warp[w]; s_warp[w] = tot; tot += c; }
        s_base = tot ? atomicAdd(qcount, tot) : 0;
      }
      __syncthreads();
      if (keep) {
        const int idx = s_base + s_warp[warp] + __popc(m & ((1u << lane) - 1));
        if (idx < qcap)
          queue[idx] = Node{sg, si.min_x + (i << h), si.min_y + (jy << h), score};
        else
          *overflow = 1;
      }
      __syncthreads();
    }
  }
}

// ---- device-resident control of the level loop ---------------------------------
// The branch-and-bound frontier sizes never visit the host between levels: every
// kernel of a level reads its chunk [start, start + n) of queue[h] and the kernel
// form (`mode`) from this block, which k_level_begin fills from the device-side queue
// counts.  Launch grids are sized for the worst case (or grid-stride), so a whole
// match batch is ONE stream of launches followed by ONE synchronisation.
enum : int {
  kCtlLeaf = 16,      // leaves recorded so far
  kCtlBest = 17,      // optimal leaves after compaction
  kCtlOverflow = 20,  // a queue / leaf buffer was too small
  kCtlItems = 24,     // work items of the current lattice launch
  kCtlStart = 25,     // first node of the current chunk in queue[h]
  kCtlCount = 26,     // nodes of the current chunk
  kCtlMode = 27,      // 0 = nothing to do, 1 = warp per parent, 2 = scan-grouped lattice
  kCtlInts = 32
};

// Branch step: one warp per parent node of level h.  Scores its children at
// level h-1, then either pushes the survivors to the next queue (h-1 >= 1) or,
// at h-1 == 0, raises the job's bound and records the leaf.
__device__ __forceinline__ void ExpandParentWarp(
    const JobDev* __restrict__ jobs, const ScanInfo* __restrict__ info,
    const short2* __restrict__ dscan, const Node nd, int h, unsigned* __restrict__ lb,
    Node* __restrict__ next, int* __restrict__ next_count, int next_cap,
    Node* __restrict__ leaves, int* __restrict__ leaf_count, int leaf_cap,
    int* __restrict__ overflow, unsigned long long* __restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const ScanInfo si = info[nd.scan];
  const JobDev& jb = jobs[si.job];
  const StackDev& st = *jb.stack;
  // bound may have risen since the node was queued
  if (!(nd.score >= OrderedToFloat(lb[si.job]))) return;
  const short2* __restrict__ pts = dscan + jb.dscan_off +
                                 static_cast<long long>(nd.scan - jb.scan_base) * jb.n;
  int sums[4];
  const unsigned valid = ScoreChildren(st, si, pts, jb.n, nd.xo, nd.yo, h, lane, sums);
  if (lane != 0) return;
  atomicAdd(&counters[0], (unsigned long long)__popc(valid));
  atomicAdd(&counters[1], 1ull);
  const int half = 1 << (h - 1);
  float sc[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) sc[t] = ToScore(st, sums[t], jb.n);
  if (h - 1 == 0) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (!((valid >> t) & 1u) || !(sc[t] > jb.min_score)) continue;
      const unsigned o = FloatToOrdered(sc[t]);
      const unsigned old = atomicMax(&lb[si.job], o);
      if (o >= old) {
        const int idx = atomicAdd(leaf_count, 1);
        if (idx < leaf_cap)
          leaves[idx] = Node{nd.scan, nd.xo + (t >> 1) * half, nd.yo + (t & 1) * half, sc[t]};
        else
          *overflow = 1;
      }
    }
  } else {
    const float bound = OrderedToFloat(lb[si.job]);
    unsigned keep = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (((valid >> t) & 1u) && sc[t] > jb.min_score && sc[t] >= bound) keep |= 1u << t;
    if (keep) {
      int idx = atomicAdd(next_count, __popc(keep));
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if ((keep >> t) & 1u) {
          if (idx < next_cap)
            next[idx] = Node{nd.scan, nd.xo + (t >> 1) * half, nd.yo + (t & 1) * half, sc[t]};
          else
            *overflow = 1;
          ++idx;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
k_expand(const JobDev* __restrict__ jobs, const ScanInfo* __restrict__ info,
         const short2* __restrict__ dscan, const Node* __restrict__ queue,
         const int* __restrict__ ctl, int h,
         unsigned* __restrict__ lb, Node* __restrict__ next, int* __restrict__ next_count,
         int next_cap, Node* __restrict__ leaves, int* __restrict__ leaf_count, int leaf_cap,
         int* __restrict__ overflow, unsigned long long* __restrict__ counters) {
  if (ctl[kCtlMode] != 1) return;
  const int count = ctl[kCtlCount];
  const Node* __restrict__ parents = queue + ctl[kCtlStart];
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; warp < count; warp += warps)
    ExpandParentWarp(jobs, info, dscan, parents[warp], h, lb, next, next_count, next_cap, leaves,
                     leaf_count, leaf_cap, overflow, counters);
}

// ---- scan-grouped branch step ------------------------------------------------
// The frontier nodes of ONE rotated scan lie on that scan's lattice of stride
// S = 2^h, so for one scan point the child-window words (StackDev::win) of lattice
// neighbours are neighbouring words of one phase tile.  Parents are first grouped by
// scan (counting sort), then one WARP handles a work item of up to 32 parents of a
// scan: point descriptors are staged once per item in shared memory and every lane
// fetches the 2 x 2 children of its parent with ONE aligned 32-bit load per point (vs.
// a point load + a scattered word per lane in the warp-per-parent form).
struct WorkItem { int scan, start, count; };

__global__ void k_level_begin(int* __restrict__ ctl, int h, int chunk_cap, int lattice_min) {
  if (threadIdx.x != 0) return;
  const int have = ctl[h];
  const int n = min(have, chunk_cap);
  ctl[h] = have - n;          // the chunk is taken from the END of the queue
  ctl[kCtlStart] = have - n;
  ctl[kCtlCount] = n;
  ctl[kCtlItems] = 0;
  ctl[kCtlMode] = n == 0 ? 0 : (n >= lattice_min ? 2 : 1);
}

__global__ void __launch_bounds__(256)
k_q_count(const Node* __restrict__ queue, const int* __restrict__ ctl,
          int* __restrict__ scan_cnt) {
  if (ctl[kCtlMode] != 2) return;
  const Node* __restrict__ nodes = queue + ctl[kCtlStart];
  const int count = ctl[kCtlCount];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    atomicAdd(&scan_cnt[nodes[i].scan], 1);
}

// Exclusive prefix sums of scan_cnt (node offsets) and of ceil(cnt / 32) (work
// items) in three steps: sums of 1024-scan blocks, a single-CTA scan of those block
// sums, and per-block scans that also emit the work items.
__device__ __forceinline__ void BlockScan2(int& ia, int& ib, int* s_wa, int* s_wb) {
  // inclusive scan of (ia, ib) over the 1024 threads of the CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int va = __shfl_up_sync(0xffffffffu, ia, o), vb = __shfl_up_sync(0xffffffffu, ib, o);
    if (lane >= o) { ia += va; ib += vb; }
  }
  if (lane == 31) { s_wa[warp] = ia; s_wb[warp] = ib; }
  __syncthreads();
  if (warp == 0) {
    int wa = s_wa[lane], wb = s_wb[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int va = __shfl_up_sync(0xffffffffu, wa, o), vb = __shfl_up_sync(0xffffffffu, wb, o);
      if (lane >= o) { wa += va; wb += vb; }
    }
    s_wa[lane] = wa;
    s_wb[lane] = wb;
  }
  __syncthreads();
  if (warp) { ia += s_wa[warp - 1]; ib += s_wb[warp - 1]; }
}

__global__ void __launch_bounds__(1024)
k_q_block_sums(const int* __restrict__ ctl, const int* __restrict__ scan_cnt, int total_scans,
               int* __restrict__ part_a, int* __restrict__ part_b) {
  __shared__ int s_wa[32], s_wb[32];
  if (ctl[kCtlMode] != 2) return;
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const int cnt = i < total_scans ? scan_cnt[i] : 0;
  int ia = cnt, ib = (cnt + 31) >> 5;
  BlockScan2(ia, ib, s_wa, s_wb);
  if (threadIdx.x == 1023) { part_a[blockIdx.x] = ia; part_b[blockIdx.x] = ib; }
}

// in-place exclusive scan of the block sums; single CTA, every thread owns a
// contiguous segment.  out[0] = total number of work items.
__global__ void __launch_bounds__(1024)
k_q_scan_parts(const int* __restrict__ ctl, int* __restrict__ part_a, int* __restrict__ part_b,
               int nb, int* __restrict__ out) {
  __shared__ int s_wa[32], s_wb[32];
  if (ctl[kCtlMode] != 2) return;
  const int seg = (nb + 1023) >> 10;
  const int lo = min(nb, static_cast<int>(threadIdx.x) * seg), hi = min(nb, lo + seg);
  int a = 0, b = 0;
  for (int i = lo; i < hi; ++i) { a += part_a[i]; b += part_b[i]; }
  int ia = a, ib = b;
  BlockScan2(ia, ib, s_wa, s_wb);
  int ea = ia - a, eb = ib - b;
  for (int i = lo; i < hi; ++i) {
    const int va = part_a[i], vb = part_b[i];
    part_a[i] = ea;
    part_b[i] = eb;
    ea += va;
    eb += vb;
  }
  if (threadIdx.x == 1023) out[0] = ib;
}

__global__ void __launch_bounds__(1024)
k_q_finish(const int* __restrict__ ctl, const int* __restrict__ scan_cnt, int total_scans,
           const int* __restrict__ part_a, const int* __restrict__ part_b,
           int* __restrict__ scan_off, WorkItem* __restrict__ items) {
  __shared__ int s_wa[32], s_wb[32];
  if (ctl[kCtlMode] != 2) return;
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const int cnt = i < total_scans ? scan_cnt[i] : 0;
  int ia = cnt, ib = (cnt + 31) >> 5;
  BlockScan2(ia, ib, s_wa, s_wb);
  if (i >= total_scans) return;
  const int off = part_a[blockIdx.x] + ia - cnt;
  const int item0 = part_b[blockIdx.x] + ib - ((cnt + 31) >> 5);
  scan_off[i] = off;
  for (int k = 0; k * 32 < cnt; ++k)
    items[item0 + k] = WorkItem{i, off + k * 32, min(32, cnt - k * 32)};
}

// Stable within every 32-node run: lanes holding nodes of the same scan get
// consecutive slots in lane order (one atomic per scan per warp), so the x-ordered
// runs produced by the push code survive the grouping.
__global__ void __launch_bounds__(256)
k_q_scatter(const Node* __restrict__ queue, const int* __restrict__ ctl,
            const int* __restrict__ scan_off, int* __restrict__ cursor,
            Node* __restrict__ sorted) {
  if (ctl[kCtlMode] != 2) return;
  const Node* __restrict__ nodes = queue + ctl[kCtlStart];
  const int count = ctl[kCtlCount];
  const int lane = threadIdx.x & 31;
  const int rounded = (count + 31) & ~31;  // whole warps take part in the match
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += gridDim.x * blockDim.x) {
    const bool ok = i < count;
    Node nd = Node{-1 - lane, 0, 0, 0.f};
    if (ok) nd = nodes[i];
    const unsigned peers = __match_any_sync(0xffffffffu, nd.scan);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (ok && lane == leader) base = atomicAdd(&cursor[nd.scan], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (ok) sorted[scan_off[nd.scan] + base + __popc(peers & ((1u << lane) - 1))] = nd;
  }
}

#ifndef CSM_LAT_MINB
#define CSM_LAT_MINB 9   // CTAs per SM the register allocation aims for
#endif
#ifndef CSM_LAT_THREADS
#define CSM_LAT_THREADS 128
#endif
constexpr int kLatThreads = CSM_LAT_THREADS;   // every warp works on its own item
constexpr int kLatChunk = 256;
template <int kUnroll>
__global__ void __launch_bounds__(kLatThreads, CSM_LAT_MINB)
k_expand_lattice(const JobDev* __restrict__ jobs, const ScanInfo* __restrict__ info,
                 const short2* __restrict__ dscan, const Node* __restrict__ sorted,
                 const WorkItem* __restrict__ items, const int* __restrict__ num_items, int h,
                 unsigned* __restrict__ lb,
                 Node* __restrict__ next, int* __restrict__ next_count, int next_cap,
                 Node* __restrict__ leaves, int* __restrict__ leaf_count, int leaf_cap,
                 int* __restrict__ overflow, unsigned long long* __restrict__ counters) {
  __shared__ __align__(16) int2 s_all[kLatThreads / 32][kLatChunk];  // {window index of lattice origin, Qy << 16 | Qx}
  // (8 B per point: a smaller shared-memory carve-out leaves more L1 for the tiles)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // (one item per warp and a grid sized for the worst case: a capped grid with an item
  // loop measured 6 % slower — the hardware CTA scheduler balances uneven items better)
  const int item = blockIdx.x * (kLatThreads / 32) + warp;
  if (item >= *num_items) return;
  int2* s_pt = s_all[warp];
  const WorkItem it = items[item];
  // With few parents, G = 2^lg lanes share one parent and split the scan points
  // (lane = parent * G + sub, sub-lane `sub` takes the point pairs sub, sub + G, ...).
  int lg = 0;
  while ((it.count << (lg + 1)) <= 32) ++lg;
  const int G = 1 << lg;
  const int pidx = lane >> lg, sub = lane & (G - 1);
  const ScanInfo si = info[it.scan];
  const JobDev& jb = jobs[si.job];
  const StackDev& st = *jb.stack;
  const int lv = h - 1;            // level of the children
  const int s = 1 << lv;           // children stride = half the parents' stride
  const int S1 = (1 << h) - 1;
  const int jd = st.win_jd[h], ids = st.win_ids[h];
  const unsigned* __restrict__ win = st.win[h];
#ifdef CSM_WIN_TILED
  const int per_phase = static_cast<int>(WinPhaseWords(jd, ids));
#endif
  const short2* __restrict__ pts = dscan + jb.dscan_off +
                                   static_cast<long long>(it.scan - jb.scan_base) * jb.n;
  const bool active = pidx < it.count;
  Node nd = Node{it.scan, si.min_x, si.min_y, 0.f};
  if (active) nd = sorted[it.start + pidx];
  // the bound may have risen since the node was queued
  const bool live = active && nd.score >= OrderedToFloat(lb[si.job]);
  const int i0 = (nd.xo - si.min_x) >> h, j0 = (nd.yo - si.min_y) >> h;  // parent lattice coords
  const int toff = j0 * ids + i0;
  const bool x2 = !(nd.xo + s > si.max_x), y2 = !(nd.yo + s > si.max_y);
  unsigned sum0 = 0, sum1 = 0, sum2 = 0, sum3 = 0;  // slots 2*ix+iy: 00, 01, 10, 11
  for (int p0 = 0; p0 < jb.n; p0 += kLatChunk) {
    __syncwarp();
    for (int t = lane; t < kLatChunk; t += 32) {
      const int p = p0 + t;
      int2 d = make_int2(0, static_cast<int>(0x80008000u));  // Qx = Qy = -32768: never in range
      if (p < jb.n) {
        const short2 c = pts[p];
        const int bx = c.x + si.min_x + s - 1, by = c.y + si.min_y + s - 1;
        const int qx = (bx >> h) + 1, qy = (by >> h) + 1;
        const int ax = bx & S1, ay = by & S1;
        if (qx > -32000 && qx < 32000 && qy > -32000 && qy < 32000)
#ifdef CSM_WIN_TILED
          d = make_int2(((ay << h) | ax) * per_phase, (qy << 16) | (qx & 0xffff));
#else
          d = make_int2((((ay << h) | ax) * jd + qy) * ids + qx, (qy << 16) | (qx & 0xffff));
#endif
      }
      s_pt[t] = d;
    }
    __syncwarp();
    if (live) {
      const int cnt = min(kLatChunk, jb.n - p0);
      unsigned r0 = 0, r1 = 0;  // packed u16 pairs: (ix 0, ix 1) of iy 0 / iy 1
      // two staged points per 16-byte shared load (entries past cnt never pass the range test)
#pragma unroll(kUnroll / 2)
      for (int t = 2 * sub; t < cnt; t += 2 * G) {
        const int4 d = *reinterpret_cast<const int4*>(s_pt + t);
        const int Ja = (d.y >> 16) + j0, Ia = static_cast<short>(d.y & 0xffff) + i0;
        const int Jb = (d.w >> 16) + j0, Ib = static_cast<short>(d.w & 0xffff) + i0;
        if (static_cast<unsigned>(Ia) < static_cast<unsigned>(ids) &&
            static_cast<unsigned>(Ja) < static_cast<unsigned>(jd)) {
#ifdef CSM_WIN_TILED
          const unsigned w = __ldg(win + (d.x + WinCell(Ia, Ja, ids)));
#else
          const unsigned w = __ldg(win + (d.x + toff));
#endif
          r0 += __byte_perm(w, 0u, 0x4140);
          r1 += __byte_perm(w, 0u, 0x4342);
        }
        if (static_cast<unsigned>(Ib) < static_cast<unsigned>(ids) &&
            static_cast<unsigned>(Jb) < static_cast<unsigned>(jd)) {
#ifdef CSM_WIN_TILED
          const unsigned w = __ldg(win + (d.z + WinCell(Ib, Jb, ids)));
#else
          const unsigned w = __ldg(win + (d.z + toff));
#endif
          r0 += __byte_perm(w, 0u, 0x4140);
          r1 += __byte_perm(w, 0u, 0x4342);
        }
      }
      sum0 += r0 & 0xffffu;   // (ix 0, iy 0)
      sum2 += r0 >> 16;       // (ix 1, iy 0)
      sum1 += r1 & 0xffffu;   // (ix 0, iy 1)
      sum3 += r1 >> 16;       // (ix 1, iy 1)
    }
  }
  // totals of the G sub-lanes (all lanes of the warp take part)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v0 = __shfl_xor_sync(0xffffffffu, sum0, o), v1 = __shfl_xor_sync(0xffffffffu, sum1, o);
    const unsigned v2 = __shfl_xor_sync(0xffffffffu, sum2, o), v3 = __shfl_xor_sync(0xffffffffu, sum3, o);
    if (o < G) { sum0 += v0; sum1 += v1; sum2 += v2; sum3 += v3; }
  }
  const bool lead = live && sub == 0;   // one lane per parent carries on
  const unsigned valid = lead ? (1u | (y2 ? 2u : 0u) | (x2 ? 4u : 0u) | ((x2 && y2) ? 8u : 0u)) : 0u;
  if (lead) {
    atomicAdd(&counters[0], (unsigned long long)__popc(valid));
    atomicAdd(&counters[1], 1ull);
  }
  const int sums[4] = {static_cast<int>(sum0), static_cast<int>(sum1), static_cast<int>(sum2),
                       static_cast<int>(sum3)};
  float sc[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) sc[t] = ToScore(st, sums[t], jb.n);
  if (lv == 0) {
    if (!lead) return;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (!((valid >> t) & 1u) || !(sc[t] > jb.min_score)) continue;
      const unsigned o = FloatToOrdered(sc[t]);
      const unsigned old = atomicMax(&lb[si.job], o);
      if (o >= old) {
        const int idx = atomicAdd(leaf_count, 1);
        if (idx < leaf_cap)
          leaves[idx] = Node{nd.scan, nd.xo + (t >> 1) * s, nd.yo + (t & 1) * s, sc[t]};
        else
          *overflow = 1;
      }
    }
    return;
  }
  // Survivors are appended row by row (all lanes' children of lattice row 2*j0, then
  // row 2*j0+1; within a row in lane order, x ascending) with one atomic per warp.
  const float bound = OrderedToFloat(lb[si.job]);
  unsigned keep = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t)
    if (((valid >> t) & 1u) && sc[t] > jb.min_score && sc[t] >= bound) keep |= 1u << t;
  // slot t = 2*ix + iy
  const unsigned m00 = __ballot_sync(0xffffffffu, keep & 1u), m10 = __ballot_sync(0xffffffffu, keep & 4u);
  const unsigned m01 = __ballot_sync(0xffffffffu, keep & 2u), m11 = __ballot_sync(0xffffffffu, keep & 8u);
  const int row0 = __popc(m00) + __popc(m10), row1 = __popc(m01) + __popc(m11);
  int base = 0;
  if (lane == 0 && row0 + row1) base = atomicAdd(next_count, row0 + row1);
  base = __shfl_sync(0xffffffffu, base, 0);
  const unsigned lt = (1u << lane) - 1u;
  int p = base + __popc(m00 & lt) + __popc(m10 & lt);
  if (keep & 1u) {
    if (p < next_cap) next[p] = Node{nd.scan, nd.xo, nd.yo, sc[0]}; else *overflow = 1;
    ++p;
  }
  if (keep & 4u) {
    if (p < next_cap) next[p] = Node{nd.scan, nd.xo + s, nd.yo, sc[2]}; else *overflow = 1;
  }
  p = base + row0 + __popc(m01 & lt) + __popc(m11 & lt);
  if (keep & 2u) {
    if (p < next_cap) next[p] = Node{nd.scan, nd.xo, nd.yo + s, sc[1]}; else *overflow = 1;
    ++p;
  }
  if (keep & 8u) {
    if (p < next_cap) next[p] = Node{nd.scan, nd.xo + s, nd.yo + s, sc[3]}; else *overflow = 1;
  }
}

// Keeps the leaves whose score equals their job's final optimum.  The leaf count is
// read on the device (grid-stride).
__global__ void __launch_bounds__(256)
k_compact_leaves(const ScanInfo* __restrict__ info, const Node* __restrict__ in,
                 const int* __restrict__ count_ptr, int cap_in, const unsigned* __restrict__ lb,
                 Node* __restrict__ out, int* __restrict__ out_count, int cap,
                 int* __restrict__ overflow) {
  const int count = min(*count_ptr, cap_in);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const Node nd = in[i];
    if (FloatToOrdered(nd.score) >= lb[info[nd.scan].job]) {
      const int idx = atomicAdd(out_count, 1);
      if (idx < cap) out[idx] = nd;
      else *overflow = 1;
    }
  }
}

// depth 1: the lowest resolution IS the leaf level (fast...2d.cc:339-343): every queued
// candidate raises its job's bound and becomes a leaf.
__global__ void __launch_bounds__(256)
k_top_as_leaves(const ScanInfo* __restrict__ info, const Node* __restrict__ queue,
                int* __restrict__ ctl, unsigned* __restrict__ lb, Node* __restrict__ leaves,
                int leaf_cap) {
  const int count = ctl[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const Node nd = queue[i];
    atomicMax(&lb[info[nd.scan].job], FloatToOrdered(nd.score));
    if (i < leaf_cap) leaves[i] = nd;
    else ctl[kCtlOverflow] = 1;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl[kCtlLeaf] = min(count, leaf_cap);
}

}  // namespace csm

// ===========================================================================
// Host side
// ===========================================================================
using namespace csm;

namespace {

// mapping/value_conversion_tables.cc:29-51 and fast...2d.cc:97-98,110-111,163-169:
// lut[v] = lround(((1 - |cost(v)|) - min_score) * (255 / (max_score - min_score)))
void BuildLut(float min_cost, float max_cost, uint8_t* lut, float* min_score, float* max_score) {
  const float lo = 1.f - max_cost;  // min_score_
  const float hi = 1.f - min_cost;  // max_score_
  *min_score = lo;
  *max_score = hi;
  const float kScale = (max_cost - min_cost) / 32766.f;
  for (int v = 0; v < 65536; ++v) {
    const uint16_t value = static_cast<uint16_t>(v) & static_cast<uint16_t>(~(1u << 15));
    float cost;
    if (value == 0) cost = max_cost;  // unknown -> max_correspondence_cost (grid_2d.cc:69-71)
    else cost = value * kScale + (min_cost - kScale);
    const float probability = 1.f - std::abs(cost);
    const long q = std::lround((probability - lo) * (255.f / (hi - lo)));
    lut[v] = static_cast<uint8_t>(q < 0 ? 0 : (q > 255 ? 255 : q));
  }
}

int DivUp(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace

// Fills every layout of the stack (levels, decimated copies, child windows) from the
// grid's cells; caller holds ctx->mu.  Shared by csm_stack2d_create and csm_stack2d_update.
static csm_status BuildStack2D(csm_stack2d* st, const uint16_t* cells) {
  Ctx* ctx = st->ctx;
  const StackDev& h = st->h;
  const int nx = h.nx, ny = h.ny, depth = h.depth, top = depth - 1;
  std::vector<uint8_t> lut(65536);
  float lo, hi;
  BuildLut(st->min_cost, st->max_cost, lut.data(), &lo, &hi);
  const size_t* win_off = st->win_off;
  // upload cells + LUT into scratch
  DevBuf& d_cells = ctx->D("stack_cells");
  DevBuf& d_lut = ctx->D("stack_lut");
  const size_t ncell = static_cast<size_t>(nx) * ny;
  CSM_TRY(d_cells.Reserve(ncell * sizeof(uint16_t)));
  CSM_TRY(d_lut.Reserve(65536));
  CSM_CUDA(cudaMemcpyAsync(d_cells.p, cells, ncell * sizeof(uint16_t), cudaMemcpyHostToDevice,
                           ctx->stream));
  CSM_CUDA(cudaMemcpyAsync(d_lut.p, lut.data(), 65536, cudaMemcpyHostToDevice, ctx->stream));
  k_stack_level0<<<DivUp(ncell, 256), 256, 0, ctx->stream>>>(
      d_cells.as<uint16_t>(), d_lut.as<uint8_t>(), st->d_levels + st->level_off[0],
      static_cast<int>(ncell));
  CSM_LAUNCH_CHECK();
  for (int l = 1; l < depth; ++l) {
    dim3 block(32, 8), grid(DivUp(h.wx[l], 32), DivUp(h.wy[l], 8));
    k_stack_double<<<grid, block, 0, ctx->stream>>>(st->d_levels + st->level_off[l - 1],
                                                    h.wx[l - 1], h.wy[l - 1],
                                                    st->d_levels + st->level_off[l], h.wx[l],
                                                    h.wy[l], 1 << (l - 1));
    CSM_LAUNCH_CHECK();
  }
  k_stack_decimate4<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(
      h.level[top], h.wx[top], h.wy[top], top, st->d_dec, h.dec_lpad[top], h.dec_id[top],
      h.dec_jd[top], h.dec_ids[top]);
  CSM_LAUNCH_CHECK();
  for (int l = 1; l < depth; ++l) {
    k_stack_window<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(
        h.level[l - 1], h.wx[l - 1], h.wy[l - 1], l, st->d_win + win_off[l], h.win_jd[l],
        h.win_ids[l]);
    CSM_LAUNCH_CHECK();
  }
  return CSM_OK;
}

extern "C" {

csm_status csm_stack2d_create(const uint16_t* cells, int32_t nx, int32_t ny, double resolution,
                              double max_x, double max_y, float min_cost, float max_cost,
                              int32_t depth, int32_t device, csm_stack2d** out) {
  CSM_REQUIRE(out != nullptr && cells != nullptr, "null pointer");
  CSM_REQUIRE(nx >= 1 && ny >= 1, "cell limits must be >= 1");  // fast...2d.cc:100-102
  CSM_REQUIRE(depth >= 1 && depth <= kMaxDepth, "branch_and_bound_depth out of range");  // :174
  CSM_REQUIRE(resolution > 0., "resolution must be > 0");
  CSM_REQUIRE(min_cost < max_cost, "min cost must be < max cost");  // grid_2d.cc:73
  CSM_REQUIRE(static_cast<long long>(nx) + (1 << (depth - 1)) < 32000 &&
              static_cast<long long>(ny) + (1 << (depth - 1)) < 32000, "grid too large");
  Ctx* ctx;
  CSM_TRY(GetCtx(device, &ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  CSM_CUDA(cudaSetDevice(device));
  std::unique_ptr<csm_stack2d> st(new csm_stack2d);
  st->ctx = ctx;
  st->min_cost = min_cost;
  st->max_cost = max_cost;
  StackDev& h = st->h;
  std::memset(&h, 0, sizeof(h));
  h.nx = nx;
  h.ny = ny;
  h.depth = depth;
  h.resolution = resolution;
  h.max_x = max_x;
  h.max_y = max_y;
  std::vector<uint8_t> lut(65536);
  BuildLut(min_cost, max_cost, lut.data(), &h.min_score, &h.max_score);
  h.k255 = (h.max_score - h.min_score) / 255.f;
  size_t total = 0;
  for (int l = 0; l < depth; ++l) {
    const int w = 1 << l;
    h.wx[l] = nx + w - 1;
    h.wy[l] = ny + w - 1;
    st->level_off[l] = total;
    total += (static_cast<size_t>(h.wx[l]) * h.wy[l] + 255) / 256 * 256;
  }
  CSM_CUDA(cudaMalloc(&st->d_levels, total));
  for (int l = 0; l < depth; ++l) h.level[l] = st->d_levels + st->level_off[l];
  // decimated, 4x byte-shifted copies of the lowest-resolution level (dense pass)
  const int top = depth - 1;
  {
    const int s = 1 << top;
    h.dec_id[top] = (h.wx[top] + s - 1) / s;
    h.dec_jd[top] = (h.wy[top] + s - 1) / s;
    h.dec_ids[top] = (h.dec_id[top] + 3 + 3) / 4 * 4;  // >= 3 zero bytes after every row
    const long long bytes = static_cast<long long>(s) * s * h.dec_jd[top] * h.dec_ids[top];
    CSM_REQUIRE(bytes < (1LL << 29), "decimated level too large");
    h.dec_lpad[top] = static_cast<int>((bytes + 32 + 15) / 16 * 16);
    CSM_CUDA(cudaMalloc(&st->d_dec, 4 * static_cast<size_t>(h.dec_lpad[top])));
    h.dec4[top] = st->d_dec;
  }
  // child-window words for every parent level (branch steps and dives)
  std::vector<size_t> win_off(depth, 0);
  size_t win_total = 0;
  for (int l = 1; l < depth; ++l) {
    const int S = 1 << l;
    h.win_ids[l] = (h.wx[l - 1] + S - 1) / S + 1;
    h.win_jd[l] = (h.wy[l - 1] + S - 1) / S + 1;
    const long long words = static_cast<long long>(S) * S * WinPhaseWords(h.win_jd[l], h.win_ids[l]);
    CSM_REQUIRE(words < (1LL << 31), "window level too large");
    win_off[l] = win_total;
    win_total += (static_cast<size_t>(words) + 63) / 64 * 64;
  }
  if (win_total) CSM_CUDA(cudaMalloc(&st->d_win, win_total * sizeof(unsigned)));
  for (int l = 1; l < depth; ++l) h.win[l] = st->d_win + win_off[l];
  for (int l = 1; l < depth; ++l) st->win_off[l] = win_off[l];
  CSM_TRY(BuildStack2D(st.get(), cells));
  CSM_CUDA(cudaMalloc(&st->d, sizeof(StackDev)));
  CSM_CUDA(cudaMemcpyAsync(st->d, &h, sizeof(StackDev), cudaMemcpyHostToDevice, ctx->stream));
  CSM_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = st.release();
  return CSM_OK;
}

csm_status csm_stack2d_destroy(csm_stack2d* stack) {
  if (!stack) return CSM_OK;
  std::lock_guard<std::mutex> lock(stack->ctx->mu);
  cudaSetDevice(stack->ctx->device);
  cudaStreamSynchronize(stack->ctx->stream);
  delete stack;  // the destructor frees the device buffers
  return CSM_OK;
}

// Incremental refresh: the submap's grid received new range data (same limits, same cost
// bounds); rebuilds every layout in place.  No match may be in flight on this stack
// (ConstraintBuilder only matches against finished submaps; a trimmed or updated submap
// goes through DeleteScanMatcher, constraints/constraint_builder_2d.cc:307-316).
csm_status csm_stack2d_update(csm_stack2d* stack, const uint16_t* cells) {
  CSM_REQUIRE(stack != nullptr && cells != nullptr, "null pointer");
  std::lock_guard<std::mutex> lock(stack->ctx->mu);
  CSM_CUDA(cudaSetDevice(stack->ctx->device));
  CSM_TRY(BuildStack2D(stack, cells));
  CSM_CUDA(cudaStreamSynchronize(stack->ctx->stream));
  return CSM_OK;
}

csm_status csm_stack2d_read_level(const csm_stack2d* stack, int32_t level, uint8_t* out,
                                  int32_t* wide_num_x, int32_t* wide_num_y) {
  CSM_REQUIRE(stack != nullptr, "null stack");
  CSM_REQUIRE(level >= 0 && level < stack->h.depth, "level out of range");
  if (wide_num_x) *wide_num_x = stack->h.wx[level];
  if (wide_num_y) *wide_num_y = stack->h.wy[level];
  if (out) {
    std::lock_guard<std::mutex> lock(stack->ctx->mu);
    CSM_CUDA(cudaSetDevice(stack->ctx->device));
    CSM_CUDA(cudaMemcpy(out, stack->h.level[level],
                        static_cast<size_t>(stack->h.wx[level]) * stack->h.wy[level],
                        cudaMemcpyDeviceToHost));
  }
  return CSM_OK;
}

csm_status csm_cloud_create(const float* xyz, int32_t n, int32_t device, csm_cloud** out) {
  CSM_REQUIRE(out != nullptr && xyz != nullptr, "null pointer");
  CSM_REQUIRE(n >= 1, "empty point cloud");
  Ctx* ctx;
  CSM_TRY(GetCtx(device, &ctx));
  std::lock_guard<std::mutex> lock(ctx->mu);
  CSM_CUDA(cudaSetDevice(device));
  std::unique_ptr<csm_cloud> c(new csm_cloud);
  c->ctx = ctx;
  c->n = n;
  c->h_xyz.assign(xyz, xyz + 3 * static_cast<size_t>(n));
  float m = 0.f;
  for (int i = 0; i < n; ++i) {
    const float x = xyz[3 * i], y = xyz[3 * i + 1];
    const float range = std::sqrt(x * x + y * y);  // head<2>().norm()
    m = std::max(range, m);
  }
  c->max_norm = m;
  const size_t need = sizeof(float) * 3 * static_cast<size_t>(n);
  for (size_t i = 0; i < ctx->cloud_pool.size(); ++i) {
    if (ctx->cloud_pool[i].second >= need && ctx->cloud_pool[i].second <= 2 * need + 4096) {
      c->d_xyz = static_cast<float*>(ctx->cloud_pool[i].first);
      c->d_bytes = ctx->cloud_pool[i].second;
      ctx->cloud_pool_bytes -= c->d_bytes;
      ctx->cloud_pool[i] = ctx->cloud_pool.back();
      ctx->cloud_pool.pop_back();
      break;
    }
  }
  if (!c->d_xyz) {
    c->d_bytes = (need + 4095) / 4096 * 4096;
    CSM_CUDA(cudaMalloc(&c->d_xyz, c->d_bytes));
  }
  CSM_CUDA(cudaMemcpyAsync(c->d_xyz, xyz, need, cudaMemcpyHostToDevice, ctx->stream));
  CSM_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = c.release();
  return CSM_OK;
}

csm_status csm_cloud_destroy(csm_cloud* cloud) {
  if (!cloud) return CSM_OK;
  std::lock_guard<std::mutex> lock(cloud->ctx->mu);
  cudaSetDevice(cloud->ctx->device);
  Ctx* ctx = cloud->ctx;
  // No match is in flight on this cloud (caller contract), so the buffer can be
  // handed to the next csm_cloud_create without a device-wide cudaFree.
  if (cloud->d_xyz && ctx->cloud_pool.size() < 4096 &&
      ctx->cloud_pool_bytes + cloud->d_bytes <= (256u << 20)) {
    ctx->cloud_pool.emplace_back(cloud->d_xyz, cloud->d_bytes);
    ctx->cloud_pool_bytes += cloud->d_bytes;
    cloud->d_xyz = nullptr;  // now owned by the pool
  } else {
    cudaStreamSynchronize(ctx->stream);
  }
  delete cloud;  // frees d_xyz unless it was pooled
  return CSM_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Batched matcher
// ---------------------------------------------------------------------------
namespace {

// SearchParameters(linear, angular, cloud, resolution)   (corr...2d.cc:27-55)
struct HostSearch {
  int num_angular;
  int num_scans;
  int lin;
  double step;
};
HostSearch MakeSearch(double linear_window, double angular_window, float cloud_max_norm,
                      double resolution) {
  float max_scan_range = 3.f * resolution;
  max_scan_range = std::max(cloud_max_norm, max_scan_range);
  const double kSafetyMargin = 1. - 1e-3;
  HostSearch s;
  s.step = kSafetyMargin *
           std::acos(1. - (resolution * resolution) / (2. * (max_scan_range * max_scan_range)));
  s.num_angular = static_cast<int>(std::ceil(angular_window / s.step));
  s.num_scans = 2 * s.num_angular + 1;
  s.lin = static_cast<int>(std::ceil(linear_window / resolution));
  return s;
}

struct TrigKey {
  const csm_cloud* cloud;
  double resolution, angular;
  bool operator<(const TrigKey& o) const {
    if (cloud != o.cloud) return cloud < o.cloud;
    if (resolution != o.resolution) return resolution < o.resolution;
    return angular < o.angular;
  }
};

struct BatchPlan {
  std::vector<JobDev> jobs;
  std::vector<HostSearch> search;
  std::vector<double> init_x, init_y, init_theta;
  long long total_scans = 0, total_points = 0, total_slots = 0;
};

struct TieLeaf { int scan, xo, yo; };

}  // namespace

// Runs a batch of independent matches.  `with_search` == false stops after
// discretisation (test hook csm_discretize2d).
static csm_status RunBatch2D(Ctx* ctx, const csm_stack2d* const* stacks, int num_stacks,
                             const csm_cloud* const* clouds, int num_clouds,
                             const csm_job2d* jobs, int num_jobs, double linear_window,
                             double angular_window, csm_result2d* results, csm_stats* total,
                             bool discretize_only, int32_t* out_dscan, int32_t* out_bounds) {
  CSM_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  static const bool timing = getenv("CSM_TIMING") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(
                        std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_phase = now();
  auto phase = [&](const char* name) {
    if (!timing) return;
    cudaStreamSynchronize(s);
    const double t = now();
    fprintf(stderr, "[csm timing] %-14s %8.3f ms\n", name, t - t_phase);
    t_phase = t;
  };
  BatchPlan plan;
  plan.jobs.resize(num_jobs);
  plan.search.resize(num_jobs);
  plan.init_x.resize(num_jobs);
  plan.init_y.resize(num_jobs);
  plan.init_theta.resize(num_jobs);

  // ---- host: SearchParameters + rotation tables (libm, like the reference) ----
  std::map<TrigKey, long long> trig_index;  // -> offset (in float2) into trig table
  std::vector<float> trig;                  // (cos, sin) pairs
  std::vector<long long> trig_off;
  for (int j = 0; j < num_jobs; ++j) {
    const csm_job2d& jb = jobs[j];
    CSM_REQUIRE(jb.stack_index >= 0 && jb.stack_index < num_stacks, "stack index");
    CSM_REQUIRE(jb.cloud_index >= 0 && jb.cloud_index < num_clouds, "cloud index");
    const csm_stack2d* st = stacks[jb.stack_index];
    const csm_cloud* cl = clouds[jb.cloud_index];
    CSM_REQUIRE(st && cl && st->ctx->device == ctx->device && cl->ctx->device == ctx->device,
                "handles must share one device");
    const StackDev& h = st->h;
    double lin = linear_window, ang = angular_window;
    double ix = jb.initial_pose[0], iy = jb.initial_pose[1], ith = jb.initial_pose[2];
    if (jb.full_submap) {  // fast...2d.cc:210-225
      lin = 1e6 * h.resolution;
      ang = M_PI;
      ix = h.max_x - 0.5 * h.resolution * h.ny;
      iy = h.max_y - 0.5 * h.resolution * h.nx;
      ith = 0.;
    }
    const HostSearch sp = MakeSearch(lin, ang, cl->max_norm, h.resolution);
    CSM_REQUIRE(sp.num_scans > 0 && sp.num_scans < (1 << 22), "angular window / step");
    plan.search[j] = sp;
    plan.init_x[j] = ix;
    plan.init_y[j] = iy;
    plan.init_theta[j] = ith;
    const TrigKey key{cl, h.resolution, ang};
    auto it = trig_index.find(key);
    if (it == trig_index.end()) {
      const long long off = static_cast<long long>(trig.size() / 2);
      // GenerateRotatedScans: delta_theta accumulates in double, is cast to float
      // for AngleAxisf, Quaternionf takes cos/sin of the float half angle.
      double delta_theta = -sp.num_angular * sp.step;
      for (int k = 0; k < sp.num_scans; ++k, delta_theta += sp.step) {
        const float ha = 0.5f * static_cast<float>(delta_theta);
        trig.push_back(std::cos(ha));
        trig.push_back(std::sin(ha));
      }
      it = trig_index.emplace(key, off).first;
    }
    JobDev& d = plan.jobs[j];
    d.stack = st->d;
    d.xyz = cl->d_xyz;
    d.trig = nullptr;  // set once the table is uploaded
    trig_off.push_back(it->second);
    d.n = cl->n;
    d.num_scans = sp.num_scans;
    d.scan_base = static_cast<int>(plan.total_scans);
    d.lin = sp.lin;
    {
      const float ha = 0.5f * static_cast<float>(ith);
      const float sn = std::sin(ha);
      d.q0w = std::cos(ha);
      d.q0x = sn * 0.f;
      d.q0y = sn * 0.f;
      d.q0z = sn * 1.f;
    }
    d.tx = static_cast<float>(ix);
    d.ty = static_cast<float>(iy);
    d.min_score = jb.min_score;
    // upper bound of lowest-resolution candidates per axis after ShrinkToFit
    // (corr...2d.cc:73-91): with c = cell of the sensor origin and e = scan radius in
    // cells, every point index lies in [c - e, c + e], so the window is at most
    //   min(lin, max(0, cells - 1 - (c - e))) + min(lin, max(0, c + e)).
    // (k_discretize re-checks the bound on the device and reports a violation.)
    const int step = 1 << (h.depth - 1);
    const long long e = static_cast<long long>(std::ceil(cl->max_norm / h.resolution)) + 4;
    auto clampll = [](double v) {
      return static_cast<long long>(std::max(-4e9, std::min(4e9, std::floor(v))));
    };
    const long long c_x = clampll((h.max_y - iy) / h.resolution - 0.5);  // index x <- world y
    const long long c_y = clampll((h.max_x - ix) / h.resolution - 0.5);
    auto span = [&](long long cells, long long c) {
      const long long lin = sp.lin;
      return std::min(lin, std::max<long long>(0, cells - 1 - (c - e))) +
             std::min(lin, std::max<long long>(0, c + e));
    };
    const long long span_x = span(h.nx, c_x), span_y = span(h.ny, c_y);
    const long long cx = (span_x + step) / step, cy = (span_y + step) / step;
    d.cap_y = static_cast<int>(cy);
    d.cap = static_cast<int>(cx * cy);
    d.dscan_off = plan.total_points;
    d.top_off = plan.total_slots;
    plan.total_scans += sp.num_scans;
    plan.total_points += static_cast<long long>(sp.num_scans) * cl->n;
    plan.total_slots += static_cast<long long>(sp.num_scans) * d.cap;
    CSM_REQUIRE(plan.total_scans < (1LL << 30), "too many scans in one batch");
  }
  // scan -> job on the host (few lookups: optimal leaves only); the per-scan tables
  // themselves are written on the device (k_scan_tables)
  std::vector<int> scan_bases(num_jobs);
  for (int j = 0; j < num_jobs; ++j) scan_bases[j] = plan.jobs[j].scan_base;
  auto job_of_scan = [&](int scan) {
    return static_cast<int>(std::upper_bound(scan_bases.begin(), scan_bases.end(), scan) -
                            scan_bases.begin()) - 1;
  };

  phase("host plan");
  // ---- device buffers ----
  DevBuf& d_trig = ctx->D("trig");
  DevBuf& d_jobs = ctx->D("jobs");
  DevBuf& d_scan_job = ctx->D("scan_job");
  DevBuf& d_slot_base = ctx->D("slot_base");
  DevBuf& d_info = ctx->D("info");
  DevBuf& d_dscan = ctx->D("dscan");
  DevBuf& d_top = ctx->D("top_sum");
  DevBuf& d_lb = ctx->D("lb");
  DevBuf& d_ctr = ctx->D("counters");
  CSM_TRY(d_trig.Reserve(trig.size() * sizeof(float)));
  CSM_TRY(d_jobs.Reserve(sizeof(JobDev) * num_jobs));
  CSM_TRY(d_scan_job.Reserve(sizeof(int) * plan.total_scans));
  CSM_TRY(d_slot_base.Reserve(sizeof(long long) * plan.total_scans));
  CSM_TRY(d_info.Reserve(sizeof(ScanInfo) * plan.total_scans));
  CSM_TRY(d_dscan.Reserve(sizeof(short2) * plan.total_points));
  CSM_TRY(d_lb.Reserve(sizeof(unsigned) * num_jobs));
  CSM_TRY(d_ctr.Reserve(sizeof(unsigned long long) * 8 + sizeof(int) * kCtlInts));
  DevBuf& d_job_best = ctx->D("job_best");
  CSM_TRY(d_job_best.Reserve(sizeof(int) * num_jobs));
  for (int j = 0; j < num_jobs; ++j) plan.jobs[j].trig = d_trig.as<float2>() + trig_off[j];

  // uploads go through pinned staging so that they are truly asynchronous (the
  // previous call on this lane has completed: every call ends with a synchronise)
  PinnedBuf& up = ctx->P("upload");
  const size_t up_trig = trig.size() * sizeof(float);
  const size_t up_jobs_off = (up_trig + 255) / 256 * 256;
  CSM_TRY(up.Reserve(up_jobs_off + sizeof(JobDev) * num_jobs));
  if (up_trig) std::memcpy(up.as<char>(), trig.data(), up_trig);
  std::memcpy(up.as<char>() + up_jobs_off, plan.jobs.data(), sizeof(JobDev) * num_jobs);
  CSM_CUDA(cudaEventRecord(ctx->ev0, s));
  if (up_trig)
    CSM_CUDA(cudaMemcpyAsync(d_trig.p, up.as<char>(), up_trig, cudaMemcpyHostToDevice, s));
  CSM_CUDA(cudaMemcpyAsync(d_jobs.p, up.as<char>() + up_jobs_off, sizeof(JobDev) * num_jobs,
                           cudaMemcpyHostToDevice, s));
  k_scan_tables<<<num_jobs, 256, 0, s>>>(d_jobs.as<JobDev>(), d_scan_job.as<int>(),
                                         d_slot_base.as<long long>(), d_lb.as<unsigned>(),
                                         d_job_best.as<int>());
  CSM_LAUNCH_CHECK();
  CSM_CUDA(cudaMemsetAsync(d_ctr.p, 0, sizeof(unsigned long long) * 8 + sizeof(int) * kCtlInts, s));
  unsigned long long* ctr = d_ctr.as<unsigned long long>();
  int* ictr = reinterpret_cast<int*>(ctr + 8);  // the level-loop control block (kCtl*)

  const int total_scans = static_cast<int>(plan.total_scans);
  ProfBegin(ctx);
  k_discretize<<<total_scans, 128, 0, s>>>(d_jobs.as<JobDev>(), d_scan_job.as<int>(),
                                           d_dscan.as<short2>(), d_info.as<ScanInfo>(), 1, ctr);
  CSM_LAUNCH_CHECK();
  ProfEnd(ctx, "k_discretize", static_cast<double>(plan.total_points));

  std::vector<ScanInfo> h_info;
  if (discretize_only) {
    h_info.resize(total_scans);
    CSM_CUDA(cudaMemcpyAsync(h_info.data(), d_info.p, sizeof(ScanInfo) * total_scans,
                             cudaMemcpyDeviceToHost, s));
    std::vector<short2> h_ds;
    if (out_dscan) {
      h_ds.resize(plan.total_points);
      CSM_CUDA(cudaMemcpyAsync(h_ds.data(), d_dscan.p, sizeof(short2) * plan.total_points,
                               cudaMemcpyDeviceToHost, s));
    }
    CSM_CUDA(cudaStreamSynchronize(s));
    for (size_t i = 0; i < h_ds.size(); ++i) {
      out_dscan[2 * i] = h_ds[i].x;
      out_dscan[2 * i + 1] = h_ds[i].y;
    }
    if (out_bounds)
      for (int i = 0; i < total_scans; ++i) {
        out_bounds[4 * i + 0] = h_info[i].min_x;
        out_bounds[4 * i + 1] = h_info[i].max_x;
        out_bounds[4 * i + 2] = h_info[i].min_y;
        out_bounds[4 * i + 3] = h_info[i].max_y;
      }
    return CSM_OK;
  }

  phase("upload+discr");
  // ---- lowest-resolution pass ----
  CSM_TRY(d_top.Reserve(sizeof(int) * plan.total_slots));
  // Wide lattices (MatchFullSubmap) take the dense decimated-grid kernel; narrow
  // ones (local windows: a few dozen candidates per scan) the gather kernel.
  static const char* force = getenv("CSM_TOP_KERNEL");  // "small" | "gather" | "tile" | "dense" (debug)
  int max_cap = 0, max_cap_x = 0, max_cap_y = 0;
  for (const JobDev& d : plan.jobs) {
    max_cap = std::max(max_cap, d.cap);
    max_cap_y = std::max(max_cap_y, d.cap_y);
    max_cap_x = std::max(max_cap_x, d.cap / std::max(1, d.cap_y));
  }
  bool use_gather_top = max_cap < 128;
  if (force && !strcmp(force, "gather")) use_gather_top = true;
  if (force && !strcmp(force, "dense")) use_gather_top = false;
  // Tile form of the dense pass: needs the whole tile in <= 4 words per lane and the
  // lattice of one scan per warp in shared memory.
  int tile_words = 0, lat_ints = 0;
  for (int j = 0; j < num_jobs; ++j) {
    const StackDev& sh = stacks[jobs[j].stack_index]->h;
    const int t = sh.depth - 1;
    tile_words = std::max(tile_words, sh.dec_jd[t] * (sh.dec_ids[t] / 4) + 1);
  }
  lat_ints = (max_cap_x + 3) / 4 * 4 * max_cap_y;
  bool use_tile_top = !use_gather_top && tile_words <= 128 && lat_ints * 4 * kTileWarps <= 96 * 1024;
  const int small_lanes = (max_cap_x + 3) / 4 * max_cap_y;  // quads per scan, a-priori bound
  bool use_small_top = small_lanes <= 32;
  if (force && !strcmp(force, "small"))
    CSM_REQUIRE(use_small_top, "CSM_TOP_KERNEL=small: more than 32 quads per scan");
  if (force && strcmp(force, "small")) use_small_top = false;
  if (use_small_top) use_gather_top = use_tile_top = false;
  if (force && !strcmp(force, "dense")) use_tile_top = false;
  if (force && !strcmp(force, "tile"))
    CSM_REQUIRE(use_tile_top, "CSM_TOP_KERNEL=tile: lattice or tile too large");
  const char* top_name = use_small_top    ? "k_score_top_small"
                         : use_gather_top ? "k_score_top_gather"
                         : use_tile_top   ? "k_score_top_tile" : "k_score_top_dense";
  ProfBegin(ctx);
  if (use_small_top) {
    const int spw = 32 / small_lanes;
    k_score_top_small<<<DivUp(DivUp(total_scans, spw), 4), 128, 4 * spw * 32 * sizeof(int2), s>>>(
        d_jobs.as<JobDev>(), d_info.as<ScanInfo>(), d_dscan.as<short2>(), d_top.as<int>(),
        d_slot_base.as<long long>(), total_scans, small_lanes);
  } else if (use_gather_top) {
    k_score_top_gather<<<DivUp(plan.total_slots * 32, 256), 256, 0, s>>>(
        d_jobs.as<JobDev>(), d_info.as<ScanInfo>(), d_dscan.as<short2>(), d_top.as<int>(),
        d_slot_base.as<long long>(), total_scans, plan.total_slots);
  } else if (use_tile_top) {
    const size_t smem = static_cast<size_t>(lat_ints) * 4 * kTileWarps;  // one lattice per warp
    // persistent warps: exactly the resident CTAs, every warp walks total_scans / warps scans
#define CSM_TILE(K)                                                                          \
    do {                                                                                     \
      if (smem > 48 * 1024)                                                                  \
        CSM_CUDA(cudaFuncSetAttribute(k_score_top_tile<K>,                                   \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                      static_cast<int>(smem)));                              \
      int per_sm = 0;                                                                        \
      CSM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_score_top_tile<K>,   \
                                                             kTileThreads, smem));           \
      const int grid = std::min(DivUp(total_scans, kTileWarps), ctx->sm_count * std::max(1, per_sm)); \
      k_score_top_tile<K><<<grid, kTileThreads, smem, s>>>(                                  \
          d_jobs.as<JobDev>(), d_info.as<ScanInfo>(), d_dscan.as<short2>(), d_top.as<int>(), \
          d_slot_base.as<long long>(), total_scans, lat_ints);                               \
    } while (0)
    if (tile_words <= 32) CSM_TILE(1);
    else if (tile_words <= 64) CSM_TILE(2);
    else if (tile_words <= 96) CSM_TILE(3);
    else CSM_TILE(4);
#undef CSM_TILE
  } else {
    const int grid = std::min(total_scans, ctx->sm_count * 128);
    // one quad per thread and extra passes for larger lattices (the real lattice is
    // only known on the device after ShrinkToFit; more quads per thread would execute
    // predicated-off work)
    k_score_top_dense<1><<<grid, kDenseThreads, 0, s>>>(
        d_jobs.as<JobDev>(), d_info.as<ScanInfo>(), d_dscan.as<short2>(), d_top.as<int>(),
        d_slot_base.as<long long>(), total_scans);
  }
  CSM_LAUNCH_CHECK();
  if (g_profile_on.load()) {
    unsigned long long c3 = 0;
    ProfStop(ctx);
    CSM_CUDA(cudaStreamSynchronize(s));
    CSM_CUDA(cudaMemcpy(&c3, ctr + 3, sizeof(c3), cudaMemcpyDeviceToHost));
    ProfCommit(ctx, top_name, static_cast<double>(c3));
  }

  phase("top pass");
  // ---- greedy dives seed the per-job bound ----
  unsigned long long prof_c0 = 0;
  auto prof_scored = [&]() -> double {  // candidates scored since the last call
    unsigned long long c0 = 0;
    cudaStreamSynchronize(s);
    cudaMemcpy(&c0, ctr, sizeof(c0), cudaMemcpyDeviceToHost);
    const double d = static_cast<double>(c0 - prof_c0);
    prof_c0 = c0;
    return d;
  };
  if (g_profile_on.load()) prof_scored();
  static const float dive_ratio = getenv("CSM_DIVE_RATIO") ? atof(getenv("CSM_DIVE_RATIO")) : 0.97f;
  ProfBegin(ctx);
  k_job_best<<<DivUp(static_cast<long long>(total_scans) * 32, 256), 256, 0, s>>>(
      d_info.as<ScanInfo>(), d_top.as<int>(), d_slot_base.as<long long>(), total_scans,
      d_job_best.as<int>());
  CSM_LAUNCH_CHECK();
  k_dive<<<DivUp(static_cast<long long>(total_scans) * 32, 256), 256, 0, s>>>(
      d_jobs.as<JobDev>(), d_info.as<ScanInfo>(), d_dscan.as<short2>(), d_top.as<int>(),
      d_slot_base.as<long long>(), total_scans, d_job_best.as<int>(), dive_ratio,
      d_lb.as<unsigned>(), ctr);
  CSM_LAUNCH_CHECK();
  if (g_profile_on.load()) {
    ProfStop(ctx);
    ProfCommit(ctx, "k_dive", prof_scored());
  }

  phase("dives");
  // ---- branch and bound: per-level queues, device-driven level loop ----
  // The first sweep hmax -> 1 is a static stream of launches: every level takes a chunk of
  // <= kChunk nodes from the end of its queue (k_level_begin), picks the kernel form on the
  // device and appends the survivors to the next queue; nothing is read back in between.
  // A queue below the top starts the sweep empty and receives <= 4 * kChunk children, so it
  // cannot overflow.  Only if a level held more than one chunk (frontiers > 8 M nodes) does
  // the host continue, deepest non-empty level first, with one read-back per extra chunk.
  int depth_max = 0;
  for (int j = 0; j < num_jobs; ++j)
    depth_max = std::max(depth_max, stacks[jobs[j].stack_index]->h.depth);
  for (int j = 0; j < num_jobs; ++j)
    CSM_REQUIRE(stacks[jobs[j].stack_index]->h.depth == depth_max,
                "internal: sub-batches are grouped by branch_and_bound_depth");
  const int hmax = depth_max - 1;
  const int kChunk = 1 << 23;
  const int kQueueCap = 4 * kChunk;
  const int kLeafCap = 1 << 22;
  const int kLatticeMin = 16384;  // smaller frontiers: the warp-per-parent kernel has a
                                  // 34-iteration critical path, the lattice kernel a 1081-iteration one
  const int kInlineBest = 4096;   // optimal leaves read back with the first (usually only) sync
  const long long top_cap_ll = std::min<long long>(plan.total_slots, 1LL << 30);
  const int top_cap = static_cast<int>(top_cap_ll);
  DevBuf& d_qtop = ctx->D("queue_top");
  DevBuf& d_q = ctx->D("queues");
  DevBuf& d_leaves = ctx->D("leaves");
  DevBuf& d_best = ctx->D("best_leaves");
  CSM_TRY(d_qtop.Reserve(sizeof(Node) * static_cast<size_t>(top_cap)));
  CSM_TRY(d_q.Reserve(sizeof(Node) * static_cast<size_t>(kQueueCap) * std::max(1, hmax)));
  CSM_TRY(d_leaves.Reserve(sizeof(Node) * static_cast<size_t>(kLeafCap)));
  CSM_TRY(d_best.Reserve(sizeof(Node) * static_cast<size_t>(kLeafCap)));
  DevBuf& d_scan_cnt = ctx->D("scan_cnt");
  DevBuf& d_scan_off = ctx->D("scan_off");
  DevBuf& d_items = ctx->D("work_items");
  DevBuf& d_sorted = ctx->D("sorted_nodes");
  CSM_TRY(d_scan_cnt.Reserve(sizeof(int) * 2 * static_cast<size_t>(total_scans)));
  CSM_TRY(d_scan_off.Reserve(sizeof(int) * (static_cast<size_t>(total_scans) +
                                            2 * (static_cast<size_t>(total_scans) / 1024 + 2))));
  CSM_TRY(d_items.Reserve(sizeof(WorkItem) * (static_cast<size_t>(kChunk) / 32 + total_scans + 2)));
  CSM_TRY(d_sorted.Reserve(sizeof(Node) * static_cast<size_t>(kChunk)));
  static const bool use_lattice = getenv("CSM_NO_LATTICE") == nullptr;
  auto queue_ptr = [&](int h) -> Node* {
    return h == hmax ? d_qtop.as<Node>() : d_q.as<Node>() + static_cast<size_t>(kQueueCap) * h;
  };
  auto queue_cap = [&](int h) { return h == hmax ? top_cap : kQueueCap; };
  int* overflow = ictr + kCtlOverflow;
  int* leaf_count = ictr + kCtlLeaf;
  int* best_count = ictr + kCtlBest;

  // read-back area (pinned): control block | counters | per-job bound | first optimal leaves
  PinnedBuf& pin = ctx->P("readback");
  const size_t rb_ctr = sizeof(int) * kCtlInts;
  const size_t rb_lb = rb_ctr + sizeof(unsigned long long) * 8;
  const size_t rb_best = (rb_lb + sizeof(unsigned) * num_jobs + 15) / 16 * 16;
  CSM_TRY(pin.Reserve(rb_best + sizeof(Node) * kInlineBest));
  int* hp = pin.as<int>();
  const unsigned long long* hctr = reinterpret_cast<const unsigned long long*>(pin.as<char>() + rb_ctr);
  const unsigned* lbh = reinterpret_cast<const unsigned*>(pin.as<char>() + rb_lb);
  const Node* best_inline = reinterpret_cast<const Node*>(pin.as<char>() + rb_best);
  int host_syncs = 0;

  k_filter_top<<<std::min(total_scans, ctx->sm_count * 16), 256, 0, s>>>(
      d_jobs.as<JobDev>(), d_info.as<ScanInfo>(), d_top.as<int>(), d_slot_base.as<long long>(),
      total_scans, d_lb.as<unsigned>(), queue_ptr(hmax), ictr + hmax, queue_cap(hmax), overflow);
  CSM_LAUNCH_CHECK();

  if (hmax == 0) {
    k_top_as_leaves<<<ctx->sm_count * 4, 256, 0, s>>>(d_info.as<ScanInfo>(), queue_ptr(0), ictr,
                                                      d_lb.as<unsigned>(), d_leaves.as<Node>(),
                                                      kLeafCap);
    CSM_LAUNCH_CHECK();
  }

  // One level step: chunk bookkeeping, grouping by scan, both kernel forms predicated on
  // the device-side mode.
  const int nb = DivUp(total_scans, 1024);
  const int sort_grid = ctx->sm_count * 8;
  const int max_items = kChunk / 32 + std::min(kChunk, total_scans) + 1;
  static const int lat_unroll = getenv("CSM_LAT_UNROLL") ? atoi(getenv("CSM_LAT_UNROLL")) : 8;
  auto level_step = [&](int h) -> csm_status {
    int* scan_cnt = d_scan_cnt.as<int>();
    int* cursor = scan_cnt + total_scans;
    int* scan_off = d_scan_off.as<int>();
    int* part_a = scan_off + total_scans;   // block sums (DivUp(total_scans, 1024) each)
    int* part_b = part_a + nb;
    Node* next = h - 1 >= 1 ? queue_ptr(h - 1) : nullptr;
    int* next_count = ictr + (h - 1 >= 1 ? h - 1 : 31);
    k_level_begin<<<1, 32, 0, s>>>(ictr, h, kChunk, use_lattice ? kLatticeMin : INT_MAX);
    CSM_LAUNCH_CHECK();
    CSM_CUDA(cudaMemsetAsync(d_scan_cnt.p, 0, sizeof(int) * 2 * total_scans, s));
    ProfBegin(ctx);
    k_q_count<<<sort_grid, 256, 0, s>>>(queue_ptr(h), ictr, scan_cnt);
    CSM_LAUNCH_CHECK();
    k_q_block_sums<<<nb, 1024, 0, s>>>(ictr, scan_cnt, total_scans, part_a, part_b);
    CSM_LAUNCH_CHECK();
    k_q_scan_parts<<<1, 1024, 0, s>>>(ictr, part_a, part_b, nb, ictr + kCtlItems);
    CSM_LAUNCH_CHECK();
    k_q_finish<<<nb, 1024, 0, s>>>(ictr, scan_cnt, total_scans, part_a, part_b, scan_off,
                                   d_items.as<WorkItem>());
    CSM_LAUNCH_CHECK();
    k_q_scatter<<<sort_grid, 256, 0, s>>>(queue_ptr(h), ictr, scan_off, cursor,
                                          d_sorted.as<Node>());
    CSM_LAUNCH_CHECK();
    if (g_profile_on.load()) {
      ProfStop(ctx);
      ProfCommit(ctx, "k_q_sort", 0.);
    }
    ProfBegin(ctx);
#define CSM_LATTICE(U)                                                                          \
    k_expand_lattice<U><<<DivUp(max_items, kLatThreads / 32), kLatThreads, 0, s>>>(             \
        d_jobs.as<JobDev>(), d_info.as<ScanInfo>(), d_dscan.as<short2>(), d_sorted.as<Node>(),  \
        d_items.as<WorkItem>(), ictr + kCtlItems, h, d_lb.as<unsigned>(), next, next_count,     \
        kQueueCap, d_leaves.as<Node>(), leaf_count, kLeafCap, overflow, ctr)
    if (lat_unroll <= 4) CSM_LATTICE(4);
    else if (lat_unroll <= 8) CSM_LATTICE(8);
    else CSM_LATTICE(16);
#undef CSM_LATTICE
    CSM_LAUNCH_CHECK();
    if (g_profile_on.load()) {
      ProfStop(ctx);
      const double c = prof_scored();
      if (c > 0.) ProfCommit(ctx, "k_expand_lattice", c);
    }
    ProfBegin(ctx);
    k_expand<<<kLatticeMin * 32 / 256, 256, 0, s>>>(
        d_jobs.as<JobDev>(), d_info.as<ScanInfo>(), d_dscan.as<short2>(), queue_ptr(h), ictr, h,
        d_lb.as<unsigned>(), next, next_count, kQueueCap, d_leaves.as<Node>(), leaf_count,
        kLeafCap, overflow, ctr);
    CSM_LAUNCH_CHECK();
    if (g_profile_on.load()) {
      ProfStop(ctx);
      const double c = prof_scored();
      if (c > 0.) ProfCommit(ctx, "k_expand", c);
    }
    return CSM_OK;
  };
  // compaction of the optimal leaves + read-back of everything the host needs
  auto collect = [&]() -> csm_status {
    CSM_CUDA(cudaMemsetAsync(best_count, 0, sizeof(int), s));
    k_compact_leaves<<<ctx->sm_count * 4, 256, 0, s>>>(
        d_info.as<ScanInfo>(), d_leaves.as<Node>(), leaf_count, kLeafCap, d_lb.as<unsigned>(),
        d_best.as<Node>(), best_count, kLeafCap, overflow);
    CSM_LAUNCH_CHECK();
    CSM_CUDA(cudaEventRecord(ctx->ev1, s));
    CSM_CUDA(cudaMemcpyAsync(pin.as<char>(), ictr, rb_ctr, cudaMemcpyDeviceToHost, s));
    CSM_CUDA(cudaMemcpyAsync(pin.as<char>() + rb_ctr, ctr, sizeof(unsigned long long) * 8,
                             cudaMemcpyDeviceToHost, s));
    CSM_CUDA(cudaMemcpyAsync(pin.as<char>() + rb_lb, d_lb.p, sizeof(unsigned) * num_jobs,
                             cudaMemcpyDeviceToHost, s));
    CSM_CUDA(cudaMemcpyAsync(pin.as<char>() + rb_best, d_best.p, sizeof(Node) * kInlineBest,
                             cudaMemcpyDeviceToHost, s));
    CSM_CUDA(cudaStreamSynchronize(s));
    ++host_syncs;
    return CSM_OK;
  };

  for (int h = hmax; h >= 1; --h) CSM_TRY(level_step(h));
  CSM_TRY(collect());
  // frontiers larger than one chunk (rare): continue deepest non-empty level first
  for (;;) {
    if (hp[kCtlOverflow]) {
      SetError("branch-and-bound queue overflow (more than 2^22 tied optimal leaves, or more "
               "than 2^30 lowest-resolution candidates in one batch)");
      return CSM_E_CAPACITY;
    }
    int h = -1;
    for (int l = 1; l <= hmax; ++l)
      if (hp[l] > 0) { h = l; break; }
    if (h < 0) break;
    if (hp[kCtlLeaf] > kLeafCap / 2) {
      // drop the leaves that are already below their job's bound (d_best holds them
      // after collect()); more than kLeafCap / 2 exactly tied optima are not supported
      if (hp[kCtlBest] > kLeafCap / 2) { SetError("too many tied leaves"); return CSM_E_CAPACITY; }
      CSM_CUDA(cudaMemcpyAsync(d_leaves.p, d_best.p, sizeof(Node) * hp[kCtlBest],
                               cudaMemcpyDeviceToDevice, s));
      CSM_CUDA(cudaMemcpyAsync(leaf_count, best_count, sizeof(int), cudaMemcpyDeviceToDevice, s));
    }
    CSM_TRY(level_step(h));
    for (int l = h - 1; l >= 1; --l) CSM_TRY(level_step(l));  // its children fit one chunk each
    CSM_TRY(collect());
  }

  phase("branch&bound");
  const int n_best = hp[kCtlBest];
  if (hctr[7]) {
    SetError("a scan point lies more than 30000 cells from the grid origin (int16 cell indices)");
    return CSM_E_CAPACITY;
  }
  if (hctr[6]) {
    SetError("internal: a scan's lowest-resolution lattice exceeds its reserved slots");
    return CSM_E_CAPACITY;
  }
  std::vector<Node> best(best_inline, best_inline + std::min(n_best, kInlineBest));
  if (n_best > kInlineBest) {  // many exactly tied optima
    best.resize(n_best);
    CSM_CUDA(cudaMemcpyAsync(best.data(), d_best.p, sizeof(Node) * n_best, cudaMemcpyDeviceToHost, s));
    CSM_CUDA(cudaStreamSynchronize(s));
    ++host_syncs;
  }

  // group optimal leaves by job
  std::vector<std::vector<TieLeaf>> per_job(num_jobs);
  for (const Node& nd : best) per_job[job_of_scan(nd.scan)].push_back(TieLeaf{nd.scan, nd.xo, nd.yo});

  // ---- tie resolution: the reference returns the first optimal leaf in DFS order ----
  int host_resolves = 0;
  long long lowest_total = 0;
  bool need_info = false;
  for (int j = 0; j < num_jobs; ++j)
    if (per_job[j].size() > 1) need_info = true;
  if (need_info) {
    h_info.resize(total_scans);
    CSM_CUDA(cudaMemcpy(h_info.data(), d_info.p, sizeof(ScanInfo) * total_scans,
                        cudaMemcpyDeviceToHost));
  }
  lowest_total = static_cast<long long>(hctr[3]);
  for (int j = 0; j < num_jobs; ++j) {
    std::vector<TieLeaf>& ties = per_job[j];
    if (ties.size() <= 1) continue;
    const int T = static_cast<int>(ties.size());
    // ancestor scores at levels 1..hmax
    std::vector<ListCand> lc;
    lc.reserve(static_cast<size_t>(T) * hmax);
    for (const TieLeaf& t : ties) {
      const ScanInfo& si = h_info[t.scan];
      for (int l = 1; l <= hmax; ++l) {
        const int ax = si.min_x + (((t.xo - si.min_x) >> l) << l);
        const int ay = si.min_y + (((t.yo - si.min_y) >> l) << l);
        lc.push_back(ListCand{t.scan, ax, ay, l});
      }
    }
    std::vector<float> anc(lc.size());
    if (!lc.empty()) {
      DevBuf& d_lc = ctx->D("tie_cands");
      DevBuf& d_ls = ctx->D("tie_scores");
      CSM_TRY(d_lc.Reserve(sizeof(ListCand) * lc.size()));
      CSM_TRY(d_ls.Reserve(sizeof(float) * lc.size()));
      CSM_CUDA(cudaMemcpyAsync(d_lc.p, lc.data(), sizeof(ListCand) * lc.size(),
                               cudaMemcpyHostToDevice, s));
      k_score_list<<<DivUp(static_cast<long long>(lc.size()) * 32, 256), 256, 0, s>>>(
          d_jobs.as<JobDev>(), d_info.as<ScanInfo>(), d_dscan.as<short2>(), d_lc.as<ListCand>(),
          static_cast<int>(lc.size()), nullptr, d_ls.as<float>());
      CSM_LAUNCH_CHECK();
      CSM_CUDA(cudaMemcpyAsync(anc.data(), d_ls.p, sizeof(float) * lc.size(),
                               cudaMemcpyDeviceToHost, s));
      CSM_CUDA(cudaStreamSynchronize(s));
    }
    // lazily computed rank of every lowest-resolution candidate of this job in
    // the reference's std::sort order (fast...2d.cc:331-332)
    std::vector<int> top_rank;
    std::vector<long long> scan_first;  // first generation index of each scan
    auto ensure_top_rank = [&]() -> csm_status {
      if (!top_rank.empty()) return CSM_OK;
      ++host_resolves;
      const JobDev& jd = plan.jobs[j];
      std::vector<int> sums(static_cast<size_t>(jd.num_scans) * jd.cap);
      CSM_CUDA(cudaMemcpy(sums.data(), d_top.as<int>() + jd.top_off, sizeof(int) * sums.size(),
                          cudaMemcpyDeviceToHost));
      const StackDev& sh = stacks[jobs[j].stack_index]->h;
      struct Item { float score; int gen; };
      std::vector<Item> items;
      scan_first.assign(jd.num_scans, 0);
      for (int k = 0; k < jd.num_scans; ++k) {
        const ScanInfo& si = h_info[jd.scan_base + k];
        scan_first[k] = static_cast<long long>(items.size());
        for (int q = 0; q < si.nxc * si.nyc; ++q) {
          const float mean = static_cast<float>(sums[static_cast<size_t>(k) * jd.cap + q]) /
                             static_cast<float>(jd.n);
          items.push_back(Item{sh.min_score + mean * sh.k255, static_cast<int>(items.size())});
        }
      }
      std::sort(items.begin(), items.end(),
                [](const Item& a, const Item& b) { return a.score > b.score; });
      top_rank.resize(items.size());
      for (size_t r = 0; r < items.size(); ++r) top_rank[items[r].gen] = static_cast<int>(r);
      return CSM_OK;
    };
    // returns true if leaf a precedes leaf b in the reference's DFS order
    csm_status err = CSM_OK;
    auto before = [&](int a, int b) -> bool {
      const TieLeaf& A = ties[a];
      const TieLeaf& B = ties[b];
      const ScanInfo& sa = h_info[A.scan];
      const ScanInfo& sb = h_info[B.scan];
      for (int l = hmax; l >= 0; --l) {
        const int ax = (A.xo - sa.min_x) >> l, ay = (A.yo - sa.min_y) >> l;
        const int bx = (B.xo - sb.min_x) >> l, by = (B.yo - sb.min_y) >> l;
        if (A.scan == B.scan && ax == bx && ay == by) continue;  // same ancestor
        const float fa = l == 0 ? 0.f : anc[static_cast<size_t>(a) * hmax + (l - 1)];
        const float fb = l == 0 ? 0.f : anc[static_cast<size_t>(b) * hmax + (l - 1)];
        if (l > 0 && fa != fb) return fa > fb;
        if (l == hmax) {
          // equal lowest-resolution scores: replay the reference's std::sort
          if (ensure_top_rank() != CSM_OK) { err = CSM_E_CUDA; return false; }
          const JobDev& jd = plan.jobs[j];
          const long long ga = scan_first[A.scan - jd.scan_base] + static_cast<long long>(ax) * sa.nyc + ay;
          const long long gb = scan_first[B.scan - jd.scan_base] + static_cast<long long>(bx) * sb.nyc + by;
          return top_rank[ga] < top_rank[gb];
        }
        // siblings: stable insertion sort keeps generation order (x outer, y inner)
        if ((ax & 1) != (bx & 1)) return (ax & 1) < (bx & 1);
        return (ay & 1) < (by & 1);
      }
      return false;
    };
    int w = 0;
    for (int t = 1; t < T; ++t)
      if (before(t, w)) w = t;
    if (err != CSM_OK) return err;
    std::swap(ties[0], ties[w]);
  }

  phase("collect+ties");
  // ---- results ----
  for (int j = 0; j < num_jobs; ++j) {
    csm_result2d& r = results[j];
    const float best_score = HostOrderedToFloat(lbh[j]);
    r.found = 0;
    r.leaves_tied = static_cast<int32_t>(per_job[j].size());
    if (!per_job[j].empty() && best_score > jobs[j].min_score) {
      const TieLeaf& t = per_job[j][0];
      const JobDev& jd = plan.jobs[j];
      const HostSearch& sp = plan.search[j];
      const double res = stacks[jobs[j].stack_index]->h.resolution;
      const int scan_index = t.scan - jd.scan_base;
      // Candidate2D (corr...2d.h:77-86): x = -y_off * res, y = -x_off * res
      const double cx = -t.yo * res, cy = -t.xo * res;
      const double orientation = (scan_index - sp.num_angular) * sp.step;
      r.found = 1;
      r.score = best_score;
      r.pose_estimate[0] = plan.init_x[j] + cx;
      r.pose_estimate[1] = plan.init_y[j] + cy;
      r.pose_estimate[2] = plan.init_theta[j] + orientation;
      r.best_scan_index = scan_index;
      r.best_x_offset = t.xo;
      r.best_y_offset = t.yo;
    }
  }
  if (total) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    total->candidates_scored += static_cast<int64_t>(hctr[0]);
    total->nodes_expanded += static_cast<int64_t>(hctr[1]);
    total->lowest_resolution_candidates += lowest_total;
    total->host_tie_resolves += host_resolves;
    total->host_syncs += host_syncs;
    total->device_ms += ms;
    if (num_jobs == 1) {
      total->num_scans = plan.search[0].num_scans;
      total->leaves_tied = results[0].leaves_tied;
      total->best_scan_index = results[0].best_scan_index;
      total->best_x_offset = results[0].best_x_offset;
      total->best_y_offset = results[0].best_y_offset;
    }
  }
  return CSM_OK;
}

extern "C" {

csm_status csm_match2d_batch(const csm_stack2d* const* stacks, int32_t num_stacks,
                             const csm_cloud* const* clouds, int32_t num_clouds,
                             const csm_job2d* jobs, int32_t num_jobs, double linear_window,
                             double angular_window, csm_result2d* results, csm_stats* total) {
  CSM_REQUIRE(stacks && clouds && jobs && results, "null pointer");
  CSM_REQUIRE(num_jobs >= 1 && num_stacks >= 1 && num_clouds >= 1, "empty batch");
  CSM_REQUIRE(stacks[0] != nullptr, "null stack");
  bool one_depth = true;
  for (int j = 0; j < num_jobs; ++j) {
    const csm_job2d& jb = jobs[j];
    CSM_REQUIRE(jb.stack_index >= 0 && jb.stack_index < num_stacks, "stack index");
    CSM_REQUIRE(jb.cloud_index >= 0 && jb.cloud_index < num_clouds, "cloud index");
    CSM_REQUIRE(stacks[jb.stack_index] && clouds[jb.cloud_index], "null handle");
    one_depth = one_depth &&
                stacks[jb.stack_index]->h.depth == stacks[jobs[0].stack_index]->h.depth;
  }
  LaneGuard guard;
  CSM_TRY(AcquireLane(stacks[0]->ctx->device, &guard));
  Ctx* ctx = guard.lane;
  if (total) std::memset(total, 0, sizeof(*total));
  // Runs jobs[0..n) (one branch_and_bound_depth) in sub-batches whose discrete-scan buffer
  // stays below ~4 GB.
  auto run = [&](const csm_job2d* js, int n, csm_result2d* rs) -> csm_status {
    const long long kMaxPoints = 1LL << 30;
    int j0 = 0;
    while (j0 < n) {
      long long pts = 0;
      int j1 = j0;
      while (j1 < n) {
        const csm_stack2d* st = stacks[js[j1].stack_index];
        const csm_cloud* cl = clouds[js[j1].cloud_index];
        const double ang = js[j1].full_submap ? M_PI : angular_window;
        const double lin = js[j1].full_submap ? 1e6 * st->h.resolution : linear_window;
        const HostSearch sp = MakeSearch(lin, ang, cl->max_norm, st->h.resolution);
        const long long add = static_cast<long long>(sp.num_scans) * cl->n;
        if (j1 > j0 && pts + add > kMaxPoints) break;
        pts += add;
        ++j1;
      }
      CSM_TRY(RunBatch2D(ctx, stacks, num_stacks, clouds, num_clouds, js + j0, j1 - j0,
                         linear_window, angular_window, rs + j0, total, false, nullptr, nullptr));
      j0 = j1;
    }
    return CSM_OK;
  };
  if (one_depth) return run(jobs, num_jobs, results);
  // Stacks of different branch_and_bound_depth (the level loop is per depth): one pass per
  // depth, results scattered back into job order.
  std::map<int, std::vector<int>> by_depth;
  for (int j = 0; j < num_jobs; ++j) by_depth[stacks[jobs[j].stack_index]->h.depth].push_back(j);
  for (const auto& kv : by_depth) {
    const std::vector<int>& idx = kv.second;
    std::vector<csm_job2d> js(idx.size());
    std::vector<csm_result2d> rs(idx.size());
    for (size_t i = 0; i < idx.size(); ++i) js[i] = jobs[idx[i]];
    CSM_TRY(run(js.data(), static_cast<int>(js.size()), rs.data()));
    for (size_t i = 0; i < idx.size(); ++i) results[idx[i]] = rs[i];
  }
  return CSM_OK;
}

csm_status csm_match2d(const csm_stack2d* stack, const float* xyz, int32_t n,
                       const double initial_pose[3], int32_t full_submap, double linear_window,
                       double angular_window, float min_score, int32_t* found, float* score,
                       double pose_estimate[3], csm_stats* stats) {
  CSM_REQUIRE(stack && xyz && found && score && pose_estimate, "null pointer");  // :232-233
  CSM_REQUIRE(full_submap || initial_pose, "null initial pose");
  csm_cloud* cloud = nullptr;
  CSM_TRY(csm_cloud_create(xyz, n, stack->ctx->device, &cloud));
  csm_job2d job;
  std::memset(&job, 0, sizeof(job));
  job.full_submap = full_submap;
  if (initial_pose) std::memcpy(job.initial_pose, initial_pose, sizeof(double) * 3);
  job.min_score = min_score;
  csm_result2d res;
  std::memset(&res, 0, sizeof(res));
  const csm_cloud* cl = cloud;
  const csm_status st = csm_match2d_batch(&stack, 1, &cl, 1, &job, 1, linear_window,
                                          angular_window, &res, stats);
  csm_cloud_destroy(cloud);
  if (st != CSM_OK) return st;
  *found = res.found;
  if (res.found) {
    *score = res.score;
    std::memcpy(pose_estimate, res.pose_estimate, sizeof(double) * 3);
  }
  return CSM_OK;
}

csm_status csm_discretize2d(const csm_stack2d* stack, const float* xyz, int32_t n,
                            const double initial_pose[3], int32_t full_submap,
                            double linear_window, double angular_window, int32_t* num_scans,
                            int32_t* discrete_scans, int32_t* bounds) {
  CSM_REQUIRE(stack && xyz && num_scans, "null pointer");
  csm_cloud* cloud = nullptr;
  CSM_TRY(csm_cloud_create(xyz, n, stack->ctx->device, &cloud));
  const double ang = full_submap ? M_PI : angular_window;
  const double lin = full_submap ? 1e6 * stack->h.resolution : linear_window;
  *num_scans = MakeSearch(lin, ang, cloud->max_norm, stack->h.resolution).num_scans;
  csm_status st = CSM_OK;
  if (discrete_scans || bounds) {
    csm_job2d job;
    std::memset(&job, 0, sizeof(job));
    job.full_submap = full_submap;
    if (initial_pose) std::memcpy(job.initial_pose, initial_pose, sizeof(double) * 3);
    const csm_cloud* cl = cloud;
    LaneGuard guard;
    st = AcquireLane(stack->ctx->device, &guard);
    if (st == CSM_OK)
      st = RunBatch2D(guard.lane, &stack, 1, &cl, 1, &job, 1, linear_window, angular_window,
                      nullptr, nullptr, true, discrete_scans, bounds);
  }
  csm_cloud_destroy(cloud);
  return st;
}

csm_status csm_score_candidates2d(const csm_stack2d* stack, int32_t level,
                                  const int32_t* discrete_scans, int32_t num_scans,
                                  int32_t n, const int32_t* candidates, int32_t num_candidates,
                                  float* scores, int32_t* sums) {
  CSM_REQUIRE(stack && discrete_scans && candidates && scores, "null pointer");
  CSM_REQUIRE(level >= 0 && level < stack->h.depth, "level out of range");
  CSM_REQUIRE(num_scans >= 1 && n >= 1 && num_candidates >= 0, "sizes");
  if (num_candidates == 0) return CSM_OK;
  LaneGuard guard;
  CSM_TRY(AcquireLane(stack->ctx->device, &guard));
  Ctx* ctx = guard.lane;
  CSM_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  // one synthetic job whose discrete scans are supplied by the caller
  JobDev jd;
  std::memset(&jd, 0, sizeof(jd));
  jd.stack = stack->d;
  jd.n = n;
  jd.num_scans = num_scans;
  std::vector<ScanInfo> info(num_scans);
  for (int k = 0; k < num_scans; ++k) {
    std::memset(&info[k], 0, sizeof(ScanInfo));
    info[k].job = 0;
  }
  std::vector<ListCand> lc(num_candidates);
  for (int c = 0; c < num_candidates; ++c) {
    CSM_REQUIRE(candidates[3 * c] >= 0 && candidates[3 * c] < num_scans, "scan_index");
    lc[c] = ListCand{candidates[3 * c], candidates[3 * c + 1], candidates[3 * c + 2], level};
  }
  DevBuf& d_jobs = ctx->D("hook_jobs");
  DevBuf& d_info = ctx->D("hook_info");
  DevBuf& d_dscan = ctx->D("hook_dscan");
  DevBuf& d_lc = ctx->D("hook_cands");
  DevBuf& d_sc = ctx->D("hook_scores");
  DevBuf& d_su = ctx->D("hook_sums");
  const size_t npts = static_cast<size_t>(num_scans) * n;
  CSM_TRY(d_jobs.Reserve(sizeof(JobDev)));
  CSM_TRY(d_info.Reserve(sizeof(ScanInfo) * num_scans));
  CSM_TRY(d_dscan.Reserve(sizeof(short2) * npts));
  std::vector<short2> h_ds(npts);
  for (size_t i = 0; i < npts; ++i)
    h_ds[i] = make_short2(
        static_cast<short>(std::max(-30000, std::min(30000, discrete_scans[2 * i]))),
        static_cast<short>(std::max(-30000, std::min(30000, discrete_scans[2 * i + 1]))));
  CSM_TRY(d_lc.Reserve(sizeof(ListCand) * num_candidates));
  CSM_TRY(d_sc.Reserve(sizeof(float) * num_candidates));
  CSM_TRY(d_su.Reserve(sizeof(int) * num_candidates));
  CSM_CUDA(cudaMemcpyAsync(d_jobs.p, &jd, sizeof(jd), cudaMemcpyHostToDevice, s));
  CSM_CUDA(cudaMemcpyAsync(d_info.p, info.data(), sizeof(ScanInfo) * num_scans,
                           cudaMemcpyHostToDevice, s));
  CSM_CUDA(cudaMemcpyAsync(d_dscan.p, h_ds.data(), sizeof(short2) * npts,
                           cudaMemcpyHostToDevice, s));
  CSM_CUDA(cudaMemcpyAsync(d_lc.p, lc.data(), sizeof(ListCand) * num_candidates,
                           cudaMemcpyHostToDevice, s));
  k_score_list<<<DivUp(static_cast<long long>(num_candidates) * 32, 256), 256, 0, s>>>(
      d_jobs.as<JobDev>(), d_info.as<ScanInfo>(), d_dscan.as<short2>(), d_lc.as<ListCand>(),
      num_candidates, d_su.as<int>(), d_sc.as<float>());
  CSM_LAUNCH_CHECK();
  CSM_CUDA(cudaMemcpyAsync(scores, d_sc.p, sizeof(float) * num_candidates,
                           cudaMemcpyDeviceToHost, s));
  if (sums)
    CSM_CUDA(cudaMemcpyAsync(sums, d_su.p, sizeof(int) * num_candidates, cudaMemcpyDeviceToHost,
                             s));
  CSM_CUDA(cudaStreamSynchronize(s));
  return CSM_OK;
}

}  // extern "C"
