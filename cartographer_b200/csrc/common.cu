#include "common.cuh"

namespace csm {

static thread_local char t_error[1024] = "";
std::atomic<int64_t> g_launches{0};

void SetError(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

static std::mutex g_ctx_mu;
static std::map<int, std::unique_ptr<Ctx>> g_ctx;

// The stream-ordered pool keeps what the workspaces release (no trimming at synchronisation
// points), so regrowths and the next call's allocations are served from the pool.
static void KeepPoolMemory(int device) {
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long threshold = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
}

csm_status GetCtx(int device, Ctx** out) {
  std::lock_guard<std::mutex> lock(g_ctx_mu);
  auto it = g_ctx.find(device);
  if (it != g_ctx.end()) {
    *out = it->second.get();
    return CSM_OK;
  }
  int count = 0;
  CSM_CUDA(cudaGetDeviceCount(&count));
  CSM_REQUIRE(device >= 0 && device < count, "device index out of range");
  CSM_CUDA(cudaSetDevice(device));
  KeepPoolMemory(device);
  std::unique_ptr<Ctx> ctx(new Ctx);
  ctx->device = device;
  cudaDeviceProp prop;
  CSM_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  CSM_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CSM_CUDA(cudaEventCreate(&ctx->ev0));
  CSM_CUDA(cudaEventCreate(&ctx->ev1));
  *out = ctx.get();
  g_ctx[device] = std::move(ctx);
  return CSM_OK;
}

static csm_status NewCtx(int device, std::unique_ptr<Ctx>* out) {
  KeepPoolMemory(device);
  std::unique_ptr<Ctx> ctx(new Ctx);
  ctx->device = device;
  cudaDeviceProp prop;
  CSM_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  CSM_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CSM_CUDA(cudaEventCreate(&ctx->ev0));
  CSM_CUDA(cudaEventCreate(&ctx->ev1));
  *out = std::move(ctx);
  return CSM_OK;
}

static std::mutex g_lane_mu;
static std::map<int, std::vector<std::unique_ptr<Ctx>>> g_lanes;
static std::atomic<unsigned> g_lane_rr{0};
constexpr size_t kMaxLanes = 16;

csm_status AcquireLane(int device, LaneGuard* out) {
  {
    std::lock_guard<std::mutex> lock(g_lane_mu);
    std::vector<std::unique_ptr<Ctx>>& lanes = g_lanes[device];
    for (auto& l : lanes) {
      std::unique_lock<std::mutex> lk(l->mu, std::try_to_lock);
      if (lk.owns_lock()) {
        out->lane = l.get();
        out->lock = std::move(lk);
        return CSM_OK;
      }
    }
    if (lanes.size() < kMaxLanes) {
      int count = 0;
      CSM_CUDA(cudaGetDeviceCount(&count));
      CSM_REQUIRE(device >= 0 && device < count, "device index out of range");
      CSM_CUDA(cudaSetDevice(device));
      std::unique_ptr<Ctx> ctx;
      CSM_TRY(NewCtx(device, &ctx));
      lanes.push_back(std::move(ctx));
      out->lane = lanes.back().get();
      out->lock = std::unique_lock<std::mutex>(out->lane->mu);
      return CSM_OK;
    }
  }
  // all lanes busy: wait for one (round robin)
  Ctx* l;
  {
    std::lock_guard<std::mutex> lock(g_lane_mu);
    std::vector<std::unique_ptr<Ctx>>& lanes = g_lanes[device];
    l = lanes[g_lane_rr.fetch_add(1) % lanes.size()].get();
  }
  out->lane = l;
  out->lock = std::unique_lock<std::mutex>(l->mu);
  return CSM_OK;
}

std::atomic<int> g_profile_on{0};
struct ProfEntry { double ms = 0; long long launches = 0; double units = 0; };
static std::mutex g_prof_mu;
static std::map<std::string, ProfEntry> g_prof;
static std::map<Ctx*, std::pair<cudaEvent_t, cudaEvent_t>> g_prof_ev;

void ProfBegin(Ctx* ctx) {
  if (!g_profile_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  auto it = g_prof_ev.find(ctx);
  if (it == g_prof_ev.end()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    it = g_prof_ev.emplace(ctx, std::make_pair(a, b)).first;
  }
  cudaEventRecord(it->second.first, ctx->stream);
}

void ProfStop(Ctx* ctx) {
  if (!g_profile_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  auto it = g_prof_ev.find(ctx);
  if (it == g_prof_ev.end()) return;
  cudaEventRecord(it->second.second, ctx->stream);
}

void ProfCommit(Ctx* ctx, const char* name, double units) {
  if (!g_profile_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  auto it = g_prof_ev.find(ctx);
  if (it == g_prof_ev.end()) return;
  cudaEventSynchronize(it->second.second);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, it->second.first, it->second.second);
  ProfEntry& e = g_prof[name];
  e.ms += ms;
  e.launches += 1;
  e.units += units;
}

}  // namespace csm

extern "C" {

csm_status csm_profile_enable(int32_t on) {
  csm::g_profile_on.store(on ? 1 : 0);
  if (on) {
    std::lock_guard<std::mutex> lock(csm::g_prof_mu);
    csm::g_prof.clear();
  }
  return CSM_OK;
}

// Writes one line per kernel: "<name> <launches> <total_ms> <units>\n".
csm_status csm_profile_read(char* buf, int32_t cap) {
  CSM_REQUIRE(buf != nullptr && cap > 0, "buffer");
  std::lock_guard<std::mutex> lock(csm::g_prof_mu);
  std::string out;
  for (const auto& kv : csm::g_prof) {
    char line[256];
    snprintf(line, sizeof(line), "%s %lld %.6f %.0f\n", kv.first.c_str(), kv.second.launches,
             kv.second.ms, kv.second.units);
    out += line;
  }
  snprintf(buf, cap, "%s", out.c_str());
  return CSM_OK;
}


csm_status csm_device_count(int32_t* count) {
  CSM_REQUIRE(count != nullptr, "count is null");
  int c = 0;
  CSM_CUDA(cudaGetDeviceCount(&c));
  *count = c;
  return CSM_OK;
}

const char* csm_last_error_string(void) { return csm::t_error; }

int64_t csm_kernel_launch_count(void) { return csm::g_launches.load(); }

}  // extern "C"
