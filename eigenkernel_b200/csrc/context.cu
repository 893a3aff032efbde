// Context: device, stream, owned allocations, timing table (the B200 side of event_logger.f90:23-65).
#include "common.cuh"

#include <algorithm>

namespace ekb {

// Device memory comes from a per-context caching arena: a solve allocates and releases ~45 GB of stage workspaces
// (cudaMalloc / cudaFree of multi-GB blocks cost tens of milliseconds each and synchronise the device), so
// released blocks are kept and handed out again to the next request of (almost) the same size.  Everything the
// library launches is ordered on ctx->stream, so reuse needs no extra synchronisation.  On an allocation failure
// the cache is emptied and the request retried.
static size_t arena_round(size_t bytes) {
  if (bytes == 0) bytes = 16;
  const size_t g = bytes >= ((size_t)1 << 20) ? ((size_t)2 << 20) : 512;
  return (bytes + g - 1) / g * g;
}

void ctx_trim(Ctx* ctx) {
  for (auto& c : ctx->cache) cudaFree(c.first);
  ctx->cache.clear();
  ctx->cached_bytes = 0;
}

int ctx_alloc(Ctx* ctx, void** p, size_t bytes) {
  *p = nullptr;
  bytes = arena_round(bytes);
  // best fit among cached blocks that waste at most 1/8 of their size
  int best = -1;
  for (int i = 0; i < (int)ctx->cache.size(); ++i) {
    const size_t sz = ctx->cache[i].second;
    if (sz >= bytes && sz - bytes <= sz / 8 && (best < 0 || sz < ctx->cache[best].second)) best = i;
  }
  if (best >= 0) {
    *p = ctx->cache[best].first;
    ctx->live.emplace_back(*p, ctx->cache[best].second);
    ctx->cached_bytes -= ctx->cache[best].second;
    ctx->cache.erase(ctx->cache.begin() + best);
    ctx->allocs.push_back(*p);
    return 0;
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess && !ctx->cache.empty()) {
    cudaGetLastError();
    cudaStreamSynchronize(ctx->stream);
    ctx_trim(ctx);
    e = cudaMalloc(p, bytes);
  }
  if (e != cudaSuccess) {
    ctx->last_cuda = e;
    ctx->last_error = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
    cudaGetLastError();
    *p = nullptr;
    return EKB_ERR_NOMEM;
  }
  ctx->live.emplace_back(*p, bytes);
  ctx->allocs.push_back(*p);
  return 0;
}

int ctx_free(Ctx* ctx, void* p) {
  if (!p) return 0;
  auto it = std::find(ctx->allocs.begin(), ctx->allocs.end(), p);
  if (it != ctx->allocs.end()) ctx->allocs.erase(it);
  size_t bytes = 0;
  for (size_t i = 0; i < ctx->live.size(); ++i)
    if (ctx->live[i].first == p) {
      bytes = ctx->live[i].second;
      ctx->live.erase(ctx->live.begin() + i);
      break;
    }
  if (bytes == 0 || !ctx->cache_enabled) {
    cudaFree(p);
    return 0;
  }
  ctx->cache.emplace_back(p, bytes);
  ctx->cached_bytes += bytes;
  return 0;
}

// Side stream (higher priority) + two reusable events: panel look-ahead experiment of sy2sb, overlapped transfers of the
// host entry points.
int ctx_ensure_aux(Ctx* ctx) {
  if (!ctx->aux_stream) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
    EKB_CUDA(cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, hi));
  }
  for (auto& e : ctx->aux_ev)
    if (!e) EKB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return 0;
}

// Same accumulate-by-name semantics as add_event (event_logger.f90:45-64).
void ctx_add_event(Ctx* ctx, const char* name, double seconds) {
  for (auto& e : ctx->events) {
    if (e.name == name) {
      e.seconds += seconds;
      e.num_repeated += 1;
      return;
    }
  }
  ctx->events.push_back(Event{name, seconds, 1});
}

StageTimer::StageTimer(Ctx* c, const char* n) : ctx(c), name(n) {
  prev_stage = ctx->cur_stage;
  ctx->cur_stage = n;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a, ctx->stream);
}
StageTimer::~StageTimer() {
  if (!done) {
    ctx->cur_stage = prev_stage;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  }
}
double StageTimer::stop() {
  if (done) return 0.0;
  done = true;
  ctx->cur_stage = prev_stage;
  cudaEventRecord(b, ctx->stream);
  cudaEventSynchronize(b);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  double s = ms * 1e-3;
  ctx_add_event(ctx, name, s);
  return s;
}

}  // namespace ekb
