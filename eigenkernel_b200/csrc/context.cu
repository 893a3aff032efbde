// Context: device, stream, owned allocations, timing table (the B200 side of event_logger.f90:23-65).
#include "common.cuh"

#include <algorithm>

namespace ekb {

int ctx_alloc(Ctx* ctx, void** p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {
    ctx->last_cuda = e;
    ctx->last_error = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
    cudaGetLastError();
    return EKB_ERR_NOMEM;
  }
  ctx->allocs.push_back(*p);
  return 0;
}

int ctx_free(Ctx* ctx, void* p) {
  if (!p) return 0;
  auto it = std::find(ctx->allocs.begin(), ctx->allocs.end(), p);
  if (it != ctx->allocs.end()) ctx->allocs.erase(it);
  cudaFree(p);
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
