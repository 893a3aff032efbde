// FP64 peak microbenchmarks: the roofline DENOMINATOR for the tensor-bound stages.
// MEASURED_PEAKS.json carries no FP64 figure, so the library measures the
// register-resident DMMA.8x8x4 issue rate and the DFMA rate itself.  Never on the solve path.
#include "common.cuh"

namespace ekb {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i * 1e-3;
  double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}

int measure_fp64_peak(Ctx* ctx, double* dmma_tflops, double* dfma_tflops) {
  double* d_out = nullptr;
  EKB_TRY(ctx_alloc(ctx, (void**)&d_out, 64));
  cudaEvent_t e0, e1;
  EKB_CUDA(cudaEventCreate(&e0));
  EKB_CUDA(cudaEventCreate(&e1));
  const int iters = 4096, ctas = ctx->num_sms * 4;
  float ms = 0.f;
  double best = 0.0;
  // The figure wanted is the issue-rate peak at boost clocks.  Two things bias a short measurement low: a cold GPU
  // sits at idle clocks for the first milliseconds of load (5 x 4 ms right after context creation gave 29.6 TF instead
  // of 37.0 on the same box), and a GPU that has been saturating the tensor pipe for a while may be power-capped to
  // ~1.57 GHz (the same 29.7 TF, seen on a 2-GPU run after a 0.3 s warm-up).  So: 64 back-to-back repetitions of
  // 4 ms, keep the BEST one -- it is taken after the ramp and before any cap.
  for (int rep = 0; rep < 64; ++rep) {
    EKB_CUDA(cudaEventRecord(e0, ctx->stream));
    dmma_peak_kernel<<<ctas, 256, 0, ctx->stream>>>(d_out, iters); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaEventRecord(e1, ctx->stream));
    EKB_CUDA(cudaEventSynchronize(e1));
    EKB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double fl = (double)ctas * 8 /*warps*/ * iters * 16.0 * (2.0 * 8 * 8 * 4);
    double tf = fl / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  *dmma_tflops = best;
  best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    EKB_CUDA(cudaEventRecord(e0, ctx->stream));
    dfma_peak_kernel<<<ctas, 256, 0, ctx->stream>>>(d_out, iters); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaEventRecord(e1, ctx->stream));
    EKB_CUDA(cudaEventSynchronize(e1));
    EKB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double fl = (double)ctas * 256 * iters * 16.0 * 2.0;
    double tf = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  *dfma_tflops = best;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ctx_free(ctx, d_out);
  return 0;
}

}  // namespace ekb
