// Microbenchmark: DMMA.8x8x4 throughput as a function of warps per SM sub-partition and independent
// accumulator chains per warp, with operands in registers or re-read from shared memory each step.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int C, bool LDS>
__global__ void k(double* out, int iters) {
  __shared__ double s[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = 1.0 + i * 1e-9;
  __syncthreads();
  double acc[C][2];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c][0] = acc[c][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  const double* p = s + (threadIdx.x & 31) * 5;
  for (int it = 0; it < iters; ++it) {
    if (LDS) {
      // one A fragment per pair of MMAs, like the back-transformation kernels
#pragma unroll
      for (int c = 0; c < C; c += 2) {
        double af = p[((it * C + c) * 37) & 2047];
        dmma(acc[c][0], acc[c][1], b, af);
        if (c + 1 < C) dmma(acc[c + 1][0], acc[c + 1][1], a, af);
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) dmma(acc[c][0], acc[c][1], a, b);
    }
  }
  double t = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) t += acc[c][0] + acc[c][1];
  if (t == 123.456) out[0] = t;
}
template <int C, bool LDS>
void run(int warps_per_smsp, double* d) {
  int threads = 128 * warps_per_smsp, iters = 20000 / C * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); k<C, LDS><<<148, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double fl = 148.0 * (threads / 32) * (double)iters * C * 512.0;
  printf("chains/warp=%2d warps/SMSP=%d lds=%d : %6.2f TF/s\n", C, warps_per_smsp, (int)LDS, fl / (best * 1e-3) / 1e12);
}
int main() {
  double* d; cudaMalloc(&d, 64);
  for (int w : {1, 2, 3, 4, 8}) { run<1, false>(w, d); run<2, false>(w, d); run<4, false>(w, d); run<8, false>(w, d); run<12, false>(w, d); run<16, false>(w, d); }
  for (int w : {1, 2, 3, 4, 8}) { run<2, true>(w, d); run<4, true>(w, d); run<8, true>(w, d); run<12, true>(w, d); }
  return 0;
}
