// Stage 2 of the two-stage tridiagonalization: symmetric band (half bandwidth b) -> tridiagonal by
// bulge chasing (Schwarz/Lang; what ELPA2's tridiag_band and LAPACK's dsytrd_sb2st do).  Together with
// sy2sb.cu this replaces pdsytrd('L'), reference src/solver_scalapack_all.f90:59, and the gather of d/e
// (allgather_row_wise, src/distribute_matrix.f90:431-478; solver_scalapack_all.f90:75-78).
//
// One CTA owns one sweep at a time and keeps the sweep's moving b x b window in shared memory; sweeps are
// pipelined across the resident CTAs through per-sweep progress counters in global memory (sweep s may
// run task t once sweep s-1 has finished task t+2).  The band lives in L2 (8*2b*n bytes = 32 MiB at
// n = 32768, b = 64); reflectors stream out to HBM in the layout the back-transformation consumes:
//   V2(r, s): column s holds the concatenated reflectors of sweep s (task t occupies rows s+1+t*b ..),
//   TAU2(t, s).
#include "common.cuh"

namespace ekb {

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// number of chase tasks of sweep s (tasks whose row block has at least 2 rows)
__host__ __device__ __forceinline__ int sb2st_num_tasks(i64 n, int b, i64 s) {
  if (s > n - 3) return 0;
  return (int)((n - 3 - s) / b) + 1;
}

template <int B>
__global__ void __launch_bounds__(B) sb2st_kernel(double* __restrict__ AB, i64 ldab, i64 n, double* __restrict__ V2,
                                                  i64 ldv, double* __restrict__ TAU2, int ldtau, int* __restrict__ prog,
                                                  double* __restrict__ d_out, double* __restrict__ e_out) {
  constexpr int LDS = B + 1;
  extern __shared__ double sm[];
  double* buf0 = sm;                 // L / B blocks alternate between buf0 and buf1
  double* buf1 = sm + B * LDS;
  double* bufD = sm + 2 * B * LDS;
  double* v = sm + 3 * B * LDS;      // B
  double* w = v + B;                 // B
  __shared__ double red[4];
  __shared__ double s_tau, s_beta, s_scal;
  const int tid = threadIdx.x;
  const int G = gridDim.x;

  for (i64 s = blockIdx.x; s <= n - 3; s += G) {
    const int ntask = sb2st_num_tasks(n, B, s);
    double* bufL = buf0;
    double* bufB = buf1;
    for (int t = 0; t < ntask; ++t) {
      const i64 r0 = s + 1 + (i64)t * B;
      const int nr = (int)min((i64)B, n - r0);               // rows of R (>= 2)
      const int nr2 = (int)max((i64)0, min((i64)B, n - (r0 + B)));  // rows of the block below
      // ---- wait for sweep s-1
      if (s > 0) {
        if (tid == 0) {
          const int need = t + 3;
          while (ld_volatile(prog + (s - 1)) < need) { __nanosleep(20); }
          __threadfence();
        }
        __syncthreads();
      }
      // ---- load
      if (t == 0) {
        // L block is just column s: rows s+1..s+nr (offsets 1..nr)
        if (tid < nr) bufL[tid] = __ldcg(AB + s * ldab + 1 + tid);
      }
      // D block: A(R,R), lower stored; mirror into full
      for (int jj = 0; jj < nr; ++jj) {
        const int ii = tid;
        if (ii >= jj && ii < nr) {
          double x = __ldcg(AB + (r0 + jj) * ldab + (ii - jj));
          bufD[jj * LDS + ii] = x;
          bufD[ii * LDS + jj] = x;
        }
      }
      // B block: A(R+B, R): element (ii,jj) at offset B + ii - jj of column r0+jj
      for (int jj = 0; jj < nr; ++jj) {
        const int ii = tid;
        if (ii < nr2) bufB[jj * LDS + ii] = __ldcg(AB + (r0 + jj) * ldab + (B + ii - jj));
      }
      __syncthreads();
      // ---- 1. reflector from x = bufL(0:nr, 0)
      {
        double xi = (tid > 0 && tid < nr) ? bufL[tid] : 0.0;
        double sq = xi * xi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if ((tid & 31) == 0) red[tid >> 5] = sq;
        __syncthreads();
        if (tid == 0) {
          double xn2 = 0.0;
          for (int q = 0; q < B / 32; ++q) xn2 += red[q];
          double alpha = bufL[0];
          double beta, tau, sc;
          if (xn2 == 0.0) {
            beta = alpha; tau = 0.0; sc = 0.0;
          } else {
            beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
            tau = (beta - alpha) / beta;
            sc = 1.0 / (alpha - beta);
          }
          s_tau = tau; s_beta = beta; s_scal = sc;
          TAU2[(i64)s * ldtau + t] = tau;
        }
        __syncthreads();
        const double sc = s_scal;
        double vi = 0.0;
        if (tid == 0) vi = 1.0;
        else if (tid < nr) vi = xi * sc;
        v[tid] = vi;
        if (tid < nr) V2[s * ldv + r0 + tid] = vi;
        if (tid == 0) bufL[0] = s_beta;
        else if (tid < nr) bufL[tid] = 0.0;
      }
      __syncthreads();
      const double tau = s_tau;
      // ---- 2. left-apply H to bufL(:, 1:B) (t >= 1), thread per column
      if (t >= 1) {
        const int jj = tid;
        if (jj >= 1) {
          double dot = 0.0;
          for (int ii = 0; ii < nr; ++ii) dot += v[ii] * bufL[jj * LDS + ii];
          dot *= tau;
          for (int ii = 0; ii < nr; ++ii) bufL[jj * LDS + ii] -= dot * v[ii];
        }
      }
      // ---- 4. two-sided on D: p = tau D v (thread per row)
      double pi = 0.0;
      if (tid < nr) {
        for (int jj = 0; jj < nr; ++jj) pi += bufD[jj * LDS + tid] * v[jj];
        pi *= tau;
      }
      {
        double pv = (tid < nr) ? pi * v[tid] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pv += __shfl_xor_sync(0xffffffffu, pv, o);
        __syncthreads();  // red reuse + bufL column updates complete
        if ((tid & 31) == 0) red[tid >> 5] = pv;
        __syncthreads();
        double ptv = 0.0;
        for (int q = 0; q < B / 32; ++q) ptv += red[q];
        const double wi = (tid < nr) ? pi - 0.5 * tau * ptv * v[tid] : 0.0;
        w[tid] = wi;
      }
      __syncthreads();
      // ---- 3. write back L block (final for this sweep)
      if (t == 0) {
        if (tid < nr) AB[s * ldab + 1 + tid] = bufL[tid];
      } else {
        for (int jj = 0; jj < B; ++jj) {
          const int ii = tid;
          if (ii < nr) AB[(r0 - B + jj) * ldab + (B + ii - jj)] = bufL[jj * LDS + ii];
        }
      }
      // D -= v w^T + w v^T (thread per row), write back lower part
      if (tid < nr) {
        const double vi = v[tid], wi = w[tid];
        for (int jj = 0; jj <= tid; ++jj) {
          double x = bufD[jj * LDS + tid] - vi * w[jj] - wi * v[jj];
          AB[(r0 + jj) * ldab + (tid - jj)] = x;
        }
      }
      // ---- 6. right-apply on B: u = B v (thread per row), B -= tau u v^T
      if (tid < nr2) {
        double u = 0.0;
        for (int jj = 0; jj < nr; ++jj) u += bufB[jj * LDS + tid] * v[jj];
        u *= tau;
        for (int jj = 0; jj < nr; ++jj) bufB[jj * LDS + tid] -= u * v[jj];
      }
      // the B block becomes the L block of task t+1 (same sweep); if there is no task t+1 it must be stored
      if (t + 1 >= ntask) {
        if (tid < nr2)
          for (int jj = 0; jj < nr; ++jj) AB[(r0 + jj) * ldab + (B + tid - jj)] = bufB[jj * LDS + tid];
      }
      // ---- publish progress
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        const int val = (t + 1 >= ntask) ? 0x3fffffff : (t + 1);
        *reinterpret_cast<volatile int*>(prog + s) = val;
      }
      double* tmp = bufL; bufL = bufB; bufB = tmp;
    }
    if (ntask == 0 && tid == 0) *reinterpret_cast<volatile int*>(prog + s) = 0x3fffffff;
  }
}

__global__ void extract_de_kernel(const double* __restrict__ AB, i64 ldab, i64 n, double* __restrict__ d, double* __restrict__ e) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  d[i] = AB[i * ldab];
  if (i < n - 1) e[i] = AB[i * ldab + 1];
}

int sb2st_max_tasks(i64 n, int b) { return (int)(n / b) + 2; }

// AB: ldab >= 2b rows (rows b+1.. must be zero on entry).  V2: n x (n-2 or more), ldv; TAU2: ldtau x (n-2).
// d (n), e (n-1) receive the tridiagonal.  prog: int workspace of n entries.
int sb2st(Ctx* ctx, i64 n, int b, double* AB, i64 ldab, double* V2, i64 ldv, double* TAU2, int ldtau, int* prog,
          double* d, double* e) {
  if (n <= 0) return 0;
  if (n >= 3) {
    EKB_CUDA(cudaMemsetAsync(prog, 0, (size_t)n * sizeof(int), ctx->stream));
    const size_t smem = (size_t)(3 * b * (b + 1) + 2 * b) * sizeof(double);
    int G = 0;
    if (b == 64) {
      EKB_CUDA(cudaFuncSetAttribute(sb2st_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0;
      EKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sb2st_kernel<64>, 64, smem));
      G = per_sm * ctx->num_sms;
    } else {
      EKB_CUDA(cudaFuncSetAttribute(sb2st_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0;
      EKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sb2st_kernel<32>, 32, smem));
      G = per_sm * ctx->num_sms;
    }
    if (G < 1) return EKB_ERR_INTERNAL;
    // all CTAs must be co-resident (spin-wait dependencies): cooperative launch enforces it
    i64 maxuse = (n / b) / 3 + 2;
    if (G > maxuse) G = (int)maxuse;
    if (G > n - 2) G = (int)(n - 2);
    void* args[] = {(void*)&AB, (void*)&ldab, (void*)&n, (void*)&V2, (void*)&ldv, (void*)&TAU2, (void*)&ldtau,
                    (void*)&prog, (void*)&d, (void*)&e};
    EKB_TRY(prof_begin(ctx, PROF_SB2ST, 12.0 * b * (double)n * (double)n));  // effective bytes, SURVEY 8(d)
    if (b == 64)
      EKB_CUDA(cudaLaunchCooperativeKernel((void*)sb2st_kernel<64>, dim3(G), dim3(64), args, smem, ctx->stream));
    else
      EKB_CUDA(cudaLaunchCooperativeKernel((void*)sb2st_kernel<32>, dim3(G), dim3(32), args, smem, ctx->stream));
    EKB_COUNT_LAUNCH(ctx);
    EKB_TRY(prof_end(ctx));
  }
  extract_de_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(AB, ldab, n, d, e); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ekb
