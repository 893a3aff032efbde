// Stage 2 of the two-stage tridiagonalization: symmetric band (half bandwidth b) -> tridiagonal by
// bulge chasing (Schwarz/Lang; what ELPA2's tridiag_band and LAPACK's dsytrd_sb2st do).  Together with
// sy2sb.cu this replaces pdsytrd('L'), reference src/solver_scalapack_all.f90:59, and the gather of d/e
// (allgather_row_wise, src/distribute_matrix.f90:431-478; solver_scalapack_all.f90:75-78).
//
// One CTA (4 b threads) owns one sweep at a time and keeps the sweep's moving b x b window in shared memory; sweeps are
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

__device__ __forceinline__ void bar_named(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// 4 B threads per CTA in four groups of B threads (whole warps).  A task's work is arranged around its ONE
// dependency on the previous sweep (the D and B blocks of the band):
//   group 0 needs nothing from the previous sweep (for t >= 1 the L block is already in shared memory): warp 0 forms
//   the reflector, then the group left-applies it to the L block and streams the block back -- all of this overlaps
//   the dependency wait and the band loads of the other groups;
//   groups 1-3: one thread polls the predecessor's progress counter, then the three groups issue ALL loads of the
//   D and B blocks before first use (one L2 round trip per task instead of 2 B dependent ones);
//   after one block barrier: group 1 forms p = tau D v and w = p - tau/2 (p.v) v, then groups 0, 1, 3 apply
//   D -= v w^T + w v^T straight back to the band (columns dealt mod 3) while group 2 right-applies the reflector to
//   the B block (which becomes the L block of the next task).
// Progress is published by thread 0 after the closing barrier with a single cumulative fence.
template <int B>
__global__ void __launch_bounds__(4 * B, 2) sb2st_kernel(double* __restrict__ AB, i64 ldab, i64 n, double* __restrict__ V2,
                                                         i64 ldv, double* __restrict__ TAU2, int ldtau,
                                                         int* __restrict__ prog, double* __restrict__ d_out,
                                                         double* __restrict__ e_out) {
  constexpr int LDS = B + 1;
  constexpr int CPT = (B + 2) / 3;  // columns of D / B loaded per thread of groups 1-3
  constexpr int EPL = B / 32;       // reflector elements per lane of warp 0
  constexpr int WPG = B / 32;       // warps per group
  extern __shared__ double sm[];
  double* buf0 = sm;                 // L / B blocks alternate between buf0 and buf1
  double* buf1 = sm + B * LDS;
  double* bufD = sm + 2 * B * LDS;
  double* v = sm + 3 * B * LDS;      // B
  double* w = v + B;                 // B
  __shared__ double red[4];
  __shared__ double s_tau;
  const int tid = threadIdx.x;
  const int grp = tid / B, li = tid % B;
  const int lane = tid & 31;
  const int G = gridDim.x;

  for (i64 s = blockIdx.x; s <= n - 3; s += G) {
    const int ntask = sb2st_num_tasks(n, B, s);
    double* bufL = buf0;
    double* bufB = buf1;
    for (int t = 0; t < ntask; ++t) {
      const i64 r0 = s + 1 + (i64)t * B;
      const int nr = (int)min((i64)B, n - r0);               // rows of R (>= 2)
      const int nr2 = (int)max((i64)0, min((i64)B, n - (r0 + B)));  // rows of the block below
      const int need = t + 3;                                // tasks 0 .. t+2 of sweep s-1 must be complete
      if (t == 0) {
        // the L block is just column s of the band, which the previous sweep has modified: wait first
        if (s > 0) {
          if (tid == 0) {
            while (ld_volatile(prog + (s - 1)) < need) {}
            __threadfence();
          }
          __syncthreads();
        }
        if (grp == 0 && li < nr) bufL[li] = __ldcg(AB + s * ldab + 1 + li);
      }
      if (grp == 0) {
        if (t == 0) bar_named(1, B);
        // ---- 1. reflector from x = bufL(0:nr, 0), by warp 0 alone
        if (tid < 32) {
          double xs[EPL];
          double sq = 0.0;
#pragma unroll
          for (int q = 0; q < EPL; ++q) {
            const int e = lane + 32 * q;
            xs[q] = (e > 0 && e < nr) ? bufL[e] : 0.0;
            sq += xs[q] * xs[q];
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
          const double alpha = bufL[0];
          double beta, tau, sc;
          if (sq == 0.0) {
            beta = alpha; tau = 0.0; sc = 0.0;
          } else {
            beta = -copysign(sqrt(alpha * alpha + sq), alpha);
            tau = (beta - alpha) / beta;
            sc = 1.0 / (alpha - beta);
          }
          __syncwarp();
          if (lane == 0) {
            s_tau = tau;
            TAU2[(i64)s * ldtau + t] = tau;
          }
#pragma unroll
          for (int q = 0; q < EPL; ++q) {
            const int e = lane + 32 * q;
            const double vi = (e == 0) ? 1.0 : (e < nr ? xs[q] * sc : 0.0);
            v[e] = vi;
            if (e < nr) {
              V2[s * ldv + r0 + e] = vi;
              bufL[e] = (e == 0) ? beta : 0.0;
            }
          }
        }
        bar_named(1, B);
        // ---- 2. left-apply H to bufL(:, 1:B) (t >= 1), thread per column; then write the L block back
        if (t >= 1) {
          const double tau = s_tau;
          const int jj = li;
          if (jj >= 1) {
            double dot = 0.0;
#pragma unroll 8
            for (int ii = 0; ii < nr; ++ii) dot += v[ii] * bufL[jj * LDS + ii];
            dot *= tau;
#pragma unroll 8
            for (int ii = 0; ii < nr; ++ii) bufL[jj * LDS + ii] -= dot * v[ii];
          }
          bar_named(1, B);
          if (li < nr) {
#pragma unroll 8
            for (int jj2 = 0; jj2 < B; ++jj2) AB[(r0 - B + jj2) * ldab + (B + li - jj2)] = bufL[jj2 * LDS + li];
          }
        } else {
          if (li < nr) AB[s * ldab + 1 + li] = bufL[li];
        }
      } else {
        // ---- wait for sweep s-1 (t >= 1; for t == 0 it happened above), then load the D and B blocks
        if (t > 0 && s > 0) {
          if (tid == B) {
            while (ld_volatile(prog + (s - 1)) < need) {}
            __threadfence();
          }
          bar_named(4, 3 * B);
        }
        double xd[CPT], xb[CPT];
        const int g3 = grp - 1;
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
          const int jj = g3 + 3 * q;
          const double* col = AB + (r0 + jj) * ldab;
          xd[q] = (jj < nr && li >= jj && li < nr) ? __ldcg(col + (li - jj)) : 0.0;
          xb[q] = (jj < nr && li < nr2) ? __ldcg(col + (B + li - jj)) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
          const int jj = g3 + 3 * q;
          if (jj < nr && li >= jj && li < nr) {  // D block: A(R,R), lower stored; mirrored into full
            bufD[jj * LDS + li] = xd[q];
            bufD[li * LDS + jj] = xd[q];
          }
          if (jj < nr && li < nr2) bufB[jj * LDS + li] = xb[q];  // B block: A(R+B, R)
        }
      }
      __syncthreads();
      const double tau = s_tau;
      if (grp == 2) {
        // ---- 4. right-apply on B: u = tau B v (thread per row), B -= u v^T
        if (li < nr2) {
          double u = 0.0;
#pragma unroll 8
          for (int jj = 0; jj < nr; ++jj) u += bufB[jj * LDS + li] * v[jj];
          u *= tau;
#pragma unroll 8
          for (int jj = 0; jj < nr; ++jj) bufB[jj * LDS + li] -= u * v[jj];
          // the B block becomes the L block of task t+1 (same sweep); if there is no task t+1 it must be stored
          if (t + 1 >= ntask) {
#pragma unroll 8
            for (int jj = 0; jj < nr; ++jj) AB[(r0 + jj) * ldab + (B + li - jj)] = bufB[jj * LDS + li];
          }
        }
      } else {
        // ---- 3. two-sided on D: p = tau D v (thread per row, group 1), w = p - tau/2 (p.v) v
        const double vl = v[li];
        if (grp == 1) {
          double pi = 0.0;
          if (li < nr) {
#pragma unroll 8
            for (int jj = 0; jj < nr; ++jj) pi += bufD[jj * LDS + li] * v[jj];
            pi *= tau;
          }
          double pv = (li < nr) ? pi * vl : 0.0;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) pv += __shfl_xor_sync(0xffffffffu, pv, o);
          if (lane == 0) red[li >> 5] = pv;
          bar_named(3, B);
          double ptv = 0.0;
#pragma unroll
          for (int q = 0; q < WPG; ++q) ptv += red[q];
          w[li] = (li < nr) ? pi - 0.5 * tau * ptv * vl : 0.0;
        }
        bar_named(2, 3 * B);
        // D -= v w^T + w v^T, lower part straight back to the band; columns dealt mod 3 to groups 0, 1, 3
        if (li < nr) {
          const double wi = w[li];
          const int c0 = grp == 0 ? 0 : (grp == 1 ? 1 : 2);
#pragma unroll 4
          for (int jj = c0; jj <= li; jj += 3)
            AB[(r0 + jj) * ldab + (li - jj)] = bufD[jj * LDS + li] - vl * w[jj] - wi * v[jj];
        }
      }
      // ---- publish progress: one barrier, then a single cumulative fence by the publishing thread
      __syncthreads();
      if (tid == 0) {
        __threadfence();
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
      EKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sb2st_kernel<64>, 256, smem));
      G = per_sm * ctx->num_sms;
    } else {
      EKB_CUDA(cudaFuncSetAttribute(sb2st_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0;
      EKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sb2st_kernel<32>, 128, smem));
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
      EKB_CUDA(cudaLaunchCooperativeKernel((void*)sb2st_kernel<64>, dim3(G), dim3(256), args, smem, ctx->stream));
    else
      EKB_CUDA(cudaLaunchCooperativeKernel((void*)sb2st_kernel<32>, dim3(G), dim3(128), args, smem, ctx->stream));
    EKB_COUNT_LAUNCH(ctx);
    EKB_TRY(prof_end(ctx));
  }
  extract_de_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(AB, ldab, n, d, e); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ekb
