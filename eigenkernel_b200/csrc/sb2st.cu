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


// ------------------------------------------------------------------------------------------ register-resident kernel
// Second-generation bulge chasing (round 2).  Same sweep-per-CTA pipeline, same band / V2 / TAU2 layouts, but built
// for LATENCY: the per-sweep critical path is what bounds this stage (n sweeps, each waiting for its predecessor),
// so every task is arranged as two short all-thread phases around two CTA barriers, nothing in the window ever
// lives in shared memory, and a sweep may follow its predecessor at a distance of TWO tasks instead of three.
//
//   * Data layout in registers: warp w owns columns [w CW, (w+1) CW) of every b x b block, lane l owns rows l (and
//     l + 32 for b = 64): all band loads / stores are 256-byte contiguous per warp instruction, each element is
//     loaded once and stored once.  The B block of task t stays in registers and IS the L block of task t+1.
//   * Task t, phase A (needs only v_t): left-apply on the L block (column dots by a transposed butterfly of warp
//     shuffles, rank-1 update, store) -- this overlaps the wait for sweep s-1; then the dependency poll (one
//     ld.acquire per warp), the loads of D_t (lower triangle) and B_t, and the partial dots of the symmetric
//     mat-vec p = tau D v and of u = tau B v (row parts through shared memory, column part by shuffles).
//   * Phase B (after the one barrier): every warp redundantly finishes p, sigma = p.v and w = p - tau/2 sigma v
//     (bit-identical in all warps), applies D -= v w^T + w v^T (lower part) and B -= u v^T in registers, stores D,
//     and warp 0 forms the NEXT reflector from column 0 of the new B block (look-ahead).
//   * Look-ahead makes the distance 2: sweep s+1 needs from task t+2 of sweep s exactly one element, the beta of that
//     task's reflector; task t+1 now computes it and writes it into the band before its completion is published.
//     (Consequently the L block is stored WITHOUT its (0,0) element: by then a later sweep may have consumed and
//     overwritten it.)  The schedule, including every read/write overlap between unordered tasks, is replayed on the
//     CPU in tests/test_host_sb2st_schedule.py.
//   * Publication: all stores of task t are issued before the barrier that opens task t+1; thread 0 then does one
//     st.release.gpu of the task count.  Consumers poll with ld.relaxed.gpu and read the band with ld.cg.
// Progress counters are polled with a strong relaxed load: ld.acquire would add an L1 invalidation (CCTL.IVALL) to
// every iteration of the spin loop -- 16 % of all stall samples in the first capture of this kernel -- and buys
// nothing here: every band element is read with ld.cg (L2, the point of coherence the counter itself is read from),
// so there is no stale L1 line to drop; the loads are issued after the loop exits (control dependency, in-order issue).
__device__ __forceinline__ int ld_poll_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Wait until *p >= need.  One poll costs an L2 round trip (~1 us here), so a single load in flight notices a
// change half a round trip late on average; four loads kept in flight, a quarter of a round trip apart, cut that to
// an eighth.  (The counter only grows, any observed value >= need is final.)
__device__ __forceinline__ void wait_progress(const int* p, int need) {
  int v0 = ld_poll_gpu(p);
  if (v0 >= need) return;
  const long long t0 = clock64();
  int v1 = ld_poll_gpu(p);
  while (clock64() - t0 < 250) {}
  int v2 = ld_poll_gpu(p);
  while (clock64() - t0 < 500) {}
  int v3 = ld_poll_gpu(p);
  for (;;) {
    if (v1 >= need) return;
    v1 = ld_poll_gpu(p);
    if (v2 >= need) return;
    v2 = ld_poll_gpu(p);
    if (v3 >= need) return;
    v3 = ld_poll_gpu(p);
    v0 = ld_poll_gpu(p);
    if (v0 >= need) return;
  }
}
__device__ __forceinline__ double ldcg_f64(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double shx(double x, int m) { return __shfl_xor_sync(0xffffffffu, x, m); }
__device__ __forceinline__ double shi(double x, int src) { return __shfl_sync(0xffffffffu, x, src); }

// Column sums over the 32 lanes by a transposed butterfly (reduce-scatter: CW/2 + CW/4 + .. exchanges, then a plain
// butterfly on the one remaining value).  Returns the total of column warp_col_of_lane<CW>(lane); the 32 / CW lanes that
// share a column hold bit-identical values.
template <int CW>
__device__ __forceinline__ int warp_col_of_lane(int lane) {
  return CW == 8 ? (lane >> 2) : (lane >> 3);
}
template <int CW>
__device__ __forceinline__ double warp_reduce_cols(const double (&y)[CW], int lane) {
  static_assert(CW == 8 || CW == 4, "columns per warp");
  double c1;
  if constexpr (CW == 8) {
    double a[4], b2[2];
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = (h4 ? y[k + 4] : y[k]) + shx(h4 ? y[k] : y[k + 4], 16);
#pragma unroll
    for (int k = 0; k < 2; ++k) b2[k] = (h3 ? a[k + 2] : a[k]) + shx(h3 ? a[k] : a[k + 2], 8);
    c1 = (h2 ? b2[1] : b2[0]) + shx(h2 ? b2[0] : b2[1], 4);
    c1 += shx(c1, 2);
    c1 += shx(c1, 1);  // column 4 bit4 + 2 bit3 + bit2 = lane >> 2
  } else {
    double a[2];
    const bool h4 = lane & 16, h3 = lane & 8;
#pragma unroll
    for (int k = 0; k < 2; ++k) a[k] = (h4 ? y[k + 2] : y[k]) + shx(h4 ? y[k] : y[k + 2], 16);
    c1 = (h3 ? a[1] : a[0]) + shx(h3 ? a[0] : a[1], 8);
    c1 += shx(c1, 4);
    c1 += shx(c1, 2);
    c1 += shx(c1, 1);  // column 2 bit4 + bit3 = lane >> 3
  }
  return c1;
}

// sum of NW values p[0], p[stride], ... as a balanced tree (depth log2 NW instead of NW dependent additions); every
// caller uses this one order, so the redundant copies of u and p in the different warps are bit-identical
template <int NW>
__device__ __forceinline__ double tree_sum(const double* p, int stride) {
  double t[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w) t[w] = p[w * stride];
#pragma unroll
  for (int m = NW / 2; m >= 1; m >>= 1)
#pragma unroll
    for (int w = 0; w < m; ++w) t[w] += t[w + m];
  return t[0];
}

__device__ __forceinline__ double warp_allreduce_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += shx(x, o);
  return x;
}

// Householder reflector (dlarfg convention) of the vector whose element `row = lane + 32 h` is x[h], rows < nrow
// valid; all lanes return the same tau and beta; v has the leading 1 explicit and zeros beyond nrow.
template <int RH>
__device__ __forceinline__ void warp_reflector(const double (&x)[RH], int nrow, int lane, double (&v)[RH], double& tau,
                                               double& beta) {
  double sq = 0.0;
#pragma unroll
  for (int h = 0; h < RH; ++h) {
    const int row = lane + 32 * h;
    if (row >= 1 && row < nrow) sq += x[h] * x[h];
  }
  sq = warp_allreduce_sum(sq);
  const double alpha = shi(x[0], 0);
  double sc;
  if (sq == 0.0) {
    beta = alpha; tau = 0.0; sc = 0.0;
  } else {
    beta = -copysign(sqrt(alpha * alpha + sq), alpha);
    tau = (beta - alpha) / beta;
    sc = 1.0 / (alpha - beta);
  }
#pragma unroll
  for (int h = 0; h < RH; ++h) {
    const int row = lane + 32 * h;
    v[h] = (row == 0) ? 1.0 : (row < nrow ? x[h] * sc : 0.0);
  }
}

// TRACE (development only, EKB200_SB2ST_TRACE=<file>): lane 0 of warps 0 and 1 of one CTA record clock64() at the
// phase boundaries of its first SB2ST_TRACE_TASKS tasks; the production instantiation carries none of it.
constexpr int SB2ST_TRACE_TASKS = 4096, SB2ST_TRACE_PTS = 8, SB2ST_TRACE_CTA = 5;
// RW: one extra warp (warp NW) does nothing but form reflectors, so that the next reflector (norm, square root, two
// divisions: the longest scalar chain of a task) runs BESIDE the rank-1/2 updates of the compute warps instead of
// after warp 0's share of them.
template <int B, int NW, bool RW, bool TRACE>
// (Tried: __maxnreg__(112) so that two 9-warp CTAs fit per SM -- at n = 32768 the pipeline wants ~260 sweeps in flight
// and one CTA per SM limits the first 40 % of the sweeps to 148.  Measured 0.460 s against 0.451 s: the two CTAs slow
// each other down by what the extra residency gains.  One CTA per SM with the registers it wants stays.)
__global__ void __launch_bounds__(32 * (NW + (RW ? 1 : 0)), ((NW <= 8 && !RW) ? 2 : 1))
sb2st_reg_kernel(double* __restrict__ AB, i64 ldab, i64 n, double* __restrict__ V2, i64 ldv, double* __restrict__ TAU2,
                 int ldtau, int* __restrict__ prog, long long* __restrict__ trace) {
  constexpr int RH = B / 32;   // rows per lane
  constexpr int CW = B / NW;   // columns per warp
  constexpr int DONE = 0x3fffffff;
  constexpr int RWARP = RW ? NW : 0;  // the warp that forms the reflectors
  static_assert(B % 32 == 0 && B % NW == 0, "geometry");
  __shared__ double s_v[2][B];
  __shared__ double s_tau[2];
  __shared__ double s_beta[2];
  __shared__ double s_x[B];        // RW: column 0 of the B block as loaded (before the right-apply)
  __shared__ double s_pu[NW][B];
  __shared__ double s_pp[NW][B];
  __shared__ double s_pc[B];
  __shared__ __align__(16) double s_y[NW][CW];  // column dots of the L block, per warp
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool compute = warp < NW;
  const int j0 = warp * CW;
  // The band is stored with exactly 2 B rows (what the solver allocates; the host falls back to the round-1 kernel
  // otherwise): every band offset below is then an immediate, which removes most of the integer arithmetic of a task
  // -- and the per-warp instruction count IS the task time here (about 2 warps per scheduler, dependent chains).
  constexpr i64 LDAB = 2 * B;
  constexpr i64 cstep = LDAB - 1;  // one column to the right along a matrix row, in band storage
  (void)ldab;
  int trace_task = 0;
  auto stamp = [&](int k, int slot) {
    if constexpr (TRACE) {
      if (blockIdx.x == SB2ST_TRACE_CTA && lane == 0 && trace_task < SB2ST_TRACE_TASKS)
        trace[((i64)trace_task * 2 + slot) * SB2ST_TRACE_PTS + k] = clock64();
    }
  };
  const int tslot = warp == 0 ? 0 : 1;
  const bool traced = warp < 2;

  for (i64 s = blockIdx.x; s <= n - 3; s += gridDim.x) {
    const int ntask = sb2st_num_tasks(n, B, s);
    // ---- prologue: reflector of task 0 from column s (needs task 0 of sweep s-1, look-ahead included)
    if (warp == RWARP) {
      if (s > 0) {
        if (lane == 0) wait_progress(prog + (s - 1), 1);
        __syncwarp();
      }
      const i64 r0 = s + 1;
      const int nr = (int)min((i64)B, n - r0);
      double x[RH], v[RH], tau, beta;
#pragma unroll
      for (int h = 0; h < RH; ++h) {
        const int row = lane + 32 * h;
        x[h] = row < nr ? __ldcg(AB + s * LDAB + 1 + row) : 0.0;
      }
      warp_reflector<RH>(x, nr, lane, v, tau, beta);
#pragma unroll
      for (int h = 0; h < RH; ++h) {
        const int row = lane + 32 * h;
        s_v[0][row] = v[h];
        if (row < nr) {
          V2[s * ldv + r0 + row] = v[h];
          AB[s * LDAB + 1 + row] = (row == 0) ? beta : 0.0;
        }
      }
      if (lane == 0) {
        s_tau[0] = tau;
        TAU2[s * (i64)ldtau] = tau;
      }
    }
    double Bt[RH][CW], Dt[RH][CW];
#pragma unroll
    for (int h = 0; h < RH; ++h)
#pragma unroll
      for (int c = 0; c < CW; ++c) Bt[h][c] = Dt[h][c] = 0.0;
    double* dprev = nullptr;  // D block of the previous task: updated in registers, all columns but the first not yet stored
    int nrprev = 0;
    // D_t goes to the band in two parts.  Column 0 (and the beta of the look-ahead) is all that the task which waits
    // for THIS task's completion reads of it; it is stored at once.  The other columns are only read one task later,
    // i.e. under the NEXT completion count, so they are stored after this task's count has been published: the
    // release then has no fresh stores to wait for (it used to sit behind 16 KB written a few hundred clocks before).
    auto store_deferred_D = [&]() {
      if (dprev == nullptr) return;
      if (nrprev == B) {
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
          for (int h = 0; h < RH; ++h) {
            const int row = lane + 32 * h, col = j0 + c;
            if (col > 0 && row >= col) dprev[c * cstep + 32 * h] = Dt[h][c];
          }
      } else {
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
          for (int h = 0; h < RH; ++h) {
            const int row = lane + 32 * h, col = j0 + c;
            if (col > 0 && col < nrprev && row >= col && row < nrprev) dprev[c * cstep + 32 * h] = Dt[h][c];
          }
      }
      dprev = nullptr;
    };

    for (int t = 0; t < ntask; ++t) {
      const int cur = t & 1;
      __syncthreads();  // v_t, tau_t visible; every store of task t-1 has been issued
      if (tid == 0 && t > 0) st_release_gpu(prog + s, t);  // tasks 0 .. t-1 complete (and beta_t is in the band)
      if (traced) stamp(0, tslot);
      if (compute) store_deferred_D();
      const double tau = s_tau[cur];
      const i64 r0 = s + 1 + (i64)t * B;
      const int nr = (int)min((i64)B, n - r0);                             // rows of R (>= 2)
      const int nr2 = (int)max((i64)0, min((i64)B, n - (r0 + B)));         // rows of the block below
      const bool last = t + 1 >= ntask;
      double vr[RH];
#pragma unroll
      for (int h = 0; h < RH; ++h) vr[h] = s_v[cur][lane + 32 * h];
      if (compute) {
        const double* vc = &s_v[cur][j0];  // v at this warp's columns: broadcast reads where needed (keeps registers free)
        // ---- L block (the B block of task t-1, columns R - B, rows R): H L, store.  Needs nothing from sweep s-1.
        if (t > 0) {
          if (RW && warp == 0) {  // column 0 became beta e1 when the reflector was formed (by the reflector warp)
            const double bt = s_beta[cur];
#pragma unroll
            for (int h = 0; h < RH; ++h) Bt[h][0] = (lane + 32 * h == 0) ? bt : 0.0;
          }
          double y[CW];
#pragma unroll
          for (int c = 0; c < CW; ++c) {
            y[c] = 0.0;
#pragma unroll
            for (int h = 0; h < RH; ++h) y[c] = fma(vr[h], Bt[h][c], y[c]);
          }
          {
            const double tot = tau * warp_reduce_cols<CW>(y, lane);
            if ((lane & (32 / CW - 1)) == 0) s_y[warp][warp_col_of_lane<CW>(lane)] = tot;
            __syncwarp();
#pragma unroll
            for (int c = 0; c < CW; c += 2) {
              const double2 t2 = *reinterpret_cast<const double2*>(&s_y[warp][c]);
              y[c] = t2.x;
              y[c + 1] = t2.y;
            }
            __syncwarp();  // s_y[warp] is rewritten by this warp in the next task
          }
          if (warp == 0) y[0] = 0.0;  // column 0 already is beta e1
          double* lp = AB + (r0 - B + j0) * LDAB + (B - j0) + lane;  // element (row = lane, col = j0)
          if (nr == B) {  // full block: unpredicated stores (all but the one element of the look-ahead)
#pragma unroll
            for (int c = 0; c < CW; ++c)
#pragma unroll
              for (int h = 0; h < RH; ++h) {
                Bt[h][c] = fma(-vr[h], y[c], Bt[h][c]);
                if (c == 0 && h == 0) {
                  if (tid != 0) lp[0] = Bt[0][0];
                } else {
                  lp[c * cstep + 32 * h] = Bt[h][c];
                }
              }
          } else {
#pragma unroll
            for (int c = 0; c < CW; ++c)
#pragma unroll
              for (int h = 0; h < RH; ++h) {
                const int row = lane + 32 * h;
                Bt[h][c] = fma(-vr[h], y[c], Bt[h][c]);
                // (0,0) = beta went to the band with the look-ahead; a later sweep may already have overwritten it
                if (row < nr && !(warp == 0 && c == 0 && row == 0)) lp[c * cstep + 32 * h] = Bt[h][c];
              }
          }
        }
        if (traced) stamp(1, tslot);
        // ---- wait for sweep s-1: tasks 0 .. t+1 complete
        if (s > 0) {
          if (lane == 0) wait_progress(prog + (s - 1), t + 2);
          __syncwarp();
        }
        if (traced) stamp(2, tslot);
        // ---- D block (lower triangle of A(R,R)) and B block (A(R+B, R))
        double* dp = AB + (r0 + j0) * LDAB - j0 + lane;  // D element (row = lane, col = j0); B element: B rows below
        if (nr == B && nr2 == B) {
          // interior task (all but the last two of a sweep): 2 RH CW unconditional loads issued back to back -- ONE
          // L2 round trip; the part of D above the diagonal reads the tail of the neighbouring band column (valid
          // memory) and is zeroed afterwards
#pragma unroll
          for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int h = 0; h < RH; ++h) {
              Dt[h][c] = ldcg_f64(dp + c * cstep + 32 * h);
              Bt[h][c] = ldcg_f64(dp + c * cstep + 32 * h + B);
            }
#pragma unroll
          for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int h = 0; h < RH; ++h)
              if (lane + 32 * h < j0 + c) Dt[h][c] = 0.0;
        } else {
#pragma unroll
          for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int h = 0; h < RH; ++h) {
              const int row = lane + 32 * h, col = j0 + c;
              Dt[h][c] = (col < nr && row >= col && row < nr) ? __ldcg(dp + c * cstep + 32 * h) : 0.0;
              Bt[h][c] = (col < nr && row < nr2) ? __ldcg(dp + c * cstep + 32 * h + B) : 0.0;
            }
        }
        // ---- partial dots: u = tau B v (rows), p = tau D v (rows from the lower triangle, columns from its transpose)
        {
          double pu[RH], pp[RH], pcq[CW];
#pragma unroll
          for (int h = 0; h < RH; ++h) {
            pu[h] = 0.0;
            pp[h] = 0.0;
#pragma unroll
            for (int c = 0; c < CW; ++c) {
              pu[h] = fma(Bt[h][c], vc[c], pu[h]);
              pp[h] = fma(Dt[h][c], vc[c], pp[h]);
            }
          }
#pragma unroll
          for (int c = 0; c < CW; ++c) {
            pcq[c] = 0.0;
#pragma unroll
            for (int h = 0; h < RH; ++h)
              if (lane + 32 * h != j0 + c) pcq[c] = fma(Dt[h][c], vr[h], pcq[c]);  // strictly lower part, transposed
          }
          if constexpr (TRACE) {
            if (pu[0] + pp[0] + pcq[0] == 1.2345e300) trace[0] = 0;  // consume the loads before the stamp
            if (traced) stamp(3, tslot);
          }
#pragma unroll
          for (int h = 0; h < RH; ++h) {
            s_pu[warp][lane + 32 * h] = pu[h];
            s_pp[warp][lane + 32 * h] = pp[h];
            if (RW && warp == 0) s_x[lane + 32 * h] = Bt[h][0];
          }
          const double ctot = warp_reduce_cols<CW>(pcq, lane);
          if ((lane & (32 / CW - 1)) == 0) s_pc[j0 + warp_col_of_lane<CW>(lane)] = ctot;
        }
        if (traced) stamp(4, tslot);
        __syncthreads();
        if (traced) stamp(5, tslot);
        // ---- every warp finishes u, p, sigma, w for its rows and columns (same operations in the same order: the
        // redundant copies are bit-identical)
        double u[RH], wv[RH], wc[CW];
        {
          double pv[RH];
          double sig = 0.0;
#pragma unroll
          for (int h = 0; h < RH; ++h) {
            const int row = lane + 32 * h;
            u[h] = tau * tree_sum<NW>(&s_pu[0][row], B);
            pv[h] = tau * (tree_sum<NW>(&s_pp[0][row], B) + s_pc[row]);
            sig = fma(pv[h], vr[h], sig);
          }
          sig = warp_allreduce_sum(sig);
#pragma unroll
          for (int h = 0; h < RH; ++h) wv[h] = fma(-0.5 * tau * sig, vr[h], pv[h]);
#pragma unroll
          for (int c = 0; c < CW; ++c) {
            const int col = j0 + c;
            wc[c] = shi((RH == 2 && col >= 32) ? wv[RH - 1] : wv[0], col & 31);
          }
        }
        // ---- D -= v w^T + w v^T (lower part, stored), B -= u v^T (kept: it is the L block of task t+1)
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
          for (int h = 0; h < RH; ++h) {
            const int row = lane + 32 * h, col = j0 + c;
            double dv = fma(-vr[h], wc[c], Dt[h][c]);
            dv = fma(-wv[h], vc[c], dv);
            Dt[h][c] = dv;
            if (col == 0 && row < nr) dp[32 * h] = dv;  // column 0 now, the rest after the completion count
            Bt[h][c] = fma(-u[h], vc[c], Bt[h][c]);
          }
        dprev = dp;
        nrprev = nr;
        if (traced) stamp(6, tslot);
        if (last) {
#pragma unroll
          for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int h = 0; h < RH; ++h) {
              const int row = lane + 32 * h, col = j0 + c;
              if (col < nr && row < nr2) dp[c * cstep + 32 * h + B] = Bt[h][c];
            }
        }
      } else {
        __syncthreads();  // the reflector warp has nothing to do before the partial dots are in shared memory
      }
      if (!last && warp == RWARP) {
        // ---- look-ahead: reflector of task t+1 from column 0 of the new B block; its beta goes to the band now
        double x[RH], vn[RH], taun, betan;
        if (RW) {
          const double v0 = s_v[cur][0];
#pragma unroll
          for (int h = 0; h < RH; ++h) {
            const int row = lane + 32 * h;
            const double a = tree_sum<NW>(&s_pu[0][row], B);
            x[h] = fma(-(tau * a), v0, s_x[row]);  // the same expression the compute warps apply to column 0
          }
        } else {
#pragma unroll
          for (int h = 0; h < RH; ++h) x[h] = Bt[h][0];
        }
        warp_reflector<RH>(x, nr2, lane, vn, taun, betan);
#pragma unroll
        for (int h = 0; h < RH; ++h) {
          const int row = lane + 32 * h;
          s_v[cur ^ 1][row] = vn[h];
          if (row < nr2) V2[s * ldv + r0 + B + row] = vn[h];
          if (!RW) Bt[h][0] = (row == 0) ? betan : 0.0;
        }
        if (lane == 0) {
          s_tau[cur ^ 1] = taun;
          s_beta[cur ^ 1] = betan;
          TAU2[s * (i64)ldtau + t + 1] = taun;
          AB[r0 * LDAB + B] = betan;
        }
        stamp(7, 0);
      }
      if constexpr (TRACE) ++trace_task;
    }
    if (compute) store_deferred_D();
    __syncthreads();
    if (tid == 0) st_release_gpu(prog + s, DONE);
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
template <int B, int NW, bool RW, bool TRACE = false>
static int sb2st_reg_launch(Ctx* ctx, i64 n, double* AB, i64 ldab, double* V2, i64 ldv, double* TAU2, int ldtau, int* prog,
                            long long* trace = nullptr) {
  auto kern = sb2st_reg_kernel<B, NW, RW, TRACE>;
  constexpr int threads = 32 * (NW + (RW ? 1 : 0));
  int per_sm = 0;
  EKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
  if (ctx->sb2st_cps > 0 && per_sm > ctx->sb2st_cps) per_sm = ctx->sb2st_cps;
  i64 G = (i64)per_sm * ctx->num_sms;
  if (G < 1) return EKB_ERR_INTERNAL;
  // a sweep follows its predecessor at a distance of 2 tasks: more than ntask(0)/2 + 2 CTAs would only wait
  const i64 maxuse = (n / B) / 2 + 2;
  if (G > maxuse) G = maxuse;
  if (G > n - 2) G = n - 2;
  void* args[] = {(void*)&AB, (void*)&ldab, (void*)&n, (void*)&V2, (void*)&ldv, (void*)&TAU2, (void*)&ldtau, (void*)&prog,
                  (void*)&trace};
  // all CTAs must be co-resident (spin-wait dependencies): the cooperative launch enforces it
  EKB_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3((unsigned)G), dim3(threads), args, 0, ctx->stream));
  EKB_COUNT_LAUNCH(ctx);
  return 0;
}

template <bool TRACE>
static int sb2st_reg_dispatch(Ctx* ctx, i64 n, int b, double* AB, i64 ldab, double* V2, i64 ldv, double* TAU2, int ldtau,
                              int* prog, long long* trace) {
  const bool wide = ctx->sb2st_warps == 16, rw = ctx->sb2st_rwarp != 0;
  if (b == 64) {
    if (wide) return rw ? sb2st_reg_launch<64, 16, true, TRACE>(ctx, n, AB, ldab, V2, ldv, TAU2, ldtau, prog, trace)
                        : sb2st_reg_launch<64, 16, false, TRACE>(ctx, n, AB, ldab, V2, ldv, TAU2, ldtau, prog, trace);
    return rw ? sb2st_reg_launch<64, 8, true, TRACE>(ctx, n, AB, ldab, V2, ldv, TAU2, ldtau, prog, trace)
              : sb2st_reg_launch<64, 8, false, TRACE>(ctx, n, AB, ldab, V2, ldv, TAU2, ldtau, prog, trace);
  }
  return rw ? sb2st_reg_launch<32, 8, true, TRACE>(ctx, n, AB, ldab, V2, ldv, TAU2, ldtau, prog, trace)
            : sb2st_reg_launch<32, 8, false, TRACE>(ctx, n, AB, ldab, V2, ldv, TAU2, ldtau, prog, trace);
}

// AB: ldab >= 2b rows (rows b+1.. must be zero on entry).  V2: n x (n-2 or more), ldv; TAU2: ldtau x (n-2).
// d (n), e (n-1) receive the tridiagonal.  prog: int workspace of n entries.
int sb2st(Ctx* ctx, i64 n, int b, double* AB, i64 ldab, double* V2, i64 ldv, double* TAU2, int ldtau, int* prog,
          double* d, double* e) {
  if (n <= 0) return 0;
  if (n >= 3 && ctx->sb2st_variant != 0 && ldab == 2 * (i64)b) {
    EKB_CUDA(cudaMemsetAsync(prog, 0, (size_t)n * sizeof(int), ctx->stream));
    EKB_TRY(prof_begin(ctx, PROF_SB2ST, 12.0 * b * (double)n * (double)n));  // effective bytes, SURVEY 8(d)
    int rc;
    const char* trace_path = getenv("EKB200_SB2ST_TRACE");
    if (trace_path && *trace_path) {  // development aid: phase timestamps of one CTA -> binary file
      const size_t cnt = (size_t)SB2ST_TRACE_TASKS * 2 * SB2ST_TRACE_PTS;
      long long* dtr = nullptr;
      EKB_TRY(ctx_alloc(ctx, (void**)&dtr, cnt * sizeof(long long)));
      EKB_CUDA(cudaMemsetAsync(dtr, 0, cnt * sizeof(long long), ctx->stream));
      rc = sb2st_reg_dispatch<true>(ctx, n, b, AB, ldab, V2, ldv, TAU2, ldtau, prog, dtr);
      std::vector<long long> htr(cnt);
      if (!rc && cudaMemcpyAsync(htr.data(), dtr, cnt * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
          cudaStreamSynchronize(ctx->stream) == cudaSuccess) {
        if (FILE* f = fopen(trace_path, "wb")) {
          fwrite(htr.data(), sizeof(long long), cnt, f);
          fclose(f);
        }
      }
      ctx_free(ctx, dtr);
    } else {
      rc = sb2st_reg_dispatch<false>(ctx, n, b, AB, ldab, V2, ldv, TAU2, ldtau, prog, nullptr);
    }
    if (rc) return rc;
    EKB_TRY(prof_end(ctx));
  } else if (n >= 3) {
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
