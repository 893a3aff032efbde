// Stage 1 of the two-stage tridiagonalization: dense symmetric -> symmetric band (half bandwidth b).
// Replaces (together with sb2st.cu) pdsytrd('L') of the reference, src/solver_scalapack_all.f90:59; the
// two-stage organisation follows the ELPA2 / eigen_sx workflow the reference offers as
// general_elpa_eigensx (src/solver_elpa_eigenexa.f90:25-198).
//
// Per panel p (columns j..j+b-1, rows j+b..n-1, m = n-j-b rows):
//   1. Householder QR of the m x b panel by ONE cooperative kernel (one grid barrier per column; the
//      update of column c and the dot products the next reflector needs are fused in a single pass).
//   2. R and the diagonal block go to band storage AB; V is made explicit in place (unit diagonal, zeros).
//   3. T (compact WY) from the Gram matrix V^T V.
//   4. X = A22 V T (SYMM on the DMMA engine reading the lower triangle only), W = X - 1/2 V T^T V^T X.
//   5. A22 -= [V W][W V]^T  (SYR2K as one K=2b GEMM, lower tiles + one tile diagonal).
// V stays in A (below the band), T per panel in T1: both are consumed by the back-transformation
// (ormtr.cu: apply_q1).
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ekb {

constexpr int QR_THREADS = 256;

// Householder QR of P (m x b, column-major, ld), b <= 64.  Grid of G CTAs, each owning a contiguous row
// range.  partial: 2 * G * 64 doubles.  Rout: b x b (ld b) receives R (upper triangle); tau: b.
// In place, rows below the diagonal of column c hold v_c (unit element implicit); the upper triangle of
// the top b x b block is left stale (fixed by fixup_extract_kernel).
// Grid barrier for a NON-cooperative launch (bar != nullptr): a monotonically growing arrival counter in global memory.
// The look-ahead runs this kernel on a side stream while the trailing update fills the chip; a cooperative grid is only
// started once ALL its CTAs fit at the same time, i.e. after the update has drained (measured: no overlap at all),
// whereas ordinary CTAs of a higher-priority stream take the SMs one by one as they fall free and wait here for the
// rest.  Safe because nothing the running update does depends on this kernel, so every CTA does get an SM.
__device__ __forceinline__ void sw_grid_barrier(unsigned* bar, unsigned nctas, unsigned& phase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++phase;
    __threadfence();
    atomicAdd(bar, 1u);
    const unsigned target = phase * nctas;
    while (*reinterpret_cast<volatile unsigned*>(bar) < target) {}
    __threadfence();
  }
  __syncthreads();
}

// Gate on the main stream: holds the trailing update back until every CTA of the side-stream panel QR is resident
// (bar[2] >= seq), so that the QR owns its few SMs BEFORE the update's thousands of CTAs flood the chip.  (Stream
// priorities alone did not do it: measured, the QR's CTAs were only placed after the update's grid had drained.)  The
// wait is bounded: if the QR cannot start, the update simply goes first and the QR runs after it, as without look-ahead.
__global__ void qr_resident_gate_kernel(const unsigned* __restrict__ bar, unsigned seq) {
  const long long t0 = clock64();
  while (*reinterpret_cast<const volatile unsigned*>(bar + 2) < seq && clock64() - t0 < (1ll << 22)) {}
}

template <int B>
__global__ void __launch_bounds__(QR_THREADS) panel_qr_kernel(double* __restrict__ P, i64 ld, int m,
                                                              double* __restrict__ tau, double* __restrict__ Rout,
                                                              double* __restrict__ partial, unsigned* __restrict__ bar,
                                                              unsigned seq) {
  cg::grid_group grid = cg::this_grid();
  unsigned phase = 0;
  if (bar && threadIdx.x == 0) {  // bar[1]: CTAs of this launch that have started; the last one opens the gate (bar[2])
    if (atomicAdd(bar + 1, 1u) == gridDim.x - 1) {
      *reinterpret_cast<volatile unsigned*>(bar + 2) = seq;
      __threadfence();
    }
  }
  __shared__ double red[QR_THREADS / 32][B];
  __shared__ double g[B];
  __shared__ double wv[B];
  __shared__ double s_scal, s_tau;
  const int G = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rpc = (m + G - 1) / G;
  const int r0 = blockIdx.x * rpc, r1 = min(m, r0 + rpc);
  const int kr = min(B, m);

  for (int j = -1; j < kr; ++j) {
    double scal = 0.0, tj = 0.0;
    if (j >= 0) {
      // reduce the partial dot products of column j with columns j..B-1 (rows > j)
      // all threads take part (QR_THREADS / B interleaved slices of the G partials per column, loads batched):
      // a single thread per column walking G dependent L2 loads used to dominate the time of a column step
      const double* pp = partial + (size_t)(j & 1) * G * B;
      {
        constexpr int NPART = QR_THREADS / B;
        const int c = tid % B, part = tid / B;
        double s = 0.0;
        if (c >= j) {
#pragma unroll 8
          for (int q = part; q < G; q += NPART) s += __ldcg(pp + q * B + c);
        }
        red[part][c] = s;
      }
      __syncthreads();
      if (tid < B) {
        constexpr int NPART = QR_THREADS / B;
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < NPART; ++q) s += red[q][tid];
        g[tid] = s;
      }
      __syncthreads();
      if (tid == 0) {
        double alpha = P[(i64)j * ld + j];
        double xn2 = g[j];
        double beta, t, sc;
        if (xn2 == 0.0) {
          beta = alpha; t = 0.0; sc = 0.0;
        } else {
          beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
          t = (beta - alpha) / beta;
          sc = 1.0 / (alpha - beta);
        }
        s_scal = sc; s_tau = t;
        if (blockIdx.x == 0) { tau[j] = t; Rout[j * B + j] = beta; }
      }
      __syncthreads();
      scal = s_scal; tj = s_tau;
      if (tid < B && tid > j) {
        // w_c = v^T P(:,c) = P(j,c) + scal * g_c ; R(j,c) = P(j,c) - tau * w_c
        double pjc = P[(i64)tid * ld + j];
        double w = pjc + scal * g[tid];
        wv[tid] = tj * w;
        if (blockIdx.x == 0) Rout[tid * B + j] = pjc - tj * w;
      }
      __syncthreads();
    }
    // pass over my rows: apply reflector j (if any), then accumulate dots of column j+1 with columns >= j+1
    double acc[B];
#pragma unroll
    for (int c = 0; c < B; ++c) acc[c] = 0.0;
    const int jn = j + 1;  // next pivot column
    if (jn < B) {
      for (int i = r0 + tid; i < r1; i += QR_THREADS) {
        double vi = 0.0;
        const bool upd = (j >= 0) && (i > j);
        if (upd) {
          vi = P[(i64)j * ld + i] * scal;
          P[(i64)j * ld + i] = vi;
        }
        double p1 = 0.0;
#pragma unroll
        for (int c = 0; c < B; ++c) {
          if (c >= jn) {
            double x = P[(i64)c * ld + i];
            if (upd) {
              x -= wv[c] * vi;
              P[(i64)c * ld + i] = x;
            }
            if (c == jn) p1 = x;
            if (i > jn) acc[c] += p1 * x;
          }
        }
      }
    } else if (j >= 0) {
      // last column: only scale it
      for (int i = r0 + tid; i < r1; i += QR_THREADS)
        if (i > j) P[(i64)j * ld + i] *= scal;
    }
    if (jn < kr) {
#pragma unroll
      for (int c = 0; c < B; ++c) {
        double v = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][c] = v;
      }
      __syncthreads();
      if (tid < B) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < QR_THREADS / 32; ++w) s += red[w][tid];
        partial[(size_t)(jn & 1) * G * B + blockIdx.x * B + tid] = s;
      }
    }
    if (bar) sw_grid_barrier(bar, (unsigned)G, phase);
    else grid.sync();
  }
  // reflectors that do not exist (m < B): tau = 0, R rows beyond m are zero
  if (blockIdx.x == 0 && tid < B && tid >= kr) tau[tid] = 0.0;
}

// ---- the same factorization with the panel RESIDENT IN SHARED MEMORY (round 2).
// panel_qr_kernel walks its rows of the panel in global memory once per column: 64 passes whose loads are L2 round
// trips (18 us per column at m = 32000: 1.15 ms per panel, 0.59 s of the n = 32768 solve, all of it with 148 - G SMs
// idle).  Here every CTA copies its <= 352 rows (176 KB) into shared memory once, does all 64 update + dot passes there
// (thread = (column, row slice): no shuffles, conflict-free columns with an odd leading dimension), and only the
// per-column exchange crosses CTAs: the partial dots of the next pivot column and that pivot's ROW (whose owner
// publishes it, since the global copy of the panel is stale until the final write-back), one software grid barrier
// per column.  Cooperative launch for co-residency; same outputs, same operation order per element as the kernel above
// except for the order of the partial sums.
constexpr int QRS_MAXROWS = 352;
constexpr int QRS_MAXG = 160;  // CTAs of the shared-memory panel QR (<= SM count; the launcher falls back above it)
template <int B>
__global__ void __launch_bounds__(QR_THREADS, 1) panel_qr_smem_kernel(double* __restrict__ P, i64 ld, int m,
                                                                      double* __restrict__ tau, double* __restrict__ Rout,
                                                                      double* __restrict__ partial,
                                                                      double* __restrict__ rowbuf /* 2 x B */,
                                                                      unsigned* __restrict__ bar) {
  extern __shared__ double ps[];  // ps[c * ldp + i]: column c, local row i
  constexpr int NQ = QR_THREADS / B;
  __shared__ double red[NQ][B];
  __shared__ double srow[B];
  const int G = gridDim.x, tid = threadIdx.x;
  const int rpc = (m + G - 1) / G;
  const int r0 = blockIdx.x * rpc, r1 = min(m, r0 + rpc);
  const int rows = max(0, r1 - r0);
  const int ldp = rpc | 1;
  const int kr = min(B, m);
  const int c = tid % B, q = tid / B;
  const int rs = (rows + NQ - 1) / NQ;
  const int i0 = min(rows, q * rs), i1 = min(rows, i0 + rs);
  unsigned phase = 0;

  for (int cc = 0; cc < B; ++cc)
    for (int i = tid; i < rows; i += QR_THREADS) ps[cc * ldp + i] = P[(i64)cc * ld + r0 + i];
  __syncthreads();

  for (int j = -1; j < kr; ++j) {
    double scal = 0.0, tj = 0.0, wc_own = 0.0;
    if (j >= 0) {
      // partial dots of column j with columns j..B-1 (rows > j), summed over the CTAs in a fixed order
      const double* pp = partial + (size_t)(j & 1) * G * B;
      const double* prow = rowbuf + (j & 1) * B;  // row j of the panel, published by its owner before the barrier
      {
        // the pivot row travels with the partial dots (same round trip), then lives in shared memory
        double rowv = 0.0;
        if (tid < B) rowv = __ldcg(prow + tid);
        // every load of this thread's share in flight at once (one L2 round trip instead of one per 8 CTAs; the sum
        // keeps its order): G <= QRS_MAXG CTAs, NQ slices
        double sacc = 0.0;
        if (c >= j) {
          constexpr int NT = (QRS_MAXG + NQ - 1) / NQ;
          double vals[NT];
#pragma unroll
          for (int u = 0; u < NT; ++u) {
            const int t = q + u * NQ;
            vals[u] = t < G ? __ldcg(pp + t * B + c) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < NT; ++u) sacc += vals[u];
        }
        red[q][c] = sacc;
        if (tid < B) srow[tid] = rowv;
      }
      __syncthreads();
      // Every thread derives the reflector's scalars and the w of ITS column from the block-level sums (same values in
      // every thread: no elected thread, no broadcast round; three block barriers per column instead of six)
      auto colsum = [&](int col) {
        double sacc2 = 0.0;
#pragma unroll
        for (int t = 0; t < NQ; ++t) sacc2 += red[t][col];
        return sacc2;
      };
      const double alpha = srow[j];
      const double xn2 = colsum(j);
      double beta;
      if (xn2 == 0.0) {
        beta = alpha; tj = 0.0; scal = 0.0;
      } else {
        beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
        tj = (beta - alpha) / beta;
        scal = 1.0 / (alpha - beta);
      }
      const double pjc = srow[c];
      const double w = pjc + scal * colsum(c);
      wc_own = tj * w;
      const int jn1 = j + 1;
      const double wn = jn1 < B ? tj * (srow[jn1] + scal * colsum(jn1)) : 0.0;
      if (blockIdx.x == 0 && q == 0) {
        if (c == j) { tau[j] = tj; Rout[j * B + j] = beta; }
        else if (c > j) Rout[c * B + j] = pjc - tj * w;
      }
      // v_j: scale column j below the pivot (the reflector, stored in place), and apply it to the next pivot column
      // at once (every thread of the main pass reads that column's UPDATED values)
      for (int i = tid; i < rows; i += QR_THREADS)
        if (r0 + i > j) {
          const double v = ps[j * ldp + i] * scal;
          ps[j * ldp + i] = v;
          if (jn1 < B) ps[jn1 * ldp + i] -= wn * v;
        }
      __syncthreads();
    }
    const int jn = j + 1;  // next pivot column
    if (jn < B) {
      // reflector j on the other columns, fused with the dots of column jn against columns >= jn
      double acc = 0.0;
      if (c >= jn) {
        const double wc = (j >= 0 && c > jn) ? wc_own : 0.0;
        const double* vj = ps + (j >= 0 ? j : 0) * ldp;
        const double* pn = ps + jn * ldp;
        double* pc = ps + c * ldp;
        for (int i = i0; i < i1; ++i) {
          const int gi = r0 + i;
          double x = pc[i];
          if (c > jn && j >= 0 && gi > j) {
            x -= wc * vj[i];
            pc[i] = x;
          }
          if (gi > jn) acc = fma(pn[i], x, acc);
        }
      }
      red[q][c] = acc;
      __syncthreads();
      if (jn < kr && tid < B) {
        double sacc = 0.0;
#pragma unroll
        for (int t = 0; t < NQ; ++t) sacc += red[t][tid];
        partial[(size_t)(jn & 1) * G * B + blockIdx.x * B + tid] = sacc;
        if (jn >= r0 && jn < r1) rowbuf[(jn & 1) * B + tid] = ps[tid * ldp + (jn - r0)];  // the owner publishes row jn
      }
    }
    sw_grid_barrier(bar, (unsigned)G, phase);
  }
  for (int cc = 0; cc < B; ++cc)
    for (int i = tid; i < rows; i += QR_THREADS) P[(i64)cc * ld + r0 + i] = ps[cc * ldp + i];
  if (blockIdx.x == 0 && tid < B && tid >= kr) tau[tid] = 0.0;
}

// One CTA per panel: write AB columns j..j+b-1 (diagonal block from A, R from Rout) and make V explicit.
__global__ void fixup_extract_kernel(double* __restrict__ A, i64 lda, i64 n, i64 j, int b, const double* __restrict__ Rout,
                                     double* __restrict__ AB, i64 ldab) {
  const int m = (int)(n - j - b);
  for (int e = threadIdx.x; e < (int)ldab * b; e += blockDim.x) {
    int d = e % (int)ldab, c = e / (int)ldab;
    i64 jc = j + c, i = jc + d;
    double v = 0.0;
    if (i < n && d <= b) {
      if (i < j + b) v = A[jc * lda + i];
      else {
        int rr = (int)(i - (j + b));
        if (rr <= c && rr < m) v = Rout[c * b + rr];
      }
    }
    AB[jc * ldab + d] = v;
  }
  double* P = A + j * lda + (j + b);
  for (int e = threadIdx.x; e < b * b; e += blockDim.x) {
    int rr = e % b, c = e / b;
    if (rr < m && rr <= c) P[(i64)c * lda + rr] = (rr == c) ? 1.0 : 0.0;
  }
}

// Tail columns (no panel below them): AB(d, jc) = A(jc + d, jc) for d <= b.
__global__ void extract_tail_kernel(const double* __restrict__ A, i64 lda, i64 n, i64 j0, int b, double* __restrict__ AB,
                                    i64 ldab) {
  i64 jc = j0 + blockIdx.x;
  if (jc >= n) return;
  for (int d = threadIdx.x; d < ldab; d += blockDim.x) {
    i64 i = jc + d;
    AB[jc * ldab + d] = (i < n && d <= b) ? A[jc * lda + i] : 0.0;
  }
}

// T = larft(forward, columnwise) from G = V^T V and tau; also usable for any b <= 64.  One CTA of 64 threads.
__global__ void build_T_kernel(const double* __restrict__ Gm, const double* __restrict__ tau, int b, double* __restrict__ T) {
  __shared__ double sT[64][65];
  const int r = threadIdx.x;
  for (int c = 0; c < b; ++c) {
    if (r < b) sT[r][c] = 0.0;
  }
  __syncthreads();
  for (int i = 0; i < b; ++i) {
    const double ti = tau[i];
    double v = 0.0;
    if (r < i) {
      for (int k = r; k < i; ++k) v += sT[r][k] * Gm[i * b + k];
      v *= -ti;
    } else if (r == i) v = ti;
    __syncthreads();
    if (r <= i && r < b) sT[r][i] = v;
    __syncthreads();
  }
  for (int c = 0; c < b; ++c)
    if (r < b) T[c * b + r] = sT[r][c];
}

// TS = T^T * S (b x b), one CTA.
__global__ void tts_kernel(const double* __restrict__ T, const double* __restrict__ S, int b, double* __restrict__ TS) {
  for (int e = threadIdx.x; e < b * b; e += blockDim.x) {
    int r = e % b, c = e / b;
    double v = 0.0;
    for (int k = 0; k <= r; ++k) v += T[r * b + k] * S[c * b + k];  // T^T(r,k) = T(k,r), k <= r
    TS[e] = v;
  }
}

// VW = [V | W], WV = [W | V]  (m x 2b, ld = ldp)
__global__ void pack_vw_kernel(const double* __restrict__ V, i64 ldv, const double* __restrict__ W, i64 ldw, int m, int b,
                               double* __restrict__ VW, double* __restrict__ WV, i64 ldp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int c = blockIdx.y;
  if (i >= m) return;
  double v = V[(i64)c * ldv + i], w = W[(i64)c * ldw + i];
  VW[(i64)c * ldp + i] = v;
  VW[(i64)(c + b) * ldp + i] = w;
  WV[(i64)c * ldp + i] = w;
  WV[(i64)(c + b) * ldp + i] = v;
}

// The shared-memory-resident QR: G = ceil(m / 256) CTAs (<= one per SM), cooperative launch, software grid barrier on
// bar[0] (zeroed here).  Returns -1 when the panel does not fit (m > 352 rows per SM): the caller then uses the
// global-memory kernel.
static int launch_panel_qr_smem(Ctx* ctx, int b, double* P, i64 ld, int m, double* tau, double* Rout, double* partial,
                                double* rowbuf, unsigned* bar) {
  int G = (m + QR_THREADS - 1) / QR_THREADS;
  if (G > ctx->num_sms) G = ctx->num_sms;
  const int rpc = (m + G - 1) / G;
  if (rpc > QRS_MAXROWS || G > QRS_MAXG) return -1;
  const size_t smem = (size_t)(rpc | 1) * b * sizeof(double);
  void* args[] = {(void*)&P, (void*)&ld, (void*)&m, (void*)&tau, (void*)&Rout, (void*)&partial, (void*)&rowbuf, (void*)&bar};
  void* kern = b == 64 ? (void*)panel_qr_smem_kernel<64> : (void*)panel_qr_smem_kernel<32>;
  static bool attr_dev[64] = {};
  if (!attr_dev[ctx->device & 63]) {
    EKB_CUDA(cudaFuncSetAttribute(panel_qr_smem_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (QRS_MAXROWS | 1) * 64 * (int)sizeof(double)));
    EKB_CUDA(cudaFuncSetAttribute(panel_qr_smem_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (QRS_MAXROWS | 1) * 32 * (int)sizeof(double)));
    attr_dev[ctx->device & 63] = true;
  }
  EKB_TRY(prof_begin(ctx, PROF_PANEL_QR, 2.0 * m * (double)b * b));
  EKB_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), ctx->stream));
  EKB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(G), dim3(QR_THREADS), args, smem, ctx->stream));
  EKB_COUNT_LAUNCH(ctx);
  return prof_end(ctx);
}

// bar == nullptr: cooperative launch (grid.sync); otherwise an ordinary launch with the software barrier on *bar.
static int launch_panel_qr(Ctx* ctx, int b, double* P, i64 ld, int m, double* tau, double* Rout, double* partial, int G,
                           unsigned* bar = nullptr, unsigned seq = 0) {
  void* args[] = {(void*)&P, (void*)&ld, (void*)&m, (void*)&tau, (void*)&Rout, (void*)&partial, (void*)&bar, (void*)&seq};
  EKB_TRY(prof_begin(ctx, PROF_PANEL_QR, 2.0 * m * (double)b * b));
  if (bar) {
    EKB_CUDA(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned), ctx->stream));  // arrivals, started (bar[2], the gate, only grows)
    if (b == 64) panel_qr_kernel<64><<<G, QR_THREADS, 0, ctx->stream>>>(P, ld, m, tau, Rout, partial, bar, seq);
    else panel_qr_kernel<32><<<G, QR_THREADS, 0, ctx->stream>>>(P, ld, m, tau, Rout, partial, bar, seq);
    EKB_CUDA(cudaGetLastError());
  } else if (b == 64) {
    EKB_CUDA(cudaLaunchCooperativeKernel((void*)panel_qr_kernel<64>, dim3(G), dim3(QR_THREADS), args, 0, ctx->stream));
  } else {
    EKB_CUDA(cudaLaunchCooperativeKernel((void*)panel_qr_kernel<32>, dim3(G), dim3(QR_THREADS), args, 0, ctx->stream));
  }
  EKB_COUNT_LAUNCH(ctx);
  return prof_end(ctx);
}

static int pick_splitk(Ctx* ctx, i64 m, i64 n, i64 k, int bm, int bn) {
  i64 tiles = (i64)cdiv(m, bm) * cdiv(n, bn);
  i64 want = (2 * ctx->num_sms + tiles - 1) / tiles;
  i64 maxs = k / 256;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  return (int)want;
}

// Workspace layout (doubles): see sy2sb_workspace_doubles.
size_t sy2sb_workspace_doubles(i64 n, int b, int num_sms) {
  i64 ldp = round_up(n, 8);
  return (size_t)ldp * b * 2      /* W0/X, (spare) */
         + (size_t)ldp * 2 * b * 2 /* VW, WV */
         + (size_t)2 * num_sms * 64 + 8 * 64 * 64 + 8 /* grid-barrier words */ + 128 /* pivot rows of the shared-memory QR */;
}

// Work issued inside this scope goes to the context's side stream (and records its profile events there).
namespace {
struct OnAuxStream {
  Ctx* ctx;
  cudaStream_t saved;
  // the split-K workspace is per stream: products on the two streams may run at the same time
  explicit OnAuxStream(Ctx* c) : ctx(c), saved(c->stream) { c->stream = c->aux_stream; swap_ws(); }
  ~OnAuxStream() { ctx->stream = saved; swap_ws(); }
  void swap_ws() {
    std::swap(ctx->splitk_ws, ctx->splitk_ws_aux);
    std::swap(ctx->splitk_ws_bytes, ctx->splitk_ws_aux_bytes);
  }
};
}  // namespace

// Panel look-ahead (option "sy2sb_lookahead", OFF by default).  Panel p+1 only needs ITS b columns of the trailing
// matrix updated, so per panel
//   main stream:  W_p (SYMM etc.) -> skinny update of the next panel's columns -> [event] -> rank-2b update of the rest
//   side stream:                                             [wait] QR, band extraction, T of panel p+1 -> [event]
// and the main stream waits for that event before the SYMM of panel p+1.  MEASURED (n = 32768, round 2,
// profiles/r02_bench_n32768_lookahead_{on,off}.json): the reduction gets SLOWER, 2.74 s -> 4.23 s.  First as a
// cooperative launch (a cooperative grid is only placed when all its CTAs fit at once, i.e. after the update has
// drained), then as an ordinary launch with a software grid barrier on a high-priority stream, then with a gate kernel
// that holds the update back until every QR CTA is resident: always ~6 ms per panel QR instead of 1.1 ms.  The QR as
// written walks the panel in global memory and is bound by L2 round trips; beside an update that moves 2-3 TB/s those
// round trips take several times longer, and the chain becomes the critical path.  Look-ahead needs a QR that does not
// depend on memory latency -- which is what panel_qr_smem_kernel below is; with it the stand-alone QR is short enough
// that the overlap is no longer worth its price.  The code path is kept for that experiment.
int sy2sb(Ctx* ctx, i64 n, int b, double* A, i64 lda, double* AB, i64 ldab, double* T1, double* work) {
  if (n <= 0) return 0;
  const i64 ldp = round_up(n, 8);
  double* X = work;                  // m x b
  double* VW = X + ldp * b * 2;      // m x 2b
  double* WV = VW + ldp * 2 * b;     // m x 2b
  double* partial = WV + ldp * 2 * b;
  double* small = partial + 2 * ctx->num_sms * 64;
  double* tau = small;               // 64
  double* Rout = small + 64;         // b*b
  double* Gm = Rout + 64 * 64;       // b*b
  double* S = Gm + 64 * 64;          // b*b
  double* TS = S + 64 * 64;          // b*b
  unsigned* qr_bar = reinterpret_cast<unsigned*>(small + 8 * 64 * 64);
  double* rowbuf = small + 8 * 64 * 64 + 8;

  bool lookahead = ctx->sy2sb_lookahead != 0 && n - b >= 4 * b;
  if (lookahead && ctx_ensure_aux(ctx) != 0) {
    cudaGetLastError();
    lookahead = false;
  }
  if (lookahead) EKB_CUDA(cudaMemsetAsync(qr_bar, 0, 4 * sizeof(unsigned), ctx->stream));
  // QR + band extraction + T of the panel at column j (m rows below the band), on the CURRENT ctx->stream
  auto factor_panel = [&](i64 j, int p, bool side) -> int {
    const i64 m = n - j - b;
    double* P = A + j * lda + (j + b);
    double* T = T1 + (size_t)p * b * b;
    int G = (int)((m + QR_THREADS - 1) / QR_THREADS);
    if (G > ctx->num_sms) G = ctx->num_sms;
    if (side) {  // beside the trailing update: a few CTAs (8 rows per thread and column step), never more than 1/8 of the chip
      int gs = (int)((m + 8 * QR_THREADS - 1) / (8 * QR_THREADS));
      gs = std::max(gs, 2);
      gs = std::min(gs, std::max(ctx->num_sms / 8, 1));
      G = std::min(G, gs);
    }
    int qrc = -1;
    if (!side && ctx->panel_qr_variant != 0)
      qrc = launch_panel_qr_smem(ctx, b, P, lda, (int)m, tau, Rout, partial, rowbuf, qr_bar + 3);
    if (qrc > 0) return qrc;
    if (qrc < 0) EKB_TRY(launch_panel_qr(ctx, b, P, lda, (int)m, tau, Rout, partial, G, side ? qr_bar : nullptr, (unsigned)(p + 1)));
    fixup_extract_kernel<<<1, 256, 0, ctx->stream>>>(A, lda, n, j, b, Rout, AB, ldab); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
    GemmP g;
    g.m = b; g.n = b; g.k = (int)m; g.A = P; g.lda = lda; g.B = P; g.ldb = lda; g.C = Gm; g.ldc = b;
    g.alpha = 1.0; g.beta = 0.0;
    EKB_TRY(gemm(ctx, GEMM_TA, g, -1, pick_splitk(ctx, b, b, m, 128, 64)));
    build_T_kernel<<<1, 64, 0, ctx->stream>>>(Gm, tau, b, T); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
    return 0;
  };

  i64 j = 0;
  int p = 0;
  bool factored = false;  // panel p already factored (on the side stream; ctx->aux_ev[0] marks its completion)
  for (;; j += b, ++p) {
    const i64 m = n - j - b;
    if (m < 2) break;
    double* P = A + j * lda + (j + b);
    double* A22 = A + (j + b) * lda + (j + b);
    double* T = T1 + (size_t)p * b * b;
    // 1.-3. panel QR, band extraction + explicit V, T
    if (factored) EKB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->aux_ev[0], 0));
    else EKB_TRY(factor_panel(j, p, false));
    factored = false;
    // 4. W0 = A22 V (symmetric, lower stored) -> VW[:, b:2b) as scratch ; X = W0 T
    GemmP g;
    double* W0 = WV;  // scratch m x b
    g.alpha = 1.0; g.beta = 0.0;
    g.m = (int)m; g.n = b; g.k = (int)m; g.A = A22; g.lda = lda; g.B = P; g.ldb = lda; g.C = W0; g.ldc = ldp;
    EKB_TRY(gemm(ctx, GEMM_SYMA, g, -1, pick_splitk(ctx, m, b, m, 128, 64)));
    g.m = (int)m; g.n = b; g.k = b; g.A = W0; g.lda = ldp; g.B = T; g.ldb = b; g.C = X; g.ldc = ldp;
    EKB_TRY(gemm(ctx, 0, g));
    // S = V^T X ; TS = T^T S ; X -= 1/2 V TS   (X becomes W)
    g.m = b; g.n = b; g.k = (int)m; g.A = P; g.lda = lda; g.B = X; g.ldb = ldp; g.C = S; g.ldc = b;
    EKB_TRY(gemm(ctx, GEMM_TA, g, -1, pick_splitk(ctx, b, b, m, 128, 64)));
    tts_kernel<<<1, 256, 0, ctx->stream>>>(T, S, b, TS); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
    g.m = (int)m; g.n = b; g.k = b; g.A = P; g.lda = lda; g.B = TS; g.ldb = b; g.C = X; g.ldc = ldp;
    g.alpha = -0.5; g.beta = 1.0;
    EKB_TRY(gemm(ctx, 0, g));
    // 5. A22 -= [V W][W V]^T
    pack_vw_kernel<<<dim3(cdiv(m, 256), b), 256, 0, ctx->stream>>>(P, lda, X, ldp, (int)m, b, VW, WV, ldp); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
    g.alpha = -1.0; g.beta = 1.0;
    const i64 mnext = m - b;  // rows of the next panel
    if (lookahead && mnext >= 2 && m > 2 * b) {
      // 5a. the next panel's columns (and the diagonal block above them) first ...
      g.m = (int)m; g.n = b; g.k = 2 * b; g.A = VW; g.lda = ldp; g.B = WV; g.ldb = ldp; g.C = A22; g.ldc = lda;
      EKB_TRY(gemm(ctx, GEMM_TB, g));
      EKB_CUDA(cudaEventRecord(ctx->aux_ev[1], ctx->stream));
      {  // ... so that panel p+1 is factored on the side stream ...
        OnAuxStream aux(ctx);
        EKB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->aux_ev[1], 0));
        EKB_TRY(factor_panel(j + b, p + 1, true));
        EKB_CUDA(cudaEventRecord(ctx->aux_ev[0], ctx->stream));
      }
      factored = true;
      qr_resident_gate_kernel<<<1, 1, 0, ctx->stream>>>(qr_bar, (unsigned)(p + 2)); EKB_COUNT_LAUNCH(ctx);
      EKB_CUDA(cudaGetLastError());
      // 5b. ... while the rest of the trailing matrix (lower tiles + one tile diagonal of A22[b:, b:]) is updated here
      g.m = (int)(m - b); g.n = (int)(m - b); g.k = 2 * b; g.A = VW + b; g.lda = ldp; g.B = WV + b; g.ldb = ldp;
      g.C = A22 + (i64)b * lda + b; g.ldc = lda;
      EKB_TRY(gemm(ctx, GEMM_TB, g, /*tri_keep=*/128));
    } else {
      g.m = (int)m; g.n = (int)m; g.k = 2 * b; g.A = VW; g.lda = ldp; g.B = WV; g.ldb = ldp; g.C = A22; g.ldc = lda;
      EKB_TRY(gemm(ctx, GEMM_TB, g, /*tri_keep=*/128));
    }
  }
  if (factored) EKB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->aux_ev[0], 0));  // (cannot happen: the last panel has no successor)
  // tail columns
  if (j < n) {
    extract_tail_kernel<<<(unsigned)(n - j), 128, 0, ctx->stream>>>(A, lda, n, j, b, AB, ldab); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ sharded variant
// The trailing matrix is dealt to the P ranks by block columns of width b, cyclically (block c -> rank c mod P),
// stored compactly: Aloc is the local piece of a 1 x P block-cyclic distribution with NB = b -- the layout
// setup_distributed_matrix (reference src/distribute_matrix.f90:92-148) builds, here with one GPU per process
// column.  Owned columns are kept COMPLETE (both triangles), so per panel
//   owner: Householder QR of its local block column, T, band columns -> one ncclBroadcast [T | AB | V];
//   all:   partial W0 = A22(:, mine) V(mine, :)  (NN GEMM)          -> one ncclAllReduce of m x b;
//          W from W0 (replicated b-wide work), A22(:, mine) -= [V W] [W(mine) V(mine)]^T  (K = 2b GEMM).
// Every rank also keeps V_p in its full-size A (below the band) and T_p in T1: apply_q1 needs all panels.

// global column of local column lc on rank r
__device__ __host__ __forceinline__ i64 cyc_global_col(i64 lc, int b, int P, int r) {
  return ((lc / b) * P + r) * (i64)b + lc % b;
}

__global__ void cyc_pack_kernel(const double* __restrict__ A, i64 lda, i64 n, int b, int P, int r, double* __restrict__ Aloc,
                                i64 ldl, i64 nloc) {
  const i64 lc = blockIdx.y;
  if (lc >= nloc) return;
  const i64 g = cyc_global_col(lc, b, P, r);
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    Aloc[lc * ldl + i] = A[g * lda + i];
}

// out(i, 0:b) = S1(g(lc0 + i) - row_base, 0:b), out(i, b:2b) = S2(same row, 0:b) (S2 may be null), i < nl
__global__ void cyc_gather_rows_kernel(const double* __restrict__ S1, i64 ld1, const double* __restrict__ S2, i64 ld2, int b,
                                       int P, int r, i64 lc0, i64 nl, i64 row_base, double* __restrict__ out, i64 ldo) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (i >= nl) return;
  const i64 row = cyc_global_col(lc0 + i, b, P, r) - row_base;
  out[(i64)c * ldo + i] = S1[(i64)c * ld1 + row];
  if (S2) out[(i64)(c + b) * ldo + i] = S2[(i64)c * ld2 + row];
}

// VW = [V | W]  (m x 2b)
__global__ void pack_vw_only_kernel(const double* __restrict__ V, i64 ldv, const double* __restrict__ W, i64 ldw, int m, int b,
                                    double* __restrict__ VW, i64 ldp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int c = blockIdx.y;
  if (i >= m) return;
  VW[(i64)c * ldp + i] = V[(i64)c * ldv + i];
  VW[(i64)(c + b) * ldp + i] = W[(i64)c * ldw + i];
}

// tail(:, c) for the owned tail columns (global col j0 + c), zero elsewhere
__global__ void cyc_tail_kernel(const double* __restrict__ Aloc, i64 ldl, i64 n, int b, int P, int r, i64 j0, i64 ncol,
                                double* __restrict__ tail, i64 ldt) {
  const i64 c = blockIdx.y;
  if (c >= ncol) return;
  const i64 g = j0 + c;
  const bool mine = (int)((g / b) % P) == r;
  const i64 lc = (g / b / P) * b + g % b;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    tail[c * ldt + i] = mine ? Aloc[lc * ldl + i] : 0.0;
}

int sy2sb_dist(Ctx* ctx, i64 n, int b, double* A, i64 lda, double* AB, i64 ldab, double* T1) {
  if (n <= 0) return 0;
  const int P = ctx->nranks, r = ctx->rank;
  const i64 nblk = (n + b - 1) / b;
  const i64 myblk = nblk > r ? (nblk - r + P - 1) / P : 0;
  i64 nloc = myblk * b;
  if (myblk > 0 && (int)((nblk - 1) % P) == r) nloc -= nblk * b - n;  // ragged last block is mine
  const i64 ldl = round_up(n, 8), ldp = ldl;
  const i64 nlmax = nloc > 0 ? nloc : 1;
  // workspaces
  double *Aloc = nullptr, *X = nullptr, *W0 = nullptr, *VW = nullptr, *Vloc = nullptr, *WVloc = nullptr, *msg = nullptr,
         *small = nullptr, *tail = nullptr;
  std::vector<void*> owned;
  auto get = [&](double** p, size_t doubles) {
    int rc = ctx_alloc(ctx, (void**)p, doubles * sizeof(double));
    if (rc == 0) owned.push_back(*p);
    return rc;
  };
  auto cleanup = [&]() {
    cudaStreamSynchronize(ctx->stream);
    for (void* q : owned) ctx_free(ctx, q);
  };
  const size_t msg_doubles = (size_t)b * b + (size_t)ldab * b + (size_t)ldp * b;
  int rc = get(&Aloc, (size_t)ldl * nlmax);
  if (!rc) rc = get(&X, (size_t)ldp * b);
  if (!rc) rc = get(&W0, (size_t)ldp * b);
  if (!rc) rc = get(&VW, (size_t)ldp * 2 * b);
  if (!rc) rc = get(&Vloc, (size_t)round_up(nlmax, 8) * b);
  if (!rc) rc = get(&WVloc, (size_t)round_up(nlmax, 8) * 2 * b);
  if (!rc) rc = get(&msg, msg_doubles);
  if (!rc) rc = get(&small, (size_t)2 * ctx->num_sms * 64 + 8 * 64 * 64 + 8 + 128);
  if (!rc) rc = get(&tail, (size_t)ldp * (b + 2));
  if (rc) { cleanup(); return rc; }
  double* partial = small;
  double* tau = partial + 2 * ctx->num_sms * 64;
  double* Rout = tau + 64;
  double* Gm = Rout + 64 * 64;
  double* S = Gm + 64 * 64;
  double* TS = S + 64 * 64;
  unsigned* qr_bar = reinterpret_cast<unsigned*>(TS + 64 * 64);  // grid-barrier word of the shared-memory panel QR
  double* rowbuf = TS + 64 * 64 + 8;                             // its published pivot rows (2 x b)
  const i64 ldvl = round_up(nlmax, 8);

  auto body = [&]() -> int {
    if (nloc > 0) {
      cyc_pack_kernel<<<dim3(std::min<i64>(cdiv(n, 256), 64), (unsigned)nloc), 256, 0, ctx->stream>>>(A, lda, n, b, P, r, Aloc,
                                                                                                    ldl, nloc);
      EKB_COUNT_LAUNCH(ctx);
      EKB_CUDA(cudaGetLastError());
    }
    i64 j = 0;
    int p = 0;
    for (;; j += b, ++p) {
      const i64 m = n - j - b;
      if (m < 2) break;
      const int owner = p % P;
      const i64 mr = round_up(m, 8);
      double* msgT = msg;
      double* msgAB = msg + (size_t)b * b;
      double* msgV = msgAB + (size_t)ldab * b;
      if (r == owner) {
        const i64 lb = p / P;
        double* Pl = Aloc + (lb * b) * ldl + (j + b);
        int G = (int)((m + QR_THREADS - 1) / QR_THREADS);
        if (G > ctx->num_sms) G = ctx->num_sms;
        int qrc = ctx->panel_qr_variant != 0
                      ? launch_panel_qr_smem(ctx, b, Pl, ldl, (int)m, tau, Rout, partial, rowbuf, qr_bar) : -1;
        if (qrc > 0) return qrc;
        if (qrc < 0) EKB_TRY(launch_panel_qr(ctx, b, Pl, ldl, (int)m, tau, Rout, partial, G));
        // band columns into the message (as if AB started at column j), V made explicit in place
        fixup_extract_kernel<<<1, 256, 0, ctx->stream>>>(Aloc + (lb * b - j) * ldl, ldl, n, j, b, Rout, msgAB - j * ldab, ldab);
        EKB_COUNT_LAUNCH(ctx);
        EKB_CUDA(cudaGetLastError());
        GemmP g;
        g.m = b; g.n = b; g.k = (int)m; g.A = Pl; g.lda = ldl; g.B = Pl; g.ldb = ldl; g.C = Gm; g.ldc = b;
        g.alpha = 1.0; g.beta = 0.0;
        EKB_TRY(gemm(ctx, GEMM_TA, g, -1, pick_splitk(ctx, b, b, m, 128, 64)));
        build_T_kernel<<<1, 64, 0, ctx->stream>>>(Gm, tau, b, msgT); EKB_COUNT_LAUNCH(ctx);
        EKB_CUDA(cudaGetLastError());
        EKB_TRY(copy_matrix(ctx, Pl, ldl, msgV, mr, m, b));
      }
      EKB_TRY(comm_bcast(ctx, msg, ((size_t)b * b + (size_t)ldab * b + (size_t)mr * b) * sizeof(double), owner));
      double* Pv = A + j * lda + (j + b);  // V_p in the full-size array (apply_q1 reads it there)
      double* T = T1 + (size_t)p * b * b;
      EKB_TRY(copy_matrix(ctx, msgV, mr, Pv, lda, m, b));
      EKB_TRY(copy_matrix(ctx, msgT, b, T, b, b, b));
      EKB_TRY(copy_matrix(ctx, msgAB, ldab, AB + j * ldab, ldab, ldab, b));
      // local trailing columns
      const i64 lb0 = p >= r ? (p - r) / P + 1 : 0;
      const i64 lc0 = lb0 * b;
      const i64 nl = nloc > lc0 ? nloc - lc0 : 0;
      double* Atr = Aloc + lc0 * ldl + (j + b);
      GemmP g;
      if (nl > 0) {
        cyc_gather_rows_kernel<<<dim3(cdiv(nl, 256), b), 256, 0, ctx->stream>>>(Pv, lda, nullptr, 0, b, P, r, lc0, nl, j + b,
                                                                               Vloc, ldvl);
        EKB_COUNT_LAUNCH(ctx);
        EKB_CUDA(cudaGetLastError());
        g.m = (int)m; g.n = b; g.k = (int)nl; g.A = Atr; g.lda = ldl; g.B = Vloc; g.ldb = ldvl; g.C = W0; g.ldc = mr;
        g.alpha = 1.0; g.beta = 0.0;
        EKB_TRY(gemm(ctx, 0, g));
      } else {
        EKB_TRY(set_zero(ctx, W0, mr, mr, b));
      }
      EKB_TRY(comm_allreduce_sum(ctx, W0, (size_t)mr * b));
      // X = W0 T ; S = V^T X ; TS = T^T S ; X -= 1/2 V TS  (X becomes W) -- replicated, identical on all ranks
      g.m = (int)m; g.n = b; g.k = b; g.A = W0; g.lda = mr; g.B = T; g.ldb = b; g.C = X; g.ldc = ldp;
      g.alpha = 1.0; g.beta = 0.0;
      EKB_TRY(gemm(ctx, 0, g));
      g.m = b; g.n = b; g.k = (int)m; g.A = Pv; g.lda = lda; g.B = X; g.ldb = ldp; g.C = S; g.ldc = b;
      EKB_TRY(gemm(ctx, GEMM_TA, g, -1, pick_splitk(ctx, b, b, m, 128, 64)));
      tts_kernel<<<1, 256, 0, ctx->stream>>>(T, S, b, TS); EKB_COUNT_LAUNCH(ctx);
      EKB_CUDA(cudaGetLastError());
      g.m = (int)m; g.n = b; g.k = b; g.A = Pv; g.lda = lda; g.B = TS; g.ldb = b; g.C = X; g.ldc = ldp;
      g.alpha = -0.5; g.beta = 1.0;
      EKB_TRY(gemm(ctx, 0, g));
      if (nl > 0) {
        pack_vw_only_kernel<<<dim3(cdiv(m, 256), b), 256, 0, ctx->stream>>>(Pv, lda, X, ldp, (int)m, b, VW, ldp);
        EKB_COUNT_LAUNCH(ctx);
        cyc_gather_rows_kernel<<<dim3(cdiv(nl, 256), b), 256, 0, ctx->stream>>>(X, ldp, Pv, lda, b, P, r, lc0, nl, j + b, WVloc,
                                                                               ldvl);
        EKB_COUNT_LAUNCH(ctx);
        EKB_CUDA(cudaGetLastError());
        // A22(:, mine) -= [V W] [W(mine) V(mine)]^T
        g.m = (int)m; g.n = (int)nl; g.k = 2 * b; g.A = VW; g.lda = ldp; g.B = WVloc; g.ldb = ldvl; g.C = Atr; g.ldc = ldl;
        g.alpha = -1.0; g.beta = 1.0;
        EKB_TRY(gemm(ctx, GEMM_TB, g));
      }
    }
    // tail columns j .. n-1 (at most b + 1 of them, possibly owned by two ranks): sum of the owners' copies
    if (j < n) {
      const i64 ncol = n - j;
      cyc_tail_kernel<<<dim3(std::min<i64>(cdiv(n, 256), 64), (unsigned)ncol), 256, 0, ctx->stream>>>(Aloc, ldl, n, b, P, r, j, ncol,
                                                                                                    tail, ldp);
      EKB_COUNT_LAUNCH(ctx);
      EKB_CUDA(cudaGetLastError());
      EKB_TRY(comm_allreduce_sum(ctx, tail, (size_t)ldp * ncol));
      extract_tail_kernel<<<(unsigned)ncol, 128, 0, ctx->stream>>>(tail - j * ldp, ldp, n, j, b, AB, ldab);
      EKB_COUNT_LAUNCH(ctx);
      EKB_CUDA(cudaGetLastError());
    }
    return 0;
  };
  rc = body();
  cleanup();
  return rc;
}

int sy2sb_num_panels(i64 n, int b) {
  int p = 0;
  for (i64 j = 0; n - j - b >= 2; j += b) ++p;
  return p;
}

}  // namespace ekb
