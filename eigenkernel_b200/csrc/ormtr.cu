// Back-transformation of the eigenvectors through both stages of the two-stage tridiagonalization:
// replaces pdormtr('L','L','N'), reference src/solver_scalapack_all.f90:115-116 (Z <- Q Z with the
// reflectors of pdsytrd).  Here Q = Q1 * Q2:
//   Q2 (band -> tridiagonal, sb2st.cu): n^2/(2b) short reflectors H(s,t) of length <= b.
//   Q1 (dense -> band, sy2sb.cu):       n/b compact-WY panels (V_p, T_p).
//
// apply_q2: reflectors of NBS consecutive sweeps at the same chase step t form a parallelogram
//   ("diamond") block G(S,t) = H(s0,t) ... H(s0+NBS-1,t) = I - V T V^T with V of (b+NBS-1) x NBS.
//   Valid application order (reflectors (s,t), (s',t') with s < s' only conflict when t' <= t):
//   sweep blocks S descending, chase steps t ascending.  Columns of Z are independent, so ONE kernel
//   applies the whole of Q2: a CTA owns KC columns of Z, keeps the moving (b+NBS-1)-row window of its slab
//   in a shared-memory ring and walks all diamond blocks, three DMMA products per block
//   (W = V^T Zw, W = T W, Zw -= V W) with structural zeros of V and T skipped at k4 granularity.
//   The packed (V,T) images are produced once by q2_pack_kernel in exactly the shared-memory layout.
// apply_q1: G panels are aggregated into one (G*b)-wide WY block (T by the block recurrence
//   T[0:c,c] = -T[0:c,0:c] (V_{<c}^T V_c) T_c), then Z -= V (T (V^T Z)) as three engine GEMMs per group,
//   groups descending.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace ekb {

// ------------------------------------------------------------------------------------------ shared helpers
__device__ __forceinline__ void dmma884_(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__host__ __device__ __forceinline__ int q2_num_tasks(i64 n, int b, i64 s) {
  if (s > n - 3) return 0;
  return (int)((n - 3 - s) / b) + 1;
}

// Compile-time geometry of the diamond blocks.
template <int B, int NBS>
struct Q2Geom {
  static constexpr int H = B + NBS - 1;            // rows of a diamond block
  static constexpr int HP = (H + 3) / 4 * 4;       // padded to the MMA k granularity
  static constexpr int LDV = HP + 4;               // smem stride of a V column (== 4 mod 16: conflict free)
  static constexpr int LDT = NBS + 4;              // smem stride of a T row
  static constexpr int BLK = NBS * LDV + NBS * LDT;  // doubles per packed block
  static constexpr int RM = 128;                   // ring modulus of the Z window (>= HP, power of two)
  static constexpr int LDZ = RM + 4;
  static_assert(HP <= RM, "window must fit the ring");
  static_assert(LDV % 16 == 4 && LDT % 16 == 4 && LDZ % 16 == 4, "bank-conflict-free strides");
};

// ------------------------------------------------------------------------------------------ q2 pack
// One CTA per diamond block (t = blockIdx.x, S = blockIdx.y): gathers V from V2 with the structural zeros
// made explicit, forms the NBS x NBS compact-WY factor T (forward, columnwise) and writes both in the
// shared-memory image consumed by q2_apply_kernel.
template <int B, int NBS>
__global__ void __launch_bounds__(128) q2_pack_kernel(const double* __restrict__ V2, i64 ldv,
                                                      const double* __restrict__ TAU2, int ldtau, i64 n,
                                                      const i64* __restrict__ blk_off, double* __restrict__ packed) {
  using G = Q2Geom<B, NBS>;
  __shared__ double Vs[NBS * G::LDV];
  __shared__ double Gs[NBS][NBS + 1];
  __shared__ double Ts[NBS][NBS + 1];
  __shared__ double taus[NBS];
  const int t = blockIdx.x, S = blockIdx.y, tid = threadIdx.x;
  const i64 s0 = (i64)S * NBS;
  if (t >= q2_num_tasks(n, B, s0)) return;
  const i64 R0 = s0 + 1 + (i64)t * B;
  for (int idx = tid; idx < NBS * G::LDV; idx += blockDim.x) {
    const int i = idx / G::LDV, r = idx % G::LDV;
    const i64 s = s0 + i;
    double v = 0.0;
    if (r < G::H && s <= n - 3) {
      const i64 r0 = R0 + i;  // first row of reflector (s, t)
      const i64 nr = min((i64)B, n - r0);
      if (nr >= 2 && r >= i && r < i + nr) v = V2[s * ldv + R0 + r];
    }
    Vs[idx] = v;
  }
  if (tid < NBS) {
    const i64 s = s0 + tid;
    double tv = 0.0;
    if (s <= n - 3) {
      const i64 nr = min((i64)B, n - (R0 + tid));
      if (nr >= 2) tv = TAU2[s * ldtau + t];
    }
    taus[tid] = tv;
  }
  __syncthreads();
  // Gram matrix (strict upper part is all that is needed)
  for (int idx = tid; idx < NBS * NBS; idx += blockDim.x) {
    const int i = idx / NBS, j = idx % NBS;
    double g = 0.0;
    if (i < j) {
      const double* vi = Vs + i * G::LDV;
      const double* vj = Vs + j * G::LDV;
      for (int r = j; r < min(G::H, i + B); ++r) g += vi[r] * vj[r];
    }
    Gs[i][j] = g;
    Ts[i][j] = 0.0;
  }
  __syncthreads();
  for (int i = 0; i < NBS; ++i) {
    const double ti = taus[i];
    double v = 0.0;
    if (tid < i) {
      for (int q = tid; q < i; ++q) v += Ts[tid][q] * Gs[q][i];
      v *= -ti;
    } else if (tid == i) v = ti;
    __syncthreads();
    if (tid <= i) Ts[tid][i] = v;
    __syncthreads();
  }
  double* out = packed + (blk_off[S] + t) * (i64)G::BLK;
  for (int idx = tid; idx < NBS * G::LDV; idx += blockDim.x) out[idx] = Vs[idx];
  double* outT = out + NBS * G::LDV;
  for (int idx = tid; idx < NBS * G::LDT; idx += blockDim.x) {
    const int i = idx / G::LDT, j = idx % G::LDT;
    outT[idx] = (j < NBS) ? Ts[i][j] : 0.0;
  }
}

// ------------------------------------------------------------------------------------------ q2 apply
constexpr int Q2_THREADS = 256;

template <int B, int NBS, int KC>
__global__ void __launch_bounds__(Q2_THREADS) q2_apply_kernel(const double* __restrict__ packed,
                                                             const i64* __restrict__ blk_off, i64 n, int nS,
                                                             double* __restrict__ Z, i64 ldz, i64 k) {
  using G = Q2Geom<B, NBS>;
  constexpr int LDW = NBS + 4;
  extern __shared__ __align__(16) double sm[];
  double* Zs = sm;                       // KC x LDZ   ring of Z rows: Zs[c*LDZ + (row & (RM-1))]
  double* Vs = Zs + KC * G::LDZ;         // NBS x LDV  Vs[i*LDV + r]
  double* Ts = Vs + NBS * G::LDV;        // NBS x LDT  Ts[i*LDT + j]
  double* Ws = Ts + NBS * G::LDT;        // KC x LDW   Ws[c*LDW + i]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lq = lane >> 2, lr = lane & 3;
  const i64 c0 = (i64)blockIdx.x * KC;
  const int ncol = (int)min((i64)KC, k - c0);
  double* Zg = Z + c0 * ldz;

  // warp tiling: 2 (M) x 4 (N) warps
  const int wm = warp & 1, wn = warp >> 1;
  constexpr int WN = KC / 4;             // columns per warp
  constexpr int NI = WN / 8;
  static_assert(KC % 32 == 0, "KC must be a multiple of 32");
  constexpr int M1 = NBS / 2, MI1 = M1 / 8;        // product 1/2: rows of W per warp
  constexpr int M3 = G::HP / 2;                     // product 3: rows of the window per warp
  static_assert(G::HP % 16 == 0, "HP must split over two warps in 8-row tiles");
  constexpr int MI3 = M3 / 8;

  for (int idx = tid; idx < KC * G::LDZ; idx += Q2_THREADS) Zs[idx] = 0.0;
  __syncthreads();

  i64 lo = 0, hi = 0;  // rows [lo, hi) of the slab are resident (and dirty) in the ring

  auto evict = [&](i64 upto) {  // write rows [lo, upto) back
    const int rows = (int)(upto - lo);
    if (rows > 0) {
      for (int idx = tid; idx < rows * ncol; idx += Q2_THREADS) {
        const int r = idx % rows, c = idx / rows;
        const i64 row = lo + r;
        if (row < n) Zg[(i64)c * ldz + row] = Zs[c * G::LDZ + (int)(row & (G::RM - 1))];
      }
      lo = upto;
    }
  };
  auto fetch = [&](i64 upto) {  // bring rows [hi, upto) in (zero beyond n)
    const int rows = (int)(upto - hi);
    if (rows > 0) {
      for (int idx = tid; idx < rows * KC; idx += Q2_THREADS) {
        const int r = idx % rows, c = idx / rows;
        const i64 row = hi + r;
        double v = 0.0;
        if (row < n && c < ncol) v = Zg[(i64)c * ldz + row];
        Zs[c * G::LDZ + (int)(row & (G::RM - 1))] = v;
      }
      hi = upto;
    }
  };

  for (int S = nS - 1; S >= 0; --S) {
    const i64 s0 = (i64)S * NBS;
    const int ntask = q2_num_tasks(n, B, s0);
    const double* pk = packed + blk_off[S] * (i64)G::BLK;
    for (int t = 0; t < ntask; ++t, pk += G::BLK) {
      const i64 R0 = s0 + 1 + (i64)t * B;
      // ---- stage operands: packed (V,T) image and the Z window [R0, R0+HP)
      {
        const double2* src = reinterpret_cast<const double2*>(pk);
        double2* dst = reinterpret_cast<double2*>(Vs);
        for (int idx = tid; idx < G::BLK / 2; idx += Q2_THREADS) dst[idx] = __ldg(src + idx);
      }
      if (t == 0) {
        evict(hi);       // flush the previous sweep block's window
        lo = hi = R0;
      } else {
        evict(R0);
      }
      __syncthreads();  // ring slots of evicted rows are reused by the rows fetched next
      fetch(R0 + G::HP);
      __syncthreads();

      // ---- product 1: W(i,c) = sum_r V(r,i) Zw(r,c);  rows i of this warp: [wm*M1, wm*M1+M1)
      {
        double acc[MI1][NI][2];
#pragma unroll
        for (int a = 0; a < MI1; ++a)
#pragma unroll
          for (int j = 0; j < NI; ++j) acc[a][j][0] = acc[a][j][1] = 0.0;
        const int i0 = wm * M1;
        const int kbeg = i0 & ~3, kend = min(G::HP, i0 + M1 + B);  // V(r,i) != 0 only for i <= r < i+B
        const double* va = Vs + (i0 + lq) * G::LDV + lr;
        const double* zb = Zs + (wn * WN + lq) * G::LDZ;
        for (int kk = kbeg; kk < kend; kk += 4) {
          double af[MI1], bf[NI];
          const int zr = (int)((R0 + kk + lr) & (G::RM - 1));
#pragma unroll
          for (int a = 0; a < MI1; ++a) af[a] = va[a * 8 * G::LDV + kk];
#pragma unroll
          for (int j = 0; j < NI; ++j) bf[j] = zb[j * 8 * G::LDZ + zr];
#pragma unroll
          for (int a = 0; a < MI1; ++a)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884_(acc[a][j][0], acc[a][j][1], bf[j], af[a]);
        }
#pragma unroll
        for (int a = 0; a < MI1; ++a)
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            const int i = i0 + a * 8 + 2 * lr, c = wn * WN + j * 8 + lq;
            *reinterpret_cast<double2*>(Ws + c * LDW + i) = make_double2(acc[a][j][0], acc[a][j][1]);
          }
      }
      __syncthreads();
      // ---- product 2: W <- T W (T upper triangular)
      {
        double acc[MI1][NI][2];
#pragma unroll
        for (int a = 0; a < MI1; ++a)
#pragma unroll
          for (int j = 0; j < NI; ++j) acc[a][j][0] = acc[a][j][1] = 0.0;
        const int i0 = wm * M1;
        const double* ta = Ts + (i0 + lq) * G::LDT + lr;
        const double* wb = Ws + (wn * WN + lq) * LDW + lr;
        for (int kk = i0; kk < NBS; kk += 4) {
          double af[MI1], bf[NI];
#pragma unroll
          for (int a = 0; a < MI1; ++a) af[a] = ta[a * 8 * G::LDT + kk];
#pragma unroll
          for (int j = 0; j < NI; ++j) bf[j] = wb[j * 8 * LDW + kk];
#pragma unroll
          for (int a = 0; a < MI1; ++a)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884_(acc[a][j][0], acc[a][j][1], bf[j], af[a]);
        }
        __syncthreads();  // every warp has finished reading W
#pragma unroll
        for (int a = 0; a < MI1; ++a)
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            const int i = i0 + a * 8 + 2 * lr, c = wn * WN + j * 8 + lq;
            *reinterpret_cast<double2*>(Ws + c * LDW + i) = make_double2(acc[a][j][0], acc[a][j][1]);
          }
      }
      __syncthreads();
      // ---- product 3: Zw(r,c) -= sum_i V(r,i) W(i,c);  rows r of this warp: [wm*M3, wm*M3+M3)
      {
        double acc[MI3][NI][2];
#pragma unroll
        for (int a = 0; a < MI3; ++a)
#pragma unroll
          for (int j = 0; j < NI; ++j) acc[a][j][0] = acc[a][j][1] = 0.0;
        const int r0w = wm * M3;
        const double* va = Vs + lr * G::LDV + r0w + lq;
        const double* wb = Ws + (wn * WN + lq) * LDW + lr;
#pragma unroll
        for (int kk = 0; kk < NBS; kk += 4) {
          double bf[NI];
#pragma unroll
          for (int j = 0; j < NI; ++j) bf[j] = wb[j * 8 * LDW + kk];
#pragma unroll
          for (int a = 0; a < MI3; ++a) {
            // rows [ra, ra+8) of the window see reflectors i with r-B < i <= r only
            const int ra = r0w + a * 8;
            if (kk <= ra + 7 && kk + 3 > ra - B) {
              const double af = va[(i64)kk * G::LDV + a * 8];
#pragma unroll
              for (int j = 0; j < NI; ++j) dmma884_(acc[a][j][0], acc[a][j][1], bf[j], af);
            }
          }
        }
#pragma unroll
        for (int a = 0; a < MI3; ++a)
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            const int r = r0w + a * 8 + 2 * lr, c = wn * WN + j * 8 + lq;
            double* zc = Zs + c * G::LDZ;
            const int q0 = (int)((R0 + r) & (G::RM - 1)), q1 = (int)((R0 + r + 1) & (G::RM - 1));
            zc[q0] -= acc[a][j][0];
            zc[q1] -= acc[a][j][1];
          }
      }
      __syncthreads();
    }
  }
  evict(min(hi, n));
}

template <int B, int NBS, int KC>
static int q2_launch(Ctx* ctx, i64 n, const double* V2, i64 ldv, const double* TAU2, int ldtau, i64 k, double* Z,
                     i64 ldz) {
  using G = Q2Geom<B, NBS>;
  const i64 nsw = n - 2;  // sweeps 0 .. n-3
  if (nsw <= 0 || k <= 0) return 0;
  const int nS = (int)((nsw + NBS - 1) / NBS);
  std::vector<i64> off(nS + 1, 0);
  for (int S = 0; S < nS; ++S) off[S + 1] = off[S] + q2_num_tasks(n, B, (i64)S * NBS);
  const i64 nblk = off[nS];
  i64* d_off = nullptr;
  double* packed = nullptr;
  EKB_TRY(ctx_alloc(ctx, (void**)&d_off, (size_t)(nS + 1) * sizeof(i64)));
  int rc = ctx_alloc(ctx, (void**)&packed, (size_t)nblk * G::BLK * sizeof(double));
  if (rc) { ctx_free(ctx, d_off); return rc; }
  auto cleanup = [&]() { cudaStreamSynchronize(ctx->stream); ctx_free(ctx, d_off); ctx_free(ctx, packed); };
  cudaError_t ce = cudaMemcpyAsync(d_off, off.data(), (size_t)(nS + 1) * sizeof(i64), cudaMemcpyHostToDevice, ctx->stream);
  if (ce == cudaSuccess) {
    const int tmax = q2_num_tasks(n, B, 0);
    q2_pack_kernel<B, NBS><<<dim3(tmax, nS), 128, 0, ctx->stream>>>(V2, ldv, TAU2, ldtau, n, d_off, packed); EKB_COUNT_LAUNCH(ctx);
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess) {
    constexpr size_t smem = (size_t)(KC * G::LDZ + NBS * G::LDV + NBS * G::LDT + KC * (NBS + 4)) * sizeof(double);
    auto kern = q2_apply_kernel<B, NBS, KC>;
    ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce == cudaSuccess) {
      kern<<<cdiv(k, KC), Q2_THREADS, smem, ctx->stream>>>(packed, d_off, n, nS, Z, ldz, k); EKB_COUNT_LAUNCH(ctx);
      ce = cudaGetLastError();
    }
  }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  cleanup();
  if (ce != cudaSuccess) {
    ctx->last_cuda = ce;
    ctx->last_error = std::string("apply_q2: ") + cudaGetErrorString(ce);
    return EKB_ERR_CUDA;
  }
  return 0;
}

// Z (n x k) <- Q2 Z with the reflectors produced by sb2st (layout documented there).
int apply_q2(Ctx* ctx, i64 n, int b, const double* V2, i64 ldv, const double* TAU2, int ldtau, i64 k, double* Z,
             i64 ldz) {
  if (b == 64) return q2_launch<64, 32, 64>(ctx, n, V2, ldv, TAU2, ldtau, k, Z, ldz);
  if (b == 32) return q2_launch<32, 32, 64>(ctx, n, V2, ldv, TAU2, ldtau, k, Z, ldz);
  return EKB_ERR_INTERNAL;
}

// ------------------------------------------------------------------------------------------ q1
constexpr int Q1_GROUP = 4;  // panels aggregated into one WY block

// Explicit zeros above every panel inside its group: rows [j0+b, j+b) of panel columns [j, j+b).
__global__ void q1_zero_above_kernel(double* __restrict__ A, i64 lda, int b, int npan, int group) {
  const int p = blockIdx.x;
  if (p >= npan) return;
  const int i = p % group;
  if (i == 0) return;
  const i64 j = (i64)p * b, j0 = (i64)(p - i) * b;
  const int rows = i * b;
  for (int idx = threadIdx.x; idx < rows * b; idx += blockDim.x) {
    const int r = idx % rows, c = idx / rows;
    A[(j + c) * lda + j0 + b + r] = 0.0;
  }
}

// Tb (W x W per group, zeroed) gets the per-panel T factors on its diagonal blocks.
__global__ void q1_init_tb_kernel(const double* __restrict__ T1, int b, int npan, int group, int W,
                                  double* __restrict__ Tb) {
  const int p = blockIdx.x;
  if (p >= npan) return;
  const int g = p / group, i = p % group;
  double* dst = Tb + (size_t)g * W * W + (size_t)(i * b) * W + i * b;
  const double* src = T1 + (size_t)p * b * b;
  for (int idx = threadIdx.x; idx < b * b; idx += blockDim.x) {
    const int r = idx % b, c = idx / b;
    dst[(size_t)c * W + r] = src[idx];
  }
}

size_t apply_q1_workspace_doubles(i64 n, int b, i64 k) {
  const int npan = sy2sb_num_panels(n, b);
  const int ng = (npan + Q1_GROUP - 1) / Q1_GROUP;
  const size_t W = (size_t)Q1_GROUP * b;
  return (size_t)ng * W * W * 2 + (size_t)ng * W * b + 2 * W * (size_t)round_up(k, 8) + 1024;
}

// Z (n x k) <- Q1 Z.  A holds the explicit V panels below the band (modified: zeros made explicit),
// T1 the per-panel compact-WY factors.
int apply_q1(Ctx* ctx, i64 n, int b, double* A, i64 lda, const double* T1, i64 k, double* Z, i64 ldz, double* work) {
  const int npan = sy2sb_num_panels(n, b);
  if (npan <= 0 || k <= 0) return 0;
  const int Gp = Q1_GROUP;
  const int ng = (npan + Gp - 1) / Gp;
  const int W = Gp * b;
  double* S = work;                              // ng x W x W
  double* Tb = S + (size_t)ng * W * W;           // ng x W x W
  double* X = Tb + (size_t)ng * W * W;           // ng x W x b
  double* Wk = X + (size_t)ng * W * b;           // W x k
  const i64 ldw = W;
  double* Wk2 = Wk + (size_t)W * round_up(k, 8);

  q1_zero_above_kernel<<<npan, 256, 0, ctx->stream>>>(A, lda, b, npan, Gp); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  EKB_CUDA(cudaMemsetAsync(Tb, 0, (size_t)ng * W * W * sizeof(double), ctx->stream));
  q1_init_tb_kernel<<<npan, 256, 0, ctx->stream>>>(T1, b, npan, Gp, W, Tb); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());

  if (Gp > 1) {
    // batched descriptors: [gram | (X_c, Tb_c) for c = 1..Gp-1]
    std::vector<GemmP> hp((size_t)ng * (1 + 2 * (Gp - 1)));
    for (int g = 0; g < ng; ++g) {
      const int p0 = g * Gp, wg = std::min(Gp, npan - p0);
      const i64 j0 = (i64)p0 * b, r0 = j0 + b, m = n - r0;
      const double* V = A + j0 * lda + r0;
      GemmP q;
      q.m = wg * b; q.n = wg * b; q.k = (int)m;
      q.A = V; q.lda = lda; q.B = V; q.ldb = lda; q.C = S + (size_t)g * W * W; q.ldc = W;
      q.alpha = 1.0; q.beta = 0.0;
      hp[g] = q;
      for (int c = 1; c < Gp; ++c) {
        GemmP x = {}, tb = {};
        if (c < wg) {
          x.m = c * b; x.n = b; x.k = b;
          x.A = S + (size_t)g * W * W + (size_t)(c * b) * W; x.lda = W;
          x.B = T1 + (size_t)(p0 + c) * b * b; x.ldb = b;
          x.C = X + (size_t)g * W * b; x.ldc = W;
          x.alpha = 1.0; x.beta = 0.0;
          tb.m = c * b; tb.n = b; tb.k = c * b;
          tb.A = Tb + (size_t)g * W * W; tb.lda = W;
          tb.B = X + (size_t)g * W * b; tb.ldb = W;
          tb.C = Tb + (size_t)g * W * W + (size_t)(c * b) * W; tb.ldc = W;
          tb.alpha = -1.0; tb.beta = 0.0;
        }
        hp[(size_t)ng * (1 + 2 * (c - 1)) + g] = x;
        hp[(size_t)ng * (2 + 2 * (c - 1)) + g] = tb;
      }
    }
    GemmP* d_gp = nullptr;
    EKB_TRY(ctx_alloc(ctx, (void**)&d_gp, hp.size() * sizeof(GemmP)));
    cudaError_t ce = cudaMemcpyAsync(d_gp, hp.data(), hp.size() * sizeof(GemmP), cudaMemcpyHostToDevice, ctx->stream);
    int rc = 0;
    if (ce != cudaSuccess) rc = EKB_ERR_CUDA;
    if (!rc) rc = gemm_batched(ctx, GEMM_TA, d_gp, ng, W, W);
    for (int c = 1; c < Gp && !rc; ++c) {
      rc = gemm_batched(ctx, 0, d_gp + (size_t)ng * (1 + 2 * (c - 1)), ng, c * b, b);
      if (!rc) rc = gemm_batched(ctx, 0, d_gp + (size_t)ng * (2 + 2 * (c - 1)), ng, c * b, b);
    }
    cudaStreamSynchronize(ctx->stream);  // hp must outlive the copy; d_gp the kernels
    ctx_free(ctx, d_gp);
    if (rc) return rc;
  }

  for (int g = ng - 1; g >= 0; --g) {
    const int p0 = g * Gp, wg = std::min(Gp, npan - p0);
    const i64 j0 = (i64)p0 * b, r0 = j0 + b, m = n - r0;
    const int w = wg * b;
    const double* V = A + j0 * lda + r0;
    double* Zs = Z + r0;
    GemmP q;
    // Wk = V^T Zs
    q.m = w; q.n = (int)k; q.k = (int)m; q.A = V; q.lda = lda; q.B = Zs; q.ldb = ldz; q.C = Wk; q.ldc = ldw;
    q.alpha = 1.0; q.beta = 0.0;
    EKB_TRY(gemm(ctx, GEMM_TA, q));
    // Wk2 = Tb Wk
    q.m = w; q.n = (int)k; q.k = w; q.A = Tb + (size_t)g * W * W; q.lda = W; q.B = Wk; q.ldb = ldw; q.C = Wk2; q.ldc = ldw;
    EKB_TRY(gemm(ctx, 0, q));
    // Zs -= V Wk2
    q.m = (int)m; q.n = (int)k; q.k = w; q.A = V; q.lda = lda; q.B = Wk2; q.ldb = ldw; q.C = Zs; q.ldc = ldz;
    q.alpha = -1.0; q.beta = 1.0;
    EKB_TRY(gemm(ctx, 0, q));
  }
  return 0;
}

}  // namespace ekb
