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
  static constexpr int BLK = 2 * NBS * LDV;        // doubles per packed block: V image, then Y = V T^T image
  static_assert(LDV % 16 == 4, "bank-conflict-free strides");
};

// ------------------------------------------------------------------------------------------ q2 pack
// One CTA per diamond block (t = blockIdx.x, S = blockIdx.y): gathers V from V2 with the structural zeros
// made explicit, forms the NBS x NBS compact-WY factor T (forward, columnwise), folds it into Y = V T^T
// (so that G Z = Z - V (Y^T Z) needs two products, not three) and writes V and Y in the shared-memory
// image consumed by q2_apply_kernel.
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
  // Y(r,i) = sum_{j >= i} V(r,j) T(i,j)
  double* outY = out + NBS * G::LDV;
  for (int idx = tid; idx < NBS * G::LDV; idx += blockDim.x) {
    const int i = idx / G::LDV, r = idx % G::LDV;
    double y = 0.0;
    for (int j = i; j < NBS; ++j) y += Vs[j * G::LDV + r] * Ts[i][j];
    outY[idx] = y;
  }
}

// ------------------------------------------------------------------------------------------ q2 apply
__device__ __forceinline__ void q2_cp_async16(double* smem, const double* g) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(g) : "memory");
}
__device__ __forceinline__ void q2_cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void q2_cp_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// One CTA owns KC = 16 NWN columns of Z and walks every diamond block (S descending, t ascending); two CTAs
// share an SM so that one CTA's window traffic overlaps the other's tensor-pipe work.  Per block:
//     W = Y^T Zw   (NBS x KC, K = HP)      Zw -= V W   (HP x KC, K = NBS)
// both on DMMA.8x8x4 with operands read from bank-conflict-free shared-memory layouts; structural zeros of
// V and Y are skipped at k4 granularity.  The window of HP rows lives in an exact ring (row -> (row+1) mod
// HP, which keeps the accumulator row pairs 16-byte aligned).  Operand staging is asynchronous and phase
// shifted so that no load latency sits on the critical path: V(t) streams into shared memory (cp.async)
// while product 1 runs, Y(t+1) while product 2 runs, and the B rows of Z that enter the next window are
// prefetched into registers during both products.  Warp grid 2 (M) x NWN (N); warp tile N = 16.
template <int B, int NBS, int NWN>
__global__ void __launch_bounds__(64 * NWN, 2) q2_apply_kernel(const double* __restrict__ packed,
                                                              const i64* __restrict__ blk_off, i64 n, int nS,
                                                              double* __restrict__ Z, i64 ldz, i64 k) {
  using G = Q2Geom<B, NBS>;
  constexpr int THREADS = 64 * NWN;
  constexpr int KC = 16 * NWN;
  constexpr int HP = G::HP;
  constexpr int LDZ = HP + 4;
  constexpr int LDW = NBS + 4;
  constexpr int IMG = NBS * G::LDV;  // doubles of one operand image (V or Y)
  static_assert(LDZ % 16 == 4 && LDW % 16 == 4, "bank-conflict-free strides");
  static_assert(HP % 2 == 0 && IMG % 2 == 0, "16-byte granularity");
  extern __shared__ __align__(16) double sm[];
  double* Zs = sm;                       // KC x LDZ   ring of Z rows: Zs[c*LDZ + (row+1) mod HP]
  double* Vs = Zs + KC * LDZ;            // NBS x LDV  Vs[i*LDV + r]
  double* Ys = Vs + IMG;                 // NBS x LDV  Ys[i*LDV + r]
  double* Ws = Ys + IMG;                 // KC x LDW   Ws[c*LDW + i]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lq = lane >> 2, lr = lane & 3;
  const i64 c0 = (i64)blockIdx.x * KC;
  const int ncol = (int)min((i64)KC, k - c0);
  double* Zg = Z + c0 * ldz;

  const int wm = warp & 1, wn = warp >> 1;
  constexpr int NI = 2;                             // 16 columns per warp
  constexpr int M1 = NBS / 2, MI1 = M1 / 8;        // product 1: rows of W per warp
  constexpr int M3 = HP / 2, MI3 = M3 / 8;         // product 2: rows of the window per warp
  static_assert(HP % 16 == 0, "HP must split over two warps in 8-row tiles");
  constexpr int ZPF = (B * KC) / THREADS;          // doubles per thread of the B entering rows
  static_assert((B * KC) % THREADS == 0, "entering rows split evenly");

  for (int idx = tid; idx < KC * LDZ; idx += THREADS) Zs[idx] = 0.0;

  i64 lo = 0, hi = 0;  // rows [lo, hi) of the slab are resident (and dirty) in the ring

  auto ring = [&](i64 row) { return (int)((row + 1) % HP); };
  auto evict = [&](i64 upto) {  // write rows [lo, upto) back
    const int rows = (int)(upto - lo);
    if (rows > 0) {
      const int base = ring(lo);
      if (rows == B) {
#pragma unroll 4
        for (int idx = tid; idx < B * ncol; idx += THREADS) {
          const int r = idx % B, c = idx / B;
          int q = base + r;
          if (q >= HP) q -= HP;
          if (lo + r < n) Zg[(i64)c * ldz + lo + r] = Zs[c * LDZ + q];
        }
      } else {
        for (int idx = tid; idx < rows * ncol; idx += THREADS) {
          const int r = idx % rows, c = idx / rows;
          int q = base + r;
          if (q >= HP) q -= HP;
          if (lo + r < n) Zg[(i64)c * ldz + lo + r] = Zs[c * LDZ + q];
        }
      }
      lo = upto;
    }
  };
  // synchronous fetch of rows [hi, upto) (the HP - B extra rows of a fresh window)
  auto fetch = [&](i64 upto) {
    const int rows = (int)(upto - hi);
    if (rows > 0) {
      const int base = ring(hi);
      for (int idx = tid; idx < rows * KC; idx += THREADS) {
        const int r = idx % rows, c = idx / rows;
        int q = base + r;
        if (q >= HP) q -= HP;
        double v = 0.0;
        if (hi + r < n && c < ncol) v = __ldcg(Zg + (i64)c * ldz + hi + r);
        Zs[c * LDZ + q] = v;
      }
      hi = upto;
    }
  };

  double zreg[ZPF];
  i64 pf_row0 = 0;  // first of the B rows of Z held in zreg
  auto prefetch_z = [&](i64 row0) {
    pf_row0 = row0;
#pragma unroll
    for (int q = 0; q < ZPF; ++q) {
      const int idx = tid + q * THREADS;
      const int r = idx % B, c = idx / B;
      double v = 0.0;
      if (row0 + r < n && c < ncol) v = __ldcg(Zg + (i64)c * ldz + row0 + r);
      zreg[q] = v;
    }
  };
  auto commit_z = [&]() {
    const int base = ring(pf_row0);
#pragma unroll
    for (int q = 0; q < ZPF; ++q) {
      const int idx = tid + q * THREADS;
      const int r = idx % B, c = idx / B;
      int p = base + r;
      if (p >= HP) p -= HP;
      Zs[c * LDZ + p] = zreg[q];
    }
  };
  auto stream_image = [&](double* dst, const double* src) {  // IMG doubles, 16 bytes per cp.async
    for (int idx = tid; idx < IMG / 2; idx += THREADS) q2_cp_async16(dst + 2 * idx, src + 2 * idx);
    q2_cp_commit();
  };

  // first block of the walk
  int S = nS - 1;
  while (S >= 0 && q2_num_tasks(n, B, (i64)S * NBS) == 0) --S;
  if (S < 0) return;
  int t = 0;
  int ntask = q2_num_tasks(n, B, (i64)S * NBS);
  const double* pk = packed + blk_off[S] * (i64)G::BLK;
  __syncthreads();
  stream_image(Ys, pk + IMG);
  prefetch_z((i64)S * NBS + 1);
  bool have_pf = true;

  while (S >= 0) {
    const i64 s0 = (i64)S * NBS;
    const i64 R0 = s0 + 1 + (i64)t * B;
    // ---- phase A: retire rows that left the window, bring the entering rows in
    if (t == 0) {
      evict(hi);
      lo = hi = R0;
    } else {
      evict(R0);
    }
    if (!have_pf) prefetch_z(hi);  // synchronous path (tiny sweep blocks only)
    q2_cp_wait_all();              // Y(t) has landed
    __syncthreads();               // (1) ring slots of evicted rows are free; Y visible to all
    commit_z();
    hi += B;
    fetch(R0 + HP);                // no-op except for a fresh window (t == 0): its last HP - B rows
    stream_image(Vs, pk);          // V(t): needed by product 2 only
    __syncthreads();               // (2) window complete
    // ---- next block of the walk; prefetch its entering rows of Z
    int nS_ = S, nt_ = t + 1;
    const double* npk = pk + G::BLK;
    if (nt_ >= ntask) {
      nS_ = S - 1;
      nt_ = 0;
      if (nS_ >= 0) npk = packed + blk_off[nS_] * (i64)G::BLK;
    }
    have_pf = false;
    if (nS_ >= 0) {
      if (nt_ > 0) {
        prefetch_z(hi);  // the next window is [R0+B, R0+B+HP): rows [hi, hi+B) enter
        have_pf = true;
      } else if (ntask >= 4) {
        // next sweep block restarts at the top: those rows were evicted >= 2 barriers ago (ntask >= 4)
        prefetch_z((i64)nS_ * NBS + 1);
        have_pf = true;
      }
    }
    const int zbase = ring(R0);  // even: R0 is odd

    // ---- product 1: W(i,c) = sum_r Y(r,i) Zw(r,c);  rows i of this warp: [wm*M1, wm*M1+M1)
    {
      double acc[MI1][NI][2];
#pragma unroll
      for (int a = 0; a < MI1; ++a)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[a][j][0] = acc[a][j][1] = 0.0;
      const int i0 = wm * M1;
      const double* ya = Ys + (i0 + lq) * G::LDV + lr;
      const double* zb = Zs + (wn * 16 + lq) * LDZ;
#pragma unroll 4
      for (int kk = i0; kk < HP; kk += 4) {  // Y(r,i) == 0 for r < i
        double af[MI1], bf[NI];
        int zr = zbase + kk + lr;
        if (zr >= HP) zr -= HP;
#pragma unroll
        for (int a = 0; a < MI1; ++a) af[a] = ya[a * 8 * G::LDV + kk];
#pragma unroll
        for (int j = 0; j < NI; ++j) bf[j] = zb[j * 8 * LDZ + zr];
#pragma unroll
        for (int a = 0; a < MI1; ++a)
#pragma unroll
          for (int j = 0; j < NI; ++j) dmma884_(acc[a][j][0], acc[a][j][1], bf[j], af[a]);
      }
#pragma unroll
      for (int a = 0; a < MI1; ++a)
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int i = i0 + a * 8 + 2 * lr, c = wn * 16 + j * 8 + lq;
          *reinterpret_cast<double2*>(Ws + c * LDW + i) = make_double2(acc[a][j][0], acc[a][j][1]);
        }
    }
    q2_cp_wait_all();   // V(t) has landed
    __syncthreads();    // (3) W and V visible; Ys is free
    if (nS_ >= 0) stream_image(Ys, npk + IMG);  // Y(t+1) streams in under product 2
    // ---- product 2: Zw(r,c) -= sum_i V(r,i) W(i,c);  rows r of this warp: [wm*M3, wm*M3+M3)
    {
      double acc[MI3][NI][2];
#pragma unroll
      for (int a = 0; a < MI3; ++a)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[a][j][0] = acc[a][j][1] = 0.0;
      const int r0w = wm * M3;
      const double* va = Vs + lr * G::LDV + r0w + lq;
      const double* wb = Ws + (wn * 16 + lq) * LDW + lr;
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
          const int r = r0w + a * 8 + 2 * lr, c = wn * 16 + j * 8 + lq;
          int q0 = zbase + r;  // even, and q0 + 1 < HP: the pair never straddles the ring seam
          if (q0 >= HP) q0 -= HP;
          double2* zp = reinterpret_cast<double2*>(Zs + c * LDZ + q0);
          double2 zv = *zp;
          zv.x -= acc[a][j][0];
          zv.y -= acc[a][j][1];
          *zp = zv;
        }
    }
    __syncthreads();    // (4) window updated; Vs is free
    // ---- advance
    if (nt_ == 0 && nS_ >= 0) ntask = q2_num_tasks(n, B, (i64)nS_ * NBS);
    S = nS_;
    t = nt_;
    pk = npk;
  }
  evict(min(hi, n));
}

template <int B, int NBS, int NWN>
static size_t q2_smem_bytes() {
  using G = Q2Geom<B, NBS>;
  constexpr int KC = 16 * NWN;
  return (size_t)(KC * (G::HP + 4) + G::BLK + KC * (NBS + 4)) * sizeof(double);
}

// Work per SM is what bounds the walk (every CTA is a serial chain over all diamond blocks): choose the slab
// width that minimises ceil(#CTA / #SM) * KC, the columns the busiest SM has to process.
template <int B, int NBS, int NWN>
static void q2_consider(Ctx* ctx, i64 k, int force_kc, long long* best_cost, int* best_nwn) {
  constexpr int KC = 16 * NWN;
  if (force_kc > 0 && force_kc != KC) return;
  const long long nct = (k + KC - 1) / KC;
  const long long cost = ((nct + ctx->num_sms - 1) / ctx->num_sms) * KC;
  if (*best_nwn == 0 || cost < *best_cost || (cost == *best_cost && NWN > *best_nwn)) {
    *best_cost = cost;
    *best_nwn = NWN;
  }
}

template <int B, int NBS, int NWN>
static cudaError_t q2_apply_launch(Ctx* ctx, const double* packed, const i64* d_off, i64 n, int nS, double* Z, i64 ldz,
                                   i64 k) {
  const size_t smem = q2_smem_bytes<B, NBS, NWN>();
  auto kern = q2_apply_kernel<B, NBS, NWN>;
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ce != cudaSuccess) return ce;
  kern<<<cdiv(k, 16 * NWN), 64 * NWN, smem, ctx->stream>>>(packed, d_off, n, nS, Z, ldz, k);
  EKB_COUNT_LAUNCH(ctx);
  return cudaGetLastError();
}

template <int B, int NBS>
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
    q2_pack_kernel<B, NBS><<<dim3(tmax, nS), 128, 0, ctx->stream>>>(V2, ldv, TAU2, ldtau, n, d_off, packed);
    EKB_COUNT_LAUNCH(ctx);
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess) {
    long long cost = 0;
    int nwn = 0;
    q2_consider<B, NBS, 2>(ctx, k, ctx->q2_kc, &cost, &nwn);
    q2_consider<B, NBS, 3>(ctx, k, ctx->q2_kc, &cost, &nwn);
    switch (nwn) {
      case 2: ce = q2_apply_launch<B, NBS, 2>(ctx, packed, d_off, n, nS, Z, ldz, k); break;
      case 3: ce = q2_apply_launch<B, NBS, 3>(ctx, packed, d_off, n, nS, Z, ldz, k); break;
      default: ce = cudaErrorInvalidValue;
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
  if (b == 64) return q2_launch<64, 32>(ctx, n, V2, ldv, TAU2, ldtau, k, Z, ldz);
  if (b == 32) return q2_launch<32, 32>(ctx, n, V2, ldv, TAU2, ldtau, k, Z, ldz);
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
