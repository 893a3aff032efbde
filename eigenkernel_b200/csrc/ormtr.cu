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
//   applies the whole of Q2: every warp owns 16 columns of Z, keeps the moving window of its columns in
//   registers as DMMA accumulator tiles and walks all diamond blocks with two products per block
//   (W = Y^T Zw with Y = V T^T, then Zw -= V W); see q2_apply_kernel.
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

// Compile-time geometry of the diamond blocks.  Window coordinates r' = row - W0 with W0 = s0 + t*B (a
// multiple of 8, so 8-row accumulator tiles never straddle a window edge and row pairs are 16-byte aligned);
// reflector i of the block occupies r' in [i+1, i+B].
template <int B, int NBS>
struct Q2Geom {
  static constexpr int HP = B + NBS;               // rows of the window
  static constexpr int TILES = HP / 8;             // 8-row accumulator tiles of the window
  static constexpr int SHIFT = B / 8;              // tiles retired (and entering) per chase step
  static constexpr int MI = NBS / 8;               // 8-reflector tiles of a block
  static constexpr int LDY = HP + 8;               // Y image: Ys[i*LDY + r']   (K-major in r')
  static constexpr int LDVT = NBS + 8;             // V image: Vt[r'*LDVT + i]  (K-major in i), holds -V
  static constexpr int IMG_Y = NBS * LDY;
  static constexpr int IMG = IMG_Y + HP * LDVT;    // doubles per packed block
  static_assert(HP % 8 == 0 && B % 8 == 0 && NBS % 8 == 0, "tile granularity");
  // 128-bit fragment loads are bank-conflict free iff the row stride is 8 mod 16 doubles
  static_assert(LDY % 16 == 8 && LDVT % 16 == 8, "bank-conflict-free strides");
};

// ------------------------------------------------------------------------------------------ q2 pack
// One CTA per diamond block (t = blockIdx.x, S = blockIdx.y): gathers V from V2 with the structural zeros
// made explicit, forms the NBS x NBS compact-WY factor T (forward, columnwise), folds it into Y = V T^T
// (so that G Z = Z - V (Y^T Z) needs two products, not three) and writes Y and -V as the shared-memory
// images consumed by q2_apply_kernel.
template <int B, int NBS>
__global__ void __launch_bounds__(128) q2_pack_kernel(const double* __restrict__ V2, i64 ldv,
                                                      const double* __restrict__ TAU2, int ldtau, i64 n,
                                                      const i64* __restrict__ blk_off, double* __restrict__ packed) {
  using G = Q2Geom<B, NBS>;
  constexpr int H = B + NBS - 1;   // rows r = r' - 1 of the parallelogram
  constexpr int LDS_ = H + 2;
  __shared__ double Vs[NBS * LDS_];  // Vs[i*LDS_ + r]
  __shared__ double Gs[NBS][NBS + 1];
  __shared__ double Ts[NBS][NBS + 1];
  __shared__ double taus[NBS];
  const int t = blockIdx.x, S = blockIdx.y, tid = threadIdx.x;
  const i64 s0 = (i64)S * NBS;
  if (t >= q2_num_tasks(n, B, s0)) return;
  const i64 R0 = s0 + 1 + (i64)t * B;
  for (int idx = tid; idx < NBS * LDS_; idx += blockDim.x) {
    const int i = idx / LDS_, r = idx % LDS_;
    const i64 s = s0 + i;
    double v = 0.0;
    if (r < H && s <= n - 3) {
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
      const double* vi = Vs + i * LDS_;
      const double* vj = Vs + j * LDS_;
      for (int r = j; r < min(H, i + B); ++r) g += vi[r] * vj[r];
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
  double* out = packed + (blk_off[S] + t) * (i64)G::IMG;
  // Y image: Y(r,i) = sum_{j >= i} V(r,j) T(i,j), stored at r' = r + 1
  for (int idx = tid; idx < G::IMG_Y; idx += blockDim.x) {
    const int i = idx / G::LDY, rp = idx % G::LDY;
    double y = 0.0;
    if (rp >= 1 && rp - 1 < H) {
      const int r = rp - 1;
      for (int j = i; j < NBS; ++j) y += Vs[j * LDS_ + r] * Ts[i][j];
    }
    out[idx] = y;
  }
  // -V image, transposed
  double* outV = out + G::IMG_Y;
  for (int idx = tid; idx < G::HP * G::LDVT; idx += blockDim.x) {
    const int rp = idx / G::LDVT, i = idx % G::LDVT;
    double v = 0.0;
    if (i < NBS && rp >= 1 && rp - 1 < H) v = -Vs[i * LDS_ + rp - 1];
    outV[idx] = v;
  }
}

// ------------------------------------------------------------------------------------------ q2 apply
// TMA bulk copy (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: one instruction moves a whole block image.
__device__ __forceinline__ unsigned q2_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void q2_mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(q2_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void q2_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(q2_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void q2_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   q2_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(q2_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void q2_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(q2_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void q2_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// One CTA owns 8 NJ NW columns of Z and walks every diamond block (S descending, t ascending).  Each WARP owns 8 NJ
// of those columns (NJ = 2 normally, 1 for narrow slabs) and keeps its HP x 8 NJ piece of the moving window in
// REGISTERS, as DMMA accumulator tiles
// (thread (lq,lr) holds rows 8 rho + 2 lr + {0,1} of column lq), for the whole walk:
//     W = Y^T Zw     the Z tiles are used directly as the k-side MMA operand: within an 8-row tile the two
//                    DMMA.8x8x4 steps take k = {2 lr} and k = {2 lr + 1} (a permutation of the summation
//                    index applied to both operands), which is exactly what each thread already holds;
//     Zw += (-V) W   W likewise goes from accumulator layout straight into the k-side operand.
// So the window never touches shared memory, there is no cross-warp exchange and no barrier between the two
// products; shared memory only carries the (Y, -V) images of the block (double buffered, each streamed by
// ONE cp.async.bulk (TMA, mbarrier completion) one block ahead), read with conflict-free 128-bit loads, 0.25 loads per DMMA.  Rows that leave the
// window are stored from registers, rows that enter are prefetched into registers one step ahead.  Every
// global element is always touched by the same thread, so program order is all the ordering the walk needs.
// RING (round 2): a (NW+1)-th warp streams the block images into a ring of THREE buffers and the compute warps
// synchronise with it only through mbarriers (full: TMA completion; empty: one arrival per compute warp), so no warp
// ever waits for another compute warp.  Round 1 double-buffered with a __syncthreads per block: 13 % of all stall
// samples sat behind that barrier, and the scheduler that holds a single warp (7 warps on 4 schedulers) idled at it.
template <int B, int NBS, int NW, bool AL16, int NJ, bool RING>
__global__ void __launch_bounds__(32 * (NW + (RING ? 1 : 0)), 1) q2_apply_kernel(const double* __restrict__ packed,
                                                             const i64* __restrict__ blk_off, i64 n, int nS,
                                                             double* __restrict__ Z, i64 ldz, i64 k) {
  using G = Q2Geom<B, NBS>;
  constexpr int TILES = G::TILES, SHIFT = G::SHIFT, KEEP = TILES - SHIFT, MI = G::MI;
  constexpr int LDY = G::LDY, LDVT = G::LDVT, IMG = G::IMG;
  static_assert(KEEP >= 0 && KEEP <= SHIFT && IMG % 2 == 0, "geometry");
  extern __shared__ __align__(16) double sm[];  // two image buffers
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lq = lane >> 2, lr = lane & 3;
  const i64 cwarp = (i64)blockIdx.x * (8 * NJ * NW) + 8 * NJ * warp;  // first column of this warp
  const bool cols_full = cwarp + 8 * NJ <= k;                     // warp-uniform
  double* zc[NJ];
  bool cv[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    cv[j] = cwarp + lq + 8 * j < k;
    zc[j] = Z + (cv[j] ? (cwarp + lq + 8 * j) * ldz : 0) + 2 * lr;  // row offset 2 lr folded in
  }
  // rows (row, row+1) of this thread's column j, row = tile row + 2 lr (even).  FAST: no bounds checks.
  auto ld2 = [&](i64 trow, int j, bool fast) -> double2 {
    const double* p = zc[j] + trow;
    if (fast) {
      if (AL16) return __ldcg(reinterpret_cast<const double2*>(p));
      return make_double2(__ldcg(p), __ldcg(p + 1));
    }
    double2 v = make_double2(0.0, 0.0);
    const i64 row = trow + 2 * lr;
    if (cv[j]) {
      if (row < n) v.x = __ldcg(p);
      if (row + 1 < n) v.y = __ldcg(p + 1);
    }
    return v;
  };
  auto st2 = [&](i64 trow, int j, bool fast, double x, double y) {
    double* p = zc[j] + trow;
    if (fast) {
      if (AL16) *reinterpret_cast<double2*>(p) = make_double2(x, y);
      else { p[0] = x; p[1] = y; }
      return;
    }
    const i64 row = trow + 2 * lr;
    if (cv[j]) {
      if (row < n) p[0] = x;
      if (row + 1 < n) p[1] = y;
    }
  };

  constexpr int NBUF = RING ? 3 : 2;
  __shared__ unsigned long long bars[NBUF], ebars[NBUF];
  constexpr unsigned IMG_BYTES = IMG * sizeof(double);
  static_assert(IMG_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < NBUF; ++q) {
      q2_mbar_init(&bars[q], 1);
      q2_mbar_init(&ebars[q], NW);
    }
    q2_fence_proxy_async();
  }
  __syncthreads();
  if (RING && warp == NW) {
    // ---- producer warp: one bulk copy per diamond block, in the order the compute warps walk them
    if (lane == 0) {
      int S = nS - 1;
      while (S >= 0 && q2_num_tasks(n, B, (i64)S * NBS) == 0) --S;
      unsigned i = 0;
      for (; S >= 0; --S) {
        const int ntask = q2_num_tasks(n, B, (i64)S * NBS);
        const double* pk = packed + blk_off[S] * (i64)IMG;
        for (int t = 0; t < ntask; ++t, pk += IMG, ++i) {
          const unsigned st = i % NBUF;
          if (i >= (unsigned)NBUF) q2_mbar_wait(&ebars[st], ((i / NBUF) - 1u) & 1u);
          q2_fence_proxy_async();
          q2_mbar_expect_tx(&bars[st], IMG_BYTES);
          q2_bulk_g2s(sm + st * IMG, pk, IMG_BYTES, &bars[st]);
        }
      }
    }
    return;
  }
  auto stream_image = [&](int b, const double* src) {  // one elected thread; completion lands on bars[b]
    if (tid == 0) {
      q2_fence_proxy_async();
      q2_mbar_expect_tx(&bars[b], IMG_BYTES);
      q2_bulk_g2s(sm + b * IMG, src, IMG_BYTES, &bars[b]);
    }
  };
  unsigned phase = 0;  // bit b: parity to wait for on bars[b]

  double zacc[TILES][NJ][2];   // the window
  double zin[SHIFT][NJ][2];    // rows entering at the next step

  int S = nS - 1;
  while (S >= 0 && q2_num_tasks(n, B, (i64)S * NBS) == 0) --S;
  if (S < 0) return;
  if (!RING) stream_image(0, packed + blk_off[S] * (i64)IMG);
  int buf = 0;
  unsigned blk = 0;  // RING: running index of the diamond block

  for (; S >= 0; --S) {
    const i64 s0 = (i64)S * NBS;
    const int ntask = q2_num_tasks(n, B, s0);
    const double* pk = packed + blk_off[S] * (i64)IMG;
    // ---- fresh window [s0, s0 + HP).  The "+ 0.0" makes the loads complete HERE (a real consumer), so the
    // walk below never waits on a scoreboard shared with its own prefetch loads.
    {
      const bool fast = cols_full && s0 + G::HP <= n;
#pragma unroll
      for (int r = 0; r < TILES; ++r)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const double2 v = ld2(s0 + 8 * r, j, fast);
          zacc[r][j][0] = v.x + 0.0;
          zacc[r][j][1] = v.y + 0.0;
        }
    }
    for (int t = 0; t < ntask; ++t, pk += IMG) {
      const i64 W0 = s0 + (i64)t * B;
      const bool last = t + 1 >= ntask;
      if (RING) {
        buf = (int)(blk % NBUF);
        q2_mbar_wait(&bars[buf], (blk / NBUF) & 1u);  // this block's images have landed (the producer runs ahead)
      } else {
        q2_mbar_wait(&bars[buf], (phase >> buf) & 1u);  // this block's images have landed (streamed one block ahead)
        phase ^= 1u << buf;
        __syncthreads();    // every warp is done with the other buffer
      }
      const double* Ys = sm + buf * IMG;
      const double* Vt = Ys + G::IMG_Y;
      if (!RING) {
        if (!last) stream_image(buf ^ 1, pk + IMG);
        else if (S > 0) stream_image(buf ^ 1, packed + blk_off[S - 1] * (i64)IMG);
      }
      if (!last) {        // rows entering the next window of this sweep block
        const bool fast = cols_full && W0 + G::HP + B <= n;
#pragma unroll
        for (int q = 0; q < SHIFT; ++q)
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const double2 v = ld2(W0 + G::HP + 8 * q, j, fast);
            zin[q][j][0] = v.x;
            zin[q][j][1] = v.y;
          }
      }
      // ---- product 1: W(i,c) = sum_r' Y(r',i) Zw(r',c)
      double wacc[MI][NJ][2];
#pragma unroll
      for (int a = 0; a < MI; ++a)
#pragma unroll
        for (int j = 0; j < NJ; ++j) wacc[a][j][0] = wacc[a][j][1] = 0.0;
      // (Measured, round 2: issuing the x halves of several accumulators before their y halves, so that no two
      // consecutive DMMAs accumulate into the same registers, made the walk 2-4 % SLOWER -- back-to-back dependent
      // DMMAs are not what holds the tensor pipe at ~70 %; the extra live fragments cost more than they gave.)
      {
        const double* yp = Ys + lq * LDY + 2 * lr;
#pragma unroll
        for (int r = 0; r < TILES; ++r)
#pragma unroll
          for (int a = 0; a < MI; ++a)
            if (r >= a) {  // Y(r',i) == 0 for r' <= i
              const double2 y = *reinterpret_cast<const double2*>(yp + a * 8 * LDY + 8 * r);
#pragma unroll
              for (int j = 0; j < NJ; ++j) {
                dmma884_(wacc[a][j][0], wacc[a][j][1], zacc[r][j][0], y.x);
                dmma884_(wacc[a][j][0], wacc[a][j][1], zacc[r][j][1], y.y);
              }
            }
      }
      // ---- product 2: Zw(r',c) += sum_i (-V)(r',i) W(i,c)
      {
        const double* vp = Vt + lq * LDVT + 2 * lr;
#pragma unroll
        for (int a = 0; a < MI; ++a)
#pragma unroll
          for (int r = 0; r < TILES; ++r)
            if (r >= a && r <= a + B / 8) {  // V(r',i) != 0 only for i < r' <= i + B
              const double2 v = *reinterpret_cast<const double2*>(vp + r * 8 * LDVT + 8 * a);
#pragma unroll
              for (int j = 0; j < NJ; ++j) {
                dmma884_(zacc[r][j][0], zacc[r][j][1], wacc[a][j][0], v.x);
                dmma884_(zacc[r][j][0], zacc[r][j][1], wacc[a][j][1], v.y);
              }
            }
      }
      if (RING) {  // this warp is done with the images of the block: one arrival on the buffer's "empty" barrier
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(q2_smem_u32(&ebars[buf])) : "memory");
        ++blk;
      }
      // ---- retire the rows that leave the window, shift, take the entering rows
      {
        const bool fast = cols_full && W0 + G::HP <= n;
        if (!last) {
#pragma unroll
          for (int r = 0; r < SHIFT; ++r)
#pragma unroll
            for (int j = 0; j < NJ; ++j) st2(W0 + 8 * r, j, fast, zacc[r][j][0], zacc[r][j][1]);
#pragma unroll
          for (int r = 0; r < KEEP; ++r)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              zacc[r][j][0] = zacc[r + SHIFT][j][0];
              zacc[r][j][1] = zacc[r + SHIFT][j][1];
            }
#pragma unroll
          for (int q = 0; q < SHIFT; ++q)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              zacc[KEEP + q][j][0] = zin[q][j][0];
              zacc[KEEP + q][j][1] = zin[q][j][1];
            }
        } else {
#pragma unroll
          for (int r = 0; r < TILES; ++r)
#pragma unroll
            for (int j = 0; j < NJ; ++j) st2(W0 + 8 * r, j, fast, zacc[r][j][0], zacc[r][j][1]);
        }
      }
      if (!RING) buf ^= 1;
    }
  }
}

// Every CTA is a serial chain over all diamond blocks, so the walk takes (waves of CTAs) x (time of one CTA).
// Measured on B200 (scripts/q2_slab_probe.py, n = 16384: 166 / 245 / 337 ms for 4 / 8 / 12 warps of 8 columns,
// 323 / 448 ms for 4 / 7 warps of 16 columns) the time of one CTA is proportional to NJ (4.4 + NW): a lone warp per
// scheduler is latency-bound, more and thinner warps interleave on the DMMA pipe.
// NJ = accumulator column tiles per warp: 2 (16 columns) normally; 1 (8 columns per warp) doubles the number of
// warps when a rank's slab is too narrow to give every scheduler of every SM a warp (multi-GPU column slabs); 14 such
// warps + the producer (480 threads, 128 registers) cover the same 112 columns per CTA as 7 warps of 16 and walk them
// 2.6 % faster (n = 32768: 3.324 -> 3.238 s, round 2) -- far less than the pre-ring model promised: with the ring the
// fixed cost per block is gone and both shapes sit at ~71 % of the DMMA rate.
// The choice is returned as NW + 100 * (NJ == 1).
template <int NW, int NJ>
static void q2_consider(Ctx* ctx, i64 k, int force_kc, long long* best_cost, int* best_nw) {
  constexpr int KC = 8 * NJ * NW;
  // q2_kc forces a kernel: 64..128 = 16-columns-per-warp with that many columns per CTA; 1000 + NW = the
  // 8-columns-per-warp kernel with NW warps
  if (force_kc >= 1000 && (NJ != 1 || force_kc != 1000 + NW)) return;
  if (force_kc > 0 && force_kc < 1000 && (force_kc != KC || NJ != 2)) return;
  const long long nct = (k + KC - 1) / KC;
  const long long cost = ((nct + ctx->num_sms - 1) / ctx->num_sms) * NJ * (44 + 10 * NW);
  const int code = NW + (NJ == 1 ? 100 : 0);
  if (*best_nw == 0 || cost < *best_cost) {
    *best_cost = cost;
    *best_nw = code;
  }
}

template <int B, int NBS, int NW, bool AL16, int NJ = 2>
static cudaError_t q2_apply_launch(Ctx* ctx, const double* packed, const i64* d_off, i64 n, int nS, double* Z, i64 ldz,
                                   i64 k) {
  using G = Q2Geom<B, NBS>;
  // the producer warp costs a warp's registers: 8 warps of 16 columns (250 registers each) leave no room for it
  constexpr bool RING = !(NW == 8 && NJ == 2);
  const size_t smem = (size_t)(RING ? 3 : 2) * G::IMG * sizeof(double);
  auto kern = q2_apply_kernel<B, NBS, NW, AL16, NJ, RING>;
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ce != cudaSuccess) return ce;
  kern<<<cdiv(k, 8 * NJ * NW), 32 * (NW + (RING ? 1 : 0)), smem, ctx->stream>>>(packed, d_off, n, nS, Z, ldz, k);
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
  int rc = ctx_alloc(ctx, (void**)&packed, (size_t)nblk * G::IMG * sizeof(double));
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
    int nw = 0;
    q2_consider<4, 2>(ctx, k, ctx->q2_kc, &cost, &nw);
    q2_consider<5, 2>(ctx, k, ctx->q2_kc, &cost, &nw);
    q2_consider<6, 2>(ctx, k, ctx->q2_kc, &cost, &nw);
    q2_consider<7, 2>(ctx, k, ctx->q2_kc, &cost, &nw);
    q2_consider<8, 2>(ctx, k, ctx->q2_kc, &cost, &nw);
    q2_consider<4, 1>(ctx, k, ctx->q2_kc, &cost, &nw);
    q2_consider<8, 1>(ctx, k, ctx->q2_kc, &cost, &nw);
    q2_consider<12, 1>(ctx, k, ctx->q2_kc, &cost, &nw);
    q2_consider<14, 1>(ctx, k, ctx->q2_kc, &cost, &nw);
    const bool al16 = ((uintptr_t)Z & 15) == 0 && (ldz & 1) == 0;
    if (!al16) nw = -4;  // 8-byte global accesses: one generic instantiation
    prof_begin(ctx, PROF_Q2_APPLY, 2.0 * (double)n * (double)n * (double)k);
    switch (nw) {
      case -4: ce = q2_apply_launch<B, NBS, 4, false>(ctx, packed, d_off, n, nS, Z, ldz, k); break;
      case 4: ce = q2_apply_launch<B, NBS, 4, true>(ctx, packed, d_off, n, nS, Z, ldz, k); break;
      case 5: ce = q2_apply_launch<B, NBS, 5, true>(ctx, packed, d_off, n, nS, Z, ldz, k); break;
      case 6: ce = q2_apply_launch<B, NBS, 6, true>(ctx, packed, d_off, n, nS, Z, ldz, k); break;
      case 7: ce = q2_apply_launch<B, NBS, 7, true>(ctx, packed, d_off, n, nS, Z, ldz, k); break;
      case 8: ce = q2_apply_launch<B, NBS, 8, true>(ctx, packed, d_off, n, nS, Z, ldz, k); break;
      case 104: ce = q2_apply_launch<B, NBS, 4, true, 1>(ctx, packed, d_off, n, nS, Z, ldz, k); break;
      case 108: ce = q2_apply_launch<B, NBS, 8, true, 1>(ctx, packed, d_off, n, nS, Z, ldz, k); break;
      case 112: ce = q2_apply_launch<B, NBS, 12, true, 1>(ctx, packed, d_off, n, nS, Z, ldz, k); break;
      case 114: ce = q2_apply_launch<B, NBS, 14, true, 1>(ctx, packed, d_off, n, nS, Z, ldz, k); break;
      default: ce = cudaErrorInvalidValue;
    }
    prof_end(ctx);
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
constexpr int Q1_GROUP = 8;  // panels aggregated into one WY block

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
