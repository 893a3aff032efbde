// FP64 GEMM engine on DMMA.8x8x4 tensor-core tiles (sm_100a).
//
// Every GEMM-shaped step of the eigensolve (PDPOTRF/PDSYGST/PDTRTRS trailing updates, the SYMM and
// SYR2K of the dense->band reduction, the D&C merge products and both back-transformations; reference
// call sites generalized_to_standard.f90:24,37,103 and solver_scalapack_all.f90:59,96,115) runs through
// this one kernel family.  tcgen05/UMMA has no f64 kind, so the Blackwell FP64 tensor path is
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) fed from shared memory that is filled by a 4-stage cp.async
// (LDGSTS) pipeline.
//
// Layout trick: the MMA is issued on the TRANSPOSED problem (MMA "row" operand = op(B)^T, "col"
// operand = op(A)^T) so that the two accumulators a thread owns are adjacent in the column-major M
// direction -> 16-byte vector stores to C.
//
// Shared-memory layouts (doubles), both bank-conflict free for the fragment pattern
// (mn = base + lane/4, k = k4 + lane%4):
//   K-major  (operand contiguous in k in global):  s[mn * (BK+4) + k]
//   MN-major (operand contiguous in m/n in global): s[k * (BMN+4) + mn]
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace ekb {

// Pipeline geometry = (BK, STAGES): k-depth of a stage and number of stages, template parameters of the kernel.  Two
// are instantiated: (16, 4) -- short prologue, best when K is small (the K = 2b updates of dense-to-band run only 8
// k-tiles) -- and (32, 3) -- half as many block barriers per DMMA; measured +3 % on 8192^3 and +9 % on the N = 64
// panel products, -14 % on the K = 128 update (profiles/r02_gemm_shapes_*.jsonl).  K-major row stride BK + 4 doubles
// (8 words mod 32: conflict-free 64-bit fragment loads per half warp for both values of BK).
constexpr int GEMM_THREADS = 256;

__device__ __forceinline__ void cp_async16(double* smem, const double* g, int bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(double* smem, const double* g, int bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// Load one (BMN x BK) operand tile.  Element (mn, k) lives at g[(mn0+mn) * s_mn + (k0+k) * s_k] where
// exactly one of the two global strides is 1 (kmajor: s_k == 1).
template <int BMN, int BK>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ g, i64 ld, bool kmajor, bool al16,
                                          int mn0, int k0, int mn_lim, int k_lim, int tid) {
  constexpr int LDK = BK + 4;
  if (kmajor) {
    if (al16) {
#pragma unroll
      for (int c = tid; c < BMN * (BK / 2); c += GEMM_THREADS) {
        int mn = c / (BK / 2), kc = (c % (BK / 2)) * 2;
        int gm = mn0 + mn, gk = k0 + kc;
        int bytes = (gm < mn_lim) ? max(0, min(16, (k_lim - gk) * 8)) : 0;
        const double* src = bytes > 0 ? g + (i64)gm * ld + gk : g;
        cp_async16(s + mn * LDK + kc, src, bytes);
      }
    } else {
#pragma unroll
      for (int c = tid; c < BMN * BK; c += GEMM_THREADS) {
        int mn = c / BK, kc = c % BK;
        int gm = mn0 + mn, gk = k0 + kc;
        int bytes = (gm < mn_lim && gk < k_lim) ? 8 : 0;
        const double* src = bytes > 0 ? g + (i64)gm * ld + gk : g;
        cp_async8(s + mn * LDK + kc, src, bytes);
      }
    }
  } else {
    constexpr int LDM = BMN + 4;
    if (al16) {
#pragma unroll
      for (int c = tid; c < BK * (BMN / 2); c += GEMM_THREADS) {
        int k = c / (BMN / 2), mc = (c % (BMN / 2)) * 2;
        int gm = mn0 + mc, gk = k0 + k;
        int bytes = (gk < k_lim) ? max(0, min(16, (mn_lim - gm) * 8)) : 0;
        const double* src = bytes > 0 ? g + (i64)gk * ld + gm : g;
        cp_async16(s + k * LDM + mc, src, bytes);
      }
    } else {
#pragma unroll
      for (int c = tid; c < BK * BMN; c += GEMM_THREADS) {
        int k = c / BMN, mc = c % BMN;
        int gm = mn0 + mc, gk = k0 + k;
        int bytes = (gk < k_lim && gm < mn_lim) ? 8 : 0;
        const double* src = bytes > 0 ? g + (i64)gk * ld + gm : g;
        cp_async8(s + k * LDM + mc, src, bytes);
      }
    }
  }
}

// ---- interior fast path: the whole (BMN x BK) tile is in range and 16-byte aligned, so a thread's copies differ
// from one another (and from one k-tile to the next) only by compile-time strides; dst/src already carry the
// thread's own offset.
__device__ __forceinline__ void cp_async16_full(unsigned s, const double* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g) : "memory");
}
template <int BMN, int BK>
__device__ __forceinline__ void load_tile_fast_km(unsigned dst, const double* __restrict__ src, i64 ld) {
  constexpr int LDK = BK + 4;
  constexpr int RPP = GEMM_THREADS / (BK / 2);  // tile rows (mn) covered per pass
#pragma unroll
  for (int q = 0; q < BMN / RPP; ++q) cp_async16_full(dst + q * RPP * LDK * 8, src + (i64)q * RPP * ld);
}
template <int BMN, int BK>
__device__ __forceinline__ void load_tile_fast_mn(unsigned dst, const double* __restrict__ src, i64 ld) {
  constexpr int KPP = GEMM_THREADS / (BMN / 2);  // k rows covered per pass
#pragma unroll
  for (int q = 0; q < BK / KPP; ++q) cp_async16_full(dst + q * KPP * (BMN + 4) * 8, src + (i64)q * KPP * ld);
}

// One k-tile of DMMAs with compile-time shared-memory strides (immediate LDS offsets, no address arithmetic).
template <int MI, int NI, int BM, int BN, int BK, bool AKM, bool BKM>
__device__ __forceinline__ void mma_ktile(double (&acc)[MI][NI][2], const double* __restrict__ tA,
                                          const double* __restrict__ tB) {
  constexpr int LDK = BK + 4;
  constexpr int a_smn = AKM ? LDK : 1, a_sk = AKM ? 1 : (BM + 4);
  constexpr int b_smn = BKM ? LDK : 1, b_sk = BKM ? 1 : (BN + 4);
#pragma unroll
  for (int kk = 0; kk < BK / 4; ++kk) {
    double af[MI], bf[NI];
#pragma unroll
    for (int i = 0; i < MI; ++i) af[i] = tA[i * 8 * a_smn + kk * 4 * a_sk];
#pragma unroll
    for (int j = 0; j < NI; ++j) bf[j] = tB[j * 8 * b_smn + kk * 4 * b_sk];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], bf[j], af[i]);
  }
}

template <int BM, int BN, int WM, int WN, bool BATCHED, int BK, int GEMM_STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(GemmP p0, const GemmP* __restrict__ batch, int flags, int tri_keep, int splitk, double* __restrict__ ws) {
  extern __shared__ __align__(16) double smem[];
  constexpr int LDK = BK + 4;
  constexpr int A_TILE = BM * LDK;  // >= BK*(BM+4)
  constexpr int B_TILE = BN * LDK;
  constexpr int MI = WM / 8, NI = WN / 8;
  constexpr int WARPS_M = BM / WM;
  static_assert((BM / WM) * (BN / WN) == GEMM_THREADS / 32, "8 warps");

  GemmP p = p0;
  if (BATCHED) p = batch[blockIdx.z];
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (m0 >= p.m || n0 >= p.n) return;
  if (tri_keep >= 0 && n0 - (m0 + BM - 1) >= tri_keep) return;

  int kt_begin = 0, kt_end = (p.k + BK - 1) / BK;
  double alpha = p.alpha, beta = p.beta;
  double* C = p.C;
  i64 ldc = p.ldc;
  if (!BATCHED && splitk > 1) {
    int per = (kt_end + splitk - 1) / splitk;
    kt_begin = blockIdx.z * per;
    kt_end = min(kt_end, kt_begin + per);
    C = ws + (i64)blockIdx.z * p.m * p.n;
    ldc = p.m;
    alpha = 1.0;
    beta = 0.0;
  }
  const int nk = max(0, kt_end - kt_begin);

  const bool ta = flags & GEMM_TA, tb = flags & GEMM_TB, syma = flags & GEMM_SYMA;
  const bool al16 = ((((uintptr_t)p.A | (uintptr_t)p.B) & 15) == 0) && ((p.lda | p.ldb) & 1) == 0;
  const bool b_kmajor = !tb;

  double* sA = smem;
  double* sB = smem + GEMM_STAGES * A_TILE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WARPS_M, wn = warp / WARPS_M;
  const int lq = lane >> 2, lr = lane & 3;

  auto a_is_kmajor = [&](int kt) -> bool {
    if (syma) return (kt * BK) >= m0 + BM;
    return ta;
  };
  // per-thread pieces of the interior fast path (see load_tile_fast_*)
  const bool a_full = al16 && (m0 + BM <= p.m);
  const bool b_full = al16 && (n0 + BN <= p.n);
  const unsigned sA32 = (unsigned)__cvta_generic_to_shared(sA), sB32 = (unsigned)__cvta_generic_to_shared(sB);
  const int r_km = tid / (BK / 2), kc_km = (tid % (BK / 2)) * 2;
  const i64 a_g_km = (i64)(m0 + r_km) * p.lda + kc_km;             // + k0
  const unsigned a_s_km = (unsigned)(r_km * LDK + kc_km) * 8u;
  const int a_k_mn = tid / (BM / 2), a_m_mn = (tid % (BM / 2)) * 2;
  const i64 a_g_mn = (i64)a_k_mn * p.lda + m0 + a_m_mn;            // + k0 * lda
  const unsigned a_s_mn = (unsigned)(a_k_mn * (BM + 4) + a_m_mn) * 8u;
  const int b_k_mn = tid / (BN / 2), b_n_mn = (tid % (BN / 2)) * 2;
  const i64 b_g = b_kmajor ? (i64)(n0 + r_km) * p.ldb + kc_km : (i64)b_k_mn * p.ldb + n0 + b_n_mn;
  const unsigned b_s = (b_kmajor ? (unsigned)(r_km * LDK + kc_km) : (unsigned)(b_k_mn * (BN + 4) + b_n_mn)) * 8u;
  auto issue = [&](int kt, int stage) {
    const int k0 = kt * BK;
    const bool kfull = k0 + BK <= p.k;
    const bool akm = a_is_kmajor(kt);
    if (a_full && kfull) {
      const unsigned st = sA32 + (unsigned)(stage * A_TILE) * 8u;
      if (akm) load_tile_fast_km<BM, BK>(st + a_s_km, p.A + a_g_km + k0, p.lda);
      else load_tile_fast_mn<BM, BK>(st + a_s_mn, p.A + a_g_mn + (i64)k0 * p.lda, p.lda);
    } else {
      load_tile<BM, BK>(sA + stage * A_TILE, p.A, p.lda, akm, al16, m0, k0, p.m, p.k, tid);
    }
    if (b_full && kfull) {
      const unsigned st = sB32 + (unsigned)(stage * B_TILE) * 8u;
      if (b_kmajor) load_tile_fast_km<BN, BK>(st + b_s, p.B + b_g + k0, p.ldb);
      else load_tile_fast_mn<BN, BK>(st + b_s, p.B + b_g + (i64)k0 * p.ldb, p.ldb);
    } else {
      load_tile<BN, BK>(sB + stage * B_TILE, p.B, p.ldb, b_kmajor, al16, n0, k0, p.n, p.k, tid);
    }
  };

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; ++s) {
    if (s < nk) issue(kt_begin + s, s);
    cp_async_commit();
  }

  // Accumulate-into-C updates (beta != 0) with alpha = +-1 -- every trailing update of the library (Cholesky, sygst,
  // the K = 2b SYR2K of dense-to-band, Q1, the triangular solves) -- start from acc = alpha beta C: the loads of C are
  // issued here, right behind the pipeline prologue, and land while the first operand tiles are in flight, and the
  // epilogue becomes store-only.  (The round-1 epilogue read each C element just before writing it: loads could not be
  // hoisted over the preceding stores, so a tile paid 32 dependent L2 round trips -- more than its whole main loop when
  // K = 128.)  Exact: alpha (alpha beta C + A B) = beta C + alpha A B for alpha = +-1.
  const bool cvec = (((uintptr_t)C & 15) == 0) && ((ldc & 1) == 0);
  if (beta != 0.0 && (alpha == 1.0 || alpha == -1.0)) {
    const double sc = alpha * beta;
    if (cvec && m0 + BM <= p.m && n0 + BN <= p.n) {
      // interior tile: straight-line code, all MI NI 128-bit loads in flight at once (with the bounds checks below every
      // load sits in its own basic block and is waited for before the next one is issued: measured 16 us per tile)
      const double* cbase = C + (i64)(n0 + wn * WN + lq) * ldc + (m0 + wm * WM + 2 * lr);
      double2 old[MI][NI];
#pragma unroll
      for (int j = 0; j < NI; ++j)
#pragma unroll
        for (int i = 0; i < MI; ++i) old[i][j] = __ldcg(reinterpret_cast<const double2*>(cbase + (i64)(j * 8) * ldc + i * 8));
#pragma unroll
      for (int j = 0; j < NI; ++j)
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          acc[i][j][0] = sc * old[i][j].x;
          acc[i][j][1] = sc * old[i][j].y;
        }
    } else
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int col = n0 + wn * WN + j * 8 + lq;
      if (col >= p.n) continue;
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int row = m0 + wm * WM + i * 8 + 2 * lr;
        if (row >= p.m) continue;
        const double* cp = C + (i64)col * ldc + row;
        if (row + 1 < p.m && cvec) {
          const double2 old = __ldcg(reinterpret_cast<const double2*>(cp));
          acc[i][j][0] = sc * old.x;
          acc[i][j][1] = sc * old.y;
        } else {
          acc[i][j][0] = sc * cp[0];
          if (row + 1 < p.m) acc[i][j][1] = sc * cp[1];
        }
      }
    }
    beta = 0.0;
  }

  // thread offsets of the fragment loads for either layout of each operand
  const int a_t_km = (wm * WM + lq) * LDK + lr, a_t_mn = (wm * WM + lq) + lr * (BM + 4);
  const int b_t = b_kmajor ? (wn * WN + lq) * LDK + lr : (wn * WN + lq) + lr * (BN + 4);

  for (int it = 0; it < nk; ++it) {
    cp_async_wait<GEMM_STAGES - 2>();
    __syncthreads();
    {
      int nx = it + GEMM_STAGES - 1;
      if (nx < nk) issue(kt_begin + nx, nx % GEMM_STAGES);
      cp_async_commit();
    }
    const int stage = it % GEMM_STAGES;
    const bool akm = a_is_kmajor(kt_begin + it);
    const double* tA = sA + stage * A_TILE + (akm ? a_t_km : a_t_mn);
    const double* tB = sB + stage * B_TILE + b_t;
    if (akm) {
      if (b_kmajor) mma_ktile<MI, NI, BM, BN, BK, true, true>(acc, tA, tB);
      else mma_ktile<MI, NI, BM, BN, BK, true, false>(acc, tA, tB);
    } else {
      if (b_kmajor) mma_ktile<MI, NI, BM, BN, BK, false, true>(acc, tA, tB);
      else mma_ktile<MI, NI, BM, BN, BK, false, false>(acc, tA, tB);
    }
  }
  cp_async_wait<0>();

  // epilogue: thread owns C(m0w + i*8 + 2*lr + {0,1}, n0w + j*8 + lq).  When C is still to be read (beta != 0 with
  // a general alpha) the MI loads of a column are issued together, ahead of that column's stores.
#pragma unroll
  for (int j = 0; j < NI; ++j) {
    const int col = n0 + wn * WN + j * 8 + lq;
    if (col >= p.n) continue;
    double2 old[MI];
    if (beta != 0.0) {
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int row = m0 + wm * WM + i * 8 + 2 * lr;
        old[i] = make_double2(0.0, 0.0);
        if (row >= p.m) continue;
        const double* cp = C + (i64)col * ldc + row;
        if (row + 1 < p.m && cvec) old[i] = *reinterpret_cast<const double2*>(cp);
        else {
          old[i].x = cp[0];
          if (row + 1 < p.m) old[i].y = cp[1];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int row = m0 + wm * WM + i * 8 + 2 * lr;
      if (row >= p.m) continue;
      double* cp = C + (i64)col * ldc + row;
      double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
      if (beta != 0.0) {
        v0 += beta * old[i].x;
        v1 += beta * old[i].y;
      }
      if (row + 1 < p.m && cvec) {
        *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
      } else {
        cp[0] = v0;
        if (row + 1 < p.m) cp[1] = v1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ TMA-fed variant
// Warp-specialised form of the same tile computation (round 2): ONE producer warp feeds the shared-memory ring with
// bulk asynchronous copies (cp.async.bulk.shared::cluster.global, the TMA engine; SASS UBLKCP) that complete on an
// mbarrier per stage; the eight consumer warps wait on that barrier, issue nothing but fragment loads and DMMAs, and
// hand the stage back through a second mbarrier.  Compared with the LDGSTS kernel above the consumers lose 16 copy
// instructions + their address arithmetic per k-tile and, more importantly, the block-wide barrier per k-tile.
// Operand tiles land in the SAME padded layouts (one copy per k-contiguous row of a K-major tile, or per k-row of an
// MN-major one: the destination of a 1-D bulk copy is free, a tensor-map box would land dense and conflict 4-way).
// Partial tiles: rows beyond the matrix are never copied and stay at the zeros written once before the first copy.
// Serves plain, symmetric-A and batched products of any offset and size (the producer picks its copy path per tile and
// zero-fills a k tail); split-K calls and products with m <= 64 stay on the LDGSTS kernel.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int BULK_CONSUMER_WARPS = 8;
constexpr int BULK_THREADS = 32 * (BULK_CONSUMER_WARPS + 1);

template <int BM, int BN, int WM, int WN, int BK, int STAGES, bool BATCHED>
__global__ void __launch_bounds__(BULK_THREADS, 1) gemm_bulk_kernel(GemmP p0, const GemmP* __restrict__ batch, int flags,
                                                                    int tri_keep, int splitk, double* __restrict__ ws) {
  extern __shared__ __align__(16) double smem[];
  GemmP p = p0;
  if (BATCHED) p = batch[blockIdx.z];
  constexpr int LDK = BK + 4;
  constexpr int A_TILE = BM * LDK;  // >= BK*(BM+4)
  constexpr int B_TILE = BN * LDK;
  constexpr int MI = WM / 8, NI = WN / 8;
  constexpr int WARPS_M = BM / WM;
  static_assert((BM / WM) * (BN / WN) == BULK_CONSUMER_WARPS, "8 consumer warps");
  __shared__ unsigned long long full_bar[STAGES], empty_bar[STAGES];

  // L2-aware tile order (GEMM_RASTER, set by gemm() for deep-k products): CTAs are dispatched in (x fastest, then y)
  // order and ~one per SM is resident, so with the plain mapping the resident set is a strip of ALL row tiles x 2-3
  // column tiles and every strip re-reads the whole of A from DRAM.  Walking the grid in groups of RASTER row tiles,
  // column by column inside a group, makes the resident set a RASTER x ~12 block of tiles.  Measured at 8192^3 (round 2,
  // profiles/r02_gemm_raster_8192_ncu.txt): DRAM read 15.8 -> 6.5 GB, same time (HBM was at 8 % of its peak either way);
  // rank-128 / rank-512 updates, whose traffic is C, LOSE 2 % with it and keep the plain order.
  constexpr int RASTER = 12;
  int tile_m = blockIdx.x, tile_n = blockIdx.y;
  if (flags & GEMM_RASTER) gemm_raster_tile(blockIdx.x + gridDim.x * blockIdx.y, gridDim.x, gridDim.y, RASTER, &tile_m, &tile_n);
  const int m0 = tile_m * BM, n0 = tile_n * BN;
  if (m0 >= p.m || n0 >= p.n) return;
  if (tri_keep >= 0 && n0 - (m0 + BM - 1) >= tri_keep) return;
  // k-tiles [kt0, kt0 + nk) of this CTA (a last partial k-tile is zero-filled by the producer); with split-K
  // (blockIdx.z = the split) the raw partial product goes to slice z of the workspace and splitk_reduce_kernel
  // applies alpha and beta
  int kt0 = 0, nk = (p.k + BK - 1) / BK;
  if (!BATCHED && splitk > 1) {
    const int per = (nk + splitk - 1) / splitk;
    kt0 = blockIdx.z * per;
    nk = max(0, min(nk, kt0 + per) - kt0);
    p.C = ws + (i64)blockIdx.z * p.m * p.n;
    p.ldc = p.m;
    p.alpha = 1.0;
    p.beta = 0.0;
  }
  const bool ta = flags & GEMM_TA, tb = flags & GEMM_TB, syma = flags & GEMM_SYMA;
  const bool b_kmajor = !tb;
  double* sA = smem;
  double* sB = smem + STAGES * A_TILE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rowsA = min(BM, p.m - m0), rowsB = min(BN, p.n - n0);
  // symmetric A (one triangle + a 128 band stored): k-tiles beyond this tile's diagonal block are read transposed
  auto a_is_kmajor = [&](int kt) -> bool { return syma ? (kt * BK) >= m0 + BM : ta; };

  // K-major operand tiles (rows of only BK doubles) are NOT worth a bulk copy each -- the copies of a warp are issued
  // one lane at a time, 128 of them per tile took longer than the tile's DMMAs (measured: 17 TF) -- so the producer
  // warp moves those with LDGSTS (16 bytes per lane, the same instruction count as the 8-warp kernel spends in total)
  // and signals them with cp.async.mbarrier.arrive.noinc: one arrival per lane on top of the expect_tx arrival.
  // (Every lane of the producer arrives once per stage, whether or not it issued LDGSTS copies: 1 + 32 arrivals.)
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < STAGES; ++st) {
      mbar_init(&full_bar[st], 33);
      mbar_init(&empty_bar[st], BULK_CONSUMER_WARPS);
    }
  }
  if (rowsA < BM || rowsB < BN) {  // partial tile: the rows no copy ever touches must read as zeros
    for (int e = tid; e < STAGES * (A_TILE + B_TILE); e += BULK_THREADS) smem[e] = 0.0;
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();

  if (warp == BULK_CONSUMER_WARPS) {
    // ---------------- producer warp
    // Copy paths per operand tile, chosen per CTA (all fill the same padded layouts):
    //   MN-major, 16-byte aligned, even row count : one bulk copy (TMA) per k-row               [the fast path]
    //   K-major, 16-byte aligned                  : LDGSTS.128, zero-filled past the end of k
    //   anything else (odd offsets / sizes of the D&C merge products, k tails of MN-major tiles): LDGSTS.64 with
    //   zero fill, still from this one warp.
    constexpr int CPR = BK / 2;        // 16-byte chunks per K-major row
    constexpr int RPP = 32 / CPR;      // rows covered by one LDGSTS.128 of the warp
    const int lrow = lane / CPR, lchk = (lane % CPR) * 2;
    const bool a_al = (((uintptr_t)p.A & 15) == 0) && ((p.lda & 1) == 0);
    const bool b_al = (((uintptr_t)p.B & 15) == 0) && ((p.ldb & 1) == 0);
    // one operand tile; base = element (row 0 of the tile, k = 0), ld = its leading dimension
    auto fill = [&](double* dst, const double* base, i64 ld, bool kmajor, bool al, int rows, int BMN, i64 k0, int kvalid,
                    unsigned long long* bar) -> unsigned {
      unsigned tx = 0;
      if (kmajor) {
        if (al) {
          const double* src = base + (i64)lrow * ld + k0 + lchk;
          unsigned d32 = smem_u32(dst + lrow * LDK + lchk);
          const int nb = max(0, min(16, (kvalid - lchk) * 8));
          for (int r = lrow; r < rows; r += RPP, src += (i64)RPP * ld, d32 += RPP * LDK * 8u) {
            if (nb == 16) cp_async16_full(d32, src);
            else asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d32), "l"(nb > 0 ? src : base), "r"(nb) : "memory");
          }
        } else {
          for (int e = lane; e < rows * BK; e += 32) {
            const int r = e / BK, kk = e % BK;
            cp_async8(dst + r * LDK + kk, kk < kvalid ? base + (i64)r * ld + k0 + kk : base, kk < kvalid ? 8 : 0);
          }
        }
      } else {
        const bool bulk = al && (rows % 2 == 0);
        if (bulk) {
          for (int kk = lane; kk < kvalid; kk += 32) bulk_g2s(dst + kk * (BMN + 4), base + (k0 + kk) * ld, (unsigned)rows * 8u, bar);
          tx = (unsigned)rows * (unsigned)kvalid * 8u;
        }
        // rows of k past the end (zero fill), or the whole tile when it cannot go by bulk copies
        for (int kk = bulk ? kvalid : 0; kk < BK; ++kk)
          for (int mm = lane; mm < rows; mm += 32)
            cp_async8(dst + kk * (BMN + 4) + mm, kk < kvalid ? base + (k0 + kk) * ld + mm : base, kk < kvalid ? 8 : 0);
      }
      return tx;
    };
    for (int it = 0; it < nk; ++it) {
      const int stage = it % STAGES;
      if (it >= STAGES) {
        mbar_wait(&empty_bar[stage], (unsigned)((it / STAGES - 1) & 1));
        // the buffer was read (and possibly zero-filled) through the generic proxy; bulk copies write it through the async one
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      }
      const bool akm = a_is_kmajor(kt0 + it);
      const i64 k0 = (i64)(kt0 + it) * BK;
      const int kvalid = (int)min((i64)BK, (i64)p.k - k0);
      // expect_tx must be posted with the byte count of the bulk copies of this stage (computed as fill does)
      const bool a_bulk = !akm && a_al && (rowsA % 2 == 0), b_bulk = !b_kmajor && b_al && (rowsB % 2 == 0);
      const unsigned bytes = (a_bulk ? (unsigned)rowsA : 0u) * kvalid * 8u + (b_bulk ? (unsigned)rowsB : 0u) * kvalid * 8u;
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[stage], bytes);
      __syncwarp();
      double* dA = sA + stage * A_TILE;
      double* dB = sB + stage * B_TILE;
      // K-major A is "rows of op(A)^T": element (m, k) at A[(m0 + m) * lda + k]; MN-major: A[k * lda + m0 + m]
      fill(dA, akm ? p.A + (i64)m0 * p.lda : p.A + m0, p.lda, akm, a_al, rowsA, BM, k0, kvalid, &full_bar[stage]);
      fill(dB, b_kmajor ? p.B + (i64)n0 * p.ldb : p.B + n0, p.ldb, b_kmajor, b_al, rowsB, BN, k0, kvalid, &full_bar[stage]);
      // this lane's arrival fires when all its LDGSTS above have landed
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&full_bar[stage])) : "memory");
    }
    return;
  }

  // ---------------- consumer warps
  const int wm = warp % WARPS_M, wn = warp / WARPS_M;
  const int lq = lane >> 2, lr = lane & 3;
  double alpha = p.alpha, beta = p.beta;
  double* C = p.C;
  const i64 ldc = p.ldc;
  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const bool cvec = (((uintptr_t)C & 15) == 0) && ((ldc & 1) == 0);
  const bool interior = cvec && rowsA == BM && rowsB == BN;
  if (beta != 0.0 && (alpha == 1.0 || alpha == -1.0)) {  // acc = alpha beta C (see gemm_kernel)
    const double sc = alpha * beta;
    if (interior) {
      const double* cbase = C + (i64)(n0 + wn * WN + lq) * ldc + (m0 + wm * WM + 2 * lr);
      double2 old[MI][NI];
#pragma unroll
      for (int j = 0; j < NI; ++j)
#pragma unroll
        for (int i = 0; i < MI; ++i) old[i][j] = __ldcg(reinterpret_cast<const double2*>(cbase + (i64)(j * 8) * ldc + i * 8));
#pragma unroll
      for (int j = 0; j < NI; ++j)
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          acc[i][j][0] = sc * old[i][j].x;
          acc[i][j][1] = sc * old[i][j].y;
        }
    } else {
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int col = n0 + wn * WN + j * 8 + lq;
        if (col >= p.n) continue;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const int row = m0 + wm * WM + i * 8 + 2 * lr;
          if (row >= p.m) continue;
          const double* cp = C + (i64)col * ldc + row;
          acc[i][j][0] = sc * cp[0];
          if (row + 1 < p.m) acc[i][j][1] = sc * cp[1];
        }
      }
    }
    beta = 0.0;
  }
  const int a_t_km = (wm * WM + lq) * LDK + lr, a_t_mn = (wm * WM + lq) + lr * (BM + 4);
  const int b_t = b_kmajor ? (wn * WN + lq) * LDK + lr : (wn * WN + lq) + lr * (BN + 4);
  for (int it = 0; it < nk; ++it) {
    const int stage = it % STAGES;
    mbar_wait(&full_bar[stage], (unsigned)((it / STAGES) & 1));
    const bool akm = a_is_kmajor(kt0 + it);
    const double* tA = sA + stage * A_TILE + (akm ? a_t_km : a_t_mn);
    const double* tB = sB + stage * B_TILE + b_t;
    if (akm) {
      if (b_kmajor) mma_ktile<MI, NI, BM, BN, BK, true, true>(acc, tA, tB);
      else mma_ktile<MI, NI, BM, BN, BK, true, false>(acc, tA, tB);
    } else {
      if (b_kmajor) mma_ktile<MI, NI, BM, BN, BK, false, true>(acc, tA, tB);
      else mma_ktile<MI, NI, BM, BN, BK, false, false>(acc, tA, tB);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
  }
  // epilogue
  if (interior && beta == 0.0) {
    double* cbase = C + (i64)(n0 + wn * WN + lq) * ldc + (m0 + wm * WM + 2 * lr);
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int i = 0; i < MI; ++i)
        *reinterpret_cast<double2*>(cbase + (i64)(j * 8) * ldc + i * 8) = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
    return;
  }
#pragma unroll
  for (int j = 0; j < NI; ++j) {
    const int col = n0 + wn * WN + j * 8 + lq;
    if (col >= p.n) continue;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int row = m0 + wm * WM + i * 8 + 2 * lr;
      if (row >= p.m) continue;
      double* cp = C + (i64)col * ldc + row;
      double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
      if (beta != 0.0) {
        v0 += beta * cp[0];
        if (row + 1 < p.m) v1 += beta * cp[1];
      }
      cp[0] = v0;
      if (row + 1 < p.m) cp[1] = v1;
    }
  }
}

template <int BM, int BN, int WM, int WN, int BK, int STAGES, bool BATCHED = false>
static int launch_bulk(Ctx* ctx, int flags, const GemmP& p, int tri_keep, const GemmP* d_batch = nullptr, int nb = 1,
                       int max_m = 0, int max_n = 0, int splitk = 1) {
  constexpr size_t smem = (size_t)STAGES * (BM + BN) * (BK + 4) * sizeof(double);
  static bool attr_dev[64] = {};
  bool& attr_set = attr_dev[ctx->device & 63];
  auto kern = gemm_bulk_kernel<BM, BN, WM, WN, BK, STAGES, BATCHED>;
  if (!attr_set) {
    EKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid(cdiv(BATCHED ? max_m : p.m, BM), cdiv(BATCHED ? max_n : p.n, BN), BATCHED ? nb : splitk);
  kern<<<grid, BULK_THREADS, smem, ctx->stream>>>(p, d_batch, flags, tri_keep, splitk, ctx->splitk_ws); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  return 0;
}

__global__ void splitk_reduce_kernel(const double* __restrict__ ws, int splitk, int m, int n, double* __restrict__ C,
                                     i64 ldc, double alpha, double beta, int tri_keep) {
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 tot = (i64)m * n;
  if (idx >= tot) return;
  int r = (int)(idx % m), c = (int)(idx / m);
  double s = 0.0;
  for (int z = 0; z < splitk; ++z) s += ws[(i64)z * tot + idx];
  double* cp = C + (i64)c * ldc + r;
  double v = alpha * s;
  if (beta != 0.0) v += beta * *cp;
  *cp = v;
}

template <int BM, int BN, int WM, int WN, bool BATCHED, int BK, int STAGES>
static int launch_cfg(Ctx* ctx, int flags, const GemmP& p, const GemmP* d_batch, int nb, int max_m, int max_n,
                      int tri_keep, int splitk) {
  constexpr size_t smem = (size_t)STAGES * (BM + BN) * (BK + 4) * sizeof(double);
  static_assert(smem <= 227 * 1024, "shared memory per CTA");
  static bool attr_dev[64] = {};  // per device: the attribute belongs to the device's context
  bool& attr_set = attr_dev[ctx->device & 63];
  auto kern = gemm_kernel<BM, BN, WM, WN, BATCHED, BK, STAGES>;
  if (!attr_set) {
    EKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid(cdiv(max_m, BM), cdiv(max_n, BN), BATCHED ? nb : splitk);
  kern<<<grid, GEMM_THREADS, smem, ctx->stream>>>(p, d_batch, flags, tri_keep, splitk, ctx->splitk_ws); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());
  return 0;
}

// Tile shape by the problem shape, pipeline geometry by the k-depth each CTA actually runs through.
template <bool BATCHED>
static int launch_shape(Ctx* ctx, int flags, const GemmP& p, const GemmP* d_batch, int nb, int max_m, int max_n, int tri_keep,
                        int splitk, bool deep) {
  if (max_n <= 64)
    return deep ? launch_cfg<128, 64, 32, 32, BATCHED, 32, 3>(ctx, flags, p, d_batch, nb, max_m, max_n, tri_keep, splitk)
                : launch_cfg<128, 64, 32, 32, BATCHED, 16, 4>(ctx, flags, p, d_batch, nb, max_m, max_n, tri_keep, splitk);
  if (!BATCHED && max_m <= 64)
    return deep ? launch_cfg<64, 128, 32, 32, BATCHED, 32, 3>(ctx, flags, p, d_batch, nb, max_m, max_n, tri_keep, splitk)
                : launch_cfg<64, 128, 32, 32, BATCHED, 16, 4>(ctx, flags, p, d_batch, nb, max_m, max_n, tri_keep, splitk);
  return deep ? launch_cfg<128, 128, 64, 32, BATCHED, 32, 3>(ctx, flags, p, d_batch, nb, max_m, max_n, tri_keep, splitk)
              : launch_cfg<128, 128, 64, 32, BATCHED, 16, 4>(ctx, flags, p, d_batch, nb, max_m, max_n, tri_keep, splitk);
}

int gemm(Ctx* ctx, int flags, const GemmP& p, int tri_keep, int splitk) {
  if (p.m <= 0 || p.n <= 0) return 0;
  const bool bulk_shape = ctx->gemm_bulk != 0 && p.m > 64 && p.k >= 64;
  // products the TMA-fed kernel takes pick their own split (option "gemm_autosplit", default on); otherwise the
  // caller's request stands, except that with a 128-row tile per SM the unsplit TMA-fed kernel beats a split one
  if (bulk_shape && tri_keep < 0 && ctx->gemm_autosplit != 0) splitk = gemm_autosplit_factor(p.m, p.n, p.k, ctx->num_sms);  // layout.h
  else if (splitk > 1 && bulk_shape && cdiv(p.m, 128) * (p.n > 64 ? cdiv(p.n, 128) : 1) >= ctx->num_sms) splitk = 1;
  if (splitk > 1) {
    int nkt = cdiv(p.k, 32);
    if (splitk > nkt) splitk = nkt > 0 ? nkt : 1;
    size_t need = (size_t)splitk * p.m * p.n * sizeof(double);
    if (need > ctx->splitk_ws_bytes) {
      if (ctx->splitk_ws) ctx_free(ctx, ctx->splitk_ws);
      ctx->splitk_ws = nullptr;
      ctx->splitk_ws_bytes = 0;
      size_t want = need < ((size_t)64 << 20) ? ((size_t)64 << 20) : need;
      EKB_TRY(ctx_alloc(ctx, (void**)&ctx->splitk_ws, want));
      ctx->splitk_ws_bytes = want;
    }
  }
  if (splitk < 1) splitk = 1;
  int rc;
  {
    // algorithmic FLOPs: 2 k per computed element; triangular updates count the kept triangle only
    double elems = (double)p.m * p.n;
    if (tri_keep >= 0) {
      const double mn = p.m < p.n ? p.m : p.n;
      elems = (double)p.m * p.n - 0.5 * mn * (mn - 1.0);
    }
    EKB_TRY(prof_begin(ctx, PROF_GEMM, 2.0 * p.k * elems));
    if (ctx->profile_gemm) ctx->prof_shape.back() = Ctx::ProfShape{p.m, p.n, p.k, flags, tri_keep, splitk};
  }
  // k-depth per CTA (after split-K) decides the pipeline geometry: >= 32 k-tiles of 32 amortise the longer prologue
  const bool deep = p.k / splitk >= 1024;
  // the TMA-fed warp-specialised kernel takes the big-tile products it supports (option "gemm_bulk", default on)
  const bool bulk_ok = bulk_shape && (splitk == 1 || ctx->gemm_autosplit != 0);
  if (bulk_ok && p.k >= 1024) flags |= GEMM_RASTER;
  if (bulk_ok && p.n > 64)
    rc = deep ? launch_bulk<128, 128, 64, 32, 32, 3>(ctx, flags, p, tri_keep, nullptr, 1, 0, 0, splitk)
              : launch_bulk<128, 128, 64, 32, 16, 4>(ctx, flags, p, tri_keep, nullptr, 1, 0, 0, splitk);
  else if (bulk_ok)
    rc = deep ? launch_bulk<128, 64, 32, 32, 32, 3>(ctx, flags, p, tri_keep, nullptr, 1, 0, 0, splitk)
              : launch_bulk<128, 64, 32, 32, 16, 4>(ctx, flags, p, tri_keep, nullptr, 1, 0, 0, splitk);
  else
    rc = launch_shape<false>(ctx, flags, p, nullptr, 1, p.m, p.n, tri_keep, splitk, deep);
  if (rc) return rc;
  if (splitk > 1) {
    i64 tot = (i64)p.m * p.n;
    splitk_reduce_kernel<<<cdiv(tot, 256), 256, 0, ctx->stream>>>(ctx->splitk_ws, splitk, p.m, p.n, p.C, p.ldc,
                                                                 p.alpha, p.beta, tri_keep);
    EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
  }
  EKB_TRY(prof_end(ctx));
  return 0;
}

int prof_begin(Ctx* ctx, int family, double work) {
  if (!ctx->profile_gemm) return 0;
  if (ctx->prof_used + 2 > ctx->prof_events.size()) {
    for (int q = 0; q < 2048; ++q) {
      cudaEvent_t ev;
      EKB_CUDA(cudaEventCreate(&ev));
      ctx->prof_events.push_back(ev);
    }
  }
  ctx->prof_flops.push_back(work);
  ctx->prof_family.push_back(family);
  ctx->prof_stage.push_back(ctx->cur_stage);
  ctx->prof_shape.push_back(Ctx::ProfShape{0, 0, 0, 0, 0, 0});
  EKB_CUDA(cudaEventRecord(ctx->prof_events[ctx->prof_used], ctx->stream));
  return 0;
}

int prof_end(Ctx* ctx) {
  if (!ctx->profile_gemm) return 0;
  EKB_CUDA(cudaEventRecord(ctx->prof_events[ctx->prof_used + 1], ctx->stream));
  ctx->prof_used += 2;
  return 0;
}

int profile_collect(Ctx* ctx, double* seconds, double* work, long long* launches) {
  EKB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int f = 0; f < PROF_FAMILIES; ++f) { seconds[f] = 0.0; work[f] = 0.0; launches[f] = 0; }
  ctx->prof_table.clear();
  // EKB200_GEMM_TRACE=<file>: one line per engine GEMM (stage m n k flags tri_keep splitk ms), appended -- the
  // per-shape view behind the (stage, family) table (scripts/gemm_trace_summary.py)
  const char* trace_path = getenv("EKB200_GEMM_TRACE");
  FILE* trace = (trace_path && *trace_path) ? fopen(trace_path, "a") : nullptr;
  struct Closer { FILE* f; ~Closer() { if (f) fclose(f); } } closer{trace};
  for (size_t q = 0; q + 1 < ctx->prof_used; q += 2) {
    float ms = 0.f;
    EKB_CUDA(cudaEventElapsedTime(&ms, ctx->prof_events[q], ctx->prof_events[q + 1]));
    const int f = ctx->prof_family[q / 2];
    if (trace && f == PROF_GEMM && q / 2 < ctx->prof_shape.size()) {
      const Ctx::ProfShape& sh = ctx->prof_shape[q / 2];
      fprintf(trace, "%s %d %d %d %d %d %d %.6f\n", ctx->prof_stage[q / 2] ? ctx->prof_stage[q / 2] : "-", sh.m, sh.n, sh.k,
              sh.flags, sh.tri, sh.splitk, (double)ms);
    }
    seconds[f] += ms * 1e-3;
    work[f] += ctx->prof_flops[q / 2];
    launches[f] += 1;
    const char* st = ctx->prof_stage[q / 2];
    Ctx::ProfRow* row = nullptr;
    for (auto& r : ctx->prof_table)
      if (r.family == f && r.stage == st) { row = &r; break; }
    if (!row) {
      ctx->prof_table.push_back(Ctx::ProfRow{st, f, 0.0, 0.0, 0});
      row = &ctx->prof_table.back();
    }
    row->seconds += ms * 1e-3;
    row->work += ctx->prof_flops[q / 2];
    row->launches += 1;
  }
  ctx->prof_used = 0;
  ctx->prof_flops.clear();
  ctx->prof_family.clear();
  ctx->prof_stage.clear();
  ctx->prof_shape.clear();
  return 0;
}

// GEMM-family view of profile_collect (kept for ekb200_gemm_profile): resets ALL families.
int gemm_profile_collect(Ctx* ctx, double* seconds, double* flops, long long* launches) {
  double s[PROF_FAMILIES], w[PROF_FAMILIES];
  long long l[PROF_FAMILIES];
  EKB_TRY(profile_collect(ctx, s, w, l));
  *seconds = s[PROF_GEMM];
  *flops = w[PROF_GEMM];
  *launches = l[PROF_GEMM];
  return 0;
}

int gemm_batched(Ctx* ctx, int flags, const GemmP* d_batch, int nb, int max_m, int max_n, int k_hint) {
  if (nb <= 0 || max_m <= 0 || max_n <= 0) return 0;
  GemmP dummy = {};
  EKB_TRY(prof_begin(ctx, PROF_GEMM_BATCHED, 0.0));  // FLOPs after deflation are only known on the device
  int rc;
  const bool deep = k_hint >= 1024;
  if (ctx->gemm_bulk != 0 && max_m > 64 && k_hint >= 64 && !(flags & GEMM_SYMA)) {
    // the D&C merge products: any offset, any size -- the producer warp picks its copy path per tile
    if (max_n > 64)
      rc = deep ? launch_bulk<128, 128, 64, 32, 32, 3, true>(ctx, flags, dummy, -1, d_batch, nb, max_m, max_n)
                : launch_bulk<128, 128, 64, 32, 16, 4, true>(ctx, flags, dummy, -1, d_batch, nb, max_m, max_n);
    else
      rc = deep ? launch_bulk<128, 64, 32, 32, 32, 3, true>(ctx, flags, dummy, -1, d_batch, nb, max_m, max_n)
                : launch_bulk<128, 64, 32, 32, 16, 4, true>(ctx, flags, dummy, -1, d_batch, nb, max_m, max_n);
  } else {
    rc = launch_shape<true>(ctx, flags, dummy, d_batch, nb, max_m, max_n, -1, 1, deep);
  }
  if (rc) return rc;
  return prof_end(ctx);
}

}  // namespace ekb
