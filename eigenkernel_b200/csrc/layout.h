// Host-only layout arithmetic shared by the CUDA library and the host-check library: how the columns of an
// n x ncols matrix are dealt to the P ranks (one B200 each).  B200 counterpart of numroc / get_local_cols
// (reference src/distribute_matrix.f90:81-89): a 1 x P grid with ONE contiguous column block per rank.
#pragma once
#include <cmath>
#include <vector>

#ifdef __CUDACC__
#define EKB_LAYOUT_HD __host__ __device__ __forceinline__
#else
#define EKB_LAYOUT_HD inline
#endif

namespace ekb {

// Slab r = columns [bounds[r], bounds[r+1]); boundaries are multiples of `gran` (trailing slabs may be short
// or empty when ncols is small).
inline void slab_bounds(long long ncols, int nranks, int gran, std::vector<long long>& bounds) {
  bounds.assign(nranks + 1, 0);
  long long chunk = (ncols + nranks - 1) / nranks;
  if (chunk < 1) chunk = 1;
  chunk = (chunk + gran - 1) / gran * gran;
  for (int r = 0; r <= nranks; ++r) {
    long long v = (long long)r * chunk;
    bounds[r] = v < ncols ? v : ncols;
  }
}

// ---- the caller's ScaLAPACK view of the result (rank-per-GPU Fortran mode).  setup_distributed_matrix
// (reference src/distribute_matrix.f90:92-148) allocates blacs%Vectors for a block-cyclic descriptor whose block size it
// CLAMPS to max(min(rows / nprow, cols / npcol), 1) (:114-120) "so that no process is empty"; on the 1 x P grid of the B200
// solvers that is floor(ncols / P), which differs from the library's own slab width (ceil(ncols / P) rounded up to 128)
// unless 128 P divides ncols.  With option "out_block" = NB the host-pointer entry points therefore deliver each rank's
// piece of the 1 x P block-cyclic distribution with block size NB (numroc columns; blocks r, r + P, ... of rank r).
inline long long reference_block_clamp(long long rows, long long cols, long long want, int nprow, int npcol) {
  long long a = rows / nprow, b = cols / npcol;
  long long mx = a < b ? a : b;
  if (mx < 1) mx = 1;
  return want > mx ? mx : want;
}
// ScaLAPACK NUMROC(n, nb, iproc, isrcproc = 0, nprocs)
inline long long numroc0(long long n, long long nb, int iproc, int nprocs) {
  const long long nblocks = n / nb;
  long long loc = (nblocks / nprocs) * nb;
  const long long extra = nblocks % nprocs;
  if (iproc < extra) loc += nb;
  else if (iproc == extra) loc += n % nb;
  return loc;
}
// global column of local column lc of rank r (block size nb, P ranks, source rank 0)
inline long long cyclic_global_col0(long long lc, long long nb, int nprocs, int r) {
  return ((lc / nb) * nprocs + r) * nb + lc % nb;
}

// ---- GEMM engine, host-side decisions (gemm.cu) ------------------------------------------------------------------
// Split-K factor for a product the TMA-fed kernel runs: one CTA per SM, so a grid of T tiles costs ceil(T / SMs) full
// rounds however empty the last one is (250 tiles of the m x 64 panel product at m = 32000: 1.69 -> 2 rounds).  Splitting
// k by s makes the rounds s times shorter and the count ceil(T s / SMs); the price is the partial-sum pass.  Modelled
// in seconds (236 GFLOP/s per SM, 4 TB/s for the partial sums), smallest s within 3 % of the best; never leaves a CTA
// less than 512 of k, never asks for more than 512 MB of partial sums.
inline int gemm_autosplit_factor(long long m, long long n, long long k, int num_sms) {
  const int bn = n > 64 ? 128 : 64;
  const double tiles = (double)((m + 127) / 128) * (double)((n + bn - 1) / bn);
  const double per_k = 2.0 * 128 * bn / 236e9;  // seconds per unit of k per CTA
  double best_t = 0.0;
  int best = 1;
  for (int s = 1; s <= 8; ++s) {
    if (s > 1 && (k / s < 512 || (double)s * (double)m * (double)n * 8.0 > 512e6)) break;
    const double rounds = std::ceil(tiles * s / num_sms);
    double t = rounds * per_k * ((double)k / s);
    if (s > 1) t += (double)(s + 2) * (double)m * (double)n * 8.0 / 4e12 + 3e-6;
    if (s == 1 || t < 0.97 * best_t) { best_t = t; best = s; }
  }
  return best;
}

// L2-aware tile order: CTA number pid (dispatch order: x fastest) -> tile (tm, tn) of a gm x gn grid, walking the grid
// in groups of `raster` row tiles, column by column inside a group.  A bijection of [0, gm gn) onto the grid.
EKB_LAYOUT_HD void gemm_raster_tile(int pid, int gm, int gn, int raster, int* tm, int* tn) {
  const int per = raster * gn;
  const int grp = pid / per, first = grp * raster;
  const int gsz = (gm - first) < raster ? (gm - first) : raster;
  const int rem = pid - grp * per;
  *tm = first + rem % gsz;
  *tn = rem / gsz;
}

// Block-column-cyclic ownership used by the sharded dense-to-band reduction: block column c (width cb) of the
// matrix belongs to rank c mod P.
inline int block_owner(long long block, int nranks) { return (int)(block % nranks); }

}  // namespace ekb
