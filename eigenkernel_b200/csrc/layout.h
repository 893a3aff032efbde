// Host-only layout arithmetic shared by the CUDA library and the host-check library: how the columns of an
// n x ncols matrix are dealt to the P ranks (one B200 each).  B200 counterpart of numroc / get_local_cols
// (reference src/distribute_matrix.f90:81-89): a 1 x P grid with ONE contiguous column block per rank.
#pragma once
#include <vector>

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

// Block-column-cyclic ownership used by the sharded dense-to-band reduction: block column c (width cb) of the
// matrix belongs to rank c mod P.
inline int block_owner(long long block, int nranks) { return (int)(block % nranks); }

}  // namespace ekb
