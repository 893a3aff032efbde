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

// Block-column-cyclic ownership used by the sharded dense-to-band reduction: block column c (width cb) of the
// matrix belongs to rank c mod P.
inline int block_owner(long long block, int nranks) { return (int)(block % nranks); }

}  // namespace ekb
