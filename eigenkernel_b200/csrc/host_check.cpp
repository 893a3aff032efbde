// Host-side check library: compiles the __host__ __device__ numerics of the CUDA path for the CPU so the
// `-m "not gpu"` tests can exercise them (against LAPACK dlaed4) without a GPU.  Test infrastructure only;
// nothing on the solve path links this.
#include "layout.h"
#include "secular.cuh"

extern "C" int ekb200_host_secular(int k, const double* d, const double* z, double rho, double* lam, int* orig,
                                   double* tau, int* iters) {
  for (int j = 0; j < k; ++j) {
    int K, it;
    double t;
    ekb::secular_root(k, j, d, z, rho, &K, &t, &it);
    lam[j] = d[K] + t;
    orig[j] = K;
    tau[j] = t;
    iters[j] = it;
  }
  return 0;
}

// bounds: nranks + 1 entries
extern "C" int ekb200_host_slab_bounds(long long ncols, int nranks, int gran, long long* bounds) {
  std::vector<long long> b;
  ekb::slab_bounds(ncols, nranks, gran, b);
  for (int r = 0; r <= nranks; ++r) bounds[r] = b[r];
  return 0;
}
