// Divide-and-conquer eigensolver for the symmetric tridiagonal matrix: replaces pdstedc('I'),
// reference src/solver_scalapack_all.f90:96-98 (eigenvalues ascending in `values`, Z = eigenvectors of T).
//
// Cuppen's method with Gu-Eisenstat stabilisation, organised for the GPU:
//   * the tree is cut to uniform depth; leaves (<= 32) are solved by implicit-shift QL, one warp per leaf;
//   * each level is processed for all its nodes at once by batched kernels, with NO host synchronisation:
//     z-vector + sort, deflation scan (sequential by nature: one warp per node, lane 0 walks the sorted
//     poles while the warp prefetches), Givens rotations of deflated pairs, secular equation ONE ROOT PER
//     THREAD (middle-way rational iteration on the offset from the nearest pole, bracketed), Loewner
//     recomputation of z, eigenvector matrix U, and the merge products Q_new = [Q1 0; 0 Q2] P U as
//     batched GEMMs on the DMMA engine that skip the structural zero blocks (LAPACK's column types 1/2/3).
//   * per-root (origin, tau) representation: d_i - lambda_j = (d_i - d_origin) - tau is evaluated on the fly
//     to high relative accuracy, so no k x k difference matrix is stored.
#include <algorithm>

#include "common.cuh"
#include "secular.cuh"

namespace ekb {

constexpr int DC_LEAF = 32;

struct DcNode {
  int off, n1, sz;
  int k, k1, k2, k3, nrot, ndefl, pad;
  double rho, tol;
};

struct DcWork {
  // length-n arrays indexed by (node offset + local index)
  double *z, *dlam, *zact, *defl_val, *rot_c, *rot_s, *lam, *tau, *zhat, *vals;
  int *indx, *act_col, *act_type, *gpos, *gcol, *defl_col, *rot_a, *rot_b, *orig, *pos;
};

// ------------------------------------------------------------------------------------------ leaves
__global__ void dc_cut_kernel(double* __restrict__ d, const double* __restrict__ e, i64 n, int nleaf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (i >= nleaf) return;
  i64 p = (i64)i * n / nleaf;
  double r = fabs(e[p - 1]);
  d[p - 1] -= r;
  d[p] -= r;
}

// One warp per leaf: implicit QL (EISPACK tql2 scheme), eigenvalues sorted ascending.
// Q block written at Q[off.., off..]; D[off..] sorted eigenvalues.
__global__ void __launch_bounds__(128) dc_leaf_kernel(const double* __restrict__ d, const double* __restrict__ e, i64 n,
                                                      int nleaf, double* __restrict__ D, double* __restrict__ Q, i64 ldq,
                                                      int* __restrict__ fail) {
  __shared__ double sz_[4][DC_LEAF][DC_LEAF + 1];
  __shared__ double sd[4][DC_LEAF + 1], se[4][DC_LEAF + 1];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int leaf = blockIdx.x * 4 + w;
  if (leaf >= nleaf) return;
  const i64 off = (i64)leaf * n / nleaf;
  const int m = (int)((i64)(leaf + 1) * n / nleaf - off);
  double(*z)[DC_LEAF + 1] = sz_[w];
  double* dd = sd[w];
  double* ee = se[w];
  for (int c = 0; c < DC_LEAF; ++c) z[lane][c] = (lane == c) ? 1.0 : 0.0;
  if (lane < m) dd[lane] = d[off + lane];
  ee[lane] = (lane < m - 1) ? e[off + lane] : 0.0;
  __syncwarp();
  // every lane runs the scalar recurrence redundantly on shared d/e (identical values), lane = row of Z
  bool bad = false;
  for (int l = 0; l < m; ++l) {
    int iter = 0;
    while (true) {
      int mm = l;
      for (; mm < m - 1; ++mm) {
        double s = fabs(dd[mm]) + fabs(dd[mm + 1]);
        if (fabs(ee[mm]) <= DC_EPS * s) break;
      }
      if (mm == l) break;
      if (++iter > 60) { bad = true; break; }
      double g = (dd[l + 1] - dd[l]) / (2.0 * ee[l]);
      double r = hypot(g, 1.0);
      g = dd[mm] - dd[l] + ee[l] / (g + copysign(r, g));
      double s = 1.0, c = 1.0, p = 0.0;
      int i = mm - 1;
      bool under = false;
      __syncwarp();
      for (; i >= l; --i) {
        double f = s * ee[i], b = c * ee[i];
        r = hypot(f, g);
        double di1 = dd[i + 1], di = dd[i];
        __syncwarp();
        if (lane == 0) ee[i + 1] = r;
        if (r == 0.0) {
          if (lane == 0) { dd[i + 1] = di1 - p; ee[mm] = 0.0; }
          under = true;
          break;
        }
        s = f / r;
        c = g / r;
        g = di1 - p;
        r = (di - g) * s + 2.0 * c * b;
        p = s * r;
        if (lane == 0) dd[i + 1] = g + p;
        g = c * r - b;
        // rotate columns i, i+1 of Z (lane = row)
        double f2 = z[lane][i + 1], zi = z[lane][i];
        z[lane][i + 1] = s * zi + c * f2;
        z[lane][i] = c * zi - s * f2;
      }
      __syncwarp();
      if (under) continue;
      if (lane == 0) { dd[l] -= p; ee[l] = g; ee[mm] = 0.0; }
      __syncwarp();
    }
    if (bad) break;
  }
  __syncwarp();
  if (bad && lane == 0) atomicAdd(fail, 1);
  // selection sort ascending (m <= 32): rank by counting
  double my = (lane < m) ? dd[lane] : 0.0;
  int rank = 0;
  for (int j = 0; j < m; ++j) {
    double o = dd[j];
    rank += (o < my || (o == my && j < lane)) ? 1 : 0;
  }
  if (lane < m) {
    D[off + rank] = my;
    for (int r = 0; r < m; ++r) Q[(off + rank) * ldq + off + r] = z[r][lane];
  }
}

// ------------------------------------------------------------------------------------------ per-level kernels
// z vector, rho, sorted permutation (merge of the two sorted child spectra), tolerance.
__global__ void __launch_bounds__(256) dc_setup_kernel(DcNode* __restrict__ nodes, const double* __restrict__ D,
                                                       const double* __restrict__ e, const double* __restrict__ Q, i64 ldq,
                                                       DcWork wk) {
  DcNode nd = nodes[blockIdx.x];
  const int off = nd.off, n1 = nd.n1, sz = nd.sz;
  const double rho_raw = e[off + n1 - 1];
  const double sgn = rho_raw < 0.0 ? -1.0 : 1.0;
  const double isq2 = 0.70710678118654752440;
  __shared__ double sred[2][8];
  double dmax = 0.0, zmax = 0.0;
  for (int i = threadIdx.x; i < sz; i += blockDim.x) {
    double zi;
    if (i < n1) zi = Q[(i64)(off + i) * ldq + off + n1 - 1] * isq2;
    else zi = Q[(i64)(off + i) * ldq + off + n1] * isq2 * sgn;
    wk.z[off + i] = zi;
    const double x = D[off + i];
    dmax = fmax(dmax, fabs(x));
    zmax = fmax(zmax, fabs(zi));
    // rank in the merged order
    int rank;
    if (i < n1) {
      int lo = 0, hi = sz - n1;  // count right elements < x
      const double* R = D + off + n1;
      while (lo < hi) { int mid = (lo + hi) >> 1; if (R[mid] < x) lo = mid + 1; else hi = mid; }
      rank = i + lo;
    } else {
      int lo = 0, hi = n1;  // count left elements <= x
      const double* L = D + off;
      while (lo < hi) { int mid = (lo + hi) >> 1; if (L[mid] <= x) lo = mid + 1; else hi = mid; }
      rank = (i - n1) + lo;
    }
    wk.indx[off + rank] = i;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    zmax = fmax(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
  }
  if ((threadIdx.x & 31) == 0) { sred[0][threadIdx.x >> 5] = dmax; sred[1][threadIdx.x >> 5] = zmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 0; q < 8; ++q) { dmax = fmax(dmax, sred[0][q]); zmax = fmax(zmax, sred[1][q]); }
    nodes[blockIdx.x].rho = 2.0 * fabs(rho_raw);
    nodes[blockIdx.x].tol = 8.0 * DC_EPS * fmax(dmax, zmax);
  }
}

// Deflation scan: one warp per node.  The warp stages 32 sorted poles at a time in shared memory, lane 0
// runs the (inherently sequential) dlaed2-style scan.
__global__ void __launch_bounds__(32) dc_deflate_kernel(DcNode* __restrict__ nodes, const double* __restrict__ D, DcWork wk) {
  DcNode nd = nodes[blockIdx.x];
  const int off = nd.off, n1 = nd.n1, sz = nd.sz;
  const double rho = nd.rho, tol = nd.tol;
  __shared__ double sd[32], sz2[32];
  __shared__ int sc[32];
  const int lane = threadIdx.x;
  int k = 0, ndefl = 0, nrot = 0, k1 = 0, k2 = 0, k3 = 0;
  bool have_p = false;
  int colp = 0, tp = 0;
  double dp = 0.0, zp = 0.0;
  for (int base = 0; base < sz; base += 32) {
    const int jj = base + lane;
    if (jj < sz) {
      int col = wk.indx[off + jj];
      sc[lane] = col;
      sd[lane] = D[off + col];
      sz2[lane] = wk.z[off + col];
    }
    __syncwarp();
    if (lane == 0) {
      const int cnt = min(32, sz - base);
      for (int q = 0; q < cnt; ++q) {
        int col = sc[q];
        double dj = sd[q], zj = sz2[q];
        int tj = col < n1 ? 1 : 3;
        if (rho * fabs(zj) <= tol) {
          wk.defl_col[off + ndefl] = col;
          wk.defl_val[off + ndefl] = dj;
          ++ndefl;
        } else if (!have_p) {
          colp = col; dp = dj; zp = zj; tp = tj; have_p = true;
        } else {
          double s = zp, c = zj;
          const double tau = hypot(c, s);
          const double t = dj - dp;
          c /= tau;
          s = -s / tau;
          if (fabs(t * c * s) <= tol) {
            zj = tau;
            wk.rot_a[off + nrot] = colp; wk.rot_b[off + nrot] = col;
            wk.rot_c[off + nrot] = c; wk.rot_s[off + nrot] = s;
            ++nrot;
            if (tj != tp) tj = 2;
            const double t2 = dp * c * c + dj * s * s;
            dj = dp * s * s + dj * c * c;
            wk.defl_col[off + ndefl] = colp;
            wk.defl_val[off + ndefl] = t2;
            ++ndefl;
          } else {
            wk.act_col[off + k] = colp; wk.dlam[off + k] = dp; wk.zact[off + k] = zp; wk.act_type[off + k] = tp;
            k1 += (tp == 1); k2 += (tp == 2); k3 += (tp == 3);
            ++k;
          }
          colp = col; dp = dj; zp = zj; tp = tj;
        }
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    if (have_p) {
      wk.act_col[off + k] = colp; wk.dlam[off + k] = dp; wk.zact[off + k] = zp; wk.act_type[off + k] = tp;
      k1 += (tp == 1); k2 += (tp == 2); k3 += (tp == 3);
      ++k;
    }
    nodes[blockIdx.x].k = k; nodes[blockIdx.x].k1 = k1; nodes[blockIdx.x].k2 = k2; nodes[blockIdx.x].k3 = k3;
    nodes[blockIdx.x].nrot = nrot; nodes[blockIdx.x].ndefl = ndefl;
  }
  __syncwarp();
  // grouped positions: type 1 first, then 2, then 3 (stable).  Warp-parallel ballot scan.
  k = __shfl_sync(0xffffffffu, k, 0);
  k1 = __shfl_sync(0xffffffffu, k1, 0);
  k2 = __shfl_sync(0xffffffffu, k2, 0);
  __threadfence_block();
  int c1 = 0, c2 = k1, c3 = k1 + k2;
  for (int base = 0; base < k; base += 32) {
    const int q = base + lane;
    const int t = (q < k) ? wk.act_type[off + q] : 0;
    const unsigned b1 = __ballot_sync(0xffffffffu, t == 1), b2 = __ballot_sync(0xffffffffu, t == 2),
                   b3 = __ballot_sync(0xffffffffu, t == 3);
    const unsigned lt = (1u << lane) - 1u;
    if (q < k) {
      int g = (t == 1) ? c1 + __popc(b1 & lt) : (t == 2) ? c2 + __popc(b2 & lt) : c3 + __popc(b3 & lt);
      wk.gpos[off + q] = g;
      wk.gcol[off + g] = wk.act_col[off + q];
    }
    c1 += __popc(b1); c2 += __popc(b2); c3 += __popc(b3);
  }
}

// The children only ever wrote their own diagonal blocks: clear the two off-diagonal blocks of every node so
// that whole-column operations (rotations, copies of deflated columns) see the structural zeros.
__global__ void __launch_bounds__(256) dc_zero_offdiag_kernel(const DcNode* __restrict__ nodes, double* __restrict__ Q, i64 ldq) {
  const DcNode nd = nodes[blockIdx.z];
  const int c = blockIdx.x;
  if (c >= nd.sz) return;
  const int r = blockIdx.y * blockDim.x + threadIdx.x;
  if (r >= nd.sz) return;
  if ((c < nd.n1) != (r < nd.n1)) Q[(i64)(nd.off + c) * ldq + nd.off + r] = 0.0;
}

// Apply the recorded Givens rotations to the columns of Q (thread per row, rotations in order).
__global__ void __launch_bounds__(256) dc_rotate_kernel(const DcNode* __restrict__ nodes, double* __restrict__ Q, i64 ldq,
                                                        DcWork wk) {
  const DcNode nd = nodes[blockIdx.y];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nd.sz || nd.nrot == 0) return;
  const int off = nd.off;
  double* row = Q + off + r;
  for (int q = 0; q < nd.nrot; ++q) {
    const int a = wk.rot_a[off + q], b = wk.rot_b[off + q];
    const double c = wk.rot_c[off + q], s = wk.rot_s[off + q];
    double x = row[(i64)(off + a) * ldq], y = row[(i64)(off + b) * ldq];
    row[(i64)(off + a) * ldq] = c * x + s * y;
    row[(i64)(off + b) * ldq] = c * y - s * x;
  }
}

// Secular equation, one root per thread (see secular.cuh).
__global__ void __launch_bounds__(128) dc_secular_kernel(const DcNode* __restrict__ nodes, DcWork wk) {
  const DcNode nd = nodes[blockIdx.y];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = nd.k;
  if (j >= k) return;
  const double* __restrict__ d = wk.dlam + nd.off;
  const double* __restrict__ z = wk.zact + nd.off;
  int K;
  double tau;
  secular_root(k, j, d, z, nd.rho, &K, &tau, nullptr);
  wk.orig[nd.off + j] = K;
  wk.tau[nd.off + j] = tau;
  wk.lam[nd.off + j] = d[K] + tau;
}

// Loewner / Gu-Eisenstat: zhat_i = sign(z_i) sqrt(| (lam_i - d_i) prod_{j != i} (lam_j - d_i)/(d_j - d_i) |).
// One CTA per pole i.
__global__ void __launch_bounds__(128) dc_lowner_kernel(const DcNode* __restrict__ nodes, DcWork wk) {
  const DcNode nd = nodes[blockIdx.y];
  const int i = blockIdx.x, k = nd.k;
  if (i >= k) return;
  const double* __restrict__ d = wk.dlam + nd.off;
  const double* __restrict__ tau = wk.tau + nd.off;
  const int* __restrict__ orig = wk.orig + nd.off;
  const double di = d[i];
  double prod = 1.0;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const double num = (d[orig[j]] - di) + tau[j];  // lam_j - d_i
    prod *= (j == i) ? num : num / (d[j] - di);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) prod *= __shfl_xor_sync(0xffffffffu, prod, o);
  __shared__ double sp[4];
  if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = prod;
  __syncthreads();
  if (threadIdx.x == 0) {
    prod = sp[0] * sp[1] * sp[2] * sp[3];
    wk.zhat[nd.off + i] = copysign(sqrt(fabs(prod)), wk.zact[nd.off + i]);
  }
}

// U(gpos(i), j) = zhat_i / (d_i - lam_j) / norm_j.  One CTA per root column j.
__global__ void __launch_bounds__(128) dc_buildU_kernel(const DcNode* __restrict__ nodes, DcWork wk, double* __restrict__ U,
                                                        i64 ldu) {
  const DcNode nd = nodes[blockIdx.y];
  const int j = blockIdx.x, k = nd.k;
  if (j >= k) return;
  const double* __restrict__ d = wk.dlam + nd.off;
  const double* __restrict__ zh = wk.zhat + nd.off;
  const int* __restrict__ gpos = wk.gpos + nd.off;
  const double dK = d[wk.orig[nd.off + j]], tj = wk.tau[nd.off + j];
  double ss = 0.0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const double q = zh[i] / ((d[i] - dK) - tj);
    ss += q * q;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  __shared__ double sp[4];
  if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = ss;
  __syncthreads();
  const double inv = 1.0 / sqrt(sp[0] + sp[1] + sp[2] + sp[3]);
  double* col = U + (i64)(nd.off + j) * ldu + nd.off;
  for (int i = threadIdx.x; i < k; i += blockDim.x) col[gpos[i]] = zh[i] / ((d[i] - dK) - tj) * inv;
}

// Pack the active columns of Q in grouped order: top rows of groups 1,2 -> W1[off.., off+g];
// bottom rows of groups 2,3 -> W1[off+n1.., off+(g-k1)].
__global__ void __launch_bounds__(256) dc_pack_kernel(const DcNode* __restrict__ nodes, DcWork wk, const double* __restrict__ Q,
                                                      i64 ldq, double* __restrict__ W1, i64 ldw) {
  const DcNode nd = nodes[blockIdx.z];
  const int g = blockIdx.x;
  if (g >= nd.k) return;
  const int off = nd.off, n1 = nd.n1, n2 = nd.sz - nd.n1;
  const int col = wk.gcol[off + g];
  const int r = blockIdx.y * blockDim.x + threadIdx.x;
  const double* src = Q + (i64)(off + col) * ldq + off;
  if (g < nd.k1 + nd.k2 && r < n1) W1[(i64)(off + g) * ldw + off + r] = src[r];
  if (g >= nd.k1 && r < n2) W1[(i64)(off + g - nd.k1) * ldw + off + n1 + r] = src[n1 + r];
}

// Root columns [crange[2t], crange[2t+1]) of node t = blockIdx.x whose sorted position falls in [lohi[2t], lohi[2t+1])
// (positions relative to the node): the roots are ascending in c, so the wanted ones form one contiguous range (a
// superset is harmless).  An empty position range gives an empty column range.
__global__ void __launch_bounds__(256) dc_colrange_kernel(const DcNode* __restrict__ nodes, DcWork wk,
                                                          const int* __restrict__ lohi, int* __restrict__ crange) {
  const int t = blockIdx.x;
  const DcNode nd = nodes[t];
  const int col_lo = lohi[2 * t], col_hi = lohi[2 * t + 1];
  __shared__ int smin, smax;
  if (threadIdx.x == 0) { smin = nd.k; smax = 0; }
  __syncthreads();
  for (int c = threadIdx.x; c < nd.k; c += blockDim.x) {
    const int p = wk.pos[nd.off + c];
    if (p >= col_lo && p < col_hi) { atomicMin(&smin, c); atomicMax(&smax, c + 1); }
  }
  __syncthreads();
  if (threadIdx.x == 0) { crange[2 * t] = smin; crange[2 * t + 1] = smax > smin ? smax : smin; }
}

// crange != nullptr (column-restricted level): node t only forms its root columns [crange[2t], crange[2t+1]).
__global__ void dc_gemm_setup_kernel(const DcNode* __restrict__ nodes, int nnodes, GemmP* __restrict__ gp, const double* W1,
                                     i64 ldw, const double* U, i64 ldu, double* W2, i64 ldw2,
                                     const int* __restrict__ crange) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnodes) return;
  const DcNode nd = nodes[t];
  const i64 off = nd.off;
  const i64 ca = crange ? crange[2 * t] : 0;
  const int ncol = crange ? crange[2 * t + 1] - crange[2 * t] : nd.k;
  GemmP a, b;
  a.m = nd.n1; a.n = ncol; a.k = nd.k1 + nd.k2;
  a.A = W1 + off * ldw + off; a.lda = ldw;
  a.B = U + (off + ca) * ldu + off; a.ldb = ldu;
  a.C = W2 + (off + ca) * ldw2 + off; a.ldc = ldw2;
  a.alpha = 1.0; a.beta = 0.0;
  b.m = nd.sz - nd.n1; b.n = ncol; b.k = nd.k2 + nd.k3;
  b.A = W1 + off * ldw + off + nd.n1; b.lda = ldw;
  b.B = U + (off + ca) * ldu + off + nd.k1; b.ldb = ldu;
  b.C = W2 + (off + ca) * ldw2 + off + nd.n1; b.ldc = ldw2;
  b.alpha = 1.0; b.beta = 0.0;
  gp[2 * t] = a;
  gp[2 * t + 1] = b;
}

// Combined values (k roots then deflated), rank sort -> pos, Dnew.
__global__ void __launch_bounds__(256) dc_rank_kernel(const DcNode* __restrict__ nodes, DcWork wk, double* __restrict__ Dnew) {
  const DcNode nd = nodes[blockIdx.y];
  const int sz = nd.sz, off = nd.off, k = nd.k;
  __shared__ double tile[256];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  auto val = [&](int q) -> double { return q < k ? wk.lam[off + q] : wk.defl_val[off + q - k]; };
  const double my = (i < sz) ? val(i) : 0.0;
  int rank = 0;
  for (int base = 0; base < sz; base += 256) {
    __syncthreads();
    if (base + threadIdx.x < sz) tile[threadIdx.x] = val(base + threadIdx.x);
    __syncthreads();
    const int cnt = min(256, sz - base);
    if (i < sz)
      for (int q = 0; q < cnt; ++q) {
        const double o = tile[q];
        rank += (o < my || (o == my && base + q < i)) ? 1 : 0;
      }
  }
  if (i < sz) {
    wk.pos[off + i] = rank;
    Dnew[off + rank] = my;
  }
}

// Qnew[:, pos[c]] = c < k ? W2[:, c] : Qcur[:, defl_col[c-k]]
__global__ void __launch_bounds__(256) dc_permute_kernel(const DcNode* __restrict__ nodes, DcWork wk, const double* __restrict__ W2,
                                                         i64 ldw2, const double* __restrict__ Qcur, i64 ldq,
                                                         double* __restrict__ Qnew, i64 ldn, const int* __restrict__ lohi) {
  const DcNode nd = nodes[blockIdx.z];
  const int col_lo = lohi ? lohi[2 * blockIdx.z] : 0, col_hi = lohi ? lohi[2 * blockIdx.z + 1] : 0x7fffffff;
  const int c = blockIdx.x;
  if (c >= nd.sz) return;
  const int off = nd.off;
  const int r = blockIdx.y * blockDim.x + threadIdx.x;
  if (r >= nd.sz) return;
  const int p = wk.pos[off + c];
  if (p < col_lo || p >= col_hi) return;
  double v;
  if (c < nd.k) v = W2[(i64)(off + c) * ldw2 + off + r];
  else v = Qcur[(i64)(off + wk.defl_col[off + c - nd.k]) * ldq + off + r];
  Qnew[(i64)(off + p) * ldn + off + r] = v;
}

// ------------------------------------------------------------------------------------------ driver
size_t stedc_workspace_bytes(i64 n) {
  const i64 ld = round_up(n, 8);
  size_t mats = (size_t)3 * ld * n * sizeof(double);  // W1, W2, Qalt
  size_t vecs = (size_t)12 * (n + 8) * sizeof(double) + (size_t)10 * (n + 8) * sizeof(int);
  size_t nodes = (size_t)(4 * (n / (DC_LEAF / 2) + 4)) * (sizeof(DcNode) + 2 * sizeof(GemmP));
  return mats + vecs + nodes + (1 << 16);
}

// d (n), e (n-1): tridiagonal (destroyed).  w (n): eigenvalues ascending.  Z (n x n, ldz): eigenvectors.
// work: stedc_workspace_bytes(n) bytes.  flops_out (optional, host): actual merge GEMM FLOPs.
int stedc(Ctx* ctx, i64 n, double* d, double* e, double* w, double* Z, i64 ldz, void* work, double* flops_out,
          i64 col_lo, i64 col_hi) {
  if (n <= 0) return 0;
  if (col_lo < 0) col_lo = 0;
  if (col_hi > n) col_hi = n;
  if (col_hi < col_lo) col_hi = col_lo;
  const bool restricted = (col_lo > 0 || col_hi < n);
  const i64 ld = round_up(n, 8);
  char* wp = (char*)work;
  auto take = [&](size_t bytes) { void* p = wp; wp += (bytes + 255) / 256 * 256; return p; };
  double* W1 = (double*)take((size_t)ld * n * 8);
  double* W2 = (double*)take((size_t)ld * n * 8);
  double* Qalt = (double*)take((size_t)ld * n * 8);
  DcWork wk;
  const size_t vb = (size_t)(n + 8) * 8, ib = (size_t)(n + 8) * 4;
  wk.z = (double*)take(vb); wk.dlam = (double*)take(vb); wk.zact = (double*)take(vb); wk.defl_val = (double*)take(vb);
  wk.rot_c = (double*)take(vb); wk.rot_s = (double*)take(vb); wk.lam = (double*)take(vb); wk.tau = (double*)take(vb);
  wk.zhat = (double*)take(vb); wk.vals = (double*)take(vb);
  double* Da = (double*)take(vb);
  double* Db = (double*)take(vb);
  wk.indx = (int*)take(ib); wk.act_col = (int*)take(ib); wk.act_type = (int*)take(ib); wk.gpos = (int*)take(ib);
  wk.gcol = (int*)take(ib); wk.defl_col = (int*)take(ib); wk.rot_a = (int*)take(ib); wk.rot_b = (int*)take(ib);
  wk.orig = (int*)take(ib); wk.pos = (int*)take(ib);

  int depth = 0;
  while (((n + ((i64)1 << depth) - 1) >> depth) > DC_LEAF) ++depth;
  const int nleaf = 1 << depth;
  // node tables for all levels
  std::vector<DcNode> hn;
  std::vector<int> lvl_start(depth + 1, 0);
  for (int l = depth - 1; l >= 0; --l) {
    lvl_start[l] = (int)hn.size();
    const int cnt = 1 << l;
    for (int i = 0; i < cnt; ++i) {
      DcNode nd = {};
      i64 o0 = (i64)i * n / cnt, o1 = (i64)(i + 1) * n / cnt, om = (i64)(2 * i + 1) * n / (2 * cnt);
      nd.off = (int)o0; nd.n1 = (int)(om - o0); nd.sz = (int)(o1 - o0);
      hn.push_back(nd);
    }
  }
  DcNode* d_nodes = (DcNode*)take(hn.size() * sizeof(DcNode) + 256);
  GemmP* d_gp = (GemmP*)take((size_t)2 * (nleaf + 2) * sizeof(GemmP));
  if (!hn.empty())
    EKB_CUDA(cudaMemcpyAsync(d_nodes, hn.data(), hn.size() * sizeof(DcNode), cudaMemcpyHostToDevice, ctx->stream));
  EKB_CUDA(cudaMemsetAsync(ctx->d_info + 1, 0, sizeof(int), ctx->stream));

  // Q buffers ping-pong between Z and Qalt so that the final level lands in Z
  double* Qcur = (depth % 2 == 0) ? Z : Qalt;
  i64 ldcur = (depth % 2 == 0) ? ldz : ld;
  double* Qnxt = (depth % 2 == 0) ? Qalt : Z;
  i64 ldnxt = (depth % 2 == 0) ? ld : ldz;
  double* Dcur = (depth % 2 == 0) ? w : Da;
  double* Dnxt = (depth % 2 == 0) ? Da : w;
  (void)Db;

  if (nleaf > 1) {
    dc_cut_kernel<<<cdiv(nleaf, 256), 256, 0, ctx->stream>>>(d, e, n, nleaf); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
  }
  dc_leaf_kernel<<<cdiv(nleaf, 4), 128, 0, ctx->stream>>>(d, e, n, nleaf, Dcur, Qcur, ldcur, ctx->d_info + 1); EKB_COUNT_LAUNCH(ctx);
  EKB_CUDA(cudaGetLastError());

  int shard_node = -1;      // level-1 node this rank formed a slab of (-1: level 1 not sharded)
  double shard_frac = 0.0;  // ... and the fraction of its columns
  for (int l = depth - 1; l >= 0; --l) {
    const int cnt = 1 << l;
    DcNode* nodes = d_nodes + lvl_start[l];
    int maxsz = 0, maxn1 = 0;
    for (int i = 0; i < cnt; ++i) {
      const DcNode& nd = hn[lvl_start[l] + i];
      maxsz = std::max(maxsz, nd.sz);
      maxn1 = std::max(maxn1, std::max(nd.n1, nd.sz - nd.n1));
    }
    dc_zero_offdiag_kernel<<<dim3(maxsz, cdiv(maxsz, 256), cnt), 256, 0, ctx->stream>>>(nodes, Qcur, ldcur); EKB_COUNT_LAUNCH(ctx);
    dc_setup_kernel<<<cnt, 256, 0, ctx->stream>>>(nodes, Dcur, e, Qcur, ldcur, wk); EKB_COUNT_LAUNCH(ctx);
    dc_deflate_kernel<<<cnt, 32, 0, ctx->stream>>>(nodes, Dcur, wk); EKB_COUNT_LAUNCH(ctx);
    dc_rotate_kernel<<<dim3(cdiv(maxsz, 256), cnt), 256, 0, ctx->stream>>>(nodes, Qcur, ldcur, wk); EKB_COUNT_LAUNCH(ctx);
    dc_secular_kernel<<<dim3(cdiv(maxsz, 128), cnt), 128, 0, ctx->stream>>>(nodes, wk); EKB_COUNT_LAUNCH(ctx);
    dc_lowner_kernel<<<dim3(maxsz, cnt), 128, 0, ctx->stream>>>(nodes, wk); EKB_COUNT_LAUNCH(ctx);
    // U lives in the (not yet written) next-level Q buffer
    dc_buildU_kernel<<<dim3(maxsz, cnt), 128, 0, ctx->stream>>>(nodes, wk, Qnxt, ldnxt); EKB_COUNT_LAUNCH(ctx);
    dc_pack_kernel<<<dim3(maxsz, cdiv(maxn1, 256), cnt), 256, 0, ctx->stream>>>(nodes, wk, Qcur, ldcur, W1, ld); EKB_COUNT_LAUNCH(ctx);
    // final positions of the merged spectrum (independent of the products, so it can steer them)
    dc_rank_kernel<<<dim3(cdiv(maxsz, 256), cnt), 256, 0, ctx->stream>>>(nodes, wk, Dnxt); EKB_COUNT_LAUNCH(ctx);
    const bool top_cut = restricted && l == 0;  // only the caller's eigenvector columns of the top merge
    // P > 1 ranks, level 1 (two merges of n/2, a quarter of all merge FLOPs, formerly replicated): rank r forms one
    // column slab of ONE of the two nodes, then the slabs are all-gathered (full columns of the block-diagonal Q) --
    // every rank ends up with the same bits it would have computed itself (the products are column-wise independent).
    const bool shard = !top_cut && l == 1 && ctx->nranks >= 2 && ctx->nranks % cnt == 0 && ctx->stedc_shard != 0 &&
                       hn[lvl_start[l]].sz >= 512;
    int* crange = nullptr;
    int* lohi = nullptr;
    int max_cols = maxsz;
    std::vector<i64> gbounds;
    if (top_cut || shard) {
      crange = ctx->d_info + 8;   // 2 ints per node of the level (<= 2 nodes)
      lohi = ctx->d_info + 16;
      int hl[4] = {0, 0, 0, 0};
      if (top_cut) {
        hl[0] = (int)col_lo; hl[1] = (int)col_hi;
        max_cols = (int)std::max<i64>(1, col_hi - col_lo);
      } else {
        const int rpn = ctx->nranks / cnt, mine = ctx->rank / rpn, q = ctx->rank % rpn;
        gbounds.assign(ctx->nranks + 1, n);
        for (int r = 0; r < ctx->nranks; ++r) {
          const DcNode& nd = hn[lvl_start[l] + r / rpn];
          std::vector<i64> sb;
          slab_bounds(nd.sz, rpn, 128, sb);
          gbounds[r] = nd.off + sb[r % rpn];
          if (r == ctx->rank) {
            hl[2 * mine] = (int)sb[q]; hl[2 * mine + 1] = (int)sb[q + 1];
            max_cols = (int)std::max<i64>(1, sb[q + 1] - sb[q]);
            shard_node = lvl_start[l] + mine;
            shard_frac = (double)(sb[q + 1] - sb[q]) / (double)nd.sz;
          }
        }
      }
      EKB_CUDA(cudaMemcpyAsync(lohi, hl, 2 * cnt * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
      EKB_CUDA(cudaStreamSynchronize(ctx->stream));  // hl is a stack array
      dc_colrange_kernel<<<cnt, 256, 0, ctx->stream>>>(nodes, wk, lohi, crange); EKB_COUNT_LAUNCH(ctx);
    }
    dc_gemm_setup_kernel<<<cdiv(cnt, 128), 128, 0, ctx->stream>>>(nodes, cnt, d_gp, W1, ld, Qnxt, ldnxt, W2, ld, crange); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
    if (col_hi > col_lo || !top_cut) EKB_TRY(gemm_batched(ctx, 0, d_gp, 2 * cnt, maxn1, max_cols, /*k_hint=*/maxn1));
    dc_permute_kernel<<<dim3(maxsz, cdiv(maxsz, 256), cnt), 256, 0, ctx->stream>>>(nodes, wk, W2, ld, Qcur, ldcur, Qnxt,
                                                                                ldnxt, lohi); EKB_COUNT_LAUNCH(ctx);
    EKB_CUDA(cudaGetLastError());
    if (shard) EKB_TRY(comm_allgather_cols(ctx, Qnxt, ldnxt, gbounds));
    std::swap(Qcur, Qnxt);
    std::swap(ldcur, ldnxt);
    std::swap(Dcur, Dnxt);
  }
  // results are in (Dcur, Qcur) == (w, Z) by construction of the ping-pong parity
  if (Qcur != Z || Dcur != w) return EKB_ERR_INTERNAL;
  // failure flag + actual FLOP count
  EKB_CUDA(cudaMemcpyAsync(ctx->h_info + 1, ctx->d_info + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (restricted)
    EKB_CUDA(cudaMemcpyAsync(ctx->h_info + 8, ctx->d_info + 8, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (flops_out && !hn.empty())
    EKB_CUDA(cudaMemcpyAsync(hn.data(), d_nodes, hn.size() * sizeof(DcNode), cudaMemcpyDeviceToHost, ctx->stream));
  EKB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (flops_out) {
    double fl = 0.0;
    for (size_t q = 0; q < hn.size(); ++q) {
      const DcNode& nd = hn[q];
      double cols = nd.k;
      if (restricted && (int)q == lvl_start[0]) cols = ctx->h_info[9] - ctx->h_info[8];
      if (shard_node >= 0 && depth >= 2 && (int)q >= lvl_start[1] && (int)q < lvl_start[1] + 2)
        cols = (int)q == shard_node ? nd.k * shard_frac : 0.0;  // this rank's share of the sharded level (estimate)
      fl += 2.0 * cols * ((double)nd.n1 * (nd.k1 + nd.k2) + (double)(nd.sz - nd.n1) * (nd.k2 + nd.k3));
    }
    *flops_out = fl;
  }
  if (ctx->h_info[1] != 0) return ctx->h_info[1];  // number of leaves whose QL iteration failed
  return 0;
}

}  // namespace ekb
