// ek_verifier_m (reference src/verifier.f90) and get_ipratios (src/distribute_matrix.f90:18-78), computed on the
// B200 (csrc/verify.cu) from what the reference routines receive: the replicated COO matrices and the type-2
// eigenpairs (each rank passes its local block column; the results come back identical on every rank).
//   eval_residual_norm   verifier.f90:75-204   -c : A_norm, mean and max of ||A x - lambda B x||_2 / ||A||_F
//   eval_orthogonality   verifier.f90:233-330  -t : || scaled Gram matrix - I ||_F over columns index1..index2
//   get_ipratios         distribute_matrix.f90:18-78 : ipratios.dat
// For synthetic inputs (nothing to scatter from COO) the matrices are regenerated in HBM and the device-resident
// entry points are used on the eigenvectors the solve left on the device.
#include <stdio.h>

#include "../include/ekb200.h"
#include "ek_app.hpp"

namespace ekapp {

ekb200_ctx* b200_context(const ek_process_t& proc);
int b200_fill_synthetic_A(ekb200_ctx* ctx, const ek_matrix_info_t& info, double* dA, int64_t ld);
int b200_fill_synthetic_B(ekb200_ctx* ctx, const ek_matrix_info_t& info, double* dB, int64_t ld);
double* b200_device_vectors(int64_t* ld);
double* b200_device_values();

static void check(int info, const char* routine) {
  if (info == 0) return;
  if (check_master()) printf("info(%s): %d\n", routine, info);
  terminate(std::string(routine) + " failed", info);
}
static ekb200_ctx* context() {
  ek_process_t proc;  // the grid of the solve: 1 x P in rank order (no second "BLACS process grid" line)
  proc.my_rank = world_rank();
  proc.n_procs = world_size();
  proc.n_procs_col = world_size();
  proc.my_proc_col = world_rank();
  return b200_context(proc);
}
static const int32_t kIjDummy[2] = {1, 1};
static const double kVDummy[1] = {0.0};

// Synthetic multi-rank runs: the device-resident checks slice the CHECKED columns by rank, which need not coincide
// with the solve's slabs, so every rank first obtains all eigenvector columns (one all-gather over NVLink).
static void ensure_all_columns(ekb200_ctx* ctx, int64_t n, int64_t nvec) {
  static bool gathered = false;
  if (gathered || world_size() <= 1) return;
  int64_t ldz = 0;
  double* dZ = b200_device_vectors(&ldz);
  check(ekb200_comm_allgather_slabs(ctx, n, nvec, dZ, ldz), "ekb200_comm_allgather_slabs");
  gathered = true;
}

struct SyntheticDev {
  ekb200_ctx* ctx;
  double *dA = nullptr, *dB = nullptr;
  int64_t ld = 0;
  SyntheticDev(ekb200_ctx* c, const ek_argument_t& arg, bool needA, bool needB) : ctx(c) {
    const int64_t n = arg.matrix_A_info.rows;
    ld = (n + 7) / 8 * 8;
    if (needA) {
      check(ekb200_dev_alloc(ctx, ld * n * 8, (void**)&dA), "ekb200_dev_alloc");
      check(b200_fill_synthetic_A(ctx, arg.matrix_A_info, dA, ld), "ekb200_fill_synthetic");
    }
    if (needB && arg.is_generalized_problem) {
      check(ekb200_dev_alloc(ctx, ld * n * 8, (void**)&dB), "ekb200_dev_alloc");
      check(b200_fill_synthetic_B(ctx, arg.matrix_B_info, dB, ld), "ekb200_fill_synthetic");
    }
  }
  ~SyntheticDev() {
    if (dA) ekb200_dev_free(ctx, dA);
    if (dB) ekb200_dev_free(ctx, dB);
  }
};

void eval_residual_norm(const ek_argument_t& arg, const ek_sparse_mat_t& matrix_A,
                        const ek_eigenpairs_types_union_t& eigenpairs, double& A_norm, double& res_norm_ave,
                        double& res_norm_max, const ek_sparse_mat_t* matrix_B) {
  const double time_start = wtime();
  if (eigenpairs.type_number != 2) terminate("eval_residual_norm: eigenpairs of type 2 (BLACS) expected", 1);
  if (arg.is_generalized_problem && !matrix_B) terminate("eval_residual_norm_blacs: matrix_B is not provided", 1);
  const ek_eigenpairs_blacs_t& ep = eigenpairs.blacs;
  const int64_t n = ep.desc[rows_], nvec = ep.desc[cols_];
  ekb200_ctx* ctx = context();
  if (arg.matrix_A_info.synthetic) {
    SyntheticDev m(ctx, arg, true, true);
    int64_t ldz = 0;
    double* dZ = b200_device_vectors(&ldz);
    ensure_all_columns(ctx, n, nvec);
    check(ekb200_eval_residual_norm_dev(ctx, n, arg.n_check_vec, m.dA, m.ld, m.dB, m.ld, b200_device_values(), dZ, ldz,
                                        &A_norm, &res_norm_ave, &res_norm_max),
          "ekb200_eval_residual_norm_dev");
  } else {
    const bool gen = arg.is_generalized_problem;
    check(ekb200_eval_residual_norm(ctx, n, nvec, arg.n_check_vec, matrix_A.num_non_zeros, matrix_A.suffix.data(),
                                    matrix_A.value.data(), gen ? matrix_B->num_non_zeros : 0,
                                    gen ? matrix_B->suffix.data() : kIjDummy, gen ? matrix_B->value.data() : kVDummy,
                                    ep.values.data(), ep.Vectors, ep.lld, &A_norm, &res_norm_ave, &res_norm_max),
          "ekb200_eval_residual_norm");
  }
  add_event("eval_residual_norm_blacs", wtime() - time_start);
}

void eval_orthogonality(const ek_argument_t& arg, const ek_eigenpairs_types_union_t& eigenpairs, double& orthogonality,
                        const ek_sparse_mat_t* matrix_B) {
  const double time_start = wtime();
  if (eigenpairs.type_number != 2) terminate("eval_orthogonality: eigenpairs of type 2 (BLACS) expected", 1);
  const ek_eigenpairs_blacs_t& ep = eigenpairs.blacs;
  if (ep.desc[block_row_] != ep.desc[block_col_])
    terminate("eval_orthogonality_blacs: anisotropic block size not supported", 1);  // verifier.f90:256-259
  const int64_t n = ep.desc[rows_], nvec = ep.desc[cols_];
  ekb200_ctx* ctx = context();
  if (arg.matrix_A_info.synthetic) {
    SyntheticDev m(ctx, arg, false, true);
    int64_t ldz = 0;
    double* dZ = b200_device_vectors(&ldz);
    ensure_all_columns(ctx, n, nvec);
    check(ekb200_eval_orthogonality_dev(ctx, n, arg.ortho_check_index_start, arg.ortho_check_index_end, dZ, ldz, m.dB,
                                        m.ld, &orthogonality),
          "ekb200_eval_orthogonality_dev");
  } else {
    const bool gen = arg.is_generalized_problem && matrix_B;
    check(ekb200_eval_orthogonality(ctx, n, nvec, arg.ortho_check_index_start, arg.ortho_check_index_end,
                                    gen ? matrix_B->num_non_zeros : 0, gen ? matrix_B->suffix.data() : kIjDummy,
                                    gen ? matrix_B->value.data() : kVDummy, ep.Vectors, ep.lld, &orthogonality),
          "ekb200_eval_orthogonality");
  }
  add_event("eval_orthogonality_blacs", wtime() - time_start);
}

void get_ipratios(const ek_argument_t& arg, const ek_process_t& proc, const ek_eigenpairs_types_union_t& eigenpairs,
                  std::vector<double>& ipratios, const ek_sparse_mat_t* matrix_B) {
  const ek_eigenpairs_blacs_t& ep = eigenpairs.blacs;
  const int64_t n = ep.desc[rows_], nvec = ep.desc[cols_];
  if (matrix_B && matrix_B->size != n) terminate("inconsistent matrix dimension", 1);  // distribute_matrix.f90:33-35
  ipratios.assign((size_t)nvec, 0.0);
  ekb200_ctx* ctx = b200_context(proc);
  if (arg.matrix_A_info.synthetic) {
    SyntheticDev m(ctx, arg, false, true);
    int64_t ldz = 0;
    double* dZ = b200_device_vectors(&ldz);
    ensure_all_columns(ctx, n, nvec);
    check(ekb200_get_ipratios_dev(ctx, n, nvec, dZ, ldz, m.dB, m.ld, ipratios.data()), "ekb200_get_ipratios_dev");
  } else {
    const bool gen = matrix_B != nullptr;
    check(ekb200_get_ipratios(ctx, n, nvec, gen ? matrix_B->num_non_zeros : 0, gen ? matrix_B->suffix.data() : kIjDummy,
                              gen ? matrix_B->value.data() : kVDummy, ep.Vectors, ep.lld, ipratios.data()),
          "ekb200_get_ipratios");
  }
}

}  // namespace ekapp
