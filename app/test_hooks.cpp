// extern "C" hooks over the host modules of ekb200_app for the CPU-only tests (tests/test_host_app_cpp.py, ctypes).
// Not part of the product boundary (that is include/ekb200.h).
#include <string.h>

#include "ek_app.hpp"

using namespace ekapp;

static int put(const std::string& s, char* out, int cap) {
  if (cap <= 0) return (int)s.size();
  const int n = (int)s.size() < cap - 1 ? (int)s.size() : cap - 1;
  memcpy(out, s.data(), n);
  out[n] = 0;
  return (int)s.size();
}

extern "C" {

int ekapp_fortran_e(double x, int width, int digits, int expw, char* out, int cap) {
  return put(fortran_e(x, width, digits, expw), out, cap);
}

// add_event for every (name, val) in order, then the log.json text for a fixed setting block
int ekapp_log_json(int n, const char** names, const double* vals, const char* command, const char* fileA,
                   const char* fileB, const char* solver, long long dimension, int block_size, char* out, int cap) {
  clear_events();
  set_world(1, 2);  // not the master: no echo on stderr
  for (int i = 0; i < n; ++i) add_event(names[i], vals[i]);
  set_world(0, 1);
  ek_argument_t arg;
  arg.command = command;
  arg.matrix_A_filename = fileA;
  arg.matrix_B_filename = fileB;
  arg.solver_type = solver;
  arg.matrix_A_info.rows = dimension;
  arg.block_size = block_size;
  return put(log_json_text(arg, events()), out, cap);
}

// mminfo + read_matrix_file; returns 0, mminfo's/open's ierr, or 1000 + code with the terminate message in msg
int ekapp_read_matrix(const char* path, int threads, long long* rows, long long* cols, long long* entries, int* ij,
                      double* v, long long cap, char* msg, int msgcap) {
  set_world(1, 2);  // quiet
  ek_matrix_info_t info;
  int ierr = wrap_mminfo(path, info);
  set_world(0, 1);
  if (ierr) return ierr;
  *rows = info.rows;
  *cols = info.cols;
  *entries = info.entries;
  if (!ij || !v) return 0;
  ek_sparse_mat_t m;
  try {
    set_world(1, 2);
    read_matrix_file(path, info, m, ierr, threads);
    set_world(0, 1);
  } catch (const Terminate& t) {
    set_world(0, 1);
    put(t.message, msg, msgcap);
    return 1000 + (t.code & 0xff);
  }
  if (ierr) return ierr;
  const long long k = m.num_non_zeros < cap ? m.num_non_zeros : cap;
  memcpy(ij, m.suffix.data(), (size_t)k * 2 * sizeof(int));
  memcpy(v, m.value.data(), (size_t)k * sizeof(double));
  return 0;
}

// -p parser: returns the number of ranges (or -1 with the terminate message)
int ekapp_parse_ranges(const char* spec, long long* ranges, int cap, char* msg, int msgcap) {
  int num = 0;
  int64_t r[kMaxNumPrintedVecsRanges][2];
  set_world(1, 2);
  try {
    arg_str_to_printed_vecs_ranges(spec, num, r);
  } catch (const Terminate& t) {
    set_world(0, 1);
    put(t.message, msg, msgcap);
    return -1;
  }
  set_world(0, 1);
  for (int i = 0; i < num && i < cap; ++i) {
    ranges[2 * i] = r[i][0];
    ranges[2 * i + 1] = r[i][1];
  }
  return num;
}
}

// print_eigenvectors on a caller-provided n x k matrix (1 x 1 grid): files <dir>/<j:08d>.dat for j in ranges
extern "C" int ekapp_print_eigenvectors(const char* dir, long long n, long long k, const double* X, long long ldx,
                                        const char* ranges_spec, int binary, int threads, char* msg, int msgcap) {
  set_world(1, 2);  // quiet
  ek_argument_t arg;
  arg.eigenvector_dir = dir;
  arg.is_binary_output = binary != 0;
  arg.io_threads = threads;
  ek_eigenpairs_types_union_t ep;
  ep.type_number = 2;
  ep.blacs.desc[rows_] = n;
  ep.blacs.desc[cols_] = k;
  ep.blacs.Vectors = const_cast<double*>(X);
  ep.blacs.lld = ldx;
  ep.blacs.loc_cols = k;
  ep.blacs.col0 = 0;
  int rc = 0;
  try {
    arg_str_to_printed_vecs_ranges(ranges_spec, arg.num_printed_vecs_ranges, arg.printed_vecs_ranges);
    print_eigenvectors(arg, ep);
  } catch (const Terminate& t) {
    put(t.message, msg, msgcap);
    rc = 1000 + (t.code & 0xff);
  }
  set_world(0, 1);
  return rc;
}
