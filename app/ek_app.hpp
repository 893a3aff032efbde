// ekb200_app -- C++ host twin of EigenKernel_App (reference src/main.f90) for the B200 solvers.
//
// The reference host code is Fortran + MPI + ScaLAPACK, none of which exist in this image, so the host side
// above the C-ABI (include/ekb200.h) is restated in C++ module by module, keeping the reference's names,
// argument meaning, output formats and error behaviour:
//
//   command_argument.cpp   ek_command_argument_m   src/command_argument.f90   (-s/-n/-c/-o/-i/-d/-p/-t/-l/-h ...)
//   matrix_io.cpp          ek_matrix_io_m + mminfo src/matrix_io.f90, src/mmio.f:341-585
//   event_logger.cpp       ek_event_logger_m + the fson printer  src/event_logger.f90, src/fson.f90:454-553
//   processes.cpp          ek_processes_m          src/processes.f90 (terminate, check_master, grid)
//   solver_main.cpp        ek_solver_main_m        src/solver_main.f90:22-100 (dispatch on -s)
//   solver_b200.cpp        ek_solver_b200_m        the new backend (twin of fortran/solver_b200.f90)
//   verifier.cpp           ek_verifier_m + get_ipratios  src/verifier.f90, src/distribute_matrix.f90:18-78
//   main.cpp               program eigbench        src/main.f90
//
// All arithmetic happens in libekb200.so (hand-written CUDA for sm_100a); this program has no CPU solve path and
// stops with `[Error] ...` when the library cannot get a GPU.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace ekapp {

// ---------------------------------------------------------------- global_variables.f90
extern int g_block_size;                 // 64, overridden by --block-size (solver_main.f90:44-46)
extern const char* const g_version;      // '20160808'
extern double g_wtime_init;              // g_mpi_wtime_init

double wtime();                          // mpi_wtime(): monotonic seconds

// ---------------------------------------------------------------- processes.f90
struct Terminate {                       // what `call terminate(msg, code)` raises (processes.f90:122-139)
  std::string message;
  int code;
};
[[noreturn]] void terminate(const std::string& error_message, int error_code);
bool check_master();
// Rank-per-GPU mode: the ranks are forked by --ngpu P before CUDA is touched; rank r drives GPU r.
struct ek_process_t {                    // processes.f90:6-9; the grid of the B200 backend is 1 x P (column slabs)
  int my_rank = 0, n_procs = 1, context = 0, n_procs_row = 1, n_procs_col = 1, my_proc_row = 0, my_proc_col = 0;
};
void set_world(int rank, int size);
int world_rank();
int world_size();
void setup_distribution(ek_process_t& proc);

// ---------------------------------------------------------------- Fortran edit descriptors as gfortran prints them
std::string fortran_e(double x, int width, int digits, int expw);   // Ew.dEe  (expw = 0: Ew.d)
std::string fortran_f(double x, int width, int digits);             // Fw.d
std::string fortran_i(long long v, int width);                      // Iw (width 0 = I0)

// ---------------------------------------------------------------- event_logger.f90
struct event_t {
  std::string name;
  int num_repeated;
  double val;
};
void add_event(const std::string& name, double val, bool to_print = true);
int num_events();
void print_events();
const std::vector<event_t>& events();    // newest name first (the reference's linked list order)
void clear_events();

// ---------------------------------------------------------------- command_argument.f90
constexpr int kMaxNumPrintedVecsRanges = 100;
struct ek_matrix_info_t {
  std::string rep, field, symm;
  int64_t rows = 0, cols = 0, entries = 0;
  // extension: `synthetic:<n>:<seed>[:<diag shift>]` in place of a path; never set for real files
  bool synthetic = false;
  uint64_t seed = 0;
  double diag_shift = 0.0;
};
struct ek_argument_t {
  std::string matrix_A_filename, matrix_B_filename;
  std::string log_filename = "log.json";
  ek_matrix_info_t matrix_A_info, matrix_B_info;
  std::string solver_type;
  std::string output_filename = "eigenvalues.dat";
  std::string ipratios_filename = "ipratios.dat";
  bool is_generalized_problem = false;
  bool is_printing_grid_mapping = false;
  bool is_dry_run = false;
  bool is_binary_output = false;
  int block_size = 0;
  int64_t n_vec = -1;
  int64_t n_check_vec = 0;
  int64_t ortho_check_index_start = 0, ortho_check_index_end = 0;
  std::string eigenvector_dir = ".";
  int num_printed_vecs_ranges = 0;
  int64_t printed_vecs_ranges[kMaxNumPrintedVecsRanges][2];
  int verbose_level = 0;
  // extensions of this twin (all default to the reference's behaviour)
  int io_threads = 0;                    // --io-threads <n>: MatrixMarket parsing / eigenvector printing threads
  int ngpu = 1;                          // --ngpu <P>: rank-per-GPU mode, P forked ranks
  std::string command;                   // argv joined by blanks (log.json "command")
};
void print_help();
void read_command_argument(int argc, char** argv, ek_argument_t& arg);
void validate_argument(const ek_argument_t& arg);
double required_memory(const ek_argument_t& arg);
void print_command_argument(const ek_argument_t& arg);
void arg_str_to_printed_vecs_ranges(const std::string& arg_str, int& num, int64_t ranges[][2]);
bool is_b200_solver(const std::string& s);
bool is_b200_select_solver(const std::string& s);
bool is_b200_generalized_solver(const std::string& s);

// log.json: the two-object tree main.f90:58-60,185-190 builds, printed byte-for-byte like fson_value_print
std::string log_json_text(const ek_argument_t& arg, const std::vector<event_t>& evs);

// ---------------------------------------------------------------- matrix_io.f90
struct ek_sparse_mat_t {                 // replicated COO, 1-based, one triangle stored
  int64_t size = 0, num_non_zeros = 0;
  std::vector<double> value;
  std::vector<int32_t> suffix;           // Fortran suffix(2, nnz): (i, j) pairs, contiguous
};
int wrap_mminfo(const std::string& filename, ek_matrix_info_t& minfo);   // returns mminfo's ierr
void read_matrix_file(const std::string& filename, const ek_matrix_info_t& info, ek_sparse_mat_t& matrix, int& ierr,
                      int threads = 0);

// ---------------------------------------------------------------- eigenpairs_types.f90 + descriptor_parameters.f90
enum { dtype_ = 0, context_ = 1, rows_ = 2, cols_ = 3, block_row_ = 4, block_col_ = 5, rsrc_ = 6, csrc_ = 7,
       local_rows_ = 8, desc_size = 9 };
struct ek_eigenpairs_blacs_t {
  std::vector<double> values;
  int64_t desc[desc_size] = {0};
  // local piece: lld x loc_cols, column-major; columns [col0, col0 + loc_cols) of the global n x n_vec matrix
  double* Vectors = nullptr;             // pinned host memory owned by the backend context
  int64_t lld = 0, loc_cols = 0, col0 = 0;
};
struct ek_eigenpairs_types_union_t {
  int type_number = 0;
  ek_eigenpairs_blacs_t blacs;
};
void print_eigenvectors(const ek_argument_t& arg, const ek_eigenpairs_types_union_t& eigenpairs);

// ---------------------------------------------------------------- solver_main.f90 / the b200 backend
void eigen_solver(ek_argument_t& arg, const ek_sparse_mat_t& matrix_A, ek_eigenpairs_types_union_t& eigenpairs,
                  ek_process_t& proc, const ek_sparse_mat_t* matrix_B);
void solve_with_b200(const ek_argument_t& arg, int64_t n, const ek_process_t& proc, const ek_sparse_mat_t& matrix_A,
                     ek_eigenpairs_types_union_t& eigenpairs, const ek_sparse_mat_t* matrix_B);
void b200_finalize();                    // releases the backend context (end of main)

// ---------------------------------------------------------------- verifier.f90 / distribute_matrix.f90:18-78
void eval_residual_norm(const ek_argument_t& arg, const ek_sparse_mat_t& matrix_A,
                        const ek_eigenpairs_types_union_t& eigenpairs, double& A_norm, double& res_norm_ave,
                        double& res_norm_max, const ek_sparse_mat_t* matrix_B);
void eval_orthogonality(const ek_argument_t& arg, const ek_eigenpairs_types_union_t& eigenpairs, double& orthogonality,
                        const ek_sparse_mat_t* matrix_B);
void get_ipratios(const ek_argument_t& arg, const ek_process_t& proc, const ek_eigenpairs_types_union_t& eigenpairs,
                  std::vector<double>& ipratios, const ek_sparse_mat_t* matrix_B);

}  // namespace ekapp
