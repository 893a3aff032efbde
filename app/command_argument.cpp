// ek_command_argument_m (reference src/command_argument.f90): option parsing, validation, the configuration
// print-out and the memory estimate, with the B200 solver names added in the four places SURVEY 8(b) lists
// (print_help :55-68, validate_argument :140-173 and :188-195, required_memory :321-334).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ek_app.hpp"

namespace ekapp {

// names of the reference whose back-ends (ScaLAPACK / ELPA / EigenExa) are not part of this build: they are
// recognised and refused the way the *_dummy.f90 twins refuse (solver_elpa_dummy.f90:14-22)
static const char* const kReferenceOnly[] = {
    "lapack", "scalapack", "scalapack_select", "general_scalapack", "general_scalapack_select", "eigensx",
    "general_scalapack_eigensx", "general_scalapack_eigens", "general_elpa_scalapack", "general_elpa1",
    "general_elpa2", "general_elpa_eigensx", "general_elpa_eigens", "general_scalapacknew_eigens"};

bool is_b200_select_solver(const std::string& s) { return s == "b200_select" || s == "general_b200_select"; }
bool is_b200_generalized_solver(const std::string& s) {
  return s == "general_b200" || s == "general_b200_select" || s == "general_b200inv";
}
bool is_b200_solver(const std::string& s) { return s == "b200" || s == "b200_select" || is_b200_generalized_solver(s); }
static bool is_reference_only(const std::string& s) {
  for (const char* k : kReferenceOnly)
    if (s == k) return true;
  return false;
}

void print_help() {
  if (!check_master()) return;
  puts(" Usage: eigen_test -s <solver_type> <options> <matrix_A> [<matrix_B>]");
  puts(" Solver types are:");
  puts("   b200 (standard)");
  puts("   b200_select (standard, selecting)");
  puts("   general_b200 (generalized)");
  puts("   general_b200_select (generalized, selecting)");
  puts("   general_b200inv (generalized, explicit-inverse reduction)");
  puts(" Options are:");
  puts("   -n <num>  (available with selecting solvers) Compute only <num> eigenpairs in ascending order of their eigenvalues");
  puts("   -c <num>  Consider only <num> eigenvectors in residual norm checking. Default is 0. Set -1 to consider all the vectors");
  puts("   -o <file>  Set output file name for eigenvalues to <file>");
  puts("   -i <file>  Set output file name for ipratios to <file>");
  puts("   -d <dir>  Set output files directory for eigenvectors to <dir>");
  puts("   -p <num1>,<num2>  Specify range of the number of eigenvectors to be output");
  puts("   -t <num1>,<num2>  Consider eigenvectors indexed <num1> to <num2>(included) in orthogonality checking");
  puts("   -l <file>  Set output file name for elapse time log to <file>");
  puts("   -h  Print this help and exit");
  puts("   --block-size <n>  Change block size in block cyclic distribution");
  puts("   --dry-run  Read command arguments and matrix files and instantly exit");
  puts("   --print-grid-mapping  Print which process is assigned to each coordinate in BLACS grid");
  puts("   --binary  Write eigenvectors as Fortran unformatted records");
  puts("   --ngpu <P>  (this build) one rank per B200, P of them on this box");
  puts("   --io-threads <n>  (this build) threads for MatrixMarket parsing and eigenvector printing");
  puts(" A matrix may be given as synthetic:<n>:<seed>[:<diagonal shift>] instead of a MatrixMarket file");
  fflush(stdout);
}

// Fortran list-directed `read (str, *) integer`: leading blanks, optional sign, digits; anything else is an error
static bool read_int(const std::string& s, int64_t& out) {
  const char* p = s.c_str();
  while (*p == ' ') ++p;
  char* end = nullptr;
  long long v = strtoll(p, &end, 10);
  if (end == p) return false;
  while (*end == ' ') ++end;
  if (*end != 0 && *end != ',' && *end != '/') return false;
  out = v;
  return true;
}
static int64_t read_int_or_terminate(const std::string& s, const char* what) {
  int64_t v = 0;
  if (!read_int(s, v)) terminate(std::string("read_command_argument: invalid number for ") + what, 1);
  return v;
}

// command_argument.f90:271-315: `a[-b][,c[-d]]...`, at most kMaxNumPrintedVecsRanges ranges
void arg_str_to_printed_vecs_ranges(const std::string& arg_str, int& num, int64_t ranges[][2]) {
  num = 0;
  size_t k1 = 0;
  const size_t len = arg_str.size();
  while (true) {
    size_t comma = arg_str.find(',', k1);
    if (comma == k1) terminate("arg_str_to_printed_vecs_ranges: invalid comma placement", 1);
    const size_t k2 = comma == std::string::npos ? len : comma;  // [k1, k2)
    const std::string part = arg_str.substr(k1, k2 - k1);
    if (num + 1 > kMaxNumPrintedVecsRanges) {
      printf(" arg_str_to_printed_vecs_ranges: too many ranges %d (> %d)\n", num + 1, kMaxNumPrintedVecsRanges);
      terminate("arg_str_to_printed_vecs_ranges: too many ranges", 1);
    }
    size_t hy = part.find('-');
    if (hy == 0) terminate("arg_str_to_printed_vecs_ranges: invalid hyphen placement", 1);
    int64_t a, b;
    if (hy == std::string::npos) {
      a = b = read_int_or_terminate(part, "-p");
    } else {
      a = read_int_or_terminate(part.substr(0, hy), "-p");
      b = read_int_or_terminate(part.substr(hy + 1), "-p");
    }
    ranges[num][0] = a;
    ranges[num][1] = b;
    ++num;
    if (comma == std::string::npos || comma + 1 >= len) break;
    k1 = comma + 1;
  }
  if (check_master()) {
    printf(" arg_str_to_printed_vecs_ranges: num_printed_vecs_ranges %d\n", num);
    for (int i = 0; i < num; ++i)
      printf(" arg_str_to_printed_vecs_ranges: range %d : %lld - %lld\n", i + 1, (long long)ranges[i][0],
             (long long)ranges[i][1]);
  }
}

// `synthetic:<n>:<seed>[:<shift>]`: the dense counter-hash matrices that cannot go through a
// MatrixMarket file at n = 32768.  Fills the info block the way mminfo would.
static bool parse_synthetic(const std::string& name, ek_matrix_info_t& info) {
  if (name.compare(0, 10, "synthetic:") != 0) return false;
  std::vector<std::string> f;
  size_t p = 10;
  while (true) {
    size_t q = name.find(':', p);
    f.push_back(name.substr(p, q == std::string::npos ? std::string::npos : q - p));
    if (q == std::string::npos) break;
    p = q + 1;
  }
  if (f.size() < 2 || f.size() > 3) terminate("read_command_argument: synthetic:<n>:<seed>[:<shift>] expected", 1);
  int64_t n = 0, seed = 0;
  if (!read_int(f[0], n) || !read_int(f[1], seed) || n < 1) terminate("read_command_argument: bad synthetic matrix spec", 1);
  info.rep = "coordinate";
  info.field = "real";
  info.symm = "symmetric";
  info.rows = info.cols = n;
  info.entries = 0;  // nothing is read from disk
  info.synthetic = true;
  info.seed = (uint64_t)seed;
  info.diag_shift = f.size() == 3 ? atof(f[2].c_str()) : 0.0;
  return true;
}

static void load_info(const std::string& filename, ek_matrix_info_t& info) {
  if (parse_synthetic(filename, info)) return;
  int ierr = 0;
  if (check_master()) ierr = wrap_mminfo(filename, info);
  else ierr = wrap_mminfo(filename, info);  // every forked rank reads the header itself (replaces bcast_matrix_info)
  if (ierr != 0) terminate("mminfo " + filename + " failed", ierr);
}

void read_command_argument(int argc, char** argv, ek_argument_t& arg) {
  for (int i = 0; i < argc; ++i) {
    arg.command += argv[i];
    if (i + 1 < argc) arg.command += ' ';
  }
  auto next = [&](int& argi) -> std::string {
    ++argi;
    return argi < argc ? std::string(argv[argi]) : std::string();  // get_command_argument past the end gives blanks
  };
  for (int argi = 1; argi < argc; ++argi) {
    const std::string a = argv[argi];
    if (!a.empty() && a[0] == '-') {
      const std::string opt = a.substr(1);
      if (opt == "s") arg.solver_type = next(argi);
      else if (opt == "n") arg.n_vec = read_int_or_terminate(next(argi), "-n");
      else if (opt == "c") arg.n_check_vec = read_int_or_terminate(next(argi), "-c");
      else if (opt == "o") arg.output_filename = next(argi);
      else if (opt == "i") arg.ipratios_filename = next(argi);
      else if (opt == "d") arg.eigenvector_dir = next(argi);
      else if (opt == "p") arg_str_to_printed_vecs_ranges(next(argi), arg.num_printed_vecs_ranges, arg.printed_vecs_ranges);
      else if (opt == "t") {
        const std::string s = next(argi);
        size_t c = s.find(',');
        if (c == std::string::npos) terminate("read_command_argument: wrong format for -t option", 1);
        arg.ortho_check_index_start = read_int_or_terminate(s.substr(0, c), "-t");
        arg.ortho_check_index_end = read_int_or_terminate(s.substr(c + 1), "-t");
      } else if (opt == "v") arg.verbose_level = 1;
      else if (opt == "l") arg.log_filename = next(argi);
      else if (opt == "h") {
        print_help();
        terminate("read_command_argument: help printed", 0);
      } else if (opt == "-block-size") arg.block_size = (int)read_int_or_terminate(next(argi), "--block-size");
      else if (opt == "-dry-run") arg.is_dry_run = true;
      else if (opt == "-print-grid-mapping") arg.is_printing_grid_mapping = true;
      else if (opt == "-binary") arg.is_binary_output = true;
      else if (opt == "-ngpu") arg.ngpu = (int)read_int_or_terminate(next(argi), "--ngpu");
      else if (opt == "-io-threads") arg.io_threads = (int)read_int_or_terminate(next(argi), "--io-threads");
      else {
        print_help();
        terminate("read_command_argument: unknown option" + a, 1);
      }
    } else if (arg.matrix_A_filename.empty()) {
      arg.matrix_A_filename = a;  // the first non-option argument is the (left) input matrix
    } else {
      arg.matrix_B_filename = a;
    }
  }
  if (arg.matrix_A_filename.empty()) terminate("read_command_argument: Matrix A file not specified", 1);
  arg.is_generalized_problem = !arg.matrix_B_filename.empty();
  load_info(arg.matrix_A_filename, arg.matrix_A_info);
  if (arg.is_generalized_problem) load_info(arg.matrix_B_filename, arg.matrix_B_info);
  if (arg.n_vec == -1) arg.n_vec = arg.matrix_A_info.rows;  // unspecified on the command line
  if (arg.n_check_vec == -1) arg.n_check_vec = arg.n_vec;
}

void validate_argument(const ek_argument_t& arg) {
  const int64_t dim = arg.matrix_A_info.rows;
  bool is_size_valid = dim == arg.matrix_A_info.cols;
  if (arg.is_generalized_problem)
    is_size_valid = is_size_valid && dim == arg.matrix_B_info.rows && dim == arg.matrix_B_info.cols;
  if (!is_size_valid) terminate("validate_argument: Matrix dimension mismatch", 1);
  if (arg.is_generalized_problem && arg.matrix_A_info.synthetic != arg.matrix_B_info.synthetic)
    terminate("validate_argument: synthetic and file matrices cannot be mixed", 1);

  const std::string& st = arg.solver_type;
  bool is_solver_valid = false;
  if (is_b200_solver(st)) {
    is_solver_valid = is_b200_generalized_solver(st) == arg.is_generalized_problem;
  } else if (is_reference_only(st)) {
    terminate("eigen_solver: solver '" + st + "' is not supported in this build", 1);
  } else {
    terminate("validate_argument: Unknown solver '" + st + "'", 1);
  }
  if (!is_solver_valid) {
    if (arg.is_generalized_problem)
      terminate("validate_argument: solver '" + st + "' is not for generalized eigenvalue problem", 1);
    else
      terminate("validate_argument: solver '" + st + "' is not for standard eigenvalue problem", 1);
  }
  const bool is_n_vec_valid = is_b200_select_solver(st) ? true : arg.n_vec == dim;
  if (!is_n_vec_valid)
    terminate("validate_argument: Solver '" + st + "' does not support partial eigenvalue computation", 1);
  if (is_b200_select_solver(st) && !(arg.n_vec > 0 && arg.n_vec <= dim))
    terminate("validate_argument: Specified number with -n option is not valid", 1);
  for (int i = 0; i < arg.num_printed_vecs_ranges; ++i) {
    const int64_t a = arg.printed_vecs_ranges[i][0], b = arg.printed_vecs_ranges[i][1];
    if (a < 0 || b < 0 || b > arg.n_vec || a > b)
      terminate("validate_argument: Specified numbers with -p option are not valid", 1);
  }
  if (arg.n_check_vec < 0 || arg.n_check_vec > arg.n_vec)
    terminate("validate_argument: Specified numbers with -c option are not valid", 1);
  if (arg.ortho_check_index_start < 0 || arg.ortho_check_index_end < 0 || arg.ortho_check_index_end > arg.n_vec ||
      arg.ortho_check_index_start > arg.ortho_check_index_end)
    terminate("validate_argument: Specified numbers with -t option are not valid", 1);
  if (arg.ngpu < 1 || (arg.ngpu & (arg.ngpu - 1)) != 0 || arg.ngpu > 8)
    terminate("validate_argument: --ngpu must be 1, 2, 4 or 8", 1);
}

// command_argument.f90:222-268,318-336: bytes per process; the B200 names follow the parallel formulas with
// n_procs = number of GPUs (host memory for the replicated inputs and the local eigenvector piece).
double required_memory(const ek_argument_t& arg) {
  const double dim = (double)arg.matrix_A_info.rows;
  const double np = (double)(arg.ngpu > 0 ? arg.ngpu : 1);
  if (!is_b200_solver(arg.solver_type)) return -1.0;  // unknown for this solver
  if (arg.is_generalized_problem) {
    double num_double = (double)(arg.matrix_A_info.entries + arg.matrix_B_info.entries);
    num_double += dim * dim * 3.0 / np;  // the input matrices (A and B) and the eigenvectors
    return 8.0 * num_double;
  }
  double num_double = (double)arg.matrix_A_info.entries;
  num_double += dim * dim * 2.0 / np;    // the input matrix and the eigenvectors
  return 8.0 * num_double;
}

static void print_matrix_info(const char* name, const ek_matrix_info_t& info) {
  printf("matrix %s field: %s\n", name, info.field.c_str());
  printf("matrix %s symm: %s\n", name, info.symm.c_str());
  printf("matrix %s rows: %lld\n", name, (long long)info.rows);
  printf("matrix %s cols: %lld\n", name, (long long)info.cols);
  printf("matrix %s entries: %lld\n", name, (long long)info.entries);
}

void print_command_argument(const ek_argument_t& arg) {
  printf(arg.is_generalized_problem ? "problem type: generalized\n" : "problem type: standard\n");
  printf("matrix A file: %s\n", arg.matrix_A_filename.c_str());
  print_matrix_info("A", arg.matrix_A_info);
  if (arg.is_generalized_problem) {
    printf("matrix B file: %s\n", arg.matrix_B_filename.c_str());
    print_matrix_info("B", arg.matrix_B_info);
  }
  printf("solver: %s\n", arg.solver_type.c_str());
  printf("eigenvalues output file: %s\n", arg.output_filename.c_str());
  printf("ipratios output file: %s\n", arg.ipratios_filename.c_str());
  printf("required eigenpairs: %lld\n", (long long)arg.n_vec);
  printf("verified eigenpairs: %lld\n", (long long)arg.n_check_vec);
  printf("log output file: %s\n", arg.log_filename.c_str());
}

}  // namespace ekapp
