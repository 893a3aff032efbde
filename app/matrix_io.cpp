// ek_matrix_io_m (reference src/matrix_io.f90) and the header probe mminfo (src/mmio.f:341-585).
//   wrap_mminfo        command_argument.f90:89-103 -> mminfo: banner + size line, same ierr codes
//   read_matrix_file   matrix_io.f90:22-144: coordinate body, list-directed `i j value` records, range-checked,
//                      same events (read_matrix_file:allocate / :header / :value / read_matrix_file)
//   print_eigenvectors matrix_io.f90:173-285: <dir>/<j:08d>.dat, '(I8," ",I8," ",E26.16e3)' per element, or one
//                      Fortran unformatted sequential record with --binary
// the body is parsed by several threads (the file is split at record boundaries, records are counted
// per chunk, then every chunk is parsed into its final position), and eigenvector files are written by several
// threads, one file each -- at n = 32768 text I/O otherwise costs more than the solve.
#include <ctype.h>
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <thread>

#include "ek_app.hpp"

namespace ekapp {

static std::string lower(std::string s) {
  for (auto& c : s) c = (char)tolower((unsigned char)c);
  return s;
}
static std::vector<std::string> words(const std::string& line) {
  std::vector<std::string> w;
  size_t i = 0;
  while (i < line.size()) {
    while (i < line.size() && isspace((unsigned char)line[i])) ++i;
    size_t j = i;
    while (j < line.size() && !isspace((unsigned char)line[j])) ++j;
    if (j > i) w.push_back(line.substr(i, j - i));
    i = j;
  }
  return w;
}
static bool getline_file(FILE* f, std::string& out) {
  out.clear();
  int c;
  bool any = false;
  while ((c = fgetc(f)) != EOF) {
    any = true;
    if (c == '\n') break;
    out.push_back((char)c);
  }
  if (!out.empty() && out.back() == '\r') out.pop_back();
  return any;
}

// mminfo's return codes (mmio.f:413-583): 1 not a 'matrix', 3 no lines, 4 no data, 5/6 wrong size line,
// 7 bad banner, 8 bad representation, 9/10 bad field, 11 bad symmetry; open failure returns errno like iostat.
int wrap_mminfo(const std::string& filename, ek_matrix_info_t& minfo) {
  FILE* f = fopen(filename.c_str(), "r");
  if (!f) return errno ? errno : 2;
  std::string line;
  int ierr = 0;
  do {
    if (!getline_file(f, line)) { puts(" Premature end-of-file.\n No lines found."); ierr = 3; break; }
    std::vector<std::string> w = words(line);
    if (w.size() < 5 || w[0] != "%%MatrixMarket") {
      printf(" Invalid matrix header: %s\n Correct header format:\n %%%%MatrixMarket type representation field symmetry\n\n"
             " Check specification and try again.\n", line.c_str());
      ierr = 7;
      break;
    }
    if (lower(w[1]) != "matrix") {
      printf(" Invalid matrix type: %s\n This reader only understands type 'matrix'.\n", w[1].c_str());
      ierr = 1;
      break;
    }
    minfo.rep = lower(w[2]);
    minfo.field = lower(w[3]);
    minfo.symm = lower(w[4]);
    if (minfo.rep != "coordinate" && minfo.rep != "array") {
      printf(" '%s' representation not recognized.\n Recognized representations:\n    array\n    coordinate\n", minfo.rep.c_str());
      ierr = 8;
      break;
    }
    const std::string& fd = minfo.field;
    if (minfo.rep == "coordinate" && fd != "integer" && fd != "real" && fd != "complex" && fd != "pattern") {
      printf(" '%s' field is not recognized.\n Recognized fields:\n    real\n    complex\n    integer\n    pattern\n", fd.c_str());
      ierr = 9;
      break;
    }
    if (minfo.rep == "array" && fd != "integer" && fd != "real" && fd != "complex") {
      printf(" '%s' arrays are not recognized.\n Recognized fields:\n    real\n    complex\n    integer\n", fd.c_str());
      ierr = 10;
      break;
    }
    const std::string& sy = minfo.symm;
    if (sy != "general" && sy != "symmetric" && sy != "hermitian" && sy != "skew-symmetric") {
      printf(" '%s' symmetry is not recognized.\n Recognized symmetries:\n    general\n    symmetric\n    hermitian\n"
             "    skew-symmetric\n", sy.c_str());
      ierr = 11;
      break;
    }
    bool got = false;
    while (getline_file(f, line)) {
      if (line.empty() || line[0] != '%') { got = true; break; }
    }
    if (!got) { puts(" Premature end-of-file.\n No data found."); ierr = 4; break; }
    w = words(line);
    if (minfo.rep == "array") {
      if (w.size() != 2) {
        printf(" Size info inconsistant with representation.\n Array matrices need exactly 2 size descriptors.\n %zu were found.\n", w.size());
        ierr = 5;
        break;
      }
      minfo.rows = atoll(w[0].c_str());
      minfo.cols = atoll(w[1].c_str());
      if (sy == "symmetric" || sy == "hermitian") minfo.entries = (minfo.rows * minfo.cols - minfo.rows) / 2 + minfo.rows;
      else if (sy == "skew-symmetric") minfo.entries = (minfo.rows * minfo.cols - minfo.rows) / 2;
      else minfo.entries = minfo.rows * minfo.cols;
    } else {
      if (w.size() != 3) {
        printf(" Size info inconsistant with representation.\n Coordinate matrices need exactly 3 size descriptors.\n %zu were found.\n", w.size());
        ierr = 6;
        break;
      }
      minfo.rows = atoll(w[0].c_str());
      minfo.cols = atoll(w[1].c_str());
      minfo.entries = atoll(w[2].c_str());
    }
  } while (false);
  fclose(f);
  return ierr;
}

// ---------------------------------------------------------------- body parser
// One list-directed record `i j value`: items separated by blanks and/or one comma; the value may use a D exponent.
// Returns 0, or 1 for a malformed record.  *pp is advanced past the record's newline.
static inline bool is_sep(char c) { return c == ' ' || c == '\t' || c == ',' || c == '\r'; }

static int parse_record(const char*& p, const char* end, int64_t& i, int64_t& j, double& v) {
  auto skip = [&]() { while (p < end && is_sep(*p)) ++p; };
  auto parse_int = [&](int64_t& out) -> bool {
    skip();
    if (p >= end || *p == '\n') return false;
    bool neg = false;
    if (*p == '+' || *p == '-') { neg = *p == '-'; ++p; }
    if (p >= end || !isdigit((unsigned char)*p)) return false;
    int64_t x = 0;
    while (p < end && isdigit((unsigned char)*p)) { x = x * 10 + (*p - '0'); ++p; }
    if (p < end && !is_sep(*p) && *p != '\n') return false;
    out = neg ? -x : x;
    return true;
  };
  if (!parse_int(i) || !parse_int(j)) return 1;
  skip();
  if (p >= end || *p == '\n') return 1;
  char buf[64];
  int q = 0;
  while (p < end && !is_sep(*p) && *p != '\n' && q < 63) {
    char c = *p++;
    buf[q++] = (c == 'D' || c == 'd') ? 'E' : c;
  }
  buf[q] = 0;
  char* e = nullptr;
  v = strtod(buf, &e);
  if (e == buf || *e != 0) return 1;
  while (p < end && *p != '\n') ++p;  // further items of the record are ignored
  if (p < end) ++p;
  return 0;
}

static inline bool blank_line(const char* p, const char* end) {
  while (p < end && *p != '\n') {
    if (!is_sep(*p)) return false;
    ++p;
  }
  return true;
}

void read_matrix_file(const std::string& filename, const ek_matrix_info_t& info, ek_sparse_mat_t& matrix, int& ierr,
                      int threads) {
  const double time_start = wtime();
  double time_start_part = time_start;
  ierr = 0;
  if (check_master()) printf("start reading matrix file %s\n", filename.c_str());
  matrix.size = info.rows;
  matrix.num_non_zeros = info.entries;
  if (info.synthetic) {  // nothing on disk: the matrix is generated in HBM by the backend
    matrix.value.clear();
    matrix.suffix.clear();
    add_event("read_matrix_file", wtime() - time_start);
    return;
  }
  if (info.rep != "coordinate") terminate("read_matrix_file: only coordinate format is supported", 1);
  try {
    matrix.suffix.assign((size_t)2 * info.entries, 0);
    matrix.value.assign((size_t)info.entries, 0.0);
  } catch (...) {
    ierr = 1;
    return;
  }
  double time_end = wtime();
  add_event("read_matrix_file:allocate", time_end - time_start_part);
  time_start_part = time_end;

  int fd = open(filename.c_str(), O_RDONLY);
  if (fd < 0) { ierr = errno ? errno : 2; return; }
  struct stat st;
  fstat(fd, &st);
  const size_t fsize = (size_t)st.st_size;
  const char* base = fsize ? (const char*)mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
  if (fsize && base == MAP_FAILED) { close(fd); ierr = errno ? errno : 5; return; }
  const char* end = base + fsize;
  // read_matrix_file_header (matrix_io.f90:73-88): line 1, then lines starting with '%', then the size line
  const char* p = base;
  auto next_line = [&](const char* q) { while (q < end && *q != '\n') ++q; return q < end ? q + 1 : end; };
  p = next_line(p);
  while (p < end && *p == '%') p = next_line(p);
  p = next_line(p);  // the size line itself (already known from mminfo)
  time_end = wtime();
  add_event("read_matrix_file:header", time_end - time_start_part);
  time_start_part = time_end;

  const int64_t nnz = info.entries;
  int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (T < 1) T = 1;
  if ((size_t)(end - p) < ((size_t)1 << 20) || nnz < 4096) T = 1;
  if (T > 64) T = 64;
  // chunk boundaries on record starts
  std::vector<const char*> cb(T + 1);
  cb[0] = p;
  cb[T] = end;
  for (int t = 1; t < T; ++t) {
    const char* q = p + (size_t)(end - p) * t / T;
    if (q < cb[t - 1]) q = cb[t - 1];
    cb[t] = (q == p) ? p : next_line(q - 1);
  }
  std::vector<int64_t> count(T + 1, 0);
  {
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t]() {
        int64_t c = 0;
        for (const char* q = cb[t]; q < cb[t + 1]; q = next_line(q))
          if (!blank_line(q, cb[t + 1])) ++c;
        count[t + 1] = c;
      });
    for (auto& x : th) x.join();
  }
  for (int t = 0; t < T; ++t) count[t + 1] += count[t];
  std::atomic<int> bad_format{0}, bad_range{0};
  const int64_t size = info.rows;
  {
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t]() {
        int64_t line = count[t];
        const char* q = cb[t];
        const char* qe = cb[t + 1];
        while (q < qe && line < nnz) {
          if (blank_line(q, qe)) { q = next_line(q); continue; }  // list-directed input skips empty records
          int64_t i, j;
          double v;
          if (parse_record(q, qe, i, j, v)) { bad_format.store(1); return; }
          if (i < 1 || i > size || j < 1 || j > size) { bad_range.store(1); return; }
          matrix.value[(size_t)line] = v;
          matrix.suffix[(size_t)2 * line] = (int32_t)i;
          matrix.suffix[(size_t)2 * line + 1] = (int32_t)j;
          ++line;
        }
      });
    for (auto& x : th) x.join();
  }
  const int64_t found = count[T];
  if (base) munmap((void*)base, fsize);
  close(fd);
  if (bad_format.load() || found < nnz)  // a short file ends the reference's read with iostat /= 0
    terminate("read_matrix_file_value: invalid format of matrix value", bad_format.load() ? 5010 : -1);
  if (bad_range.load()) terminate("read_matrix_file_value: index of matrix out of range", 0 + 1);
  time_end = wtime();
  add_event("read_matrix_file:value", time_end - time_start_part);
  add_event("read_matrix_file", time_end - time_start);
}

// ---------------------------------------------------------------- eigenvector files
static void write_vector_file(const std::string& filename, const double* col, int64_t m, int64_t j, bool binary) {
  FILE* f = fopen(filename.c_str(), binary ? "wb" : "w");
  if (!f) {
    printf(" iostat: %d\n", errno);
    terminate("print_eigenvectors: cannot open " + filename, errno ? errno : 1);
  }
  if (binary) {
    // Fortran unformatted sequential record (gfortran): 4-byte length marker, payload, marker
    const int32_t mark = (int32_t)(m * 8);
    fwrite(&mark, 4, 1, f);
    fwrite(col, 8, (size_t)m, f);
    fwrite(&mark, 4, 1, f);
  } else {
    std::string buf;
    buf.reserve((size_t)m * 45);
    const std::string js = fortran_i(j, 8);
    for (int64_t i = 0; i < m; ++i) {
      buf += fortran_i(i + 1, 8);
      buf += ' ';
      buf += js;
      buf += ' ';
      buf += fortran_e(col[i], 26, 16, 3);
      buf += '\n';
    }
    fwrite(buf.data(), 1, buf.size(), f);
  }
  fclose(f);
}

void print_eigenvectors(const ek_argument_t& arg, const ek_eigenpairs_types_union_t& eigenpairs) {
  const double time_start = wtime();
  if (eigenpairs.type_number == 1) {
    terminate("print_eigenvectors: printer for a local matrix not implemented yet", 1);
  } else if (eigenpairs.type_number == 2) {
    const ek_eigenpairs_blacs_t& ep = eigenpairs.blacs;
    const int64_t m = ep.desc[rows_];
    // the rank that owns column j writes its file (the reference ships the column to the printing process instead)
    std::vector<int64_t> mine;
    for (int i = 0; i < arg.num_printed_vecs_ranges; ++i)
      for (int64_t j = arg.printed_vecs_ranges[i][0]; j <= arg.printed_vecs_ranges[i][1]; ++j)
        if (j >= 1 && j - 1 >= ep.col0 && j - 1 < ep.col0 + ep.loc_cols) mine.push_back(j);
    int T = arg.io_threads > 0 ? arg.io_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if ((int64_t)T > (int64_t)mine.size()) T = (int)mine.size();
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    std::string fail_msg;
    int fail_code = 0;
    auto work = [&]() {
      while (true) {
        const size_t q = next.fetch_add(1);
        if (q >= mine.size() || failed.load()) return;
        const int64_t j = mine[q];
        char num[32];
        snprintf(num, sizeof num, "%08lld", (long long)j);
        const std::string filename = arg.eigenvector_dir + "/" + num + ".dat";
        try {
          write_vector_file(filename, ep.Vectors + (size_t)(j - 1 - ep.col0) * ep.lld, m, j, arg.is_binary_output);
        } catch (const Terminate& t) {
          if (!failed.exchange(1)) { fail_msg = t.message; fail_code = t.code; }
          return;
        }
      }
    };
    for (int64_t j : mine)
      fprintf(stderr, "[Event%s] print eigenvector %lld on process (0, %d)\n", fortran_f(wtime() - g_wtime_init, 16, 6).c_str(),
              (long long)j, world_rank());
    if (T <= 1) {
      work();
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < T; ++t) th.emplace_back(work);
      for (auto& x : th) x.join();
    }
    if (failed.load()) terminate(fail_msg, fail_code);
  }
  add_event("print_eigenvectors", wtime() - time_start);
}

}  // namespace ekapp
