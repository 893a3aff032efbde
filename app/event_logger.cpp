// ek_event_logger_m (reference src/event_logger.f90:23-141), the Fortran edit descriptors the output files are
// written with, and the fson printer (src/fson.f90:454-553) for the tree main.f90 builds.
//   add_event          : accumulate by name, a NEW name goes to the FRONT of the list; the master echoes
//                        `[Event<t F16.6>] <name>,<val E24.16e3>` on stderr unless to_print is false
//   print_events       : oldest first, list-directed `name  count  value` (called by terminate)
//   log_json_text      : {"setting": {...}, "events": [...]} with 2-space indent, integers I0, reals E24.16e3
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "ek_app.hpp"

namespace ekapp {

// ---------------------------------------------------------------- edit descriptors
static std::string right_align(const std::string& body, int width) {
  if (width <= 0 || (int)body.size() == width) return body;
  if ((int)body.size() > width) return std::string(width, '*');  // Fortran fills an overflowing field with '*'
  return std::string(width - body.size(), ' ') + body;
}

// Ew.dEe as gfortran prints it: mantissa in [0.1, 1), d digits after the point, e exponent digits
// (E26.16e3: '   -0.1121921212197622E+001').  expw = 0 means plain Ew.d: two exponent digits, or three with the
// 'E' dropped when |exponent| > 99.
std::string fortran_e(double x, int width, int digits, int expw) {
  if (isnan(x)) return right_align("NaN", width);
  if (isinf(x)) return right_align(x < 0 ? "-Infinity" : "Infinity", width);
  char mant[64];
  int e10 = 0;
  if (x == 0.0) {
    memset(mant, '0', digits);
    mant[digits] = 0;
  } else {
    char buf[80];
    snprintf(buf, sizeof buf, "%.*e", digits - 1, fabs(x));  // d.ddd...e+XX, correctly rounded to `digits` digits
    char* ep = strchr(buf, 'e');
    e10 = atoi(ep + 1) + 1;
    int q = 0;
    for (char* p = buf; p < ep; ++p)
      if (*p != '.') mant[q++] = *p;
    mant[q] = 0;
  }
  char ex[16];
  const int ae = e10 < 0 ? -e10 : e10;
  if (expw > 0) {
    snprintf(ex, sizeof ex, "E%c%0*d", e10 < 0 ? '-' : '+', expw, ae);
  } else if (ae <= 99) {
    snprintf(ex, sizeof ex, "E%c%02d", e10 < 0 ? '-' : '+', ae);
  } else {
    snprintf(ex, sizeof ex, "%c%03d", e10 < 0 ? '-' : '+', ae);
  }
  std::string body = std::string(signbit(x) ? "-" : "") + "0." + mant + ex;
  if ((int)body.size() > width && width > 0) body = std::string(signbit(x) ? "-" : "") + "." + mant + ex;  // optional 0 dropped
  return right_align(body, width);
}

std::string fortran_f(double x, int width, int digits) {
  char buf[400];
  snprintf(buf, sizeof buf, "%.*f", digits, x);
  return right_align(buf, width);
}

std::string fortran_i(long long v, int width) {
  char buf[32];
  snprintf(buf, sizeof buf, "%lld", v);
  return width > 0 ? right_align(buf, width) : std::string(buf);
}

// ---------------------------------------------------------------- event list
static std::vector<event_t> s_events;  // index 0 = newest name (head of the reference's linked list)

void add_event(const std::string& name, double val, bool to_print) {
  if (to_print && check_master()) {
    const double t = wtime() - g_wtime_init;
    fprintf(stderr, "[Event%s] %s,%s\n", fortran_f(t, 16, 6).c_str(), name.c_str(), fortran_e(val, 24, 16, 3).c_str());
  }
  for (auto& e : s_events)
    if (e.name == name) {
      e.num_repeated += 1;
      e.val += val;
      return;
    }
  s_events.insert(s_events.begin(), event_t{name, 1, val});
}

int num_events() { return (int)s_events.size(); }
const std::vector<event_t>& events() { return s_events; }
void clear_events() { s_events.clear(); }

// event_logger.f90:79-101: `print *, trim(name), num_repeated, val`, oldest first.  gfortran's list-directed
// output: a leading blank, the string, the integer in 12 columns, the real(8) as 1PG25.17-like with 17 digits.
void print_events() {
  for (int i = (int)s_events.size() - 1; i >= 0; --i) {
    const event_t& e = s_events[i];
    char real[96];
    const double a = fabs(e.val);
    if (e.val == 0.0 || (a >= 0.1 && a < 1e16)) {
      int lead = a < 1.0 ? 0 : (int)floor(log10(a)) + 1;  // digits before the point
      int dec = 17 - (lead > 0 ? lead : 1);
      if (dec < 0) dec = 0;
      snprintf(real, sizeof real, "%.*f    ", dec, e.val);
    } else {
      // list-directed output uses d.dddE+eee (one digit before the point)
      char tmp[64];
      snprintf(tmp, sizeof tmp, "%.16E", e.val);
      char* ep = strchr(tmp, 'E');
      int ex = atoi(ep + 1);
      *ep = 0;
      snprintf(real, sizeof real, "%sE%c%03d", tmp, ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
    }
    printf(" %s%s%s\n", e.name.c_str(), fortran_i(e.num_repeated, 12).c_str(), right_align(real, 26).c_str());
  }
  fflush(stdout);
}

// ---------------------------------------------------------------- log.json
static std::string json_string(const std::string& s) { return "\"" + s + "\""; }  // fson prints strings unescaped

std::string log_json_text(const ek_argument_t& arg, const std::vector<event_t>& evs) {
  std::string o = "{\n";
  o += "  \"setting\": {\n";
  o += "    \"version\": " + json_string(g_version) + ",\n";
  o += "    \"command\": " + json_string(arg.command) + ",\n";
  o += "    \"matrix_A_filename\": " + json_string(arg.matrix_A_filename) + ",\n";
  o += "    \"matrix_B_filename\": " + json_string(arg.matrix_B_filename) + ",\n";
  o += "    \"log_filename\": " + json_string(arg.log_filename) + ",\n";
  o += "    \"dimension\": " + fortran_i(arg.matrix_A_info.rows, 0) + ",\n";
  o += "    \"solver\": " + json_string(arg.solver_type) + ",\n";
  o += "    \"g_block_size\": " + fortran_i(g_block_size, 0) + ",\n";
  o += "    \"block_size\": " + fortran_i(arg.block_size, 0) + "\n";
  o += "  },\n";
  o += "  \"events\": [\n";
  for (size_t q = 0; q < evs.size(); ++q) {
    o += "    {\n";
    o += "      \"name\": " + json_string(evs[q].name) + ",\n";
    o += "      \"num_repeated\": " + fortran_i(evs[q].num_repeated, 0) + ",\n";
    o += "      \"val\": " + fortran_e(evs[q].val, 24, 16, 3) + "\n";
    o += std::string("    }") + (q + 1 < evs.size() ? "," : "") + "\n";
  }
  o += "  ]\n";
  o += "}\n";
  return o;
}

}  // namespace ekapp
