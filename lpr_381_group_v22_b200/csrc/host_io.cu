// host_io.cu -- the data formats either side of the pivot path, native host code:
//   * IO/InputFileParser.cs:19-68 (`max|min c...` / `a... rel rhs` / sign line) and the CLI's extra rows
//     (Program.cs:114-124, :511-535) -> lpr_model, built straight into a device tableau (lpr_tab_create_from_model:
//     no List<Constraint> -> double[,] -> H2D detour), plus a dense binary model format for the synthetic configs;
//   * Utilities/TableIterationFormater.cs:22-48 and NumFormat.N3 (Simplex/RevisedPrimalSimplexSolver.cs:451-465):
//     the text snapshot of a tableau with .NET Framework number formatting; for a device tableau the rows are
//     streamed D2H in blocks (pinned double buffer) while host threads format the previous block.
// No CUDA kernel lives here; the only device work is the row-block copy of lpr_tab_format.
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <sys/stat.h>
#include <string>
#include <thread>
#include <vector>

#include "tableau.cuh"

using namespace lpr;

struct lpr_model {
  bool loaded = false;       // ReadInputFile returned normally after parsing (not after one of its two early returns)
  std::string message;       // what ReadInputFile wrote to the console
  std::string problem_type;  // first token of the objective line, lower-cased
  std::vector<double> objective;
  struct Row {
    std::vector<double> coef;
    std::string relation;
    double rhs = 0.0;
  };
  std::vector<Row> rows;
  std::vector<std::string> signs;
};

namespace {

// ---- text helpers --------------------------------------------------------------------------------------
// char.IsWhiteSpace for the ASCII range plus NBSP/NEL (String.Trim() trims all Unicode white space; the multi-byte
// ones are matched as UTF-8 sequences below)
inline bool is_ws(unsigned char c) { return c == ' ' || (c >= 0x09 && c <= 0x0D); }
size_t utf8_ws_len(const std::string& s, size_t i) {  // length of the white-space character starting at i, 0 = none
  const unsigned char c = (unsigned char)s[i];
  if (is_ws(c)) return 1;
  if (c == 0xC2 && i + 1 < s.size() && ((unsigned char)s[i + 1] == 0x85 || (unsigned char)s[i + 1] == 0xA0)) return 2;
  if (c == 0xE2 && i + 2 < s.size()) {
    const unsigned char a = (unsigned char)s[i + 1], b = (unsigned char)s[i + 2];
    if (a == 0x80 && ((b >= 0x80 && b <= 0x8A) || b == 0xA8 || b == 0xA9 || b == 0xAF)) return 3;
    if (a == 0x81 && b == 0x9F) return 3;
  }
  if (c == 0xE1 && i + 2 < s.size() && (unsigned char)s[i + 1] == 0x9A && (unsigned char)s[i + 2] == 0x80) return 3;
  if (c == 0xE3 && i + 2 < s.size() && (unsigned char)s[i + 1] == 0x80 && (unsigned char)s[i + 2] == 0x80) return 3;
  return 0;
}
std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b) {
    const size_t l = utf8_ws_len(s, a);
    if (!l) break;
    a += l;
  }
  while (b > a) {  // step back over one white-space character at a time
    size_t l = 0;
    for (size_t w = 1; w <= 3 && w <= b - a; w++)
      if (utf8_ws_len(s, b - w) == w) l = w;
    if (!l) break;
    b -= l;
  }
  return s.substr(a, b - a);
}
std::vector<std::string> split_space(const std::string& s, bool remove_empty) {  // String.Split(' ')
  std::vector<std::string> out;
  size_t a = 0;
  while (true) {
    const size_t b = s.find(' ', a);
    std::string tok = s.substr(a, b == std::string::npos ? std::string::npos : b - a);
    if (!remove_empty || !tok.empty()) out.push_back(std::move(tok));
    if (b == std::string::npos) break;
    a = b + 1;
  }
  return out;
}
std::vector<std::string> read_all_lines(const char* text, size_t len) {  // File.ReadAllLines: \r\n, \n or \r
  std::vector<std::string> lines;
  size_t a = 0;
  if (len >= 3 && (unsigned char)text[0] == 0xEF && (unsigned char)text[1] == 0xBB && (unsigned char)text[2] == 0xBF) a = 3;
  std::string cur;
  bool pending = false;
  for (size_t i = a; i < len; i++) {
    const char c = text[i];
    if (c == '\r' || c == '\n') {
      lines.push_back(cur);
      cur.clear();
      pending = false;
      if (c == '\r' && i + 1 < len && text[i + 1] == '\n') i++;
    } else {
      cur.push_back(c);
      pending = true;
    }
  }
  if (pending) lines.push_back(cur);
  return lines;
}
std::string to_lower_ascii(std::string s) {
  for (char& c : s)
    if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
  return s;
}

// double.Parse(s, CultureInfo.InvariantCulture): NumberStyles.Float | AllowThousands -- white space around, one
// leading sign, digits with ',' group separators in the integer part, '.', exponent; or the NaN / Infinity symbols.
bool parse_net_double(const std::string& tok, double* out) {
  // Number parsing only skips the ASCII white space 0x09-0x0D and 0x20 (not the Unicode set String.Trim() removes)
  size_t a = 0, b = tok.size();
  while (a < b && is_ws((unsigned char)tok[a])) a++;
  while (b > a && is_ws((unsigned char)tok[b - 1])) b--;
  const std::string s = tok.substr(a, b - a);
  if (s.empty()) return false;
  if (s == "NaN") { *out = NAN; return true; }
  if (s == "Infinity") { *out = INFINITY; return true; }
  if (s == "-Infinity") { *out = -INFINITY; return true; }
  std::string clean;
  size_t i = 0;
  if (s[i] == '+' || s[i] == '-') clean.push_back(s[i++]);
  int digits = 0;
  while (i < s.size() && ((s[i] >= '0' && s[i] <= '9') || (s[i] == ',' && digits > 0))) {  // a group separator needs a digit before it
    if (s[i] != ',') { clean.push_back(s[i]); digits++; }
    i++;
  }
  if (i < s.size() && s[i] == '.') {
    clean.push_back(s[i++]);
    while (i < s.size() && s[i] >= '0' && s[i] <= '9') { clean.push_back(s[i++]); digits++; }
  }
  if (!digits) return false;
  if (i < s.size() && (s[i] == 'e' || s[i] == 'E')) {
    size_t j = i + 1;
    std::string ex = "e";
    if (j < s.size() && (s[j] == '+' || s[j] == '-')) ex.push_back(s[j++]);
    int ed = 0;
    while (j < s.size() && s[j] >= '0' && s[j] <= '9') { ex.push_back(s[j++]); ed++; }
    if (!ed) return false;
    clean += ex;
    i = j;
  }
  if (i != s.size()) return false;
  errno = 0;
  char* end = nullptr;
  const double v = strtod(clean.c_str(), &end);
  if (!end || *end != '\0') return false;
  if (std::isinf(v)) return false;  // OverflowException on .NET Framework
  *out = v;
  return true;
}

// ReadInputFile body once the lines are in memory (:27-66)
int parse_lines(const std::vector<std::string>& lines, lpr_model* m) {
  if (lines.size() < 3) {
    m->message = "The input file is not formatted correctly.";
    return LPR_OK;
  }
  const std::vector<std::string> obj = split_space(trim(lines[0]), false);
  m->problem_type = to_lower_ascii(obj[0]);
  for (size_t i = 1; i < obj.size(); i++) {
    double v;
    if (!parse_net_double(obj[i], &v))
      return fail(LPR_E_BADARG, "FormatException: objective coefficient %zu ('%s'): input string was not in a correct format",
                  i, obj[i].c_str());
    m->objective.push_back(v);
  }
  const size_t n = m->objective.size();
  for (size_t li = 1; li + 1 < lines.size(); li++) {
    const std::vector<std::string> parts = split_space(trim(lines[li]), true);
    // tokens are consumed left to right like the reference's loop (:50-57): a malformed coefficient throws
    // FormatException before a missing one throws IndexOutOfRangeException
    auto missing = [&](size_t idx) {
      return fail(LPR_E_BADARG, "IndexOutOfRangeException: constraint line %zu has %zu tokens, token %zu is needed", li,
                  parts.size(), idx + 1);
    };
    lpr_model::Row row;
    row.coef.resize(n);
    for (size_t j = 0; j < n; j++) {
      if (j >= parts.size()) return missing(j);
      if (!parse_net_double(parts[j], &row.coef[j]))
        return fail(LPR_E_BADARG, "FormatException: constraint line %zu, coefficient %zu ('%s')", li, j + 1, parts[j].c_str());
    }
    if (n >= parts.size()) return missing(n);
    row.relation = parts[n];
    if (n + 1 >= parts.size()) return missing(n + 1);
    if (!parse_net_double(parts[n + 1], &row.rhs))
      return fail(LPR_E_BADARG, "FormatException: constraint line %zu, right-hand side ('%s')", li, parts[n + 1].c_str());
    m->rows.push_back(std::move(row));
  }
  m->signs = split_space(trim(lines.back()), false);
  m->message = "Your file was read and is in the correct format!";
  m->loaded = true;
  return LPR_OK;
}

int copy_out(const std::string& s, char* out, int cap) {
  if (!out || cap < 1) return fail(LPR_E_BADARG, "null output buffer");
  if ((size_t)cap < s.size() + 1) return fail(LPR_E_CAPACITY, "output buffer too small (%zu bytes needed)", s.size() + 1);
  memcpy(out, s.c_str(), s.size() + 1);
  return LPR_OK;
}

// ---- .NET Framework number formatting -------------------------------------------------------------------
// Double -> NUMBER with 15 significant digits (what every double.ToString of the Framework starts from): digits
// d[0..14] and `scale` such that value = 0.d0d1... x 10^scale.
struct NetNumber {
  bool neg = false;
  int scale = 0;
  std::string digits;  // no trailing zeros; empty = zero
};
NetNumber to_number15(double x) {
  NetNumber nb;
  nb.neg = std::signbit(x);
  char buf[40];
  snprintf(buf, sizeof buf, "%.14e", fabs(x));  // d.dddddddddddddde+XX
  nb.digits.push_back(buf[0]);
  nb.digits.append(buf + 2, 14);
  nb.scale = atoi(buf + 17) + 1;
  while (!nb.digits.empty() && nb.digits.back() == '0') nb.digits.pop_back();
  if (nb.digits.empty()) nb.scale = 0;
  return nb;
}
// RoundNumber(number, pos): keep `pos` digits, round half up on the digit string
void round_number(NetNumber& nb, int pos) {
  if (pos < 0) {
    nb.digits.clear();
  } else if ((size_t)pos < nb.digits.size()) {
    const bool up = nb.digits[pos] >= '5';
    nb.digits.resize(pos);
    if (up) {
      int i = pos - 1;
      while (i >= 0 && nb.digits[i] == '9') i--;
      if (i < 0) {
        nb.digits = "1";
        nb.scale++;
      } else {
        nb.digits[i]++;
        nb.digits.resize(i + 1);
      }
    }
  }
  while (!nb.digits.empty() && nb.digits.back() == '0') nb.digits.pop_back();
  if (nb.digits.empty()) {
    nb.scale = 0;
    nb.neg = false;  // the Framework prints a rounded-to-zero negative without its sign
  }
}
// "F<decimals>"
void append_fixed(std::string& out, double x, int decimals) {
  if (std::isnan(x)) { out += "NaN"; return; }
  if (std::isinf(x)) { out += x > 0 ? "Infinity" : "-Infinity"; return; }
  NetNumber nb = to_number15(x);
  round_number(nb, nb.scale + decimals);
  if (nb.neg) out.push_back('-');
  if (nb.scale > 0) {
    for (int i = 0; i < nb.scale; i++) out.push_back((size_t)i < nb.digits.size() ? nb.digits[i] : '0');
  } else {
    out.push_back('0');
  }
  if (decimals > 0) {
    out.push_back('.');
    for (int k = 0; k < decimals; k++) {
      const int idx = nb.scale + k;  // digit index of the k-th decimal
      out.push_back(idx >= 0 && (size_t)idx < nb.digits.size() ? nb.digits[idx] : '0');
    }
  }
}
// double.ToString() ("G", 15 significant digits): NumFormat.N3's integral branch and CanonicalFormConverter
void append_general(std::string& out, double x) {
  if (std::isnan(x)) { out += "NaN"; return; }
  if (std::isinf(x)) { out += x > 0 ? "Infinity" : "-Infinity"; return; }
  NetNumber nb = to_number15(x);
  if (nb.digits.empty()) { out.push_back('0'); return; }
  if (nb.neg) out.push_back('-');
  if (nb.scale > 15 || nb.scale < -3) {  // scientific unless -5 < exponent < 15 (exponent = scale - 1): d.dddE+XX
    out.push_back(nb.digits[0]);
    if (nb.digits.size() > 1) {
      out.push_back('.');
      out.append(nb.digits, 1, std::string::npos);
    }
    char e[16];
    snprintf(e, sizeof e, "E%c%02d", nb.scale - 1 < 0 ? '-' : '+', abs(nb.scale - 1));
    out += e;
    return;
  }
  if (nb.scale <= 0) {
    out += "0.";
    out.append((size_t)(-nb.scale), '0');
    out += nb.digits;
  } else {
    for (int i = 0; i < nb.scale; i++) out.push_back((size_t)i < nb.digits.size() ? nb.digits[i] : '0');
    if (nb.digits.size() > (size_t)nb.scale) {
      out.push_back('.');
      out.append(nb.digits, nb.scale, std::string::npos);
    }
  }
}
// custom format "0.###"
void append_custom3(std::string& out, double x) {
  if (std::isnan(x)) { out += "NaN"; return; }
  if (std::isinf(x)) { out += x > 0 ? "Infinity" : "-Infinity"; return; }
  NetNumber nb = to_number15(x);
  round_number(nb, nb.scale + 3);
  if (nb.neg) out.push_back('-');
  if (nb.scale > 0) {
    for (int i = 0; i < nb.scale; i++) out.push_back((size_t)i < nb.digits.size() ? nb.digits[i] : '0');
  } else {
    out.push_back('0');
  }
  std::string dec;
  for (int k = 0; k < 3; k++) {
    const int idx = nb.scale + k;
    dec.push_back(idx >= 0 && (size_t)idx < nb.digits.size() ? nb.digits[idx] : '0');
  }
  while (!dec.empty() && dec.back() == '0') dec.pop_back();
  if (!dec.empty()) {
    out.push_back('.');
    out += dec;
  }
}
// Math.Round(x, 3, MidpointRounding.AwayFromZero) of the Framework: scale, split, bump on |fraction| >= 0.5, unscale
double net_round3_away(double x) {
  if (fabs(x) < 1e16) {
    double v = x * 1e3, ip;
    const double fr = modf(v, &ip);
    if (fabs(fr) >= 0.5) ip += (fr > 0) - (fr < 0);
    x = ip / 1e3;
  }
  return x;
}
void append_n3(std::string& out, double x) {  // NumFormat.N3 :455-465
  if (fabs(x) < 1e-12) x = 0.0;
  const double r = net_round3_away(x);
  const double ri = std::nearbyint(r);  // Math.Round(r): half to even
  if (fabs(r - ri) < 1e-12)
    append_general(out, ri);
  else
    append_custom3(out, r);
}

// rows [r0, r1) of the table body (row 0 is the "Z" row)
void format_rows(std::string& out, const double* tab, int64_t ld, int r0, int r1, int cols, int row_base,
                 const char* const* labels, int n_labels) {
  for (int i = r0; i < r1; i++) {
    const int gi = row_base + i;
    if (gi == 0) {
      out += "Z\t";
    } else {
      if (labels && n_labels >= gi && labels[gi - 1])
        out += labels[gi - 1];
      else
        out += std::to_string(gi);
      out.push_back('\t');
    }
    const double* row = tab + (size_t)i * ld;
    for (int j = 0; j < cols; j++) {
      append_fixed(out, row[j], 3);
      out.push_back('\t');
    }
    out += "\r\n";
  }
}
// the same, split over host threads (each formats a contiguous span of rows)
void format_rows_mt(std::string& out, const double* tab, int64_t ld, int nrows, int cols, int row_base,
                    const char* const* labels, int n_labels) {
  const int64_t cells = (int64_t)nrows * cols;
  int nt = (int)std::min<int64_t>(std::max(1u, std::min(16u, std::thread::hardware_concurrency())), cells / 16384 + 1);
  nt = std::max(1, std::min(nt, nrows));
  if (nt == 1) {
    format_rows(out, tab, ld, 0, nrows, cols, row_base, labels, n_labels);
    return;
  }
  std::vector<std::string> parts(nt);
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++) {
    const int a = (int)((int64_t)nrows * t / nt), b = (int)((int64_t)nrows * (t + 1) / nt);
    th.emplace_back([&, t, a, b] {
      parts[t].reserve((size_t)(b - a) * cols * 8);
      format_rows(parts[t], tab, ld, a, b, cols, row_base, labels, n_labels);
    });
  }
  for (auto& x : th) x.join();
  for (auto& p : parts) out += p;
}
void format_header(std::string& out, int cols, int num_original_vars, const char* title) {
  out += "\n";
  out += title ? title : "";
  out += ":\r\n";
  out.append(80, '-');
  out += "\r\n";
  out += "Table\t";
  for (int j = 0; j < num_original_vars; j++) {
    out += "x" + std::to_string(j + 1);
    out.push_back('\t');
  }
  for (int j = num_original_vars; j < cols - 1; j++) {
    out += "t" + std::to_string(j - num_original_vars + 1);
    out.push_back('\t');
  }
  out += "RHS\r\n";
}

thread_local std::string g_text;

// FormatCoeff :95-98
void append_coeff(std::string& out, double c) {
  if (c >= 0) out += "+ ";
  append_general(out, c);
}
// CanonicalFormConverter.CanonicalFormForFile :57-93 (AppendLine = "\r\n", the literal "\n"s are the source's)
void append_canonical_form(std::string& sb, const lpr_model* m) {
  sb += "\n=== Canonical Form ===\r\n";
  sb += "Z ";
  for (size_t i = 0; i < m->objective.size(); i++) {
    append_coeff(sb, m->objective[i] * -1);
    sb += "x" + std::to_string(i + 1) + " ";
  }
  sb += "= 0\n";
  for (size_t i = 0; i < m->rows.size(); i++) {
    const auto& r = m->rows[i];
    for (size_t j = 0; j < r.coef.size(); j++) {
      append_coeff(sb, r.coef[j]);
      sb += "x" + std::to_string(j + 1) + " ";
    }
    sb += "+ S" + std::to_string(i + 1) + " ";
    sb += "= ";
    append_general(sb, r.rhs);
    sb += "\n";
  }
  sb += "\nSign Restrictions: ";
  for (size_t i = 0; i < m->signs.size(); i++) sb += "x" + std::to_string(i + 1) + ": " + m->signs[i] + " ";
  sb += "\n======================\n\r\n";
}
std::string timestamp_or_now(const char* ts) {  // DateTime.Now:yyyy-MM-dd HH:mm:ss
  if (ts) return ts;
  char buf[32];
  const time_t t = time(nullptr);
  tm lt;
  localtime_r(&t, &lt);
  strftime(buf, sizeof buf, "%Y-%m-%d %H:%M:%S", &lt);
  return buf;
}
void append_final_results(std::string& sb, double final_z, const double* x, int n_x) {  // :66-73, :109-116
  sb += "=== Final Results ===\r\n";
  sb += "Z* = ";
  append_n3(sb, final_z);
  sb += "\r\n";
  for (int i = 0; x && i < n_x; i++) {
    sb += "x" + std::to_string(i + 1) + " = ";
    append_n3(sb, x[i]);
    sb += "\r\n";
  }
}
// EnsureDirectory + WriteToFile :122-137: File.WriteAllText / AppendAllText with Encoding.UTF8 -- the byte-order mark
// is written when the file starts empty, never when text is appended to existing content
int write_text_file(const char* path, const std::string& content, int append) {
  std::string p(path);
  for (size_t i = 1; i < p.size(); i++) {  // Directory.CreateDirectory of the parent, recursively
    if (p[i] == '/') {
      const std::string dir = p.substr(0, i);
      if (mkdir(dir.c_str(), 0777) != 0 && errno != EEXIST) return fail(LPR_E_BADARG, "cannot create directory '%s'", dir.c_str());
    }
  }
  bool bom = true;
  if (append) {
    struct stat st;
    if (stat(path, &st) == 0) bom = st.st_size == 0;
    else append = 0;
  }
  FILE* f = fopen(path, append ? "ab" : "wb");
  if (!f) return fail(LPR_E_BADARG, "cannot open '%s' for writing", path);
  bool ok = true;
  if (bom) ok = fwrite("\xEF\xBB\xBF", 1, 3, f) == 3;
  ok = ok && (content.empty() || fwrite(content.data(), 1, content.size(), f) == content.size());
  ok = (fclose(f) == 0) && ok;
  return ok ? LPR_OK : fail(LPR_E_BADARG, "short write to '%s'", path);
}
  // result of the last lpr_fmt_table / lpr_tab_format call of this thread

int model_to_arrays(const lpr_model* m, std::vector<double>& coef, std::vector<int>& cnt, std::vector<int>& rel,
                    std::vector<double>& rhs, int* stride) {
  const size_t n = m->objective.size();
  size_t st = std::max<size_t>(n, 1);
  for (const auto& r : m->rows) st = std::max(st, r.coef.size());
  coef.assign(m->rows.size() * st, 0.0);
  cnt.resize(m->rows.size());
  rel.resize(m->rows.size());
  rhs.resize(m->rows.size());
  for (size_t i = 0; i < m->rows.size(); i++) {
    const auto& r = m->rows[i];
    std::copy(r.coef.begin(), r.coef.end(), coef.begin() + i * st);
    cnt[i] = (int)r.coef.size();
    rel[i] = r.relation == ">=" ? LPR_REL_GE : (r.relation == "=" ? LPR_REL_EQ : LPR_REL_LE);
    rhs[i] = r.rhs;
  }
  *stride = (int)st;
  return LPR_OK;
}

constexpr char kMagic[8] = {'L', 'P', 'R', 'M', 'O', 'D', '1', '\0'};

}  // namespace

extern "C" {

// ---- model ------------------------------------------------------------------------------------------------
int lpr_model_parse_text(const char* text, int64_t len, lpr_model** out) {
  if (!out || (!text && len > 0) || len < 0) return fail(LPR_E_BADARG, "bad arguments");
  *out = nullptr;
  lpr_model* m = new lpr_model();
  const int rc = parse_lines(read_all_lines(text, (size_t)len), m);
  if (rc) {
    delete m;
    return rc;
  }
  *out = m;
  return LPR_OK;
}
int lpr_model_parse_file(const char* path, lpr_model** out) {
  if (!out || !path) return fail(LPR_E_BADARG, "bad arguments");
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) {  // :21-25: message, no exception, parser left empty
    lpr_model* m = new lpr_model();
    m->message = "Sorry, we can't find your file, please check it's in the right folser";
    *out = m;
    return LPR_OK;
  }
  std::string text;
  char buf[1 << 16];
  size_t k;
  while ((k = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, k);
  fclose(f);
  return lpr_model_parse_text(text.data(), (int64_t)text.size(), out);
}
int lpr_model_from_dense(int n, int m, const double* objective, const double* coef, const int* relation, const double* rhs,
                         int is_maximization, lpr_model** out) {
  if (!out || n < 0 || m < 0 || (n > 0 && !objective) || (m > 0 && (!rhs || (n > 0 && !coef))))
    return fail(LPR_E_BADARG, "bad dense model (n=%d m=%d)", n, m);
  lpr_model* md = new lpr_model();
  md->loaded = true;
  md->problem_type = is_maximization ? "max" : "min";
  md->objective.assign(objective, objective + n);
  md->rows.resize(m);
  for (int i = 0; i < m; i++) {
    md->rows[i].coef.assign(coef + (size_t)i * n, coef + (size_t)(i + 1) * n);
    const int r = relation ? relation[i] : LPR_REL_LE;
    md->rows[i].relation = r == LPR_REL_GE ? ">=" : (r == LPR_REL_EQ ? "=" : "<=");
    md->rows[i].rhs = rhs[i];
  }
  *out = md;
  return LPR_OK;
}
int lpr_model_destroy(lpr_model* m) {
  delete m;
  return LPR_OK;
}
int lpr_model_info(const lpr_model* m, int* loaded, int* n, int* n_constraints, int* n_signs) {
  if (!m) return fail(LPR_E_BADARG, "null model");
  if (loaded) *loaded = m->loaded ? 1 : 0;
  if (n) *n = (int)m->objective.size();
  if (n_constraints) *n_constraints = (int)m->rows.size();
  if (n_signs) *n_signs = (int)m->signs.size();
  return LPR_OK;
}
int lpr_model_problem_type(const lpr_model* m, char* out, int cap) {
  if (!m) return fail(LPR_E_BADARG, "null model");
  return copy_out(m->problem_type, out, cap);
}
int lpr_model_message(const lpr_model* m, char* out, int cap) {
  if (!m) return fail(LPR_E_BADARG, "null model");
  return copy_out(m->message, out, cap);
}
int lpr_model_objective(const lpr_model* m, double* c) {
  if (!m || (!c && !m->objective.empty())) return fail(LPR_E_BADARG, "null argument");
  std::copy(m->objective.begin(), m->objective.end(), c);
  return LPR_OK;
}
int lpr_model_constraint(const lpr_model* m, int i, double* coef, int cap, int* count, char* relation, int rel_cap,
                         double* rhs) {
  if (!m || i < 0 || (size_t)i >= m->rows.size()) return fail(LPR_E_BADARG, "bad constraint index %d", i);
  const auto& r = m->rows[i];
  if (count) *count = (int)r.coef.size();
  if (coef) {
    if ((size_t)cap < r.coef.size()) return fail(LPR_E_CAPACITY, "coefficient buffer too small (%zu needed)", r.coef.size());
    std::copy(r.coef.begin(), r.coef.end(), coef);
  }
  if (rhs) *rhs = r.rhs;
  if (relation) return copy_out(r.relation, relation, rel_cap);
  return LPR_OK;
}
int lpr_model_sign(const lpr_model* m, int j, char* out, int cap) {
  if (!m || j < 0 || (size_t)j >= m->signs.size()) return fail(LPR_E_BADARG, "bad sign restriction index %d", j);
  return copy_out(m->signs[j], out, cap);
}
int lpr_model_add_cli_bound_rows(lpr_model* m) {  // Program.cs:114-124 / :372-382
  if (!m) return fail(LPR_E_BADARG, "null model");
  const size_t n = m->objective.size();
  for (size_t i = 0; i < n; i++) {
    lpr_model::Row r;
    r.coef.assign(n + 3, 0.0);
    r.coef[i] = 1.0;
    r.coef[n + 1] = 1.0;  // the stray 1 of SURVEY Q1 (beyond the n coefficients the solver reads)
    r.relation = "<=";
    r.rhs = 1.0;
    m->rows.push_back(std::move(r));
  }
  return LPR_OK;
}
int lpr_model_add_upper_bound_rows(lpr_model* m) {  // Program.cs:511-535
  if (!m) return fail(LPR_E_BADARG, "null model");
  if (m->signs.empty()) return LPR_OK;
  const size_t n = m->objective.size();
  for (size_t j = 0; j < n; j++) {
    std::string s;
    for (char c : m->signs[std::min(j, m->signs.size() - 1)])
      if (c != ' ') s.push_back(c);
    const std::string low = to_lower_ascii(s);
    const bool is_bin = low.find("bin") != std::string::npos;
    const bool upper1 = s.find("\xE2\x89\xA4" "1") != std::string::npos || s.find("<=1") != std::string::npos;
    if (is_bin || upper1) {
      lpr_model::Row r;
      r.coef.assign(n, 0.0);
      r.coef[j] = 1.0;
      r.relation = "<=";
      r.rhs = 1.0;
      m->rows.push_back(std::move(r));
    }
  }
  return LPR_OK;
}
int lpr_tab_create_from_model(int device, const lpr_model* m, int is_maximization, lpr_tab** out) {
  if (!m || !out) return fail(LPR_E_BADARG, "null argument");
  if (!m->loaded) return fail(LPR_E_STATE, "the model was not loaded (%s)", m->message.c_str());
  std::vector<double> coef, rhs;
  std::vector<int> cnt, rel;
  int stride = 1;
  model_to_arrays(m, coef, cnt, rel, rhs, &stride);
  return lpr_tab_create_primal(device, (int)m->objective.size(), (int)m->rows.size(), m->objective.data(), coef.data(), stride,
                               cnt.data(), rel.data(), rhs.data(), is_maximization, out);
}
// dense binary model: magic, int32 n, m, n_signs, type length; type bytes; per sign: int32 length + bytes;
// objective n doubles; per row: int32 count, int32 relation length, relation bytes, rhs, count doubles
int lpr_model_save_binary(const lpr_model* m, const char* path) {
  if (!m || !path) return fail(LPR_E_BADARG, "null argument");
  FILE* f = fopen(path, "wb");
  if (!f) return fail(LPR_E_BADARG, "cannot open '%s' for writing", path);
  bool ok = fwrite(kMagic, 1, 8, f) == 8;
  auto w32 = [&](int32_t v) { ok = ok && fwrite(&v, 4, 1, f) == 1; };
  auto wstr = [&](const std::string& s) {
    w32((int32_t)s.size());
    ok = ok && (s.empty() || fwrite(s.data(), 1, s.size(), f) == s.size());
  };
  w32((int32_t)m->objective.size());
  w32((int32_t)m->rows.size());
  w32((int32_t)m->signs.size());
  w32(m->loaded ? 1 : 0);
  wstr(m->problem_type);
  for (const auto& s : m->signs) wstr(s);
  ok = ok && (m->objective.empty() || fwrite(m->objective.data(), 8, m->objective.size(), f) == m->objective.size());
  for (const auto& r : m->rows) {
    w32((int32_t)r.coef.size());
    wstr(r.relation);
    ok = ok && fwrite(&r.rhs, 8, 1, f) == 1;
    ok = ok && (r.coef.empty() || fwrite(r.coef.data(), 8, r.coef.size(), f) == r.coef.size());
  }
  ok = (fclose(f) == 0) && ok;
  return ok ? LPR_OK : fail(LPR_E_BADARG, "short write to '%s'", path);
}
int lpr_model_load_binary(const char* path, lpr_model** out) {
  if (!path || !out) return fail(LPR_E_BADARG, "null argument");
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return fail(LPR_E_BADARG, "cannot open '%s'", path);
  lpr_model* m = new lpr_model();
  bool ok = true;
  char magic[8];
  ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, kMagic, 8) == 0;
  auto r32 = [&]() -> int32_t {
    int32_t v = 0;
    ok = ok && fread(&v, 4, 1, f) == 1;
    return v;
  };
  auto rstr = [&]() -> std::string {
    const int32_t l = r32();
    std::string s;
    if (ok && l >= 0 && l < (1 << 20)) {
      s.resize(l);
      ok = l == 0 || fread(&s[0], 1, l, f) == (size_t)l;
    } else {
      ok = false;
    }
    return s;
  };
  const int32_t n = r32(), rows = r32(), ns = r32(), loaded = r32();
  ok = ok && n >= 0 && rows >= 0 && ns >= 0;
  if (ok) {
    m->loaded = loaded != 0;
    m->problem_type = rstr();
    for (int i = 0; ok && i < ns; i++) m->signs.push_back(rstr());
    m->objective.resize(n);
    ok = ok && (n == 0 || fread(m->objective.data(), 8, n, f) == (size_t)n);
    for (int i = 0; ok && i < rows; i++) {
      lpr_model::Row r;
      const int32_t cnt = r32();
      r.relation = rstr();
      ok = ok && cnt >= 0 && cnt < (1 << 28) && fread(&r.rhs, 8, 1, f) == 1;
      if (ok) {
        r.coef.resize(cnt);
        ok = cnt == 0 || fread(r.coef.data(), 8, cnt, f) == (size_t)cnt;
      }
      if (ok) m->rows.push_back(std::move(r));
    }
  }
  fclose(f);
  if (!ok) {
    delete m;
    return fail(LPR_E_BADARG, "'%s' is not a dense binary model file (or is truncated)", path);
  }
  *out = m;
  return LPR_OK;
}

// ---- RevisedPrimalSimplexSolver.CaptureSnapshot (Simplex/RevisedPrimalSimplexSolver.cs:294-387) --------------------
// Pure text: every number is already computed.  sb.AppendLine = "\r\n" (.NET Framework on Windows), VarLabel :289-292,
// NumFormat.N3 :455-465.  rcS may be NULL (= -y, :100-102 / :224-227).
}  // extern "C"
namespace lpr {
void format_revised_snapshot(std::string& sb, const char* title, bool is_min, int m, int n, const double* y,
                             const double* rcX, const double* rcS, int entering, double rc_pre, const double* u_pre,
                             const double* ratios_pre, const int* basis_pre, int leave_row, int leave_var_pre,
                             double z_working, double z_original, const double* BinvA, int64_t ldBA,
                             const double* Binv, int64_t ldB, const double* xB, const int* basis_post) {
  const char* NL = "\r\n";
  auto label = [&](int idx) {
    return idx < n ? "x" + std::to_string(idx + 1) : "S" + std::to_string(idx - n + 1);
  };
  auto joined = [&](const double* v, int cnt, bool neg) {
    for (int i = 0; i < cnt; i++) {
      if (i) sb.push_back('\t');
      append_n3(sb, neg ? -v[i] : v[i]);
    }
  };
  sb += title ? title : "";
  sb += NL;
  sb += "Current Tableau (Revised Simplex)";
  sb += NL;
  sb += is_min ? "Problem type: MIN (solving by MAX of -c)" : "Problem type: MAX";
  sb += NL;
  sb += NL;
  sb += "Dual prices (y = c_B^T B^{-1}):";
  sb += NL;
  joined(y, m, false);
  sb += NL;
  sb += NL;
  sb += "Reduced costs:";
  sb += NL;
  sb += "  x: ";
  joined(rcX, n, false);
  sb += NL;
  sb += "  s: ";
  if (rcS) joined(rcS, m, false); else joined(y, m, true);
  sb += NL;
  sb += NL;
  if (entering >= 0) {
    const std::string el = label(entering);
    sb += "Entering variable (chosen pre-pivot): " + el + "  (reduced cost pre = ";
    append_n3(sb, rc_pre);
    sb += ")";
    sb += NL;
    sb += "Direction u = B^{-1} a_enter (pre-pivot):";
    sb += NL;
    joined(u_pre, m, false);
    sb += NL;
    sb += NL;
    sb += "Ratio test (xB_i / u_i; \xE2\x88\x9E if u_i \xE2\x89\xA4 0)  [labels = pre-pivot basis]:";
    sb += NL;
    for (int i = 0; i < m; i++) {
      sb += label(basis_pre[i]) + ": ";
      if (std::isinf(ratios_pre[i]) && ratios_pre[i] > 0) sb += "\xE2\x88\x9E"; else append_n3(sb, ratios_pre[i]);
      sb += NL;
    }
    if (leave_row >= 0 && leave_var_pre >= 0) {
      sb += "Pivot (pre\xE2\x86\x92post): " + label(leave_var_pre) + "  \xE2\x86\x92  " + el + "    (pivot = ";
      append_n3(sb, u_pre[leave_row]);
      sb += ")";
      sb += NL;
      sb += NL;
    }
  }
  sb += "Working objective Z_working (maxified): ";
  append_n3(sb, z_working);
  sb += NL;
  sb += is_min ? "Original objective Z_original (MIN): " : "Original objective Z_original (MAX): ";
  append_n3(sb, z_original);
  sb += NL;
  sb += NL;
  sb += "Table\t";
  for (int j = 0; j < n; j++) sb += "x" + std::to_string(j + 1) + "\t";
  for (int j = 0; j < m; j++) sb += "S" + std::to_string(j + 1) + "\t";
  sb += "RHS";
  sb += NL;
  sb += "Z~\t";
  for (int j = 0; j < n; j++) {
    append_n3(sb, rcX[j]);
    sb.push_back('\t');
  }
  for (int j = 0; j < m; j++) {
    append_n3(sb, rcS ? rcS[j] : -y[j]);
    sb.push_back('\t');
  }
  append_n3(sb, z_working);
  sb += NL;
  for (int i = 0; i < m; i++) {
    sb += label(basis_post[i]) + "\t";
    for (int j = 0; j < n; j++) {
      append_n3(sb, BinvA[(size_t)i * ldBA + j]);
      sb.push_back('\t');
    }
    for (int j = 0; j < m; j++) {
      append_n3(sb, Binv[(size_t)i * ldB + j]);
      sb.push_back('\t');
    }
    append_n3(sb, xB[i]);
    sb += NL;
  }
  sb += "Basic Variables: ";
  for (int i = 0; i < m; i++) {
    if (i) sb += ", ";
    sb += label(basis_post[i]);
  }
  sb += NL;
}
std::string& thread_text() { return g_text; }
}  // namespace lpr
extern "C" {

/* $"{x:F6}" etc. with the .NET Framework rules (the "Final Tableau (Optimal)" summary must not use printf) */
int lpr_fmt_fixed(double x, int decimals, char* out, int cap) {
  if (decimals < 0 || decimals > 15) return fail(LPR_E_BADARG, "decimals out of range");
  std::string s;
  append_fixed(s, x, decimals);
  return copy_out(s, out, cap);
}

// ---- formatting ---------------------------------------------------------------------------------------------
int lpr_fmt_f3(double x, char* out, int cap) {
  std::string s;
  append_fixed(s, x, 3);
  return copy_out(s, out, cap);
}
int lpr_fmt_n3(double x, char* out, int cap) {
  std::string s;
  append_n3(s, x);
  return copy_out(s, out, cap);
}
int lpr_fmt_table(const double* tab, int rows, int cols, int64_t ld, int num_original_vars, const char* title,
                  const char* const* row_labels, int n_labels, const char** text, int64_t* len) {
  if (!tab || rows < 1 || cols < 1 || ld < cols || !text) return fail(LPR_E_BADARG, "bad table arguments");
  g_text.clear();
  g_text.reserve((size_t)rows * cols * 8 + 256);
  format_header(g_text, cols, num_original_vars, title);
  format_rows_mt(g_text, tab, ld, rows, cols, 0, row_labels, n_labels);
  *text = g_text.c_str();
  if (len) *len = (int64_t)g_text.size();
  return LPR_OK;
}
int lpr_tab_format(lpr_tab* h, int num_original_vars, const char* title, const char* const* row_labels, int n_labels,
                   const char** text, int64_t* len) {
  if (!h || !text) return fail(LPR_E_BADARG, "null argument");
  int rc = select_device(h->device);
  if (rc) return rc;
  const int R = h->R, Cc = h->C;
  // row blocks of about 8 MB, two pinned buffers: block b+1 crosses PCIe while block b is formatted
  const int br = (int)std::max<int64_t>(1, std::min<int64_t>(R, (8 << 20) / ((int64_t)Cc * 8)));
  double* pin[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  auto cleanup = [&] {
    for (int i = 0; i < 2; i++) {
      if (pin[i]) cudaFreeHost(pin[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
    }
  };
  for (int i = 0; i < 2; i++) {
    if (cudaMallocHost(&pin[i], sizeof(double) * (size_t)br * Cc) != cudaSuccess ||
        cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      cleanup();
      return fail(LPR_E_NOMEM, "snapshot staging allocation failed");
    }
  }
  g_text.clear();
  g_text.reserve((size_t)R * Cc * 8 + 256);
  format_header(g_text, Cc, num_original_vars, title);
  const int nblk = (R + br - 1) / br;
  auto issue = [&](int b) -> cudaError_t {
    const int r0 = b * br, nr = std::min(br, R - r0);
    cudaError_t e = cudaMemcpy2DAsync(pin[b & 1], sizeof(double) * Cc, h->T + (size_t)r0 * h->ld, sizeof(double) * h->ld,
                                      sizeof(double) * Cc, nr, cudaMemcpyDeviceToHost, h->stream);
    return e == cudaSuccess ? cudaEventRecord(ev[b & 1], h->stream) : e;
  };
  cudaError_t e = issue(0);
  for (int b = 0; b < nblk && e == cudaSuccess; b++) {
    if ((e = cudaEventSynchronize(ev[b & 1])) != cudaSuccess) break;
    if (b + 1 < nblk && (e = issue(b + 1)) != cudaSuccess) break;
    const int r0 = b * br, nr = std::min(br, R - r0);
    format_rows_mt(g_text, pin[b & 1], Cc, nr, Cc, r0, row_labels, n_labels);
  }
  if (e != cudaSuccess) {
    cudaStreamSynchronize(h->stream);
    cleanup();
    return fail(LPR_E_CUDA, "snapshot copy failed: %s", cudaGetErrorString(e));
  }
  cleanup();
  *text = g_text.c_str();
  if (len) *len = (int64_t)g_text.size();
  return LPR_OK;
}


// ---- result files (IO/OutputFileWrite.cs:16-137, Utilities/CanonicalFormConverter.cs:57-98) -------------------------
int lpr_fmt_general(double x, char* out, int cap) {  // double.ToString()
  std::string s;
  append_general(s, x);
  return copy_out(s, out, cap);
}
int lpr_model_canonical_form(const lpr_model* m, const char** text, int64_t* len) {
  if (!m || !text) return fail(LPR_E_BADARG, "null argument");
  g_text.clear();
  append_canonical_form(g_text, m);
  *text = g_text.c_str();
  if (len) *len = (int64_t)g_text.size();
  return LPR_OK;
}
int lpr_out_write_full_results(const char* path, const char* solver_used, const lpr_model* m, const char* const* snapshots,
                               int n_snapshots, double final_z, const double* x, int n_x, int append, const char* timestamp) {
  if (!path || !m || n_snapshots < 0 || (n_snapshots > 0 && !snapshots)) return fail(LPR_E_BADARG, "bad arguments");
  std::string sb;
  sb += "============================================================\r\n";
  sb += std::string("Solver: ") + (solver_used ? solver_used : "") + "\r\n";
  sb += "Problem type: " + m->problem_type + "\r\n";
  sb += "Timestamp: " + timestamp_or_now(timestamp) + "\r\n";
  sb += "============================================================\r\n";
  append_canonical_form(sb, m);
  if (n_snapshots > 0) {
    sb += "=== Iteration Snapshots ===\r\n";
    for (int i = 0; i < n_snapshots; i++) {
      sb += "--- Iteration " + std::to_string(i + 1) + " ---\r\n";
      sb += snapshots[i] ? snapshots[i] : "";
      sb += "\r\n";
    }
    sb += "\r\n";
  }
  append_final_results(sb, final_z, x, n_x);
  return write_text_file(path, sb, append);
}
int lpr_out_write_snapshots_only(const char* path, const char* solver_used, const char* const* snapshots, int n_snapshots,
                                 double final_z, const double* x, int n_x, int append, const char* timestamp) {
  if (!path || n_snapshots < 0 || (n_snapshots > 0 && !snapshots)) return fail(LPR_E_BADARG, "bad arguments");
  std::string sb;
  sb += "============================================================\r\n";
  sb += std::string("Solver: ") + (solver_used ? solver_used : "") + "\r\n";
  sb += "Timestamp: " + timestamp_or_now(timestamp) + "\r\n";
  sb += "============================================================\r\n";
  if (n_snapshots > 0) {
    sb += "=== Solver Log ===\r\n";
    for (int i = 0; i < n_snapshots; i++) {
      const std::string s = snapshots[i] ? snapshots[i] : "";
      sb += s + "\r\n";
      if (s.empty() || s.back() != '\n') sb += "\r\n";
    }
  }
  append_final_results(sb, final_z, x, n_x);
  return write_text_file(path, sb, append);
}

}  // extern "C"
