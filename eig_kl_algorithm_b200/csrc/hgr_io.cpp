// hgr_io.cpp -- the text formats of the reference's file surface, host side.
//   .hgr input            cEIG.cpp:178-182,94-101 ; cKL.cpp:92-115
//   pre_saved_EIG file    written by cEIG.cpp:213-220, read by cKL.cpp:155-174
//   KL trace              cKL.cpp:315,380
#include "internal.h"
#include <cerrno>
#include <cstdlib>
#include <fstream>
#include <memory>

namespace eigkl {

namespace {
struct FileCloser { void operator()(FILE *f) const { if (f) fclose(f); } };
using File = std::unique_ptr<FILE, FileCloser>;

std::string slurp(const char *path) {
  File f(fopen(path, "rb"));
  if (!f) throw Error(EIGKL_E_IO, std::string("Error opening input file: ") + path);
  std::string s;
  char buf[1 << 16];
  size_t got;
  while ((got = fread(buf, 1, sizeof(buf), f.get())) > 0) s.append(buf, got);
  return s;
}
inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }
}  // namespace

// Header "<nets> <nodes>" (further header tokens ignored), then exactly <nets> lines; each line is
// read like `while (ss >> node)`: whitespace separated unsigned integers, stopping at the first
// token that is not one.  Ids are 1-based in the file, 0-based here.
void parse_hgr(const char *path, HostHgr &out) {
  const std::string text = slurp(path);
  const char *p = text.data(), *end = p + text.size();
  auto read_uint = [&](long long &v) -> bool {
    while (p < end && is_blank(*p)) ++p;
    if (p >= end || *p < '0' || *p > '9') return false;
    long long x = 0;
    while (p < end && *p >= '0' && *p <= '9') { x = x * 10 + (*p - '0'); if (x > (1ll << 40)) x = (1ll << 40); ++p; }
    v = x;
    return true;
  };
  long long nets = 0, nodes = 0;
  if (!read_uint(nets) || !read_uint(nodes)) throw Error(EIGKL_E_FORMAT, std::string("bad .hgr header in ") + path);
  if (nodes <= 0 || nodes > 2147483647ll / 2 || nets > 2147483647ll / 2) throw Error(EIGKL_E_FORMAT, "unsupported .hgr dimensions");
  while (p < end && *p != '\n') ++p;
  if (p < end) ++p;
  out.n_nets = (int32_t)nets;
  out.n_nodes = (int32_t)nodes;
  out.net_off.assign((size_t)nets + 1, 0);
  out.pins.clear();
  out.pins.reserve(text.size() / 3);
  for (long long e = 0; e < nets; ++e) {
    while (p < end && *p != '\n') {
      long long v;
      if (read_uint(v)) {
        if (v < 1 || v > nodes) throw Error(EIGKL_E_FORMAT, "pin id out of range [1, nodes] in net " + std::to_string(e + 1));
        out.pins.push_back((int32_t)(v - 1));
      } else {
        while (p < end && *p != '\n') ++p;       // non-numeric token ends the net, as operator>> would
      }
    }
    if (p < end) ++p;
    out.net_off[(size_t)e + 1] = (int64_t)out.pins.size();
  }
}

// lambda2, median, then "i\tside\tv_i", floats with 12 significant digits (ostream << setprecision(12))
void write_eig_file(const char *path, double lambda2, double median, const double *vec, int32_t n) {
  File f(fopen(path, "w"));
  if (!f) throw Error(EIGKL_E_IO, std::string("Error opening output file: ") + path);
  std::string buf;
  buf.reserve((size_t)n * 32 + 64);
  char tmp[96];
  int len = snprintf(tmp, sizeof(tmp), "%.12g\n%.12g\n", lambda2, median);
  buf.append(tmp, (size_t)len);
  for (int32_t i = 0; i < n; ++i) {
    len = snprintf(tmp, sizeof(tmp), "%d\t%d\t%.12g\n", i, (median > vec[i]) ? 1 : 0, vec[i]);
    buf.append(tmp, (size_t)len);
  }
  if (fwrite(buf.data(), 1, buf.size(), f.get()) != buf.size()) throw Error(EIGKL_E_IO, std::string("short write: ") + path);
}

// cKL.cpp:162-173: two lines skipped, then "node side weight" per line; nodes are appended to
// remain[side] in FILE order.  `side` gets the per-node side; the orders are returned so that a file
// that is not in ascending node order keeps the reference's tie-breaking.
void read_eig_file(const char *path, int32_t n, std::vector<uint8_t> &side) {
  std::ifstream in(path);
  if (!in.is_open()) throw Error(EIGKL_E_IO, "Error: EIG file not found");
  std::string line;
  std::getline(in, line);
  std::getline(in, line);
  side.assign((size_t)n, 0xFF);
  int64_t rows = 0;
  long long prev = -1;
  while (std::getline(in, line)) {
    char *q = nullptr;
    const char *s = line.c_str();
    errno = 0;
    long long node = strtoll(s, &q, 10);
    if (q == s) continue;                                      // blank line
    const char *s2 = q;
    long long sd = strtoll(s2, &q, 10);
    if (q == s2) throw Error(EIGKL_E_FORMAT, std::string("malformed EIG row in ") + path);
    if (node < 0 || node >= n || (sd != 0 && sd != 1)) throw Error(EIGKL_E_FORMAT, std::string("EIG row out of range in ") + path);
    if (node <= prev) throw Error(EIGKL_E_FORMAT, std::string("EIG rows are not in ascending node order in ") + path);
    prev = node;
    side[(size_t)node] = (uint8_t)sd;
    ++rows;
  }
  if (rows != n) throw Error(EIGKL_E_FORMAT, std::string("EIG file has ") + std::to_string(rows) + " rows, expected " + std::to_string(n));
}

// row 0 "0\t<cut>\t0", then "<iter>\t<cut>\t<gain>" with default ostream float formatting (%g, 6 digits)
void write_trace_file(const char *path, const eigkl_trace *t) {
  if (!t || !t->cut || !t->gain) throw Error(EIGKL_E_ARG, "eigkl_write_trace: trace needs cut and gain");
  File f(fopen(path, "w"));
  if (!f) throw Error(EIGKL_E_IO, "Error: Cannot open output file");
  std::string buf;
  char tmp[96];
  int len = snprintf(tmp, sizeof(tmp), "0\t%g\t0\n", (double)t->cut[0]);
  buf.append(tmp, (size_t)len);
  for (int64_t i = 1; i <= t->swaps; ++i) {
    len = snprintf(tmp, sizeof(tmp), "%lld\t%g\t%g\n", (long long)i, (double)t->cut[i], (double)t->gain[i]);
    buf.append(tmp, (size_t)len);
  }
  if (fwrite(buf.data(), 1, buf.size(), f.get()) != buf.size()) throw Error(EIGKL_E_IO, std::string("short write: ") + path);
}

}  // namespace eigkl
