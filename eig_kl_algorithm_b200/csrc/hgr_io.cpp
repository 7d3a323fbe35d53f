// hgr_io.cpp -- the text formats of the reference's file surface, host side.
//   .hgr input            cEIG.cpp:178-182,94-101 ; cKL.cpp:92-115
//   pre_saved_EIG file    written by cEIG.cpp:213-220, read by cKL.cpp:155-174
//   KL trace              cKL.cpp:315,380
#include "internal.h"
#include <cerrno>
#include <cstdlib>
#include <fstream>
#include <memory>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <thread>

namespace eigkl {

namespace {
struct FileCloser { void operator()(FILE *f) const { if (f) fclose(f); } };
using File = std::unique_ptr<FILE, FileCloser>;

// read-only view of a whole file: mmap, no copy (a 2 M-node .hgr is 38 MB; the reference reads it through
// ifstream + stringstream line by line, cKL.cpp:92-115)
struct Mapped {
  const char *p = nullptr;
  size_t n = 0;
  int fd = -1;
  bool mapped = false;
  std::string owned;                       // fallback when the file cannot be mapped (a pipe, an empty file)
  Mapped(const char *path, const char *what, bool append_path) {
    fd = open(path, O_RDONLY);
    if (fd < 0) throw Error(EIGKL_E_IO, append_path ? std::string(what) + path : std::string(what));
    struct stat sb;
    if (fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
      void *m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m != MAP_FAILED) {
        madvise(m, (size_t)sb.st_size, MADV_SEQUENTIAL);
        p = static_cast<const char *>(m); n = (size_t)sb.st_size; mapped = true;
        return;
      }
    }
    char buf[1 << 16];
    ssize_t got;
    while ((got = read(fd, buf, sizeof(buf))) > 0) owned.append(buf, (size_t)got);
    p = owned.data(); n = owned.size();
  }
  ~Mapped() {
    if (mapped) munmap(const_cast<char *>(p), n);
    if (fd >= 0) close(fd);
  }
  Mapped(const Mapped &) = delete;
  Mapped &operator=(const Mapped &) = delete;
};

inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

int io_threads(size_t bytes) {
  if (const char *e = getenv("EIGKL_IO_THREADS")) return std::max(1, atoi(e));
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t by_size = bytes / ((size_t)1 << 20) + 1;              // one thread per MiB of text
  return (int)std::min<size_t>(std::min<size_t>(hw, 16), by_size);
}

// [begin, end) cut into `parts` pieces that start right after a newline
std::vector<const char *> split_at_newlines(const char *begin, const char *end, int parts) {
  std::vector<const char *> cut((size_t)parts + 1);
  cut[0] = begin;
  cut[(size_t)parts] = end;
  for (int t = 1; t < parts; ++t) {
    const char *q = begin + (size_t)(end - begin) * (size_t)t / (size_t)parts;
    if (q < cut[(size_t)t - 1]) q = cut[(size_t)t - 1];
    while (q < end && *q != '\n') ++q;
    if (q < end) ++q;
    cut[(size_t)t] = q;
  }
  return cut;
}

template <typename F>
void run_parallel(int parts, F &&f) {
  if (parts <= 1) { f(0); return; }
  std::vector<std::thread> th;
  th.reserve((size_t)parts - 1);
  for (int t = 1; t < parts; ++t) th.emplace_back([&f, t] { f(t); });
  f(0);
  for (auto &x : th) x.join();
}

// One line of a .hgr body, read like `while (ss >> node)` (cKL.cpp:108-111): whitespace separated unsigned
// integers, stopping at the first token that is not one.  emit(v) gets every id; returns the line's end.
template <typename Emit>
inline const char *scan_net_line(const char *p, const char *end, Emit &&emit) {
  while (p < end && *p != '\n') {
    while (p < end && is_blank(*p)) ++p;
    if (p >= end || *p == '\n') break;
    if (*p < '0' || *p > '9') {                   // non-numeric token ends the net, as operator>> would
      while (p < end && *p != '\n') ++p;
      break;
    }
    long long x = 0;
    while (p < end && *p >= '0' && *p <= '9') { x = x * 10 + (*p - '0'); if (x > (1ll << 40)) x = (1ll << 40); ++p; }
    emit(x);
  }
  return p;
}
}  // namespace

// Header "<nets> <nodes>" (further header tokens ignored), then exactly <nets> lines (missing lines are empty
// nets, extra lines are ignored).  Ids are 1-based in the file, 0-based here.
// Parallel: the body is cut at newlines into one piece per thread; pass 1 counts lines and ids per piece, a
// prefix sum gives every piece its first net and first pin slot, pass 2 parses in place.  The result does not
// depend on the number of threads (tests/test_host.py).
void parse_hgr(const char *path, HostHgr &out) {
  Mapped file(path, "Error opening input file: ", true);
  const char *p = file.p, *end = p + file.n;
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
  const int T = io_threads((size_t)(end - p));
  const std::vector<const char *> cut = split_at_newlines(p, end, T);
  std::vector<long long> lines((size_t)T + 1, 0), ids((size_t)T + 1, 0);
  // pass 1: lines and ids per piece (a last line without '\n' counts as a line when it is not empty)
  run_parallel(T, [&](int t) {
    const char *q = cut[(size_t)t], *qe = cut[(size_t)t + 1];
    long long nl = 0, ni = 0;
    while (q < qe) {
      q = scan_net_line(q, qe, [&](long long) { ++ni; });
      ++nl;
      if (q < qe) ++q;
    }
    lines[(size_t)t + 1] = nl; ids[(size_t)t + 1] = ni;
  });
  for (int t = 0; t < T; ++t) { lines[(size_t)t + 1] += lines[(size_t)t]; ids[(size_t)t + 1] += ids[(size_t)t]; }
  // ids of lines beyond <nets> are parsed into the tail of pins[] and dropped afterwards
  std::vector<int32_t> pins((size_t)ids[(size_t)T]);
  std::vector<long long> bad((size_t)T, -1);
  int64_t *off = out.net_off.data();
  run_parallel(T, [&](int t) {
    const char *q = cut[(size_t)t], *qe = cut[(size_t)t + 1];
    long long e = lines[(size_t)t], k = ids[(size_t)t];
    while (q < qe) {
      q = scan_net_line(q, qe, [&](long long v) {
        if ((v < 1 || v > nodes) && e < nets && bad[(size_t)t] < 0) bad[(size_t)t] = e;
        pins[(size_t)k++] = (int32_t)(v - 1);
      });
      ++e;
      if (e <= nets) off[e] = k;
      if (q < qe) ++q;
    }
  });
  for (int t = 0; t < T; ++t)
    if (bad[(size_t)t] >= 0) throw Error(EIGKL_E_FORMAT, "pin id out of range [1, nodes] in net " + std::to_string(bad[(size_t)t] + 1));
  const long long have = std::min<long long>(lines[(size_t)T], nets);
  for (long long e = have + 1; e <= nets; ++e) off[e] = off[have];       // missing lines: empty nets
  pins.resize((size_t)off[nets]);
  out.pins.swap(pins);
}

// lambda2, median, then "i\tside\tv_i", floats with 12 significant digits (ostream << setprecision(12), one
// `endl` flush per row in the reference, cEIG.cpp:213-220).  The rows are formatted by several threads into
// per-thread buffers and written in order with one write each.
void write_eig_file(const char *path, double lambda2, double median, const double *vec, int32_t n) {
  File f(fopen(path, "w"));
  if (!f) throw Error(EIGKL_E_IO, std::string("Error opening output file: ") + path);
  const int T = io_threads((size_t)n * 28);
  std::vector<std::string> part((size_t)T);
  run_parallel(T, [&](int t) {
    const int32_t lo = (int32_t)((int64_t)n * t / T), hi = (int32_t)((int64_t)n * (t + 1) / T);
    std::string &buf = part[(size_t)t];
    buf.reserve((size_t)(hi - lo) * 32 + 64);
    char tmp[96];
    if (t == 0) {
      const int len = snprintf(tmp, sizeof(tmp), "%.12g\n%.12g\n", lambda2, median);
      buf.append(tmp, (size_t)len);
    }
    for (int32_t i = lo; i < hi; ++i) {
      const int len = snprintf(tmp, sizeof(tmp), "%d\t%d\t%.12g\n", i, (median > vec[i]) ? 1 : 0, vec[i]);
      buf.append(tmp, (size_t)len);
    }
  });
  for (int t = 0; t < T; ++t)
    if (fwrite(part[(size_t)t].data(), 1, part[(size_t)t].size(), f.get()) != part[(size_t)t].size())
      throw Error(EIGKL_E_IO, std::string("short write: ") + path);
}

// cKL.cpp:162-173: two lines skipped, then "node side weight" per line; nodes are appended to
// remain[side] in FILE order.  `side` gets the per-node side.  Every file cEIG (or this library) writes lists
// the nodes in ascending order; for any other file `ascending` comes back false and order0 / order1 hold
// remain[0] / remain[1] in file order, so that the pair selection keeps the reference's tie-breaking.
// Every node must appear exactly once (the reference would silently drop a missing node from both lists).
// The rows are parsed by several threads (pieces cut at newlines), then merged in file order.
void read_eig_file(const char *path, int32_t n, std::vector<uint8_t> &side, std::vector<int32_t> &order0,
                   std::vector<int32_t> &order1, bool &ascending) {
  Mapped file(path, "Error: EIG file not found", false);       // the reference's message, cKL.cpp:157-160
  const char *p = file.p, *end = p + file.n;
  for (int skip = 0; skip < 2; ++skip) {                        // lambda2, median
    while (p < end && *p != '\n') ++p;
    if (p < end) ++p;
  }
  const int T = io_threads((size_t)(end - p));
  const std::vector<const char *> cut = split_at_newlines(p, end, T);
  struct Piece { std::vector<int32_t> node; std::vector<uint8_t> sd; int err = 0; };
  std::vector<Piece> pc((size_t)T);
  run_parallel(T, [&](int t) {
    const char *q = cut[(size_t)t], *qe = cut[(size_t)t + 1];
    Piece &P = pc[(size_t)t];
    P.node.reserve((size_t)(qe - q) / 20 + 16); P.sd.reserve((size_t)(qe - q) / 20 + 16);
    auto read_int = [&](long long &v) -> bool {                 // as strtoll: optional sign, decimal digits
      while (q < qe && is_blank(*q)) ++q;
      bool neg = false;
      if (q < qe && (*q == '-' || *q == '+')) { neg = *q == '-'; ++q; }
      if (q >= qe || *q < '0' || *q > '9') return false;
      long long x = 0;
      while (q < qe && *q >= '0' && *q <= '9') { x = x * 10 + (*q - '0'); if (x > (1ll << 40)) x = (1ll << 40); ++q; }
      v = neg ? -x : x;
      return true;
    };
    while (q < qe) {
      long long node = 0, sd = 0;
      if (read_int(node)) {
        if (!read_int(sd)) P.err |= 1;
        else if (node < 0 || node >= n || (sd != 0 && sd != 1)) P.err |= 2;
        else { P.node.push_back((int32_t)node); P.sd.push_back((uint8_t)sd); }
      }                                                         // else: blank line
      while (q < qe && *q != '\n') ++q;
      if (q < qe) ++q;
    }
  });
  side.assign((size_t)n, 0xFF);
  ascending = true;
  order0.clear(); order1.clear();
  order0.reserve((size_t)n / 2 + 1); order1.reserve((size_t)n / 2 + 1);
  int64_t rows = 0;
  long long prev = -1;
  for (int t = 0; t < T; ++t) {
    const Piece &P = pc[(size_t)t];
    if (P.err & 1) throw Error(EIGKL_E_FORMAT, std::string("malformed EIG row in ") + path);
    if (P.err & 2) throw Error(EIGKL_E_FORMAT, std::string("EIG row out of range in ") + path);
    for (size_t i = 0; i < P.node.size(); ++i) {
      const int32_t node = P.node[i];
      if (side[(size_t)node] != 0xFF) throw Error(EIGKL_E_FORMAT, std::string("EIG file lists a node twice: ") + path);
      if (node <= prev) ascending = false;
      prev = node;
      side[(size_t)node] = P.sd[i];
      (P.sd[i] ? order1 : order0).push_back(node);
      ++rows;
    }
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

// "<node>\t<side>" per line, nodes ascending (0-based, like the EIG file): the partition the reference computes but
// never writes (cKL.cpp:395-405)
void write_partition_file(const char *path, const uint8_t *side, int32_t n) {
  File f(fopen(path, "w"));
  if (!f) throw Error(EIGKL_E_IO, std::string("Error: Cannot open output file: ") + path);
  const int T = io_threads((size_t)n * 10);
  std::vector<std::string> part((size_t)T);
  run_parallel(T, [&](int t) {
    const int32_t lo = (int32_t)((int64_t)n * t / T), hi = (int32_t)((int64_t)n * (t + 1) / T);
    std::string &buf = part[(size_t)t];
    buf.reserve((size_t)(hi - lo) * 12 + 16);
    char tmp[32];
    for (int32_t i = lo; i < hi; ++i) {
      const int len = snprintf(tmp, sizeof(tmp), "%d\t%d\n", i, (int)(side[i] & 1u));
      buf.append(tmp, (size_t)len);
    }
  });
  for (int t = 0; t < T; ++t)
    if (fwrite(part[(size_t)t].data(), 1, part[(size_t)t].size(), f.get()) != part[(size_t)t].size())
      throw Error(EIGKL_E_IO, std::string("short write: ") + path);
}

}  // namespace eigkl
