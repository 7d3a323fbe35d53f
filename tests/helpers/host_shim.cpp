// host_shim.cpp -- TEST ONLY: exposes the product's host-side code (text I/O, dense eigen-solver,
// the libstdc++ order replay shared with the kernels) to the CPU test-suite, which has no GPU and
// therefore cannot create an eigkl_handle.  Built by tests/helpers/build_helpers.py with g++.
#include "internal.h"
#include "stl_order.h"
#include <unordered_map>
#include <unordered_set>
#include <vector>

using namespace eigkl;

extern "C" {

int shim_parse_hgr(const char *path, int32_t *n_nodes, int32_t *n_nets, int64_t *net_off /*cap*/, int64_t cap_off,
                   int32_t *pins, int64_t cap_pins, char *err, int errlen) {
  try {
    HostHgr h;
    parse_hgr(path, h);
    *n_nodes = h.n_nodes; *n_nets = h.n_nets;
    if ((int64_t)h.net_off.size() > cap_off || (int64_t)h.pins.size() > cap_pins) return -100;
    std::copy(h.net_off.begin(), h.net_off.end(), net_off);
    std::copy(h.pins.begin(), h.pins.end(), pins);
    return 0;
  } catch (const Error &e) { snprintf(err, errlen, "%s", e.what()); return e.code; }
}
int shim_write_eig(const char *path, double lambda2, double median, const double *vec, int32_t n) {
  try { write_eig_file(path, lambda2, median, vec, n); return 0; } catch (const Error &e) { return e.code; }
}
int shim_read_eig(const char *path, int32_t n, uint8_t *side, char *err, int errlen) {
  try {
    std::vector<uint8_t> s;
    std::vector<int32_t> o0, o1;
    bool asc = true;
    read_eig_file(path, n, s, o0, o1, asc);
    std::copy(s.begin(), s.end(), side);
    return 0;
  } catch (const Error &e) { snprintf(err, errlen, "%s", e.what()); return e.code; }
}
// as above, plus remain[0] ++ remain[1] in file order (order must hold n ints); returns 1 when the file is ascending, 0 when not
int shim_read_eig_orders(const char *path, int32_t n, uint8_t *side, int32_t *order, int32_t *n0) {
  try {
    std::vector<uint8_t> s;
    std::vector<int32_t> o0, o1;
    bool asc = true;
    read_eig_file(path, n, s, o0, o1, asc);
    std::copy(s.begin(), s.end(), side);
    std::copy(o0.begin(), o0.end(), order);
    std::copy(o1.begin(), o1.end(), order + o0.size());
    *n0 = (int32_t)o0.size();
    return asc ? 1 : 0;
  } catch (const Error &e) { return e.code; }
}
int shim_sym_eig(int n, double *a, double *evals) {
  try { sym_eig(n, a, evals); return 0; } catch (const Error &e) { return e.code; }
}
int shim_sym_top_eig(int n, const double *a, int k, double *theta, double *Y) {
  try { sym_top_eig(n, a, k, theta, Y); return 0; } catch (const Error &e) { return e.code; }
}
int shim_tridiag_top(int n, const double *d, const double *e, int k, double *theta, double *Y) {
  try { tridiag_top_eig(n, d, e, k, theta, Y); return 0; } catch (const Error &e2) { return e2.code; }
}
// product's order replay (stl_order.h), host instantiation
int shim_write_partition(const char *path, const uint8_t *side, int32_t n) {
  try { write_partition_file(path, side, n); return 0; } catch (const Error &e) { return e.code; }
}
void shim_stl_order(const uint32_t *keys, int32_t n, int32_t *order) {
  std::vector<int32_t> next((size_t)std::max(n, 1)), bkt(stl_final_buckets((uint32_t)n));
  int32_t head;
  stl_replay_inserts(n, [&](int32_t i) { return keys[i]; }, next.data(), bkt.data(), head);
  int32_t i = 0;
  for (int32_t p = head; p >= 0; p = next[p]) order[i++] = p;
}
// the REAL containers of this toolchain's libstdc++ (what the reference's behaviour is defined by)
void shim_real_unordered_map_order(const uint32_t *keys, int32_t n, uint32_t *out_keys) {
  std::unordered_map<uint32_t, float> m;
  for (int32_t i = 0; i < n; ++i) m[keys[i]] += 1.0f;               // cKL.cpp:128
  int32_t i = 0;
  for (const auto &kv : m) out_keys[i++] = kv.first;
}
void shim_real_unordered_set_order(const uint32_t *keys, int32_t n, uint32_t *out_keys) {
  std::vector<uint32_t> v(keys, keys + n);
  std::unordered_set<uint32_t> s(v.begin(), v.end());                // cKL.cpp:201
  int32_t i = 0;
  for (uint32_t k : s) out_keys[i++] = k;
}
uint32_t shim_final_buckets(uint32_t n) { return stl_final_buckets(n); }
uint32_t shim_real_bucket_count(uint32_t n) {
  std::unordered_map<uint32_t, float> m;
  for (uint32_t i = 0; i < n; ++i) m[i * 2654435761u] += 1.0f;
  return (uint32_t)m.bucket_count();
}

}  // extern "C"
