// assemble.cu -- hypergraph -> clique-model matrices on the GPU, as a sort + segmented reduce over
// net pins (north-star subsystem 1).
//
//   pins --expand--> P = sum k(k-1)/2 pairs (a<b, key = a:b, value = pair index in file order)
//        --stable radix sort by key--> runs of equal (a,b) in file order
//        --segmented reduce--> U unique edges with  wL = sum 2.0/k   (fp64, cEIG.cpp:110)
//                                                   wA = sum 1.0f/(k-1) (fp32, in file order, cKL.cpp:117,128)
//                                                   first = pair index of the first occurrence
//   L  (cEIG.cpp:86-133): rows ascending by column, diagonal = -(row sum of off-diagonals)
//   A  (cKL.cpp:84-149 in the traversal order of cKL.cpp:225-251): row v = forward neighbours (b>v)
//      in std::unordered_map iteration order (replayed from their first-occurrence order, stl_order.h)
//      followed by backward neighbours (a<v) ascending.
//
// All kernels are integer/byte streaming work bound by HBM/L2; no tensor-core shaped step exists.
#include "internal.h"
#include "device_utils.cuh"
#include "stl_order.h"
#include <algorithm>
#include <cstdlib>

namespace eigkl {

constexpr int TPB = 256;
static inline unsigned grid_for(int64_t n, int tpb = TPB) { return (unsigned)((n + tpb - 1) / tpb > 0 ? (n + tpb - 1) / tpb : 1); }

// ------------------------------------------------------------------------------------------------
// also validates the offsets: a non-monotonic net_off would give negative net sizes, and k(k-1)/2 is positive for
// negative k too, so the pair expansion would index pins[] out of bounds (err bit 4 -> EIGKL_E_FORMAT)
__global__ void net_pair_count_kernel(const int64_t *__restrict__ net_off, int32_t n_nets,
                                      int64_t *__restrict__ cnt, int *__restrict__ err) {
  int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_nets) {
    int64_t k = net_off[e + 1] - net_off[e];
    if (k < 0) { atomicOr(err, 4); k = 0; }
    cnt[e] = k * (k - 1) / 2;
  }
}

__global__ void validate_pins_kernel(const int32_t *__restrict__ pins, int64_t n_pins, int32_t n_nodes,
                                     int *__restrict__ err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pins) {
    int32_t p = pins[i];
    if (p < 0 || p >= n_nodes) atomicOr(err, 1);
  }
}

// one thread per pair: locate the net (binary search on pair_off), decode (j,l) from the running
// index inside the net (row-major upper triangle: (0,1),(0,2),...,(1,2),... as cKL.cpp:119-120)
__global__ void expand_pairs_kernel(const int64_t *__restrict__ net_off, const int32_t *__restrict__ pins,
                                    const int64_t *__restrict__ pair_off, int32_t n_nets, int64_t n_pairs,
                                    int node_bits, unsigned long long *__restrict__ keys,
                                    uint32_t *__restrict__ vals, int *__restrict__ err) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  int32_t lo = 0, hi = n_nets;            // last e with pair_off[e] <= p
  while (hi - lo > 1) {
    int32_t mid = (lo + hi) >> 1;
    if (pair_off[mid] <= p) lo = mid; else hi = mid;
  }
  const int32_t e = lo;
  const int64_t k = net_off[e + 1] - net_off[e];
  const int64_t q = p - pair_off[e];
  // j = largest j with j*(2k-j-1)/2 <= q
  const double t = (double)(2 * k - 1);
  int64_t j = (int64_t)floor((t - sqrt(t * t - 8.0 * (double)q)) * 0.5);
  if (j < 0) j = 0;
  if (j > k - 2) j = k - 2;
  while (j > 0 && j * (2 * k - j - 1) / 2 > q) --j;
  while (j < k - 2 && (j + 1) * (2 * k - j - 2) / 2 <= q) ++j;
  const int64_t l = q - j * (2 * k - j - 1) / 2 + j + 1;
  int32_t a = pins[net_off[e] + j], b = pins[net_off[e] + l];
  if (a == b) atomicOr(err, 2);           // duplicate pin in a net: no reference behaviour is pinned
  if (a > b) { int32_t s = a; a = b; b = s; }
  keys[p] = ((unsigned long long)(uint32_t)a << node_bits) | (unsigned long long)(uint32_t)b;
  vals[p] = (uint32_t)p;
}

__global__ void head_flag_kernel(const unsigned long long *__restrict__ keys, int64_t n, int32_t *__restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one thread per run head: walk the run (file order, because the sort is stable) and accumulate
__global__ void segment_reduce_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals,
                                      const int32_t *__restrict__ uid, int64_t n, int node_bits,
                                      const int64_t *__restrict__ net_off, const int64_t *__restrict__ pair_off,
                                      int32_t n_nets, int32_t *__restrict__ ua, int32_t *__restrict__ ub,
                                      float *__restrict__ uwA, double *__restrict__ uwL, uint32_t *__restrict__ ufirst) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = keys[i];
  if (i > 0 && keys[i - 1] == key) return;
  float sA = 0.0f;      // map value-initialises to 0.0f, then += w      cKL.cpp:128
  double sL = 0.0;
  for (int64_t r = i; r < n && keys[r] == key; ++r) {
    const int64_t p = vals[r];
    int32_t lo = 0, hi = n_nets;
    while (hi - lo > 1) {
      int32_t mid = (lo + hi) >> 1;
      if (pair_off[mid] <= p) lo = mid; else hi = mid;
    }
    const int64_t k = net_off[lo + 1] - net_off[lo];
    sA = __fadd_rn(sA, __fdiv_rn(1.0f, (float)(k - 1)));          // cKL.cpp:117
    sL = __dadd_rn(sL, __ddiv_rn(2.0, (double)k));                // cEIG.cpp:110
  }
  const int32_t u = uid[i];
  ua[u] = (int32_t)(key >> node_bits);
  ub[u] = (int32_t)(key & ((1ull << node_bits) - 1ull));
  uwA[u] = sA;
  uwL[u] = sL;
  ufirst[u] = vals[i];
}

// start[v] = first index i in [0,n) with sorted[i] >= v   (v in [0, n_nodes])
__global__ void lower_bound_kernel(const int32_t *__restrict__ sorted, int64_t n, int32_t n_nodes,
                                   int32_t *__restrict__ start) {
  int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v > n_nodes) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted[mid] < v) lo = mid + 1; else hi = mid;
  }
  start[v] = (int32_t)lo;
}

__global__ void make_bkey_kernel(const int32_t *__restrict__ ub, int64_t U, unsigned long long *__restrict__ keys,
                                 uint32_t *__restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < U) { keys[i] = (unsigned long long)(uint32_t)ub[i]; vals[i] = (uint32_t)i; }
}
__global__ void keys_to_i32_kernel(const unsigned long long *__restrict__ keys, int64_t n, int32_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)keys[i];
}

// ------------------------------------------------------------------------------------------------
void upload_pins(eigkl_handle *h, int32_t n_nodes, int32_t n_nets, const int64_t *net_off, const int32_t *pins) {
  EIGKL_REQUIRE(n_nodes > 0 && n_nets >= 0 && net_off && (pins || net_off[n_nets] == 0), EIGKL_E_ARG, "eigkl_set_pins: bad arguments");
  EIGKL_REQUIRE(net_off[0] == 0, EIGKL_E_ARG, "eigkl_set_pins: net_off[0] must be 0");
  const int64_t n_pins = net_off[n_nets];
  auto &g = h->hg;
  g.loaded = false;
  h->ue.valid = false;
  h->ueL.valid = false;
  h->order.valid = false;
  h->L.valid = false;
  h->A.valid = false;
  h->kl.have_partition = false;
  h->eig.have_vector = h->eig.have_median = false;
  g.n_nodes = n_nodes; g.n_nets = n_nets; g.n_pins = n_pins;
  g.net_off.alloc((size_t)n_nets + 1);
  g.pins.alloc((size_t)n_pins);
  g.pair_off.alloc((size_t)n_nets + 1);
  EIGKL_CUDA(cudaMemcpyAsync(g.net_off.p, net_off, ((size_t)n_nets + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
  if (n_pins) EIGKL_CUDA(cudaMemcpyAsync(g.pins.p, pins, (size_t)n_pins * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  auto &err = h->scr.err; err.alloc(1);
  EIGKL_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), h->stream));
  if (n_pins) { validate_pins_kernel<<<grid_for(n_pins), TPB, 0, h->stream>>>(g.pins.p, n_pins, n_nodes, err.p); h->launches++; }
  if (n_nets) { net_pair_count_kernel<<<grid_for(n_nets), TPB, 0, h->stream>>>(g.net_off.p, n_nets, g.pair_off.p, err.p); h->launches++; }
  exclusive_scan_i64(h, g.pair_off.p, g.pair_off.p, n_nets);
  int herr = 0;
  int64_t P = 0;
  EIGKL_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  EIGKL_CUDA(cudaMemcpyAsync(&P, g.pair_off.p + n_nets, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  EIGKL_CUDA(cudaStreamSynchronize(h->stream));
  EIGKL_REQUIRE((herr & 4) == 0, EIGKL_E_FORMAT, "net_off is not non-decreasing");
  EIGKL_REQUIRE(herr == 0, EIGKL_E_FORMAT, "pin id out of range [1, nodes]");
  EIGKL_REQUIRE(P < (int64_t)2147483647, EIGKL_E_ARG, "more than 2^31-1 clique pairs");
  g.n_pairs = P;
  g.loaded = true;
  h->stats.n_nodes = n_nodes; h->stats.n_nets = n_nets; h->stats.n_pins = n_pins; h->stats.n_pairs = P;
}

void build_unique_edges(eigkl_handle *h, UniqueEdges &ue, const int32_t *pins) {
  auto &g = h->hg;
  EIGKL_REQUIRE(g.loaded, EIGKL_E_ARG, "no hypergraph loaded");
  if (ue.valid) return;
  const int64_t P = g.n_pairs;
  const int32_t n = g.n_nodes;
  const int nb = bits_for((uint64_t)(n > 1 ? n - 1 : 1));
  auto &e = h->eig;
  for (int i = 0; i < 2; ++i) { e.sortkey[i].ensure((size_t)std::max<int64_t>(P, n) + 1); e.sortval[i].ensure((size_t)std::max<int64_t>(P, n) + 1); }
  unsigned long long *keys[2] = {e.sortkey[0].p, e.sortkey[1].p};
  uint32_t *vals[2] = {e.sortval[0].p, e.sortval[1].p};
  auto &err = h->scr.err; err.alloc(1);
  EIGKL_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), h->stream));
  int64_t U = 0;
  if (P > 0) {
    expand_pairs_kernel<<<grid_for(P), TPB, 0, h->stream>>>(g.net_off.p, pins, g.pair_off.p, g.n_nets, P, nb, keys[0], vals[0], err.p);
    h->launches++;
    const int cur = radix_sort_kv(h, keys, vals, P, 2 * nb);
    auto &flag = h->scr.i32a; flag.alloc((size_t)P + 1);
    head_flag_kernel<<<grid_for(P), TPB, 0, h->stream>>>(keys[cur], P, flag.p);
    h->launches++;
    exclusive_scan_i32(h, flag.p, flag.p, P);
    int32_t U32 = 0; int herr = 0;
    EIGKL_CUDA(cudaMemcpyAsync(&U32, flag.p + P, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    EIGKL_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    EIGKL_CUDA(cudaStreamSynchronize(h->stream));
    EIGKL_REQUIRE(herr == 0, EIGKL_E_FORMAT, "a net lists the same pin twice");
    U = U32;
    ue.a.alloc((size_t)U); ue.b.alloc((size_t)U); ue.wA.alloc((size_t)U); ue.wL.alloc((size_t)U); ue.first.alloc((size_t)U);
    segment_reduce_kernel<<<grid_for(P), TPB, 0, h->stream>>>(keys[cur], vals[cur], flag.p, P, nb, g.net_off.p, g.pair_off.p,
                                                             g.n_nets, ue.a.p, ue.b.p, ue.wA.p, ue.wL.p, ue.first.p);
    h->launches++;
  }
  ue.U = U;
  ue.fstart.alloc((size_t)n + 1);
  ue.bstart.alloc((size_t)n + 1);
  ue.perm_b.alloc((size_t)U);
  lower_bound_kernel<<<grid_for(n + 1), TPB, 0, h->stream>>>(ue.a.p, U, n, ue.fstart.p);
  h->launches++;
  if (U > 0) {
    make_bkey_kernel<<<grid_for(U), TPB, 0, h->stream>>>(ue.b.p, U, keys[0], vals[0]);
    h->launches++;
    const int cur = radix_sort_kv(h, keys, vals, U, nb);
    EIGKL_CUDA(cudaMemcpyAsync(ue.perm_b.p, vals[cur], (size_t)U * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
    auto &bs = h->scr.i32b; bs.alloc((size_t)U);
    keys_to_i32_kernel<<<grid_for(U), TPB, 0, h->stream>>>(keys[cur], U, bs.p);
    lower_bound_kernel<<<grid_for(n + 1), TPB, 0, h->stream>>>(bs.p, U, n, ue.bstart.p);
    h->launches += 2;
  } else {
    EIGKL_CUDA(cudaMemsetAsync(ue.bstart.p, 0, ((size_t)n + 1) * sizeof(int32_t), h->stream));
  }
  EIGKL_CUDA(cudaGetLastError());
  ue.valid = true;
}

// ------------------------------------------------------------------------------------------------
// Laplacian
// ------------------------------------------------------------------------------------------------
__global__ void lap_degree_kernel(const int32_t *__restrict__ fstart, const int32_t *__restrict__ bstart, int32_t n,
                                  int32_t *__restrict__ deg) {
  int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) deg[v] = (fstart[v + 1] - fstart[v]) + (bstart[v + 1] - bstart[v]) + 1;
}
__global__ void lap_fill_fwd_kernel(const int32_t *__restrict__ ua, const int32_t *__restrict__ ub, const double *__restrict__ uwL,
                                    int64_t U, const int32_t *__restrict__ fstart, const int32_t *__restrict__ bstart,
                                    const int32_t *__restrict__ rowptr, int32_t *__restrict__ col, double *__restrict__ val) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= U) return;
  const int32_t a = ua[i];
  const int64_t pos = (int64_t)rowptr[a] + (bstart[a + 1] - bstart[a]) + 1 + (i - fstart[a]);
  col[pos] = ub[i];
  val[pos] = -uwL[i];                                            // cEIG.cpp:114
}
__global__ void lap_fill_bwd_kernel(const int32_t *__restrict__ ua, const int32_t *__restrict__ ub, const double *__restrict__ uwL,
                                    const uint32_t *__restrict__ perm_b, int64_t U, const int32_t *__restrict__ bstart,
                                    const int32_t *__restrict__ rowptr, int32_t *__restrict__ col, double *__restrict__ val) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= U) return;
  const uint32_t e = perm_b[j];
  const int32_t b = ub[e];
  const int64_t pos = (int64_t)rowptr[b] + (j - bstart[b]);
  col[pos] = ua[e];
  val[pos] = -uwL[e];                                            // cEIG.cpp:115
}
// diagonal = -(sum of the row's off-diagonals, ascending column order)      cEIG.cpp:127-130
// also reduces min / max of the diagonal (spectrum bounds for the Chebyshev filter of the Fiedler solve)
__global__ void lap_diag_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ bstart, int32_t n,
                                int32_t *__restrict__ col, double *__restrict__ val, unsigned long long *__restrict__ minmax) {
  int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long kmin = 0ull, kmax = 0ull;      // max over ~orderable(d) == min over d
  if (v < n) {
    const int32_t lo = rowptr[v], hi = rowptr[v + 1];
    const int32_t dpos = lo + (bstart[v + 1] - bstart[v]);
    double s = 0.0;
    for (int32_t i = lo; i < hi; ++i)
      if (i != dpos) s += val[i];
    col[dpos] = v;
    val[dpos] = -s;
    const unsigned long long k = double_orderable(-s);
    kmin = ~k; kmax = k;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long a = __shfl_xor_sync(FULL_MASK, kmin, o), b = __shfl_xor_sync(FULL_MASK, kmax, o);
    kmin = a > kmin ? a : kmin; kmax = b > kmax ? b : kmax;
  }
  if ((threadIdx.x & 31) == 0) { atomicMax(minmax, kmin); atomicMax(minmax + 1, kmax); }
}
__global__ void diag_decode_kernel(const unsigned long long *__restrict__ minmax, double *__restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = double_from_orderable(~minmax[0]); out[1] = double_from_orderable(minmax[1]); }
}
// blk_row[b] = first row r in [row_lo, row_hi] with cost(r) >= b*chunk, where
// cost(r) = rowptr[r] - rowptr[row_lo] + row_weight * (r - row_lo)   (non-zeros, plus a per-row charge)
// (b in [0, n_blocks]; blk_row[n_blocks] = row_hi)
__global__ void row_blocks_kernel(const int32_t *__restrict__ rowptr, int32_t row_lo, int32_t row_hi, int64_t chunk,
                                  int32_t n_blocks, int32_t *__restrict__ blk_row, int64_t row_weight = 0) {
  int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_blocks) return;
  if (b == n_blocks) { blk_row[b] = row_hi; return; }
  const int64_t base = (int64_t)rowptr[row_lo];
  const int64_t x = (int64_t)b * chunk;
  int32_t lo = row_lo, hi = row_hi;
  while (lo < hi) {
    int32_t mid = (lo + hi) >> 1;
    if ((int64_t)rowptr[mid] - base + row_weight * (int64_t)(mid - row_lo) < x) lo = mid + 1; else hi = mid;
  }
  blk_row[b] = lo;
}

__global__ void blk_info_kernel(const int32_t *__restrict__ blk_row, const int32_t *__restrict__ rowptr, int32_t n_blocks,
                                int4 *__restrict__ info) {
  int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n_blocks) {
    const int32_t r0 = blk_row[b], r1 = blk_row[b + 1];
    info[b] = make_int4(r0, r1, rowptr[r0], rowptr[r1]);
  }
}

// res_check[0] = longest span, [1] = most rows over the resident row blocks
__global__ void blk_check_kernel(const int4 *__restrict__ info, int32_t n_blocks, int32_t *__restrict__ out) {
  int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n_blocks) {
    const int4 i = info[b];
    atomicMax(out, i.w - i.z);
    atomicMax(out + 1, i.y - i.x);
  }
}

// Row-block size in non-zeros.  The circuits here are small next to a B200 (ibm01: 0.23 M non-zeros vs
// 0.3 M resident threads), so the kernels are latency bound: the chunk is sized to spread the matrix
// over the CTAs resident at once (6 per SM for the SpMV kernel, 8 for the D-value kernel) in ONE wave,
// between 256 and 2048 non-zeros (the staging capacities in spmv.cu / kl.cu are 4096).
static int64_t pick_chunk(const eigkl_handle *h, int64_t nnz, int ctas_per_sm) {
  if (const char *ev = getenv("EIGKL_CHUNK")) return std::max<int64_t>(64, atoll(ev));   // tuning aid
  const int64_t target_ctas = (int64_t)h->sm_count * ctas_per_sm;
  int64_t c = ceil_div(std::max<int64_t>(nnz, 1), target_ctas);
  c = ceil_div(c, 256) * 256;
  return std::min<int64_t>(2048, std::max<int64_t>(256, c));
}

// ---- node order of the EIG stage ---------------------------------------------------------------------
// The x gathers of the SpMV are its traffic: in the file's numbering nearly every gather of ibm10 is a new
// 32-byte sector (0.88 distinct sectors per non-zero within 128 consecutive rows).  Numbering the nodes by
// the first net that mentions them puts net-mates next to each other: 0.17 distinct sectors per non-zero
// (reverse Cuthill-McKee gives 0.18 and costs a BFS), i.e. the gathers hit L1 instead of L2.
__global__ void first_net_kernel(const int64_t *__restrict__ net_off, const int32_t *__restrict__ pins, int32_t n_nets,
                                 int64_t n_pins, uint32_t *__restrict__ first) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pins) return;
  int32_t lo = 0, hi = n_nets;            // last e with net_off[e] <= p
  while (hi - lo > 1) {
    int32_t mid = (lo + hi) >> 1;
    if (net_off[mid] <= p) lo = mid; else hi = mid;
  }
  atomicMin(&first[pins[p]], (uint32_t)lo);
}
__global__ void order_keys_kernel(const uint32_t *__restrict__ first, int32_t n, uint32_t n_nets, unsigned long long *__restrict__ keys,
                                  uint32_t *__restrict__ vals) {
  int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) { keys[v] = (unsigned long long)min(first[v], n_nets); vals[v] = (uint32_t)v; }   // isolated nodes last
}
__global__ void order_finish_kernel(const uint32_t *__restrict__ sorted, int32_t n, int32_t *__restrict__ perm, int32_t *__restrict__ inv) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int32_t v = (int32_t)sorted[i]; perm[i] = v; inv[v] = i; }
}
__global__ void identity_order_kernel(int32_t n, int32_t *__restrict__ perm, int32_t *__restrict__ inv) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { perm[i] = i; inv[i] = i; }
}
__global__ void relabel_pins_kernel(const int32_t *__restrict__ pins, const int32_t *__restrict__ inv, int64_t n_pins, int32_t *__restrict__ out) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_pins) out[p] = inv[pins[p]];
}

static void build_node_order(eigkl_handle *h) {
  auto &g = h->hg;
  auto &o = h->order;
  if (o.valid) return;
  const int32_t n = g.n_nodes;
  o.active = !(h->opts.flags & EIGKL_F_NATURAL_ORDER) && g.n_pins > 0;
  o.perm.alloc((size_t)n); o.inv.alloc((size_t)n);
  if (!o.active) {
    identity_order_kernel<<<grid_for(n), TPB, 0, h->stream>>>(n, o.perm.p, o.inv.p);
    h->launches++;
    o.valid = true;
    return;
  }
  o.first.alloc((size_t)n); o.pins.alloc((size_t)g.n_pins);
  auto &e = h->eig;
  for (int i = 0; i < 2; ++i) { e.sortkey[i].ensure((size_t)std::max<int64_t>(g.n_pairs, n) + 1); e.sortval[i].ensure((size_t)std::max<int64_t>(g.n_pairs, n) + 1); }
  unsigned long long *keys[2] = {e.sortkey[0].p, e.sortkey[1].p};
  uint32_t *vals[2] = {e.sortval[0].p, e.sortval[1].p};
  EIGKL_CUDA(cudaMemsetAsync(o.first.p, 0xFF, (size_t)n * sizeof(uint32_t), h->stream));
  first_net_kernel<<<grid_for(g.n_pins), TPB, 0, h->stream>>>(g.net_off.p, g.pins.p, g.n_nets, g.n_pins, o.first.p);
  order_keys_kernel<<<grid_for(n), TPB, 0, h->stream>>>(o.first.p, n, (uint32_t)g.n_nets, keys[0], vals[0]);
  const int cur = radix_sort_kv(h, keys, vals, n, bits_for((uint64_t)std::max(g.n_nets, 1)));
  order_finish_kernel<<<grid_for(n), TPB, 0, h->stream>>>(vals[cur], n, o.perm.p, o.inv.p);
  relabel_pins_kernel<<<grid_for(g.n_pins), TPB, 0, h->stream>>>(g.pins.p, o.inv.p, g.n_pins, o.pins.p);
  h->launches += 4;
  EIGKL_CUDA(cudaGetLastError());
  o.valid = true;
}

// row blocks of the resident filter: cost = non-zeros + rows, cut every `chunk`
void resident_row_blocks(eigkl_handle *h, int64_t chunk) {
  auto &L = h->L;
  row_blocks_kernel<<<grid_for(L.res_blocks + 1), TPB, 0, h->stream>>>(L.rowptr.p, 0, L.n, chunk, L.res_blocks, L.res_row.p, 1);
  blk_info_kernel<<<grid_for(L.res_blocks), TPB, 0, h->stream>>>(L.res_row.p, L.rowptr.p, L.res_blocks, reinterpret_cast<int4 *>(L.res_info.p));
  blk_check_kernel<<<grid_for(L.res_blocks), TPB, 0, h->stream>>>(reinterpret_cast<const int4 *>(L.res_info.p), L.res_blocks, L.res_check.p);
  EIGKL_CUDA(cudaGetLastError());
  h->launches += 3;
}

void assemble_laplacian(eigkl_handle *h) {
  build_node_order(h);
  auto &ue = h->order.active ? h->ueL : h->ue;
  build_unique_edges(h, ue, h->order.active ? h->order.pins.p : h->hg.pins.p);
  auto &L = h->L;
  const int32_t n = h->hg.n_nodes;
  const int64_t U = ue.U;
  L.valid = false;
  L.n = n;
  L.nnz = 2 * U + n;
  EIGKL_REQUIRE(L.nnz < (int64_t)2147483647, EIGKL_E_ARG, "Laplacian has more than 2^31-1 non-zeros");
  L.rowptr.alloc((size_t)n + 1);
  L.col.alloc((size_t)L.nnz);
  L.val.alloc((size_t)L.nnz);
  lap_degree_kernel<<<grid_for(n), TPB, 0, h->stream>>>(ue.fstart.p, ue.bstart.p, n, L.rowptr.p);
  h->launches++;
  exclusive_scan_i32(h, L.rowptr.p, L.rowptr.p, n);
  if (U > 0) {
    lap_fill_fwd_kernel<<<grid_for(U), TPB, 0, h->stream>>>(ue.a.p, ue.b.p, ue.wL.p, U, ue.fstart.p, ue.bstart.p, L.rowptr.p, L.col.p, L.val.p);
    lap_fill_bwd_kernel<<<grid_for(U), TPB, 0, h->stream>>>(ue.a.p, ue.b.p, ue.wL.p, ue.perm_b.p, U, ue.bstart.p, L.rowptr.p, L.col.p, L.val.p);
    h->launches += 2;
  }
  L.diag_minmax.alloc(4);
  EIGKL_CUDA(cudaMemsetAsync(L.diag_minmax.p, 0, 4 * sizeof(unsigned long long), h->stream));
  lap_diag_kernel<<<grid_for(n), TPB, 0, h->stream>>>(L.rowptr.p, ue.bstart.p, n, L.col.p, L.val.p, L.diag_minmax.p);
  diag_decode_kernel<<<1, 32, 0, h->stream>>>(L.diag_minmax.p, reinterpret_cast<double *>(L.diag_minmax.p + 2));
  double dmm[2] = {0, 0};
  EIGKL_CUDA(cudaMemcpyAsync(dmm, L.diag_minmax.p + 2, sizeof(dmm), cudaMemcpyDeviceToHost, h->stream));
  cheb_resident_plan(h);     // resident polynomial filter (spmv.cu): plan enqueued, checked after the sync
  EIGKL_CUDA(cudaStreamSynchronize(h->stream));
  L.diag_min = dmm[0]; L.diag_max = dmm[1];
  cheb_resident_plan_finish(h);
  // The matrix itself is assembled in full on every rank (~1 ms, and it keeps the assembly free of collectives).
  // With several ranks: a matrix that fits one chip is SOLVED on every rank as well; a larger one is cut into
  // nnz-balanced row ranges, and SpMV and every Lanczos vector are row-partitioned (dist.cu).
  const int64_t nnz_local = dist_decide(h);
  // which SpMV kernel runs this matrix: the flat kernel (shared-memory staged row blocks, one round of loads
  // per block; blocks that hold a row too long for the staging buffer fall back to warp-per-row inside it).
  // Measured warm, flat vs sub-warp: ibm01 5.1 vs 7.3 us, ibm10 10.3 vs 14.4, industry2 10.4 vs 12.4,
  // synthetic x1 10.3 vs 15.3.  EIGKL_SPMV_MODE = 1 / 2 select the older staged / sub-warp kernels.
  L.flat = (h->spmv_mode == 3) || (h->spmv_mode == 0);
  int64_t chunk = pick_chunk(h, nnz_local, L.flat ? 4 : 6);
  if (L.flat) chunk = std::min<int64_t>(chunk, 1792);     // 2048-product staging buffer minus slack for the last row
  L.n_blocks = (int32_t)std::max<int64_t>(1, ceil_div(nnz_local, chunk));
  L.blk_row.alloc((size_t)L.n_blocks + 1);
  L.blk_info.alloc((size_t)4 * L.n_blocks + 4);
  row_blocks_kernel<<<grid_for(L.n_blocks + 1), TPB, 0, h->stream>>>(L.rowptr.p, L.row_lo, L.row_hi, chunk, L.n_blocks, L.blk_row.p);
  blk_info_kernel<<<grid_for(L.n_blocks), TPB, 0, h->stream>>>(L.blk_row.p, L.rowptr.p, L.n_blocks, reinterpret_cast<int4 *>(L.blk_info.p));
  h->launches++;
  h->launches += 3;
  EIGKL_CUDA(cudaGetLastError());
  h->stats.dist_ranks = 1; h->stats.dist_rows = n; h->stats.dist_halo = 0; h->stats.dist_exports = 0;
  dist_plan(h);              // halo / export lists of this rank's rows (row-partitioned mode only)
  L.valid = true;
  h->stats.nnz_laplacian = L.nnz;
  // algorithmic bytes of one SpMV: nnz*(8 val + 4 col) + n*(4 rowptr + 8 x + 8 y)   (SURVEY.md 8d); row-partitioned:
  // this rank's share (its rows' entries, its rows of x / y) -- the halo values it receives are counted as x reads
  h->stats.bytes_spmv = h->dist.valid ? (double)h->dist.nnz_l * 12.0 + (double)h->dist.nl * 20.0 + (double)h->stats.dist_halo * 8.0
                                      : (double)L.nnz * 12.0 + (double)n * 20.0;
}

// ------------------------------------------------------------------------------------------------
// KL graph
// ------------------------------------------------------------------------------------------------
__global__ void kl_degree_kernel(const int32_t *__restrict__ fstart, const int32_t *__restrict__ bstart, int32_t n,
                                 int32_t *__restrict__ deg, int32_t *__restrict__ scratch_need) {
  int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) {
    const int32_t fd = fstart[v + 1] - fstart[v];
    deg[v] = fd + (bstart[v + 1] - bstart[v]);
    scratch_need[v] = fd > 1 ? fd + (int32_t)stl_final_buckets((uint32_t)fd) : 0;
  }
}
__global__ void kl_first_key_kernel(const int32_t *__restrict__ ua, const uint32_t *__restrict__ ufirst, int64_t U,
                                    int pair_bits, unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < U) {
    keys[i] = ((unsigned long long)(uint32_t)ua[i] << pair_bits) | (unsigned long long)ufirst[i];
    vals[i] = (uint32_t)i;
  }
}
struct FwdKey {
  const uint32_t *by_first; const int32_t *ub; int32_t f0;
  EIGKL_HD uint32_t operator()(int32_t i) const { return (uint32_t)ub[by_first[f0 + i]]; }
};
// one thread per row: replay the unordered_map inserts of the row's forward keys (first-occurrence
// order) and emit them in iteration order                                   cKL.cpp:128, 230-236
__global__ void kl_fill_fwd_kernel(const uint32_t *__restrict__ by_first /* edge ids sorted by (a, first) */,
                                   const int32_t *__restrict__ ub, const float *__restrict__ uwA,
                                   const int32_t *__restrict__ fstart, const int32_t *__restrict__ rowptr,
                                   const int32_t *__restrict__ scratch_off, int32_t *__restrict__ scratch, int32_t n,
                                   int32_t *__restrict__ fwd_end, int32_t *__restrict__ col, float *__restrict__ w) {
  int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int32_t f0 = fstart[v], d = fstart[v + 1] - f0, r0 = rowptr[v];
  fwd_end[v] = r0 + d;
  if (d == 0) return;
  if (d == 1) { const uint32_t e = by_first[f0]; col[r0] = ub[e]; w[r0] = uwA[e]; return; }
  int32_t *next = scratch + scratch_off[v];
  int32_t *bkt = next + d;
  int32_t head;
  stl_replay_inserts(d, FwdKey{by_first, ub, f0}, next, bkt, head);
  int32_t pos = r0;
  for (int32_t p = head; p >= 0; p = next[p]) {
    const uint32_t e = by_first[f0 + p];
    col[pos] = ub[e];
    w[pos] = uwA[e];
    ++pos;
  }
}
__global__ void kl_fill_bwd_kernel(const int32_t *__restrict__ ua, const int32_t *__restrict__ ub, const float *__restrict__ uwA,
                                   const uint32_t *__restrict__ perm_b, int64_t U, const int32_t *__restrict__ bstart,
                                   const int32_t *__restrict__ fwd_end, int32_t *__restrict__ col, float *__restrict__ w) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= U) return;
  const uint32_t e = perm_b[j];
  const int32_t b = ub[e];
  const int64_t pos = (int64_t)fwd_end[b] + (j - bstart[b]);
  col[pos] = ua[e];
  w[pos] = uwA[e];
}

void assemble_kl_graph(eigkl_handle *h) {
  build_unique_edges(h, h->ue, h->hg.pins.p);
  auto &ue = h->ue;
  auto &A = h->A;
  const int32_t n = h->hg.n_nodes;
  const int64_t U = ue.U;
  A.valid = false;
  A.n = n;
  A.nnz = 2 * U;
  EIGKL_REQUIRE(A.nnz < (int64_t)2147483647, EIGKL_E_ARG, "KL graph has more than 2^31-1 entries");
  A.rowptr.alloc((size_t)n + 1);
  A.fwd_end.alloc((size_t)n);
  A.col.alloc((size_t)A.nnz);
  A.w.alloc((size_t)A.nnz);
  auto &soff = h->scr.i32a; soff.alloc((size_t)n + 1);
  kl_degree_kernel<<<grid_for(n), TPB, 0, h->stream>>>(ue.fstart.p, ue.bstart.p, n, A.rowptr.p, soff.p);
  h->launches++;
  exclusive_scan_i32(h, A.rowptr.p, A.rowptr.p, n);
  exclusive_scan_i32(h, soff.p, soff.p, n);
  int32_t scratch_total = 0;
  EIGKL_CUDA(cudaMemcpyAsync(&scratch_total, soff.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  EIGKL_CUDA(cudaStreamSynchronize(h->stream));
  auto &scratch = h->scr.i32b; scratch.alloc((size_t)scratch_total + 1);
  if (U > 0) {
    auto &e = h->eig;
    unsigned long long *keys[2] = {e.sortkey[0].p, e.sortkey[1].p};
    uint32_t *vals[2] = {e.sortval[0].p, e.sortval[1].p};
    const int nb = bits_for((uint64_t)(n > 1 ? n - 1 : 1));
    const int pb = bits_for((uint64_t)(h->hg.n_pairs > 1 ? h->hg.n_pairs - 1 : 1));
    kl_first_key_kernel<<<grid_for(U), TPB, 0, h->stream>>>(ue.a.p, ue.first.p, U, pb, keys[0], vals[0]);
    h->launches++;
    const int cur = radix_sort_kv(h, keys, vals, U, nb + pb);
    kl_fill_fwd_kernel<<<grid_for(n, 128), 128, 0, h->stream>>>(vals[cur], ue.b.p, ue.wA.p, ue.fstart.p, A.rowptr.p, soff.p,
                                                               scratch.p, n, A.fwd_end.p, A.col.p, A.w.p);
    kl_fill_bwd_kernel<<<grid_for(U), TPB, 0, h->stream>>>(ue.a.p, ue.b.p, ue.wA.p, ue.perm_b.p, U, ue.bstart.p, A.fwd_end.p, A.col.p, A.w.p);
    h->launches += 2;
  } else {
    EIGKL_CUDA(cudaMemsetAsync(A.fwd_end.p, 0, (size_t)n * sizeof(int32_t), h->stream));
  }
  // the D-value kernel stages a block's 2048 signed weights: leave room for the block's last row
  // (blocks are cut by non-zeros PLUS rows: a node without neighbours has an empty row, and a run of those must not
  // overflow the kernel's 2048 staged row offsets)
  const int64_t chunk = std::min<int64_t>(pick_chunk(h, A.nnz + n, 8), 1792);
  A.n_blocks = (int32_t)std::max<int64_t>(1, ceil_div(A.nnz + n, chunk));
  A.blk_row.alloc((size_t)A.n_blocks + 1);
  row_blocks_kernel<<<grid_for(A.n_blocks + 1), TPB, 0, h->stream>>>(A.rowptr.p, 0, n, chunk, A.n_blocks, A.blk_row.p, 1);
  h->launches++;
  EIGKL_CUDA(cudaGetLastError());
  A.valid = true;
  A.nb_valid = false;
  A.info_valid = false;
  h->stats.nnz_kl = A.nnz;
  // algorithmic bytes of one full D-value pass: nnz*(4 w + 4 col) + n*(4 rowptr + 1 side + 4 out)  
  h->stats.bytes_dvalues = (double)A.nnz * 8.0 + (double)n * 9.0;
}

}  // namespace eigkl
