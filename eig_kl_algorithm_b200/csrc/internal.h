// internal.h -- shared declarations of libeigkl.so (host side).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <algorithm>
#include <string>
#include <vector>
#include "eigkl.h"

namespace eigkl {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define EIGKL_CUDA(call)                                                                         \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      throw ::eigkl::Error(EIGKL_E_CUDA, std::string("CUDA error in ") + __FILE__ + ":" +        \
                                             std::to_string(__LINE__) + ": " + cudaGetErrorString(e__)); \
  } while (0)

#define EIGKL_REQUIRE(cond, code, msg)                                                           \
  do {                                                                                           \
    if (!(cond)) throw ::eigkl::Error((code), (msg));                                            \
  } while (0)

// device buffer owned by the handle (plain cudaMalloc; sizes here are a few GB at most of 180 GB)
template <typename T>
struct DBuf {
  T *p = nullptr;
  size_t n = 0;
  DBuf() = default;
  DBuf(const DBuf &) = delete;
  DBuf &operator=(const DBuf &) = delete;
  DBuf(DBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DBuf &operator=(DBuf &&o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; n = 0;
  }
  // grow-only: device allocations are never returned inside the hot path (cudaFree synchronises the
  // device and cudaMalloc costs 0.1-100 ms, more than a whole assembly stage)
  void alloc(size_t count) {
    if (count == 0) count = 1;
    if (count <= n) return;
    release();
    EIGKL_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
    n = count;
  }
  void ensure(size_t count) { alloc(count); }
  size_t bytes() const { return n * sizeof(T); }
};

// pinned host buffer
template <typename T>
struct HBuf {
  T *p = nullptr;
  size_t n = 0;
  HBuf() = default;
  HBuf(const HBuf &) = delete;
  HBuf &operator=(const HBuf &) = delete;
  ~HBuf() { if (p) cudaFreeHost(p); }
  void ensure(size_t count) {
    if (count <= n) return;
    if (p) cudaFreeHost(p);
    p = nullptr;
    EIGKL_CUDA(cudaMallocHost((void **)&p, count * sizeof(T)));
    n = count;
  }
};

struct StageTimer {
  cudaEvent_t a = nullptr, b = nullptr;
  void init() {
    EIGKL_CUDA(cudaEventCreate(&a));
    EIGKL_CUDA(cudaEventCreate(&b));
  }
  void destroy() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    a = b = nullptr;
  }
  void start(cudaStream_t s) { EIGKL_CUDA(cudaEventRecord(a, s)); }
  void stop(cudaStream_t s) { EIGKL_CUDA(cudaEventRecord(b, s)); }
  double ms() {
    float t = 0.f;
    EIGKL_CUDA(cudaEventSynchronize(b));
    EIGKL_CUDA(cudaEventElapsedTime(&t, a, b));
    return (double)t;
  }
};

// per kernel-class profiler (EIGKL_F_PROFILE): a pool of event pairs, resolved lazily
struct KernelProfiler {
  bool on = false;
  std::vector<cudaEvent_t> ev;   // pairs
  std::vector<int> cls;
  std::vector<int> weight;       // launches covered by the pair (a group of back-to-back launches of one class)
  int suppress = 0;              // >0: inner begin/end calls are ignored (the caller brackets a group)
  size_t used = 0;
  double ms[8] = {0};
  int64_t cnt[8] = {0};
  void begin(int c, cudaStream_t s, int w = 1);
  void end(cudaStream_t s);
  void resolve();
  void reset();
  ~KernelProfiler();
};
enum { KC_SPMV = 0, KC_MULTIDOT = 1, KC_UPDATE = 2, KC_RESTART = 3, KC_DVALUES = 4, KC_COMM = 5, KC_PUSH = 6 };

// ---------------------------------------------------------------------------------------------------
// device-resident problem state
// ---------------------------------------------------------------------------------------------------
struct Hypergraph {            // input pins (device) + sizes
  int32_t n_nodes = 0, n_nets = 0;
  int64_t n_pins = 0, n_pairs = 0;
  DBuf<int64_t> net_off;       // n_nets+1
  DBuf<int32_t> pins;          // n_pins
  DBuf<int64_t> pair_off;      // n_nets+1 : exclusive scan of k(k-1)/2
  bool loaded = false;
};

struct UniqueEdges {           // result of sort + segmented reduce: unique pairs a<b, sorted by (a,b)
  int64_t U = 0;
  DBuf<int32_t> a, b;
  DBuf<float> wA;              // sum of 1.0f/(k-1) in file order (fp32)           cKL.cpp:117,128
  DBuf<double> wL;             // sum of 2.0/k (fp64)                              cEIG.cpp:110
  DBuf<uint32_t> first;        // pair index of the first occurrence (file order)
  DBuf<int32_t> fstart;        // n+1 : first edge with a == v
  DBuf<uint32_t> perm_b;       // U : edge ids sorted by (b, a)
  DBuf<int32_t> bstart;        // n+1 : first position in perm_b with b == v
  bool valid = false;
};

// node relabelling used by the EIG stage only (the KL stage keeps the file's ids: its fp32 summation order
// and tie-breaks are defined by them): nodes sorted by the first net that mentions them
struct NodeOrder {
  DBuf<int32_t> perm;          // new id -> file id
  DBuf<int32_t> inv;           // file id -> new id
  DBuf<int32_t> pins;          // the pins relabelled
  DBuf<uint32_t> first;        // scratch: first net of every node
  bool active = false;         // false: identity (EIGKL_F_NATURAL_ORDER)
  bool valid = false;
};

struct LaplacianCsr {          // fp64, symmetric, rows ascending by column, diagonal included
  int32_t n = 0;
  int64_t nnz = 0;
  DBuf<int32_t> rowptr, col;
  DBuf<double> val;
  // this rank's rows [row_lo, row_hi) (everything when nranks == 1) cut into row blocks of the
  // adaptive SpMV: block b owns rows [blk_row[b], blk_row[b+1])
  int32_t row_lo = 0, row_hi = 0;
  int32_t n_blocks = 0;
  DBuf<int32_t> blk_row;
  DBuf<int32_t> blk_info;      // 4 ints per row block: r0, r1, first entry, end entry (the flat SpMV's descriptor)
  bool flat = false;           // row blocks are sized for / run by spmv_flat_kernel
  // resident polynomial filter (single rank): at most one row block per SM, small enough that its matrix
  // entries stay in registers / shared memory across all SpMVs of one filter application
  DBuf<int32_t> res_info;      // 4 ints per resident block, as blk_info
  DBuf<int32_t> res_row;       // first row of each resident block
  DBuf<int32_t> res_check;     // [0] = longest block span, [1] = most rows in a block, [2] = largest halo
  int32_t res_check_host[4] = {0, 0, 0, 0};
  DBuf<uint16_t> res_src;      // per entry: index into the owning CTA's x cache, bit 15 = row start
  DBuf<int32_t> res_halo_ids, res_halo_cnt;   // per block: sorted distinct columns outside its rows
  int64_t res_chunk = 0;
  int res_k = 0;               // entries per thread of the resident kernel (4, 8 or 16)
  DBuf<unsigned long long> res_ll;   // halo exchange buffer: 2 slots x n x {lo, tag, hi, tag}
  uint32_t res_tag = 0;        // tags handed out so far (monotonic; y_k of a launch carries res_tag + k)
  int32_t res_blocks = 0;
  bool res_ok = false;
  DBuf<unsigned long long> diag_minmax;   // [0] = orderable(min L_ii) complemented, [1] = orderable(max L_ii)
  double diag_min = 0, diag_max = 0;      // spectrum bounds: lambda_max <= 2 max L_ii, lambda_2 <= n/(n-1) min L_ii
  bool valid = false;
};

// ---------------------------------------------------------------------------------------------------
// Row-partitioned Lanczos across ranks (nranks > 1 and the matrix does not fit the chip).
// Rank r owns rows [cuts[r], cuts[r+1]) -- nnz-balanced cuts on multiples of 32 -- of L and of every Lanczos
// vector.  A rank's SpMV input lives in ONE buffer of R slots x n_pad doubles: slot `me` holds its own rows,
// slot p != me holds, packed in ascending column order, the HALO it needs from rank p (the distinct columns in
// p's range that its rows reference).  L is symmetric, so "rows of p that q references" can be computed by p
// from its own rows alone: no plan exchange between ranks.  Producers push their export rows straight into the
// consumers' slots over NVLink (peer-mapped memory) from the SpMV epilogue, then raise a flag (dist.cu, spmv.cu).
// ---------------------------------------------------------------------------------------------------
constexpr int EIGKL_MAX_RANKS = 16;
struct DistPlan {
  int R = 1, me = 0;
  int32_t cuts[EIGKL_MAX_RANKS + 1] = {0};
  int32_t n_pad = 0;             // slot size: the largest row count of a rank, rounded up to 32
  int32_t nl = 0;                // rows of this rank
  int32_t e_lo = 0;              // rowptr[cuts[me]]: first entry of this rank's rows
  int64_t nnz_l = 0;
  DBuf<int32_t> col_c;           // nnz_l: index into the x buffer (slot * n_pad + position) of every local entry
  DBuf<uint32_t> bm_halo;        // one bit per column: referenced by my rows and owned by another rank
  DBuf<int32_t> pre_halo;        // exclusive popcount prefix over bm_halo's words
  DBuf<uint32_t> bm_exp;         // R bitmaps over my rows: row is referenced by rank q
  DBuf<int32_t> pre_exp;
  DBuf<int32_t> exp_ids;         // R x n_pad: my rows (local offsets) that rank q needs, ascending
  DBuf<int32_t> exp_cnt;         // R
  DBuf<int32_t> blk_exp;         // (n_blocks + 1) x R: export rows of rank q below the row block's first row
  int32_t exp_cnt_host[EIGKL_MAX_RANKS] = {0};
  int32_t halo_cnt_host[EIGKL_MAX_RANKS] = {0};
  bool valid = false;
};

// Peer-mapped exchange arena (cudaIpc): [flags | w0 | w1 | w2 | stage], each vector R x n_pad doubles.  Grow-only;
// the mappings are re-opened only when the arena grows.
struct PeerArena {
  void *base = nullptr;
  size_t bytes = 0;
  size_t vec_bytes = 0;          // bytes of one vector buffer in the current layout
  void *peer[EIGKL_MAX_RANKS] = {nullptr};   // peer[q]: rank q's arena in this process' address space
  DBuf<unsigned long long> dev_ptrs;         // the same table on the device
  DBuf<int> err;                 // [0] != 0: a flag wait (1) or an all-reduce wait (2) timed out
  uint32_t seq = 0;              // halo productions issued so far (identical on every rank)
  uint32_t red_seq = 0;          // one-shot all-reduces issued so far (identical on every rank)
  int state = 0;                 // 0 untried, 1 usable, -1 peer mapping unavailable (every rank then solves replicated)
};
constexpr int DIST_RED_MAX = 128;                                      // values per one-shot all-reduce
constexpr size_t PEER_RED_OFFSET = 4096;                               // reduce areas: 2 x EIGKL_MAX_RANKS x DIST_RED_MAX x 16 bytes behind the flags
constexpr size_t PEER_FLAGS_BYTES = PEER_RED_OFFSET + (size_t)2 * EIGKL_MAX_RANKS * DIST_RED_MAX * 16;   // flags + reduce areas; the vectors follow

struct KlCsr {                 // fp32, symmetric, rows in reference traversal order
  int32_t n = 0;
  int64_t nnz = 0;
  DBuf<int32_t> rowptr, fwd_end, col;
  DBuf<float> w;
  int32_t n_blocks = 0;        // row blocks of the staged D-value kernel
  DBuf<int32_t> blk_row;
  DBuf<int32_t> nb;            // 2 per entry: rowptr[col], rowptr[col + 1] (the shared-memory swap loop's one-load lookup)
  bool nb_valid = false;
  DBuf<int32_t> blk_info;      // 4 ints per row block of the D-value kernel: r0, r1, first entry, end entry
  bool info_valid = false;
  bool valid = false;
};

struct KlState {
  DBuf<uint8_t> state;         // bit0 = side, bit1 = locked
  DBuf<uint32_t> side_bits;    // the sides as a bitmap (1 bit per node), packed when a partition is set: the D-value kernel's gather source
  DBuf<uint32_t> rank;         // position in remain[side] (ties -> lowest rank)
  DBuf<float> val;             // connections(v)
  DBuf<unsigned long long> tile_key;   // 2 * n_tiles
  DBuf<uint32_t> tile_stamp;
  DBuf<int32_t> order0, order1;        // remain[0], remain[1]
  int64_t n0 = 0, n1 = 0;
  bool ascending = true;       // remain orders are ascending ids (the -EIG branch)
  bool have_partition = false;
  bool consumed = false;       // a pass ran over this partition: state[] carries lock bits and swapped sides, orders are stale
  // trace on device
  DBuf<float> t_cut, t_gain;
  DBuf<int32_t> t_n1, t_n2;
  DBuf<int64_t> ctrl;          // [0] swaps, [1] status
  DBuf<int64_t> mctrl;         // multi-rank loop state (KlCtrl, 32 bytes)
  int64_t swaps = 0;
  int64_t cap = 0;
};

struct EigState {
  int32_t n = 0, ncv = 0;
  size_t ld = 0;               // leading dimension of the basis (n rounded up)
  DBuf<double> V[2];           // two banks of ld*(ncv+1)
  int bank = 0;
  DBuf<double> w[3];           // Lanczos / Chebyshev recurrence work vectors (rotating)
  DBuf<double> partial;        // multidot / norm partials
  DBuf<double> hcoef;          // 2*(ncv+1) : h of pass 1 and 2
  DBuf<double> alpha, beta;    // ncv each
  DBuf<double> scal;           // [0] norm^2, [1] 1/beta, ...
  DBuf<unsigned int> counters; // last-block-done counters
  DBuf<int> flag;              // [0] = 1: second Gram-Schmidt pass of the current step is skipped
  HBuf<double> snap;           // pinned snapshot of alpha / beta for the deferred convergence checks
  DBuf<double> gs_partial;     // fused Gram-Schmidt kernel: per-CTA partial dot products / norms
  DBuf<unsigned int> gs_sync;  // [0] grid-barrier arrivals (monotonic within a solve), [1] last-CTA ticket
  DBuf<double> Y;              // ncv*ncv restart coefficients (column major, ld = ncv)
  DBuf<double> xfull;          // nranks * n_pad : all-gathered SpMV input (multi-rank only)
  DBuf<double> fiedler;        // n : result vector, file node ids
  DBuf<double> fiedler_perm;   // n (nranks * n_pad when multi-rank) : result vector in the solver's node order
  DBuf<uint8_t> side;          // n : partition from the Fiedler vector
  DBuf<unsigned long long> sortkey[2];
  DBuf<uint32_t> sortval[2];
  double lambda2 = 0, median = 0;
  bool have_vector = false, have_median = false;
};

}  // namespace eigkl

// the opaque handle
struct eigkl_handle {
  eigkl_opts opts{};
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::string err;
  eigkl_stats stats{};
  eigkl::StageTimer timer;
  eigkl::KernelProfiler prof;
  eigkl::Hypergraph hg;
  eigkl::UniqueEdges ue;       // edges in the file's node ids (KL graph; Laplacian when the order is natural)
  eigkl::UniqueEdges ueL;      // edges in the relabelled ids (Laplacian)
  eigkl::NodeOrder order;
  eigkl::LaplacianCsr L;
  eigkl::DistPlan dist;
  eigkl::PeerArena arena;
  int dist_mode = 0;           // EIGKL_DIST: 0 auto (replicate when the matrix fits the chip), 1 always row-partition, 2 always replicate
  int kl_dist = 0;             // EIGKL_KL_DIST=1: KL partitioned by node range with one NCCL arg-max per swap (tested option)
  eigkl::KlCsr A;
  eigkl::KlState kl;
  eigkl::EigState eig;
  int spmv_mode = 0;           // EIGKL_SPMV_MODE: 0 auto (flat), 1 staged, 2 sub-warp, 3 flat (tuning aid)
  int spmv_pdl = 1;            // EIGKL_SPMV_PDL=0 disables programmatic dependent launch of the SpMV chain
  int spmv_resident = 1;       // EIGKL_SPMV_RESIDENT=0: one launch per SpMV even when the matrix fits on chip
  int gs_fused = 1;            // EIGKL_GS_FUSED=0: Gram-Schmidt as separate multidot / update launches
  // per-handle (= per-device) one-time kernel attribute set-up, and what the device allows
  bool attr_kl_local = false, attr_gs = false, attr_resident = false;
  unsigned attr_kl_flat = 0;   // one bit per variant of the flat KL loop whose shared-memory limit has been raised
  int coop_ok = -1;            // -1 unknown, 0/1: cudaDevAttrCooperativeLaunch
  int kl_local = 1;            // EIGKL_KL_LOCAL=0: never run the swap loop with its state in shared memory
  int kl_flat = 1;             // EIGKL_KL_FLAT=0: the warp-per-row form of the shared-memory swap loop (kl_loop_local_kernel)
  void *nccl_comm = nullptr;   // ncclComm_t when nranks > 1
  void *l2_flush = nullptr;    // >L2 scratch for eigkl_time_kernel
  // scratch of the sort / scan primitives
  eigkl::DBuf<int32_t> sort_hist;
  eigkl::DBuf<int64_t> scan_tmp;
  // stage-local temporaries (grow-only, reused across calls; stream order keeps reuse safe)
  struct {
    eigkl::DBuf<int32_t> i32a, i32b;
    eigkl::DBuf<uint32_t> u32a, u32b;
    eigkl::DBuf<uint8_t> u8a;
    eigkl::DBuf<int> err;
  } scr;
  int64_t launches = 0;
};

namespace eigkl {

// ---- primitives (scan_sort.cu) --------------------------------------------------------------------
// exclusive scan of n int64 values (in -> out, may alias); returns nothing, total = out[n] if with_total
void exclusive_scan_i64(eigkl_handle *h, const int64_t *in, int64_t *out, int64_t n);
void exclusive_scan_i32(eigkl_handle *h, const int32_t *in, int32_t *out, int64_t n);
// stable LSD radix sort of (key, val) pairs on bits [0, nbits) of the key.  keys/vals are double
// buffers; returns the index (0/1) of the buffer holding the result.
int radix_sort_kv(eigkl_handle *h, unsigned long long *keys[2], uint32_t *vals[2], int64_t n, int nbits);
int bits_for(uint64_t max_value);

// ---- assembly (assemble.cu) ------------------------------------------------------------------------
void upload_pins(eigkl_handle *h, int32_t n_nodes, int32_t n_nets, const int64_t *net_off, const int32_t *pins);
void build_unique_edges(eigkl_handle *h, UniqueEdges &ue, const int32_t *pins);   // sort + segmented reduce over net pins
void assemble_laplacian(eigkl_handle *h);
void assemble_kl_graph(eigkl_handle *h);

// ---- EIG (spmv.cu, lanczos.cu, eig_solver.cpp) ------------------------------------------------------
void spmv_launch(eigkl_handle *h, const double *x, double *y, const double *scale_inv /*device or null*/,
                 double *store_scaled /*or null*/);
// the whole Chebyshev recurrence (deg SpMVs) as ONE cooperative launch; out_idx[k] = index into w[] of y_{k+1}
bool cheb_resident_usable(const eigkl_handle *h);
void cheb_resident_launch(eigkl_handle *h, const double *x_in, const double *scale, double *v_store, double *const w[3],
                          const unsigned char *out_idx, int deg, double fc, double fe);
bool device_cooperative(eigkl_handle *h);        // lanczos.cu
void spmv_resident_print_phases();
void cheb_resident_plan(eigkl_handle *h);          // enqueue (no sync)
void cheb_resident_plan_finish(eigkl_handle *h);   // after the stream has been synchronised
void resident_row_blocks(eigkl_handle *h, int64_t chunk);   // assemble.cu
struct SpmvDist {              // row-partitioned SpMV: which halo production the input carries / the output becomes
  uint32_t wait_seq;           // the input buffer's halo slots are complete once every peer's flag reaches this
  uint32_t push_seq;           // 0: do not push; else export rows of y go to the peers' buffer `out_buf` under this number
  int out_buf;
};
void spmv_launch_ex(eigkl_handle *h, const double *xg, const double *xl, const double *z, double *y, const double *scale_inv,
                    double *store_scaled, double ca, double cb, double cg, const SpmvDist *dist = nullptr);
// ---- row-partitioned mode (dist.cu) ------------------------------------------------------------------
inline int dist_ranks(const eigkl_handle *h);                      // 1 = every rank solves the whole problem
int64_t dist_decide(eigkl_handle *h);                             // assemble_laplacian: replicate or cut rows; local nnz
void dist_plan(eigkl_handle *h);                                  // halo / export plan once the SpMV row blocks exist
double *dist_buf(eigkl_handle *h, int b);                         // arena vector b (0..2 = w, 3 = stage), R x n_pad doubles
inline double *dist_own(eigkl_handle *h, int b);
void dist_raise_flags(eigkl_handle *h, uint32_t seq);             // spmv.cu: one-warp kernel, release.sys store of seq into every peer's flag word
uint32_t dist_push(eigkl_handle *h, int b);                       // push the own slot's export rows of buffer b; returns its seq
void dist_allreduce_sum(eigkl_handle *h, double *buf, size_t count);   // in place, <= DIST_RED_MAX doubles, over the arena (no NCCL)
void dist_stage_load(eigkl_handle *h, const double *src_local);   // stage.own = src (then dist_push(h, 3))
void dist_check(eigkl_handle *h);                                 // throws when a flag wait timed out
void dist_gather_full(eigkl_handle *h, const double *slice, double *full_natural);   // NCCL all-gather + un-slotting
void peer_arena_destroy(eigkl_handle *h);
void fiedler_solve(eigkl_handle *h);
void partition_from_fiedler(eigkl_handle *h);
void sym_eig(int n, double *a, double *evals);   // dense symmetric eigen-solver (host)
void tridiag_top_eig(int n, const double *d, const double *e, int k, double *theta, double *Y);   // k largest pairs
void sym_top_eig(int n, const double *a, int k, double *theta, double *Y);   // k largest pairs of a dense symmetric matrix

// ---- KL (kl.cu) ---------------------------------------------------------------------------------------
void kl_set_partition(eigkl_handle *h, const uint8_t *side_host, const int32_t *order0, int64_t n0,
                      const int32_t *order1, int64_t n1, bool ascending);
void kl_set_partition_device(eigkl_handle *h, const uint8_t *side_dev);   // ascending orders, built on device
void kl_dvalues(eigkl_handle *h);                 // val[] for every node from the current sides
float kl_cut0(eigkl_handle *h);
void kl_run(eigkl_handle *h);
int64_t kl_rollback(eigkl_handle *h, float *best_cut);

// ---- text I/O (hgr_io.cpp) -----------------------------------------------------------------------------
struct HostHgr {
  int32_t n_nodes = 0, n_nets = 0;
  std::vector<int64_t> net_off;
  std::vector<int32_t> pins;
};
void parse_hgr(const char *path, HostHgr &out);
void write_eig_file(const char *path, double lambda2, double median, const double *vec, int32_t n);
void read_eig_file(const char *path, int32_t n, std::vector<uint8_t> &side, std::vector<int32_t> &order0,
                   std::vector<int32_t> &order1, bool &ascending);
void write_trace_file(const char *path, const eigkl_trace *t);
void write_partition_file(const char *path, const uint8_t *side, int32_t n);

// ---- comm (comm.cpp) --------------------------------------------------------------------------------------
void comm_init(eigkl_handle *h);
void comm_destroy(eigkl_handle *h);
void comm_unique_id(void *id128);
void comm_allreduce_sum_f64(eigkl_handle *h, double *buf, size_t count);            // in place, on h->stream
void comm_allreduce_max_u64(eigkl_handle *h, unsigned long long *buf, size_t count);
void comm_allgather_f64(eigkl_handle *h, const double *send, double *recv, size_t count_per_rank);
void comm_broadcast_bytes(eigkl_handle *h, void *buf, size_t bytes, int root);
void comm_allgather_bytes(eigkl_handle *h, const void *send, void *recv, size_t bytes_per_rank);
void comm_allreduce_min_i32(eigkl_handle *h, int32_t *buf, size_t count);

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int dist_ranks(const eigkl_handle *h) { return h->dist.valid ? h->dist.R : 1; }
inline double *dist_own(eigkl_handle *h, int b) { return dist_buf(h, b) + (size_t)h->dist.me * (size_t)h->dist.n_pad; }

// 1-D row partition of n rows over nranks ranks: equal blocks of n_pad rows (a multiple of 32, so that
// one ncclAllGather of n_pad values per rank rebuilds a full vector in place: global row g lives at
// index g of the gathered buffer); the last ranks may own fewer (or zero) real rows.
inline void row_partition(int32_t n, int nranks, int rank, int32_t *lo, int32_t *hi, int32_t *n_pad) {
  int64_t per = ceil_div(n, nranks);
  per = ceil_div(per, 32) * 32;
  const int64_t l = std::min<int64_t>(n, per * rank), hh = std::min<int64_t>(n, per * (rank + 1));
  *lo = (int32_t)l; *hi = (int32_t)hh; *n_pad = (int32_t)per;
}

}  // namespace eigkl
