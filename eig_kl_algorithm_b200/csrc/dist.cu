// dist.cu -- the row-partitioned (multi-GPU) form of the Fiedler solve: partition, halo plan, peer-mapped arena.
//
// SURVEY.md 8(e) / north star: "the Laplacian and all vectors are row-partitioned across the 8 B200s of one box".
// Policy (eigkl_opts.nranks > 1):
//   * the matrix fits ONE chip's registers + shared memory (cheb_resident_plan succeeds: every shipped circuit)
//     -> every rank runs the complete single-GPU solve ("replicas only"): an ibm10 SpMV is ~6 us on
//     chip, less than any exchange between GPUs, so partitioning it can only lose (round 1 measured 4.2x slower
//     on 8 GPUs than on 1);
//   * otherwise (the 2 M-node synthetic: 157 MB matrix, 1.6 GB basis, HBM-bound) rows are cut into R nnz-balanced
//     ranges and every Lanczos vector is partitioned alike.  The reference has no multi-GPU path at all.
// What replaces the per-SpMV full-vector ncclAllGather of round 1:
//   * x as a rank needs it is ONE buffer of R slots x n_pad doubles -- its own rows in slot `me`, and in slot p the
//     distinct columns of p's range that its rows reference, PACKED in ascending order (the halo).  The column ids
//     of the local CSR entries are rewritten once per matrix into indices of that buffer (col_c);
//   * L is symmetric: q references my row i  <=>  my row i has an entry in q's column range.  So a rank derives
//     both its halo lists and its export lists from its own rows -- the plan needs no communication;
//   * the buffers live in a cudaIpc-mapped arena; the SpMV kernel that produces y_k writes its export rows
//     straight into the consumers' slots over NVLink (coalesced 8-byte stores into packed slots), the last CTA to
//     finish raises one flag per peer (release, system scope), and the SpMV that consumes y_k polls its own flags
//     (acquire) after its constant prologue: transfer and signalling ride inside the compute kernels, no
//     collective call, no host involvement (spmv.cu: spmv_dist_kernel).  Only values that are needed travel.
// The three dot-product reductions per Lanczos step stay ncclAllReduce calls (<= 101 doubles each).
#include "internal.h"
#include "device_utils.cuh"
#include <algorithm>
#include <cstdlib>

namespace eigkl {

namespace {
constexpr int TPB = 256;
inline unsigned grid_for(int64_t n, int tpb = TPB) { return (unsigned)std::max<int64_t>(1, (n + tpb - 1) / tpb); }

struct Cuts { int32_t c[EIGKL_MAX_RANKS + 1]; int R; };
__device__ __forceinline__ int rank_of(const Cuts &k, int32_t row) {
  int r = 0;
#pragma unroll 1
  while (r + 1 < k.R && row >= k.c[r + 1]) ++r;
  return r;
}
}  // namespace

// cut r = the row (a multiple of 32) closest to holding r/R of the solve's bytes: 12 per non-zero (value + column) and
// DIST_ROW_BYTES per row -- the SpMV's 28 (row pointer, x, z, y) plus the row's share of the Gram-Schmidt passes, which
// stream the basis: 4 passes x ~50 columns x 8 bytes per Lanczos step = ~100 bytes per row and SpMV at filter degree 16.
// Balancing the non-zeros alone gave one of two ranks 50 % more rows at 2 M nodes (multidot 69 vs 52 us).
constexpr long long DIST_ROW_BYTES = 128;
__global__ void dist_cuts_kernel(const int32_t *__restrict__ rowptr, int32_t n, int R, int32_t *__restrict__ cuts /* R+1, then R+1 entry offsets */) {
  const int r = threadIdx.x;
  if (r > R) return;
  int32_t row;
  if (r == 0) row = 0;
  else if (r == R) row = n;
  else {
    const int64_t total = 12ll * rowptr[n] + DIST_ROW_BYTES * n;
    const int64_t target = total * r / R;
    int32_t lo = 0, hi = n;
    while (lo < hi) {
      const int32_t mid = (lo + hi) >> 1;
      if (12ll * rowptr[mid] + DIST_ROW_BYTES * mid < target) lo = mid + 1; else hi = mid;
    }
    row = (int32_t)min((int64_t)n, ((int64_t)lo + 16) / 32 * 32);
  }
  cuts[r] = row;
  cuts[R + 1 + r] = rowptr[row];
}

// one thread per local row: mark the remote columns it references (my halo) and, per remote owner q, that q
// references this row (my exports to q) -- the same fact seen from the other side of the symmetric matrix
__global__ void dist_mark_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, Cuts k, int me, int32_t Wl1,
                                 uint32_t *__restrict__ bm_halo, uint32_t *__restrict__ bm_exp) {
  const int32_t il = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t lo = k.c[me], nl = k.c[me + 1] - lo;
  if (il >= nl) return;
  const int32_t i = lo + il;
  unsigned mask = 0u;
  for (int32_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
    const int32_t j = col[e];
    if (j >= lo && j < lo + nl) continue;
    atomicOr(&bm_halo[j >> 5], 1u << (j & 31));
    mask |= 1u << rank_of(k, j);
  }
  while (mask) {
    const int q = __ffs(mask) - 1;
    mask &= mask - 1;
    atomicOr(&bm_exp[(size_t)q * Wl1 + (il >> 5)], 1u << (il & 31));
  }
}
__global__ void dist_popc_kernel(const uint32_t *__restrict__ bm, int64_t words, int32_t *__restrict__ out) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < words) out[w] = __popc(bm[w]);
}
// index of every local entry in the x buffer: own rows -> slot me at the row's offset; remote -> slot p, packed
__global__ void dist_remap_kernel(const int32_t *__restrict__ col, int32_t e_lo, int64_t nnz_l, Cuts k, int me, int32_t n_pad,
                                  const uint32_t *__restrict__ bm_halo, const int32_t *__restrict__ pre_halo,
                                  int32_t *__restrict__ col_c) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz_l) return;
  const int32_t j = col[e_lo + e];
  const int p = rank_of(k, j);
  int32_t idx;
  if (p == me) idx = me * n_pad + (j - k.c[me]);
  else idx = p * n_pad + (pre_halo[j >> 5] - pre_halo[k.c[p] >> 5]) + __popc(bm_halo[j >> 5] & ((1u << (j & 31)) - 1u));
  col_c[e] = idx;
}
__global__ void dist_export_fill_kernel(const uint32_t *__restrict__ bm_exp, const int32_t *__restrict__ pre_exp, int R, int32_t Wl1,
                                        int32_t n_pad, int32_t *__restrict__ exp_ids, int32_t *__restrict__ exp_cnt,
                                        const int32_t *__restrict__ pre_halo, Cuts k, int32_t *__restrict__ halo_cnt) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < R) {
    exp_cnt[t] = pre_exp[(size_t)(t + 1) * Wl1 - 1] - pre_exp[(size_t)t * Wl1];     // the last word of a bitmap is always empty
    const int32_t w0 = k.c[t] >> 5, w1 = (t + 1 == R) ? (k.c[R] + 31) >> 5 : k.c[t + 1] >> 5;
    halo_cnt[t] = pre_halo[w1] - pre_halo[w0];
  }
  if (t >= (int64_t)R * Wl1) return;
  const int q = (int)(t / Wl1);
  const int32_t w = (int32_t)(t % Wl1);
  uint32_t bits = bm_exp[t];
  int32_t pos = pre_exp[t] - pre_exp[(size_t)q * Wl1];
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    exp_ids[(size_t)q * n_pad + pos++] = (w << 5) + b;
  }
}
// export rows of rank q below each row block's first row: the block's segment of q's export list is [blk_exp[b][q], blk_exp[b+1][q])
__global__ void dist_blk_exp_kernel(const int32_t *__restrict__ blk_row, int32_t n_blocks, int32_t row_lo, int R, int32_t Wl1,
                                    const uint32_t *__restrict__ bm_exp, const int32_t *__restrict__ pre_exp,
                                    int32_t *__restrict__ blk_exp) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)(n_blocks + 1) * R) return;
  const int32_t b = (int32_t)(t / R);
  const int q = (int)(t % R);
  const int32_t r = blk_row[b] - row_lo;
  const size_t w = (size_t)q * Wl1 + (r >> 5);
  blk_exp[t] = pre_exp[w] - pre_exp[(size_t)q * Wl1] + __popc(bm_exp[w] & ((1u << (r & 31)) - 1u));
}

// ---------------------------------------------------------------------------------------------------
// the peer-mapped arena
// ---------------------------------------------------------------------------------------------------
static void arena_close(eigkl_handle *h) {
  auto &a = h->arena;
  const int R = h->opts.nranks, me = h->opts.rank;
  for (int q = 0; q < R && q < EIGKL_MAX_RANKS; ++q) {
    if (q != me && a.peer[q]) cudaIpcCloseMemHandle(a.peer[q]);
    a.peer[q] = nullptr;
  }
}
void peer_arena_destroy(eigkl_handle *h) {
  arena_close(h);
  if (h->arena.base) cudaFree(h->arena.base);
  h->arena.base = nullptr;
  h->arena.bytes = 0;
}

// Collective over the ranks.  Returns false (on every rank alike) when peer mapping is not available.
static bool peer_arena_ensure(eigkl_handle *h, size_t vec_bytes) {
  auto &a = h->arena;
  const int R = h->opts.nranks, me = h->opts.rank;
  if (a.state < 0) return false;
  const size_t need = PEER_FLAGS_BYTES + 4 * vec_bytes;
  a.vec_bytes = vec_bytes;
  if (a.state == 1 && need <= a.bytes) return true;
  cudaStream_t st = h->stream;
  DBuf<int32_t> flag; flag.alloc(1);
  auto agree = [&](int ok) -> bool {              // min over the ranks; doubles as a barrier
    int32_t v = ok;
    EIGKL_CUDA(cudaMemcpyAsync(flag.p, &v, sizeof(v), cudaMemcpyHostToDevice, st));
    comm_allreduce_min_i32(h, flag.p, 1);
    EIGKL_CUDA(cudaMemcpyAsync(&v, flag.p, sizeof(v), cudaMemcpyDeviceToHost, st));
    EIGKL_CUDA(cudaStreamSynchronize(st));
    return v != 0;
  };
  // nobody may still map the old arena when it is freed
  arena_close(h);
  agree(1);
  if (a.base) { cudaFree(a.base); a.base = nullptr; a.bytes = 0; }
  int ok = 1;
  const size_t bytes = need + need / 4;
  if (getenv("EIGKL_NO_PEER") != nullptr) ok = 0;
  if (ok && cudaMalloc(&a.base, bytes) != cudaSuccess) { cudaGetLastError(); a.base = nullptr; ok = 0; }
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) {
    EIGKL_CUDA(cudaMemsetAsync(a.base, 0, bytes, st));
    if (cudaIpcGetMemHandle(&mine, a.base) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  DBuf<unsigned char> dh; dh.alloc((size_t)64 * R);
  std::vector<cudaIpcMemHandle_t> all((size_t)R);
  EIGKL_CUDA(cudaMemcpyAsync(dh.p + (size_t)64 * me, &mine, 64, cudaMemcpyHostToDevice, st));
  comm_allgather_bytes(h, dh.p + (size_t)64 * me, dh.p, 64);
  EIGKL_CUDA(cudaMemcpyAsync(all.data(), dh.p, (size_t)64 * R, cudaMemcpyDeviceToHost, st));
  EIGKL_CUDA(cudaStreamSynchronize(st));
  const bool everyone_allocated = agree(ok);
  if (everyone_allocated) {
    for (int q = 0; q < R; ++q) {
      if (q == me) { a.peer[q] = a.base; continue; }
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[(size_t)q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; p = nullptr; }
      a.peer[q] = p;
    }
  }
  const bool usable = everyone_allocated && agree(ok);
  if (!usable) {
    arena_close(h);
    agree(1);
    if (a.base) { cudaFree(a.base); a.base = nullptr; }
    a.bytes = 0;
    a.state = -1;
    return false;
  }
  a.bytes = bytes;
  a.state = 1;
  a.seq = 0;
  a.red_seq = 0;
  a.dev_ptrs.alloc(EIGKL_MAX_RANKS);
  unsigned long long tbl[EIGKL_MAX_RANKS] = {0};
  for (int q = 0; q < R; ++q) tbl[q] = (unsigned long long)(uintptr_t)a.peer[q];
  EIGKL_CUDA(cudaMemcpyAsync(a.dev_ptrs.p, tbl, sizeof(tbl), cudaMemcpyHostToDevice, st));
  a.err.alloc(4);
  EIGKL_CUDA(cudaMemsetAsync(a.err.p, 0, 4 * sizeof(int), st));
  EIGKL_CUDA(cudaStreamSynchronize(st));
  return true;
}

double *dist_buf(eigkl_handle *h, int b) {
  return reinterpret_cast<double *>(static_cast<char *>(h->arena.base) + PEER_FLAGS_BYTES + (size_t)b * h->arena.vec_bytes);
}

// ---------------------------------------------------------------------------------------------------
// decide (after the resident plan is known) and cut the rows; called by assemble_laplacian before it builds the
// SpMV row blocks.  Sets L.row_lo / row_hi and returns the number of local non-zeros.
// ---------------------------------------------------------------------------------------------------
int64_t dist_decide(eigkl_handle *h) {
  auto &L = h->L;
  auto &D = h->dist;
  D.valid = false;
  const int R = h->opts.nranks, me = h->opts.rank;
  const int32_t n = L.n;
  L.row_lo = 0; L.row_hi = n;
  D.R = 1; D.me = 0;
  if (R <= 1) return L.nnz;
  EIGKL_REQUIRE(R <= EIGKL_MAX_RANKS, EIGKL_E_ARG, "at most 16 ranks");
  const bool want = h->dist_mode == 1 || (h->dist_mode == 0 && !L.res_ok);
  if (!want || n < 64 * R) return L.nnz;
  cudaStream_t st = h->stream;
  auto &cd = h->scr.i32a; cd.alloc((size_t)2 * (R + 1));
  dist_cuts_kernel<<<1, 32, 0, st>>>(L.rowptr.p, n, R, cd.p);
  h->launches++;
  int32_t hc[2 * (EIGKL_MAX_RANKS + 1)];
  EIGKL_CUDA(cudaMemcpyAsync(hc, cd.p, (size_t)2 * (R + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  EIGKL_CUDA(cudaStreamSynchronize(st));
  int32_t most = 0;
  for (int r = 0; r <= R; ++r) {
    if (r > 0 && hc[r] < hc[r - 1]) hc[r] = hc[r - 1];          // monotone whatever the rounding did
    D.cuts[r] = hc[r];
    if (r > 0) most = std::max(most, hc[r] - hc[r - 1]);
  }
  D.n_pad = (int32_t)(ceil_div(std::max(most, 32), 32) * 32);
  // the arena: collective, identical decision on every rank
  if (!peer_arena_ensure(h, (size_t)R * (size_t)D.n_pad * sizeof(double))) return L.nnz;
  D.R = R; D.me = me;
  D.nl = D.cuts[me + 1] - D.cuts[me];
  D.e_lo = hc[R + 1 + me];
  D.nnz_l = (int64_t)hc[R + 1 + me + 1] - hc[R + 1 + me];
  L.row_lo = D.cuts[me]; L.row_hi = D.cuts[me + 1];
  D.valid = true;
  return D.nnz_l;
}

// halo / export plan of this rank (after the SpMV row blocks exist)
void dist_plan(eigkl_handle *h) {
  auto &L = h->L;
  auto &D = h->dist;
  if (!D.valid) return;
  cudaStream_t st = h->stream;
  const int R = D.R, me = D.me;
  const int32_t n = L.n;
  const int32_t W = (n + 31) / 32;
  const int32_t Wl1 = D.n_pad / 32 + 1;                   // words per export bitmap, the last one always empty
  Cuts k;
  for (int r = 0; r <= EIGKL_MAX_RANKS; ++r) k.c[r] = r <= R ? D.cuts[r] : n;
  k.R = R;
  D.bm_halo.alloc((size_t)W + 1); D.pre_halo.alloc((size_t)W + 2);
  D.bm_exp.alloc((size_t)R * Wl1); D.pre_exp.alloc((size_t)R * Wl1 + 1);
  D.exp_ids.alloc((size_t)R * D.n_pad); D.exp_cnt.alloc(2 * EIGKL_MAX_RANKS);
  D.col_c.alloc((size_t)std::max<int64_t>(D.nnz_l, 1));
  D.blk_exp.alloc((size_t)(L.n_blocks + 1) * R);
  EIGKL_CUDA(cudaMemsetAsync(D.bm_halo.p, 0, ((size_t)W + 1) * sizeof(uint32_t), st));
  EIGKL_CUDA(cudaMemsetAsync(D.bm_exp.p, 0, (size_t)R * Wl1 * sizeof(uint32_t), st));
  if (D.nl > 0) dist_mark_kernel<<<grid_for(D.nl), TPB, 0, st>>>(L.rowptr.p, L.col.p, k, me, Wl1, D.bm_halo.p, D.bm_exp.p);
  dist_popc_kernel<<<grid_for(W + 1), TPB, 0, st>>>(D.bm_halo.p, W + 1, D.pre_halo.p);
  exclusive_scan_i32(h, D.pre_halo.p, D.pre_halo.p, W + 1);
  dist_popc_kernel<<<grid_for((int64_t)R * Wl1), TPB, 0, st>>>(D.bm_exp.p, (int64_t)R * Wl1, D.pre_exp.p);
  exclusive_scan_i32(h, D.pre_exp.p, D.pre_exp.p, (int64_t)R * Wl1);
  if (D.nnz_l > 0)
    dist_remap_kernel<<<grid_for(D.nnz_l), TPB, 0, st>>>(L.col.p, D.e_lo, D.nnz_l, k, me, D.n_pad, D.bm_halo.p, D.pre_halo.p, D.col_c.p);
  dist_export_fill_kernel<<<grid_for((int64_t)R * Wl1), TPB, 0, st>>>(D.bm_exp.p, D.pre_exp.p, R, Wl1, D.n_pad, D.exp_ids.p, D.exp_cnt.p,
                                                                    D.pre_halo.p, k, D.exp_cnt.p + EIGKL_MAX_RANKS);
  dist_blk_exp_kernel<<<grid_for((int64_t)(L.n_blocks + 1) * R), TPB, 0, st>>>(L.blk_row.p, L.n_blocks, L.row_lo, R, Wl1, D.bm_exp.p,
                                                                             D.pre_exp.p, D.blk_exp.p);
  h->launches += 6;
  int32_t cnt[2 * EIGKL_MAX_RANKS];
  EIGKL_CUDA(cudaMemcpyAsync(cnt, D.exp_cnt.p, sizeof(cnt), cudaMemcpyDeviceToHost, st));
  EIGKL_CUDA(cudaStreamSynchronize(st));
  EIGKL_CUDA(cudaGetLastError());
  for (int q = 0; q < R; ++q) { D.exp_cnt_host[q] = cnt[q]; D.halo_cnt_host[q] = cnt[EIGKL_MAX_RANKS + q]; }
  int64_t halo = 0, exp = 0;
  for (int q = 0; q < R; ++q) { halo += D.halo_cnt_host[q]; exp += D.exp_cnt_host[q]; }
  h->stats.dist_ranks = R;
  h->stats.dist_rows = D.nl;
  h->stats.dist_halo = halo;
  h->stats.dist_exports = exp;
}

// ---------------------------------------------------------------------------------------------------
// stand-alone halo push (the first SpMV of a filter application, whose input comes out of the Gram-Schmidt
// kernels, and the few SpMVs on basis columns): packed export rows -> the peers' slot `me`, then the flags
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
halo_push_kernel(const double *__restrict__ own, const int32_t *__restrict__ exp_ids, const int32_t *__restrict__ exp_cnt,
                 const unsigned long long *__restrict__ peers, size_t buf_off, int32_t n_pad, int me) {
  const int q = blockIdx.y;
  if (q == me) return;
  const int32_t cnt = exp_cnt[q];
  double *dst = reinterpret_cast<double *>(peers[q] + buf_off) + (size_t)me * n_pad;
  const int32_t *ids = exp_ids + (size_t)q * n_pad;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) dst[i] = own[ids[i]];
}

uint32_t dist_push(eigkl_handle *h, int b) {
  auto &D = h->dist;
  auto &a = h->arena;
  const uint32_t seq = ++a.seq;
  int32_t most = 1;
  for (int q = 0; q < D.R; ++q) most = std::max(most, D.exp_cnt_host[q]);
  dim3 grid((unsigned)std::min<int64_t>(256, ceil_div(most, 1024)), (unsigned)D.R);
  const size_t off = PEER_FLAGS_BYTES + (size_t)b * a.vec_bytes;
  h->prof.begin(KC_PUSH, h->stream);
  halo_push_kernel<<<grid, 256, 0, h->stream>>>(dist_own(h, b), D.exp_ids.p, D.exp_cnt.p, a.dev_ptrs.p, off, D.n_pad, D.me);
  h->launches++;
  dist_raise_flags(h, seq);                        // the stores of the kernel above are complete when this one starts
  h->prof.end(h->stream);
  return seq;
}

void dist_stage_load(eigkl_handle *h, const double *src_local) {
  if (h->dist.nl > 0)
    EIGKL_CUDA(cudaMemcpyAsync(dist_own(h, 3), src_local, (size_t)h->dist.nl * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
}

void dist_check(eigkl_handle *h) {
  if (!h->dist.valid) return;
  int e[4] = {0, 0, 0, 0};
  EIGKL_CUDA(cudaMemcpyAsync(e, h->arena.err.p, sizeof(e), cudaMemcpyDeviceToHost, h->stream));
  EIGKL_CUDA(cudaStreamSynchronize(h->stream));
  if (e[0] != 0) {
    EIGKL_CUDA(cudaMemsetAsync(h->arena.err.p, 0, sizeof(e), h->stream));
    throw Error(EIGKL_E_NCCL, e[0] == 2 ? "row-partitioned solve: a peer's partial sums did not arrive (all-reduce wait timed out; ranks out of step?)"
                                        : "row-partitioned SpMV: a peer's halo did not arrive (flag wait timed out; ranks out of step?)");
  }
}


// ---------------------------------------------------------------------------------------------------
// One-shot all-reduce of <= 128 doubles over the peer-mapped arena (the Gram-Schmidt coefficients and the norms of a
// Lanczos step; an ncclAllReduce of that size measured 35 us per call, 280 calls per solve at 2 M nodes).  Every rank
// stores its partials into slot `me` of every rank's reduce area as 16-byte {lo, tag, hi, tag} words (8-byte stores are
// single-copy atomic, over NVLink too: a reader that sees both tags has the value -- no fence, no flag), then polls the
// R slots of its own area and adds them in rank order: the same sum, bit for bit, on every rank.  Two areas alternate:
// a rank can be at most one reduction ahead of a peer that has not yet read the previous one.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DIST_RED_MAX)
dist_allreduce_kernel(double *__restrict__ buf, int count, const unsigned long long *__restrict__ peers, int me, int R, uint32_t tag, int parity,
                      int *__restrict__ err) {
  const int t = threadIdx.x;
  if (t >= count) return;
  const double mine = buf[t];
  const unsigned long long bits = (unsigned long long)__double_as_longlong(mine);
  const unsigned long long w0 = (bits & 0xFFFFFFFFull) | ((unsigned long long)tag << 32), w1 = (bits >> 32) | ((unsigned long long)tag << 32);
  const size_t slot = PEER_RED_OFFSET + (((size_t)parity * EIGKL_MAX_RANKS + me) * DIST_RED_MAX + t) * 16;
  for (int q = 0; q < R; ++q) {
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(peers[q] + slot);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
  }
  double sum = 0.0;
  const long long t0 = clock64();
  for (int r = 0; r < R; ++r) {
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(peers[me] + PEER_RED_OFFSET +
                                                                                 (((size_t)parity * EIGKL_MAX_RANKS + r) * DIST_RED_MAX + t) * 16);
    unsigned long long a, b;
    for (;;) {
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
      if ((uint32_t)(a >> 32) == tag && (uint32_t)(b >> 32) == tag) break;
      if (clock64() - t0 > 6000000000ll) { atomicExch(err, 2); break; }       // ~3 s: ranks out of step; reported by dist_check
    }
    sum += __longlong_as_double((long long)((a & 0xFFFFFFFFull) | (b << 32)));
  }
  buf[t] = sum;
}

void dist_allreduce_sum(eigkl_handle *h, double *buf, size_t count) {
  auto &a = h->arena;
  EIGKL_REQUIRE(count <= (size_t)DIST_RED_MAX, EIGKL_E_ARG, "dist_allreduce_sum: too many values");
  const uint32_t tag = ++a.red_seq;
  dist_allreduce_kernel<<<1, DIST_RED_MAX, 0, h->stream>>>(buf, (int)count, a.dev_ptrs.p, h->dist.me, h->dist.R, tag, (int)(tag & 1u), a.err.p);
  h->launches++;
}

// slice (n_pad doubles, this rank's rows) of every rank -> the full vector in natural row order on every rank
__global__ void unslot_kernel(const double *__restrict__ slots, Cuts k, int32_t n_pad, int32_t n, double *__restrict__ out) {
  const int32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int r = rank_of(k, g);
  out[g] = slots[(size_t)r * n_pad + (g - k.c[r])];
}
void dist_gather_full(eigkl_handle *h, const double *slice, double *full_natural) {
  auto &D = h->dist;
  auto &tmp = h->eig.xfull;
  tmp.ensure((size_t)D.R * D.n_pad);
  comm_allgather_f64(h, slice, tmp.p, (size_t)D.n_pad);
  Cuts k;
  for (int r = 0; r <= EIGKL_MAX_RANKS; ++r) k.c[r] = r <= D.R ? D.cuts[r] : h->L.n;
  k.R = D.R;
  unslot_kernel<<<grid_for(h->L.n), TPB, 0, h->stream>>>(tmp.p, k, D.n_pad, h->L.n, full_natural);
  h->launches++;
}

}  // namespace eigkl
