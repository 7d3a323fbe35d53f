// kl.cu -- the Kernighan-Lin pass on the GPU (north-star subsystem 3), bit-exact with cKL.cpp.
//
// Reference being replaced: KL(), cKL.cpp:288-390 with connections() 225-251, calCutSize() 199-223,
// the pair selection 337-360, swip() 274-286 and updateAffectedNodeGains() 253-272; and the
// reference's own GPU attempt, connectionsKernel + 6 host<->device crossings per swap (gKL.cu:104-145,
// 188-227, 417-549).
//
// Design (B200-first):
//  * D-value gain kernel (kl_dvalues): CSR rows in the reference's traversal order, cut into row blocks
//    of ~2048 entries; a CTA streams col/w coalesced, gathers the neighbour's side byte, stages the
//    SIGNED weight (+w external, -w internal) in shared memory, then one thread per row adds its
//    entries strictly in order into two fp32 accumulators (E, I) -- the order and the two-accumulator
//    shape are what make the result bit-identical to cKL.cpp:225-251.
//  * The whole swap loop is ONE persistent kernel on one thread-block cluster (1..16 CTAs x 1024
//    threads) -- KL is latency bound, so cluster barriers (~0.2 us) replace the
//    reference's per-swap memcpys and a grid-wide sync.  Per swap:
//      S1  argmax pair: max over per-tile cached keys (tile = 256 nodes);
//          key = orderable(D) : ~position, so the max is the first maximum in remain[] order
//      S2  gain = (D1 - D2) - 2 w(a,b); cut -= gain; trace row; termination counter
//      S3  lock a, b, flip sides; every neighbour of a or b gets its D recomputed FROM SCRATCH by a
//          warp (32 entries at a time, coalesced load + side gather, then the ordered two-accumulator
//          sum replayed through warp shuffles)
//      S4  tiles that contain a touched node are rescanned (first warp to stamp the tile does it)
//    Nothing returns to the host until the pass ends; the trace is buffered in HBM.
//    Three forms of that loop, byte-identical in their traces: kl_loop_kernel (state in global memory, 1..16-CTA cluster:
//    graphs beyond 2 M nodes, or kl_cluster asked for), kl_loop_local_kernel (one CTA, keys and side bits in shared
//    memory, a warp per neighbour row) and kl_loop_flat_kernel (the default up to 2 M nodes: the same shared-memory state,
//    the neighbour rows laid out as one flat list of entries; see the comment above it).
//  * Initial cut (kl_cut0): the reference's one-thread evaluation order, including the iteration order
//    of its unordered_set of right nodes, rebuilt with sorts (stl_order.h explains the rule).
#include "internal.h"
#include "device_utils.cuh"
#include "stl_order.h"
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace eigkl {

constexpr int KL_TILE = 256;            // nodes per argmax tile
constexpr int KLD_THREADS = 256;
constexpr int KLD_STAGE = 4096;         // staged signed weights per CTA (16 KB)
constexpr int KL_LOOP_THREADS = 1024;
constexpr int KL_MAX_CLUSTER = 16;
constexpr int TPB = 256;
static inline unsigned grid_for(int64_t n, int tpb = TPB) { return (unsigned)std::max<int64_t>(1, (n + tpb - 1) / tpb); }

#define ST_SIDE 1u
#define ST_LOCK 2u

// ---------------------------------------------------------------------------------------------------
// ordered two-accumulator sum of one row by a warp                       cKL.cpp:225-251
//   E = sum of w to neighbours NOT in the left side, I = sum of w to neighbours in the left side,
//   each added strictly in row order; returns E - I (all lanes).
//   ov_a / ov_b: nodes whose side is taken as 1 / 0 regardless of state[] (the pair being swapped).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_row_value(const int32_t *__restrict__ col, const float *__restrict__ w,
                                                const uint8_t *state, int32_t lo, int32_t hi, int32_t ov_a,
                                                int32_t ov_b, int lane) {
  float E = 0.0f, I = 0.0f;
  int32_t i = lo + lane;
  int32_t c = 0;
  float ww = 0.0f;
  if (i < hi) { c = __ldg(col + i); ww = __ldg(w + i); }
  for (int32_t base = lo; base < hi; base += 32) {
    const bool valid = (base + lane) < hi;
    // prefetch the next 32 entries while this chunk is being summed
    int32_t cn = 0;
    float wn = 0.0f;
    const int32_t in = base + 32 + lane;
    if (in < hi) { cn = __ldg(col + in); wn = __ldg(w + in); }
    float x = 0.0f;
    if (valid) {
      unsigned s;
      if (c == ov_a) s = 1u;
      else if (c == ov_b) s = 0u;
      else s = (unsigned)__ldcg(state + c) & ST_SIDE;
      x = s ? ww : -ww;
    }
    const int cnt = min(32, hi - base);
    for (int t = 0; t < cnt; ++t) {
      const float xt = __shfl_sync(FULL_MASK, x, t);
      E = __fadd_rn(E, fmaxf(xt, 0.0f));
      I = __fadd_rn(I, fmaxf(-xt, 0.0f));
    }
    c = cn; ww = wn;
  }
  return __fsub_rn(E, I);
}

// ---------------------------------------------------------------------------------------------------
// D-values of every node (cKL.cpp:318-321)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KLD_THREADS)
dvalues_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, const float *__restrict__ w,
               const uint8_t *__restrict__ state, float *__restrict__ val, const int32_t *__restrict__ blk_row) {
  __shared__ float sv[KLD_STAGE];
  const int tid = threadIdx.x;
  const int32_t r0 = blk_row[blockIdx.x], r1 = blk_row[blockIdx.x + 1];
  if (r0 >= r1) return;
  const int32_t e0 = rowptr[r0], e1 = rowptr[r1];
  if (e1 - e0 <= KLD_STAGE) {
#pragma unroll 4
    for (int32_t i = e0 + tid; i < e1; i += KLD_THREADS) {
      const float ww = w[i];
      sv[i - e0] = (state[col[i]] & ST_SIDE) ? ww : -ww;
    }
    __syncthreads();
    for (int32_t r = r0 + tid; r < r1; r += KLD_THREADS) {
      const int32_t lo = rowptr[r] - e0, hi = rowptr[r + 1] - e0;
      float E = 0.0f, I = 0.0f;
      for (int32_t i = lo; i < hi; ++i) {
        const float x = sv[i];
        E = __fadd_rn(E, fmaxf(x, 0.0f));
        I = __fadd_rn(I, fmaxf(-x, 0.0f));
      }
      val[r] = __fsub_rn(E, I);
    }
  } else {                              // a row longer than the staging buffer shares this block
    const int lane = tid & 31, warp = tid >> 5;
    for (int32_t r = r0 + warp; r < r1; r += KLD_THREADS / 32) {
      const float v = warp_row_value(col, w, state, rowptr[r], rowptr[r + 1], -1, -1, lane);
      if (lane == 0) val[r] = v;
    }
  }
}

// After a pass, state[] holds lock bits and the swapped sides while order0 / order1 / rank / n0 / n1 still
// describe the partition the pass started from.  The reference's KL() rebuilds remain[] / split[] from the current
// sides on every call (cKL.cpp:290-301), so the next pass / cut / D-value request does the same here: the final
// sides become a fresh partition with ascending remain[] lists and no locks.
__global__ void mask_side_kernel(const uint8_t *__restrict__ state, int32_t n, uint8_t *__restrict__ out) {
  int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) out[v] = state[v] & ST_SIDE;
}
void kl_set_partition_device(eigkl_handle *h, const uint8_t *side_dev);
static void kl_rearm(eigkl_handle *h) {
  auto &k = h->kl;
  if (!k.have_partition || !k.consumed) return;
  const int32_t n = h->hg.n_nodes;
  auto &side = h->scr.u8a; side.alloc((size_t)n);
  mask_side_kernel<<<grid_for(n), TPB, 0, h->stream>>>(k.state.p, n, side.p);
  h->launches++;
  kl_set_partition_device(h, side.p);
}

// Second form (default).  The first one gathers one side BYTE per entry -- a 32-byte sector of L2 traffic per
// useful byte, 352 MB on top of the 88 MB of (col, w) at 2 M nodes -- after two dependent loads just to find its
// row block (0.22 of the HBM peak there, 0.11 on ibm10).  Here:
//   * the sides are a BITMAP (1 bit per node, packed when the partition is set: 252 KB at 2 M nodes), read
//     through the read-only path: after the first touches an SM's L1 holds it, so the gathers stop reaching L2;
//   * a CTA reads ONE int4 descriptor (rows and entry range), then issues all its loads in one round -- 8
//     (col, w) pairs per thread, the row pointers into shared memory -- then the bit gathers, stages the signed
//     weights, and one thread per row adds its run in order (two fp32 accumulators, cKL.cpp:225-251).
constexpr int KLD_K = 8;
__global__ void __launch_bounds__(KLD_THREADS)
dvalues_bits_kernel(const int4 *__restrict__ info, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                    const float *__restrict__ w, const uint32_t *__restrict__ side_bits, const uint8_t *__restrict__ state,
                    float *__restrict__ val) {
  __shared__ float sv[KLD_THREADS * KLD_K];
  __shared__ int32_t rp[KLD_THREADS * KLD_K + 1];
  const int tid = threadIdx.x;
  const int4 bi = __ldg(info + blockIdx.x);
  const int32_t r0 = bi.x, r1 = bi.y, e0 = bi.z, e1 = bi.w;
  if (r0 >= r1) return;
  const int32_t span = e1 - e0, nrows = r1 - r0;
  if (span <= KLD_THREADS * KLD_K) {
    int32_t c[KLD_K];
    float ww[KLD_K];
#pragma unroll
    for (int k = 0; k < KLD_K; ++k) {
      const int32_t i = tid + k * KLD_THREADS;
      c[k] = -1; ww[k] = 0.0f;
      if (i < span) { c[k] = __ldcs(col + e0 + i); ww[k] = __ldcs(w + e0 + i); }
    }
    for (int32_t rr = tid; rr <= nrows; rr += KLD_THREADS) rp[rr] = __ldg(rowptr + r0 + rr) - e0;
#pragma unroll
    for (int k = 0; k < KLD_K; ++k) {
      const int32_t i = tid + k * KLD_THREADS;
      if (c[k] >= 0) sv[i] = ((__ldg(side_bits + (c[k] >> 5)) >> (c[k] & 31)) & 1u) ? ww[k] : -ww[k];
    }
    __syncthreads();
    for (int32_t rr = tid; rr < nrows; rr += KLD_THREADS) {
      const int32_t lo = rp[rr], hi = rp[rr + 1];
      float E = 0.0f, I = 0.0f;
      for (int32_t i = lo; i < hi; ++i) {
        const float x = sv[i];
        E = __fadd_rn(E, fmaxf(x, 0.0f));
        I = __fadd_rn(I, fmaxf(-x, 0.0f));
      }
      val[r0 + rr] = __fsub_rn(E, I);
    }
  } else {                              // a row longer than the staging buffer lives here
    const int lane = tid & 31, warp = tid >> 5;
    for (int32_t r = r0 + warp; r < r1; r += KLD_THREADS / 32) {
      const float v = warp_row_value(col, w, state, rowptr[r], rowptr[r + 1], -1, -1, lane);
      if (lane == 0) val[r] = v;
    }
  }
}
__global__ void kl_blk_info_kernel(const int32_t *__restrict__ blk_row, const int32_t *__restrict__ rowptr, int32_t n_blocks,
                                   int4 *__restrict__ info) {
  const int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const int32_t r0 = blk_row[b], r1 = blk_row[b + 1];
  info[b] = make_int4(r0, r1, rowptr[r0], rowptr[r1]);
}
// 32 nodes per warp: the side bits of the state bytes, one word per ballot
__global__ void pack_side_bits_kernel(const uint8_t *__restrict__ state, int32_t n, uint32_t *__restrict__ out) {
  const int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned bit = (v < n) ? ((unsigned)state[v] & ST_SIDE) : 0u;
  const unsigned word = __ballot_sync(FULL_MASK, bit != 0u);
  if ((threadIdx.x & 31) == 0 && (v >> 5) <= ((n - 1) >> 5)) out[v >> 5] = word;
}
static void kl_pack_sides(eigkl_handle *h) {
  auto &k = h->kl;
  const int32_t n = h->hg.n_nodes;
  k.side_bits.ensure((size_t)(n + 31) / 32 + 1);
  pack_side_bits_kernel<<<grid_for((int64_t)((n + 31) / 32) * 32), TPB, 0, h->stream>>>(k.state.p, n, k.side_bits.p);
  h->launches++;
}

void kl_dvalues(eigkl_handle *h) {
  auto &A = h->A;
  auto &k = h->kl;
  EIGKL_REQUIRE(A.valid, EIGKL_E_ARG, "KL graph not assembled");
  EIGKL_REQUIRE(k.have_partition, EIGKL_E_ARG, "no partition set");
  kl_rearm(h);
  static const bool old_form = getenv("EIGKL_DVALUES_BYTES") != nullptr;     // tuning aid: the byte-gather kernel
  if (!old_form && !A.info_valid) {
    A.blk_info.alloc((size_t)4 * A.n_blocks + 4);
    kl_blk_info_kernel<<<grid_for(A.n_blocks), TPB, 0, h->stream>>>(A.blk_row.p, A.rowptr.p, A.n_blocks, reinterpret_cast<int4 *>(A.blk_info.p));
    h->launches++;
    A.info_valid = true;
  }
  h->prof.begin(KC_DVALUES, h->stream);
  if (old_form)
    dvalues_kernel<<<(unsigned)A.n_blocks, KLD_THREADS, 0, h->stream>>>(A.rowptr.p, A.col.p, A.w.p, k.state.p, k.val.p, A.blk_row.p);
  else
    dvalues_bits_kernel<<<(unsigned)A.n_blocks, KLD_THREADS, 0, h->stream>>>(reinterpret_cast<const int4 *>(A.blk_info.p), A.rowptr.p, A.col.p,
                                                                          A.w.p, k.side_bits.p, k.state.p, k.val.p);
  h->prof.end(h->stream);
  h->launches++;
  EIGKL_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------
// partition set-up
// ---------------------------------------------------------------------------------------------------
__global__ void side_flag_kernel(const uint8_t *__restrict__ side, int32_t n, int32_t *__restrict__ is0, int *__restrict__ err) {
  int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) {
    const uint8_t s = side[v];
    if (s > 1) atomicOr(err, 1);
    is0[v] = (s == 0) ? 1 : 0;
  }
}
__global__ void build_orders_kernel(const uint8_t *__restrict__ side, const int32_t *__restrict__ pos0, int32_t n,
                                    int32_t *__restrict__ order0, int32_t *__restrict__ order1, uint8_t *__restrict__ state,
                                    uint32_t *__restrict__ rank) {
  int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int32_t p0 = pos0[v];
  if (side[v] == 0) { order0[p0] = v; rank[v] = (uint32_t)p0; state[v] = 0; }
  else              { order1[v - p0] = v; rank[v] = (uint32_t)(v - p0); state[v] = ST_SIDE; }
}
__global__ void apply_order_kernel(const int32_t *__restrict__ order, int64_t cnt, uint8_t s, int32_t n, uint8_t *__restrict__ state,
                                   uint32_t *__restrict__ rank, int32_t *__restrict__ seen, int *__restrict__ err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  const int32_t v = order[i];
  if (v < 0 || v >= n) { atomicOr(err, 1); return; }
  if (atomicAdd(&seen[v], 1) != 0) atomicOr(err, 2);
  state[v] = s;
  rank[v] = (uint32_t)i;
}

static void kl_alloc_state(eigkl_handle *h, int32_t n) {
  auto &k = h->kl;
  k.state.ensure((size_t)n); k.rank.ensure((size_t)n); k.val.ensure((size_t)n);
  k.order0.ensure((size_t)n); k.order1.ensure((size_t)n);
  const int64_t n_tiles = ceil_div(n, KL_TILE);
  k.tile_key.ensure((size_t)(2 * n_tiles + 2 * KL_MAX_CLUSTER + 8));
  k.tile_stamp.ensure((size_t)n_tiles + 1);
  k.ctrl.ensure(64);
}

void kl_set_partition_device(eigkl_handle *h, const uint8_t *side_dev) {
  const int32_t n = h->hg.n_nodes;
  auto &k = h->kl;
  kl_alloc_state(h, n);
  auto &pos = h->scr.i32a; pos.alloc((size_t)n + 1);
  auto &err = h->scr.err; err.alloc(1);
  EIGKL_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), h->stream));
  side_flag_kernel<<<grid_for(n), TPB, 0, h->stream>>>(side_dev, n, pos.p, err.p);
  exclusive_scan_i32(h, pos.p, pos.p, n);
  build_orders_kernel<<<grid_for(n), TPB, 0, h->stream>>>(side_dev, pos.p, n, k.order0.p, k.order1.p, k.state.p, k.rank.p);
  h->launches += 2;
  int32_t n0 = 0; int herr = 0;
  EIGKL_CUDA(cudaMemcpyAsync(&n0, pos.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  EIGKL_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  EIGKL_CUDA(cudaStreamSynchronize(h->stream));
  EIGKL_REQUIRE(herr == 0, EIGKL_E_FORMAT, "partition side not in {0,1}");
  k.n0 = n0; k.n1 = n - n0; k.ascending = true; k.have_partition = true; k.consumed = false;
  kl_pack_sides(h);
}

void kl_set_partition(eigkl_handle *h, const uint8_t *side_host, const int32_t *order0, int64_t n0,
                      const int32_t *order1, int64_t n1, bool ascending) {
  EIGKL_REQUIRE(h->hg.loaded, EIGKL_E_ARG, "no hypergraph loaded");
  const int32_t n = h->hg.n_nodes;
  auto &k = h->kl;
  if (ascending) {
    EIGKL_REQUIRE(side_host != nullptr, EIGKL_E_ARG, "side is NULL");
    auto &side = h->scr.u8a; side.alloc((size_t)n);
    EIGKL_CUDA(cudaMemcpyAsync(side.p, side_host, (size_t)n, cudaMemcpyHostToDevice, h->stream));
    kl_set_partition_device(h, side.p);
    return;
  }
  EIGKL_REQUIRE(order0 && order1 && n0 >= 0 && n1 >= 0 && n0 + n1 == n, EIGKL_E_ARG, "orders must cover every node once");
  kl_alloc_state(h, n);
  auto &seen = h->scr.i32a; seen.alloc((size_t)n);
  auto &err = h->scr.err; err.alloc(1);
  EIGKL_CUDA(cudaMemsetAsync(seen.p, 0, (size_t)n * sizeof(int32_t), h->stream));
  EIGKL_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), h->stream));
  EIGKL_CUDA(cudaMemcpyAsync(k.order0.p, order0, (size_t)n0 * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  EIGKL_CUDA(cudaMemcpyAsync(k.order1.p, order1, (size_t)n1 * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  if (n0) apply_order_kernel<<<grid_for(n0), TPB, 0, h->stream>>>(k.order0.p, n0, 0, n, k.state.p, k.rank.p, seen.p, err.p);
  if (n1) apply_order_kernel<<<grid_for(n1), TPB, 0, h->stream>>>(k.order1.p, n1, ST_SIDE, n, k.state.p, k.rank.p, seen.p, err.p);
  h->launches += 2;
  int herr = 0;
  EIGKL_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  EIGKL_CUDA(cudaStreamSynchronize(h->stream));
  EIGKL_REQUIRE(herr == 0, EIGKL_E_ARG, "orders must cover every node exactly once");
  k.n0 = n0; k.n1 = n1; k.ascending = false; k.have_partition = true; k.consumed = false;
  kl_pack_sides(h);
}

// ---------------------------------------------------------------------------------------------------
// initial cut in the reference's one-thread order                         cKL.cpp:199-223
// ---------------------------------------------------------------------------------------------------
__global__ void set_first_kernel(const int32_t *__restrict__ seq, int32_t L, uint32_t B, uint32_t *__restrict__ first) {
  int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < L) atomicMin(&first[(uint32_t)seq[t] % B], (uint32_t)t);
}
__global__ void set_keys_kernel(const int32_t *__restrict__ seq, int32_t L, uint32_t B, const uint32_t *__restrict__ first,
                                unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals) {
  int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < L) {
    // reversed sequence, stable sort by descending first-hit time of the bucket (see stl_order.h)
    const uint32_t f = first[(uint32_t)seq[t] % B];
    keys[L - 1 - t] = (unsigned long long)((uint32_t)(L - 1) - f);
    vals[L - 1 - t] = (uint32_t)seq[t];
  }
}
__global__ void set_rank_kernel(const uint32_t *__restrict__ ordered, int32_t L, uint32_t *__restrict__ set_rank) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L) set_rank[ordered[i]] = (uint32_t)i;
}
// cut entries of the left rows: key = position of the row in remain[0] : forward/backward : order inside
__global__ void cut_entries_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ fwd_end,
                                   const int32_t *__restrict__ col, const uint8_t *__restrict__ state,
                                   const uint32_t *__restrict__ rank, const uint32_t *__restrict__ set_rank, int32_t n,
                                   int qb, unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals,
                                   unsigned int *__restrict__ count) {
  // one warp per row
  const int lane = threadIdx.x & 31;
  const int32_t v = (int32_t)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (v >= n) return;
  if (state[v] & ST_SIDE) return;                                   // only rows of remain[0]
  const int32_t lo = rowptr[v], fe = fwd_end[v], hi = rowptr[v + 1];
  const unsigned long long rowkey = (unsigned long long)rank[v] << (qb + 1);
  for (int32_t i = lo + lane; i < hi; i += 32) {
    const int32_t c = col[i];
    if (!(state[c] & ST_SIDE)) continue;                            // rightNodes.count(neighbor)
    unsigned long long key;
    if (i < fe) key = rowkey | (unsigned long long)(uint32_t)(i - lo);                       // map order
    else        key = rowkey | (1ull << qb) | (unsigned long long)set_rank[c];               // set order
    const unsigned int slot = atomicAdd(count, 1u);
    keys[slot] = key;
    vals[slot] = (uint32_t)i;
  }
}
// strictly sequential fp32 sum of w[vals[i]], i ascending (the order calCutSize adds the cut weights in,
// cKL.cpp:199-223).  Warp 0 adds one 4096-value chunk from shared memory, 4-5 cycles per value (shuffle +
// dependent FADD), while the other 31 warps gather the next chunk (two dependent loads per value) into the
// second buffer: the sum runs at the FADD chain's pace instead of two L2 round trips per 32 values
// (ibm10: 568 -> ~40 us, the largest item of the KL set-up).
constexpr int OS_THREADS = 1024;
constexpr int OS_CHUNK = 4096;
__global__ void __launch_bounds__(OS_THREADS)
ordered_sum_kernel(const uint32_t *__restrict__ vals, const float *__restrict__ w, int64_t m, float *__restrict__ out) {
  __shared__ float buf[2][OS_CHUNK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t i = tid; i < min((int64_t)OS_CHUNK, m); i += OS_THREADS) buf[0][i] = w[vals[i]];
  __syncthreads();
  float s = 0.0f;
  int c = 0;
  for (int64_t base = 0; base < m; base += OS_CHUNK, c ^= 1) {
    if (warp == 0) {
      const int cnt = (int)min((int64_t)OS_CHUNK, m - base);
      for (int j = 0; j < cnt; j += 32) {
        const float x = (j + lane < cnt) ? buf[c][j + lane] : 0.0f;
        const int k = min(32, cnt - j);
        for (int t = 0; t < k; ++t) s = __fadd_rn(s, __shfl_sync(FULL_MASK, x, t));
      }
    } else {
      const int64_t nb = base + OS_CHUNK;
      const int64_t cnt = min((int64_t)OS_CHUNK, m - nb);
      for (int64_t i = tid - 32; i < cnt; i += OS_THREADS - 32) buf[c ^ 1][i] = w[vals[nb + i]];
    }
    __syncthreads();
  }
  if (tid == 0) out[0] = s;
}

float kl_cut0(eigkl_handle *h) {
  auto &A = h->A;
  auto &k = h->kl;
  auto &e = h->eig;
  EIGKL_REQUIRE(A.valid, EIGKL_E_ARG, "KL graph not assembled");
  EIGKL_REQUIRE(k.have_partition, EIGKL_E_ARG, "no partition set");
  kl_rearm(h);
  const int32_t n = A.n;
  const int32_t n1 = (int32_t)k.n1;
  cudaStream_t st = h->stream;
  const size_t need = (size_t)std::max<int64_t>(std::max<int64_t>(A.nnz, n), 16) + 1;
  for (int i = 0; i < 2; ++i) { e.sortkey[i].ensure(need); e.sortval[i].ensure(need); }
  unsigned long long *keys[2] = {e.sortkey[0].p, e.sortkey[1].p};
  uint32_t *vals[2] = {e.sortval[0].p, e.sortval[1].p};
  auto &set_rank = h->scr.u32a; set_rank.alloc((size_t)n);
  auto &seq = h->scr.i32a; seq.alloc((size_t)std::max(n1, 1));
  // (1) iteration order of unordered_set<uint32_t>(remain[1].begin(), remain[1].end())
  if (n1 > 0) {
    const uint32_t Bfinal = stl_final_buckets((uint32_t)n1);
    auto &first = h->scr.u32b; first.alloc(Bfinal);
    int32_t done = 0;                                   // elements already in seq
    for (int lv = 0; lv < STL_CHAIN_LEN; ++lv) {
      const uint32_t B = stl_bucket_chain(lv);
      const int32_t L = (int32_t)std::min<int64_t>(n1, B);
      // append remain[1][done .. L)
      EIGKL_CUDA(cudaMemcpyAsync(seq.p + done, k.order1.p + done, (size_t)(L - done) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
      EIGKL_CUDA(cudaMemsetAsync(first.p, 0xFF, (size_t)B * sizeof(uint32_t), st));
      set_first_kernel<<<grid_for(L), TPB, 0, st>>>(seq.p, L, B, first.p);
      set_keys_kernel<<<grid_for(L), TPB, 0, st>>>(seq.p, L, B, first.p, keys[0], vals[0]);
      h->launches += 2;
      const int cur = radix_sort_kv(h, keys, vals, L, bits_for((uint64_t)std::max(L - 1, 1)));
      EIGKL_CUDA(cudaMemcpyAsync(seq.p, vals[cur], (size_t)L * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
      done = L;
      if ((int64_t)B >= n1) break;
    }
    set_rank_kernel<<<grid_for(n1), TPB, 0, st>>>(reinterpret_cast<const uint32_t *>(seq.p), n1, set_rank.p);
    h->launches++;
  }
  // (2) cut entries keyed by evaluation order, (3) sort, (4) ordered sum
  EIGKL_CUDA(cudaMemsetAsync(k.ctrl.p + 4, 0, sizeof(int64_t), st));
  unsigned int *count = reinterpret_cast<unsigned int *>(k.ctrl.p + 4);
  const int qb = bits_for((uint64_t)std::max(n - 1, 1));
  cut_entries_kernel<<<grid_for((int64_t)n * 32), TPB, 0, st>>>(A.rowptr.p, A.fwd_end.p, A.col.p, k.state.p, k.rank.p, set_rank.p, n, qb,
                                                              keys[0], vals[0], count);
  h->launches++;
  unsigned int m = 0;
  EIGKL_CUDA(cudaMemcpyAsync(&m, count, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  EIGKL_CUDA(cudaStreamSynchronize(st));
  const int kb = bits_for((uint64_t)std::max<int64_t>(k.n0 - 1, 1)) + 1 + qb;
  const int cur = radix_sort_kv(h, keys, vals, m, kb);
  float *out = reinterpret_cast<float *>(k.ctrl.p + 5);
  ordered_sum_kernel<<<1, OS_THREADS, 0, st>>>(vals[cur], A.w.p, (int64_t)m, out);
  h->launches++;
  float cut = 0.0f;
  EIGKL_CUDA(cudaMemcpyAsync(&cut, out, sizeof(float), cudaMemcpyDeviceToHost, st));
  EIGKL_CUDA(cudaStreamSynchronize(st));
  EIGKL_CUDA(cudaGetLastError());
  return cut;
}

// ---------------------------------------------------------------------------------------------------
// tile keys
// ---------------------------------------------------------------------------------------------------
// [lo, hi): the nodes this rank owns (everything on one GPU); other nodes of the tile are ignored.
// All 24 loads of a lane are issued before the first use: one L2 round trip per rescan.
__device__ __forceinline__ void tile_scan(const uint8_t *state, const float *val, const uint32_t *__restrict__ rank, int32_t lo,
                                          int32_t hi, int32_t tile, int lane, unsigned long long &k0, unsigned long long &k1) {
  k0 = 0ull; k1 = 0ull;
  const int32_t base = tile * KL_TILE;
  unsigned st[KL_TILE / 32];
  float vv[KL_TILE / 32];
  uint32_t rk[KL_TILE / 32];
#pragma unroll
  for (int r = 0; r < KL_TILE / 32; ++r) {
    const int32_t u = base + r * 32 + lane;
    const bool in = (u >= lo && u < hi);
    st[r] = in ? (unsigned)__ldcg(state + u) : ST_LOCK;
    vv[r] = in ? __ldcg(val + u) : 0.0f;
    rk[r] = in ? __ldg(rank + u) : 0u;
  }
#pragma unroll
  for (int r = 0; r < KL_TILE / 32; ++r) {
    if (!(st[r] & ST_LOCK)) {
      const unsigned long long low = (unsigned long long)(0xFFFFFFFFu - rk[r]);
      if (st[r] & ST_SIDE) { const unsigned long long key = ((unsigned long long)float_orderable(-vv[r]) << 32) | low; k1 = key > k1 ? key : k1; }
      else                 { const unsigned long long key = ((unsigned long long)float_orderable(vv[r]) << 32) | low;  k0 = key > k0 ? key : k0; }
    }
  }
  k0 = warp_max_u64(k0);
  k1 = warp_max_u64(k1);
}
__global__ void tile_init_kernel(const uint8_t *__restrict__ state, const float *__restrict__ val, const uint32_t *__restrict__ rank,
                                 int32_t own_lo, int32_t own_hi, int32_t n_tiles, unsigned long long *__restrict__ tile_key,
                                 uint32_t *__restrict__ tile_stamp) {
  const int lane = threadIdx.x & 31;
  const int32_t tile = (int32_t)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (tile >= n_tiles) return;
  unsigned long long k0, k1;
  tile_scan(state, val, rank, own_lo, own_hi, tile, lane, k0, k1);
  if (lane == 0) { tile_key[2 * tile] = k0; tile_key[2 * tile + 1] = k1; tile_stamp[tile] = 0u; }
}

__device__ __forceinline__ float float_from_orderable(uint32_t u) {
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(u);
}

// ---------------------------------------------------------------------------------------------------
// the persistent swap loop
// ---------------------------------------------------------------------------------------------------
struct KlLoopParams {
  int32_t n, n_tiles;
  const int32_t *rowptr, *col;
  const float *w;
  uint8_t *state;
  const uint32_t *rank;
  float *val;
  unsigned long long *tile_key;     // 2*n_tiles, then the exchange area 2*KL_MAX_CLUSTER
  uint32_t *tile_stamp;
  const int32_t *order0, *order1;
  float *t_cut, *t_gain;
  int32_t *t_n1, *t_n2;
  int64_t *ctrl;
  float cut0;
  uint32_t term_limit;
  int64_t n0, n1;
  int32_t own_lo, own_hi;           // nodes whose D-values / tile keys this rank maintains
  struct KlCtrl *mctrl;             // multi-rank: loop state carried between the per-swap launches
};

struct KlCtrl {                     // multi-rank loop state (device memory)
  float cut;
  uint32_t term, iter;
  int32_t done;
  long long rem0, rem1;
};

// multi-rank S1: best pair over THIS rank's tiles -> exch[0..1] (then ncclAllReduce(max) across ranks: C3)
__global__ void __launch_bounds__(KL_LOOP_THREADS)
kl_select_kernel(const unsigned long long *__restrict__ tile_key, int32_t tile_lo, int32_t tile_hi,
                 unsigned long long *__restrict__ exch, const KlCtrl *__restrict__ mctrl) {
  __shared__ unsigned long long red0[32], red1[32];
  if (mctrl->done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long k0 = 0ull, k1 = 0ull;
  for (int32_t t = tile_lo + tid; t < tile_hi; t += KL_LOOP_THREADS) {
    const unsigned long long a0 = __ldcg(tile_key + 2 * (size_t)t), a1 = __ldcg(tile_key + 2 * (size_t)t + 1);
    k0 = a0 > k0 ? a0 : k0;
    k1 = a1 > k1 ? a1 : k1;
  }
  k0 = warp_max_u64(k0);
  k1 = warp_max_u64(k1);
  if (lane == 0) { red0[warp] = k0; red1[warp] = k1; }
  __syncthreads();
  if (warp == 0) {
    k0 = warp_max_u64(red0[lane]);
    k1 = warp_max_u64(red1[lane]);
    if (lane == 0) { exch[0] = k0; exch[1] = k1; }
  }
}

// MULTI = false: the whole pass in one launch.  MULTI = true: ONE swap per launch -- the pair comes from
// exch[] (all-reduced over the ranks), the loop state from *p.mctrl, and only owned nodes are updated.
template <bool MULTI>
__global__ void __launch_bounds__(KL_LOOP_THREADS, 1) kl_loop_kernel(const KlLoopParams p) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned nc = cluster.num_blocks();
  const unsigned cr = cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned gwarp = cr * (KL_LOOP_THREADS / 32) + warp;
  const unsigned total_warps = nc * (KL_LOOP_THREADS / 32);
  const unsigned gtid = cr * KL_LOOP_THREADS + tid;
  const unsigned total_threads = nc * KL_LOOP_THREADS;
  unsigned long long *exch = p.tile_key + 2 * (size_t)p.n_tiles;

  __shared__ unsigned long long red0[32], red1[32];
  __shared__ unsigned long long sh_best[2];
  __shared__ float sh_cut;
  __shared__ uint32_t sh_term, sh_iter;
  __shared__ int sh_done;
  __shared__ long long sh_rem0, sh_rem1;
  if (tid == 0) {
    if (MULTI) {
      sh_cut = p.mctrl->cut; sh_term = p.mctrl->term; sh_iter = p.mctrl->iter; sh_done = p.mctrl->done;
      sh_rem0 = p.mctrl->rem0; sh_rem1 = p.mctrl->rem1;
    } else {
      sh_cut = p.cut0; sh_term = 0; sh_iter = 0; sh_done = 0; sh_rem0 = p.n0; sh_rem1 = p.n1;
      if (p.n0 <= 0 || p.n1 <= 0) sh_done = 1;
    }
  }
  __syncthreads();
  uint32_t it_local = sh_iter;
  // phase clocks of the swap loop (thread 0 only; a handful of clock reads per swap): published in ctrl[8..13]
  long long tph[6] = {0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
#define KL_PHASE(i) do { if (tid == 0) { const long long t_ = clock64(); tph[i] += t_ - tprev; tprev = t_; } } while (0)

  while (!sh_done) {
    // ---- S1: best pair over the cached tile keys ------------------------------------------------
    unsigned long long k0 = 0ull, k1 = 0ull;
    if (MULTI) {
      if (tid == 0) { sh_best[0] = __ldcg(exch); sh_best[1] = __ldcg(exch + 1); }
    } else {
    for (uint32_t t = gtid; t < (uint32_t)p.n_tiles; t += total_threads) {
      const unsigned long long a0 = __ldcg(p.tile_key + 2 * (size_t)t), a1 = __ldcg(p.tile_key + 2 * (size_t)t + 1);
      k0 = a0 > k0 ? a0 : k0;
      k1 = a1 > k1 ? a1 : k1;
    }
    k0 = warp_max_u64(k0);
    k1 = warp_max_u64(k1);
    if (lane == 0) { red0[warp] = k0; red1[warp] = k1; }
    __syncthreads();
    if (warp == 0) {
      k0 = warp_max_u64(red0[lane]);
      k1 = warp_max_u64(red1[lane]);
      if (lane == 0) {
        if (nc > 1) { __stcg(exch + 2 * cr, k0); __stcg(exch + 2 * cr + 1, k1); }
        else { sh_best[0] = k0; sh_best[1] = k1; }
      }
    }
    if (nc > 1) {
      cluster.sync();
      if (warp == 0) {
        k0 = lane < (int)nc ? __ldcg(exch + 2 * lane) : 0ull;
        k1 = lane < (int)nc ? __ldcg(exch + 2 * lane + 1) : 0ull;
        k0 = warp_max_u64(k0);
        k1 = warp_max_u64(k1);
        if (lane == 0) { sh_best[0] = k0; sh_best[1] = k1; }
      }
    }
    }   // !MULTI
    __syncthreads();
    KL_PHASE(0);                                     // S1: tile-key reduction
    const unsigned long long b0 = sh_best[0], b1 = sh_best[1];
    if (b0 == 0ull || b1 == 0ull) {                  // no selectable node on one side (cKL.cpp:387-389)
      if (tid == 0) sh_done = 1;
      __syncthreads();
      break;
    }
    ++it_local;                                      // swap number, kept in a register by every thread
    const int32_t a = __ldg(p.order0 + (0xFFFFFFFFu - (uint32_t)(b0 & 0xFFFFFFFFull)));
    const int32_t b = __ldg(p.order1 + (0xFFFFFFFFu - (uint32_t)(b1 & 0xFFFFFFFFull)));
    const int32_t alo = __ldg(p.rowptr + a), ahi = __ldg(p.rowptr + a + 1);
    const int32_t blo = __ldg(p.rowptr + b), bhi = __ldg(p.rowptr + b + 1);
    const int32_t da = ahi - alo, items = da + (bhi - blo);
    KL_PHASE(1);                                     // decode (a, b), their row pointers
    constexpr int WORKERS = KL_LOOP_THREADS / 32 - 1;    // warp 31 of every CTA is the bookkeeper
    const unsigned gworker = cr * WORKERS + warp, total_workers = nc * WORKERS;
    int32_t vkeep[4];
    if (warp == WORKERS) {
      // ---- S2 (off the critical path): gain, cut, trace, termination.  Every CTA's bookkeeper computes
      //      the same values; CTA 0's also publishes them.
      float wab = 0.0f;
      for (int32_t i = alo + lane; i < ahi; i += 32)
        if (__ldg(p.col + i) == b) wab = __ldg(p.w + i);             // getEdgeWeight, cKL.cpp:75-82
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wab = fmaxf(wab, __shfl_xor_sync(FULL_MASK, wab, o));   // weights are > 0
      if (lane == 0) {
        const float maxGain = float_from_orderable((uint32_t)(b0 >> 32));
        const float minGain = __fsub_rn(0.0f, float_from_orderable((uint32_t)(b1 >> 32)));
        const float gain = __fsub_rn(__fsub_rn(maxGain, minGain), __fmul_rn(2.0f, wab));    // cKL.cpp:360
        const float cut = __fsub_rn(sh_cut, gain);                                          // cKL.cpp:362
        sh_cut = cut;
        sh_iter = it_local;
        if (cr == 0) {
          p.t_cut[it_local] = cut; p.t_gain[it_local] = gain; p.t_n1[it_local] = a; p.t_n2[it_local] = b;
          __stcg(p.state + a, (uint8_t)(ST_SIDE | ST_LOCK));        // swip, cKL.cpp:274-286
          p.state[b] = (uint8_t)(ST_LOCK);
        }
        if (gain <= 0.0f) { if (++sh_term > p.term_limit) sh_done = 1; }   // cKL.cpp:382-386
        else sh_term = 0;
        if (--sh_rem0 == 0) sh_done = 1;
        if (--sh_rem1 == 0) sh_done = 1;
      }
    } else {
      // ---- S3: recompute D of every neighbour of a or b from scratch (cKL.cpp:253-272) ---------------
      int kept = 0;
      for (int32_t it = gworker; it < items; it += total_workers, ++kept) {
        const int32_t v = __ldg(p.col + (it < da ? alo + it : blo + (it - da)));
        if (kept < 4) vkeep[kept] = v;
        if (MULTI && (v < p.own_lo || v >= p.own_hi)) continue;     // another rank owns this D-value
        const float nv = warp_row_value(p.col, p.w, p.state, __ldg(p.rowptr + v), __ldg(p.rowptr + v + 1), a, b, lane);
        if (lane == 0) __stcg(p.val + v, nv);
      }
    }
    KL_PHASE(2);                                     // S3 as seen by warp 0 (its own rows)
    if (nc == 1) __syncthreads(); else cluster.sync();   // one CTA: a block barrier orders the .cg accesses
    KL_PHASE(3);                                     // wait for the slowest S3 warp / the bookkeeper
    // ---- S4: rescan the tiles that contain a touched node (the bookkeeper takes the tiles of a and b) ----
    const uint32_t stamp = it_local;
    {
      // Every warp first claims ALL its tiles at once (lane k stamps the tile of the warp's k-th node: one
      // atomic round trip for the whole batch instead of one per node), then rescans the claimed ones.
      // The bookkeeper's batch is {a, b}; a worker's batch is the nodes it recomputed in S3.
      const bool keeper = (warp == WORKERS);
      const int32_t n_mine = keeper ? ((cr == 0) ? 2 : 0)
                                    : (items > (int32_t)gworker ? (items - 1 - (int32_t)gworker) / (int32_t)total_workers + 1 : 0);
      for (int32_t b0i = 0; b0i < n_mine; b0i += 32) {
        const int32_t kidx = b0i + lane;
        int32_t tile = -1;
        if (kidx < n_mine) {
          int32_t v;
          if (keeper) v = kidx ? b : a;
          else {
            const int32_t it = (int32_t)gworker + kidx * (int32_t)total_workers;
            v = (kidx < 4) ? vkeep[kidx] : __ldg(p.col + (it < da ? alo + it : blo + (it - da)));
          }
          if (!(MULTI && (v < p.own_lo || v >= p.own_hi))) tile = v / KL_TILE;
        }
        bool claimed = false;
        if (tile >= 0) claimed = atomicExch(p.tile_stamp + tile, stamp) != stamp;
        unsigned todo = __ballot_sync(FULL_MASK, claimed);
        while (todo) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const int32_t t = __shfl_sync(FULL_MASK, tile, src);
          unsigned long long t0, t1;
          tile_scan(p.state, p.val, p.rank, p.own_lo, p.own_hi, t, lane, t0, t1);
          if (lane == 0) { __stcg(p.tile_key + 2 * (size_t)t, t0); __stcg(p.tile_key + 2 * (size_t)t + 1, t1); }
        }
      }
    }
    KL_PHASE(4);                                     // S4 as seen by warp 0
    if (nc == 1) __syncthreads(); else cluster.sync();
    KL_PHASE(5);                                     // wait for the slowest S4 warp
    if (MULTI) break;                                  // one swap per launch
  }
  if (cr == 0 && tid == 0) {
    if (MULTI) {
      p.mctrl->cut = sh_cut; p.mctrl->term = sh_term; p.mctrl->iter = sh_iter; p.mctrl->done = sh_done;
      p.mctrl->rem0 = sh_rem0; p.mctrl->rem1 = sh_rem1;
    }
    p.ctrl[0] = (int64_t)sh_iter; p.ctrl[1] = 1;
    if (!MULTI)
      for (int i = 0; i < 6; ++i) p.ctrl[8 + i] = tph[i];
  }
#undef KL_PHASE
}


// ---------------------------------------------------------------------------------------------------
// The swap loop with its working state in shared memory (single rank, n <= KL_LOCAL_MAX_N).
//
// kl_loop_kernel above is a chain of ~10 dependent L2 round trips per swap (tile keys -> order[] -> row
// pointers -> neighbour id -> its row pointers -> its entries -> their side bytes -> D store | stamp
// atomics -> tile rescan loads -> key stores) at ~1 us each from one SM: 8.8 us per swap on ibm10
// whatever the bandwidth.  One CTA can hold what the chain keeps re-fetching:
//   * the tile keys and the rescan stamps / work list (20 bytes per 256 nodes),
//   * the side / locked bits of EVERY node, 2 bits each (ibm10: 17 KB; 128 KB at the 524 288-node limit),
// an auxiliary array gives, for every CSR entry, the row extent of the neighbour it names (one load yields
// v, rowptr[v], rowptr[v+1]), a warp issues the loads of ALL its neighbour rows before it sums the first,
// and a recomputed D-value updates its tile's key with one shared-memory atomicMax -- a tile is rescanned
// only when the node that held its best key was itself touched or locked (2-4 tiles per swap, spread over
// the warps, instead of every touched tile).  Per swap: row pointers of (a, b) -> neighbour entries -> the
// neighbours' rows | D-values of the few tiles to rescan: four round trips.
// When the remain[] orders are ascending node ids (the -EIG start, cKL.cpp:155-174) the tie-break field of a
// key is the node id itself, so the selected key names the node without the order[] / rank[] loads.
// The arithmetic (order of the fp32 additions, first-max selection, gain, termination) is unchanged:
// traces stay byte-identical to cKL.cpp's.
// ---------------------------------------------------------------------------------------------------
constexpr int32_t KL_LOCAL_MAX_N = 524288;            // tile keys + 2 bits per node in shared memory
constexpr int32_t KL_LOCAL_GBITS_MAX_N = 2097152;     // tile keys in shared memory, state bytes in global memory

__device__ __forceinline__ unsigned bits_get(const uint32_t *bits, int32_t u) { return (bits[u >> 4] >> ((u & 15) * 2)) & 3u; }
// GBITS (graphs whose 2 bits per node do not fit beside the tile keys, up to KL_LOCAL_GBITS_MAX_N nodes): the
// side / locked state is read from the global state bytes instead -- one more dependent load in the row sums
template <bool GBITS>
__device__ __forceinline__ unsigned kl_state_get(const uint32_t *bits, const uint8_t *state, int32_t u) {
  return GBITS ? ((unsigned)__ldcg(state + u) & 3u) : bits_get(bits, u);
}

// as warp_row_value, sides from the shared-memory bits; (c, ww) = this lane's entry of the row's first 32, already
// loaded.  The two ordered sums are independent chains -- E only ever grows by the external weights, I by the
// internal ones (the other accumulator gets + 0.0f, which changes nothing) -- so the weights are compacted in row
// order into two per-warp lists (ballot + popc) and lane 0 adds the E list while lane 1 adds the I list: the
// same additions in the same order, in max(#E, #I) steps of LDS + FADD instead of 32 steps of shuffle + two
// FMNMX + two FADD executed by the whole warp (ncu: that replay was ~70% of the loop's 17 K warp instructions
// per swap, on a kernel whose issue slots are 41% busy).
template <bool GBITS>
__device__ __forceinline__ float warp_row_value_local(const int32_t *__restrict__ col, const float *__restrict__ w,
                                                      const uint32_t *bits, const uint8_t *state, int32_t lo, int32_t hi,
                                                      int32_t ov_a, int32_t ov_b, int lane, int32_t c, float ww,
                                                      float *wsm /* 64 floats of this warp */) {
  float acc = 0.0f;                                  // lane 0: E, lane 1: I
  const unsigned lt = (1u << lane) - 1u;
  for (int32_t base = lo; base < hi; base += 32) {
    const bool valid = (base + lane) < hi;
    int32_t cn = 0;
    float wn = 0.0f;
    const int32_t in = base + 32 + lane;
    if (in < hi) { cn = __ldg(col + in); wn = __ldg(w + in); }   // next 32 entries while this chunk is summed
    bool ext = false;
    if (valid) {
      if (c == ov_a) ext = true;
      else if (c == ov_b) ext = false;
      else ext = (kl_state_get<GBITS>(bits, state, c) & ST_SIDE) != 0u;
    }
    const unsigned P = __ballot_sync(FULL_MASK, valid && ext), N = __ballot_sync(FULL_MASK, valid && !ext);
    if (valid) wsm[ext ? __popc(P & lt) : 32 + __popc(N & lt)] = ww;
    __syncwarp();
    if (lane < 2) {
      const float *src = wsm + 32 * lane;
      const int cnt = __popc(lane ? N : P);
      for (int t = 0; t < cnt; ++t) acc = __fadd_rn(acc, src[t]);
    }
    __syncwarp();
    c = cn; ww = wn;
  }
  const float E = __shfl_sync(FULL_MASK, acc, 0), I = __shfl_sync(FULL_MASK, acc, 1);
  return __fsub_rn(E, I);
}

template <bool ASC>
__device__ __forceinline__ unsigned long long kl_key(float v, unsigned side, uint32_t ident) {
  return ((unsigned long long)float_orderable(side ? -v : v) << 32) | (unsigned long long)(0xFFFFFFFFu - ident);
}

// best keys of one tile (both sides); one L2 round trip (the D-values, and the ranks unless ASC)
template <bool ASC, bool GBITS>
__device__ __forceinline__ void tile_scan_local(const uint32_t *bits, const uint8_t *state, const float *val,
                                                const uint32_t *__restrict__ rank, int32_t n, int32_t tile, int lane,
                                                unsigned long long *keys) {
  unsigned long long k0 = 0ull, k1 = 0ull;
  const int32_t base = tile * KL_TILE;
  float vv[KL_TILE / 32];
  uint32_t id[KL_TILE / 32];
  unsigned sg[KL_TILE / 32];
#pragma unroll
  for (int r = 0; r < KL_TILE / 32; ++r) {
    const int32_t u = base + r * 32 + lane;
    vv[r] = u < n ? __ldcg(val + u) : 0.0f;
    id[r] = ASC ? (uint32_t)u : (u < n ? __ldg(rank + u) : 0u);
    sg[r] = (GBITS && u < n) ? ((unsigned)__ldcg(state + u) & 3u) : ST_LOCK;   // same round trip as the D-values
  }
#pragma unroll
  for (int r = 0; r < KL_TILE / 32; ++r) {
    const int32_t u = base + r * 32 + lane;
    const unsigned st = GBITS ? sg[r] : (u < n ? bits_get(bits, u) : ST_LOCK);
    if (!(st & ST_LOCK)) {
      const unsigned long long key = kl_key<ASC>(vv[r], st & ST_SIDE, id[r]);
      if (st & ST_SIDE) k1 = key > k1 ? key : k1;
      else k0 = key > k0 ? key : k0;
    }
  }
  k0 = warp_max_u64(k0);
  k1 = warp_max_u64(k1);
  if (lane == 0) { keys[2 * tile] = k0; keys[2 * tile + 1] = k1; }
}

struct KlLocalParams {
  int32_t n, n_tiles;
  const int32_t *rowptr, *col;
  const int2 *nb;                   // per entry: row extent of the neighbour it names
  const float *w;
  uint8_t *state;
  const uint32_t *rank;
  const int32_t *order0, *order1;
  float *val;
  float *t_cut, *t_gain;
  int32_t *t_n1, *t_n2;
  int64_t *ctrl;
  float cut0;
  uint32_t term_limit;
  int64_t n0, n1;
  int clocks;
};

__global__ void nb_extent_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, int64_t nnz, int2 *__restrict__ nb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) {
    const int32_t v = col[i];
    nb[i] = make_int2(rowptr[v], rowptr[v + 1]);
  }
}

template <bool ASC, bool GBITS>
__global__ void __launch_bounds__(KL_LOOP_THREADS, 1) kl_loop_local_kernel(const KlLocalParams p) {
  extern __shared__ __align__(16) unsigned char kl_sm[];
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(kl_sm);      // 2 * n_tiles
  uint32_t *stamps = reinterpret_cast<uint32_t *>(keys + 2 * (size_t)p.n_tiles);  // n_tiles
  int32_t *list = reinterpret_cast<int32_t *>(stamps + p.n_tiles);                // n_tiles: tiles to rescan this swap
  uint32_t *bits = reinterpret_cast<uint32_t *>(list + p.n_tiles);                // ceil(n / 16)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ unsigned long long red0[32], red1[32];
  __shared__ unsigned long long sh_best[2];
  __shared__ float sh_wsm[KL_LOOP_THREADS / 32][64];   // per warp: the E and I weight lists of the row being summed
  __shared__ float sh_cut;
  __shared__ uint32_t sh_term, sh_iter;
  __shared__ int sh_done, sh_nlist;
  __shared__ long long sh_rem0, sh_rem1;
  if (tid == 0) {
    sh_cut = p.cut0; sh_term = 0; sh_iter = 0; sh_rem0 = p.n0; sh_rem1 = p.n1; sh_nlist = 0;
    sh_done = (p.n0 <= 0 || p.n1 <= 0) ? 1 : 0;
  }
  // side / locked bits of every node: 16 nodes per word
  const int32_t n_words = GBITS ? 0 : (p.n + 15) >> 4;
  for (int32_t wd = tid; wd < n_words; wd += KL_LOOP_THREADS) {
    uint32_t v = 0u;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int32_t u = wd * 16 + q;
      const unsigned st = u < p.n ? (unsigned)p.state[u] & 3u : ST_LOCK;
      v |= st << (2 * q);
    }
    bits[wd] = v;
  }
  for (int32_t t = tid; t < p.n_tiles; t += KL_LOOP_THREADS) stamps[t] = 0u;
  __syncthreads();
  for (int32_t t = warp; t < p.n_tiles; t += KL_LOOP_THREADS / 32)
    tile_scan_local<ASC, GBITS>(bits, p.state, p.val, p.rank, p.n, t, lane, keys);
  __syncthreads();
  uint32_t it_local = 0;
  // phase clocks of thread 0 (EIGKL_KL_PHASES=1): one predicated clock read per phase
  long long tph[6] = {0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
#define KL_PHASE(i) do { if (p.clocks && tid == 0) { const long long t_ = clock64(); tph[i] += t_ - tprev; tprev = t_; } } while (0)
  // finer probes (thread 0, shared-memory accumulators so that they cost no registers): ctrl[16 + i]
  __shared__ long long sh_fine[16];
  __shared__ long long sh_fprev;
  if (tid == 0) { for (int i = 0; i < 16; ++i) sh_fine[i] = 0; sh_fprev = clock64(); }
#define KL_FINE(i) do { if (p.clocks && tid == 0) { const long long t_ = clock64(); sh_fine[i] += t_ - sh_fprev; sh_fprev = t_; } } while (0)

  while (!sh_done) {
    // ---- S1: best pair over the tile keys (shared memory) ----
    unsigned long long k0 = 0ull, k1 = 0ull;
    for (int32_t t = tid; t < p.n_tiles; t += KL_LOOP_THREADS) {
      const unsigned long long a0 = keys[2 * t], a1 = keys[2 * t + 1];
      k0 = a0 > k0 ? a0 : k0;
      k1 = a1 > k1 ? a1 : k1;
    }
    k0 = warp_max_u64(k0);
    k1 = warp_max_u64(k1);
    if (lane == 0) { red0[warp] = k0; red1[warp] = k1; }
    if (tid == 0) sh_nlist = 0;
    KL_FINE(0);
    __syncthreads();
    KL_FINE(1);
    if (warp == 0) {
      k0 = warp_max_u64(red0[lane]);
      k1 = warp_max_u64(red1[lane]);
      if (lane == 0) { sh_best[0] = k0; sh_best[1] = k1; }
    }
    KL_FINE(2);
    __syncthreads();
    KL_FINE(3);
    const unsigned long long b0 = sh_best[0], b1 = sh_best[1];
    if (b0 == 0ull || b1 == 0ull) {                  // no selectable node on one side (cKL.cpp:387-389)
      if (tid == 0) sh_done = 1;
      __syncthreads();
      break;
    }
    KL_PHASE(0);
    ++it_local;
    const uint32_t ia = 0xFFFFFFFFu - (uint32_t)(b0 & 0xFFFFFFFFull), ib = 0xFFFFFFFFu - (uint32_t)(b1 & 0xFFFFFFFFull);
    const int32_t a = ASC ? (int32_t)ia : __ldg(p.order0 + ia);
    const int32_t b = ASC ? (int32_t)ib : __ldg(p.order1 + ib);
    const int32_t alo = __ldg(p.rowptr + a), ahi = __ldg(p.rowptr + a + 1);
    const int32_t blo = __ldg(p.rowptr + b), bhi = __ldg(p.rowptr + b + 1);
    const int32_t da = ahi - alo, items = da + (bhi - blo);
    KL_PHASE(1);
    KL_FINE(4);
    constexpr int WORKERS = KL_LOOP_THREADS / 32 - 1;    // warp 31 is the bookkeeper
    const uint32_t stamp = it_local;
    if (warp == WORKERS) {
      // ---- S2 (off the critical path): gain, cut, trace, termination, lock-and-swap ----
      float wab = 0.0f;
      for (int32_t i = alo + lane; i < ahi; i += 32)
        if (__ldg(p.col + i) == b) wab = __ldg(p.w + i);             // getEdgeWeight, cKL.cpp:75-82
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wab = fmaxf(wab, __shfl_xor_sync(FULL_MASK, wab, o));   // weights are > 0
      if (lane == 0) {
        const float maxGain = float_from_orderable((uint32_t)(b0 >> 32));
        const float minGain = __fsub_rn(0.0f, float_from_orderable((uint32_t)(b1 >> 32)));
        const float gain = __fsub_rn(__fsub_rn(maxGain, minGain), __fmul_rn(2.0f, wab));    // cKL.cpp:360
        const float cut = __fsub_rn(sh_cut, gain);                                          // cKL.cpp:362
        sh_cut = cut;
        sh_iter = it_local;
        p.t_cut[it_local] = cut; p.t_gain[it_local] = gain; p.t_n1[it_local] = a; p.t_n2[it_local] = b;
        __stcg(p.state + a, (uint8_t)(ST_SIDE | ST_LOCK));           // swip, cKL.cpp:274-286
        __stcg(p.state + b, (uint8_t)(ST_LOCK));
        // the workers take a's and b's sides from (a, b) directly, never from these words
        if (!GBITS) {
          bits[a >> 4] = (bits[a >> 4] & ~(3u << ((a & 15) * 2))) | ((ST_SIDE | ST_LOCK) << ((a & 15) * 2));
          bits[b >> 4] = (bits[b >> 4] & ~(3u << ((b & 15) * 2))) | (ST_LOCK << ((b & 15) * 2));
        }
        // a and b held the best keys of their tiles: those tiles are rescanned
        if (atomicExch(stamps + a / KL_TILE, stamp) != stamp) list[atomicAdd(&sh_nlist, 1)] = a / KL_TILE;
        if (atomicExch(stamps + b / KL_TILE, stamp) != stamp) list[atomicAdd(&sh_nlist, 1)] = b / KL_TILE;
        if (gain <= 0.0f) { if (++sh_term > p.term_limit) sh_done = 1; }   // cKL.cpp:382-386
        else sh_term = 0;
        if (--sh_rem0 == 0) sh_done = 1;
        if (--sh_rem1 == 0) sh_done = 1;
      }
    } else {
      // ---- S3: recompute D of every neighbour of a or b from scratch (cKL.cpp:253-272) ----
      // lane k holds the warp's k-th neighbour (id, row extent, rank): ONE round trip for all of them; then the
      // first 32 entries of four rows at a time are in flight before the first row is summed
      const int32_t n_mine = items > warp ? (items - 1 - warp) / WORKERS + 1 : 0;
      for (int32_t k0i = 0; k0i < n_mine; k0i += 32) {
        int32_t my_v = 0;
        int2 my_ext = make_int2(0, 0);
        const int32_t kidx = k0i + lane;
        if (kidx < n_mine) {
          const int32_t it = warp + kidx * WORKERS;
          const int32_t e = it < da ? alo + it : blo + (it - da);
          my_v = __ldg(p.col + e);
          my_ext = __ldg(p.nb + e);
        }
        uint32_t my_id = (uint32_t)my_v;
        if (!ASC && kidx < n_mine) my_id = __ldg(p.rank + my_v);
        const int cnt = min(32, n_mine - k0i);
        if (p.clocks && tid == 0 && my_ext.x + my_v == -12345) sh_fine[15] = 1;     // a use of the loaded values: the probe below sees their arrival
        KL_FINE(5);
        for (int j0 = 0; j0 < cnt; j0 += 4) {
          int32_t lo[4], hi[4], c[4];
          float ww[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = min(j0 + u, cnt - 1);
            lo[u] = __shfl_sync(FULL_MASK, my_ext.x, j);
            hi[u] = __shfl_sync(FULL_MASK, my_ext.y, j);
            if (j0 + u >= cnt) hi[u] = lo[u];
            c[u] = 0; ww[u] = 0.0f;
            if (lo[u] + lane < hi[u]) { c[u] = __ldg(p.col + lo[u] + lane); ww[u] = __ldg(p.w + lo[u] + lane); }
          }
          if (p.clocks && tid == 0 && c[0] + c[1] + c[2] + c[3] == -12345) sh_fine[15] = 2;
          KL_FINE(6);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (j0 + u < cnt) {                                    // warp-uniform
              const int32_t v = __shfl_sync(FULL_MASK, my_v, j0 + u);
              const uint32_t vid = __shfl_sync(FULL_MASK, my_id, j0 + u);
              const float nv = warp_row_value_local<GBITS>(p.col, p.w, bits, p.state, lo[u], hi[u], a, b, lane, c[u], ww[u], sh_wsm[warp]);
              KL_FINE(7);
              if (lane == 0) {
                __stcg(p.val + v, nv);
                const unsigned st = kl_state_get<GBITS>(bits, p.state, v);
                if (!(st & ST_LOCK) && v != a && v != b) {
                  // the tile's key: raised in place, unless v itself held it (then the tile is rescanned)
                  const int32_t tile = v / KL_TILE;
                  unsigned long long *kp = keys + 2 * tile + (st & ST_SIDE);
                  const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(kp);
                  if (cur != 0ull && (uint32_t)(cur & 0xFFFFFFFFull) == 0xFFFFFFFFu - vid) {
                    if (atomicExch(stamps + tile, stamp) != stamp) list[atomicAdd(&sh_nlist, 1)] = tile;
                  } else {
                    atomicMax(kp, kl_key<ASC>(nv, st & ST_SIDE, vid));
                  }
                }
              }
              KL_FINE(8);
            }
          }
        }
      }
    }
    KL_PHASE(2);
    KL_FINE(9);
    __syncthreads();
    KL_PHASE(3);
    KL_FINE(10);
    // ---- S4: rescan the few tiles whose best node was touched or locked, one per warp ----
    {
      const int nl = sh_nlist;
      if (p.clocks && tid == 0) sh_fine[14] += nl;
      for (int q = warp; q < nl; q += KL_LOOP_THREADS / 32)
        tile_scan_local<ASC, GBITS>(bits, p.state, p.val, p.rank, p.n, list[q], lane, keys);
    }
    KL_PHASE(4);
    KL_FINE(11);
    __syncthreads();
    KL_PHASE(5);
    KL_FINE(12);
  }
  if (tid == 0) {
    p.ctrl[0] = (int64_t)sh_iter; p.ctrl[1] = 1;
    if (p.clocks) {
      for (int i = 0; i < 6; ++i) p.ctrl[8 + i] = tph[i];
      for (int i = 0; i < 16; ++i) p.ctrl[16 + i] = sh_fine[i];
    }
  }
#undef KL_PHASE
}

// ---------------------------------------------------------------------------------------------------
// The swap loop, third form ("flat"): the same state in shared memory as kl_loop_local_kernel, laid out by ENTRY.
// ncu on the earlier forms (profiles/r02/kl_flat_v1_ncu_lines.txt): the loop is not waiting for memory -- an L2 hit is 282
// cycles from a lone CTA (tools/micro/chase.cu) and the four dependent trips of a swap explain ~1 100 of its
// ~10 000 cycles -- it is issue- and barrier-bound: 14 800 warp instructions per swap, a third of them the per-warp
// slot bookkeeping that spread ~40 neighbour rows over 15 warps, and a third of all warp cycles spent waiting at
// block barriers for the slowest warp (the 256-entry rescans of the two winners' tiles, the group refolds).
// This form does the least work per swap that the arithmetic allows:
//   * ITEMS: the neighbour list N(a) ++ N(b), one item per thread, dense from thread 0 (ibm10: 41 items = 2
//     warps).  One load gives (v, rowptr[v], rowptr[v+1]); a warp scan + the warps' totals give every row's
//     offset in ONE flat list of entries.
//   * ENTRIES: thread g handles flat entry g (and g + 416, ...): a 4-step search over the warp totals (shuffles)
//     and a 5-step search over its warp's offsets (shared memory) name the row, then ONE coalesced round of loads
//     (col, w), the side bit from shared memory, and the signed weight goes to stage[g].  No per-row loops, no
//     slots: ~60 instructions per thread whatever the row lengths.
//   * SUMS: the item's thread adds its row's segment of stage[] serially -- the fp32 order of cKL.cpp:225-251 --
//     all rows concurrently, and publishes the D-value (tile key raised in place, late rescan only when the
//     holder of a tile's best key lost ground).
//   * the tiles of a and b always need new keys (their best nodes were just locked).  Two otherwise idle warps
//     load those tiles' D-values at the START of the swap and reduce them to a base key per side, leaving out
//     the nodes this swap recomputes (an exclusion bitmap the item threads fill); recomputed nodes of those
//     tiles go to a per-tile key by atomicMax.  After the barrier the tile's key is max(base, published): two
//     shared-memory words instead of a 256-entry rescan behind the barrier.
//   * the same two warps then refold the groups of those tiles; when no late rescan was requested (9 swaps in
//     10) that is the whole epilogue: one barrier instead of three.
// The D-values and state bytes are read and written with plain loads and stores here: one CTA, ordered by its own
// barriers, needs no L1 bypass (the .cg forms of the cluster kernel compile to STRONG.GPU accesses, and a strong byte
// load in the middle of the item phase cost ~2 400 cycles per swap when the state bytes live in global memory).
// Hub swaps (more neighbours than row threads, or more entries than the staging buffer) take the warp-per-row
// replay.  The arithmetic is untouched: traces stay byte-identical (test_kl_loop_variants_byte_exact).
// ---------------------------------------------------------------------------------------------------
constexpr int KLF_THREADS = 512;
constexpr int KLF_WARPS = KLF_THREADS / 32;
constexpr int KLF_ROW_WARPS = 13;               // warps 0..12: items, entries, sums
constexpr int KLF_E0 = 13, KLF_E1 = 14;         // the early-tile warps (tile of a, tile of b)
constexpr int KLF_BK = 15;                      // the bookkeeper
constexpr int KLF_ROW_THREADS = KLF_ROW_WARPS * 32;
constexpr int KLF_MAXN = KLF_ROW_THREADS;       // |N(a)| + |N(b)| handled on the flat path
constexpr int KLF_ENT = 4096;                   // staged entries per swap
constexpr int KLF_LCAP = KLF_MAXN + 8;          // late rescans / dirty groups per swap
constexpr int KL_GROUP = 32;                    // tiles per group key

struct KlFlatSmem {                             // fixed part, placed after the size-dependent arrays
  float stage[KLF_ENT];
  int32_t it_lo[KLF_ROW_THREADS + 1];           // first entry of the item's row
  int32_t it_len[KLF_ROW_THREADS + 1];          // its length
  int32_t it_inc[KLF_ROW_THREADS];              // inclusive prefix of the row lengths inside the item's warp
  unsigned long long best[2];                   // the pair selection of the NEXT swap when best_stamp names this one
  uint32_t best_stamp;
  int32_t wtot[16];                             // entries of the items of each row warp
  unsigned long long pkey[2][2][2];             // [swap parity][tile of a / tile of b][side]: best published key
  uint32_t excl[2][2][KL_TILE / 32];            // [swap parity][tile of a / b]: nodes recomputed in this swap
  int32_t list[2][KLF_LCAP];                    // late rescans of a swap, double-buffered by the swap's parity: the counters
  int32_t dlist[2][KLF_LCAP + 8];               // of the NEXT swap are cleared while this swap's lists are still being read
  float wab;
  float cut;
  uint32_t term, iter;
  int done, nlist[2], ndirty[2];
  long long rem0, rem1;
  long long fine[16];
  long long fprev;
};

struct KlfCtx {
  unsigned long long *keys, *gkeys;
  uint32_t *stamps;
  KlFlatSmem *S;
  float *val;
  int32_t a, b, ta, tb;
  uint32_t stamp;
  int par;
};
__device__ __forceinline__ uint32_t *key_lo(unsigned long long *k) { return reinterpret_cast<uint32_t *>(k); }       // ~position
__device__ __forceinline__ uint32_t *key_hi(unsigned long long *k) { return reinterpret_cast<uint32_t *>(k) + 1; }   // orderable(D)
template <bool ASC>
__device__ __forceinline__ void klf_publish(const KlfCtx &X, int32_t v, uint32_t vid, unsigned st, float nv) {
  __stcg(X.val + v, nv);
  if ((st & ST_LOCK) || v == X.a || v == X.b) return;
  const int32_t tile = v / KL_TILE;
  KlFlatSmem &S = *X.S;
  const unsigned long long nk = kl_key<ASC>(nv, st & ST_SIDE, vid);
  if (tile == X.ta || tile == X.tb) {                 // the early warp of that tile folds this in after the barrier
    atomicMax(&S.pkey[X.par][tile == X.ta ? 0 : 1][st & ST_SIDE], nk);
    return;
  }
  unsigned long long *kp = X.keys + 2 * tile + (st & ST_SIDE);
  const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(kp);
  if (cur != 0ull && (uint32_t)(cur & 0xFFFFFFFFull) == 0xFFFFFFFFu - vid && nk < cur) {
    // v held the tile's best key and lost ground: the tile is rescanned
    if (atomicExch(X.stamps + tile, X.stamp) != X.stamp) {
      const int slot = atomicAdd(&S.nlist[X.par], 1);
      if (slot < KLF_LCAP) S.list[X.par][slot] = tile;            // beyond the list (a hub swap): every tile is rescanned
    }
  } else if (atomicMax(kp, nk) < nk) {
    if (X.gkeys != nullptr) atomicMax(X.gkeys + 2 * (tile / KL_GROUP) + (st & ST_SIDE), nk);
  }
}
template <bool ASC, bool GBITS>
__device__ __noinline__ void klf_row_replay(const KlLocalParams &p, const KlfCtx &X, const uint32_t *bits, int32_t v, uint32_t vid,
                                            int32_t lo, int32_t hi, float *wsm) {
  const int lane = threadIdx.x & 31;
  int32_t c = 0;
  float ww = 0.0f;
  if (lo + lane < hi) { c = __ldg(p.col + lo + lane); ww = __ldg(p.w + lo + lane); }
  const float r = warp_row_value_local<GBITS>(p.col, p.w, bits, p.state, lo, hi, X.a, X.b, lane, c, ww, wsm);
  if (lane == 0) klf_publish<ASC>(X, v, vid, kl_state_get<GBITS>(bits, p.state, v), r);
}
template <bool ASC, bool GBITS>
__device__ __noinline__ void klf_tile_rescan(const KlLocalParams &p, const uint32_t *bits, int32_t tile, unsigned long long *keys) {
  tile_scan_local<ASC, GBITS>(bits, p.state, p.val, p.rank, p.n, tile, threadIdx.x & 31, keys);
}
__device__ __forceinline__ void klf_bar_items() { asm volatile("bar.sync 1, %0;" ::"n"((KLF_ROW_WARPS + 2) * 32) : "memory"); }   // row + early warps
__device__ __forceinline__ void klf_bar_rows() { asm volatile("bar.sync 2, %0;" ::"n"(KLF_ROW_WARPS * 32) : "memory"); }          // row warps
__device__ __forceinline__ void klf_bar_early_arrive() { __threadfence_block(); asm volatile("bar.arrive 3, 64;" ::: "memory"); }
__device__ __forceinline__ void klf_bar_early_sync() { asm volatile("bar.sync 3, 64;" ::: "memory"); }

// ONE: up to 64 tiles (16 384 nodes: fract, ibm01, industry2) the pair selection scans the tile keys themselves -- one
// reduction instead of group fold + global fold on the path between two swaps, and no group keys to raise at a publish
template <bool ASC, bool GBITS, bool CLOCKS, bool ONE>
__global__ void __launch_bounds__(KLF_THREADS, 1) kl_loop_flat_kernel(const KlLocalParams p) {
  extern __shared__ __align__(16) unsigned char kl_sm[];
  const int32_t n_groups = (p.n_tiles + KL_GROUP - 1) / KL_GROUP;
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(kl_sm);       // 2 * n_tiles
  unsigned long long *gkeys = keys + 2 * (size_t)p.n_tiles;                        // 2 * n_groups
  KlFlatSmem &S = *reinterpret_cast<KlFlatSmem *>(gkeys + 2 * (size_t)n_groups);
  uint32_t *stamps = reinterpret_cast<uint32_t *>(&S + 1);                         // n_tiles
  uint32_t *gstamp = stamps + p.n_tiles;                                           // n_groups
  uint32_t *bits = gstamp + n_groups;                                              // ceil(n / 16) unless GBITS
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    S.cut = p.cut0; S.term = 0; S.iter = 0; S.rem0 = p.n0; S.rem1 = p.n1; S.wab = 0.0f;
    S.nlist[0] = S.nlist[1] = 0; S.ndirty[0] = S.ndirty[1] = 0;
    S.done = (p.n0 <= 0 || p.n1 <= 0) ? 1 : 0;
    for (int i = 0; i < 16; ++i) S.fine[i] = 0;
    S.fprev = clock64();
  }
  const int32_t n_words = GBITS ? 0 : (p.n + 15) >> 4;
  for (int32_t wd = tid; wd < n_words; wd += KLF_THREADS) {
    uint32_t v = 0u;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int32_t u = wd * 16 + q;
      const unsigned st = u < p.n ? (unsigned)p.state[u] & 3u : ST_LOCK;
      v |= st << (2 * q);
    }
    bits[wd] = v;
  }
  for (int32_t t = tid; t < p.n_tiles; t += KLF_THREADS) stamps[t] = 0u;
  for (int32_t g = tid; g < n_groups; g += KLF_THREADS) gstamp[g] = 0u;
  if (tid < 8) (&S.pkey[0][0][0])[tid] = 0ull;
  if (tid < 2 * 2 * (KL_TILE / 32)) (&S.excl[0][0][0])[tid] = 0u;
  if (tid < 16) S.wtot[tid] = 0;
  if (tid == 0) { S.best_stamp = 0xFFFFFFFFu; S.it_lo[KLF_ROW_THREADS] = 0; S.it_len[KLF_ROW_THREADS] = 0; }
  __syncthreads();
  for (int32_t t = warp; t < p.n_tiles; t += KLF_WARPS) klf_tile_rescan<ASC, GBITS>(p, bits, t, keys);
  __syncthreads();
  // group keys of every group
  auto group_fold = [&](int32_t g) {
    const int32_t t = g * KL_GROUP + lane;
    unsigned long long k0 = 0ull, k1 = 0ull;
    if (t < p.n_tiles) { k0 = keys[2 * t]; k1 = keys[2 * t + 1]; }
    k0 = warp_max_u64(k0);
    k1 = warp_max_u64(k1);
    if (lane == 0) { gkeys[2 * g] = k0; gkeys[2 * g + 1] = k1; }
  };
  for (int32_t g = warp; g < n_groups; g += KLF_WARPS) group_fold(g);
  __syncthreads();
  const unsigned long long *sel = ONE ? keys : gkeys;               // what the selection scans: tile keys or group keys
  const int32_t n_sel = ONE ? p.n_tiles : n_groups;
#define KLF_FINE(i) do { if (CLOCKS && tid == 0) { const long long t_ = clock64(); S.fine[i] += t_ - S.fprev; S.fprev = t_; } } while (0)
  uint32_t it_local = 0;
  while (true) {
    // ---- S1: every warp folds the group keys (identical result in every warp, no barrier) ----
    if (S.done) break;
    unsigned long long b0, b1;
    if (S.best_stamp == it_local) {                    // the warp that refolded the groups of the last swap left the answer
      b0 = S.best[0]; b1 = S.best[1];
    } else {
      unsigned long long k0 = 0ull, k1 = 0ull;
#pragma unroll 1
      for (int32_t g = lane; g < n_sel; g += 32) {
        const unsigned long long a0 = sel[2 * g], a1 = sel[2 * g + 1];
        k0 = a0 > k0 ? a0 : k0;
        k1 = a1 > k1 ? a1 : k1;
      }
      b0 = warp_max_u64(k0); b1 = warp_max_u64(k1);
    }
    if (b0 == 0ull || b1 == 0ull) break;               // no selectable node on one side (cKL.cpp:387-389)
    KLF_FINE(0);
    ++it_local;
    const uint32_t stamp = it_local;
    const int par = (int)(stamp & 1u);
    const uint32_t ia = 0xFFFFFFFFu - (uint32_t)(b0 & 0xFFFFFFFFull), ib = 0xFFFFFFFFu - (uint32_t)(b1 & 0xFFFFFFFFull);
    const int32_t a = ASC ? (int32_t)ia : __ldg(p.order0 + ia);
    const int32_t b = ASC ? (int32_t)ib : __ldg(p.order1 + ib);
    const int32_t ta = a / KL_TILE, tb = b / KL_TILE;
    // ---- early loads: the tiles of a and b (warps E0 / E1), before anything else is in flight ----
    float ev[KL_TILE / 32];
    uint32_t eid[KL_TILE / 32];
    unsigned esg[KL_TILE / 32];
    const bool early = (warp == KLF_E0) || (warp == KLF_E1 && tb != ta);
    const int32_t et = (warp == KLF_E1) ? tb : ta;
    const int es = (warp == KLF_E1) ? 1 : 0;
    if (early) {
#pragma unroll
      for (int r = 0; r < KL_TILE / 32; ++r) {
        const int32_t u = et * KL_TILE + r * 32 + lane;
        ev[r] = u < p.n ? p.val[u] : 0.0f;
        eid[r] = ASC ? (uint32_t)u : (u < p.n ? __ldg(p.rank + u) : 0u);
        esg[r] = (GBITS && u < p.n) ? (unsigned)p.state[u] : ST_LOCK;       // masked where it is used
      }
    }
    const int32_t alo = __ldg(p.rowptr + a), ahi = __ldg(p.rowptr + a + 1);
    const int32_t blo = __ldg(p.rowptr + b), bhi = __ldg(p.rowptr + b + 1);
    const int32_t da = ahi - alo, items = da + (bhi - blo);
    KLF_FINE(1);
    // the context of the out-of-line hub routines is built only where they are called: it lives in local memory
    auto make_ctx = [&](KlfCtx &X) {
      X.keys = keys; X.gkeys = ONE ? nullptr : gkeys; X.stamps = stamps; X.S = &S; X.val = p.val;
      X.a = a; X.b = b; X.ta = ta; X.tb = tb; X.stamp = stamp; X.par = par;
    };
    unsigned long long base0 = 0ull, base1 = 0ull;     // early warps: best keys of the tile's nodes this swap leaves alone
    if (warp == KLF_BK) {
      // ---- S2, concurrently with S3: gain, cut, trace, termination, lock-and-swap (the row warps take the sides of a
      //      and b from (a, b) themselves, never from the words updated here) ----
      float wab = 0.0f;
      for (int32_t i = alo + lane; i < ahi; i += 32)
        if (__ldg(p.col + i) == b) wab = __ldg(p.w + i);                                     // getEdgeWeight, cKL.cpp:75-82
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wab = fmaxf(wab, __shfl_xor_sync(FULL_MASK, wab, o));   // weights are > 0
      if (lane == 0) {
        const float maxGain = float_from_orderable((uint32_t)(b0 >> 32));
        const float minGain = __fsub_rn(0.0f, float_from_orderable((uint32_t)(b1 >> 32)));
        const float gain = __fsub_rn(__fsub_rn(maxGain, minGain), __fmul_rn(2.0f, wab));      // cKL.cpp:360
        const float cut = __fsub_rn(S.cut, gain);                                            // cKL.cpp:362
        S.cut = cut; S.iter = it_local;
        p.t_cut[it_local] = cut; p.t_gain[it_local] = gain; p.t_n1[it_local] = a; p.t_n2[it_local] = b;
        if (gain <= 0.0f) { if (++S.term > p.term_limit) S.done = 1; }                       // cKL.cpp:382-386
        else S.term = 0;
      } else if (lane == 1) {
        p.state[a] = (uint8_t)(ST_SIDE | ST_LOCK);                                           // swip, cKL.cpp:274-286
        if (!GBITS) bits[a >> 4] = (bits[a >> 4] & ~(3u << ((a & 15) * 2))) | ((ST_SIDE | ST_LOCK) << ((a & 15) * 2));
        if (--S.rem0 == 0) S.done = 1;
      } else if (lane == 2) {
        __stcg(p.state + b, (uint8_t)(ST_LOCK));
        if (--S.rem1 == 0) S.done = 1;
        S.nlist[par ^ 1] = 0; S.ndirty[par ^ 1] = 0;               // the next swap's lists (nobody touches them now)
      } else if (lane >= 4 && lane < 8) {
        (&S.pkey[par ^ 1][0][0])[lane - 4] = 0ull;
      } else if (lane >= 8 && lane < 8 + 2 * (KL_TILE / 32)) {
        (&S.excl[par ^ 1][0][0])[lane - 8] = 0u;
      }
      __syncwarp();
      if (!GBITS && lane == 2) {                                    // after lane 1's word update (a and b may share a word)
        bits[b >> 4] = (bits[b >> 4] & ~(3u << ((b & 15) * 2))) | (ST_LOCK << ((b & 15) * 2));
      }
    } else if (warp >= KLF_ROW_WARPS) {
      // ---- the early warps: wait for the exclusion bitmap, fold the tile's untouched nodes into a base key per side ----
      klf_bar_items();
      if (early) {
#pragma unroll
        for (int r = 0; r < KL_TILE / 32; ++r) {
          const int idx = r * 32 + lane;
          const int32_t u = et * KL_TILE + idx;
          if (u >= p.n || u == a || u == b) continue;
          if ((S.excl[par][es][r] >> lane) & 1u) continue;          // recomputed in this swap: arrives through pkey
          const unsigned st = GBITS ? (esg[r] & 3u) : bits_get(bits, u);
          if (st & ST_LOCK) continue;
          const unsigned long long key = kl_key<ASC>(ev[r], st & ST_SIDE, eid[r]);
          if (st & ST_SIDE) base1 = key > base1 ? key : base1; else base0 = key > base0 ? key : base0;
        }
        base0 = warp_max_u64(base0);
        base1 = warp_max_u64(base1);
      }
    } else if (items <= KLF_MAXN) {
      // ---- S3, flat.  ITEMS: this thread's neighbour, its row extent, the row's offset in the flat entry list ----
      int32_t my_v = 0, my_len = 0;
      uint32_t my_id = 0u;
      unsigned my_st = ST_LOCK;
      int32_t my_lo = 0;
      if (tid < items) {
        const int32_t e = tid < da ? alo + tid : blo + (tid - da);
        my_v = __ldg(p.col + e);
        const int2 ext = __ldg(p.nb + e);
        my_lo = ext.x; my_len = ext.y - ext.x;
        if (CLOCKS && my_len >= 0) KLF_FINE(11);
        my_id = ASC ? (uint32_t)my_v : __ldg(p.rank + my_v);
        if (GBITS) my_st = (unsigned)p.state[my_v];                   // needed (and masked) at the publish: in flight across the entry phase
        const int32_t tv = my_v / KL_TILE;
        if (tv == ta) atomicOr(&S.excl[par][0][(my_v % KL_TILE) >> 5], 1u << (my_v & 31));
        else if (tv == tb) atomicOr(&S.excl[par][1][(my_v % KL_TILE) >> 5], 1u << (my_v & 31));
      }
      int inc = my_len;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(FULL_MASK, inc, d);
        if (lane >= d) inc += t;
      }
      if (CLOCKS && inc >= 0) KLF_FINE(12);
      S.it_lo[tid] = my_lo;
      S.it_len[tid] = my_len;
      S.it_inc[tid] = inc;
      if (lane == 31) S.wtot[warp] = inc;
      KLF_FINE(2);
      klf_bar_items();                                   // items, offsets and the exclusion bitmap are in
      KLF_FINE(3);
      // inclusive prefix of the warps' totals: lane j holds the entries of warps 0..j
      int winc = lane < KLF_ROW_WARPS ? S.wtot[lane] : 0;
#pragma unroll
      for (int d = 1; d < 16; d <<= 1) {
        const int t = __shfl_up_sync(FULL_MASK, winc, d);
        if (lane >= d) winc += t;
      }
      const int total = __shfl_sync(FULL_MASK, winc, KLF_ROW_WARPS - 1);
      const int my_base = __shfl_sync(FULL_MASK, winc, warp) - __shfl_sync(FULL_MASK, inc, 31);   // first flat entry of this warp's items
      float *stg = S.stage;
      if (total <= KLF_ENT) {
        // ---- ENTRIES: thread t owns the K consecutive flat entries from t*K: ONE search names the row of the first (4
        //      shuffle steps over the warps' totals, 5 shared-memory steps over that warp's offsets), the rest follow by
        //      walking; the loads of a chunk of 4 are in flight together; the signed weights go to stage[] ----
        const int K = (total + KLF_ROW_THREADS - 1) / KLF_ROW_THREADS;         // 1 .. KLF_ENT / KLF_ROW_THREADS + 1
        const int32_t g = tid * K;
        int w = 0, wb = 0;                                                      // the item warp that holds entry g, the entries below it
        const int n_item_warps = (items + 31) >> 5;                             // 2 on ibm10, 1 on the synthetic circuits
#pragma unroll 1
        for (int cand = 1; cand < n_item_warps; ++cand) {
          const int below = __shfl_sync(FULL_MASK, winc, cand - 1);             // entries of warps 0..cand-1
          if (below <= g) { w = cand; wb = below; }
        }
        if (g < total) {
          const int32_t local = g - wb;
          const int32_t *incs = S.it_inc + w * 32;
          int k = 0;
#pragma unroll
          for (int step = 16; step > 0; step >>= 1) {
            const int cand = k + step;
            if (incs[cand - 1] <= local) k = cand;
          }
          int32_t item = w * 32 + k;
          int32_t pos = local - (k > 0 ? incs[k - 1] : 0);
          int32_t row_lo = S.it_lo[item], row_len = S.it_len[item];
          if (CLOCKS && row_len >= 0) KLF_FINE(13);
          const int32_t mine = min(K, total - g);
          for (int j0 = 0; j0 < mine; j0 += 4) {
            int32_t e4[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              e4[jj] = -1;
              if (j0 + jj < mine) {
                e4[jj] = row_lo + pos;
                if (++pos == row_len) { ++item; row_lo = S.it_lo[item]; row_len = S.it_len[item]; pos = 0; }
              }
            }
            int32_t c4[4];
            float w4[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              c4[jj] = -1; w4[jj] = 0.0f;
              if (e4[jj] >= 0) { c4[jj] = __ldg(p.col + e4[jj]); w4[jj] = __ldg(p.w + e4[jj]); }
            }
            unsigned s4[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              s4[jj] = 0u;
              if (c4[jj] >= 0) {
                if (c4[jj] == a) s4[jj] = 1u;                          // the pair being swapped: sides after the swap
                else if (c4[jj] == b) s4[jj] = 0u;
                else s4[jj] = GBITS ? ((unsigned)p.state[c4[jj]] & ST_SIDE) : (bits_get(bits, c4[jj]) & ST_SIDE);
              }
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              if (c4[jj] >= 0) stg[g + j0 + jj] = s4[jj] ? w4[jj] : -w4[jj];
          }
        }
        KLF_FINE(4);
        klf_bar_rows();                                  // every signed weight of this swap is staged
        KLF_FINE(5);
        // ---- SUMS: the two ordered sums of cKL.cpp:225-251 -- E over the external weights, I over the internal ones, row
        //      order -- then the key.  A 64-bit shared-memory atomicMax is a compare-and-swap loop (~350 cycles measured,
        //      tools/micro/prims.cu); the keys are raised in two passes of native 32-bit atomics instead: the D words
        //      first (a raiser clears the id word), then, behind a barrier, the id words of whoever holds the final D ----
        bool pub = false;
        uint32_t khi = 0u, klo = 0u;
        unsigned long long *kp = nullptr, *gp = nullptr;
        if (tid < items) {
          my_st = GBITS ? (my_st & 3u) : bits_get(bits, my_v);
          float E = 0.0f, I = 0.0f;
          const float *src = stg + my_base + (inc - my_len);
          // batches of 8, the next batch's loads in flight while this one is added; a batch is padded with +0.0f, which
          // changes neither sum (E, I >= +0)
          float cur[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) cur[q] = q < my_len ? src[q] : 0.0f;
          for (int t = 0; t < my_len; t += 8) {
            float nxt[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) nxt[q] = (t + 8 + q) < my_len ? src[t + 8 + q] : 0.0f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              E = __fadd_rn(E, fmaxf(cur[q], 0.0f));
              I = __fadd_rn(I, fmaxf(-cur[q], 0.0f));
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) cur[q] = nxt[q];
          }
          const float nv = __fsub_rn(E, I);
          if (CLOCKS && nv != 1e30f) KLF_FINE(15);
          p.val[my_v] = nv;
          if (!((my_st & ST_LOCK) || my_v == a || my_v == b)) {
            const unsigned sd = my_st & ST_SIDE;
            khi = float_orderable(sd ? -nv : nv);
            klo = 0xFFFFFFFFu - my_id;
            const int32_t tile = my_v / KL_TILE;
            if (tile == ta || tile == tb) {               // the early warp of that tile folds this in after the barrier
              kp = &S.pkey[par][tile == ta ? 0 : 1][sd];
              atomicMax(key_hi(kp), khi);
              pub = true;
            } else {
              kp = keys + 2 * tile + sd;
              const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(kp);
              const unsigned long long nk = ((unsigned long long)khi << 32) | klo;
              if (cur != 0ull && (uint32_t)(cur & 0xFFFFFFFFull) == klo && nk < cur) {
                // v held the tile's best key and lost ground: the tile is rescanned
                if (atomicExch(stamps + tile, stamp) != stamp) {
                  const int slot = atomicAdd(&S.nlist[par], 1);
                  if (slot < KLF_LCAP) S.list[par][slot] = tile;
                }
              } else {
                pub = true;
                if (!ONE) gp = gkeys + 2 * (tile / KL_GROUP) + sd;
                if (atomicMax(key_hi(kp), khi) < khi) {
                  *reinterpret_cast<volatile uint32_t *>(key_lo(kp)) = 0u;
                  if (!ONE && atomicMax(key_hi(gp), khi) < khi) *reinterpret_cast<volatile uint32_t *>(key_lo(gp)) = 0u;
                }
              }
            }
          }
        }
        klf_bar_rows();                                  // every D word is final
        if (pub) {
          if (*reinterpret_cast<volatile uint32_t *>(key_hi(kp)) == khi) atomicMax(key_lo(kp), klo);
          if (gp != nullptr && *reinterpret_cast<volatile uint32_t *>(key_hi(gp)) == khi) atomicMax(key_lo(gp), klo);
        }
        KLF_FINE(6);
      } else {
        // more entries than the staging buffer holds (a hub among the neighbours): warp-per-row replay over the item list
        KlfCtx X;
        make_ctx(X);
        float *wsm = S.stage + warp * 64;
        for (int32_t it = warp; it < items; it += KLF_ROW_WARPS) {
          const int32_t e = it < da ? alo + it : blo + (it - da);
          const int32_t v = __ldg(p.col + e);
          const int2 ext = __ldg(p.nb + e);
          const uint32_t vid = ASC ? (uint32_t)v : __ldg(p.rank + v);
          klf_row_replay<ASC, GBITS>(p, X, bits, v, vid, ext.x, ext.y, wsm);
        }
      }
    } else {
      // ---- more neighbours than row threads (industry2-class hubs): exclusion marks, then warp-per-row over the list ----
      for (int32_t it = tid; it < items; it += KLF_ROW_THREADS) {
        const int32_t e = it < da ? alo + it : blo + (it - da);
        const int32_t v = __ldg(p.col + e);
        const int32_t tv = v / KL_TILE;
        if (tv == ta) atomicOr(&S.excl[par][0][(v % KL_TILE) >> 5], 1u << (v & 31));
        else if (tv == tb) atomicOr(&S.excl[par][1][(v % KL_TILE) >> 5], 1u << (v & 31));
      }
      klf_bar_items();
      KlfCtx X;
      make_ctx(X);
      float *wsm = S.stage + warp * 64;
      for (int32_t it = warp; it < items; it += KLF_ROW_WARPS) {
        const int32_t e = it < da ? alo + it : blo + (it - da);
        const int32_t v = __ldg(p.col + e);
        const int2 ext = __ldg(p.nb + e);
        const uint32_t vid = ASC ? (uint32_t)v : __ldg(p.rank + v);
        klf_row_replay<ASC, GBITS>(p, X, bits, v, vid, ext.x, ext.y, wsm);
      }
    }
    KLF_FINE(7);
    __syncthreads();                                   // (A) every D-value, published key and rescan request of this swap is in
    KLF_FINE(8);
    // ---- S4: the early tiles' keys = max(base, published); group keys; late rescans (rare) ----
    const int nl = S.nlist[par];
    const int32_t gA = ta / KL_GROUP, gB = tb / KL_GROUP;
    if (early && lane == 0) {
      const unsigned long long q0 = S.pkey[par][es][0], q1 = S.pkey[par][es][1];
      keys[2 * et] = base0 > q0 ? base0 : q0;
      keys[2 * et + 1] = base1 > q1 ? base1 : q1;
    }
    if (nl == 0) {
      // common case: only the groups of the two early tiles changed downwards; their warps refold them
      if (warp == KLF_E1 && early) {
        __syncwarp();
        if (!ONE && gB != gA) group_fold(gB);
        klf_bar_early_arrive();
      } else if (warp == KLF_E0) {
        __syncwarp();
        if (ONE) {
          if (tb != ta) klf_bar_early_sync();
        } else {
          if (tb != ta && gB == gA) klf_bar_early_sync();
          group_fold(gA);
          if (tb != ta && gB != gA) klf_bar_early_sync();
        }
        __syncwarp();
        // the next swap's pair selection, while the other warps are on their way to the barrier
        unsigned long long k0 = 0ull, k1 = 0ull;
#pragma unroll 1
        for (int32_t g = lane; g < n_sel; g += 32) {
          const unsigned long long a0 = sel[2 * g], a1 = sel[2 * g + 1];
          k0 = a0 > k0 ? a0 : k0;
          k1 = a1 > k1 ? a1 : k1;
        }
        k0 = warp_max_u64(k0); k1 = warp_max_u64(k1);
        if (lane == 0) { S.best[0] = k0; S.best[1] = k1; S.best_stamp = stamp; }
      }
    } else if (nl > KLF_LCAP) {
      // more rescan requests than the list holds (only a hub swap can do that): rebuild every key
      for (int32_t t = warp; t < p.n_tiles; t += KLF_WARPS) klf_tile_rescan<ASC, GBITS>(p, bits, t, keys);
      __syncthreads();
      if (!ONE)
        for (int32_t g = warp; g < n_groups; g += KLF_WARPS) group_fold(g);
    } else {
      if (early) {
        if (!ONE && lane == 0) {
          const int32_t g = et / KL_GROUP;
          if (atomicExch(gstamp + g, stamp) != stamp) S.dlist[par][atomicAdd(&S.ndirty[par], 1)] = g;
        }
      } else if (warp < KLF_ROW_WARPS) {
        for (int q = warp; q < nl; q += KLF_ROW_WARPS) {
          const int32_t t = S.list[par][q];
          klf_tile_rescan<ASC, GBITS>(p, bits, t, keys);
          if (!ONE && lane == 0) {
            const int32_t g = t / KL_GROUP;
            if (atomicExch(gstamp + g, stamp) != stamp) S.dlist[par][atomicAdd(&S.ndirty[par], 1)] = g;
          }
        }
      }
      __syncthreads();                                 // (B) tile keys final
      if (!ONE) {
        const int nd = S.ndirty[par];
        for (int q = warp; q < nd; q += KLF_WARPS) group_fold(S.dlist[par][q]);
      }
    }
    if (CLOCKS && tid == 0) S.fine[14] += nl;
    KLF_FINE(9);
    __syncthreads();                                   // (C) group keys final
    KLF_FINE(10);
  }
  __syncthreads();
  if (tid == 0) {
    p.ctrl[0] = (int64_t)S.iter; p.ctrl[1] = 1;
    if (CLOCKS)
      for (int i = 0; i < 16; ++i) p.ctrl[16 + i] = S.fine[i];
  }
#undef KLF_FINE
}

// Best-prefix rollback (the classic KL pass ends by keeping only the swaps up to the best cut -- the
// reference tracks minCutSize, cKL.cpp:363, but never rolls back nor saves the partition, cKL.cpp:395-405): swaps
// best+1 .. swaps of the last pass are undone on the device.  Returns the kept row (first minimum of the cut column).
__global__ void kl_undo_kernel(const int32_t *__restrict__ n1, const int32_t *__restrict__ n2, int64_t from, int64_t to,
                               uint8_t *__restrict__ state) {
  const int64_t i = from + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > to) return;
  state[n1[i]] = (uint8_t)ST_LOCK;                   // back on side 0 (still marked: the pass is over)
  state[n2[i]] = (uint8_t)(ST_SIDE | ST_LOCK);       // back on side 1
}
int64_t kl_rollback(eigkl_handle *h, float *best_cut) {
  auto &k = h->kl;
  EIGKL_REQUIRE(k.have_partition && k.consumed, EIGKL_E_ARG, "eigkl_kl_rollback: no finished KL pass to roll back");
  std::vector<float> cut((size_t)k.swaps + 1);
  EIGKL_CUDA(cudaMemcpyAsync(cut.data(), k.t_cut.p, cut.size() * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  EIGKL_CUDA(cudaStreamSynchronize(h->stream));
  int64_t best = 0;
  for (int64_t i = 1; i <= k.swaps; ++i)
    if (cut[(size_t)i] < cut[(size_t)best]) best = i;                 // first minimum
  if (best < k.swaps) {
    kl_undo_kernel<<<grid_for(k.swaps - best), TPB, 0, h->stream>>>(k.t_n1.p, k.t_n2.p, best + 1, k.swaps, k.state.p);
    h->launches++;
    EIGKL_CUDA(cudaGetLastError());
  }
  k.swaps = best;                                                    // a second rollback is a no-op
  if (best_cut) *best_cut = cut[(size_t)best];
  return best;
}

void kl_run(eigkl_handle *h) {
  auto &A = h->A;
  auto &k = h->kl;
  EIGKL_REQUIRE(A.valid, EIGKL_E_ARG, "eigkl_kl_run: call eigkl_assemble_kl_graph first");
  EIGKL_REQUIRE(k.have_partition, EIGKL_E_ARG, "eigkl_kl_run: no initial partition");
  kl_rearm(h);
  const int32_t n = A.n;
  cudaStream_t st = h->stream;
  const int64_t cap = std::min(k.n0, k.n1) + 1;
  k.t_cut.ensure((size_t)cap); k.t_gain.ensure((size_t)cap); k.t_n1.ensure((size_t)cap); k.t_n2.ensure((size_t)cap);
  k.cap = cap;
  // set-up: initial cut (row 0), D-values of every node, tile keys
  h->timer.start(st);
  const float cut0 = kl_cut0(h);                                     // cKL.cpp:306
  kl_dvalues(h);                                                     // cKL.cpp:318-321
  const int32_t n_tiles = (int32_t)ceil_div(n, KL_TILE);
  // Several ranks: the pass is latency-bound and sequential (one CTA at 4.4 us per swap), so every rank runs the
  // complete pass on its own replica -- bit-identical by construction.  EIGKL_KL_DIST=1 keeps the partitioned
  // variant (D-values / tile keys by node range, one NCCL arg-max all-reduce per swap: 24 us per swap measured).
  const int R = h->kl_dist ? h->opts.nranks : 1;
  int32_t own_lo = 0, own_hi = n, own_pad = 0;
  if (R > 1) row_partition(n, R, h->opts.rank, &own_lo, &own_hi, &own_pad);
  // state in shared memory (one CTA) whenever it fits, unless a cluster size was asked for explicitly
  const bool local = R == 1 && h->kl_local && h->opts.kl_cluster <= 0 && n <= KL_LOCAL_GBITS_MAX_N;
  const bool gbits = local && (n > KL_LOCAL_MAX_N || getenv("EIGKL_KL_GBITS") != nullptr);
  const bool flat = local && h->kl_flat;
  if (!local) {
    tile_init_kernel<<<grid_for((int64_t)n_tiles * 32), TPB, 0, st>>>(k.state.p, k.val.p, k.rank.p, own_lo, own_hi, n_tiles, k.tile_key.p, k.tile_stamp.p);
    h->launches++;
  } else if (!A.nb_valid) {
    A.nb.alloc((size_t)2 * A.nnz + 2);
    nb_extent_kernel<<<grid_for(A.nnz), TPB, 0, st>>>(A.rowptr.p, A.col.p, A.nnz, reinterpret_cast<int2 *>(A.nb.p));
    h->launches++;
    A.nb_valid = true;
  }
  const float zero = 0.0f; const int32_t neg = -1;
  EIGKL_CUDA(cudaMemcpyAsync(k.t_cut.p, &cut0, sizeof(float), cudaMemcpyHostToDevice, st));
  EIGKL_CUDA(cudaMemcpyAsync(k.t_gain.p, &zero, sizeof(float), cudaMemcpyHostToDevice, st));
  EIGKL_CUDA(cudaMemcpyAsync(k.t_n1.p, &neg, sizeof(int32_t), cudaMemcpyHostToDevice, st));
  EIGKL_CUDA(cudaMemcpyAsync(k.t_n2.p, &neg, sizeof(int32_t), cudaMemcpyHostToDevice, st));
  EIGKL_CUDA(cudaMemsetAsync(k.ctrl.p, 0, 4 * sizeof(int64_t), st));
  EIGKL_CUDA(cudaMemsetAsync(k.ctrl.p + 40, 0, 8 * sizeof(int64_t), st));
  h->timer.stop(st);
  h->stats.ms_kl_setup = h->timer.ms();

  KlLoopParams p;
  p.n = n; p.n_tiles = n_tiles;
  p.rowptr = A.rowptr.p; p.col = A.col.p; p.w = A.w.p;
  p.state = k.state.p; p.rank = k.rank.p; p.val = k.val.p;
  p.tile_key = k.tile_key.p; p.tile_stamp = k.tile_stamp.p;
  p.order0 = k.order0.p; p.order1 = k.order1.p;
  p.t_cut = k.t_cut.p; p.t_gain = k.t_gain.p; p.t_n1 = k.t_n1.p; p.t_n2 = k.t_n2.p;
  p.ctrl = k.ctrl.p;
  p.cut0 = cut0;
  p.term_limit = (uint32_t)std::log2((double)n) + 5;                 // cKL.cpp:303
  p.n0 = k.n0; p.n1 = k.n1;
  p.own_lo = own_lo; p.own_hi = own_hi;
  p.mctrl = nullptr;

  int nc = h->opts.kl_cluster;
  // measured (tools/kl_sweep.py): the swap loop is a chain of ~10 dependent L2 round trips, so up to
  // ibm10's size one CTA (block barriers only) beats any cluster (11.6 vs 13.6 us/swap); the cluster only
  // pays once the per-swap tile scan is large
  if (nc <= 0) nc = (n <= 131072) ? 1 : 8;
  EIGKL_REQUIRE(nc == 1 || nc == 2 || nc == 4 || nc == 8 || nc == 16, EIGKL_E_ARG, "kl_cluster must be 1, 2, 4, 8 or 16");
  if (nc > 8) {
    EIGKL_CUDA(cudaFuncSetAttribute(kl_loop_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    EIGKL_CUDA(cudaFuncSetAttribute(kl_loop_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)nc);
  cfg.blockDim = dim3(KL_LOOP_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)nc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  h->timer.start(st);
  if (local) {
    KlLocalParams q;
    q.n = n; q.n_tiles = n_tiles;
    q.rowptr = A.rowptr.p; q.col = A.col.p; q.nb = reinterpret_cast<const int2 *>(A.nb.p); q.w = A.w.p;
    q.state = k.state.p; q.rank = k.rank.p; q.val = k.val.p; q.order0 = k.order0.p; q.order1 = k.order1.p;
    q.t_cut = k.t_cut.p; q.t_gain = k.t_gain.p; q.t_n1 = k.t_n1.p; q.t_n2 = k.t_n2.p;
    q.ctrl = k.ctrl.p; q.cut0 = cut0; q.term_limit = p.term_limit; q.n0 = k.n0; q.n1 = k.n1;
    q.clocks = getenv("EIGKL_KL_PHASES") != nullptr ? 1 : 0;
    const size_t smem = (size_t)n_tiles * (16 + 4 + 4) + (gbits ? 0 : (size_t)((n + 15) / 16) * 4) + 16;
    if (!h->attr_kl_local) {
      const size_t max_smem = std::max((size_t)(KL_LOCAL_MAX_N / KL_TILE) * 24 + (size_t)(KL_LOCAL_MAX_N / 16) * 4,
                                       (size_t)(KL_LOCAL_GBITS_MAX_N / KL_TILE) * 24) + 16;
      EIGKL_CUDA(cudaFuncSetAttribute(kl_loop_local_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
      EIGKL_CUDA(cudaFuncSetAttribute(kl_loop_local_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
      EIGKL_CUDA(cudaFuncSetAttribute(kl_loop_local_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
      EIGKL_CUDA(cudaFuncSetAttribute(kl_loop_local_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
      h->attr_kl_local = true;
    }
    nc = 1;
    if (flat) {
      const size_t n_groups = (size_t)ceil_div(n_tiles, KL_GROUP);
      const size_t fsmem = (size_t)n_tiles * 20 + n_groups * 20 + sizeof(KlFlatSmem) + (gbits ? 0 : (size_t)((n + 15) / 16) * 4) + 16;
      const bool one = !gbits && n_tiles <= 64;       // the selection scans the tile keys directly (measured: 3 % on ibm01; at 272 tiles already 5 % slower)
      const int variant = (one ? 8 : 0) | (k.ascending ? 4 : 0) | (gbits ? 2 : 0) | (q.clocks ? 1 : 0);
      auto launch_flat = [&](auto kern) {
        if (!(h->attr_kl_flat & (1u << variant))) {
          EIGKL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          h->attr_kl_flat |= 1u << variant;
        }
        kern<<<1, KLF_THREADS, fsmem, st>>>(q);       // (147 sleeping "company" CTAs were tried against the lone-CTA issue throttle: no effect)
      };
      EIGKL_REQUIRE(fsmem <= 227 * 1024, EIGKL_E_ARG, "KL flat loop: shared-memory plan exceeds the SM");
      switch (variant) {
        case 0: launch_flat(kl_loop_flat_kernel<false, false, false, false>); break;
        case 1: launch_flat(kl_loop_flat_kernel<false, false, true, false>); break;
        case 2: launch_flat(kl_loop_flat_kernel<false, true, false, false>); break;
        case 3: launch_flat(kl_loop_flat_kernel<false, true, true, false>); break;
        case 4: launch_flat(kl_loop_flat_kernel<true, false, false, false>); break;
        case 5: launch_flat(kl_loop_flat_kernel<true, false, true, false>); break;
        case 6: launch_flat(kl_loop_flat_kernel<true, true, false, false>); break;
        case 7: launch_flat(kl_loop_flat_kernel<true, true, true, false>); break;
        case 8: launch_flat(kl_loop_flat_kernel<false, false, false, true>); break;
        case 9: launch_flat(kl_loop_flat_kernel<false, false, true, true>); break;
        case 12: launch_flat(kl_loop_flat_kernel<true, false, false, true>); break;
        default: launch_flat(kl_loop_flat_kernel<true, false, true, true>); break;
      }
    } else if (k.ascending) {
      if (gbits) kl_loop_local_kernel<true, true><<<1, KL_LOOP_THREADS, smem, st>>>(q);
      else kl_loop_local_kernel<true, false><<<1, KL_LOOP_THREADS, smem, st>>>(q);
    } else {
      if (gbits) kl_loop_local_kernel<false, true><<<1, KL_LOOP_THREADS, smem, st>>>(q);
      else kl_loop_local_kernel<false, false><<<1, KL_LOOP_THREADS, smem, st>>>(q);
    }
    EIGKL_CUDA(cudaGetLastError());
    h->launches++;
  } else if (R == 1) {
    EIGKL_CUDA(cudaLaunchKernelEx(&cfg, kl_loop_kernel<false>, p));
    h->launches++;
  } else {
    // multi-rank: D-values and tile keys are partitioned by node range, the side
    // bytes are replicated; per swap: local arg-max -> ncclAllReduce(max, 2 x uint64 packed keys) ->
    // every rank applies the swap to its replica and recomputes the D-values it owns.  Swaps are
    // issued in batches; the done flag is read back once per batch.
    k.mctrl.ensure(8);
    KlCtrl init{cut0, 0u, 0u, (k.n0 <= 0 || k.n1 <= 0) ? 1 : 0, (long long)k.n0, (long long)k.n1};
    EIGKL_CUDA(cudaMemcpyAsync(k.mctrl.p, &init, sizeof(init), cudaMemcpyHostToDevice, st));
    p.mctrl = reinterpret_cast<KlCtrl *>(k.mctrl.p);
    unsigned long long *exch = k.tile_key.p + 2 * (size_t)n_tiles;
    const int32_t tile_lo = own_lo / KL_TILE, tile_hi = (int32_t)ceil_div(own_hi, KL_TILE);
    const int batch = 64;
    const int64_t max_swaps = std::min(k.n0, k.n1);
    int32_t done = init.done;
    for (int64_t issued = 0; !done && issued < max_swaps + batch; issued += batch) {
      for (int i = 0; i < batch; ++i) {
        kl_select_kernel<<<1, KL_LOOP_THREADS, 0, st>>>(k.tile_key.p, tile_lo, own_hi > own_lo ? tile_hi : tile_lo, exch, p.mctrl);
        comm_allreduce_max_u64(h, exch, 2);
        EIGKL_CUDA(cudaLaunchKernelEx(&cfg, kl_loop_kernel<true>, p));
        h->launches += 2;
      }
      EIGKL_CUDA(cudaMemcpyAsync(&done, reinterpret_cast<char *>(k.mctrl.p) + offsetof(KlCtrl, done), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      EIGKL_CUDA(cudaStreamSynchronize(st));
    }
  }
  h->timer.stop(st);
  int64_t ctrl[48] = {0};
  EIGKL_CUDA(cudaMemcpyAsync(ctrl, k.ctrl.p, sizeof(ctrl), cudaMemcpyDeviceToHost, st));
  EIGKL_CUDA(cudaStreamSynchronize(st));
  EIGKL_CUDA(cudaGetLastError());
  h->stats.ms_kl_loop = h->timer.ms();
  EIGKL_REQUIRE(ctrl[1] == 1, EIGKL_E_CUDA, "KL kernel did not complete");
  k.swaps = ctrl[0];
  k.consumed = true;
  h->stats.kl_swaps = k.swaps;
  if (getenv("EIGKL_KL_PHASES") && R == 1 && k.swaps > 0) {
    static const char *nm[6] = {"S1 reduce", "decode + row pointers", "S3 own rows", "S3 barrier wait", "S4 own tiles", "S4 barrier wait"};
    fprintf(stderr, "[eigkl] KL phases, cycles per swap (thread 0):");
    for (int i = 0; i < 6; ++i) fprintf(stderr, " %s %.0f;", nm[i], (double)ctrl[8 + i] / (double)k.swaps);
    fprintf(stderr, "\n");
    if (local && flat) {
      static const char *fn[11] = {"S1 (pair selection)", "decode+rowptr (+early loads issued)", "item loads + scan", "items barrier", "entry search + loads + stage",
                                   "entries barrier", "row sums + two-pass publish", "S3 tail", "barrier A", "early keys + group refold (+ late rescans)", "barrier C"};
      fprintf(stderr, "[eigkl] KL flat-loop probes, cycles per swap (thread 0):");
      for (int i = 0; i < 11; ++i) fprintf(stderr, " %s %.0f;", fn[i], (double)ctrl[16 + i] / (double)k.swaps);
      fprintf(stderr, " [sub-probes: item loads arrived %.0f; scan done %.0f; entry search done %.0f; sums done %.0f]", (double)ctrl[16 + 11] / (double)k.swaps,
              (double)ctrl[16 + 12] / (double)k.swaps, (double)ctrl[16 + 13] / (double)k.swaps, (double)ctrl[16 + 15] / (double)k.swaps);
      fprintf(stderr, " late rescans per swap %.2f\n", (double)ctrl[16 + 14] / (double)k.swaps);
    } else if (local) {
      static const char *fn[13] = {"S1 local max", "S1 barrier 1", "S1 final max", "S1 barrier 2", "decode+rowptr", "nbr entries arrive",
                                   "row loads arrive (per batch)", "row sums (all rows)", "row epilogues (all rows)", "S3 tail", "S3 barrier",
                                   "S4 rescans", "S4 barrier"};
      fprintf(stderr, "[eigkl] KL fine probes, cycles per swap (thread 0):");
      for (int i = 0; i < 13; ++i) fprintf(stderr, " %s %.0f;", fn[i], (double)ctrl[16 + i] / (double)k.swaps);
      fprintf(stderr, " tiles rescanned per swap %.2f\n", (double)ctrl[16 + 14] / (double)k.swaps);
    }
  }
  h->stats.kl_cluster = nc;
  h->stats.kl_threads = nc * KL_LOOP_THREADS;
  h->stats.kl_local = local ? (gbits ? 2 : 1) : 0;
  h->stats.kl_flat = flat ? 1 : 0;
  h->stats.kl_threads = flat ? KLF_THREADS : nc * KL_LOOP_THREADS;
}

}  // namespace eigkl
