// spmv.cu -- fp64 CSR SpMV y = L x for the Lanczos solver (north-star subsystem 2).
//
// Replaces Spectra's SparseSymMatProd (cEIG.cpp:194, serial Eigen product) and is the B200 answer to
// the reference's only GPU SpMV, sparseMVKernel (gKL2.cu:65-89: thread-per-row, scalar, fp32).
//
// Three kernels, chosen per matrix at assembly:
//   cheb_resident_kernel  (single rank, the matrix fits the chip's registers + shared memory): a whole filter
//                         application, d SpMVs, per launch -- see "Resident polynomial filter" below;
//   spmv_flat_kernel      one launch per SpMV, the whole row block in one round of loads (default otherwise);
//   spmv_adaptive_kernel  one launch per SpMV, lanes-per-row chosen per block (EIGKL_SPMV_MODE = 1 / 2).
// Layout: CSR with int32 rowptr/col and fp64 values, rows cut into row blocks of ~SPMV_CHUNK
// non-zeros (blk_row, built at assembly).  spmv_adaptive_kernel, one CTA per row block:
//   * stream mode (very short rows, mean < 4 entries): all threads stream val/col fully coalesced,
//     gather x through the read-only path, stage the products in shared memory, then 1..32 threads
//     per row reduce each row from shared memory;
//   * vector mode: 2..32 lanes per row chosen from the block's mean row length -- a full warp per row
//     for high-degree rows (industry2's 100..900-entry rows) -- lanes stride the row, shuffle reduction.
//     (ncu on ibm10: the staged variant stalls on mio_throttle/barrier and its 32 KB/CTA of shared memory
//     leaves no L1 for the x gathers, which are 70% of the L2 sector traffic.)
// Fused epilogue for the (Chebyshev-filtered) Lanczos recurrence:
//     y[r] = ca * s * (L x)[r] + cb * s * x[r] + cg * z[r],   s = *scale (1/beta of the previous step)
// and, optionally, v_store[r] = s * x[r], so the three-term recurrence T_{k+1} = 2 t T_k - T_{k-1} and
// the normalisation v_j = w/beta are applied by the SpMV itself (no separate axpy/scale kernels).
// Bound: HBM (or L2 when the matrix fits the 126 MB L2): nnz*12 + n*20 bytes per launch.
#include "internal.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "device_utils.cuh"

namespace eigkl {

constexpr int SPMV_THREADS = 256;
constexpr int SPMV_STAGE = 4096;        // staging capacity in products (32 KB)

struct SpmvEpilogue {
  const double *xl;      // this rank's slice of x (row r at xl[r - off])
  const double *z;       // slice of the vector two steps back in the recurrence (may be null)
  double *y;             // output slice
  double ca, cb, cg;     // already multiplied by *scale where the formula asks for it
  int32_t off;
  // the row's own x and z values: loaded when the row starts, so that they are not one more dependent
  // round trip after the reduction
  __device__ __forceinline__ double prefetch(int32_t r) const {
    double t = 0.0;
    if (cb != 0.0) t = cb * xl[r - off];
    if (z) t += cg * z[r - off];
    return t;
  }
  __device__ __forceinline__ void emit(int32_t r, double lx, double pre) const { y[r - off] = ca * lx + pre; }
  __device__ __forceinline__ void emit(int32_t r, double lx) const { emit(r, lx, prefetch(r)); }
};

// L lanes cooperate on one row: lanes stride the row (coalesced across the sub-warp and, because
// consecutive sub-warps own consecutive rows, across the warp), then a shuffle reduction of width L.
// Rows much longer than the block's mean (a 574-entry row among 20-entry rows in ibm10; the 900-entry rows
// of industry2) would serialise dozens of dependent load rounds on a few lanes, so they are set aside in a
// shared-memory list and handled afterwards by whole warps with 4 independent loads per lane in flight
// (measured on industry2: 26.7 -> 14.4 us per SpMV; unrolling the sub-warp loop 4-fold in place instead
// gave 28.7 us).
constexpr int SPMV_MAX_LONG = 64;
template <int L>
__device__ __forceinline__ void rows_subwarp(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                             const double *__restrict__ val, const double *__restrict__ x,
                                             const SpmvEpilogue &ep, int32_t r0, int32_t r1, int tid,
                                             int32_t *long_rows, int *n_long) {
  constexpr int ROWS_PER_ITER = SPMV_THREADS / L;
  constexpr int LONG_LEN = (L >= 32) ? 512 : 8 * L;
  const int sub = tid & (L - 1);
  if (tid == 0) *n_long = 0;
  __syncthreads();
  for (int32_t base = r0; base < r1; base += ROWS_PER_ITER) {      // block-uniform trip count
    const int32_t r = base + tid / L;
    double s = 0.0, pre = 0.0;
    bool deferred = false;
    if (r < r1) {
      const int32_t lo = rowptr[r], hi = rowptr[r + 1];
      if (sub == 0) pre = ep.prefetch(r);
      if (hi - lo > LONG_LEN) {
        // only the lanes of this row's sub-warp are guaranteed to be here
        const unsigned sub_mask = (L >= 32) ? FULL_MASK : (((1u << L) - 1u) << ((tid & 31) & ~(L - 1)));
        int slot = SPMV_MAX_LONG;
        if (sub == 0) slot = atomicAdd(n_long, 1);
        slot = __shfl_sync(sub_mask, slot, (tid & 31) & ~(L - 1));
        if (slot < SPMV_MAX_LONG) {
          if (sub == 0) long_rows[slot] = r;
          deferred = true;
        }
      }
      if (!deferred) {
        int32_t i = lo + sub;
        for (; i + L < hi; i += 2 * L) {                           // two independent gathers in flight
          const double a0 = __ldcs(val + i), a1 = __ldcs(val + i + L);          // matrix: streamed (evict-first),
          const double x0 = __ldg(x + __ldcs(col + i)), x1 = __ldg(x + __ldcs(col + i + L));   // x: kept in L1
          s += a0 * x0;
          s += a1 * x1;
        }
        if (i < hi) s += __ldcs(val + i) * __ldg(x + __ldcs(col + i));
      }
    }
#pragma unroll
    for (int o = L >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
    if (r < r1 && sub == 0 && !deferred) ep.emit(r, s, pre);
  }
  __syncthreads();
  const int nl = min(*n_long, SPMV_MAX_LONG);
  const int lane = tid & 31, warp = tid >> 5;
  for (int q = warp; q < nl; q += SPMV_THREADS / 32) {             // one warp per long row
    const int32_t r = long_rows[q];
    const int32_t lo = rowptr[r], hi = rowptr[r + 1];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int32_t i = lo + lane;
    for (; i + 96 < hi; i += 128) {
      const double a0 = __ldcs(val + i), a1 = __ldcs(val + i + 32), a2 = __ldcs(val + i + 64), a3 = __ldcs(val + i + 96);
      const int32_t c0 = __ldcs(col + i), c1 = __ldcs(col + i + 32), c2 = __ldcs(col + i + 64), c3 = __ldcs(col + i + 96);
      s0 += a0 * __ldg(x + c0); s1 += a1 * __ldg(x + c1); s2 += a2 * __ldg(x + c2); s3 += a3 * __ldg(x + c3);
    }
    for (; i < hi; i += 32) s0 += __ldcs(val + i) * __ldg(x + __ldcs(col + i));
    const double s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) ep.emit(r, s);
  }
}

// WITH_STREAM: the kernel variant that carries the 32 KB staging buffer (launched only when the matrix
// has row blocks short enough to want it, so that vector-only matrices keep their L1 for the x gathers).
// mode: 0 = choose per row block, 1 = always stream (shared-memory staged), 2 = always sub-warp/warp per row
template <bool WITH_STREAM>
__global__ void __launch_bounds__(SPMV_THREADS, WITH_STREAM ? 4 : 6)    // 6 CTAs/SM (<= 42 registers): the row blocks are sized for ONE wave
spmv_adaptive_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                     const double *__restrict__ val, const double *__restrict__ x, const double *__restrict__ xl,
                     const double *__restrict__ z, double *__restrict__ y,
                     const int32_t *__restrict__ blk_row, const double *__restrict__ scale,
                     double *__restrict__ v_store, int32_t row_offset, int mode, double ca, double cb, double cg) {
  __shared__ double prod[WITH_STREAM ? SPMV_STAGE : 1];
  __shared__ int32_t long_rows[SPMV_MAX_LONG];
  __shared__ int n_long;
  const int tid = threadIdx.x;
  const int32_t r0 = blk_row[blockIdx.x], r1 = blk_row[blockIdx.x + 1];
  if (r0 >= r1) return;
  const double sc = scale ? __ldg(scale) : 1.0;
  const int32_t e0 = rowptr[r0], e1 = rowptr[r1];
  const int32_t span = e1 - e0, nrows = r1 - r0;
  // rows are global ids; xl, z, y and v_store are this rank's slices
  const SpmvEpilogue ep{xl, z, y, ca * sc, cb * sc, cg, row_offset};
  if (v_store) {
    for (int32_t r = r0 + tid; r < r1; r += SPMV_THREADS) v_store[r - row_offset] = xl[r - row_offset] * sc;
  }
  int mean = span / nrows;
  if (mode >= 100) mean = mode - 100;    // tuning aid: EIGKL_SPMV_MODE = 100 + forced mean row length
  const bool stream = WITH_STREAM && ((mode == 1) || (mode == 0 && mean < 4));
  if (stream && span <= SPMV_STAGE) {
#pragma unroll 4
    for (int32_t i = tid; i < span; i += SPMV_THREADS) prod[i] = val[e0 + i] * __ldg(x + col[e0 + i]);
    __syncthreads();
    int tpr = 1;                                     // threads per row, power of two <= 32
    while (tpr < 32 && nrows * tpr * 2 <= SPMV_THREADS) tpr <<= 1;
    const int rows_per_iter = SPMV_THREADS / tpr;
    const int sub = tid & (tpr - 1);
    for (int32_t base = 0; base < nrows; base += rows_per_iter) {   // warp-uniform trip count
      const int32_t rr = base + tid / tpr;
      double s = 0.0;
      if (rr < nrows) {
        const int32_t lo = rowptr[r0 + rr] - e0, hi = rowptr[r0 + rr + 1] - e0;
        for (int32_t i = lo + sub; i < hi; i += tpr) s += prod[i];
      }
      for (int o = tpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
      if (rr < nrows && sub == 0) ep.emit(r0 + rr, s);
    }
  } else if (mean <= 4) {
    rows_subwarp<2>(rowptr, col, val, x, ep, r0, r1, tid, long_rows, &n_long);
  } else if (mean <= 8) {
    rows_subwarp<4>(rowptr, col, val, x, ep, r0, r1, tid, long_rows, &n_long);
  } else if (mean <= 16) {
    rows_subwarp<8>(rowptr, col, val, x, ep, r0, r1, tid, long_rows, &n_long);
  } else if (mean <= 48) {
    rows_subwarp<16>(rowptr, col, val, x, ep, r0, r1, tid, long_rows, &n_long);
  } else {
    rows_subwarp<32>(rowptr, col, val, x, ep, r0, r1, tid, long_rows, &n_long);   // warp per row (industry2-class rows)
  }
}

// ---------------------------------------------------------------------------------------------------
// "flat" variant: the whole row block in ONE round of loads.  Dependent round trips, not bytes, set the time
// on the L2-resident circuits (ncu: no unit above 40 %, long_scoreboard stalls; the sub-warp kernel walks a
// block in ~5 serial row iterations of 3-4 dependent loads each).  Here a CTA reads its block descriptor
// (one int4: rows and entry range, precomputed), then issues ALL its loads at once -- up to FLAT_K
// (value, column) pairs per thread, the block's row pointers into shared memory, the epilogue operands --
// then all x gathers, stages the products in shared memory and lets one thread per row add its run.
// Three dependent round trips per SpMV.  A block whose span exceeds the staging capacity (a row of
// > ~2000 entries) falls back to warp-per-row.
// ---------------------------------------------------------------------------------------------------
constexpr int FLAT_K = 8;
constexpr int FLAT_CAP = SPMV_THREADS * FLAT_K;       // 2048 products, 16 KB
__global__ void __launch_bounds__(SPMV_THREADS, 4)
spmv_flat_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, const double *__restrict__ val,
                 const double *__restrict__ x, const double *__restrict__ xl, const double *__restrict__ z,
                 double *__restrict__ y, const int4 *__restrict__ blk_info, const double *__restrict__ scale,
                 double *__restrict__ v_store, int32_t row_offset, double ca, double cb, double cg) {
  __shared__ double prod[FLAT_CAP];
  __shared__ int32_t rp[FLAT_CAP + 1];
  __shared__ int32_t long_rows[SPMV_MAX_LONG];
  __shared__ int n_long;
  const int tid = threadIdx.x;
  // Programmatic dependent launch: the next kernel in the stream may start launching now; everything this
  // kernel reads before griddepcontrol.wait (block descriptor, row pointers, matrix entries) is constant,
  // so that prologue overlaps the tail of the previous kernel (whose output x this SpMV gathers).
  asm volatile("griddepcontrol.launch_dependents;");
  const int4 info = __ldg(blk_info + blockIdx.x);
  const int32_t r0 = info.x, r1 = info.y, e0 = info.z, e1 = info.w;
  if (r0 >= r1) return;
  const int32_t span = e1 - e0, nrows = r1 - r0;
  if (span > FLAT_CAP) {                               // a very long row lives here
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const double sc = scale ? __ldg(scale) : 1.0;
    const SpmvEpilogue ep{xl, z, y, ca * sc, cb * sc, cg, row_offset};
    if (v_store)
      for (int32_t r = r0 + tid; r < r1; r += SPMV_THREADS) v_store[r - row_offset] = xl[r - row_offset] * sc;
    rows_subwarp<32>(rowptr, col, val, x, ep, r0, r1, tid, long_rows, &n_long);
    return;
  }
  double a[FLAT_K];
  int32_t c[FLAT_K];
#pragma unroll
  for (int k = 0; k < FLAT_K; ++k) {
    const int32_t i = tid + k * SPMV_THREADS;
    a[k] = 0.0; c[k] = -1;
    if (i < span) { a[k] = __ldcs(val + e0 + i); c[k] = __ldcs(col + e0 + i); }
  }
  for (int32_t rr = tid; rr <= nrows; rr += SPMV_THREADS) rp[rr] = rowptr[r0 + rr] - e0;
  asm volatile("griddepcontrol.wait;" ::: "memory");   // the previous kernel (producer of x, z, *scale) is complete
  const double sc = scale ? __ldg(scale) : 1.0;
  const SpmvEpilogue ep{xl, z, y, ca * sc, cb * sc, cg, row_offset};
  double pre0 = 0.0, xs0 = 0.0;
  if (tid < nrows) {
    pre0 = ep.prefetch(r0 + tid);
    if (v_store) xs0 = xl[r0 + tid - row_offset] * sc;
  }
#pragma unroll
  for (int k = 0; k < FLAT_K; ++k) {
    const int32_t i = tid + k * SPMV_THREADS;
    if (c[k] >= 0) prod[i] = a[k] * __ldg(x + c[k]);
  }
  __syncthreads();
  for (int32_t rr = tid; rr < nrows; rr += SPMV_THREADS) {
    const int32_t lo = rp[rr], hi = rp[rr + 1];
    double s0 = 0.0, s1 = 0.0;
    int32_t i = lo;
    for (; i + 1 < hi; i += 2) { s0 += prod[i]; s1 += prod[i + 1]; }
    if (i < hi) s0 += prod[i];
    const int32_t r = r0 + rr;
    if (rr == tid) {
      ep.emit(r, s0 + s1, pre0);
      if (v_store) v_store[r - row_offset] = xs0;
    } else {
      ep.emit(r, s0 + s1);
      if (v_store) v_store[r - row_offset] = xl[r - row_offset] * sc;
    }
  }
}


// ---------------------------------------------------------------------------------------------------
// Row-partitioned SpMV with the halo exchange fused in (multi-GPU, dist.cu explains the layout).
// The flat kernel above, with three changes:
//   * the gather source is this rank's x buffer (own rows + packed halos), indexed by the pre-computed col_c;
//   * after its constant prologue (and the dependent-launch wait) a CTA polls this rank's flags until every peer's
//     push of the input vector (number wait_seq) has landed -- peers write the halo slots over NVLink;
//   * the epilogue keeps the block's y values in shared memory and one warp per peer copies the block's segment of
//     that peer's export list straight into the peer's slot `me` of the OUTPUT buffer (packed, coalesced 8-byte
//     stores through the peer mapping) while other blocks still compute; a one-warp successor kernel then raises
//     this rank's flag on every peer (release at system scope).  Values that nobody needs never travel.
// ---------------------------------------------------------------------------------------------------
struct DistArgs {
  const int32_t *col_c;
  int32_t e_lo;
  const unsigned int *flags;           // this rank's flags: flags[p] = last production rank p has pushed completely
  const unsigned long long *peers;     // arena base of every rank in this process' address space
  size_t out_off;                      // byte offset of the output vector inside an arena
  const int32_t *exp_ids, *blk_exp;
  int32_t n_pad;
  int me, R;
  uint32_t wait_seq, push_seq;
  int *err;
  int diag;                            // EIGKL_DIST_DIAG (timing experiments, results invalid): 1 no export stores, 4 no flag wait
};

__device__ __forceinline__ void dist_wait_flags(const DistArgs &D) {
  if (D.diag & 4) return;
  if (threadIdx.x < (unsigned)D.R && (int)threadIdx.x != D.me) {
    const unsigned int *f = D.flags + threadIdx.x;
    unsigned int v;
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int)(v - D.wait_seq) >= 0) break;
      if (clock64() - t0 > 6000000000ll) { atomicExch(D.err, 1); break; }      // ~3 s: ranks out of step; reported by dist_check
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(SPMV_THREADS, 4)
spmv_dist_kernel(const int32_t *__restrict__ rowptr, const double *__restrict__ val, const double *x, const double *xl,
                 const double *__restrict__ z, double *__restrict__ y, const int4 *__restrict__ blk_info,
                 const double *__restrict__ scale, double *__restrict__ v_store, int32_t row_offset, double ca, double cb,
                 double cg, const DistArgs D) {
  __shared__ double prod[FLAT_CAP];
  __shared__ double ys[FLAT_CAP];
  __shared__ int32_t rp[FLAT_CAP + 1];
  __shared__ int32_t long_rows[SPMV_MAX_LONG];
  __shared__ int n_long;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  asm volatile("griddepcontrol.launch_dependents;");
  const int4 info = __ldg(blk_info + blockIdx.x);
  const int32_t r0 = info.x, r1 = info.y, e0 = info.z, e1 = info.w;
  const int32_t span = e1 - e0, nrows = r1 - r0;
  const bool flat = nrows > 0 && span <= FLAT_CAP;
  double a[FLAT_K];
  int32_t c[FLAT_K];
  if (flat) {
#pragma unroll
    for (int k = 0; k < FLAT_K; ++k) {
      const int32_t i = tid + k * SPMV_THREADS;
      a[k] = 0.0; c[k] = -1;
      if (i < span) { a[k] = __ldcs(val + e0 + i); c[k] = __ldcs(D.col_c + (e0 - D.e_lo) + i); }
    }
    for (int32_t rr = tid; rr <= nrows; rr += SPMV_THREADS) rp[rr] = rowptr[r0 + rr] - e0;
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  dist_wait_flags(D);
  const double sc = scale ? __ldg(scale) : 1.0;
  const SpmvEpilogue ep{xl, z, y, ca * sc, cb * sc, cg, row_offset};
  if (flat) {
    double pre0 = 0.0, xs0 = 0.0;
    if (tid < nrows) {
      pre0 = ep.prefetch(r0 + tid);
      if (v_store) xs0 = xl[r0 + tid - row_offset] * sc;
    }
#pragma unroll
    for (int k = 0; k < FLAT_K; ++k) {
      const int32_t i = tid + k * SPMV_THREADS;
      if (c[k] >= 0) prod[i] = a[k] * __ldcg(x + c[k]);            // halo slots are written by peers: L2 is the coherence point
    }
    __syncthreads();
    for (int32_t rr = tid; rr < nrows; rr += SPMV_THREADS) {
      const int32_t lo = rp[rr], hi = rp[rr + 1];
      double s0 = 0.0, s1 = 0.0;
      int32_t i = lo;
      for (; i + 1 < hi; i += 2) { s0 += prod[i]; s1 += prod[i + 1]; }
      if (i < hi) s0 += prod[i];
      const int32_t r = r0 + rr;
      const double yv = ep.ca * (s0 + s1) + (rr == tid ? pre0 : ep.prefetch(r));
      y[r - row_offset] = yv;
      ys[rr] = yv;
      if (v_store) v_store[r - row_offset] = (rr == tid) ? xs0 : xl[r - row_offset] * sc;
    }
  } else if (nrows > 0) {                                         // a row too long for the staging buffer lives here
    if (v_store)
      for (int32_t r = r0 + tid; r < r1; r += SPMV_THREADS) v_store[r - row_offset] = xl[r - row_offset] * sc;
    // the sub-warp routine gathers through the read-only path with the ORIGINAL column ids: not usable on the
    // packed buffer, so this (rare) block walks its rows warp by warp over col_c
    for (int32_t r = r0 + warp; r < r1; r += SPMV_THREADS / 32) {
      const int32_t lo = rowptr[r], hi = rowptr[r + 1];
      double s = 0.0;
      for (int32_t i = lo + lane; i < hi; i += 32) s += __ldcs(val + i) * __ldcg(x + __ldcs(D.col_c + (i - D.e_lo)));
      s = warp_sum(s);
      if (lane == 0) ep.emit(r, s);
    }
  }
  (void)long_rows; (void)n_long;
  if (D.push_seq == 0 || (D.diag & 1)) return;
  __syncthreads();
  // ---- push this block's export rows: every thread takes entries of the block's segments of the peers' export lists
  //      (coalesced 8-byte stores through the peer mapping).  No fence, no flag here: the stores of a kernel are complete
  //      when it ends, and dist_flag_kernel, the next launch in the stream, raises this rank's flag on every peer ----
  for (int q = 0; q < D.R; ++q) {
    if (q == D.me) continue;
    const int32_t s = D.blk_exp[(size_t)blockIdx.x * D.R + q], e = D.blk_exp[(size_t)(blockIdx.x + 1) * D.R + q];
    double *dst = reinterpret_cast<double *>(D.peers[q] + D.out_off) + (size_t)D.me * D.n_pad;
    const int32_t *ids = D.exp_ids + (size_t)q * D.n_pad;
    for (int32_t i = s + tid; i < e; i += SPMV_THREADS) {
      const int32_t rl = ids[i];                                   // local row offset
      dst[i] = flat ? ys[rl - (r0 - row_offset)] : __ldcg(y + rl);
    }
  }
}

// raises this rank's flag on every peer: "my export rows of production `seq` have landed".  Launched as an ordinary
// stream successor of the kernel that stored them (spmv_dist_kernel / halo_push_kernel), i.e. after all of its stores --
// peer stores included -- are complete; the release at system scope orders the flag behind them for the peers' acquire.
// That replaces a system-scope fence in each of the SpMV's ~2 700 CTAs (38 of 104 us per SpMV on two GPUs at 2 M nodes).
__global__ void dist_flag_kernel(const unsigned long long *__restrict__ peers, int me, int R, uint32_t seq) {
  asm volatile("griddepcontrol.launch_dependents;");               // the next SpMV may start its constant prologue
  if (threadIdx.x < (unsigned)R && (int)threadIdx.x != me) {
    unsigned int *flag = reinterpret_cast<unsigned int *>(peers[threadIdx.x]) + me;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
  }
}
void dist_raise_flags(eigkl_handle *h, uint32_t seq) {
  dist_flag_kernel<<<1, 32, 0, h->stream>>>(h->arena.dev_ptrs.p, h->dist.me, h->dist.R, seq);
  h->launches++;
}

// ---------------------------------------------------------------------------------------------------
// Resident polynomial filter: the d SpMVs of one Chebyshev filter application as ONE cooperative launch.
//
// On the shipped circuits a SpMV launch is ~8 us of which the bytes explain ~3: the rest is launch/drain and
// dependent round trips, paid d = 16 times per Lanczos step.  A circuit's Laplacian (ibm10: 1.5 M entries,
// 18 MB) does not fit one SM but it does fit the CHIP: 148 SMs x (256 KB registers + 227 KB shared memory).
// So: one CTA per SM, each owning one row block for the whole recurrence,
//   * the block's values in registers (K consecutive entries per thread) with their 15-bit source index
//     and a row-start bit packed two to a register;
//   * x as this CTA needs it in shared memory: its own rows of y_{k-1} (it produced them) followed by the
//     block's HALO, the distinct columns outside its row range.  The halo list is built once per matrix
//     (cheb_resident_plan); per SpMV a CTA loads each halo value ONCE from L2 instead of once per entry
//     (the first-net node order makes a block reference ~0.2 distinct remote values per entry);
//   * its own rows of y_{k-2} in shared memory, so the recurrence's epilogue reads no global memory.
// A thread adds its K products in registers; rows that cross threads are finished by a segmented scan of
// the threads' open sums (five shuffle steps per warp, one shared-memory hop across warps), so the work
// per SpMV is the same for every thread whatever the row lengths are.  The SpMVs are separated by a grid
// barrier (release-add + acquire-poll on one counter) instead of a kernel boundary.  x is re-written inside
// the kernel, so the halo is read with ordinary coherent loads (no ld.global.nc); the acquire at the
// barrier invalidates L1.
// Not used when the matrix does not fit (cheb_resident_plan decides) or with more than one rank (the halo
// exchange between SpMVs is a NCCL call): those take one spmv_flat_kernel launch per SpMV.
// ---------------------------------------------------------------------------------------------------
constexpr int RES_THREADS = 768;
constexpr int RES_WARPS = RES_THREADS / 32;
constexpr int RES_KMAX = 16;
constexpr int RES_CAP = RES_THREADS * RES_KMAX;       // 12288 entry slots per CTA; a block's span is <= RES_CAP - 1
constexpr int RES_MAXROWS = RES_CAP / 2;              // the builder charges a row like one more entry, so this always holds
constexpr int RES_MAXHALO = 6144;
constexpr int RES_STAGE = RES_CAP + RES_THREADS + 8;   // padded staging of the prologue
static_assert(RES_MAXROWS + RES_MAXHALO <= 32767, "source index must fit 15 bits");
static_assert(RES_STAGE <= RES_MAXROWS + RES_MAXHALO + RES_MAXROWS, "value staging aliases the x cache and the y_{k-2} rows");
static_assert(RES_STAGE * 2 <= RES_MAXROWS * 8, "index staging aliases the row sums");
constexpr size_t RES_SMEM = (size_t)(RES_MAXROWS + RES_MAXHALO) * 8 + 2 * (size_t)RES_MAXROWS * 8 + (size_t)RES_MAXHALO * 4 +
                            (size_t)RES_WARPS * 16;
constexpr int PLAN_THREADS = 1024;
constexpr int32_t PLAN_MAX_N = 800000;                // bitmap + prefix of the plan kernel: n/4 bytes of shared memory

struct ResidentArgs {
  const double *val;
  const uint16_t *src;          // per entry: index into the CTA's x cache | 0x8000 at a row start
  const int4 *info;
  const int32_t *halo_ids, *halo_cnt;
  const double *x_in, *scale;
  double *v_store;
  double *w[3];
  uint4 *ll;                    // two slots of n {lo, tag, hi, tag} words: y_k of every row, published for the halos
  uint32_t tag_base;            // y_k of this launch carries tag tag_base + k
  int32_t n;
  int deg;
  double fc, fe;
  int phases;
  unsigned char out_idx[64];
};

// Halo exchange in the style of NCCL's LL protocol: a value travels as two 8-byte words {low half, tag},
// {high half, tag}; 8-byte stores are single-copy atomic, so a reader that sees both tags equal to the one it
// waits for has the value, without any fence or barrier.  Two slots (tag parity) are enough: L is symmetric,
// so a CTA can only publish y_{k+1} after every CTA that reads its rows has published y_k, i.e. has already
// consumed its y_{k-1}, the value being overwritten.
__device__ __forceinline__ void ll_store(uint4 *p, double v, uint32_t tag) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((uint32_t)__double2loint(v)), "r"(tag),
               "r"((uint32_t)__double2hiint(v)), "r"(tag)
               : "memory");
}
__device__ __forceinline__ uint4 ll_load(const uint4 *p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// per-phase clocks of CTA 0 (EIGKL_RES_PHASES=1 prints them after the solve): halo, products, row sums, barrier
__device__ unsigned long long res_phase_clk[5];
#define RES_CLK(slot, t0)                                                         \
  if (A.phases && blockIdx.x == 0 && tid == 0) {                                  \
    const long long t1 = clock64();                                               \
    res_phase_clk[slot] += (unsigned long long)(t1 - t0);                         \
    t0 = t1;                                                                      \
  }

// One CTA per resident block, once per matrix: the block's halo (sorted distinct remote columns) and, per
// entry, where its x value will sit in the CTA's shared-memory x cache (own rows first, then the halo).
__global__ void __launch_bounds__(PLAN_THREADS)
res_plan_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, const int4 *__restrict__ info, int32_t n,
                uint16_t *__restrict__ src, int32_t *__restrict__ halo_ids, int32_t *__restrict__ halo_cnt,
                int32_t *__restrict__ check) {
  extern __shared__ uint32_t plan_sm[];
  const int32_t W = (n + 31) >> 5;
  uint32_t *bitmap = plan_sm, *prefix = plan_sm + W, *wsum = prefix + W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int4 bi = info[blockIdx.x];
  const int32_t r0 = bi.x, r1 = bi.y, e0 = bi.z, span = bi.w - bi.z, nrows = bi.y - bi.x;
  for (int32_t w = tid; w < W; w += PLAN_THREADS) bitmap[w] = 0u;
  __syncthreads();
  for (int32_t i = tid; i < span; i += PLAN_THREADS) {
    const int32_t c = col[e0 + i];
    if (c < r0 || c >= r1) atomicOr(&bitmap[c >> 5], 1u << (c & 31));
  }
  __syncthreads();
  const int32_t chunk = (W + PLAN_THREADS - 1) / PLAN_THREADS;
  const int32_t w0 = min(W, tid * chunk), w1 = min(W, w0 + chunk);
  uint32_t mine = 0;
  for (int32_t w = w0; w < w1; ++w) mine += __popc(bitmap[w]);
  uint32_t inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(FULL_MASK, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint32_t before = 0;
  for (int q = 0; q < warp; ++q) before += wsum[q];
  uint32_t run = before + inc - mine;
  for (int32_t w = w0; w < w1; ++w) {
    prefix[w] = run;
    uint32_t bits = bitmap[w];
    while (bits) {
      const int b = __ffs(bits) - 1;
      if (run < (uint32_t)RES_MAXHALO) halo_ids[(size_t)blockIdx.x * RES_MAXHALO + run] = (w << 5) + b;
      ++run;
      bits &= bits - 1;
    }
  }
  if (tid == PLAN_THREADS - 1) {
    halo_cnt[blockIdx.x] = (int32_t)run;
    atomicMax(check + 2, (int32_t)run);
  }
  __syncthreads();
  for (int32_t i = tid; i < span; i += PLAN_THREADS) {
    const int32_t c = col[e0 + i];
    uint32_t idx;
    if (c >= r0 && c < r1) idx = (uint32_t)(c - r0);
    else idx = (uint32_t)nrows + prefix[c >> 5] + __popc(bitmap[c >> 5] & ((1u << (c & 31)) - 1u));
    src[e0 + i] = (uint16_t)(idx & 0x7FFFu);
  }
  __syncthreads();
  for (int32_t rr = tid; rr < nrows; rr += PLAN_THREADS) src[rowptr[r0 + rr]] |= (uint16_t)0x8000u;
}

template <int K, int T>
__global__ void __launch_bounds__(T, 1) cheb_resident_kernel(const ResidentArgs A) {
  extern __shared__ __align__(16) unsigned char res_smem[];
  double *xc = reinterpret_cast<double *>(res_smem);              // x cache: own rows of y_{k-1}, then the halo
  double *zc = xc + RES_MAXROWS + RES_MAXHALO;                    // own rows of y_{k-2}
  double *rsum = zc + RES_MAXROWS;                                // (L y_{k-1}) of the own rows, as the threads finish them
  double *wout = rsum + RES_MAXROWS;                              // per warp: sum of the warp's open segment
  int32_t *wflag = reinterpret_cast<int32_t *>(wout + (T / 32)); // per warp: a row starts inside the warp / start count
  int32_t *hid = wflag + 2 * (T / 32);                           // halo column ids
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int4 info = A.info[blockIdx.x];
  const int32_t r0 = info.x, e0 = info.z;
  const int32_t nrows = info.y - info.x, span = info.w - info.z;
  const int32_t H = A.halo_cnt[blockIdx.x];
  // ---- prologue: the block's entries, coalesced into a padded staging area, then K consecutive per thread ----
  {
    double *stg_a = xc;
    uint16_t *stg_s = reinterpret_cast<uint16_t *>(rsum);
    for (int32_t i = tid; i < T * K; i += T) {
      double v = 0.0;
      uint16_t q = (i == span) ? (uint16_t)0x8000u : (uint16_t)0;  // sentinel: the row after the last one starts here
      if (i < span) { v = __ldcs(A.val + e0 + i); q = __ldcs(A.src + e0 + i); }
      const int32_t pos = i + i / K;
      stg_a[pos] = v;
      stg_s[pos] = q;
    }
    const int32_t *hsrc = A.halo_ids + (size_t)blockIdx.x * RES_MAXHALO;
    for (int32_t j = tid; j < H; j += T) hid[j] = hsrc[j];
  }
  __syncthreads();
  double a[K];
  uint32_t sp[K / 2];                                             // two 16-bit source words per register
  uint32_t flags = 0u;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int32_t pos = tid * (K + 1) + k;
    a[k] = xc[pos];
    const uint32_t q = reinterpret_cast<const uint16_t *>(rsum)[pos];
    if (k & 1) sp[k / 2] |= q << 16; else sp[k / 2] = q;
    flags |= (q >> 15) << k;
  }
  // first row that starts at or after this thread's entries = number of row starts before them
  int32_t first_row;
  {
    const int cnt = __popc(flags);
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(FULL_MASK, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) wflag[(T / 32) + warp] = inc;
    __syncthreads();                                              // also: the staging area has been consumed
    int before = 0;
    for (int q = 0; q < warp; ++q) before += wflag[(T / 32) + q];
    first_row = before + inc - cnt;
  }
  const double sc = A.scale ? *A.scale : 1.0;
  for (int32_t rr = tid; rr < nrows; rr += T) {
    const double y0 = sc * A.x_in[r0 + rr];                       // T_0 x = the normalised vector
    xc[rr] = y0;
    zc[rr] = 0.0;
    if (A.v_store) A.v_store[r0 + rr] = y0;
  }
  const unsigned wm = __ballot_sync(FULL_MASK, flags != 0u);      // lanes of this warp in which a row starts
  const unsigned below_eq = wm & ((2u << lane) - 1u), below = wm & ((1u << lane) - 1u);
  const int seg = below_eq ? 31 - __clz(below_eq) : 0;            // first lane of this lane's scan segment
  if (lane == 0) wflag[warp] = wm != 0u;
  double ca = -1.0 / A.fe, cb = A.fc / A.fe, cg = 0.0;            // y1 = (c y0 - L y0) / e
  double *yout = A.w[A.out_idx[A.deg - 1]];
  constexpr int HU = (K == 16) ? 2 : 4;                             // halo values in flight per thread (registers are tight at K = 16)
  long long tclk = clock64();
  for (int kk = 1; kk <= A.deg; ++kk) {
    if (kk == 1) {
      for (int32_t j0 = tid; j0 < H; j0 += T * HU) {
        double v[HU];
#pragma unroll
        for (int u = 0; u < HU; ++u) {
          const int32_t j = j0 + u * T;
          v[u] = j < H ? A.x_in[hid[j]] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < HU; ++u) {
          const int32_t j = j0 + u * T;
          if (j < H) xc[nrows + j] = sc * v[u];
        }
      }
    } else {
      const uint32_t tag = A.tag_base + (uint32_t)(kk - 1);       // wait for y_{kk-1} of the rows in the halo
      const uint4 *src = A.ll + (size_t)(tag & 1u) * (size_t)A.n;
      for (int32_t j0 = tid; j0 < H; j0 += T * HU) {
        uint4 v[HU];
        const uint4 *ptr[HU];
#pragma unroll
        for (int u = 0; u < HU; ++u) {
          const int32_t j = j0 + u * T;
          ptr[u] = j < H ? src + hid[j] : nullptr;
        }
        bool ok;
        do {
          ok = true;
#pragma unroll
          for (int u = 0; u < HU; ++u)
            if (ptr[u]) v[u] = ll_load(ptr[u]);
#pragma unroll
          for (int u = 0; u < HU; ++u)
            if (ptr[u]) ok = ok && v[u].y == tag && v[u].w == tag;
        } while (!ok);
#pragma unroll
        for (int u = 0; u < HU; ++u) {
          const int32_t j = j0 + u * T;
          if (j < H) xc[nrows + j] = __hiloint2double((int)v[u].z, (int)v[u].x);
        }
      }
    }
    __syncthreads();
    RES_CLK(0, tclk);
    double run = 0.0, head = 0.0;
    int j = 0;
    constexpr int XB = K >= 8 ? 8 : K;                            // x values in flight per thread
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += XB) {
      double xv[XB];
#pragma unroll
      for (int u = 0; u < XB; ++u) {                              // all loads of the batch before any store: the
        const int k = k0 + u;                                     // compiler cannot prove rsum[] and xc[] apart
        const uint32_t q = (k & 1) ? (sp[k / 2] >> 16) : (sp[k / 2] & 0xFFFFu);
        xv[u] = xc[q & 0x7FFFu];
      }
#pragma unroll
      for (int u = 0; u < XB; ++u) {
        const int k = k0 + u;
        if ((flags >> k) & 1u) {
          if (j == 0) head = run;
          else rsum[first_row + j - 1] = run;                     // a row that lies entirely inside this thread
          run = 0.0;
          ++j;
        }
        run = fma(a[k], xv[u], run);
      }
    }
    // segmented inclusive scan of the open sums: inc = sum of `run` over [segment start, lane]
    double inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double t = __shfl_up_sync(FULL_MASK, inc, d);
      if (lane - d >= seg) inc += t;
    }
    if (lane == 31) wout[warp] = inc;
    double cin = __shfl_up_sync(FULL_MASK, inc, 1);
    if (lane == 0) cin = 0.0;
    __syncthreads();
    RES_CLK(1, tclk);
    if (flags != 0u && first_row >= 1) {
      if (below == 0u) {                                          // the row began in an earlier warp
        for (int wq = warp - 1; wq >= 0; --wq) {
          cin += wout[wq];
          if (wflag[wq]) break;
        }
      }
      rsum[first_row - 1] = cin + head;
    }
    __syncthreads();
    const uint32_t tag_out = A.tag_base + (uint32_t)kk;
    uint4 *dst = A.ll + (size_t)(tag_out & 1u) * (size_t)A.n + r0;
    for (int32_t rr = tid; rr < nrows; rr += T) {       // y_k = ca L y_{k-1} + cb y_{k-1} + cg y_{k-2}
      const double xo = xc[rr];
      const double yn = ca * rsum[rr] + cb * xo + cg * zc[rr];
      if (kk < A.deg) ll_store(dst + rr, yn, tag_out);            // intermediate vectors exist only for the halos
      else yout[r0 + rr] = yn;
      zc[rr] = xo;
      xc[rr] = yn;
    }
    RES_CLK(2, tclk);
    if (A.phases && blockIdx.x == 0 && tid == 0) res_phase_clk[4] += 1;
    ca = -2.0 / A.fe; cb = 2.0 * A.fc / A.fe; cg = -1.0;
  }
}

void spmv_resident_print_phases() {
  if (!getenv("EIGKL_RES_PHASES")) return;
  unsigned long long c[5] = {0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(c, res_phase_clk, sizeof(c)) != cudaSuccess || c[4] == 0) return;
  fprintf(stderr, "resident filter, CTA 0, cycles per SpMV: halo wait+load %.0f  products+scan %.0f  rows %.0f  (%llu SpMVs)\n",
          (double)c[0] / c[4], (double)c[1] / c[4], (double)c[2] / c[4], c[4]);
  unsigned long long z[5] = {0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(res_phase_clk, z, sizeof(z));
}

// Decides whether this matrix runs resident and builds its plan (row blocks, halo lists, source indices).
// Called by assemble_laplacian before its one host synchronisation; cheb_resident_plan_finish reads the
// checks back after it.
void cheb_resident_plan(eigkl_handle *h) {
  auto &L = h->L;
  const int32_t n = L.n;
  L.res_ok = false;
  L.res_blocks = 0;
  L.res_k = 0;
  // a row is charged like one more entry, so that a block never holds more rows than half its entry capacity
  const int64_t cost = L.nnz + (int64_t)n;
  // with several ranks the plan is still made: a matrix that fits the chip is solved replicated on every rank
  if ((h->opts.nranks != 1 && h->dist_mode == 1) || !h->spmv_resident || n > PLAN_MAX_N || cost > (int64_t)(RES_CAP - 1) * h->sm_count) return;
  int64_t min_chunk = 1024;
  if (const char *ev = getenv("EIGKL_RES_CHUNK")) min_chunk = std::max<int64_t>(64, atoll(ev));   // tuning aid
  int64_t chunk = std::max<int64_t>(min_chunk, ceil_div(cost, (int64_t)h->sm_count));
  chunk = ceil_div(chunk, 32) * 32;
  L.res_blocks = (int32_t)std::max<int64_t>(1, ceil_div(cost, chunk));
  L.res_chunk = chunk;
  L.res_info.alloc((size_t)4 * L.res_blocks + 4);
  L.res_row.alloc((size_t)L.res_blocks + 1);
  L.res_check.alloc(4);
  L.res_ll.alloc((size_t)4 * n + 8);                // 2 slots x n x 16 bytes
  L.res_src.alloc((size_t)L.nnz + 8);
  L.res_halo_ids.alloc((size_t)L.res_blocks * RES_MAXHALO);
  L.res_halo_cnt.alloc((size_t)L.res_blocks);
  EIGKL_CUDA(cudaMemsetAsync(L.res_check.p, 0, 4 * sizeof(int32_t), h->stream));
  EIGKL_CUDA(cudaMemsetAsync(L.res_ll.p, 0, (size_t)4 * n * sizeof(unsigned long long), h->stream));
  L.res_tag = 0;
  resident_row_blocks(h, chunk);                                   // assemble.cu: res_row, res_info, res_check[0..1]
  const size_t plan_smem = ((size_t)2 * ((n + 31) / 32) + 64) * sizeof(uint32_t);
  if (!h->attr_resident) {
    EIGKL_CUDA(cudaFuncSetAttribute(res_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(((size_t)2 * ((PLAN_MAX_N + 31) / 32) + 64) * sizeof(uint32_t))));
    EIGKL_CUDA(cudaFuncSetAttribute(cheb_resident_kernel<4, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RES_SMEM));
    EIGKL_CUDA(cudaFuncSetAttribute(cheb_resident_kernel<8, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RES_SMEM));
    EIGKL_CUDA(cudaFuncSetAttribute(cheb_resident_kernel<16, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RES_SMEM));
    EIGKL_CUDA(cudaFuncSetAttribute(cheb_resident_kernel<24, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RES_SMEM));
    h->attr_resident = true;
  }
  res_plan_kernel<<<(unsigned)L.res_blocks, PLAN_THREADS, plan_smem, h->stream>>>(
      L.rowptr.p, L.col.p, reinterpret_cast<const int4 *>(L.res_info.p), n, L.res_src.p, L.res_halo_ids.p, L.res_halo_cnt.p,
      L.res_check.p);
  EIGKL_CUDA(cudaGetLastError());
  h->launches += 1;
  EIGKL_CUDA(cudaMemcpyAsync(L.res_check_host, L.res_check.p, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
}

void cheb_resident_plan_finish(eigkl_handle *h) {
  auto &L = h->L;
  if (L.res_blocks <= 0) return;
  const int32_t max_span = L.res_check_host[0], max_rows = L.res_check_host[1], max_halo = L.res_check_host[2];
  // 768 threads x 4 or 8 entries; the largest blocks run as 512 threads x 24 entries (128 registers per thread
  // instead of 80: industry2 5.6 vs 6.8 ms per solve, ibm10 equal).  K = 16 (768 threads) is kept for tuning.
  int k = 0;
  if (max_span <= RES_THREADS * 4 - 1) k = 4;
  else if (max_span <= RES_THREADS * 8 - 1) k = 8;
  else if (max_span <= RES_CAP - 1) k = 24;
  if (const char *ev = getenv("EIGKL_RES_K")) {                   // tuning aid
    const int want = atoi(ev);
    const int cap = (want == 24 ? 512 : RES_THREADS) * want - 1;
    if ((want == 4 || want == 8 || want == 16 || want == 24) && k != 0 && max_span <= cap) k = want;
  }
  L.res_k = k;
  if (getenv("EIGKL_RES_PHASES"))
    fprintf(stderr, "resident plan: n %d nnz %lld blocks %d chunk %lld max span %d rows %d halo %d -> K %d\n", L.n, (long long)L.nnz,
            L.res_blocks, (long long)L.res_chunk, max_span, max_rows, max_halo, k);
  L.res_ok = k != 0 && max_rows > 0 && max_rows <= RES_MAXROWS && max_halo <= RES_MAXHALO;
}

bool cheb_resident_usable(const eigkl_handle *h) {
  return h->spmv_resident && !h->dist.valid && h->L.valid && h->L.res_ok;
}

void cheb_resident_launch(eigkl_handle *h, const double *x_in, const double *scale, double *v_store, double *const w[3],
                          const unsigned char *out_idx, int deg, double fc, double fe) {
  auto &L = h->L;
  EIGKL_REQUIRE(cheb_resident_usable(h) && deg >= 1 && deg <= 64, EIGKL_E_ARG, "resident filter not available");
  ResidentArgs A{};
  A.val = L.val.p; A.src = L.res_src.p;
  A.info = reinterpret_cast<const int4 *>(L.res_info.p);
  A.halo_ids = L.res_halo_ids.p; A.halo_cnt = L.res_halo_cnt.p;
  A.x_in = x_in; A.scale = scale; A.v_store = v_store;
  for (int b = 0; b < 3; ++b) A.w[b] = w[b];
  A.ll = reinterpret_cast<uint4 *>(L.res_ll.p);
  A.tag_base = L.res_tag;
  A.n = L.n;
  A.deg = deg; A.fc = fc; A.fe = fe;
  static const bool phases = getenv("EIGKL_RES_PHASES") != nullptr;
  A.phases = phases ? 1 : 0;
  for (int k = 0; k < deg; ++k) A.out_idx[k] = out_idx[k];
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)L.res_blocks);
  cfg.blockDim = dim3(L.res_k == 24 ? 512 : RES_THREADS);
  cfg.dynamicSmemBytes = RES_SMEM;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;       // every CTA resident at once: a CTA waits on its neighbours' values
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (L.res_k == 4) EIGKL_CUDA(cudaLaunchKernelEx(&cfg, cheb_resident_kernel<4, 768>, A));
  else if (L.res_k == 8) EIGKL_CUDA(cudaLaunchKernelEx(&cfg, cheb_resident_kernel<8, 768>, A));
  else if (L.res_k == 16) EIGKL_CUDA(cudaLaunchKernelEx(&cfg, cheb_resident_kernel<16, 768>, A));
  else EIGKL_CUDA(cudaLaunchKernelEx(&cfg, cheb_resident_kernel<24, 512>, A));
  L.res_tag += (uint32_t)deg;
  h->launches++;
}

// y = ca*s*(L x) + cb*s*x + cg*z for this rank's rows (s = *scale_inv or 1).
//   xg: full-length x (global column ids, the gather source); xl: this rank's slice of the same vector;
//   z, y, store_scaled: rank-local slices.
void spmv_launch_ex(eigkl_handle *h, const double *xg, const double *xl, const double *z, double *y, const double *scale_inv,
                    double *store_scaled, double ca, double cb, double cg, const SpmvDist *dist) {
  auto &L = h->L;
  EIGKL_REQUIRE(L.valid, EIGKL_E_ARG, "Laplacian not assembled");
  if (h->dist.valid) {
    // row-partitioned: launched even by a rank without rows, because the flag protocol counts every rank
    EIGKL_REQUIRE(dist != nullptr && L.flat, EIGKL_E_ARG, "row-partitioned SpMV needs its exchange arguments");
    auto &P = h->dist;
    DistArgs D{};
    D.col_c = P.col_c.p; D.e_lo = P.e_lo;
    D.flags = reinterpret_cast<const unsigned int *>(h->arena.base);
    D.peers = h->arena.dev_ptrs.p;
    D.out_off = PEER_FLAGS_BYTES + (size_t)dist->out_buf * h->arena.vec_bytes;
    D.exp_ids = P.exp_ids.p; D.blk_exp = P.blk_exp.p;
    D.n_pad = P.n_pad; D.me = P.me; D.R = P.R;
    D.wait_seq = dist->wait_seq; D.push_seq = dist->push_seq;
    D.err = h->arena.err.p;
    static const int diag = getenv("EIGKL_DIST_DIAG") ? atoi(getenv("EIGKL_DIST_DIAG")) : 0;
    D.diag = diag;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)L.n_blocks);
    cfg.blockDim = dim3(SPMV_THREADS);
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = h->spmv_pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const int4 *info = reinterpret_cast<const int4 *>(L.blk_info.p);
    const int32_t *rp = L.rowptr.p;
    const double *vl = L.val.p;
    h->prof.begin(KC_SPMV, h->stream);
    EIGKL_CUDA(cudaLaunchKernelEx(&cfg, spmv_dist_kernel, rp, vl, xg, xl, z, y, info, scale_inv, store_scaled, L.row_lo, ca, cb, cg, D));
    h->launches++;
    if (dist->push_seq != 0) dist_raise_flags(h, dist->push_seq);
    h->prof.end(h->stream);
    return;
  }
  if (L.row_hi <= L.row_lo) return;
  h->prof.begin(KC_SPMV, h->stream);
  const bool with_stream = h->spmv_mode == 1 || (h->spmv_mode == 0 && L.nnz < 6 * (int64_t)L.n);
  if (L.flat) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)L.n_blocks);
    cfg.blockDim = dim3(SPMV_THREADS);
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = h->spmv_pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const int4 *info = reinterpret_cast<const int4 *>(L.blk_info.p);
    const int32_t *rp = L.rowptr.p, *cl = L.col.p;
    const double *vl = L.val.p;
    EIGKL_CUDA(cudaLaunchKernelEx(&cfg, spmv_flat_kernel, rp, cl, vl, xg, xl, z, y, info, scale_inv, store_scaled, L.row_lo, ca, cb, cg));
  } else if (with_stream)
    spmv_adaptive_kernel<true><<<(unsigned)L.n_blocks, SPMV_THREADS, 0, h->stream>>>(L.rowptr.p, L.col.p, L.val.p, xg, xl, z, y, L.blk_row.p,
                                                                                   scale_inv, store_scaled, L.row_lo, h->spmv_mode, ca, cb, cg);
  else
    spmv_adaptive_kernel<false><<<(unsigned)L.n_blocks, SPMV_THREADS, 0, h->stream>>>(L.rowptr.p, L.col.p, L.val.p, xg, xl, z, y, L.blk_row.p,
                                                                                    scale_inv, store_scaled, L.row_lo, h->spmv_mode, ca, cb, cg);
  h->prof.end(h->stream);
  h->launches++;
}

// plain product y = L x (x full-length, y the rank's slice)
void spmv_launch(eigkl_handle *h, const double *x, double *y, const double *scale_inv, double *store_scaled) {
  spmv_launch_ex(h, x, x + h->L.row_lo, nullptr, y, scale_inv, store_scaled, 1.0, 0.0, 0.0);
}

}  // namespace eigkl
