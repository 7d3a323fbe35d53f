// spmv.cu -- fp64 CSR SpMV y = L x for the Lanczos solver (north-star subsystem 2).
//
// Replaces Spectra's SparseSymMatProd (cEIG.cpp:194, serial Eigen product) and is the B200 answer to
// the reference's only GPU SpMV, sparseMVKernel (gKL2.cu:65-89: thread-per-row, scalar, fp32).
//
// Layout: CSR with int32 rowptr/col and fp64 values, rows cut into row blocks of ~SPMV_CHUNK
// non-zeros (blk_row, built at assembly).  One CTA per row block:
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

// y = ca*s*(L x) + cb*s*x + cg*z for this rank's rows (s = *scale_inv or 1).
//   xg: full-length x (global column ids, the gather source); xl: this rank's slice of the same vector;
//   z, y, store_scaled: rank-local slices.
void spmv_launch_ex(eigkl_handle *h, const double *xg, const double *xl, const double *z, double *y, const double *scale_inv,
                    double *store_scaled, double ca, double cb, double cg) {
  auto &L = h->L;
  EIGKL_REQUIRE(L.valid, EIGKL_E_ARG, "Laplacian not assembled");
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
