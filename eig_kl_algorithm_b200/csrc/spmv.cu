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
// Fused epilogue/prologue for Lanczos: y = (L x) * (*scale) and, optionally, v_store = x * (*scale)
// for the block's own rows, so the basis vector v_j = w/beta is written by the SpMV that consumes it.
// Bound: HBM (or L2 when the matrix fits the 126 MB L2): nnz*12 + n*20 bytes per launch.
#include "internal.h"
#include "device_utils.cuh"

namespace eigkl {

constexpr int SPMV_THREADS = 256;
constexpr int SPMV_STAGE = 4096;        // staging capacity in products (32 KB)

// L lanes cooperate on one row: lanes stride the row (coalesced across the sub-warp and, because
// consecutive sub-warps own consecutive rows, across the warp), then a shuffle reduction of width L.
template <int L>
__device__ __forceinline__ void rows_subwarp(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                             const double *__restrict__ val, const double *__restrict__ x,
                                             double *__restrict__ y, int32_t r0, int32_t r1, double sc, int tid) {
  constexpr int ROWS_PER_ITER = SPMV_THREADS / L;
  const int sub = tid & (L - 1);
  for (int32_t base = r0; base < r1; base += ROWS_PER_ITER) {      // block-uniform trip count
    const int32_t r = base + tid / L;
    double s = 0.0;
    if (r < r1) {
      const int32_t lo = rowptr[r], hi = rowptr[r + 1];
      int32_t i = lo + sub;
      for (; i + L < hi; i += 2 * L) {                             // two independent gathers in flight
        const double a0 = val[i], a1 = val[i + L];
        const double x0 = __ldg(x + col[i]), x1 = __ldg(x + col[i + L]);
        s += a0 * x0;
        s += a1 * x1;
      }
      if (i < hi) s += val[i] * __ldg(x + col[i]);
    }
#pragma unroll
    for (int o = L >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
    if (r < r1 && sub == 0) y[r] = s * sc;
  }
}

// WITH_STREAM: the kernel variant that carries the 32 KB staging buffer (launched only when the matrix
// has row blocks short enough to want it, so that vector-only matrices keep their L1 for the x gathers).
// mode: 0 = choose per row block, 1 = always stream (shared-memory staged), 2 = always sub-warp/warp per row
template <bool WITH_STREAM>
__global__ void __launch_bounds__(SPMV_THREADS)
spmv_adaptive_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                     const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y,
                     const int32_t *__restrict__ blk_row, const double *__restrict__ scale,
                     double *__restrict__ v_store, int32_t row_offset, int mode) {
  __shared__ double prod[WITH_STREAM ? SPMV_STAGE : 1];
  const int tid = threadIdx.x;
  const int32_t r0 = blk_row[blockIdx.x], r1 = blk_row[blockIdx.x + 1];
  if (r0 >= r1) return;
  const double sc = scale ? __ldg(scale) : 1.0;
  const int32_t e0 = rowptr[r0], e1 = rowptr[r1];
  const int32_t span = e1 - e0, nrows = r1 - r0;
  y -= row_offset;                       // rows are global ids; y and v_store are this rank's slices
  if (v_store) {
    for (int32_t r = r0 + tid; r < r1; r += SPMV_THREADS) v_store[r - row_offset] = __ldg(x + r) * sc;
  }
  const int mean = span / nrows;
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
      if (rr < nrows && sub == 0) y[r0 + rr] = s * sc;
    }
  } else if (mean <= 4) {
    rows_subwarp<2>(rowptr, col, val, x, y, r0, r1, sc, tid);
  } else if (mean <= 8) {
    rows_subwarp<4>(rowptr, col, val, x, y, r0, r1, sc, tid);
  } else if (mean <= 16) {
    rows_subwarp<8>(rowptr, col, val, x, y, r0, r1, sc, tid);
  } else if (mean <= 48) {
    rows_subwarp<16>(rowptr, col, val, x, y, r0, r1, sc, tid);
  } else {
    rows_subwarp<32>(rowptr, col, val, x, y, r0, r1, sc, tid);   // warp per row (industry2-class rows)
  }
}

// x: full-length vector (global column ids); y / store_scaled: this rank's row slice
void spmv_launch(eigkl_handle *h, const double *x, double *y, const double *scale_inv, double *store_scaled) {
  auto &L = h->L;
  EIGKL_REQUIRE(L.valid, EIGKL_E_ARG, "Laplacian not assembled");
  if (L.row_hi <= L.row_lo) return;
  h->prof.begin(KC_SPMV, h->stream);
  const bool with_stream = h->spmv_mode == 1 || (h->spmv_mode == 0 && L.nnz < 6 * (int64_t)L.n);
  if (with_stream)
    spmv_adaptive_kernel<true><<<(unsigned)L.n_blocks, SPMV_THREADS, 0, h->stream>>>(L.rowptr.p, L.col.p, L.val.p, x, y, L.blk_row.p,
                                                                                   scale_inv, store_scaled, L.row_lo, h->spmv_mode);
  else
    spmv_adaptive_kernel<false><<<(unsigned)L.n_blocks, SPMV_THREADS, 0, h->stream>>>(L.rowptr.p, L.col.p, L.val.p, x, y, L.blk_row.p,
                                                                                    scale_inv, store_scaled, L.row_lo, h->spmv_mode);
  h->prof.end(h->stream);
  h->launches++;
}

}  // namespace eigkl
