// spmv.cu -- fp64 CSR SpMV y = L x for the Lanczos solver (north-star subsystem 2).
//
// Replaces Spectra's SparseSymMatProd (cEIG.cpp:194, serial Eigen product) and is the B200 answer to
// the reference's only GPU SpMV, sparseMVKernel (gKL2.cu:65-89: thread-per-row, scalar, fp32).
//
// Layout: CSR with int32 rowptr/col and fp64 values, rows cut into row blocks of ~SPMV_CHUNK
// non-zeros (blk_row, built at assembly).  One CTA per row block:
//   * stream mode (block's non-zeros fit the staging buffer): all threads stream val/col fully
//     coalesced, gather x through the read-only path, stage the products in shared memory, then
//     1..32 threads per row (chosen from the block's row count) reduce each row from shared memory;
//   * vector mode (a row longer than the buffer, e.g. industry2's 900-entry rows sharing a block):
//     one warp per row, lanes stride the row, warp-shuffle reduction.
// Fused epilogue/prologue for Lanczos: y = (L x) * (*scale) and, optionally, v_store = x * (*scale)
// for the block's own rows, so the basis vector v_j = w/beta is written by the SpMV that consumes it.
// Bound: HBM (or L2 when the matrix fits the 126 MB L2): nnz*12 + n*20 bytes per launch.
#include "internal.h"
#include "device_utils.cuh"

namespace eigkl {

constexpr int SPMV_THREADS = 256;
constexpr int SPMV_CHUNK = 2048;        // target non-zeros per row block (must match assemble.cu)
constexpr int SPMV_STAGE = 4096;        // staging capacity in products (32 KB)

__global__ void __launch_bounds__(SPMV_THREADS)
spmv_adaptive_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                     const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y,
                     const int32_t *__restrict__ blk_row, const double *__restrict__ scale,
                     double *__restrict__ v_store, int32_t row_offset) {
  __shared__ double prod[SPMV_STAGE];
  const int tid = threadIdx.x;
  const int32_t r0 = blk_row[blockIdx.x], r1 = blk_row[blockIdx.x + 1];
  if (r0 >= r1) return;
  const double sc = scale ? __ldg(scale) : 1.0;
  const int32_t e0 = rowptr[r0], e1 = rowptr[r1];
  const int32_t span = e1 - e0, nrows = r1 - r0;
  if (v_store) {
    for (int32_t r = r0 + tid; r < r1; r += SPMV_THREADS) v_store[r] = __ldg(x + row_offset + r) * sc;
  }
  if (span <= SPMV_STAGE) {
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
  } else {
    const int lane = tid & 31, warp = tid >> 5;
    for (int32_t r = r0 + warp; r < r1; r += SPMV_THREADS / 32) {
      const int32_t lo = rowptr[r], hi = rowptr[r + 1];
      double s = 0.0;
      for (int32_t i = lo + lane; i < hi; i += 32) s += val[i] * __ldg(x + col[i]);
      s = warp_sum(s);
      if (lane == 0) y[r] = s * sc;
    }
  }
}

void spmv_launch(eigkl_handle *h, const double *x, double *y, const double *scale_inv, double *store_scaled) {
  auto &L = h->L;
  EIGKL_REQUIRE(L.valid, EIGKL_E_ARG, "Laplacian not assembled");
  h->prof.begin(KC_SPMV, h->stream);
  spmv_adaptive_kernel<<<(unsigned)L.n_blocks, SPMV_THREADS, 0, h->stream>>>(L.rowptr.p, L.col.p, L.val.p, x, y, L.blk_row.p,
                                                                           scale_inv, store_scaled, 0);
  h->prof.end(h->stream);
  h->launches++;
}

}  // namespace eigkl
