// lanczos.cu -- fp64 thick-restart Lanczos Fiedler solver (north-star subsystem 2).
//
// Replaces the Spectra call in cEIG.cpp:194-207 (SymEigsSolver, nev = 2, ncv = min(100, n/2),
// SmallestAlge, tol 1e-10, maxit 1000): the two algebraically smallest eigenpairs of the clique
// Laplacian, of which the larger (lambda2, Fiedler vector) is reported.
//
// Device side, one Lanczos step j (all on one stream, no host round trip inside a restart cycle):
//   spmv      w' = L (w / beta_{j-1}),  v_j = w / beta_{j-1} stored by the same kernel      (spmv.cu)
//   multidot  h  = V_{0..j}^T w'        column groups x row chunks, warp-shuffle + block reduce,
//                                       last finishing CTA folds the partials in a fixed order
//   update    w' -= V_{0..j} h          one row per thread, coalesced across the basis columns
//   multidot, update again              ("twice is enough" full re-orthogonalisation), the second
//                                       update also reduces |w'|^2 -> beta_j, 1/beta_j
// Host side, once per cycle: ncv alphas/betas come back, the small projected eigenproblem is solved
// (dense_eig.cpp), convergence is tested as |beta_m y_last| < tol * max(eps^(2/3), |theta|) for
// both wanted pairs, and the basis is compressed to `keep` Ritz vectors by one tall-skinny
// V <- V Y kernel (restart).
// Every reduction runs in a fixed order, so results are bit-reproducible run to run.
// Bound: HBM/L2 bandwidth; the re-orthogonalisation streams 4*j*n*8 bytes per step, the SpMV
// nnz*12 + n*20.  No dense contraction worth a tensor core: the only GEMM-shaped piece is the
// n x ncv by ncv x keep restart, run once per ~80 steps.
#include "internal.h"
#include "device_utils.cuh"
#include <algorithm>
#include <cmath>
#include <cfloat>
#include <cstdlib>

namespace eigkl {

constexpr int LZ_THREADS = 256;
constexpr int MD_COLS = 8;          // basis columns per multidot CTA

// ---------------------------------------------------------------------------------------------------
__global__ void fill_random_kernel(double *__restrict__ w, int32_t n, uint64_t seed, int32_t row_offset) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(row_offset + i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  w[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

// block-level sum of one double per thread (256 threads); result valid in thread 0
__device__ __forceinline__ double block_sum_256(double v, double *sm /* 8 doubles */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = lane < (LZ_THREADS / 32) ? sm[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// h[c] = sum_i V[c*ld + i] * w[i], c in [0, ncols).
// grid = (row chunks of MD_ROWS, column groups of MD_COLS): every thread owns two adjacent rows (one
// 16-byte load per column) and MD_COLS columns, so a step with j columns exposes n/2 * j/8 threads with
// 9 independent loads each -- enough bytes in flight to cover the L2/HBM latency even at n = 12 K.
// partial[c * nbx + bx]; the last CTA of a column group (ticket counter per group) folds that group's
// partials, one warp per column, in a fixed order.
constexpr int MD_ROWS = 2 * LZ_THREADS;
__global__ void __launch_bounds__(LZ_THREADS)
multidot_kernel(const double *__restrict__ V, size_t ld, int ncols, const double *__restrict__ w, int32_t n,
                double *__restrict__ partial, unsigned int *__restrict__ counters, double *__restrict__ h_out,
                const int *__restrict__ skip_flag) {
  __shared__ double sm[LZ_THREADS / 32][MD_COLS];
  __shared__ bool am_last;
  if (skip_flag && *skip_flag) return;      // second Gram-Schmidt pass not needed (see update_kernel)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = blockIdx.y * MD_COLS;
  const int nc = min(MD_COLS, ncols - c0);
  const unsigned nbx = gridDim.x;
  const int32_t i = (blockIdx.x * LZ_THREADS + tid) * 2;
  double acc[MD_COLS];
#pragma unroll
  for (int g = 0; g < MD_COLS; ++g) acc[g] = 0.0;
  if (i + 1 < n) {
    const double2 w2 = *reinterpret_cast<const double2 *>(w + i);
    double2 v[MD_COLS];
#pragma unroll
    for (int g = 0; g < MD_COLS; ++g)
      if (g < nc) v[g] = *reinterpret_cast<const double2 *>(V + (size_t)(c0 + g) * ld + i);
#pragma unroll
    for (int g = 0; g < MD_COLS; ++g)
      if (g < nc) acc[g] = v[g].x * w2.x + v[g].y * w2.y;
  } else if (i < n) {
    const double w1 = w[i];
#pragma unroll
    for (int g = 0; g < MD_COLS; ++g)
      if (g < nc) acc[g] = V[(size_t)(c0 + g) * ld + i] * w1;
  }
#pragma unroll
  for (int g = 0; g < MD_COLS; ++g) {
    const double s = warp_sum(acc[g]);
    if (lane == 0) sm[warp][g] = s;
  }
  __syncthreads();
  if (tid < MD_COLS) {
    double s = 0.0;
#pragma unroll
    for (int wi = 0; wi < LZ_THREADS / 32; ++wi) s += sm[wi][tid];
    if (tid < nc) partial[(size_t)(c0 + tid) * nbx + blockIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) am_last = (atomicInc(counters + blockIdx.y, nbx - 1) == nbx - 1);
  __syncthreads();
  if (!am_last) return;
  __threadfence();
  if (warp < nc) {                        // one warp per column of this group
    const int c = c0 + warp;
    double s = 0.0;
    for (unsigned bx = lane; bx < nbx; bx += 32) s += __ldcg(&partial[(size_t)c * nbx + bx]);
    s = warp_sum(s);
    if (lane == 0) h_out[c] = s;
  }
}

// w[i] -= sum_c V[c*ld+i] * h[c], plus the reductions that finish a Gram-Schmidt pass.
//
// One row per thread; the column loop is unrolled 16-fold with all loads issued before the FMAs, so a
// thread keeps 16 independent 8-byte loads in flight (at n = 69 K that is 9 MB in flight chip-wide).
// (A variant with 8 column slices per row and one load round was measured 25% slower on ibm10: 111
// registers, 2 CTAs/SM, 4 waves.)
//
// pass 1 (DGKS test, the criterion ARPACK/Spectra-class solvers use): with V orthonormal,
//   |w_before|^2 = |w_after|^2 + |h|^2.  If |w_after| > eta * |w_before| the first pass lost at most
//   eps/eta of orthogonality and the second pass is skipped: flag = 1, beta_j = |w_after|, alpha_j = h[j].
//   Otherwise flag = 0 and pass 2 (multidot + update) runs and finalises alpha_j = h1[j] + h2[j], beta_j.
// pass 2 kernels return at once when flag == 1.
// single == 0 (multi-rank): the kernel only leaves its local |w|^2 in scal[0]; the decision / beta are
// taken by decide_kernel / beta_kernel after the all-reduce.
constexpr int UP_ROWS = LZ_THREADS;       // rows per CTA
constexpr int UP_UNROLL = 16;
__global__ void __launch_bounds__(LZ_THREADS)
update_kernel(const double *__restrict__ V, size_t ld, int ncols, double *__restrict__ w, int32_t n,
              const double *__restrict__ h, double *__restrict__ partial, unsigned int *__restrict__ counter,
              double *__restrict__ scal, double *__restrict__ beta_out, double *__restrict__ alpha_out,
              const double *__restrict__ h_prev, int j, int pass, int single, double eta2, int *__restrict__ flag) {
  extern __shared__ double hs[];            // ncols
  __shared__ double sm[8];
  __shared__ bool am_last;
  if (pass == 2 && *flag) return;           // second pass not needed (uniform for the whole grid)
  const int tid = threadIdx.x;
  for (int c = tid; c < ncols; c += LZ_THREADS) hs[c] = h[c];
  __syncthreads();
  double nrm = 0.0;
  for (int32_t i = blockIdx.x * LZ_THREADS + tid; i < n; i += gridDim.x * LZ_THREADS) {
    double s = w[i];
    const double *vp = V + i;
    int c = 0;
    for (; c + UP_UNROLL <= ncols; c += UP_UNROLL) {
      double v[UP_UNROLL];
#pragma unroll
      for (int u = 0; u < UP_UNROLL; ++u) v[u] = vp[(size_t)(c + u) * ld];
#pragma unroll
      for (int u = 0; u < UP_UNROLL; ++u) s -= v[u] * hs[c + u];
    }
    for (; c + 4 <= ncols; c += 4) {
      const double v0 = vp[(size_t)(c + 0) * ld], v1 = vp[(size_t)(c + 1) * ld];
      const double v2 = vp[(size_t)(c + 2) * ld], v3 = vp[(size_t)(c + 3) * ld];
      s -= v0 * hs[c + 0]; s -= v1 * hs[c + 1]; s -= v2 * hs[c + 2]; s -= v3 * hs[c + 3];
    }
    for (; c < ncols; ++c) s -= vp[(size_t)c * ld] * hs[c];
    w[i] = s;
    nrm += s * s;
  }
  const double bs = block_sum_256(nrm, sm);
  if (tid == 0) partial[blockIdx.x] = bs;
  __threadfence();
  if (tid == 0) am_last = (atomicInc(counter, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (!am_last) return;
  __threadfence();
  double s = 0.0;
  for (unsigned b = tid; b < gridDim.x; b += LZ_THREADS) s += __ldcg(&partial[b]);   // fixed-order fold
  s = block_sum_256(s, sm);
  double hh = 0.0;
  if (pass == 1) {
    for (int c = tid; c < ncols; c += LZ_THREADS) hh += hs[c] * hs[c];
    hh = block_sum_256(hh, sm);
  }
  if (tid == 0) {
    scal[0] = s;
    if (single) {
      const bool skip = (pass == 1) && (s > eta2 * (s + hh));
      if (pass == 1) *flag = skip ? 1 : 0;
      if (pass == 2 || skip) {
        const double beta = sqrt(s);
        scal[1] = beta > 0.0 ? 1.0 / beta : 0.0;
        beta_out[j] = beta;
        alpha_out[j] = (pass == 2) ? h_prev[j] + hs[j] : hs[j];
      }
    }
  }
}

// multi-rank, after the all-reduce of |w|^2 (scal[0]): the pass-1 decision ...
__global__ void decide_kernel(double *__restrict__ scal, const double *__restrict__ h1, int ncols, double eta2,
                              int *__restrict__ flag, double *__restrict__ beta_out, double *__restrict__ alpha_out, int j) {
  __shared__ double sm[8];
  double hh = 0.0;
  for (int c = threadIdx.x; c < ncols; c += LZ_THREADS) hh += h1[c] * h1[c];
  hh = block_sum_256(hh, sm);
  if (threadIdx.x == 0) {
    const double s = scal[0];
    const bool skip = s > eta2 * (s + hh);
    *flag = skip ? 1 : 0;
    if (skip) {
      const double beta = sqrt(s);
      scal[1] = beta > 0.0 ? 1.0 / beta : 0.0;
      beta_out[j] = beta;
      alpha_out[j] = h1[j];
    }
  }
}
// ... and the pass-2 finalisation (also used for the norm of the start / Ritz vector: flag == nullptr)
__global__ void beta_kernel(double *__restrict__ scal, double *__restrict__ beta_out, double *__restrict__ alpha_out,
                            const double *__restrict__ h1, const double *__restrict__ h2, int j, const int *__restrict__ flag) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (flag && *flag) return;
    const double beta = sqrt(scal[0]);
    scal[1] = beta > 0.0 ? 1.0 / beta : 0.0;
    if (beta_out) beta_out[j] = beta;
    if (alpha_out) alpha_out[j] = h1[j] + h2[j];
  }
}

// norm of w -> scal[0] = |w|^2, scal[1] = 1/|w|   (start vector)
__global__ void __launch_bounds__(LZ_THREADS)
norm_kernel(const double *__restrict__ w, int32_t n, double *__restrict__ partial, unsigned int *__restrict__ counter,
            double *__restrict__ scal) {
  __shared__ double sm[8];
  __shared__ bool am_last;
  double nrm = 0.0;
  for (int32_t i = blockIdx.x * LZ_THREADS + threadIdx.x; i < n; i += gridDim.x * LZ_THREADS) nrm += w[i] * w[i];
  const double bs = block_sum_256(nrm, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = bs;
  __threadfence();
  if (threadIdx.x == 0) am_last = (atomicInc(counter, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (!am_last) return;
  __threadfence();
  double s = 0.0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += LZ_THREADS) s += __ldcg(&partial[b]);
  s = block_sum_256(s, sm);
  if (threadIdx.x == 0) { scal[0] = s; scal[1] = s > 0.0 ? 1.0 / sqrt(s) : 0.0; }
}

__global__ void scale_store_kernel(const double *__restrict__ w, const double *__restrict__ scale, double *__restrict__ out, int32_t n) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = w[i] * __ldg(scale);
}

// out[:, c] = sum_j Vin[:, j] * Y[j, c]  for c in [c0, c0+RS_COLS) (Y column-major m x kk, ld m)
constexpr int RS_COLS = 16;
__global__ void __launch_bounds__(LZ_THREADS)
restart_kernel(const double *__restrict__ Vin, size_t ld, int m, const double *__restrict__ Y, int kk,
               double *__restrict__ Vout, size_t ld_out, int32_t n) {
  extern __shared__ double Ys[];            // m * RS_COLS
  const int c0 = blockIdx.y * RS_COLS;
  const int nc = min(RS_COLS, kk - c0);
  for (int t = threadIdx.x; t < m * RS_COLS; t += LZ_THREADS) {
    const int j = t / RS_COLS, g = t % RS_COLS;
    Ys[t] = (g < nc) ? Y[(size_t)(c0 + g) * m + j] : 0.0;
  }
  __syncthreads();
  const int32_t i = blockIdx.x * LZ_THREADS + threadIdx.x;
  if (i >= n) return;
  double acc[RS_COLS];
#pragma unroll
  for (int g = 0; g < RS_COLS; ++g) acc[g] = 0.0;
  for (int j = 0; j < m; ++j) {
    const double v = Vin[(size_t)j * ld + i];
#pragma unroll
    for (int g = 0; g < RS_COLS; ++g) acc[g] += v * Ys[j * RS_COLS + g];
  }
#pragma unroll
  for (int g = 0; g < RS_COLS; ++g)
    if (g < nc) Vout[(size_t)(c0 + g) * ld_out + i] = acc[g];
}

// ---------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------
namespace {

struct LzCtx {
  eigkl_handle *h;
  int32_t n;        // global rows
  int32_t nl;       // rows owned by this rank
  int32_t row_lo;
  int R;            // ranks
  int m;
  size_t ld;        // leading dimension of the local basis slice (= n_pad)
  int gx_md, gx_up;
  double eta2;      // DGKS threshold squared
};

void launch_multidot(LzCtx &c, const double *V, int ncols, const double *w, double *h_out, int pass) {
  auto &e = c.h->eig;
  if (c.nl > 0) {
    dim3 grid((unsigned)c.gx_md, (unsigned)ceil_div(ncols, MD_COLS));
    c.h->prof.begin(KC_MULTIDOT, c.h->stream);
    multidot_kernel<<<grid, LZ_THREADS, 0, c.h->stream>>>(V, c.ld, ncols, w, c.nl, e.partial.p, e.counters.p + 8, h_out,
                                                        pass == 2 ? e.flag.p : nullptr);
    c.h->prof.end(c.h->stream);
    c.h->launches++;
    c.h->stats.bytes_multidot_total += c.h->prof.on ? ((double)ncols * c.nl * 8.0 + (double)c.nl * 8.0) : 0.0;
  } else {
    EIGKL_CUDA(cudaMemsetAsync(h_out, 0, (size_t)ncols * sizeof(double), c.h->stream));
  }
  if (c.R > 1) comm_allreduce_sum_f64(c.h, h_out, (size_t)ncols);      // C2: Lanczos dot products
}
void launch_update(LzCtx &c, const double *V, int ncols, double *w, const double *hcoef, const double *h_prev, int j, int pass) {
  auto &e = c.h->eig;
  const int single = c.R == 1 ? 1 : 0;
  if (c.nl > 0) {
    c.h->prof.begin(KC_UPDATE, c.h->stream);
    update_kernel<<<(unsigned)c.gx_up, LZ_THREADS, (size_t)ncols * sizeof(double), c.h->stream>>>(
        V, c.ld, ncols, w, c.nl, hcoef, e.partial.p, e.counters.p + 1, e.scal.p, e.beta.p, e.alpha.p, h_prev, j, pass, single,
        c.eta2, e.flag.p);
    c.h->prof.end(c.h->stream);
    c.h->launches++;
    c.h->stats.bytes_update_total += c.h->prof.on ? ((double)ncols * c.nl * 8.0 + (double)c.nl * 16.0) : 0.0;
  } else {
    EIGKL_CUDA(cudaMemsetAsync(e.scal.p, 0, sizeof(double), c.h->stream));
  }
  if (!single) {
    comm_allreduce_sum_f64(c.h, e.scal.p, 1);
    if (pass == 1) decide_kernel<<<1, LZ_THREADS, 0, c.h->stream>>>(e.scal.p, hcoef, ncols, c.eta2, e.flag.p, e.beta.p, e.alpha.p, j);
    else beta_kernel<<<1, 32, 0, c.h->stream>>>(e.scal.p, e.beta.p, e.alpha.p, h_prev, hcoef, j, e.flag.p);
    c.h->launches++;
  }
}
void launch_norm(LzCtx &c, const double *w) {
  auto &e = c.h->eig;
  if (c.nl > 0) {
    norm_kernel<<<(unsigned)c.gx_up, LZ_THREADS, 0, c.h->stream>>>(w, c.nl, e.partial.p, e.counters.p + 1, e.scal.p);
    c.h->launches++;
  } else {
    EIGKL_CUDA(cudaMemsetAsync(e.scal.p, 0, sizeof(double), c.h->stream));
  }
  if (c.R > 1) {
    comm_allreduce_sum_f64(c.h, e.scal.p, 1);
    beta_kernel<<<1, 32, 0, c.h->stream>>>(e.scal.p, nullptr, nullptr, nullptr, nullptr, 0, nullptr);
    c.h->launches++;
  }
}

// one Lanczos step j: x_local (this rank's slice of the un-normalised v_j, with *scale = 1/|v_j|) -> V[:, j];
// leaves the un-normalised v_{j+1} slice in w_out and 1/beta_j in scal[1]
void lanczos_step(LzCtx &c, double *V, int j, const double *x_local, const double *scale, bool store, double *w_out) {
  auto &e = c.h->eig;
  const double *x = x_local;
  if (c.R > 1) {                                                      // C1: SpMV halo exchange
    comm_allgather_f64(c.h, x_local, e.xfull.p, c.ld);
    x = e.xfull.p;
  }
  spmv_launch(c.h, x, w_out, scale, store ? V + (size_t)j * c.ld : nullptr);
  double *h1 = e.hcoef.p, *h2 = e.hcoef.p + (c.m + 1);
  launch_multidot(c, V, j + 1, w_out, h1, 1);
  launch_update(c, V, j + 1, w_out, h1, nullptr, j, 1);      // sets flag = 1 when the second pass can be skipped
  launch_multidot(c, V, j + 1, w_out, h2, 2);
  launch_update(c, V, j + 1, w_out, h2, h1, j, 2);
}

}  // namespace

void fiedler_solve(eigkl_handle *h) {
  auto &L = h->L;
  auto &e = h->eig;
  EIGKL_REQUIRE(L.valid, EIGKL_E_ARG, "eigkl_fiedler: call eigkl_assemble_laplacian first");
  const int32_t n = L.n;
  const int nev = 2;
  int m = h->opts.ncv > 0 ? h->opts.ncv : std::min(100, n / 2);        // cEIG.cpp:195
  EIGKL_REQUIRE(m > nev && m <= n, EIGKL_E_ARG, "eigkl_fiedler: need nev < ncv <= n (graph too small)");
  const double tol = h->opts.tol > 0 ? h->opts.tol : 1e-10;
  const int maxit = h->opts.max_restarts > 0 ? h->opts.max_restarts : 1000;
  const double eps23 = std::pow(DBL_EPSILON, 2.0 / 3.0);
  cudaStream_t st = h->stream;

  LzCtx c;
  c.h = h; c.n = n; c.m = m; c.R = h->opts.nranks;
  int32_t lo, hi, n_pad;
  row_partition(n, c.R, h->opts.rank, &lo, &hi, &n_pad);
  EIGKL_REQUIRE(lo == L.row_lo && hi == L.row_hi, EIGKL_E_ARG, "row partition changed since assembly");
  c.nl = hi - lo; c.row_lo = lo;
  c.ld = (size_t)n_pad;
  c.gx_md = (int)std::max<int64_t>(1, ceil_div(c.nl, MD_ROWS));
  c.gx_up = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(c.nl, UP_ROWS), 8 * h->sm_count));
  {
    // second Gram-Schmidt pass only when |w_after| <= eta |w_before| (eta = 1/sqrt(2) is the classical
    // "twice is enough" bound; EIGKL_DGKS_ETA overrides, 1.0 forces both passes every step)
    double eta = 0.70710678118654752;
    if (const char *ev = getenv("EIGKL_DGKS_ETA")) eta = atof(ev);
    c.eta2 = eta * eta;
  }
  e.n = n; e.ncv = m; e.ld = c.ld;
  for (int b = 0; b < 2; ++b) { e.V[b].ensure(c.ld * (size_t)(m + 1)); e.w[b].ensure(c.ld); }
  e.partial.ensure((size_t)std::max<int64_t>((int64_t)c.gx_md * (m + 1), c.gx_up) + 8);
  e.hcoef.ensure(2 * (size_t)(m + 1));
  e.alpha.ensure((size_t)m); e.beta.ensure((size_t)m);
  e.scal.ensure(8);
  e.flag.ensure(2);
  EIGKL_CUDA(cudaMemsetAsync(e.flag.p, 0, 2 * sizeof(int), st));
  const int n_counters = 8 + (m + 1 + MD_COLS - 1) / MD_COLS + 1;
  e.counters.ensure((size_t)n_counters);
  e.Y.ensure((size_t)m * m);
  e.fiedler.ensure(c.ld * (size_t)c.R);
  if (c.R > 1) e.xfull.ensure(c.ld * (size_t)c.R);
  EIGKL_CUDA(cudaMemsetAsync(e.counters.p, 0, (size_t)n_counters * sizeof(unsigned int), st));
  for (int b = 0; b < 2; ++b) EIGKL_CUDA(cudaMemsetAsync(e.w[b].p, 0, c.ld * sizeof(double), st));
  const double one = 1.0;
  EIGKL_CUDA(cudaMemcpyAsync(e.scal.p + 2, &one, sizeof(double), cudaMemcpyHostToDevice, st));   // scal[2] = 1.0

  // start vector (Spectra: SimpleRandom residual, uniform in [-0.5, 0.5); ours is a seeded splitmix64 of the
  // GLOBAL row id, so the vector does not depend on the number of ranks)
  if (c.nl > 0) {
    fill_random_kernel<<<(unsigned)ceil_div(c.nl, LZ_THREADS), LZ_THREADS, 0, st>>>(e.w[0].p, c.nl, h->opts.seed + 0x9E3779B97F4A7C15ull, lo);
    h->launches++;
  }
  launch_norm(c, e.w[0].p);

  std::vector<double> T((size_t)m * m, 0.0), Yh((size_t)m * m), theta(m), alpha(m), beta(m), Ycm;
  int k = 0, cur = 0, bank = 0, it = 0, nmv = 0;
  double res[2] = {0, 0};
  bool converged = false;
  double beta_m = 0.0;
  for (it = 0; it < maxit; ++it) {
    double *V = e.V[bank].p;
    for (int j = k; j < m; ++j) {
      if (j == k && it > 0) {
        // right after a restart V[:, k] already holds the normalised v_m
        lanczos_step(c, V, j, V + (size_t)k * c.ld, e.scal.p + 2, false, e.w[cur ^ 1].p);
      } else {
        lanczos_step(c, V, j, e.w[cur].p, e.scal.p + 1, true, e.w[cur ^ 1].p);
      }
      cur ^= 1;
      ++nmv;
    }
    // v_m = w / beta_m into column m
    if (c.nl > 0) {
      scale_store_kernel<<<(unsigned)ceil_div(c.nl, LZ_THREADS), LZ_THREADS, 0, st>>>(e.w[cur].p, e.scal.p + 1, V + (size_t)m * c.ld, c.nl);
      h->launches++;
    }
    EIGKL_CUDA(cudaMemcpyAsync(alpha.data(), e.alpha.p, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
    EIGKL_CUDA(cudaMemcpyAsync(beta.data(), e.beta.p, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
    EIGKL_CUDA(cudaStreamSynchronize(st));
    for (int j = k; j < m; ++j) {
      T[(size_t)j * m + j] = alpha[j];
      if (j + 1 < m) { T[(size_t)j * m + j + 1] = beta[j]; T[(size_t)(j + 1) * m + j] = beta[j]; }
    }
    beta_m = beta[m - 1];
    EIGKL_REQUIRE(std::isfinite(beta_m), EIGKL_E_NOCONV, "eigkl_fiedler: Lanczos breakdown (non-finite beta)");
    Yh = T;
    sym_eig(m, Yh.data(), theta.data());
    int nconv = 0;
    for (int i = 0; i < nev; ++i) {
      res[i] = std::fabs(beta_m * Yh[(size_t)(m - 1) * m + i]);
      if (res[i] < tol * std::max(eps23, std::fabs(theta[i]))) ++nconv;
    }
    if (nconv == nev) { converged = true; ++it; break; }
    if (it == maxit - 1) { ++it; break; }
    int kk = h->opts.keep > 0 ? h->opts.keep : std::max(nev + nconv, m / 5);
    kk = std::max(nev, std::min(kk, m - 2));
    // V_new[:, 0:kk] = V[:, 0:m] Y[:, 0:kk];  V_new[:, kk] = v_m   (rank-local: no communication)
    Ycm.assign((size_t)m * kk, 0.0);
    for (int cc = 0; cc < kk; ++cc)
      for (int j = 0; j < m; ++j) Ycm[(size_t)cc * m + j] = Yh[(size_t)j * m + cc];
    EIGKL_CUDA(cudaMemcpyAsync(e.Y.p, Ycm.data(), Ycm.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    double *Vn = e.V[bank ^ 1].p;
    if (c.nl > 0) {
      dim3 grid((unsigned)ceil_div(c.nl, LZ_THREADS), (unsigned)ceil_div(kk, RS_COLS));
      h->prof.begin(KC_RESTART, st);
      restart_kernel<<<grid, LZ_THREADS, (size_t)m * RS_COLS * sizeof(double), st>>>(V, c.ld, m, e.Y.p, kk, Vn, c.ld, c.nl);
      h->prof.end(st);
      h->launches++;
    }
    EIGKL_CUDA(cudaMemcpyAsync(Vn + (size_t)kk * c.ld, V + (size_t)m * c.ld, c.ld * sizeof(double), cudaMemcpyDeviceToDevice, st));
    EIGKL_CUDA(cudaStreamSynchronize(st));    // Ycm is reused next cycle
    std::fill(T.begin(), T.end(), 0.0);
    for (int cc = 0; cc < kk; ++cc) {
      T[(size_t)cc * m + cc] = theta[cc];
      const double s = beta_m * Yh[(size_t)(m - 1) * m + cc];
      T[(size_t)kk * m + cc] = s;
      T[(size_t)cc * m + kk] = s;
    }
    bank ^= 1;
    k = kk;
  }
  // Ritz vector of the larger wanted value (index 1): cEIG.cpp:205-207 takes evalues(0)/evecs.col(0)
  // of Spectra's descending result, i.e. lambda2
  {
    Ycm.assign((size_t)m, 0.0);
    for (int j = 0; j < m; ++j) Ycm[j] = Yh[(size_t)j * m + 1];
    EIGKL_CUDA(cudaMemcpyAsync(e.Y.p, Ycm.data(), (size_t)m * sizeof(double), cudaMemcpyHostToDevice, st));
    double *slice = (c.R > 1) ? e.w[1].p : e.fiedler.p;
    if (c.nl > 0) {
      dim3 grid((unsigned)ceil_div(c.nl, LZ_THREADS), 1);
      restart_kernel<<<grid, LZ_THREADS, (size_t)m * RS_COLS * sizeof(double), st>>>(e.V[bank].p, c.ld, m, e.Y.p, 1, e.w[0].p, c.ld, c.nl);
      h->launches++;
    }
    launch_norm(c, e.w[0].p);
    if (c.nl > 0) {
      scale_store_kernel<<<(unsigned)ceil_div(c.nl, LZ_THREADS), LZ_THREADS, 0, st>>>(e.w[0].p, e.scal.p + 1, slice, c.nl);
      h->launches++;
    }
    if (c.R > 1) comm_allgather_f64(h, slice, e.fiedler.p, c.ld);     // every rank ends with the full vector
    EIGKL_CUDA(cudaStreamSynchronize(st));
  }
  EIGKL_CUDA(cudaGetLastError());
  e.lambda2 = theta[1];
  e.have_vector = true;
  e.have_median = false;
  e.bank = bank;
  auto &s = h->stats;
  s.ncv = m; s.matvecs = nmv; s.restarts = it; s.converged = converged ? 1 : 0;
  s.resid_est[0] = res[0]; s.resid_est[1] = res[1];
  s.lambda[0] = theta[0]; s.lambda[1] = theta[1];
  if (!converged) throw Error(EIGKL_E_NOCONV, "eigkl_fiedler: not converged within max_restarts");
}

// ---------------------------------------------------------------------------------------------------
// median + sides (cEIG.cpp:55-65, 218) on the device: radix sort of the order-preserving bit pattern
// ---------------------------------------------------------------------------------------------------
__global__ void median_keys_kernel(const double *__restrict__ v, int32_t n, unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = double_orderable(v[i]); vals[i] = (uint32_t)i; }
}
__global__ void median_pick_kernel(const unsigned long long *__restrict__ sorted, int32_t n, double *__restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (n % 2 != 0) out[0] = double_from_orderable(sorted[n / 2]);
    else out[0] = (double_from_orderable(sorted[(n - 1) / 2]) + double_from_orderable(sorted[n / 2])) / 2.0;
  }
}
__global__ void side_kernel(const double *__restrict__ v, int32_t n, const double *__restrict__ median, uint8_t *__restrict__ side) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) side[i] = (__ldg(median) > v[i]) ? 1 : 0;            // cEIG.cpp:218
}

void partition_from_fiedler(eigkl_handle *h) {
  auto &e = h->eig;
  EIGKL_REQUIRE(e.have_vector, EIGKL_E_ARG, "no Fiedler vector: call eigkl_fiedler first");
  const int32_t n = e.n;
  cudaStream_t st = h->stream;
  for (int i = 0; i < 2; ++i) { e.sortkey[i].ensure((size_t)n + 1); e.sortval[i].ensure((size_t)n + 1); }
  e.side.ensure((size_t)n);
  unsigned long long *keys[2] = {e.sortkey[0].p, e.sortkey[1].p};
  uint32_t *vals[2] = {e.sortval[0].p, e.sortval[1].p};
  median_keys_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(e.fiedler.p, n, keys[0], vals[0]);
  const int cur = radix_sort_kv(h, keys, vals, n, 64);
  median_pick_kernel<<<1, 32, 0, st>>>(keys[cur], n, e.scal.p + 3);
  side_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(e.fiedler.p, n, e.scal.p + 3, e.side.p);
  h->launches += 3;
  EIGKL_CUDA(cudaMemcpyAsync(&e.median, e.scal.p + 3, sizeof(double), cudaMemcpyDeviceToHost, st));
  EIGKL_CUDA(cudaStreamSynchronize(st));
  EIGKL_CUDA(cudaGetLastError());
  e.have_median = true;
}

}  // namespace eigkl
