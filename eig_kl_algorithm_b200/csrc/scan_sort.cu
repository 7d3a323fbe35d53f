// scan_sort.cu -- hand-written device primitives used by the assembly stage:
//   * hierarchical exclusive scan (int32 / int64)
//   * stable LSD radix sort of (uint64 key, uint32 value) pairs, 8 bits per pass
// Both are HBM/L2-bound integer kernels; nothing here is GEMM-shaped.
#include "internal.h"
#include "device_utils.cuh"

namespace eigkl {

// ------------------------------------------------------------------------------------------------
// exclusive scan: out[i] = sum_{k<i} in[k], i in [0, n]   (n+1 outputs, out[n] = total)
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_kernel(const T *in, T *out, T *__restrict__ tile_total, int64_t n) {
  __shared__ T warp_tot[SCAN_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  T v[SCAN_ITEMS];
  T run = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? in[base + i] : (T)0;
    run += v[i];
  }
  // inclusive scan of thread totals within the warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(FULL_MASK, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  T woff = 0;
  for (int w = 0; w < warp; ++w) woff += warp_tot[w];
  T excl = woff + inc - run;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = excl;
    excl += v[i];
  }
  if (threadIdx.x == SCAN_THREADS - 1) tile_total[blockIdx.x] = woff + inc;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(T *__restrict__ out, const T *__restrict__ tile_off, int64_t n, int64_t n_tiles) {
  // adds the scanned tile totals; the thread owning index n also writes the grand total out[n]
  const int64_t tile = blockIdx.x;
  const T off = tile_off[tile];
  const int64_t base = tile * SCAN_TILE;
  for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_THREADS) {
    int64_t idx = base + i;
    if (idx < n) out[idx] += off;
  }
  if (tile == n_tiles - 1 && threadIdx.x == 0) out[n] = tile_off[n_tiles];
}

template <typename T>
static void exclusive_scan_impl(eigkl_handle *h, const T *in, T *out, int64_t n, T *tmp, size_t tmp_elems) {
  if (n <= 0) {
    EIGKL_CUDA(cudaMemsetAsync(out, 0, sizeof(T), h->stream));
    return;
  }
  const int64_t n_tiles = ceil_div(n, SCAN_TILE);
  EIGKL_REQUIRE((size_t)(2 * n_tiles + 2) <= tmp_elems, EIGKL_E_ARG, "scan scratch too small");
  T *tot = tmp;                      // n_tiles (+1 after scanning)
  T *rest = tmp + n_tiles + 1;
  scan_tile_kernel<T><<<(unsigned)n_tiles, SCAN_THREADS, 0, h->stream>>>(in, out, tot, n);
  h->launches++;
  if (n_tiles == 1) {
    // out[n] = total
    EIGKL_CUDA(cudaMemcpyAsync(out + n, tot, sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
    return;
  }
  exclusive_scan_impl<T>(h, tot, tot, n_tiles, rest, tmp_elems - (size_t)(n_tiles + 1));
  scan_add_kernel<T><<<(unsigned)n_tiles, SCAN_THREADS, 0, h->stream>>>(out, tot, n, n_tiles);
  h->launches++;
  EIGKL_CUDA(cudaGetLastError());
}

static size_t scan_tmp_elems(int64_t n) {
  size_t tot = 0;
  int64_t t = n;
  do {
    t = ceil_div(t, SCAN_TILE);
    tot += (size_t)(2 * t + 4);
  } while (t > 1);
  return tot + 8;
}

void exclusive_scan_i64(eigkl_handle *h, const int64_t *in, int64_t *out, int64_t n) {
  size_t need = scan_tmp_elems(n);
  h->scan_tmp.ensure(need);
  exclusive_scan_impl<int64_t>(h, in, out, n, h->scan_tmp.p, h->scan_tmp.n);
}
void exclusive_scan_i32(eigkl_handle *h, const int32_t *in, int32_t *out, int64_t n) {
  size_t need = scan_tmp_elems(n);
  h->scan_tmp.ensure(need);   // int64 storage reused as int32
  exclusive_scan_impl<int32_t>(h, in, out, n, reinterpret_cast<int32_t *>(h->scan_tmp.p), h->scan_tmp.n * 2);
}

int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b) != 0) ++b;
  return b;
}

// ------------------------------------------------------------------------------------------------
// radix sort
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const unsigned long long *__restrict__ keys, int64_t n, int shift, int32_t *__restrict__ hist,
               int64_t n_tiles) {
  __shared__ int32_t cnt[RS_RADIX];
  cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int r = 0; r < RS_ITEMS; ++r) {
    int64_t idx = base + (int64_t)r * RS_THREADS + threadIdx.x;
    if (idx < n) atomicAdd(&cnt[(unsigned)(keys[idx] >> shift) & 255u], 1);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = cnt[threadIdx.x];   // digit-major
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const unsigned long long *__restrict__ kin, const uint32_t *__restrict__ vin,
                  unsigned long long *__restrict__ kout, uint32_t *__restrict__ vout, int64_t n, int shift,
                  const int32_t *__restrict__ hist_scanned, int64_t n_tiles) {
  __shared__ int32_t base[RS_RADIX];       // global start of this tile's run of each digit
  __shared__ int32_t running[RS_RADIX];    // elements of each digit already placed by earlier rounds
  __shared__ int32_t wcnt[RS_WARPS][RS_RADIX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  base[tid] = hist_scanned[(int64_t)tid * n_tiles + blockIdx.x];
  running[tid] = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) wcnt[w][tid] = 0;
  __syncthreads();
  const int64_t tbase = (int64_t)blockIdx.x * RS_TILE;
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t idx = tbase + (int64_t)r * RS_THREADS + tid;
    const bool valid = idx < n;
    unsigned long long k = valid ? kin[idx] : 0ull;
    uint32_t v = valid ? vin[idx] : 0u;
    const unsigned digit = valid ? ((unsigned)(k >> shift) & 255u) : 256u;
    const unsigned peers = __match_any_sync(FULL_MASK, digit);
    const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank_in_warp == 0) wcnt[warp][digit] = __popc(peers);
    __syncthreads();
    if (valid) {
      int pre = 0;
      for (int w = 0; w < warp; ++w) pre += wcnt[w][digit];
      const int64_t pos = (int64_t)base[digit] + running[digit] + pre + rank_in_warp;
      kout[pos] = k;
      vout[pos] = v;
    }
    __syncthreads();
    {
      int tot = 0;
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) { tot += wcnt[w][tid]; wcnt[w][tid] = 0; }
      running[tid] += tot;
    }
    __syncthreads();
  }
}

int radix_sort_kv(eigkl_handle *h, unsigned long long *keys[2], uint32_t *vals[2], int64_t n, int nbits) {
  if (n <= 1 || nbits <= 0) return 0;
  EIGKL_REQUIRE(n < (int64_t)2147483647, EIGKL_E_ARG, "radix sort: more than 2^31-1 elements");
  const int64_t n_tiles = ceil_div(n, RS_TILE);
  const int64_t hn = n_tiles * RS_RADIX;
  h->sort_hist.ensure((size_t)hn + 1);
  int cur = 0;
  for (int shift = 0; shift < nbits; shift += 8) {
    rs_hist_kernel<<<(unsigned)n_tiles, RS_THREADS, 0, h->stream>>>(keys[cur], n, shift, h->sort_hist.p, n_tiles);
    h->launches++;
    exclusive_scan_i32(h, h->sort_hist.p, h->sort_hist.p, hn);
    rs_scatter_kernel<<<(unsigned)n_tiles, RS_THREADS, 0, h->stream>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1],
                                                                     n, shift, h->sort_hist.p, n_tiles);
    h->launches++;
    cur ^= 1;
  }
  EIGKL_CUDA(cudaGetLastError());
  return cur;
}

}  // namespace eigkl
