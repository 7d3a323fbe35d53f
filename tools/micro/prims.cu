// Latency of the on-chip primitives the KL swap loop chains together, from one CTA of 512 threads (tuning aid).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/prims tools/micro/prims.cu && /tmp/prims
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__global__ void __launch_bounds__(512, 1) prims(long long *out, unsigned *sink, int active_warps) {
  __shared__ unsigned sm[1024];
  __shared__ unsigned long long sm64[64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 1024; i += 512) sm[i] = (i * 7 + 1) & 1023;
  if (tid < 64) sm64[tid] = 0;
  __syncthreads();
  unsigned x = tid;
  long long t0, t1;
  if (warp < active_warps) {
    // 0: dependent redux.sync.max
    t0 = clock64();
    for (int i = 0; i < N; ++i) x = __reduce_max_sync(0xffffffffu, x + lane) ^ 1u;
    t1 = clock64(); if (tid == 0) out[0] = t1 - t0;
    // 1: dependent shfl
    t0 = clock64();
    for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1u;
    t1 = clock64(); if (tid == 0) out[1] = t1 - t0;
    // 2: dependent LDS
    x &= 1023u;
    t0 = clock64();
    for (int i = 0; i < N; ++i) x = sm[x];
    t1 = clock64(); if (tid == 0) out[2] = t1 - t0;
    // 3: dependent ballot
    t0 = clock64();
    for (int i = 0; i < N; ++i) x = __ballot_sync(0xffffffffu, (x >> lane) & 1u) + 3u;
    t1 = clock64(); if (tid == 0) out[3] = t1 - t0;
    // 4: shared atomicMax u64 (returning), lane-distinct addresses
    unsigned long long y = x;
    t0 = clock64();
    for (int i = 0; i < N; ++i) y = atomicMax(&sm64[(lane + (unsigned)y) & 63], y + i) + 1ull;
    t1 = clock64(); if (tid == 0) out[4] = t1 - t0;
    x += (unsigned)y;
    // 5: dependent FADD
    float f = (float)x;
    t0 = clock64();
    for (int i = 0; i < N; ++i) f = __fadd_rn(f, 1.0f);
    t1 = clock64(); if (tid == 0) out[5] = t1 - t0;
    x += (unsigned)f;
    // 6: 64-bit butterfly max (5 rounds of 2 shuffles)
    y = x * 0x9E3779B97F4A7C15ull;
    t0 = clock64();
    for (int i = 0; i < N / 8; ++i) {
      for (int o = 16; o > 0; o >>= 1) { const unsigned long long z = __shfl_xor_sync(0xffffffffu, y, o); y = z > y ? z : y; }
      y += lane;
    }
    t1 = clock64(); if (tid == 0) out[6] = (t1 - t0) * 8;
    x += (unsigned)y;
  }
  // 7: block barrier, all 16 warps
  __syncthreads();
  t0 = clock64();
  for (int i = 0; i < N; ++i) __syncthreads();
  t1 = clock64(); if (tid == 0) out[7] = t1 - t0;
  // 8: clock64 read + shared accumulate (the probe the KL loop uses)
  t0 = clock64();
  for (int i = 0; i < N; ++i) { const long long t = clock64(); sm64[1] += t - (long long)sm64[2]; sm64[2] = t; }
  t1 = clock64(); if (tid == 0) out[8] = t1 - t0;
  if (x == 0x12345u) *sink = x;
}
int main() {
  long long *out; unsigned *sink;
  cudaMallocManaged(&out, 16 * 8); cudaMalloc(&sink, 4);
  const char *nm[] = {"redux.sync.max u32", "shfl.xor", "LDS (dependent)", "ballot", "shared atomicMax u64", "FADD", "64-bit butterfly max (whole)", "__syncthreads (512 thr)", "clock probe"};
  for (int aw : {1, 4, 16}) {
    prims<<<1, 512>>>(out, sink, aw);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed\n"); return 1; }
    printf("active warps %2d:", aw);
    for (int i = 0; i < 9; ++i) printf("  %s %.1f", nm[i], (double)out[i] / N);
    printf("\n");
  }
  return 0;
}
