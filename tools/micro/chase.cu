// Pointer-chase latency from a lone CTA at natural clocks (tuning aid for the KL swap loop's dependent-load chain).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/chase tools/micro/chase.cu && /tmp/chase
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void warm(const unsigned *p, size_t n, unsigned *out) {
  unsigned s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += p[i];
  if (s == 0x12345678u) *out = s;
}
// mode 0: ld.global.nc (L1 allocating)  1: ld.global.cg  2: ld.global.nc with 32 lanes chasing independent chains
template <int MODE>
__global__ void chase(const unsigned *p, unsigned start, int hops, long long *cycles, unsigned *sink) {
  unsigned i = start + (MODE == 2 ? threadIdx.x * 7919u : 0u);
  if (MODE != 2 && threadIdx.x != 0) return;
  const long long t0 = clock64();
  for (int h = 0; h < hops; ++h) {
    if (MODE == 1) i = __ldcg(p + i); else i = __ldg(p + i);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  if (i == 0xFFFFFFFFu) *sink = i;
}
__global__ void busy(volatile int *stop, unsigned *sink) {     // company CTAs: spin on arithmetic until told to stop
  unsigned x = threadIdx.x;
  while (!*stop) { for (int k = 0; k < 1000; ++k) x = x * 1664525u + 1013904223u; }
  if (x == 1) *sink = x;
}

int main() {
  const size_t sizes_mb[] = {8, 35, 100, 400};
  long long *cyc; unsigned *sink; int *stop;
  CK(cudaMallocManaged(&cyc, 8)); CK(cudaMalloc(&sink, 4)); CK(cudaMallocHost(&stop, 4));
  cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
  for (size_t mb : sizes_mb) {
    const size_t n = mb * 1024 * 1024 / 4;
    std::vector<unsigned> h(n);
    // random cyclic permutation with stride >= 128 B: each hop lands in a new line
    std::vector<unsigned> order(n / 64);
    for (size_t i = 0; i < order.size(); ++i) order[i] = (unsigned)i;
    srand(1);
    for (size_t i = order.size() - 1; i > 0; --i) { size_t j = ((size_t)rand() * 32768u + rand()) % (i + 1); std::swap(order[i], order[j]); }
    for (size_t i = 0; i < n; ++i) h[i] = (unsigned)((i + 64) % n);
    for (size_t i = 0; i < order.size(); ++i) h[(size_t)order[i] * 64] = order[(i + 1) % order.size()] * 64;
    unsigned *d; CK(cudaMalloc(&d, n * 4));
    CK(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice));
    for (int company = 0; company < 2; ++company) {
      *stop = 0;
      if (company) busy<<<147, 256, 0, s2>>>(stop, sink);
      for (int mode = 0; mode < 3; ++mode) {
        warm<<<592, 256, 0, s1>>>(d, n, sink);
        const int hops = 2000;
        for (int rep = 0; rep < 2; ++rep) {
          if (mode == 0) chase<0><<<1, 32, 0, s1>>>(d, order[0] * 64, hops, cyc, sink);
          if (mode == 1) chase<1><<<1, 32, 0, s1>>>(d, order[0] * 64, hops, cyc, sink);
          if (mode == 2) chase<2><<<1, 32, 0, s1>>>(d, order[0] * 64, hops, cyc, sink);
          CK(cudaStreamSynchronize(s1));
        }
        printf("buffer %4zu MB  %s  mode %d (%s): %.0f cycles per dependent hop\n", mb, company ? "147 busy CTAs" : "lone CTA     ", mode,
               mode == 0 ? "ld.nc 1 lane " : mode == 1 ? "ld.cg 1 lane " : "ld.nc 32 lanes", (double)*cyc / hops);
      }
      *stop = 1;
      CK(cudaDeviceSynchronize());
    }
    {   // warmed by OTHER SMs only: every line read once by a full-grid kernel, then ONE chase (no self-warming repeat)
      unsigned *junk; CK(cudaMalloc(&junk, (size_t)512 << 20));
      CK(cudaMemset(junk, 7, (size_t)512 << 20));
      warm<<<592, 256, 0, s1>>>(d, n, sink);
      chase<1><<<1, 32, 0, s1>>>(d, order[0] * 64, 2000, cyc, sink);
      CK(cudaStreamSynchronize(s1));
      printf("buffer %4zu MB  warmed by a full-grid read, first chase (ld.cg 1 lane): %.0f cycles per dependent hop\n", mb, (double)*cyc / 2000);
      chase<1><<<1, 32, 0, s1>>>(d, order[0] * 64, 2000, cyc, sink);
      CK(cudaStreamSynchronize(s1));
      printf("buffer %4zu MB  same chase again: %.0f cycles per dependent hop\n", mb, (double)*cyc / 2000);
      CK(cudaFree(junk));
    }
    {   // cold: flush L2 with a 512 MB write, then chase once (every hop is a DRAM access); and the same with a 64-bit
        // load by 32 lanes over 256 contiguous bytes (the KL item load)
      unsigned *junk; CK(cudaMalloc(&junk, (size_t)512 << 20));
      for (int mode = 0; mode < 3; ++mode) {
        CK(cudaMemset(junk, mode, (size_t)512 << 20));
        const int hops = 2000;
        if (mode == 0) chase<0><<<1, 32, 0, s1>>>(d, order[0] * 64, hops, cyc, sink);
        if (mode == 1) chase<1><<<1, 32, 0, s1>>>(d, order[0] * 64, hops, cyc, sink);
        if (mode == 2) chase<2><<<1, 32, 0, s1>>>(d, order[0] * 64, hops, cyc, sink);
        CK(cudaStreamSynchronize(s1));
        printf("buffer %4zu MB  cold (L2 flushed)  mode %d: %.0f cycles per dependent hop\n", mb, mode, (double)*cyc / hops);
      }
      CK(cudaFree(junk));
    }
    CK(cudaFree(d));
  }
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("SM clock (attribute) %d kHz\n", clk);
  return 0;
}
