// Dependent-load latency on this GPU as a function of footprint (pointer chase, one warp, one lane active) and the
// latency of one "gather row, then gather children" trip pattern as used by the select walk.  Diagnostic only.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__global__ void chase(const unsigned* __restrict__ next, int steps, unsigned start, long long* cycles, unsigned* sink) {
  unsigned i = start;
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) i = next[i];
  long long t1 = clock64();
  if (threadIdx.x == 0) { *cycles = t1 - t0; *sink = i; }
}

// many warps chasing independent chains at once (like 1024 trees): per-step latency under that concurrency
__global__ void chase_many(const unsigned* __restrict__ next, int steps, unsigned n, long long* cycles, unsigned* sink) {
  unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  unsigned i = (w * 2654435761u) % n;
  long long t0 = clock64();
  for (int s = 0; s < steps; ++s) i = next[i];
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) { cycles[w] = t1 - t0; sink[w] = i; }
}

// dependent-issue latency of the ALU / warp-collective ops the select loop chains
__global__ void ops(long long* out, float x, int iters) {
  const unsigned FULL = 0xffffffffu;
  float f = x + threadIdx.x;
  unsigned u = threadIdx.x * 2654435761u + 1u;
  long long t[8];
  t[0] = clock64();
  for (int i = 0; i < iters; ++i) u = __reduce_max_sync(FULL, u ^ (unsigned)i) + threadIdx.x;
  t[1] = clock64();
  for (int i = 0; i < iters; ++i) u = __shfl_sync(FULL, u + i, (u >> 3) & 31);
  t[2] = clock64();
  for (int i = 0; i < iters; ++i) f = __fdiv_rn(f + 1.0f, 1.0001f + f * 0.5f);
  t[3] = clock64();
  for (int i = 0; i < iters; ++i) f = __fsqrt_rn(f + 2.0f);
  t[4] = clock64();
  for (int i = 0; i < iters; ++i) f = __fadd_rn(__fmul_rn(f, 1.0001f), 0.5f);
  t[5] = clock64();
  for (int i = 0; i < iters; ++i) u = __ballot_sync(FULL, (u + i) & 1) + threadIdx.x;
  t[6] = clock64();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 6; ++k) out[k] = t[k + 1] - t[k];
    out[7] = (long long)f + u;
  }
}

// launch overhead: N empty dependent kernels back to back, in a graph
__global__ void empty_kernel(int* p) { if (p && threadIdx.x == 12345) *p = 1; }
struct Big { char pad[960]; };
__global__ void empty_big(Big b, int* p) { if (p && threadIdx.x == 12345) *p = b.pad[5]; }

int main() {
  {
    long long* d; cudaMalloc(&d, 64); long long h[8];
    ops<<<1, 32>>>(d, 1.0f, 1000); ops<<<1, 32>>>(d, 1.0f, 1000);
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("dependent-issue cycles/op: redux %.1f  shfl %.1f  fdiv_rn %.1f  fsqrt_rn %.1f  fmul+fadd %.1f  ballot %.1f\n",
           h[0] / 1000.0, h[1] / 1000.0, h[2] / 1000.0, h[3] / 1000.0, h[4] / 1000.0, h[5] / 1000.0);
    cudaStream_t st; cudaStreamCreate(&st);
    for (int variant = 0; variant < 2; ++variant) {
      cudaGraph_t g; cudaGraphExec_t ge;
      cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
      Big big = {};
      for (int i = 0; i < 1000; ++i) {
        if (variant == 0) empty_kernel<<<512, 64, 0, st>>>(nullptr); else empty_big<<<512, 64, 0, st>>>(big, nullptr);
      }
      cudaStreamEndCapture(st, &g);
      cudaGraphInstantiate(&ge, g, 0);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
      cudaEventRecord(e0, st); cudaGraphLaunch(ge, st); cudaEventRecord(e1, st); cudaStreamSynchronize(st);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("graph of 1000 dependent empty kernels (%s params, 512x64): %.2f us per kernel\n", variant ? "960 B" : "8 B", ms);
    }
  }
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("%s SMs=%d clock=%d MHz L2=%d MB\n", prop.name, prop.multiProcessorCount, clk_khz / 1000, prop.l2CacheSize >> 20);
  long long* d_cycles; unsigned* d_sink;
  cudaMalloc(&d_cycles, 8 * 65536); cudaMalloc(&d_sink, 4 * 65536);
  size_t sizes_mb[] = {1, 8, 32, 64, 96, 128, 256, 1024};
  for (size_t mb : sizes_mb) {
    size_t n = mb * (1 << 20) / 128;  // one element per 128 B line
    std::vector<unsigned> perm(n);
    for (size_t i = 0; i < n; ++i) perm[i] = (unsigned)i;
    unsigned seed = 12345;
    for (size_t i = n - 1; i > 0; --i) { seed = seed * 1664525u + 1013904223u; size_t j = seed % (i + 1); std::swap(perm[i], perm[j]); }
    std::vector<unsigned> next(n * 32, 0);
    for (size_t i = 0; i < n; ++i) next[(size_t)perm[i] * 32] = perm[(i + 1) % n] * 32;  // cycle through all lines
    unsigned* d_next; cudaMalloc(&d_next, n * 128);
    cudaMemcpy(d_next, next.data(), n * 128, cudaMemcpyHostToDevice);
    int steps = 20000;
    long long c1 = 0, c2 = 0;
    chase<<<1, 1>>>(d_next, steps, perm[0] * 32, d_cycles, d_sink);  // warm
    cudaMemcpy(&c1, d_cycles, 8, cudaMemcpyDeviceToHost);
    chase<<<1, 1>>>(d_next, steps, perm[0] * 32, d_cycles, d_sink);
    cudaMemcpy(&c2, d_cycles, 8, cudaMemcpyDeviceToHost);
    // 1024 warps concurrently
    int warps = 1024;
    chase_many<<<warps / 2, 64>>>(d_next, 2000, (unsigned)n, d_cycles, d_sink);
    chase_many<<<warps / 2, 64>>>(d_next, 2000, (unsigned)n, d_cycles, d_sink);
    std::vector<long long> cyc(warps);
    cudaMemcpy(cyc.data(), d_cycles, 8 * warps, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto c : cyc) avg += (double)c; avg /= warps * 2000.0;
    printf("footprint %5zu MB: single chain %7.1f cyc/load (first pass %7.1f)   1024 warps: %7.1f cyc/load\n", mb,
           (double)c2 / steps, (double)c1 / steps, avg);
    cudaFree(d_next);
  }
  return 0;
}
