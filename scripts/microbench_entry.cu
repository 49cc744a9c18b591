// What the FIRST memory round trip of a kernel costs when every warp touches many different arrays (pages),
// as k_sim does at entry (work buffers + tree arrays), vs later trips.  Diagnostic only.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

struct Ptrs { const int* a[16]; };

// each warp: lane 0..31 loads one word from each of `na` arrays at a per-warp offset, then a dependent second trip
__global__ void entry_trips(Ptrs p, int na, size_t stride_words, long long* out, int* sink, int kernel_tag) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  int acc = 0;
  for (int i = 0; i < na; ++i) acc += p.a[i][(size_t)w * stride_words + lane];
  acc = __shfl_sync(0xffffffffu, acc, 0);  // forces completion
  long long t1 = clock64();
  int acc2 = 0;
  for (int i = 0; i < na; ++i) acc2 += p.a[i][(size_t)w * stride_words + 64 + ((acc + lane) & 31)];
  acc2 = __shfl_sync(0xffffffffu, acc2, 0);
  long long t2 = clock64();
  int acc3 = 0;
  for (int i = 0; i < na; ++i) acc3 += p.a[i][(size_t)w * stride_words + 128 + ((acc2 + lane) & 31)];
  acc3 = __shfl_sync(0xffffffffu, acc3, 0);
  long long t3 = clock64();
  if (lane == 0) { out[3 * w] = t1 - t0; out[3 * w + 1] = t2 - t1; out[3 * w + 2] = t3 - t2; sink[w] = acc3 + kernel_tag; }
}

int main() {
  const int W = 1024;                 // warps (trees)
  const size_t stride_words = 4096;   // 16 KB per warp per array -> 16 MB per array
  Ptrs p;
  int* base; cudaMalloc(&base, 16 * W * stride_words * 4 + (64 << 20));
  cudaMemset(base, 0, 16 * W * stride_words * 4);
  long long* d_out; int* d_sink; cudaMalloc(&d_out, 24 * W); cudaMalloc(&d_sink, 4 * W);
  std::vector<long long> h(3 * W);
  for (int na : {1, 4, 8, 16}) {
    for (int mode = 0; mode < 2; ++mode) {  // 0: arrays far apart (own pages); 1: all "arrays" interleaved inside each warp's 16 KB block
      for (int i = 0; i < 16; ++i) p.a[i] = mode == 0 ? base + (size_t)i * W * stride_words : base + i * 256;
      const size_t sw = mode == 0 ? stride_words : stride_words;  // same per-warp stride
      for (int rep = 0; rep < 3; ++rep) entry_trips<<<W / 2, 64>>>(p, na, sw, d_out, d_sink, rep);
      cudaMemcpy(h.data(), d_out, 24 * W, cudaMemcpyDeviceToHost);
      double s[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
      for (int w = 0; w < W; ++w) for (int k = 0; k < 3; ++k) { s[k] += h[3 * w + k]; if (h[3 * w + k] > mx[k]) mx[k] = h[3 * w + k]; }
      printf("na=%2d %-12s: trip1 avg %6.0f max %6.0f | trip2 avg %6.0f max %6.0f | trip3 avg %6.0f max %6.0f cycles\n", na,
             mode == 0 ? "separate" : "interleaved", s[0] / W, mx[0], s[1] / W, mx[1], s[2] / W, mx[2]);
    }
  }
  return 0;
}
