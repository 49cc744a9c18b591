// Front-end costs a short kernel pays on every launch: kernel-parameter (constant bank) misses and instruction-cache
// misses.  Diagnostic only; results are quoted in DESIGN.md.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

struct Big { int v[240]; };  // 960 B of parameters, all zero

// chain of dependent parameter reads, each 64 B (16 ints) further: every read is the first touch of its line
__global__ void param_chain(Big b, long long* out) {
  long long t0 = clock64();
  int idx = 0;
#pragma unroll 1
  for (int i = 0; i < 14; ++i) idx = idx + 16 + b.v[idx];  // value is 0; the compiler cannot know
  long long t1 = clock64();
  int idx2 = idx - 14 * 16;
#pragma unroll 1
  for (int i = 0; i < 14; ++i) idx2 = idx2 + 16 + b.v[idx2];  // same lines again: hits
  long long t2 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = idx + idx2; }
}

// same number of parameter lines, touched independently (loads can overlap)
__global__ void param_parallel(Big b, long long* out) {
  long long t0 = clock64();
  int s = 0;
#pragma unroll
  for (int i = 0; i < 14; ++i) s += b.v[16 * i + (threadIdx.x & 1)];
  if (s == 12345) out[3] = s;
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[2] = s; }
}

// straight-line code of a given size: N dependent FMAs, fully unrolled (16 B of SASS each)
template <int N>
__global__ void code_body(float x, long long* out, float* sink) {
  long long t0 = clock64();
  float a = x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < N; ++i) a = fmaf(a, 1.0001f, 0.5f + i);
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (a == 1234.5f) *sink = a;
}
__global__ void other_kernel(float x, float* sink) {  // something else in between (like the leaf kernel)
  float a = x;
#pragma unroll
  for (int i = 0; i < 512; ++i) a = fmaf(a, 1.0002f, 0.25f + i);
  if (a == 1234.5f) *sink = a;
}

int main() {
  long long* d; cudaMalloc(&d, 64); float* sink; cudaMalloc(&sink, 4);
  long long h[8];
  Big big = {};
  for (int rep = 0; rep < 3; ++rep) {
    param_chain<<<512, 64>>>(big, d);
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("param chain: 14 first-touch lines %lld cyc (%.0f / line), same lines again %lld cyc (%.0f / line)\n", h[0], h[0] / 14.0, h[1], h[1] / 14.0);
  }
  for (int rep = 0; rep < 3; ++rep) {
    param_parallel<<<512, 64>>>(big, d);
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("param parallel: 14 lines touched independently %lld cyc\n", h[0]);
  }
  for (int rep = 0; rep < 4; ++rep) {
    code_body<2048><<<512, 64>>>(1.0f, d, sink);
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("code 2048 FMAs (32 KB), launch %d back to back: %lld cyc (%.2f / instr)\n", rep, h[0], h[0] / 2048.0);
  }
  for (int rep = 0; rep < 4; ++rep) {
    other_kernel<<<512, 64>>>(1.0f, sink);
    code_body<2048><<<512, 64>>>(1.0f, d, sink);
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("code 2048 FMAs (32 KB), after another kernel: %lld cyc (%.2f / instr)\n", h[0], h[0] / 2048.0);
  }
  for (int rep = 0; rep < 3; ++rep) {
    other_kernel<<<512, 64>>>(1.0f, sink);
    code_body<512><<<512, 64>>>(1.0f, d, sink);
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("code 512 FMAs (8 KB), after another kernel: %lld cyc (%.2f / instr)\n", h[0], h[0] / 512.0);
  }
  for (int rep = 0; rep < 3; ++rep) {
    other_kernel<<<512, 64>>>(1.0f, sink);
    code_body<8192><<<512, 64>>>(1.0f, d, sink);
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("code 8192 FMAs (128 KB), after another kernel: %lld cyc (%.2f / instr)\n", h[0], h[0] / 8192.0);
  }
  return 0;
}
